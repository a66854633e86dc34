#!/bin/bash
# Build an experimental variant of the library: the listed sources are recompiled with the extra -D flags, every other object is
# taken from the regular build (csrc/build/*.o; run `python -m imagerestoration_development_unrolling_b200.build` first).
#   tools/build_variant.sh variants/ww5.so weights_walk.cu -DWW_FWD_MINB=5
# Select it with GLRGTV_LIB=variants/ww5.so (tools/*.py, bench.py).
set -e
out=$1; shift
srcs=$1; shift
cd "$(dirname "$0")/.."
C=imagerestoration_development_unrolling_b200/csrc
objs=""
for o in $C/build/*.o; do
  b=$(basename $o .o)
  if echo ",$srcs," | grep -q ",$b.cu,"; then
    v=/tmp/variant_${b}_$$.o
    /usr/local/cuda/bin/nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-O3 --expt-relaxed-constexpr "$@" -c $C/$b.cu -o $v
    objs="$objs $v"
  else
    objs="$objs $o"
  fi
done
mkdir -p "$(dirname $out)"
/usr/local/cuda/bin/nvcc -shared -o $out $objs -gencode arch=compute_100a,code=sm_100a
