#!/bin/bash
# build an experimental variant of the library with extra -D flags: tools/build_variant.sh out.so -DGLR_TH=16 ...
set -e
out=$1; shift
cd "$(dirname "$0")/.."
CUT=$(python -c "from imagerestoration_development_unrolling_b200.build import cutlass_include as c; print(c())")
objs=""
for f in imagerestoration_development_unrolling_b200/csrc/*.cu; do
  o=/tmp/variant_$(basename $f .cu)_$$.o
  /usr/local/cuda/bin/nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-O3 --expt-relaxed-constexpr -I$CUT "$@" -c $f -o $o &
  objs="$objs $o"
done
wait
/usr/local/cuda/bin/nvcc -shared -o $out $objs -gencode arch=compute_100a,code=sm_100a
rm -f $objs
