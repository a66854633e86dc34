#!/bin/sh
# Vendor the UNMODIFIED reference module files this path's baselines import into oracle/_ref/ (git-ignored, but shipped to
# the GPU box with the gpurun snapshot, like the built .so files).  Test / baseline infrastructure only: nothing in the
# product package imports oracle/.  Run from the repo root in the container that has /root/reference:
#     sh oracle/vendor_ref.sh
# Sources (copied byte for byte, never edited):
#   exploration/GGTV_GGLR_v1.0/deep_multiscale_GGLR_GGTV_v1x0.py            (V1X0 == LIB v13 == v22: the hot path, bench + GPU baseline)
#   exploration/model_multiscale_mixture_GLR/lib/model_GLR_GTV_deep_v7.py   (older family, config 3)
#   exploration/model_multiscale_mixture_GLR/lib/model_GLR_GTV_deep_v1.py   (three-block chain of config 3)
set -e
REF=${REF:-/root/reference}
DST="$(dirname "$0")/_ref"
if [ ! -d "$REF" ]; then
    echo "vendor_ref: $REF not present (GPU box): keeping what is in $DST" >&2
    exit 0
fi
mkdir -p "$DST"
cp "$REF/exploration/GGTV_GGLR_v1.0/deep_multiscale_GGLR_GGTV_v1x0.py" "$DST/"
cp "$REF/exploration/model_multiscale_mixture_GLR/lib/model_GLR_GTV_deep_v7.py" "$DST/"
cp "$REF/exploration/model_multiscale_mixture_GLR/lib/model_GLR_GTV_deep_v1.py" "$DST/"
( cd "$DST" && md5sum *.py > MD5SUMS )
echo "vendor_ref: $(ls "$DST" | tr '\n' ' ')"
