#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY -- builds the *reference's own* vren CUDA extension as the
# GPU-side oracle ("GPU-ref A" in BASELINE.md).  Nothing under mf-nerf_b200/ may load it.
#
# The reference sources are compiled from where they lie (/root/reference/models/csrc).
# torch 2.11 no longer converts DeprecatedTypeProperties -> ScalarType, so the 13
# `AT_DISPATCH_*(x.type(), ...)` sites are patched to `x.scalar_type()` on a SCRATCH copy
# under /tmp (never in the repo, never in /root/reference).  The python module is renamed
# vren_ref so it can sit beside our own `vren` drop-in.  Output: oracle/_ref/vren_ref*.so
# (git-ignored, still shipped to the GPU box by gpurun).  Compile flags are the
# reference's own (-O2, default -fmad=true) because marcher bit-exactness depends on them.
set -euo pipefail
REF=${REF:-/root/reference}
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/_ref"
if [ ! -d "$REF/models/csrc" ]; then
  echo "[build_ref_vren] $REF not present (GPU box?) - using prebuilt files in $OUT"; exit 0
fi
mkdir -p "$OUT"
# the frozen-API python layer, staged (git-ignored) so GPU-side integration tests can drive the
# UNMODIFIED reference rendering/custom_functions/networks/losses/utils through our shims.
mkdir -p "$OUT/refpy/models"
cp "$REF/models/__init__.py" "$REF/models/custom_functions.py" "$REF/models/networks.py" "$REF/models/rendering.py" "$OUT/refpy/models/"
cp "$REF/losses.py" "$REF/utils.py" "$OUT/refpy/"
if ls "$OUT"/vren_ref*.so >/dev/null 2>&1 && [ -z "${FORCE:-}" ]; then
  echo "[build_ref_vren] already built: $(ls "$OUT"/vren_ref*.so)"; exit 0
fi
SCRATCH=$(mktemp -d /tmp/vren_ref_build.XXXXXX)
cp -r "$REF/models/csrc/." "$SCRATCH/"
cd "$SCRATCH"
sed -i -E 's/([A-Za-z_]+)\.type\(\), "/\1.scalar_type(), "/' *.cu
sed -i -E "s/name='vren'/name='vren_ref'/g" setup.py
TORCH_CUDA_ARCH_LIST="10.0" MAX_JOBS=${MAX_JOBS:-6} python setup.py build_ext --inplace > "$OUT/build.log" 2>&1
cp vren_ref*.so "$OUT/"
rm -rf "$SCRATCH"
echo "[build_ref_vren] built $(ls "$OUT"/vren_ref*.so)"
