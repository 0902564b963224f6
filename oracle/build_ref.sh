#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY -- builds the UNMODIFIED reference pointops CUDA launchers
# (libs/pointops/src/**/**_cuda_kernel.cu, compiled where they lie under /root/reference)
# into oracle/_ref/libpointops_ref.so for sm_100a.  Nothing from the reference is copied
# into this repository; oracle/_ref/ is git-ignored but travels to the GPU box.
# The launchers are `extern "C"` raw-pointer functions (e.g. sampling_cuda_kernel.h:13), so
# no torch runtime is linked; torch headers are only needed because the reference headers
# include them for the at::Tensor shims (which we do not compile).
set -euo pipefail
REF=${REF_ROOT:-/root/reference}/libs/pointops/src
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/_ref"
if [ ! -d "$REF" ]; then
  echo "[build_ref] $REF not present (GPU box?) -- keeping prebuilt $OUT" >&2
  exit 0
fi
mkdir -p "$OUT"
PY=${PYTHON:-python}
mkdir -p "$OUT"
TORCH_INC=$($PY - <<'PY'
import torch.utils.cpp_extension as C, sysconfig
print(" ".join("-I"+p for p in C.include_paths() + [sysconfig.get_paths()["include"]]))
PY
)
SRCS=$(ls "$REF"/*/*_cuda_kernel.cu)
# incremental: skip when the .so is newer than every reference source (the torch headers the
# reference's own headers drag in make each file take ~1.5 min to compile).
if [ -f "$OUT/libpointops_ref.so" ] && [ -z "$(find $SRCS "$REF/cuda_utils.h" -newer "$OUT/libpointops_ref.so")" ]; then
  echo "[build_ref] $OUT/libpointops_ref.so up to date"; exit 0
fi
OBJS=""
for s in $SRCS; do
  o="$OUT/$(basename "$s" .cu).o"; OBJS="$OBJS $o"
  nvcc -O2 -c -Xcompiler -fPIC -std=c++17 -gencode arch=compute_100a,code=sm_100a \
    $TORCH_INC -o "$o" "$s" &
done
wait
nvcc -shared -o "$OUT/libpointops_ref.so" $OBJS
rm -f $OBJS
echo "[build_ref] built $OUT/libpointops_ref.so"
