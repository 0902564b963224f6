#!/usr/bin/env bash
# Installs the UNMODIFIED reference into baseline/_ref (git-ignored, travels to the GPU box) for the R-GPU
# baseline (tools/bench_ref_gpu.py):
#   * the `src` package (setup.py at the reference root) -- pure Python;
#   * `pointops` (libs/pointops/setup.py): the reference's own Python wrappers + its pybind11 CUDA extension
#     `pointops._C`, compiled by its own setup.py for sm_100 (TORCH_CUDA_ARCH_LIST=10.0).
# /root/reference is read-only, so both are installed from a copy under /tmp.  No index access
# (--no-index --no-build-isolation --no-deps): lightning / hydra / spconv are NOT installed -- the R-GPU tool
# imports the policy modules it needs and stubs the `src.utils` package init (which imports lightning), exactly
# like oracle/gen_golden_act.py does.  Nothing from the reference enters git history.
set -euo pipefail
REF=${REF_ROOT:-/root/reference}
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/_ref"
if [ ! -d "$REF" ]; then echo "[install_ref] $REF not present -- keeping $OUT" >&2; exit 0; fi
if [ -f "$OUT/.installed" ] && [ -z "$(find "$REF/src" "$REF/libs/pointops" -newer "$OUT/.installed" -name '*.py' -o -newer "$OUT/.installed" -name '*.cu' | head -1)" ]; then
  echo "[install_ref] $OUT up to date"; exit 0
fi
TMP=$(mktemp -d /tmp/pcm_ref_XXXX)
TMP2=$(mktemp -d /tmp/pcm_refops_XXXX)
cp -r "$REF/src" "$REF/setup.py" "$TMP/"
cp -r "$REF/libs/pointops/." "$TMP2/"
mkdir -p "$OUT"
PY=${PYTHON:-python}
$PY -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --upgrade --target "$OUT" "$TMP" 2>&1 | tail -2
(cd "$TMP2" && TORCH_CUDA_ARCH_LIST=10.0 MAX_JOBS=8 $PY -m pip install --no-index --no-build-isolation --no-deps \
   --find-links /opt/wheelhouse --upgrade --target "$OUT" . 2>&1 | tail -2)
rm -rf "$TMP" "$TMP2"
touch "$OUT/.installed"
echo "[install_ref] installed: $(ls "$OUT" | tr '\n' ' ')"
