#!/bin/bash
# Installs the UNMODIFIED reference package into baseline/_ref (git-ignored; it travels to the GPU box with
# the gpurun snapshot) so that the -m gpu tests can run the reference's own krr_approx.py on the CUDA
# engine (tests/test_gpu_fit.py::test_reference_krr_approx_on_cuda_engine).  TEST INFRASTRUCTURE ONLY.
#
# The reference's pyproject.toml asks for the `hatchling` build backend, which is not in this image, so
# the install goes through its setup.py from a scratch copy (/root/reference is read-only); no source
# file is altered.  Outcome recorded in DESIGN.md section 8.
set -euo pipefail
ROOT="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
REF="${1:-/root/reference}"
[ -d "${REF}/sobolev_alignment" ] || { echo "no reference tree at ${REF}: nothing to install"; exit 0; }
TMP="$(mktemp -d)"
trap 'rm -rf "${TMP}"' EXIT
cp -r "${REF}" "${TMP}/ref"
rm -f "${TMP}/ref/pyproject.toml"
rm -rf "${ROOT}/baseline/_ref"
python -m pip install --quiet --no-index --no-build-isolation --no-deps --target "${ROOT}/baseline/_ref" "${TMP}/ref"
diff -rq -x __pycache__ "${REF}/sobolev_alignment" "${ROOT}/baseline/_ref/sobolev_alignment"
echo "installed unmodified reference into ${ROOT}/baseline/_ref"
