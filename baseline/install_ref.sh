#!/bin/bash
# Install the UNMODIFIED reference (cvg/GeoCalib, /root/reference) into the git-ignored baseline/_ref
# so that it travels to the GPU box with gpurun (bench.py --impl reference, tests/test_dropin_gpu.py).
# The source tree is read-only, so pip builds from a copy under /tmp; --no-deps because the wheelhouse
# has no torch wheel (torch is already in the image; kornia / matplotlib are absent and only needed by
# ImagePreprocessor.resize / viz, see baseline/ref_import.py).
set -e
cd "$(dirname "$0")/.."
[ -d /root/reference ] || { echo "no /root/reference here: keeping the existing baseline/_ref"; exit 0; }
rm -rf /tmp/gclm_refcopy baseline/_ref
cp -r /root/reference /tmp/gclm_refcopy
python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
  --target baseline/_ref /tmp/gclm_refcopy
rm -rf /tmp/gclm_refcopy
