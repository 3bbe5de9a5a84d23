#!/bin/bash
# Install the UNMODIFIED reference (mcgillmrl/prob_mbrl, read-only at /root/reference) into the git-ignored
# baseline/_ref so that it travels to the GPU box with the repo snapshot (it is NOT gpurun-ignored):
#   * `bench.py --impl reference` times the reference's own algorithms.mc_pilco from there;
#   * tests/test_gpu_acceptance.py and baseline/run_example.py run the reference's own modules and
#     examples/deep_pilco_*.py (copied verbatim next to the package: pip installs only the package) on the
#     fused backend.
# Nothing under baseline/_ref is tracked by git and nothing in prob_mbrl_b200/ reads it.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${1:-/root/reference}"
[ -d "$REF/prob_mbrl" ] || { echo "no reference tree at $REF"; exit 1; }
TMP="$(mktemp -d)"
cp -r "$REF" "$TMP/src"                      # the mount is read-only: pip builds its wheel in a copy
rm -rf "$HERE/_ref"
python -m pip install -q --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
    --target "$HERE/_ref" "$TMP/src"
mkdir -p "$HERE/_ref/examples"
cp "$REF"/examples/deep_pilco_no_mm.py "$REF"/examples/deep_pilco_mm.py "$HERE/_ref/examples/"
rm -rf "$TMP"
echo "reference installed at $HERE/_ref"
