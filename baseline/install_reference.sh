#!/usr/bin/env bash
# Installs the UNMODIFIED reference into baseline/_ref (git-ignored; travels to the GPU box with the gpurun snapshot).
#  1. /root/reference is read-only and setup.py builds in-tree -> install from a /tmp copy
#  2. dependency resolution fails offline (bittensor==6.10.1, mlflow, ... unavailable) -> --no-deps
#  3. the reference's packaging omits hivetrain/utils (no __init__.py, so find_packages() skips it; upstream only works
#     with `pip install -e .`): the directory is copied verbatim next to the installed package.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
SRC="${1:-/root/reference}"
rm -rf /tmp/ref_copy && cp -r "$SRC" /tmp/ref_copy
rm -rf "$HERE/_ref"
python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target "$HERE/_ref" /tmp/ref_copy
cp -r "$SRC/hivetrain/utils" "$HERE/_ref/hivetrain/utils"
diff -rq "$SRC/hivetrain" "$HERE/_ref/hivetrain" | grep -v "Only in $SRC/hivetrain: docs" | grep -v __pycache__ || true
echo "reference installed into $HERE/_ref"
