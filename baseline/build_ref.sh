#!/bin/bash
# Install the UNMODIFIED reference (ysig/GraKeL at /root/reference) into baseline/_ref for the CPU arm of
# bench.py (`--impl reference`, `cpu_baseline`).  /root/reference is read-only and the build writes into the
# source tree, so a copy under /tmp is installed.  Build container only: the GPU box receives the installed
# tree with the gpurun snapshot (baseline/_ref is git-ignored, not gpurun-ignored).
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
SRC=${1:-/root/reference}
TMP=$(mktemp -d /tmp/grakel_ref.XXXXXX)
cp -r "$SRC" "$TMP/ref"
chmod -R u+w "$TMP/ref"
rm -rf "$HERE/_ref"
python -m pip install --no-index --no-build-isolation --find-links /opt/wheelhouse --no-deps --target "$HERE/_ref" "$TMP/ref"
rm -rf "$TMP"
python -c "import sys; sys.path.insert(0, '$HERE/_ref'); import grakel; print('installed grakel', grakel.__version__, 'into', grakel.__file__)"
