#!/bin/bash
# Installs the UNMODIFIED reference (nanograv/PTMCMCSampler at /root/reference) into the git-ignored
# baseline/_ref/ so that `bench.py --impl reference` can time it on the benchmark host (the directory travels
# to the GPU box with the snapshot, like the built .so files).  Offline: no index, no dependency resolution
# (numpy / scipy are in the image).  setuptools_scm cannot write PTMCMCSampler/version.py without git
# metadata, so the one-line file the package's own build would generate is written here.
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
SRC=${1:-/root/reference}
[ -d "$SRC" ] || { echo "no reference at $SRC"; exit 0; }
TMP=$(mktemp -d)
cp -r "$SRC" "$TMP/ref"
rm -rf "$ROOT/baseline/_ref"
python -m pip install -q --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
    --target "$ROOT/baseline/_ref" "$TMP/ref" || { mkdir -p "$ROOT/baseline/_ref"; cp -r "$SRC/PTMCMCSampler" "$ROOT/baseline/_ref/"; }
[ -f "$ROOT/baseline/_ref/PTMCMCSampler/version.py" ] || echo 'version = "0+ref.dd837f9"' > "$ROOT/baseline/_ref/PTMCMCSampler/version.py"
rm -rf "$TMP"
echo "reference installed in $ROOT/baseline/_ref"
