#!/usr/bin/env bash
# Install the UNMODIFIED reference (arnab39/equiadapt at /root/reference) into baseline/_ref (git-ignored, travels to the
# GPU box with the gpurun snapshot).  The source tree is read-only and setuptools writes build files next to setup.py,
# so the install runs from a copy under /tmp; kornia / e2cnn / torch_scatter / omegaconf are absent from the offline
# wheelhouse, hence --no-deps (bench.py's reference arm puts oracle/shims in front of it, see oracle/shims/README.md).
set -euo pipefail
ROOT="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
SRC="${1:-/root/reference}"
TMP="$(mktemp -d)"
cp -r "$SRC" "$TMP/reference"
rm -rf "$ROOT/baseline/_ref"
python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
    --target "$ROOT/baseline/_ref" "$TMP/reference"
rm -rf "$TMP"
PYTHONPATH="$ROOT/oracle/shims:$ROOT/baseline/_ref" python -c "import equiadapt, sys; print('baseline/_ref ok:', equiadapt.__file__)"
