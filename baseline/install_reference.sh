#!/bin/bash
# Install the UNMODIFIED reference (brandondube/prysm v0.22, pure Python) into baseline/_ref/ (git-ignored, travels to the
# GPU box with the gpurun snapshot).  The prescribed command
#   python -m pip install --no-index --no-build-isolation --find-links /opt/wheelhouse --target baseline/_ref /root/reference
# fails in this image because the reference's build backend (hatchling) is not installed and there is no network.
# The package is pure Python and its wheel target is `packages = ["prysm"]`, so the same files are installed by
# building from a scratch copy under /tmp whose [build-system] stanza names setuptools instead; not one line of
# the prysm/ package is touched (checked below with diff -r).
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
SRC="${1:-/root/reference}"
DST="$HERE/_ref"
if python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target "$DST" "$SRC" >/tmp/ref_install.log 2>&1; then
  echo "installed with the reference's own build backend"; exit 0
fi
TMP="$(mktemp -d /tmp/prysm_src.XXXXXX)"
cp -r "$SRC/prysm" "$TMP/prysm"
cp "$SRC/LICENSE.md" "$SRC/README.md" "$TMP/" 2>/dev/null || true
cat > "$TMP/pyproject.toml" <<'TOML'
[build-system]
requires = ["setuptools"]
build-backend = "setuptools.build_meta"
[project]
name = "prysm"
version = "0.22"
dependencies = []
[tool.setuptools.packages.find]
include = ["prysm*"]
[tool.setuptools.package-data]
"*" = ["*"]
TOML
rm -rf "$DST"
python -m pip install --no-index --no-build-isolation --no-deps --target "$DST" "$TMP" 2>&1 | tail -2
diff -r -q "$SRC/prysm" "$DST/prysm" -x __pycache__ && echo "baseline/_ref/prysm is identical to $SRC/prysm"
rm -rf "$TMP"
