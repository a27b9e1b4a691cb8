#!/usr/bin/env bash
# Installs the UNMODIFIED reference package (laplace-torch, /root/reference) into baseline/_ref (git-ignored, travels to
# the GPU box) so that tests / bench.py can drive the real `laplace.Laplace(...)` front end with backend=B200GGN there.
#
# The reference's build backend (pdm-backend) is not in the offline wheelhouse, so the copy under /tmp gets its
# [build-system] table pointed at setuptools -- packaging metadata only, no library source is touched -- and is installed
# with --no-deps (its third-party curvature libraries are not installable offline; laplace_b200/compat.py registers
# placeholders for them at import time).
set -euo pipefail
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
SRC="${1:-/root/reference}"
[ -d "$SRC/laplace" ] || { echo "no reference tree at $SRC"; exit 0; }
TMP="$(mktemp -d /tmp/lpb_ref.XXXXXX)"
cp -r "$SRC/laplace" "$SRC/pyproject.toml" "$SRC/README.md" "$TMP/" 2>/dev/null || cp -r "$SRC/laplace" "$SRC/pyproject.toml" "$TMP/"
python - "$TMP/pyproject.toml" <<'PY'
import sys
p = sys.argv[1]
s = open(p).read()
s = s.replace('requires = ["pdm-backend"]', 'requires = ["setuptools"]').replace('build-backend = "pdm.backend"', 'build-backend = "setuptools.build_meta"')
s = s.replace('py-modules = ["laplace"]', 'packages = ["laplace", "laplace.curvature", "laplace.utils"]')
open(p, "w").write(s)
PY
rm -rf "$ROOT/baseline/_ref"
mkdir -p "$ROOT/baseline"
python -m pip install --quiet --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target "$ROOT/baseline/_ref" "$TMP"
rm -rf "$TMP"
ls "$ROOT/baseline/_ref"
