#!/usr/bin/env bash
# Installs the UNMODIFIED reference (makgyver/gossipy) into baseline/_ref for `bench.py --impl reference`.
# The reference ships no setup.py / pyproject.toml, so pip has nothing to build from /root/reference
# directly; we copy the tree to /tmp, add a minimal setup.py NEXT TO the untouched package sources and
# install with the offline flags the task prescribes (--no-deps: its pins are years old).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
SRC="${1:-/root/reference}"
TMP="$(mktemp -d /tmp/gossipy_ref.XXXXXX)"
cp -r "$SRC/gossipy" "$TMP/gossipy"
cat > "$TMP/setup.py" <<'PY'
from setuptools import setup, find_packages
setup(name="gossipy-dfl", version="0.0.1", packages=find_packages(include=["gossipy", "gossipy.*"]))
PY
rm -rf "$HERE/_ref"
python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
    --target "$HERE/_ref" "$TMP" 2>&1 | tail -3
diff -r "$SRC/gossipy" "$HERE/_ref/gossipy" --exclude=__pycache__ && echo "reference installed unmodified into $HERE/_ref"
rm -rf "$TMP"
