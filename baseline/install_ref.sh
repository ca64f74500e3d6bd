#!/bin/bash
# Installs the UNMODIFIED reference package (torch_geometric 2.9.0) into baseline/_ref (git-ignored, travels with
# gpurun).  /root/reference is read-only and its build backend (flit_core) is not in this image, so the install runs
# from a copy under /tmp whose pyproject.toml [build-system] table alone is pointed at setuptools (present here); no
# file of the torch_geometric/ package is touched.  --no-deps: pyparsing is absent (never imported by the package).
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
SRC=${1:-/root/reference}
[ -d "$SRC/torch_geometric" ] || { echo "no reference at $SRC"; exit 0; }
TMP=$(mktemp -d /tmp/refcopy.XXXXXX)
cp -r "$SRC/torch_geometric" "$SRC/pyproject.toml" "$SRC/README.md" "$SRC/LICENSE" "$TMP/"
python - "$TMP/pyproject.toml" <<'PY'
import re, sys
p = sys.argv[1]
t = open(p).read()
t = re.sub(r'\[build-system\].*?(?=\n\[)', '[build-system]\nrequires=["setuptools"]\nbuild-backend="setuptools.build_meta"\n', t, count=1, flags=re.S)
t = re.sub(r'\n\[tool\.flit[^\]]*\].*?(?=\n\[|\Z)', '\n', t, flags=re.S)
t += '\n[tool.setuptools.packages.find]\ninclude=["torch_geometric*"]\n[tool.setuptools.package-data]\n"*"=["*.jinja", "*.json", "*.yaml", "*.yml", "*.txt"]\n'
open(p, "w").write(t)
PY
rm -rf "$ROOT/baseline/_ref"
python -m pip install --no-index --no-build-isolation --find-links /opt/wheelhouse --no-deps --target "$ROOT/baseline/_ref" "$TMP" 
rm -rf "$TMP"
# every file of the installed package must equal the reference's
(cd "$SRC" && find torch_geometric -type f \( -name '*.py' -o -name '*.jinja' \) | sort | while read f; do cmp -s "$f" "$ROOT/baseline/_ref/$f" || echo "DIFF $f"; done) | tee /tmp/ref_diff.txt
[ -s /tmp/ref_diff.txt ] && { echo "installed package differs from the reference"; exit 1; }
echo "baseline/_ref installed: $(find "$ROOT/baseline/_ref/torch_geometric" -name '*.py' | wc -l) python files identical to $SRC"
