#!/usr/bin/env bash
# Installs the UNMODIFIED reference package into baseline/_ref (git-ignored, travels to the GPU box with gpurun) for
# `bench.py --impl reference`.  The reference builds with hatchling, which this image does not have, so the install runs
# from a copy under /tmp whose pyproject.toml names setuptools as the build backend instead; the package sources
# (vector_quantization/**) are installed byte for byte (checked below).  Dependencies are not resolved (--no-deps):
# torch / einops / rich are in the image; `einx` is not and cannot be installed -- the reference uses exactly one einx
# call (residual_vq.py:117, get_codes_from_indices), for which baseline/einx_standin/einx.py provides a stand-in that
# bench.py puts on sys.path.  Nothing of the reference is committed.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${VQB_REFERENCE:-/root/reference}"
TMP="$(mktemp -d /tmp/vqref.XXXXXX)"
cp -r "$REF"/. "$TMP"/
python - "$TMP/pyproject.toml" <<'PY'
import re, sys
p = sys.argv[1]
s = open(p).read()
s = re.sub(r'\[build-system\]\nrequires = \["hatchling"\]\nbuild-backend = "hatchling.build"',
           '[build-system]\nrequires = ["setuptools"]\nbuild-backend = "setuptools.build_meta"\n\n'
           '[tool.setuptools.packages.find]\ninclude = ["vector_quantization*"]', s)
open(p, "w").write(s)
PY
rm -rf "$HERE/_ref"
python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target "$HERE/_ref" "$TMP" \
  > "$HERE/_ref_install.log" 2>&1 || { tail -20 "$HERE/_ref_install.log"; exit 1; }
diff -r "$REF/vector_quantization" "$HERE/_ref/vector_quantization" -x __pycache__ && echo "baseline/_ref: sources identical to $REF"
rm -rf "$TMP"
