"""Randomised differential check of the search (tools/fuzz_search.py): tensor-core path == exact scan, bit for bit,
over random shapes / dtypes / metrics / scales, duplicated codes, zero and outlier rows, with and without the
in-kernel latent conversion."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("seed", [11, 12])
def test_fuzz_search_equals_exact_scan(seed):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "fuzz_search.py"), "150", str(seed)],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "0 mismatches" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]
