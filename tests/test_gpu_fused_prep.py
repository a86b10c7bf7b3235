"""In-kernel conversion of 16-bit latents (search_tc.cu, CONV; opt-in flag VQB_SEARCH_FUSED_PREP): the search kernel
converts the rows itself behind a sampled bound of the row norms.  Indices and exact scores must equal the exact scan bit for bit -- also for rows that the
sample never saw and that exceed its bound (they must come back through the exact rescan), for all-zero rows, for ragged
row counts and for several codebooks -- and must equal what the separate prepare pass gives."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _case(dtype, H, N, K, d, cos, outliers, seed):
    n_bad = 0
    from vqb200 import ops
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(seed)
    x = torch.randn(H, N, d, generator=g, device=dev)
    if outliers:
        # the strided sample takes rows 0, stride, 2 stride, ...: put rows far above its bound in between, a block of
        # zero rows (padding in real batches), and a few tiny rows
        stride = max((H * N) // 16384, 1)
        flat = x.view(-1, d)
        bad = torch.arange(1, H * N, 997, device=dev)
        bad = bad[(bad % stride) != 0] if stride > 1 else bad[:0]
        flat[bad[:50]] *= 40.0
        n_bad = int(bad[:50].numel())
        flat[5000:5100] = 0.0
        flat[7001:7011] *= 1e-6
    x = x.to(dtype)
    c = torch.randn(H, K, d, generator=g, device=dev) * 0.5
    if cos:
        c = torch.nn.functional.normalize(c, dim=-1)
    cache = ops.prepare_codebook(c, cos)
    # leave another batch's row scales / operands behind in the shared search workspace: nothing of them may be read
    ops.search((torch.randn(H, N, d, generator=g, device=dev) * 300.0).to(dtype), c, cache, cos, fused_prep=True)
    idx, score, ws = ops.search(x, c, cache, cos, want_score=True, fused_prep=True)
    st = ops.search_stats(ws)
    ex, es, _ = ops.search(x, c, None, cos, force_exact=True, want_score=True)
    return idx, score, ex, es, st, n_bad


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("H,N,K,d,cos,outliers", [
    (1, 40000, 8192, 256, False, False),      # ragged last row tile
    (1, 40000, 8192, 256, False, True),
    (2, 16500, 6144, 256, False, True),       # two codebooks, padded codes? (6144 = 24 x 256)
    (1, 33000, 8192, 248, False, True),       # d not a multiple of 64 (zero padded to 256)
    (1, 40000, 8192, 256, True, True),        # dot metric, no bias k-step (aug mode 2)
    (1, 40000, 24576, 64, False, True),       # one stage per tile
])
def test_fused_prep_equals_exact_scan(dtype, H, N, K, d, cos, outliers):
    idx, score, ex, es, st, n_bad = _case(dtype, H, N, K, d, cos, outliers, 5)
    assert st["tensor_core_pass"] == 1
    assert torch.equal(idx, ex), f"{int((idx != ex).sum())} rows differ"
    assert torch.equal(score, es)
    # the rows above the sampled bound went through the exact rescan
    assert st["rescanned_rows"] >= n_bad, (st, n_bad)


def test_fused_prep_equals_separate_prepare_pass():
    """Same search with the in-kernel conversion and with the separate prepare pass: identical indices and scores, and
    the launch lists differ exactly as they should (sample kernel instead of the prepare kernel)."""
    from torch.profiler import ProfilerActivity, profile
    from vqb200 import ops
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(9)
    x = torch.randn(1, 65536, 256, generator=g, device=dev).bfloat16()
    c = torch.randn(1, 8192, 256, generator=g, device=dev) * 0.5
    cache = ops.prepare_codebook(c, False)
    res = {}
    for fused in (True, False):
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            idx, score, ws = ops.search(x, c, cache, False, want_score=True, fused_prep=fused)
            torch.cuda.synchronize()
        names = " ".join(e.key for e in prof.key_averages())
        res[fused] = (idx.clone(), score.clone(), names)
    assert torch.equal(res[True][0], res[False][0]) and torch.equal(res[True][1], res[False][1])
    if os.environ.get("VQB_FUSED_PREP") is None:
        assert "sample_bound_kernel" in res[True][2] and "prepare_latents" not in res[True][2], res[True][2]
        assert "prepare_latents" in res[False][2] and "sample_bound_kernel" not in res[False][2], res[False][2]
