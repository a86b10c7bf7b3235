"""Cold-start check: the very first large search of a fresh process (and a few more on fresh data) against the exact
CUDA scan on a large row sample.  No CPU oracle involved: whatever differs here is the tensor-core path's doing."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "vector-quantization-by-ml_b200"))
import torch
from vqb200 import CodebookParams, VectorQuantize, ops
dev = torch.device("cuda:0")
# the kind of small work the test-suite does before its first full-size search
vq = VectorQuantize(dim=64, codebook_params=CodebookParams(dim=64, codebook_size=128)).to(dev).train()
vq(torch.randn(4, 500, 64, device=dev))
N, K, d = 1 << 20, 8192, 256
total_bad = 0
for seed in (11, 12, 13, 14):
    g = torch.Generator(device=dev).manual_seed(seed)
    x = torch.randn(1, N, d, generator=g, device=dev).bfloat16().contiguous()
    c = torch.randn(1, K, d, generator=g, device=dev) * 0.5
    cache = ops.prepare_codebook(c, False)
    idx, _, ws = ops.search(x, c, cache, False)
    st = ops.search_stats(ws)
    # seed 11 (the full-size test's inputs): EVERY row against the exact scan; the others: a 65536-row sample
    n_chk = N if seed == 11 else 65536
    rows = torch.randperm(N, generator=torch.Generator().manual_seed(seed))[:n_chk].to(dev)
    ex, es, _ = ops.search(x[:, rows].contiguous(), c, None, False, force_exact=True, want_score=True)
    bad = (idx[0, rows] != ex[0]).nonzero().flatten()
    total_bad += int(bad.numel())
    print(f"seed {seed}: {int(bad.numel())} of {n_chk} rows differ from the exact scan; stats {st}", flush=True)
    for b in bad[:6].tolist():
        r = int(rows[b])
        _, s_got, _ = ops.search(x[:, r:r + 1].contiguous(), c[:, int(idx[0, r]):int(idx[0, r]) + 1].contiguous(), None, False,
                                 force_exact=True, want_score=True)
        print(f"   row {r}: got {int(idx[0, r])} (exact score {float(s_got[0, 0]):.7f}) exact {int(ex[0, b])} (score {float(es[0, b]):.7f})")
print("first_search_check:", "OK" if total_bad == 0 else f"{total_bad} MISMATCHES")
