"""Bring-up check for the tcgen05 search kernel: TC path vs the exact CUDA-core scan.
Run under `timeout`; prints one line per shape."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "vector-quantization-by-ml_b200"))
import torch

from vqb200 import ops

dev = torch.device("cuda:0")
shapes = [  # H, N, K, d, cosine, cb_scale
    (1, 128, 256, 64, False, 0.5),
    (1, 1024, 512, 256, False, 0.5),
    (1, 1000, 300, 100, False, 0.5),
    (2, 777, 520, 96, True, 0.5),
    (1, 4096, 1024, 512, False, 0.5),
    (1, 1024, 512, 256, False, None),   # default-init scale codebook (near ties)
    (1, 40000, 8192, 256, False, 0.5),
]
BF16 = False
if len(sys.argv) > 1 and sys.argv[1] == "big":
    shapes = [(1, 1 << 20, 8192, 256, False, 0.5)]
    BF16 = True
for (H, N, K, d, cos, scale) in shapes:
    g = torch.Generator().manual_seed(N + K)
    x = torch.randn(H, N, d, generator=g)
    if BF16:
        x = x.bfloat16()
    if scale is None:
        c = (torch.rand(H, K, d, generator=g) * 2 - 1) * (6.0 / (K * d)) ** 0.5
    else:
        c = torch.randn(H, K, d, generator=g) * scale
    xd, cd = x.to(dev), c.to(dev)
    cache = ops.prepare_codebook(cd, cos)
    torch.cuda.synchronize()
    t0 = time.time()
    idx_tc, _, ws = ops.search(xd, cd, cache, cos)
    torch.cuda.synchronize()
    t1 = time.time()
    st = ops.search_stats(ws)
    idx_ex, sc_ex, _ = ops.search(xd, cd, cache, cos, force_exact=True, want_score=True)
    torch.cuda.synchronize()
    t2 = time.time()
    neq = int((idx_tc != idx_ex).sum())
    msg = f"H={H} N={N} K={K} d={d} cos={cos} scale={scale}: tc_vs_exact mismatches={neq} stats={st} t_tc={t1-t0:.4f}s t_exact={t2-t1:.3f}s"
    if N * K <= 1 << 24:
        sim = (torch.einsum('hnd,hkd->hnk', x.float(), c) if cos else -torch.cdist(x.float(), c))
        ref = sim.argmax(-1)
        top2 = sim.topk(2, -1).values
        gap = (top2[..., 0] - top2[..., 1]).abs() / top2[..., 0].abs().clamp_min(1e-30)
        bad = (idx_tc.cpu() != ref) & (gap >= 1e-6)
        msg += f" | vs cpu: mismatches={int((idx_tc.cpu() != ref).sum())} non-exempt={int(bad.sum())}"
    print(msg, flush=True)
    if neq:
        rows = (idx_tc != idx_ex).nonzero()[:5]
        for h, r in rows.tolist():
            print("   row", h, r, "tc", int(idx_tc[h, r]), "exact", int(idx_ex[h, r]), "score", float(sc_ex[h, r]), flush=True)
# per-kernel breakdown for the last shape
from torch.profiler import ProfilerActivity, profile
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        ops.search(xd, cd, cache, cos)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=12, max_name_column_width=70), flush=True)
# timing loop for the last shape
torch.cuda.synchronize()
for it in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ops.search(xd, cd, cache, cos)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"search {N}x{K}x{d}: {ms:.3f} ms -> {2*N*K*d/ms/1e9:.1f} TFLOP/s", flush=True)
