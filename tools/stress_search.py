"""Repeats searches of several shapes through ONE shared workspace and compares every result with the first of its
shape: the answer must not depend on timing (shared running minima, atomics, pair-list order) or on what an earlier
search left in the workspace.  usage: python tools/stress_search.py [reps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "vector-quantization-by-ml_b200"))
import torch
from vqb200 import ops
dev = torch.device("cuda:0")
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
g = torch.Generator(device=dev).manual_seed(11)
cases = []
for (N, K, d, dt, cos) in [(1 << 20, 8192, 256, torch.bfloat16, False), (65536, 8192, 256, torch.float32, False),
                           (1 << 18, 1024, 512, torch.float32, False), (1 << 19, 16384, 512, torch.float32, True),
                           (300001, 1000, 72, torch.float32, False), (1 << 21, 8192, 64, torch.float32, False)]:
    x = torch.randn(1, N, d, generator=g, device=dev)
    c = torch.randn(1, K, d, generator=g, device=dev) * 0.5
    if cos:
        x = ops.l2norm_rows(x); c = torch.nn.functional.normalize(c, dim=-1)
    if len(cases) == 1:      # exact codebook rows as latents (the idempotence leg of the full-size test)
        x = c[:, torch.randint(0, K, (N,), generator=g, device=dev)].contiguous()
    x = x.to(dt).contiguous()
    cache = ops.prepare_codebook(c, cos)
    ref, rs, ws = ops.search(x, c, cache, cos, want_score=True)
    cases.append((x, c, cache, cos, ref.clone(), rs.clone(), ops.search_stats(ws)))
bad = 0
for r in range(reps):
    for ci, (x, c, cache, cos, ref, rs, st0) in enumerate(cases):
        want = ((r + ci) % 2 == 0)
        idx, sc, ws = ops.search(x, c, cache, cos, want_score=want)
        diff = (idx != ref).nonzero()
        if diff.numel() or (want and not torch.equal(sc, rs)):
            bad += 1
            rows = diff[:, 1][:6].tolist()
            print(f"rep {r} case {ci}: {diff.shape[0]} rows differ, e.g. {rows}; stats {ops.search_stats(ws)} first {st0}", flush=True)
            for row in rows[:3]:
                ex, es, _ = ops.search(x[:, row:row + 1].contiguous(), c, None, cos, force_exact=True, want_score=True)
                print(f"   row {row}: first {int(ref[0, row])} now {int(idx[0, row])} exact {int(ex[0, 0])} score {float(es[0, 0])}")
print(f"stress_search: {bad} mismatching searches in {reps} x {len(cases)}; first-run stats {[c[6] for c in cases]}")
