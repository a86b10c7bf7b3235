"""Fills the search workspace with garbage before a search and compares with the result of a clean run: nothing in the
workspace may be read before it is written."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "vector-quantization-by-ml_b200"))
import torch
from vqb200 import ops, _lib as L
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(11)
for (N, K, d, dt) in [(1 << 20, 8192, 256, torch.bfloat16), (1 << 18, 1024, 512, torch.float32), (300001, 1000, 72, torch.float32)]:
    x = torch.randn(1, N, d, generator=g, device=dev).to(dt).contiguous()
    c = torch.randn(1, K, d, generator=g, device=dev) * 0.5
    cache = ops.prepare_codebook(c, False)
    ref, _, ws = ops.search(x, c, cache, False)
    ref = ref.clone()
    for fill in ("0xff", "random", "0x7f", "zeros"):
        if fill == "random":
            ws.random_(0, 256)
        elif fill == "zeros":
            ws.zero_()
        else:
            ws.fill_(int(fill, 16))
        idx, _, ws2 = ops.search(x, c, cache, False)
        assert ws2.data_ptr() == ws.data_ptr()
        nd = int((idx != ref).sum())
        print(f"N={N} K={K} d={d} fill={fill}: {nd} rows differ; stats {ops.search_stats(ws2)}", flush=True)
        if nd:
            rows = (idx != ref).nonzero()[:, 1][:5].tolist()
            for row in rows:
                ex, es, _ = ops.search(x[:, row:row + 1].contiguous(), c, None, False, force_exact=True, want_score=True)
                print(f"   row {row}: clean {int(ref[0, row])} poisoned {int(idx[0, row])} exact {int(ex[0, 0])}")
