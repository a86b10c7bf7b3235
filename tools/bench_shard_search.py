"""One rank's search of the sharded config 5 on ONE GPU (4M fp32 latents x d=64 against an 8192-code shard, exact
score of every row wanted): time of the whole `ops.search(..., want_score=True)` and of its kernels (torch profiler)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "vector-quantization-by-ml_b200"))
import torch
from torch.profiler import ProfilerActivity, profile
from vqb200 import ops
dev = torch.device("cuda:0")
N, K, d = (int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (1 << 22, 8192, 64)
g = torch.Generator(device=dev).manual_seed(0)
x = torch.randn(1, N, d, generator=g, device=dev)
c = torch.randn(1, K, d, generator=g, device=dev) * 0.5
cache = ops.prepare_codebook(c, False)
for _ in range(3):
    ops.search(x, c, cache, False, want_score=True, idx_offset=K)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    ops.search(x, c, cache, False, want_score=True, idx_offset=K)
e1.record(); torch.cuda.synchronize()
print(f"search(want_score) N={N} K={K} d={d}: {e0.elapsed_time(e1) / 5:.3f} ms", flush=True)
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        ops.search(x, c, cache, False, want_score=True, idx_offset=K)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=10, max_name_column_width=60))
