"""Config 5: large-codebook VQ, K=65536 x d=64, 4M latents, codebook rows sharded over the ranks with a cross-GPU
(score, index) min-reduction.  Launch: torchrun --nproc-per-node W --master-addr 127.0.0.1 tools/bench_sharded.py
Every rank holds all latents (replicated) and K/W codes; prints one JSON line from rank 0."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "vector-quantization-by-ml_b200"))
import torch
import torch.distributed as dist

from vqb200 import ShardedCodebook, ops

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
N, K, d = 1 << 22, 65536, 64
g = torch.Generator(device=dev).manual_seed(7)            # same latents and codebook on every rank
x = torch.randn(N, d, generator=g, device=dev)
full = torch.randn(K, d, generator=g, device=dev) * 0.5
sh = ShardedCodebook(d, K).to(dev)
sh.load_full_codebook(full)
sh.train()


def step():
    return sh(x)


for _ in range(3):
    step()
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
steps = 10
e0.record()
for _ in range(steps):
    q, idx, commit = step()
e1.record()
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
t = torch.tensor([e0.elapsed_time(e1) / steps], device=dev, dtype=torch.float64)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
# parity of the sharded argmin against an un-sharded exact scan on a sample of rows
sample = torch.arange(0, N, N // 4096, device=dev)
cur = sh.gather_full_codebook()[None].contiguous()
q2, idx2, _ = None, None, None
sh.eval()
_, idx_eval, _ = sh(x)
ref, _, _ = ops.search(x[sample][None].contiguous(), cur, None, False, force_exact=True)
ok = bool(torch.equal(idx_eval[sample], ref[0]))
if rank == 0:
    ms = float(t[0])
    print(json.dumps({"config": "C5 sharded codebook", "n_gpus": world, "N": N, "K": K, "d": d, "codes_per_gpu": K // world,
                      "step_ms": ms, "lookups_per_s": N / ms * 1e3, "sample_matches_unsharded_exact_scan": ok,
                      "step": "shard search + min-key all_reduce + codebook all_gather + gather/ST/loss + shard EMA"}),
          flush=True)
if world > 1:
    dist.destroy_process_group()
