"""Per-kernel breakdown of one full training step (VectorQuantize.forward) at a named config.
usage: python tools/profile_step.py [c2|c3|c4|c5|c1]
       python -m torch.distributed.run --nproc-per-node W --master-addr 127.0.0.1 tools/profile_step.py c5
(under torchrun the codebook of c5 is sharded over the ranks and rank 0 prints its own kernel table, NCCL included)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "vector-quantization-by-ml_b200"))
import torch
from torch.profiler import ProfilerActivity, profile

from vqb200 import CodebookParams, KmeansParameters, ResidualVQ, VectorQuantize, ops

cfg = sys.argv[1] if len(sys.argv) > 1 else "c2"
world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(0)
g = torch.Generator(device=dev).manual_seed(1)


def seed_codebook(cb, scale=0.5, l2=False):
    c = torch.randn(cb.embeddings.shape, generator=g, device=dev) * scale
    if l2:
        c = torch.nn.functional.normalize(c, dim=-1)
    cb.embeddings.copy_(c); cb.embed_avg.copy_(c); cb.cluster_size.fill_(1.0); cb.invalidate_cache()


if cfg == "c2":
    mod = VectorQuantize(dim=256, codebook_params=CodebookParams(dim=256, codebook_size=8192)).to(dev)
    seed_codebook(mod._codebook)
    x = torch.randn(1024, 1024, 256, generator=g, device=dev).bfloat16()
elif cfg == "c3":
    mod = VectorQuantize(dim=512, codebook_params=CodebookParams(dim=512, codebook_size=16384, use_cosine_sim=True,
                         transform_input="l2norm", weights_regularization="l2norm")).to(dev)
    seed_codebook(mod._codebook, l2=True)
    x = torch.randn(512, 1024, 512, generator=g, device=dev)
elif cfg == "c4":
    mod = ResidualVQ(dim=512, num_quantizers=8, codebook_params=CodebookParams(dim=512, codebook_size=1024)).to(dev)
    for i, l in enumerate(mod.layers):
        seed_codebook(l._codebook, scale=0.5 / (1.5 ** i))
    x = torch.randn(64, 4096, 512, generator=g, device=dev)
elif cfg == "c5":
    mod = VectorQuantize(dim=64, codebook_params=CodebookParams(dim=64, codebook_size=65536, threshold_ema_dead_code=0),
                         sync_codebook=world > 1).to(dev)
    mod._codebook.load_full_codebook(torch.randn(1, 65536, 64, generator=torch.Generator().manual_seed(0)) * 0.5)
    mod._codebook.sharded_input = "replicated"
    x = torch.randn(4096, 1024, 64, generator=torch.Generator(device=dev).manual_seed(4321), device=dev)
else:
    mod = VectorQuantize(dim=256, codebook_params=CodebookParams(dim=256, codebook_size=512, threshold_ema_dead_code=0)).to(dev)
    x = torch.randn(1, 1024, 256, generator=g, device=dev)
mod.train()
with torch.no_grad():
    for _ in range(3):
        mod(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        mod(x)
    e1.record()
    torch.cuda.synchronize()
    if rank == 0:
        print(f"{cfg} (world {world}): {e0.elapsed_time(e1) / 5:.3f} ms per step", flush=True)
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(3):
            mod(x)
        torch.cuda.synchronize()
if rank == 0:
    print("(3 profiled steps: divide the totals by 3)")
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=64), flush=True)
if rank == 0 and os.environ.get("VQB_TIMELINE"):
    # device timeline of the LAST profiled step: every kernel with its start offset, duration and the idle gap before it
    from torch.autograd import DeviceType
    evs = sorted((e for e in prof.events() if e.device_type == DeviceType.CUDA), key=lambda e: e.time_range.start)
    per = len(evs) // 3
    last = evs[-per:]
    t0, prev_end, idle = last[0].time_range.start, last[0].time_range.start, 0.0
    print(f"timeline of one step ({per} device activities): start_us  dur_us  gap_before_us  name")
    for e in last:
        gap = e.time_range.start - prev_end
        idle += max(gap, 0.0)
        print(f"  {e.time_range.start - t0:9.1f} {e.time_range.end - e.time_range.start:9.1f} {gap:7.1f}  {e.name[:90]}")
        prev_end = max(prev_end, e.time_range.end)
    print(f"span {prev_end - t0:.1f} us, idle between kernels {idle:.1f} us")
if world > 1:
    dist.destroy_process_group()
