"""C4 step time with the dead-code check (one host sync per level, as the reference does) and without it."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "vector-quantization-by-ml_b200"))
import torch
from vqb200 import CodebookParams, ResidualVQ
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1)
x = torch.randn(64, 4096, 512, generator=g, device=dev)
for thr in (2, 0):
    torch.manual_seed(0)
    rvq = ResidualVQ(dim=512, num_quantizers=8, codebook_params=CodebookParams(dim=512, codebook_size=1024,
                     threshold_ema_dead_code=thr)).to(dev).train()
    for i, l in enumerate(rvq.layers):
        cb = l._codebook
        c = torch.randn(cb.embeddings.shape, generator=g, device=dev) * 0.5 / 1.4 ** i
        cb.embeddings.copy_(c); cb.embed_avg.copy_(c); cb.cluster_size.fill_(100.0); cb.invalidate_cache()
    with torch.no_grad():
        for _ in range(3):
            rvq(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            rvq(x)
        e1.record(); torch.cuda.synchronize()
    print(f"threshold_ema_dead_code={thr}: {e0.elapsed_time(e1) / 5:.3f} ms per step", flush=True)
