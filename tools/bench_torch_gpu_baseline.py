""""also time the reference on the B200 with torch CUDA ops (row-chunked) -- that, not the CPU, is the
real bar for 'beat the reference on the same box'".  The reference's training forward (codebooks.py:350-435 +
vector_quantize_pytorch.py:261-273,362) written with the same torch calls on CUDA tensors, in row chunks because the
N x K fp32 similarities / one-hot of the full C2 batch need > 160 GiB; EMA statistics are accumulated over the chunks
and the refresh runs once.  Measurement aid only (not product code, does not touch oracle/)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "vector-quantization-by-ml_b200"))
import torch
import torch.nn.functional as F

dev = torch.device("cuda:0")
torch.backends.cuda.matmul.allow_tf32 = False
N, K, d, CH = 1 << 20, 8192, 256, 32768
g = torch.Generator(device=dev).manual_seed(1)
x = torch.randn(N, d, generator=g, device=dev).bfloat16()
emb = torch.randn(1, K, d, generator=g, device=dev) * 0.5
embed_avg, cluster_size = emb.clone(), torch.ones(1, K, device=dev)


def step():
    counts = torch.zeros(1, K, device=dev)
    sums = torch.zeros(1, K, d, device=dev)
    loss = torch.zeros((), device=dev)
    for r0 in range(0, N, CH):
        flat = x[r0:r0 + CH].float()[None]                               # codebooks.py:354
        sim = -torch.cdist(flat, emb)                                    # :386
        ind = sim.argmax(-1)                                             # utils/general.py:128
        onehot = F.one_hot(ind, K).type(flat.dtype)                      # :129
        quant = torch.einsum("h n c, h c d -> h n d", onehot, emb)       # codebooks.py:393-395
        out = flat + (quant - flat)                                      # vector_quantize_pytorch.py:273
        loss = loss + F.mse_loss(quant, flat, reduction="sum")           # :362 (mean over the whole batch below)
        counts += onehot.sum(1)                                          # codebooks.py:408
        sums += torch.einsum("h n d, h n c -> h c d", flat, onehot)      # :413
        del sim, onehot, quant, out
    cluster_size.lerp_(counts, 0.2)                                      # :411
    embed_avg.lerp_(sums, 0.2)                                           # :417
    cs = (cluster_size + 1e-5) / (cluster_size.sum(-1, keepdim=True) + K * 1e-5) * cluster_size.sum(-1, keepdim=True)
    emb.copy_(embed_avg / cs[..., None])                                 # :419-425
    return loss / (N * d)


for _ in range(2):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
print(json.dumps({"what": "reference algorithm with torch CUDA ops (fp32, TF32 off), row chunks of %d" % CH,
                  "config": "C2 N=1048576 K=8192 d=256", "ms_per_step": ms, "lookups_per_s": N / ms * 1e3}))
