"""Multi-GPU parity check, launched as: torchrun --nproc-per-node W --master-addr 127.0.0.1 tools/mgpu_check.py
 1. data parallel: W ranks, each quantising its slice, end up with the codebook a single process gets on the
    concatenated batch (cluster_size exact, embeddings/embed_avg 1e-6), identical on every rank, incl. expiry.
 2. sharded codebook: indices equal the un-sharded search; each shard's EMA result equals the matching rows of the
    un-sharded update.
Prints 'MGPU OK' from rank 0 on success."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "vector-quantization-by-ml_b200"))
import torch
import torch.distributed as dist

from vqb200 import CodebookParams, KmeansParameters, ShardedCodebook, VectorQuantize, ops

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def check(cond, msg):
    t = torch.tensor([0 if cond else 1], device=dev)
    dist.all_reduce(t)
    if int(t.item()):
        raise SystemExit(f"[rank {rank}] FAILED: {msg}")


# ---------------- 1. data parallel ----------------
K, d, n_per = 512, 64, 4096
g = torch.Generator().manual_seed(3)
x_all = torch.randn(world, n_per, d, generator=g)
c0 = torch.randn(1, K, d, generator=g) * 0.5
for thr in (0, 2):
    torch.manual_seed(0)
    dp = VectorQuantize(dim=d, codebook_params=CodebookParams(dim=d, codebook_size=K, threshold_ema_dead_code=thr,
                                                              kmeans_params=KmeansParameters()), sync_codebook=True).to(dev)
    single = VectorQuantize(dim=d, codebook_params=CodebookParams(dim=d, codebook_size=K, threshold_ema_dead_code=0),
                            sync_codebook=False).to(dev)
    for m in (dp, single):
        cb = m._codebook
        cb.embeddings.copy_(c0); cb.embed_avg.copy_(c0); cb.cluster_size.fill_(1.0 if thr == 0 else 0.5); cb.invalidate_cache()
        m.train()
    assert dp._codebook.use_ddp
    with torch.no_grad():
        q, ind, loss = dp(x_all[rank][None].to(dev))
        qs, inds, _ = single(x_all.reshape(1, -1, d).to(dev))
    check(torch.equal(ind[0], inds[0, rank * n_per:(rank + 1) * n_per]), "DP indices differ from single-process")
    if thr == 0:
        check(torch.equal(dp._codebook.cluster_size, single._codebook.cluster_size), "DP cluster_size")
        check(rel(dp._codebook.embed_avg, single._codebook.embed_avg) < 1e-6, "DP embed_avg")
        check(rel(dp._codebook.embeddings, single._codebook.embeddings) < 1e-6, "DP embeddings")
    # replicas identical on every rank (also after dead-code replacement)
    for name in ("embeddings", "embed_avg", "cluster_size"):
        mine = getattr(dp._codebook, name).contiguous()
        ref = mine.clone()
        dist.broadcast(ref, src=0)
        check(torch.equal(mine, ref), f"replicas diverged: {name} (thr={thr})")

# ---------------- 1b. distributed_replace_codes=False (reference codebooks.py:238-239) ----------------
# every rank samples its own replacement rows and the replacements are their mean over the ranks: replicas must stay
# identical, and the replaced codes must differ from any single rank's rows (they are means of W rows)
torch.manual_seed(1 + rank)            # different local draws on every rank
dpm = VectorQuantize(dim=d, codebook_params=CodebookParams(dim=d, codebook_size=K, threshold_ema_dead_code=2,
                                                           kmeans_params=KmeansParameters(),
                                                           distributed_replace_codes=False), sync_codebook=True).to(dev)
cbm = dpm._codebook
cbm.embeddings.copy_(c0); cbm.embed_avg.copy_(c0); cbm.cluster_size.fill_(0.5); cbm.invalidate_cache()
dpm.train()
with torch.no_grad():
    dpm(x_all[rank][None].to(dev))
for name in ("embeddings", "embed_avg", "cluster_size"):
    mine = getattr(cbm, name).contiguous()
    ref = mine.clone()
    dist.broadcast(ref, src=0)
    check(torch.equal(mine, ref), f"replicas diverged with distributed_replace_codes=False: {name}")

# ---------------- 2. sharded codebook ----------------
K, d, N = 4096, 64, 20000
g = torch.Generator().manual_seed(9)
x = torch.randn(N, d, generator=g).to(dev)
full = (torch.randn(K, d, generator=g) * 0.5).to(dev)
sh = ShardedCodebook(d, K).to(dev)
sh.load_full_codebook(full)
sh.train()
q, gidx, commit = sh(x)
cache = ops.prepare_codebook(full[None].contiguous(), False)
ref_idx, _, ws = ops.search(x[None], full[None].contiguous(), cache, False)
check(torch.equal(gidx, ref_idx[0]), f"sharded indices differ: {int((gidx != ref_idx[0]).sum())}")
check(torch.equal(q, x + (full[ref_idx[0]] - x)), "sharded quantize not bit-exact")
stats = ops.ema_reduce(x[None], ref_idx, None, K, bound_ws=ws)
cs, ea, em = torch.ones(1, K, device=dev), full[None].clone(), full[None].clone()
ops.ema_apply(stats, cs, ea, em, 1 - 0.8, 1e-5, False)
sl = slice(sh.offset, sh.offset + sh.shard_size)
check(torch.equal(sh.cluster_size[0], cs[0, sl]), "sharded cluster_size")
check(rel(sh.embed_avg[0], ea[0, sl]) < 1e-6 and rel(sh.embeddings[0], em[0, sl]) < 1e-5, "sharded EMA buffers")
if rank == 0:
    print("MGPU OK", flush=True)
dist.destroy_process_group()
