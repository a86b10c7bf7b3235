"""Multi-GPU parity check, launched as: torchrun --nproc-per-node W --master-addr 127.0.0.1 tools/mgpu_check.py
 1. data parallel: W ranks, each quantising its slice, end up with the codebook a single process gets on the
    concatenated batch (cluster_size exact, embeddings/embed_avg 1e-6), identical on every rank, incl. expiry.
 1c. the same data-parallel step against the CPU ORACLE on the rank-concatenated batch.
 2. sharded codebook through the module API (Codebook.sharded; replicated and all_gather inputs) against the CPU
    ORACLE on the un-sharded codebook: indices, bit-exact quantize, every shard's EMA + dead-code expiry result equals
    the matching rows of the reference update, state_dict holds the full tensors.
Prints 'MGPU OK' from rank 0 on success."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "vector-quantization-by-ml_b200"))
import torch
import torch.distributed as dist

from oracle import vq_oracle as O  # noqa: E402  (test tool: the oracle is the checker)
from vqb200 import CodebookParams, KmeansParameters, VectorQuantize, distributed as D, ops

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

# ---------------- 1c. data parallel against the ORACLE (single process on the rank-concatenated batch) ----------------
torch.manual_seed(0)
dpo = VectorQuantize(dim=d, codebook_params=CodebookParams(dim=d, codebook_size=K, threshold_ema_dead_code=0),
                     sync_codebook=True).to(dev).train()
cbo = dpo._codebook
cbo.embeddings.copy_(c0); cbo.embed_avg.copy_(c0); cbo.cluster_size.fill_(1.0); cbo.invalidate_cache()
with torch.no_grad():
    qd, indd, _ = dpo(x_all[rank][None].to(dev))
sto = O.CodebookState(c0.clone(), c0.clone(), torch.ones(1, K))
qo, io, lo, exo = O.vq_forward(sto, x_all.reshape(1, -1, d), O.VQOpts(codebook=O.CodebookOpts(threshold_ema_dead_code=0)),
                               want_gap=True)
rows = slice(rank * n_per, (rank + 1) * n_per)
diff = indd[0].cpu() != io[0, rows]
check(bool((exo["top2_rel_gap"][0, rows][diff] < 1e-6).all()), "DP indices differ from the ORACLE outside its 1e-6 window")
if not bool(diff.any()):
    check(torch.equal(qd[0].cpu(), qo[0, rows]), "DP quantize vs oracle")
    check(torch.equal(cbo.cluster_size.cpu(), sto.cluster_size), "DP cluster_size vs oracle")
    check(rel(cbo.embeddings.cpu(), sto.embeddings) < 1e-5, "DP embeddings vs oracle")

# ---------------- 1d. ResidualVQ data parallel: CUDA graph (with the captured statistics all_reduce) == eager ----------------
from vqb200 import ResidualVQ  # noqa: E402


def _make_rvq(thr):
    torch.manual_seed(0)
    m = ResidualVQ(dim=64, num_quantizers=4, codebook_params=CodebookParams(dim=64, codebook_size=96,
                                                                            threshold_ema_dead_code=thr),
                   sync_codebook=True).to(dev).train()
    gg = torch.Generator().manual_seed(3)
    for li, layer in enumerate(m.layers):
        c = (torch.randn(1, 96, 64, generator=gg) * (0.6 / 1.5 ** li)).to(dev)
        cb = layer._codebook
        cb.embeddings.copy_(c); cb.embed_avg.copy_(c); cb.cluster_size.fill_(1.0); cb.invalidate_cache()
    return m


for thr in (0, 2):
    eager, graphed = _make_rvq(thr), _make_rvq(thr)
    graphed.enable_cuda_graph(data_parallel=True)
    gg = torch.Generator().manual_seed(11 + rank)
    x_e = torch.empty(3, 700, 64, device=dev)
    x_g = torch.empty(3, 700, 64, device=dev)
    for step in range(5):
        xx = torch.randn(3, 700, 64, generator=gg)
        x_e.copy_(xx); x_g.copy_(xx)
        torch.manual_seed(100 + step)
        with torch.no_grad():
            qe, ie, le = eager(x_e)
        torch.manual_seed(100 + step)
        with torch.no_grad():
            qg, ig, lg = graphed(x_g)
        check(torch.equal(qe, qg) and torch.equal(ie, ig) and torch.equal(le, lg), f"RVQ DP graph step {step} thr {thr}: outputs")
        for le_, lg_ in zip(eager.layers, graphed.layers):
            for name in ("embeddings", "embed_avg", "cluster_size"):
                check(torch.equal(getattr(le_._codebook, name), getattr(lg_._codebook, name)),
                      f"RVQ DP graph step {step} thr {thr}: {name}")
    check(any(isinstance(v, tuple) for v in graphed._graphs.values()), "RVQ DP graph was never captured")
    # a captured graph holds NCCL kernels of this communicator: release it before the process group goes away
    # (destroy_process_group with such a graph alive never returned on two B200s)
    graphed.enable_cuda_graph(False)
    del eager, graphed
import gc  # noqa: E402
gc.collect()
torch.cuda.synchronize()
if rank == 0:
    print("1d. ResidualVQ data-parallel CUDA graph == eager: ok", flush=True)

# ---------------- 2. sharded codebook, through the module API, against the ORACLE (un-sharded reference path) -------
# codebooks.py:350-435 on the whole codebook is the oracle; the sharded path must give its indices (outside the
# reference's own 1e-6 ties), bit-exact quantize, and on every rank the matching rows of its EMA + expiry result
D.SHARD_MIN_CODES = 4096
K, d, N = 4096, 64, 20000
g = torch.Generator().manual_seed(9)
x = torch.randn(N, d, generator=g)
full = torch.randn(K, d, generator=g) * 0.5
full[3000] = full[11]                 # duplicate in another shard: the lowest index must win every row
full[100:104] *= 30.0                 # never chosen -> die -> replaced by the rows rank 0 draws
for mode in ("replicated", "all_gather"):
    torch.manual_seed(0)
    vq = VectorQuantize(dim=d, codebook_params=CodebookParams(dim=d, codebook_size=K, threshold_ema_dead_code=2),
                        sync_codebook=True).to(dev).train()
    cb = vq._codebook
    check(cb.sharded and cb.embeddings.shape[1] == K // world, "Codebook did not shard")
    cb.sharded_input = mode
    cb.load_full_codebook(full)
    n_loc = N // world
    x_in = x if mode == "replicated" else x[rank * n_loc:(rank + 1) * n_loc]
    x_ref = x if mode == "replicated" else x[:n_loc * world]
    from vqb200.codebook import Codebook
    Codebook._draw_rows = staticmethod(lambda n, m, device: (torch.randperm(n)[:m] if n >= m else torch.randint(0, n, (m,))).to(device))
    torch.manual_seed(21)
    with torch.no_grad():
        q, gidx, loss = vq(x_in[None].to(dev))
    st = O.CodebookState(full[None].clone(), full[None].clone(), torch.ones(1, K))
    torch.manual_seed(21)
    qo, io, lo, ex = O.vq_forward(st, x_ref[None], O.VQOpts(codebook=O.CodebookOpts(threshold_ema_dead_code=2)), want_gap=True)
    rows = slice(0, N) if mode == "replicated" else slice(rank * n_loc, (rank + 1) * n_loc)
    diff = gidx[0].cpu() != io[0, rows]
    check(bool((ex["top2_rel_gap"][0, rows][diff] < 1e-6).all()),
          f"sharded ({mode}) indices differ from the oracle outside its 1e-6 window: {int(diff.sum())}")
    check(bool((gidx != 3000).all()), "duplicate code: the higher index won")
    same = ~diff
    check(torch.equal(q[0].cpu()[same], qo[0, rows][same]), f"sharded ({mode}) quantize not bit-exact")
    sl = slice(cb.shard_offset, cb.shard_offset + cb.shard_size)
    n_flip = torch.tensor([int(diff.sum())], device=dev)
    dist.all_reduce(n_flip)
    if int(n_flip) == 0:
        if mode == "replicated":
            check(abs(float(loss) - float(lo)) <= 1e-5 * float(lo), "sharded loss")
        check(torch.equal(cb.cluster_size[0].cpu(), st.cluster_size[0, sl]), f"sharded ({mode}) cluster_size")
        check(rel(cb.embed_avg[0].cpu(), st.embed_avg[0, sl]) < 1e-5 and rel(cb.embeddings[0].cpu(), st.embeddings[0, sl]) < 1e-5,
              f"sharded ({mode}) EMA buffers / expiry")
        check(int((st.cluster_size[0] == 2.0).sum()) >= 4, "expiry did not fire in the sharded scenario")
    sd = vq.state_dict()
    check(tuple(sd["_codebook.embeddings"].shape) == (1, K, d), "state_dict must hold the full codebook")
if rank == 0:
    print("MGPU OK", flush=True)
torch.cuda.synchronize()
dist.barrier()
dist.destroy_process_group()
