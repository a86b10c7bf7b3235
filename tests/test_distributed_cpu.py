"""world_size-2 Gloo tests (CPU) of the multi-GPU host logic: replica-consistent replacement sampling, the packed
statistics all_reduce used by data parallel, and the min-key merge used by the sharded codebook."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _pack(score: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """Pure-torch mirror of vqb_minkey_pack (csrc/ema.cu), test-side only."""
    u = score.view(torch.int32).to(torch.int64) & 0xFFFFFFFF
    neg = (u >> 31) == 1
    u = torch.where(neg, (~u) & 0xFFFFFFFF, u | 0x80000000)
    u = u ^ 0x80000000
    key = (u << 32) | (idx & 0xFFFFFFFF)
    return torch.where(key >= (1 << 63), key - (1 << 64), key) if False else (key - ((key >> 63) << 64))


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "vector-quantization-by-ml_b200"))
    from vqb200 import distributed as D
    from vqb200.codebook import Codebook
    res = {}
    # 1. replica-consistent sampling: same vectors on both ranks, every vector is a row of some rank's data
    torch.manual_seed(100 + rank)
    local = torch.randn(50 + 10 * rank, 8) + 100 * rank
    torch.manual_seed(7)          # rank 0's multinomial + both ranks' local draws
    s = D.sample_vectors_distributed(local, 23, Codebook._draw_rows)
    gathered = [torch.empty_like(s) for _ in range(world)]
    dist.all_gather(gathered, s)
    res["same"] = all(torch.equal(gathered[0], g) for g in gathered)
    res["shape"] = tuple(s.shape)
    alls = [torch.empty(50 + 10 * r, 8) for r in range(world)]
    for r in range(world):
        buf = local if r == rank else alls[r]
        dist.broadcast(buf, src=r)
        alls[r] = buf
    union = torch.cat(alls)
    res["member"] = bool(((s[:, None, :] == union[None]).all(-1).any(-1)).all())
    # ... and they are exactly the rows one process would draw from the concatenated batch with the same generator
    torch.manual_seed(7)
    res["as_single"] = torch.equal(s, union[Codebook._draw_rows(union.shape[0], 23, union.device)])
    # 2. data parallel: ONE packed all_reduce of (H,K,d+1) statistics == statistics of the concatenated batch
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 64, 4, generator=g)                      # (rank, rows, d)
    idx = torch.randint(0, 6, (2, 64), generator=g)
    def stats_of(xr, ir):
        oh = torch.nn.functional.one_hot(ir, 6).float()
        return torch.cat([torch.einsum("nd,nc->cd", xr, oh), oh.sum(0)[:, None]], -1)[None]
    st = stats_of(x[rank], idx[rank])
    D.all_reduce_sum(st)
    ref = stats_of(x.reshape(-1, 4), idx.reshape(-1))
    res["stats"] = bool(torch.allclose(st, ref, rtol=1e-6, atol=1e-6)) and torch.equal(st[..., -1], ref[..., -1])
    # 3. sharded codebook merge: min over ranks of (score, global index) keys, lowest index on ties
    score = torch.tensor([[1.5, -2.0, 0.0, 3.0, -0.0, 7.0], [1.5, -2.5, 0.0, 2.0, 0.0, 7.0]])[rank]
    gidx = torch.tensor([[3, 9, 4, 1, 2, 5], [11, 8, 12, 10, 15, 13]])[rank]
    keys = _pack(score, gidx)
    D.merge_min_keys(keys)
    res["win_idx"] = (keys & 0xFFFFFFFF).tolist()
    # 4. distributed_replace_codes=False: locally sampled replacements averaged over the ranks
    loc = torch.full((3, 4), float(rank + 1))
    res["mean"] = D.maybe_distributed_mean(loc).tolist()
    if rank == 0:
        torch.save(res, out)
    dist.destroy_process_group()


def test_two_rank_gloo_host_logic(tmp_path):
    out = str(tmp_path / "res.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    res = torch.load(out)
    assert res["same"] and res["shape"] == (23, 8) and res["member"] and res["as_single"]
    assert res["stats"]
    # scores: equal 1.5 -> lower index 3; -2.5 < -2.0 -> 8; 0.0 tie -> 4; 2.0 < 3.0 -> 10; -0.0 vs 0.0: -0.0 sorts first -> 2; 7 tie -> 5
    assert res["win_idx"] == [3, 8, 4, 10, 2, 5]
    assert res["mean"] == [[1.5] * 4] * 3


def _dp_worker(rank, world, port, out):
    """Data parallel through the PRODUCT's orchestration (Codebook._run, expire_codes_, _kmeans_init) on Gloo, with the
    tensor-level wrappers replaced by plain-torch statements (tests/cpu_kernels.py)."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, os.path.join(root, "tests"), os.path.join(root, "vector-quantization-by-ml_b200")):
        sys.path.insert(0, p)
    import cpu_kernels
    from oracle import vq_oracle as O
    from vqb200 import CodebookParams, KmeansParameters, VectorQuantize, _lib, ops
    cpu_kernels.install(ops, _lib)
    res = {}
    K, d, n_per = 48, 8, 160
    g = torch.Generator().manual_seed(3)
    x_all = torch.randn(world, n_per, d, generator=g)
    c0 = torch.randn(1, K, d, generator=g) * 0.5

    def replicas_identical(cb):
        ok = True
        for name in ("embeddings", "embed_avg", "cluster_size"):
            mine = getattr(cb, name).detach().clone().contiguous()
            ref = mine.clone()
            dist.broadcast(ref, src=0)
            ok &= torch.equal(mine, ref)
        return ok

    def make(thr, cs0=None, **cp):
        torch.manual_seed(0)
        vq = VectorQuantize(dim=d, codebook_params=CodebookParams(dim=d, codebook_size=K, threshold_ema_dead_code=thr,
                                                                  kmeans_params=KmeansParameters(), **cp),
                            sync_codebook=True).train()
        cb = vq._codebook
        with torch.no_grad():
            cb.embeddings.copy_(c0); cb.embed_avg.copy_(c0)
            cb.cluster_size.fill_(cs0 if cs0 is not None else (1.0 if thr == 0 else 0.5))
        return vq, cb

    # 1. EMA step without expiry: every rank ends with the codebook a single process gets on the concatenated batch
    vq, cb = make(0)
    assert cb.use_ddp
    with torch.no_grad():
        q, ind, loss = vq(x_all[rank][None])
    st = O.CodebookState(c0.clone(), c0.clone(), torch.ones(1, K))
    _, ind_all, _, _ = O.vq_forward(st, x_all.reshape(1, -1, d), O.VQOpts(codebook=O.CodebookOpts(threshold_ema_dead_code=0)))
    res["ind"] = torch.equal(ind[0], ind_all[0, rank * n_per:(rank + 1) * n_per])
    res["cs"] = torch.equal(cb.cluster_size, st.cluster_size)
    res["emb"] = float((cb.embeddings - st.embeddings).abs().max() / st.embeddings.abs().max()) < 1e-6
    res["same0"] = replicas_identical(cb)
    # which codes die in cases 2 and 3: the same step without expiry, from the same start
    twin, cbt = make(0, cs0=0.5)
    with torch.no_grad():
        twin(x_all[rank][None])
    dead = cbt.cluster_size[0] < 2
    # 2. expiry, replacements sampled from the union of the ranks' batches (distributed_replace_codes=True)
    vq, cb = make(2)
    torch.manual_seed(11)
    with torch.no_grad():
        vq(x_all[rank][None])
    replaced = dead
    res["reset2"] = bool((cb.cluster_size[0][dead] == 2.0).all()) and \
        torch.equal(cb.embeddings[0][~dead], cbt.embeddings[0][~dead])
    res["n_replaced"] = int(replaced.sum())
    res["same2"] = replicas_identical(cb)
    union = x_all.reshape(-1, d)
    res["member2"] = bool((cb.embeddings[0][replaced][:, None, :] == union[None]).all(-1).any(-1).all())
    # 3. distributed_replace_codes=False: locally sampled rows, averaged over the ranks (reference codebooks.py:238-239)
    vq, cb = make(2, distributed_replace_codes=False)
    torch.manual_seed(100 + rank)                      # the local draws differ between the ranks
    with torch.no_grad():
        vq(x_all[rank][None])
    replaced = dead
    res["same3"] = replicas_identical(cb)
    res["n_replaced3"] = int(replaced.sum())
    rows = cb.embeddings[0][replaced]
    res["member3"] = bool((rows[:, None, :] == union[None]).all(-1).any(-1).any())     # means of two rows: not rows
    halves = (x_all[0][:, None, :] + x_all[1][None, :, :]) / 2                        # all pair means (world = 2)
    res["pairmean3"] = bool(((rows[:, None, None, :] - halves[None]).abs().amax(-1) < 1e-6).flatten(1).any(-1).all())
    # 4. kmeans init under data parallel (sync'd): identical centroids on every rank
    torch.manual_seed(0)
    vqk = VectorQuantize(dim=d, codebook_params=CodebookParams(dim=d, codebook_size=12, threshold_ema_dead_code=0,
                                                               initialization_by_kmeans=True,
                                                               kmeans_params=KmeansParameters()),
                         sync_codebook=True).train()
    torch.manual_seed(5)
    with torch.no_grad():
        vqk(x_all[rank][None])
    res["same_kmeans"] = replicas_identical(vqk._codebook) and bool(vqk._codebook.is_initialized)
    # 5. sharded codebook THROUGH THE MODULE API (Codebook.sharded: rows split over the ranks, cross-rank (score, index)
    #    min-key merge): indices equal the un-sharded reference path, every shard's EMA + expiry result equals the
    #    matching rows of the un-sharded update, kmeans init on shards, state_dict round trip with full-size tensors
    from vqb200 import distributed as D
    D.SHARD_MIN_CODES = 32
    Ks, Ns = 64, 300
    gs = torch.Generator().manual_seed(9)
    xs = torch.randn(world, Ns, d, generator=gs)
    full = torch.randn(Ks, d, generator=gs) * 0.5
    full[40] = full[7]                                   # a duplicated code across the two shards: lowest index wins
    full[50:53] *= 30.0                                  # never chosen: these die and are replaced (expiry on a shard)

    def sharded_vq(thr, mode, **cp):
        torch.manual_seed(0)
        vq = VectorQuantize(dim=d, codebook_params=CodebookParams(dim=d, codebook_size=Ks, threshold_ema_dead_code=thr,
                                                                  **cp), sync_codebook=True).train()
        vq._codebook.sharded_input = mode
        return vq, vq._codebook

    for mode in ("replicated", "all_gather"):
        vq, cb = sharded_vq(2, mode)
        res[f"sh_is_sharded_{mode}"] = bool(cb.sharded) and cb.embeddings.shape == (1, Ks // world, d)
        cb.load_full_codebook(full)
        x_in = xs[0][None] if mode == "replicated" else xs[rank][None]
        x_ref = xs[0][None] if mode == "replicated" else xs.reshape(1, -1, d)
        torch.manual_seed(21)
        with torch.no_grad():
            qs, gidx, loss = vq(x_in)
        st5 = O.CodebookState(full[None].clone(), full[None].clone(), torch.ones(1, Ks))
        torch.manual_seed(21)
        q5, i5, l5, _ = O.vq_forward(st5, x_ref, O.VQOpts(codebook=O.CodebookOpts(threshold_ema_dead_code=2)))
        rows = slice(0, Ns) if mode == "replicated" else slice(rank * Ns, (rank + 1) * Ns)
        sl = slice(cb.shard_offset, cb.shard_offset + cb.shard_size)
        res[f"sh_idx_{mode}"] = torch.equal(gidx[0], i5[0, rows]) and bool((gidx != 40).all())
        res[f"sh_q_{mode}"] = torch.equal(qs[0], q5[0, rows])
        if mode == "replicated":
            res["sh_loss"] = bool(torch.allclose(loss, l5, rtol=1e-6))
        res[f"sh_cs_{mode}"] = torch.equal(cb.cluster_size[0], st5.cluster_size[0, sl])
        res[f"sh_emb_{mode}"] = float((cb.embeddings[0] - st5.embeddings[0, sl]).abs().max()) < 1e-5 and \
            float((cb.embed_avg[0] - st5.embed_avg[0, sl]).abs().max()) < 1e-5
        res[f"sh_expired_{mode}"] = int((st5.cluster_size[0] == 2.0).sum())
        sd = vq.state_dict()
        res[f"sh_sd_{mode}"] = tuple(sd["_codebook.embeddings"].shape) == (1, Ks, d) and \
            float((sd["_codebook.embeddings"] - st5.embeddings).abs().max()) < 1e-5
        vq2, cb2 = sharded_vq(2, mode)
        vq2.load_state_dict(sd)
        res[f"sh_load_{mode}"] = torch.equal(cb2.embeddings, cb.embeddings) and torch.equal(cb2.cluster_size, cb.cluster_size)
    # kmeans init on shards == the reference's kmeans on the un-sharded centroids (rank 0's draw)
    vqk, cbk = sharded_vq(0, "replicated", initialization_by_kmeans=True, kmeans_params=KmeansParameters())
    torch.manual_seed(33)
    with torch.no_grad():
        _, ik, _ = vqk(xs[0][None])
    stk = O.CodebookState.fresh(1, Ks, d, kmeans_init=True)
    torch.manual_seed(33)
    _, iko, _, _ = O.vq_forward(stk, xs[0][None], O.VQOpts(codebook=O.CodebookOpts(threshold_ema_dead_code=0)))
    sl = slice(cbk.shard_offset, cbk.shard_offset + cbk.shard_size)
    res["sh_kmeans"] = torch.equal(ik, iko) and torch.equal(cbk.cluster_size[0], stk.cluster_size[0, sl]) and \
        float((cbk.embeddings[0] - stk.embeddings[0, sl]).abs().max()) < 1e-5
    if rank == 0:
        torch.save(res, out)
    dist.destroy_process_group()


def test_two_rank_gloo_data_parallel_orchestration(tmp_path):
    out = str(tmp_path / "dp.pt")
    mp.spawn(_dp_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    res = torch.load(out)
    assert res["ind"] and res["cs"] and res["emb"] and res["same0"], res
    assert res["n_replaced"] > 0 and res["same2"] and res["member2"] and res["reset2"], res
    assert res["n_replaced3"] > 0 and res["same3"] and not res["member3"] and res["pairmean3"], res
    assert res["same_kmeans"], res
    for mode in ("replicated", "all_gather"):
        for key in ("is_sharded", "idx", "q", "cs", "emb", "sd", "load"):
            assert res[f"sh_{key}_{mode}"], (key, mode, res)
        assert res[f"sh_expired_{mode}"] >= 3, res
    assert res["sh_loss"] and res["sh_kmeans"], res


def _rvq_dp_worker(rank, world, port, out):
    """ResidualVQ data parallel through the FUSED level loop (`_fused_levels`: per-level statistics all_reduce launched
    asynchronously, refreshes applied after the loop) on Gloo: every rank must end with the codebooks one process gets on
    the rank-concatenated batch (the oracle's rvq_forward), with and without dead-code expiry."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, os.path.join(root, "tests"), os.path.join(root, "vector-quantization-by-ml_b200")):
        sys.path.insert(0, p)
    import cpu_kernels
    from oracle import vq_oracle as O
    from vqb200 import CodebookParams, ResidualVQ, _lib, ops, rvq as RVQ
    cpu_kernels.install(ops, _lib)
    real_can_fuse = RVQ.ResidualVQ._can_fuse
    fused_calls = []

    class _OnDevice:
        def __init__(self, t):
            self.is_cuda, self.ndim, self.requires_grad, self.shape = True, t.ndim, t.requires_grad, t.shape

    def can_fuse(self, x, dropout_active):
        ok = real_can_fuse(self, _OnDevice(x), dropout_active)
        fused_calls.append(ok)
        return ok
    RVQ.ResidualVQ._can_fuse = can_fuse
    res = {}
    Q, K, d, b, n = 3, 40, 8, 2, 90
    g = torch.Generator().manual_seed(4)
    x_all = torch.randn(world * b, n, d, generator=g)
    cbs = [torch.randn(1, K, d, generator=g) * (0.6 / 1.5 ** li) for li in range(Q)]
    for thr in (0, 2):
        torch.manual_seed(0)
        m = ResidualVQ(dim=d, num_quantizers=Q, codebook_params=CodebookParams(dim=d, codebook_size=K,
                                                                               threshold_ema_dead_code=thr),
                       sync_codebook=True).train()
        sts = []
        for layer, c in zip(m.layers, cbs):
            cb = layer._codebook
            with torch.no_grad():
                cb.embeddings.copy_(c); cb.embed_avg.copy_(c); cb.cluster_size.fill_(1.0)
            sts.append(O.CodebookState(c.clone(), c.clone(), torch.ones(1, K)))
        with torch.no_grad():
            q, ind, loss = m(x_all[rank * b:(rank + 1) * b])
        opts = O.VQOpts(codebook=O.CodebookOpts(threshold_ema_dead_code=0))
        qo, io, lo, _ = O.rvq_forward(sts, x_all, opts, training=True)
        res[f"ind{thr}"] = torch.equal(ind, io[rank * b:(rank + 1) * b])
        res[f"q{thr}"] = torch.equal(q, qo[rank * b:(rank + 1) * b])
        same = True
        for layer, st in zip(m.layers, sts):
            cb = layer._codebook
            if thr == 0:
                same &= torch.equal(cb.cluster_size, st.cluster_size)
                same &= float((cb.embeddings - st.embeddings).abs().max() / st.embeddings.abs().max()) < 1e-6
            for name in ("embeddings", "embed_avg", "cluster_size"):      # replicas identical, expiry included
                mine = getattr(cb, name).detach().clone().contiguous()
                ref = mine.clone()
                dist.broadcast(ref, src=0)
                same &= torch.equal(mine, ref)
            if thr == 2:
                same &= bool((cb.cluster_size >= 2.0 - 1e-6).all())       # every dead code was replaced
        res[f"books{thr}"] = same
    res["fused"] = all(fused_calls) and len(fused_calls) == 2
    if rank == 0:
        torch.save(res, out)
    dist.destroy_process_group()


def test_two_rank_gloo_residual_vq_fused_loop_data_parallel(tmp_path):
    out = str(tmp_path / "rvq_dp.pt")
    mp.spawn(_rvq_dp_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    res = torch.load(out)
    assert all(res.values()), res
