"""Parity at BASELINE.json's full sizes through size-independent properties (the CPU oracle cannot hold N x K there):
  * a random sample of rows is checked against the CPU oracle (indices identical except rows whose reference fp32
    top-2 gap is < 1e-6 relative);
  * idempotence: the quantised vectors are assigned to their own codes;
  * quantize is bit-exact given the indices: q == fl(x + fl(C[idx] - x)) (same IEEE ops, evaluated by torch);
  * EMA statistics: counts sum to N, per-code sums add up to the column sums of x (linearity), bitwise reproducible.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _sample_check(x, c, idx, cos, n_sample=2048, seed=0):
    from oracle import vq_oracle as O
    g = torch.Generator().manual_seed(seed)
    rows = torch.randperm(x.shape[1], generator=g)[:n_sample]
    xs = x[0, rows.to(x.device)].float().cpu()[None]
    sim = O.similarities(xs, c.cpu(), cos)
    ref = sim.argmax(-1)[0]
    top2 = sim.topk(2, -1).values[0]
    gap = (top2[:, 0] - top2[:, 1]).abs() / top2[:, 0].abs().clamp_min(1e-30)
    got = idx[0, rows.to(idx.device)].cpu()
    bad = (got != ref) & (gap >= 1e-6)
    if bool(bad.any()):
        # STRICT (north star): a row may differ from the reference only inside the reference's own fp32 1e-6 window.
        # Round 1 accepted rows just outside it when the CUDA index was the fp64 argmin, on the hypothesis that the
        # oracle's own fp32 noise mis-orders them; tools/oracle_noise.py measured that noise on the CPU
        # (profiles/r02_oracle_noise.jsonl: at every named shape, under 6 thread / MKL-kernel settings, each row where
        # the oracle differs from fp64 has an oracle gap < 1e-6 and misses the true winner by <= 1.6e-7) -- the
        # hypothesis does not hold, so there is no arbitration any more.  Everything needed to diagnose is printed.
        br = rows[bad]
        xs64, c64 = x[0, br.to(x.device)].double().cpu(), c[0].double().cpu()
        s64 = xs64 @ c64.T if cos else -torch.cdist(xs64, c64)
        arg64 = s64.argmax(-1)
        t2 = s64.topk(2, -1).values
        gap64 = (t2[:, 0] - t2[:, 1]).abs() / t2[:, 0].abs().clamp_min(1e-30)
        from vqb200 import ops
        ex, _, _ = ops.search(x[:, br.to(x.device)].contiguous(), c, None, cos, force_exact=True)
        again, _, ws = ops.search(x, c, ops.prepare_codebook(c, cos), cos)
        detail = [(int(r), int(got[bad][i]), int(ref[bad][i]), int(arg64[i]), int(ex[0, i]), int(again[0, int(r)]),
                   float(gap[bad][i]), float(gap64[i])) for i, r in enumerate(br[:8])]
        raise AssertionError(f"{int(bad.sum())} sampled rows differ from the oracle outside its 1e-6 window; "
                             f"(row, got, oracle, fp64 argmin, exact scan, second search, oracle gap, fp64 gap): "
                             f"{detail}; whole batch vs second search: {int((again != idx).sum())} rows differ; "
                             f"stats {ops.search_stats(ws)}")
    return int((got != ref).sum())


def _stats_props(x, idx, K):
    from vqb200 import ops
    a = ops.ema_reduce(x, idx, None, K)
    b = ops.ema_reduce(x, idx, None, K)
    assert torch.equal(a, b), "EMA statistics not bitwise reproducible"
    d = x.shape[-1]
    assert float(a[..., d].sum()) == float(x.shape[1])
    assert torch.equal(a[0, :, d], torch.bincount(idx[0], minlength=K).float())
    col = x[0].double().sum(0)
    got = a[0, :, :d].double().sum(0)
    assert float((got - col).abs().max()) <= 1e-5 * float(col.abs().max()) + 1e-3


@pytest.mark.parametrize("name,N,K,d,cos,dtype", [
    ("C2", 1 << 20, 8192, 256, False, torch.bfloat16),
    ("C3", 1 << 19, 16384, 512, True, torch.float32),
    ("C5", 1 << 22, 65536, 64, False, torch.float32),
])
def test_search_gather_ema_at_full_size(name, N, K, d, cos, dtype):
    from vqb200 import ops
    g = torch.Generator(device=DEV).manual_seed(11)
    x = torch.randn(1, N, d, generator=g, device=DEV)
    if cos:
        x = ops.l2norm_rows(x)
    x = x.to(dtype).contiguous()
    c = torch.randn(1, K, d, generator=g, device=DEV) * 0.5
    if cos:
        c = torch.nn.functional.normalize(c, dim=-1)
    cache = ops.prepare_codebook(c, cos)
    idx, _, ws = ops.search(x, c, cache, cos)
    st = ops.search_stats(ws)
    assert st["tensor_core_pass"] == 1
    assert int(idx.min()) >= 0 and int(idx.max()) < K
    _sample_check(x, c, idx, cos)
    # quantize: bit-exact identity + idempotence of the assignment
    q, loss = ops.gather_st_loss(x, c, idx, None, True, True)
    xf = x.float()
    cq = c[0][idx[0]][None]
    assert torch.equal(q, xf + (cq - xf)), "quantize not bit-exact given the indices"
    ref_loss = ((cq - xf).double() ** 2).mean()
    assert abs(float(loss[0]) - float(ref_loss)) <= 1e-5 * float(ref_loss)
    sub = torch.arange(0, N, max(1, N // 65536), device=DEV)
    idx2, _, _ = ops.search(cq[:, sub].contiguous(), c, cache, cos)
    same = idx2[0] == idx[0, sub]
    if not bool(same.all()):     # only exact duplicate / tied codes may differ
        a, b = c[0][idx2[0][~same]], c[0][idx[0, sub][~same]]
        assert float((a - b).abs().max()) < 1e-6
    _stats_props(x, idx, K)
    del q, cq, xf
    torch.cuda.empty_cache()


def test_rvq_c4_shape_fused_equals_generic():
    """C4: ResidualVQ 8 levels x K=1024 x d=512 (a slice of the 64x4096 batch): fused level kernel == generic loop."""
    from vqb200 import CodebookParams, ResidualVQ
    torch.manual_seed(0)
    mods = []
    for _ in range(2):
        m = ResidualVQ(dim=512, num_quantizers=8, codebook_params=CodebookParams(dim=512, codebook_size=1024,
                                                                                 threshold_ema_dead_code=0)).to(DEV)
        mods.append(m)
    g = torch.Generator(device=DEV).manual_seed(5)
    for li in range(8):
        c = torch.randn(1, 1024, 512, generator=g, device=DEV) * (0.5 / 1.4 ** li)
        for m in mods:
            cb = m.layers[li]._codebook
            cb.embeddings.copy_(c); cb.embed_avg.copy_(c); cb.cluster_size.fill_(1.0); cb.invalidate_cache()
    x = torch.randn(8, 4096, 512, generator=g, device=DEV)
    mods[0].use_fused_levels, mods[1].use_fused_levels = True, False
    outs = []
    for m in mods:
        m.train()
        with torch.no_grad():
            outs.append(m(x))
    (qa, ia, la), (qb, ib, lb) = outs
    assert torch.equal(ia, ib), f"{int((ia != ib).sum())} index mismatches between fused and generic RVQ"
    assert torch.equal(qa, qb)
    assert torch.allclose(la, lb, rtol=1e-6)
    for a, b in zip(mods[0].layers, mods[1].layers):
        assert torch.equal(a._codebook.cluster_size, b._codebook.cluster_size)
        assert torch.equal(a._codebook.embeddings, b._codebook.embeddings)
