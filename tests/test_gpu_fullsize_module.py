"""Full-size parity of the MODULE path against the row-chunked CPU oracle (VERDICT r01 "next" 1b).

Unlike tests/test_gpu_fullsize.py (op level, 2048 sampled rows, properties), these drive `VectorQuantize.forward` /
`ResidualVQ.forward` at BASELINE.json's shapes and compare EVERY index and every EMA buffer with
`oracle.vq_oracle` run on the host with `row_chunk` (the reference materialises N x K; chunking rows is bitwise neutral
for cdist / einsum):

  C2  VectorQuantize(256, K=8192), 2^20 bf16 latents, one training step          (codebooks.py:350-435)
  C3  cosine + l2norm, K=16384, d=512, 2^19 latents, THREE training steps with threshold_ema_dead_code=2 and codes that
      really die (duplicates + far-away codes), so the expiry of evolved statistics is compared too   (:230-255)
  C4  ResidualVQ 8 x K=1024 x d=512 on an 8 x 4096 slice of the 64 x 4096 batch, one training step (residual_vq.py:212-243)

Bars (north star): indices identical except rows whose reference fp32 top-2 gap is < 1e-6; quantize bit-exact given
the indices and identical codebooks; loss / embed_avg / embeddings <= 1e-5 relative; cluster_size exact.
After each step the oracle's buffers are copied into the module ("teacher forcing"), so every step is compared on
identical state and an exempt tie in one step cannot blur the next.

The oracle needs tens of seconds of host CPU per step at these sizes: the tests are marked `slow_oracle` as well and
can be deselected with `-m "gpu and not slow_oracle"` while iterating.
"""
import os

import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.slow_oracle]
DEV = "cuda:0"
REL = 1e-5


def _cpu_draw(num_rows, m, device):
    if num_rows >= m:
        return torch.randperm(num_rows)[:m].to(device)
    return torch.randint(0, num_rows, (m,)).to(device)


@pytest.fixture(autouse=True)
def _patch_draw(monkeypatch):
    from vqb200 import codebook
    monkeypatch.setattr(codebook.Codebook, "_draw_rows", staticmethod(_cpu_draw))
    torch.set_num_threads(os.cpu_count() or 1)


def _rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def _exempt_only(flat_x, emb, cos, got, ref, what):
    """Every row where `got != ref` must lie inside the reference's own fp32 1e-6 top-2 window.  Returns the flipped
    row ids.  flat_x (N,d) fp32 CPU = what the reference's search saw, emb (K,d)."""
    from oracle import vq_oracle as O
    flips = (got != ref).nonzero().flatten()
    if flips.numel() == 0:
        return flips
    sim = O.similarities(flat_x[flips][None], emb[None], cos)[0]
    top2 = sim.topk(2, -1).values
    gap = (top2[:, 0] - top2[:, 1]).abs() / top2[:, 0].abs().clamp_min(1e-30)
    bad = gap >= 1e-6
    assert not bool(bad.any()), (
        f"{what}: {int(bad.sum())} of {got.numel()} rows differ from the oracle OUTSIDE its 1e-6 window; "
        f"(row, got, oracle, oracle gap): "
        f"{[(int(flips[i]), int(got[flips[i]]), int(ref[flips[i]]), float(gap[i])) for i in bad.nonzero().flatten()[:8]]}")
    return flips


def _compare_buffers(cb, st, flips_codes, what):
    """cluster_size exact, embed_avg / embeddings <= 1e-5 -- on the codes no exempt tie touched (all of them when
    there was none)."""
    cs, ea, em = cb.cluster_size.cpu(), cb.embed_avg.cpu(), cb.embeddings.detach().cpu()
    keep = torch.ones(cs.shape[-1], dtype=torch.bool)
    if flips_codes.numel():
        keep[flips_codes] = False
    assert torch.equal(cs[0][keep], st.cluster_size[0][keep]), \
        f"{what}: cluster_size differs on {int((cs[0][keep] != st.cluster_size[0][keep]).sum())} codes"
    assert _rel(ea[0][keep], st.embed_avg[0][keep]) <= REL, f"{what}: embed_avg {_rel(ea[0][keep], st.embed_avg[0][keep])}"
    assert _rel(em[0][keep], st.embeddings[0][keep]) <= REL, f"{what}: embeddings {_rel(em[0][keep], st.embeddings[0][keep])}"


def _sync_module_to_oracle(cb, st):
    with torch.no_grad():
        cb.embeddings.copy_(st.embeddings.to(cb.embeddings.device))
        cb.embed_avg.copy_(st.embed_avg.to(cb.embed_avg.device))
        cb.cluster_size.copy_(st.cluster_size.to(cb.cluster_size.device))
    cb.invalidate_cache()


def test_c2_vector_quantize_training_step_matches_chunked_oracle():
    from oracle import vq_oracle as O
    from vqb200 import CodebookParams, VectorQuantize
    K, d, shape = 8192, 256, (1024, 1024, 256)
    torch.manual_seed(0)
    vq = VectorQuantize(dim=d, codebook_params=CodebookParams(dim=d, codebook_size=K, threshold_ema_dead_code=2)).to(DEV)
    g = torch.Generator().manual_seed(21)
    c = torch.randn(1, K, d, generator=g) * 0.5
    c[0, 4000:4008] = c[0, 17]                  # duplicates: lowest index wins every row, the copies die and expire
    c[0, 5000:5004] *= 40.0                     # far-away codes: never chosen
    cs0 = torch.ones(1, K)
    x = torch.randn(shape, generator=g).bfloat16()
    cb = vq._codebook
    st = O.CodebookState(c.clone(), c.clone(), cs0.clone())
    _sync_module_to_oracle(cb, st)
    vq.train()
    torch.manual_seed(77)
    with torch.no_grad():
        q, ind, loss = vq(x.to(DEV))
    torch.cuda.synchronize()
    torch.manual_seed(77)
    opts = O.VQOpts(codebook=O.CodebookOpts(threshold_ema_dead_code=2))
    qo, io, lo, _ = O.vq_forward(st, x, opts, training=True, row_chunk=16384)
    got, ref = ind.cpu().reshape(-1), io.reshape(-1)
    flips = _exempt_only(x.reshape(-1, d).float(), c[0], False, got, ref, "C2")
    same = got == ref
    qc = q.cpu().reshape(-1, d)
    assert torch.equal(qc[same], qo.reshape(-1, d)[same]), "C2: quantize not bit-exact given the indices"
    assert abs(float(loss) - float(lo)) <= REL * abs(float(lo)), (float(loss), float(lo))
    touched = torch.cat([got[flips], ref[flips]]).unique()
    _compare_buffers(cb, st, touched, "C2")
    # the dead codes (duplicates, far-away codes) were replaced by the same batch rows
    assert torch.equal(cb.cluster_size.cpu()[0, 4000:4008], torch.full((8,), 2.0))
    print(f"C2 module parity: {int(flips.numel())} exempt ties of {got.numel()} rows")


def test_c3_cosine_three_training_steps_with_expiry_match_chunked_oracle():
    from oracle import vq_oracle as O
    from vqb200 import CodebookParams, VectorQuantize
    K, d, shape = 16384, 512, (512, 1024, 512)
    torch.manual_seed(0)
    cp = CodebookParams(dim=d, codebook_size=K, threshold_ema_dead_code=2, use_cosine_sim=True,
                        transform_input="l2norm", weights_regularization="l2norm")
    vq = VectorQuantize(dim=d, codebook_params=cp).to(DEV)
    g = torch.Generator().manual_seed(31)
    c = torch.nn.functional.normalize(torch.randn(1, K, d, generator=g), dim=-1)
    c[0, 9000:9016] = c[0, 5]                   # duplicates die (the lowest index takes their rows)
    cs0 = torch.ones(1, K)
    cb = vq._codebook
    st = O.CodebookState(c.clone(), c.clone(), cs0.clone())
    _sync_module_to_oracle(cb, st)
    vq.train()
    opts = O.VQOpts(input_l2norm=True, codebook=O.CodebookOpts(threshold_ema_dead_code=2, use_cosine_sim=True,
                                                               weights_l2norm=True))
    x0 = torch.randn(shape, generator=g)
    dirs = torch.nn.functional.normalize(torch.randn(3, d, generator=g), dim=-1)
    fired = []
    for step in range(3):
        # non-stationary batch: every step the latents lean towards another direction (cos ~ 0.7), so the codes on the
        # far side starve and die, are replaced by batch rows, and those replacements starve in the NEXT step's batch:
        # expiry fires on evolved statistics in all three steps
        x = x0 + (d ** 0.5) * dirs[step]
        emb_before = st.embeddings.clone()
        torch.manual_seed(100 + step)
        with torch.no_grad():
            q, ind, loss = vq(x.to(DEV))
        torch.cuda.synchronize()
        torch.manual_seed(100 + step)
        cs_pre = st.cluster_size.clone()
        qo, io, lo, _ = O.vq_forward(st, x, opts, training=True, row_chunk=8192)
        got, ref = ind.cpu().reshape(-1), io.reshape(-1)
        xn = O.l2norm(x.reshape(-1, d))
        flips = _exempt_only(xn, emb_before[0], True, got, ref, f"C3 step {step}")
        same = got == ref
        assert _rel(q.cpu().reshape(-1, d)[same], qo.reshape(-1, d)[same]) <= REL, f"C3 step {step}: quantize"
        assert abs(float(loss) - float(lo)) <= REL * abs(float(lo)), (step, float(loss), float(lo))
        touched = torch.cat([got[flips], ref[flips]]).unique()
        _compare_buffers(cb, st, touched, f"C3 step {step}")
        # how many codes the reference expired in this step: cluster_size == reset exactly and changed embeddings
        expired = int(((st.cluster_size[0] == 2.0) & (0.8 * cs_pre[0] < 2.0)).sum())
        fired.append(expired)
        print(f"C3 step {step}: {int(flips.numel())} exempt ties, ~{expired} codes expired, loss {float(lo):.6f}")
        _sync_module_to_oracle(cb, st)
    assert all(n > 0 for n in fired), f"the C3 scenario was meant to make dead-code expiry fire in every step: {fired}"


def test_c4_residual_vq_slice_matches_oracle():
    from oracle import vq_oracle as O
    from vqb200 import CodebookParams, ResidualVQ
    Q, K, d = 8, 1024, 512
    torch.manual_seed(0)
    rvq = ResidualVQ(dim=d, num_quantizers=Q,
                     codebook_params=CodebookParams(dim=d, codebook_size=K, threshold_ema_dead_code=2)).to(DEV)
    g = torch.Generator().manual_seed(41)
    states = []
    for li, layer in enumerate(rvq.layers):
        c = torch.randn(1, K, d, generator=g) * (0.5 / 1.4 ** li)
        c[0, 700:704] = c[0, 3]
        st = O.CodebookState(c.clone(), c.clone(), torch.ones(1, K))
        _sync_module_to_oracle(layer._codebook, st)
        states.append(st)
    x = torch.randn(8, 4096, d, generator=g)
    pre = [s.embeddings.clone() for s in states]
    opts = O.VQOpts(codebook=O.CodebookOpts(threshold_ema_dead_code=2))
    for fused in (True, False):
        rvq.use_fused_levels = fused
        sts = [O.CodebookState(e.clone(), e.clone(), torch.ones(1, K)) for e in pre]
        for layer, st in zip(rvq.layers, sts):
            _sync_module_to_oracle(layer._codebook, st)
        rvq.train()
        torch.manual_seed(500)
        with torch.no_grad():
            q, ind, losses = rvq(x.to(DEV))
        torch.cuda.synchronize()
        torch.manual_seed(500)
        qo, io, lo, extras = O.rvq_forward(sts, x, opts, training=True, want_gap=True)
        got, ref = ind.cpu().reshape(-1, Q), io.reshape(-1, Q)
        alive = torch.ones(got.shape[0], dtype=torch.bool)       # rows whose residual is still identical
        n_flips = 0
        for li in range(Q):
            diff = (got[:, li] != ref[:, li]) & alive
            gap = extras[li]["top2_rel_gap"].reshape(-1)
            assert bool((gap[diff] < 1e-6).all()), \
                f"C4 level {li} (fused={fused}): {int((gap[diff] >= 1e-6).sum())} index mismatches outside the 1e-6 window"
            n_flips += int(diff.sum())
            alive &= ~diff
        assert torch.equal(q.cpu().reshape(-1, d)[alive], qo.reshape(-1, d)[alive]), \
            f"C4 (fused={fused}): quantized_out not bit-exact"
        if n_flips == 0:
            assert torch.allclose(losses.cpu(), lo, rtol=REL), (losses, lo)
            for li, (layer, st) in enumerate(zip(rvq.layers, sts)):
                _compare_buffers(layer._codebook, st, torch.empty(0, dtype=torch.long), f"C4 level {li} fused={fused}")
        print(f"C4 slice fused={fused}: {n_flips} exempt ties of {got.numel()} lookups")


def test_c4_slice_fused_backward_equals_per_level_autograd():
    """`_FusedRVQ` (fused level loop forward, ONE replay pass backward: vqb_rvq_backward) against the generic per-level
    loop under torch autograd at the C4 slice (8 x 1024 x 512): same outputs and losses, input gradient to fp32
    rounding.  The per-level loop itself is pinned to the reference's gradients by the `grads/` fixtures."""
    from vqb200 import CodebookParams, ResidualVQ
    Q, K, d = 8, 1024, 512
    g = torch.Generator().manual_seed(43)
    cbs = [torch.randn(1, K, d, generator=g) * (0.5 / 1.4 ** li) for li in range(Q)]
    x0 = torch.randn(4, 4096, d, generator=g)
    w = torch.randn(4, 4096, d, generator=g)
    res = {}
    for fused in (True, False):
        torch.manual_seed(0)
        rvq = ResidualVQ(dim=d, num_quantizers=Q,
                         codebook_params=CodebookParams(dim=d, codebook_size=K, threshold_ema_dead_code=0)).to(DEV).train()
        for layer, c in zip(rvq.layers, cbs):
            cb = layer._codebook
            with torch.no_grad():
                cb.embeddings.copy_(c); cb.embed_avg.copy_(c); cb.cluster_size.fill_(1.0)
            cb.invalidate_cache()
        rvq.use_fused_levels = fused
        x = x0.to(DEV).requires_grad_(True)
        q, ind, losses = rvq(x)
        ((q * w.to(DEV)).sum() + (losses * torch.arange(1, Q + 1, device=DEV)).sum() * 3.0).backward()
        res[fused] = (q.detach(), ind, losses.detach(), x.grad.clone(),
                      [l._codebook.embeddings.clone() for l in rvq.layers])
    qf, indf, lf, gf, ef = res[True]
    qg, indg, lg, gg, eg = res[False]
    assert torch.equal(indf, indg)
    assert torch.equal(qf, qg), "quantized_out differs between the fused and the per-level loop"
    assert torch.allclose(lf, lg, rtol=1e-6)
    assert _rel(gf.cpu(), gg.cpu()) <= 1e-6, _rel(gf.cpu(), gg.cpu())
    for a, b in zip(ef, eg):
        assert _rel(a.cpu(), b.cpu()) <= 1e-6
