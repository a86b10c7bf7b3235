"""GPU tests of the dense-similarity consumers (csrc/dense.cu behind vqb200.VectorQuantize): cross-entropy to given
indices, cross-entropy commitment loss, codebook diversity loss -- forward and input gradient.

(a) every C entry point against a plain-torch fp64 statement of the same op (tests/dense_ref.py) on ragged shapes,
    all latent dtypes, both metrics;
(b) the module against the fixtures recorded from the live reference (tests/golden/dense/): quantize bit-exact,
    indices equal, losses within 1e-5 relative, the gradient of the recorded scalar within 1e-5 of its max;
(c) the gradient of each loss ALONE against torch autograd through the CPU restatement (the recorded scalar is
    dominated by the straight-through term).
"""
import pytest
import torch

import dense_ref as R
import golden_util as gu
from test_dense_host_cpu import build_module, load_state

pytestmark = pytest.mark.gpu


def _dev():
    return torch.device("cuda:0")


def _rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


SHAPES = [
    # H, N, K, d, n_pos, cosine, dtype
    (1, 200, 100, 32, 50, False, torch.float32),
    (2, 3 * 67, 130, 30, 67, False, torch.float32),      # d % 4 != 0: scalar loads; ragged everywhere
    (1, 1000, 513, 256, 125, True, torch.float32),
    (1, 700, 257, 300, 70, False, torch.bfloat16),       # d > 256: two output chunks in the backward pass
    (3, 128, 64, 512, 16, False, torch.float16),
    (1, 320, 1000, 64, 1, False, torch.float32),         # n_pos = 1 (2-D inputs)
    (1, 4099, 2048, 128, 4099, True, torch.bfloat16),
    (2, 1, 7, 8, 1, False, torch.float32),
]


@pytest.mark.parametrize("H,N,K,d,n_pos,cosine,dtype", SHAPES)
def test_dense_entry_points_match_fp64_statement(H, N, K, d, n_pos, cosine, dtype):
    from vqb200 import ops
    dev = _dev()
    g = torch.Generator().manual_seed(H * 1000 + N + K + d)
    x = torch.randn(H, N, d, generator=g).to(dtype).to(dev)
    c_dist = (torch.randn(H, K, d, generator=g) * 0.5).to(dev)
    c_comb = (c_dist + 0.1 * torch.randn(H, K, d, generator=g).to(dev)).contiguous()
    if cosine:
        x = torch.nn.functional.normalize(x.float(), dim=-1).to(dtype)
        c_dist = torch.nn.functional.normalize(c_dist, dim=-1)
    target = torch.randint(0, K, (H, N), generator=g).to(dev)
    target[:, ::7] = -1
    table = torch.randn(n_pos, K, generator=g).to(dev)
    coef = torch.rand(H, N, generator=g).to(dev)
    x64, cd64, cc64 = x.double(), c_dist.double(), c_comb.double()

    xn2 = None if cosine else ops.dense_row_norms(x)
    cn2 = None if cosine else ops.dense_row_norms(c_dist)
    if not cosine:
        assert _rel(xn2, (x64 * x64).sum(-1)) <= 1e-6 and _rel(cn2, (cd64 * cd64).sum(-1)) <= 1e-6

    for alpha in (1.0, -3.0):
        lse, st = ops.dense_rowstats(x, xn2, c_dist, cn2, cosine, alpha, target)
        lse_r, st_r = R.rowstats(x64, cd64, cosine, alpha, target)
        assert torch.allclose(lse.double(), lse_r, rtol=1e-5, atol=2e-5 * abs(alpha)), (alpha, _rel(lse, lse_r))
        assert torch.allclose(st.double(), st_r, rtol=1e-5, atol=2e-5)

        avg = ops.dense_avgprob(x, xn2, c_dist, cn2, cosine, alpha, lse, n_pos)
        avg_r = R.avgprob(x64, cd64, cosine, alpha, lse_r, n_pos)
        assert avg.shape == (n_pos, K)
        assert torch.allclose(avg.double(), avg_r, rtol=2e-4, atol=1e-8), _rel(avg, avg_r)

        rd = ops.dense_rowdot(x, xn2, c_dist, cn2, cosine, alpha, lse, table, n_pos)
        rd_r = R.rowdot(x64, cd64, cosine, alpha, lse_r, table.double(), n_pos)
        assert torch.allclose(rd.double(), rd_r, rtol=2e-4, atol=1e-4), _rel(rd, rd_r)

        g_ce = ops.dense_backward(x, xn2, c_dist, cn2, c_comb, cosine, alpha, lse, coef, target=target)
        g_ce_r = R.backward(x64, cd64, cc64, cosine, alpha, lse_r, coef.double(), target=target)
        assert g_ce.shape == (H, N, d) and bool(torch.isfinite(g_ce).all())
        assert _rel(g_ce, g_ce_r) <= 2e-4, _rel(g_ce, g_ce_r)

        gc_ce = ops.dense_backward_codes(x, xn2, c_dist, cn2, cosine, alpha, lse, coef, target=target)
        gc_ce_r = R.backward_codes(x64, cd64, cosine, alpha, lse_r, coef.double(), target=target)
        assert gc_ce.shape == (H, K, d) and _rel(gc_ce, gc_ce_r) <= 2e-4, _rel(gc_ce, gc_ce_r)
        gc_dv = ops.dense_backward_codes(x, xn2, c_dist, cn2, cosine, alpha, lse, coef, table=table, rdot=rd,
                                         n_pos=n_pos)
        gc_dv_r = R.backward_codes(x64, cd64, cosine, alpha, lse_r, coef.double(), table=table.double(), rdot=rd_r,
                                   n_pos=n_pos)
        assert _rel(gc_dv, gc_dv_r) <= 2e-4, _rel(gc_dv, gc_dv_r)

        g_dv = ops.dense_backward(x, xn2, c_dist, cn2, c_comb, cosine, alpha, lse, coef, table=table, rdot=rd,
                                  n_pos=n_pos)
        g_dv_r = R.backward(x64, cd64, cc64, cosine, alpha, lse_r, coef.double(), table=table.double(), rdot=rd_r,
                            n_pos=n_pos)
        assert bool(torch.isfinite(g_dv).all())
        assert _rel(g_dv, g_dv_r) <= 2e-4, _rel(g_dv, g_dv_r)


def test_dense_rowstats_is_repeatable_and_handles_empty_input():
    from vqb200 import ops
    dev = _dev()
    g = torch.Generator().manual_seed(5)
    x = torch.randn(1, 3000, 64, generator=g).to(dev)
    c = torch.randn(1, 700, 64, generator=g).to(dev)
    xn2, cn2 = ops.dense_row_norms(x), ops.dense_row_norms(c)
    a = ops.dense_rowstats(x, xn2, c, cn2, False, 1.0, None)[0]
    b = ops.dense_rowstats(x, xn2, c, cn2, False, 1.0, None)[0]
    assert torch.equal(a, b)                                # fixed reduction order: bitwise repeatable
    p = ops.dense_avgprob(x, xn2, c, cn2, False, -2.0, ops.dense_rowstats(x, xn2, c, cn2, False, -2.0, None)[0], 300)
    assert torch.allclose(p.sum(-1), torch.ones(300, device=dev), atol=1e-5)
    e = torch.empty(1, 0, 64, device=dev)
    lse, _ = ops.dense_rowstats(e, ops.dense_row_norms(e), c, cn2, False, 1.0, None)
    assert lse.shape == (1, 0)


@pytest.mark.parametrize("name", gu.dense_fixture_names())
def test_dense_consumers_match_reference_fixture(name):
    fx = gu.load_dense(name)
    cfg = fx["cfg"]
    dev = _dev()
    vq = build_module(cfg).to(dev)
    vq.train(cfg["training"])
    load_state(vq, fx)
    x = fx["x"].to(dev).requires_grad_(True)
    launches0 = __import__("vqb200").ops.launch_count()
    if cfg["kind"] == "indices":
        out = vq(x, indices=fx["targets"].to(dev))
        assert isinstance(out, tuple) and len(out) == 2
        q, ce = out
        (q.sum() * 0.01 + ce * 1.3 if cfg["training"] else ce * 1.3).backward()
        assert torch.allclose(ce.detach().cpu(), fx["ce"], rtol=1e-5)
    else:
        mask = fx["mask"].to(dev) if fx["mask"] is not None else None
        q, ind, loss, bd = vq(x, mask=mask, return_loss_breakdown=True)
        ((q * fx["w"].to(dev)).sum() + loss.sum() * 1.7).backward()
        assert torch.equal(ind.cpu(), fx["indices"])        # incl. the -1 the reference writes into masked positions
        assert loss.shape == fx["loss"].shape
        assert torch.allclose(loss.detach().cpu(), fx["loss"], rtol=1e-5, atol=1e-6)
        assert torch.allclose(bd.commitment.detach().cpu(), fx["commitment"], rtol=1e-5)
        assert torch.allclose(bd.codebook_diversity.detach().cpu(), fx["codebook_diversity"], rtol=1e-5)
    assert __import__("vqb200").ops.launch_count() > launches0
    if cfg.get("l2in"):
        # the input normalisation on the device and torch's on the CPU may differ in the last bit (as in the cosine
        # cases of test_gpu_parity.py); given x^ the quantized vectors are exact
        assert _rel(q.detach().cpu(), fx["quantize"]) <= 1e-6
    else:
        assert torch.equal(q.detach().cpu(), fx["quantize"])
    assert _rel(x.grad.cpu(), fx["grad_x"]) <= 1e-5
    if cfg["training"]:
        assert torch.equal(vq._codebook.cluster_size.cpu(), fx["after"]["cluster_size"])
    assert _rel(vq._codebook.embeddings.cpu(), fx["after"]["embeddings"]) <= 1e-5


@pytest.mark.parametrize("name", [n for n in gu.dense_fixture_names() if not n.startswith("ce_indices")])
def test_dense_loss_gradient_alone_matches_cpu_autograd(name):
    """d(loss part)/dx with nothing else in the scalar, against torch autograd through the CPU restatement (which is
    itself pinned to the reference's gradients, tests/test_oracle_golden.py)."""
    from oracle import vq_oracle as O
    fx = gu.load_dense(name)
    cfg = fx["cfg"]
    dev = _dev()
    for part in (["codebook_diversity"] if cfg.get("dw", 0) > 0 else []) + \
                (["commitment"] if cfg["kind"] == "commit" or cfg.get("ce_commit") else []):
        vq = build_module(cfg).to(dev).train()
        load_state(vq, fx)
        x = fx["x"].to(dev).requires_grad_(True)
        mask = fx["mask"].to(dev) if fx["mask"] is not None else None
        _, _, _, bd = vq(x, mask=mask, return_loss_breakdown=True)
        getattr(bd, part).backward()
        st = O.CodebookState(fx["init"]["embeddings"].clone(), fx["init"]["embed_avg"].clone(),
                             fx["init"]["cluster_size"].clone())
        xo = fx["x"].clone().requires_grad_(True)
        _, _, _, parts = O.vq_forward_dense(st, xo, gu.dense_oracle_opts(cfg), training=True, mask=fx["mask"],
                                            **gu.dense_kwargs(cfg))
        parts[part].backward()
        assert float(xo.grad.abs().max()) > 0
        tol = 1e-4 if part == "codebook_diversity" else 2e-5   # the diversity weights cancel (G - sum p G)
        assert _rel(x.grad.cpu(), xo.grad) <= tol, (part, _rel(x.grad.cpu(), xo.grad))


def test_dense_consumers_medium_shape_bf16_and_eval_indices():
    """A shape with several row / code tiles per block, bf16 latents, against the fp64 statement; and the
    `indices=` branch in eval mode leaves the codebook untouched."""
    from vqb200 import CodebookParams, VectorQuantize
    dev = _dev()
    g = torch.Generator().manual_seed(21)
    B, n, d, K = 6, 333, 128, 1500
    vq = VectorQuantize(dim=d, codebook_params=CodebookParams(dim=d, codebook_size=K, threshold_ema_dead_code=0),
                        commitment_use_cross_entropy_loss=True, codebook_diversity_loss_weight=0.5,
                        codebook_diversity_temperature=1.5).to(dev).train()
    c = (torch.randn(1, K, d, generator=g) * 0.5).to(dev)
    cb = vq._codebook
    with torch.no_grad():
        cb.embeddings.copy_(c); cb.embed_avg.copy_(c); cb.cluster_size.fill_(1.0)
    cb.invalidate_cache()
    x = torch.randn(B, n, d, generator=g).to(torch.bfloat16).to(dev)
    q, ind, loss, bd = vq(x, return_loss_breakdown=True)
    x64, c64 = x.double().reshape(1, -1, d), c.double()
    s = R.sims(x64, c64, False)
    ce_r = torch.nn.functional.cross_entropy(s[0], ind.reshape(-1))
    prob = (-s * 1.5).softmax(-1).reshape(B, n, K).mean(0)
    dv_r = (prob * prob.clamp(min=1e-5).log()).sum(-1).mean()
    assert torch.allclose(bd.commitment.double(), ce_r, rtol=1e-5)
    assert torch.allclose(bd.codebook_diversity.double(), dv_r, rtol=1e-5)
    assert q.dtype == torch.float32 and ind.dtype == torch.int64 and loss.shape == (1,)

    vq.eval()
    before = cb.embeddings.clone()
    tgt = torch.randint(0, K, (B, n), generator=g).to(dev)
    q2, ce2 = vq(x, indices=tgt)
    s2 = R.sims(x64, cb.embeddings.double(), False)
    assert torch.allclose(ce2.double(), torch.nn.functional.cross_entropy(s2[0], tgt.reshape(-1)), rtol=1e-5)
    assert torch.equal(cb.embeddings, before)


def test_residual_vq_levels_carry_the_dense_losses():
    """ResidualVQ(commitment_use_cross_entropy_loss=True, codebook_diversity_loss_weight>0): the fused level loop is
    bypassed and every level's loss holds the CE commitment + diversity terms, as in the reference's loop over
    VectorQuantize (residual_vq.py:212-243); checked level by level against the CPU restatement."""
    from oracle import vq_oracle as O
    from vqb200 import CodebookParams, ResidualVQ
    dev = _dev()
    g = torch.Generator().manual_seed(9)
    B, n, d, K, Q = 3, 40, 32, 50, 3
    rvq = ResidualVQ(dim=d, num_quantizers=Q, codebook_params=CodebookParams(dim=d, codebook_size=K,
                                                                           threshold_ema_dead_code=0),
                     commitment_use_cross_entropy_loss=True, commitment_weight=0.6, codebook_diversity_loss_weight=0.4,
                     codebook_diversity_temperature=2.0, sync_codebook=False).to(dev).train()
    states = []
    for layer in rvq.layers:
        c = torch.randn(1, K, d, generator=g) * 0.6
        cb = layer._codebook
        with torch.no_grad():
            cb.embeddings.copy_(c); cb.embed_avg.copy_(c); cb.cluster_size.fill_(1.0)
        cb.invalidate_cache()
        states.append(O.CodebookState(c.clone(), c.clone(), torch.ones(1, K)))
    x = torch.randn(B, n, d, generator=g)
    out, ind, losses = rvq(x.to(dev))
    assert out.shape == (B, n, d) and ind.shape == (B, n, Q) and losses.shape == (1, Q)
    opts = O.VQOpts(commitment_weight=0.6, codebook=O.CodebookOpts(threshold_ema_dead_code=0))
    residual, total = x, 0.0
    for qi in range(Q):
        q, i, l, _ = O.vq_forward_dense(states[qi], residual, opts, training=True, ce_commit=True,
                                        diversity_weight=0.4, diversity_temperature=2.0)
        assert torch.equal(ind[..., qi].cpu(), i), f"level {qi}: indices"
        assert torch.allclose(losses[:, qi].cpu(), l, rtol=1e-5), f"level {qi}: loss {losses[:, qi]} vs {l}"
        residual = residual - q.detach()
        total = total + q.detach()
    assert _rel(out.detach().cpu(), total) <= 1e-6


@pytest.mark.parametrize("name", ["learnable_ce_commit", "learnable_ce_commit_dot", "learnable_diversity",
                                  "learnable_ce_indices"])
def test_learnable_codebook_with_dense_losses_matches_reference_fixture(name):
    """learnable codebook + CE commitment / diversity loss / CE to indices: outputs and the gradients with respect to
    the input AND the codebook (through the similarities) against the live reference
    (tests/golden/make_golden_learnable.py)."""
    import os
    from vqb200 import CodebookParams, VectorQuantize
    dev = _dev()
    fx = torch.load(os.path.join(gu.GOLDEN_DIR, "learnable", name + ".pt"), weights_only=False)
    cfg = fx["cfg"]
    cp = CodebookParams(dim=cfg["dim"], codebook_size=cfg["K"], learnable_codebook=True, ema_update=False,
                        threshold_ema_dead_code=0, use_cosine_sim=cfg.get("cosine", False))
    extra = {}
    if cfg.get("ce"):
        extra["commitment_use_cross_entropy_loss"] = True
    if cfg.get("dw"):
        extra.update(codebook_diversity_loss_weight=cfg["dw"], codebook_diversity_temperature=cfg["temp"])
    vq = VectorQuantize(dim=cfg["dim"], codebook_params=cp, commitment_weight=cfg["cw"], sync_update_v=cfg["v"],
                        sync_codebook=False, **extra).to(dev).train()
    with torch.no_grad():
        vq._codebook.embeddings.copy_(fx["init_embeddings"])
    vq._codebook.invalidate_cache()
    x = fx["x"].to(dev).requires_grad_(True)
    if cfg.get("indices"):
        q, ce = vq(x, indices=fx["targets"].to(dev))
        (q.sum() * 0.01 + ce * 1.3).backward()
        assert torch.allclose(ce.detach().cpu(), fx["ce"], rtol=1e-5)
    else:
        mask = fx["mask"].to(dev) if fx["mask"] is not None else None
        q, ind, loss = vq(x, mask=mask)
        (q * fx["w"].to(dev)).sum().add(loss.sum() * 1.7).backward()
        assert torch.equal(ind.cpu(), fx["indices"])
        assert torch.allclose(loss.detach().cpu(), fx["loss"], rtol=1e-5)
    assert torch.equal(q.detach().cpu(), fx["quantize"])
    assert _rel(x.grad.cpu(), fx["grad_x"]) <= 1e-5
    assert _rel(vq._codebook.embeddings.grad.cpu(), fx["grad_embeddings"]) <= 2e-5
    assert torch.equal(vq._codebook.embeddings.detach().cpu(), fx["init_embeddings"])
