"""Pins the CPU oracle against fixtures produced by the live reference (tests/golden/make_golden.py).

No GPU.  Bit-exact: indices, quantize, loss, cluster_size.  embed_avg/embeddings: 1e-6 relative
(the oracle accumulates embed_sum with the same einsum, so these are normally bit-exact too).
"""
import pytest
import torch

import golden_util as gu
from oracle import vq_oracle as O


@pytest.mark.parametrize("name", gu.fixture_names())
def test_oracle_matches_reference_fixture(name):
    fx = gu.load(name)
    cfg = fx["cfg"]
    opts = gu.oracle_opts(cfg)
    states = gu.oracle_states(fx)
    training = cfg.get("training", True)
    for s, step in enumerate(fx["steps"]):
        x = fx["x"] + 0.01 * s
        torch.manual_seed(fx["rng_seed"] + s)
        if cfg["kind"] == "vq":
            q, ind, loss, ex = O.vq_forward(states[0], x, opts, training=training, mask=fx["mask"], want_gap=True)
            gaps = [ex["top2_rel_gap"]]
        else:
            q, ind, loss, exs = O.rvq_forward(states, x, opts, training=training, mask=fx["mask"], want_gap=True)
            gaps = [e["top2_rel_gap"] for e in exs]
        assert torch.equal(ind, step["indices"]), f"{name} step {s}: indices differ"
        assert torch.equal(q, step["quantize"]), f"{name} step {s}: quantize not bit-exact"
        assert q.dtype == step["quantize"].dtype and ind.dtype == torch.int64
        assert loss.shape == step["loss"].shape
        assert torch.equal(loss, step["loss"]), f"{name} step {s}: loss {loss} vs {step['loss']}"
        for lvl, after in enumerate(step["after"]):
            st = states[lvl]
            assert torch.equal(st.cluster_size, after["cluster_size"]), f"{name} step {s} lvl {lvl} cluster_size"
            assert gu.rel_err(st.embed_avg, after["embed_avg"]) <= 1e-6
            assert gu.rel_err(st.embeddings, after["embeddings"]) <= 1e-6
        for g_or, g_ref in zip(gaps, step["top2_rel_gap"]):
            assert torch.allclose(g_or.reshape(-1), g_ref.reshape(-1), rtol=1e-4, atol=1e-9)


def test_row_chunked_oracle_equals_unchunked():
    """The big configs need a row-chunked oracle: chunking must not change indices/quantize."""
    fx = gu.load("c1_noexpire")
    opts = gu.oracle_opts(fx["cfg"])
    a, b = gu.oracle_states(fx)[0], gu.oracle_states(fx)[0]
    qa, ia, la, _ = O.vq_forward(a, fx["x"], opts)
    qb, ib, lb, _ = O.vq_forward(b, fx["x"], opts, row_chunk=100)
    assert torch.equal(ia, ib) and torch.equal(qa, qb) and torch.equal(la, lb)
    assert torch.equal(a.cluster_size, b.cluster_size)
    assert gu.rel_err(b.embed_avg, a.embed_avg) <= 1e-6
    assert gu.rel_err(b.embeddings, a.embeddings) <= 1e-6


LEARNABLE = ["learnable_plain", "learnable_masked_v", "inplace_sgd", "inplace_sgd_masked", "learnable_ce_commit",
             "learnable_ce_commit_dot", "learnable_diversity", "learnable_ce_indices"]


@pytest.mark.parametrize("name", LEARNABLE)
def test_oracle_learnable_codebook_matches_reference_fixture(name):
    """The learnable-codebook restatement reproduces the live reference's outputs AND its gradients with respect to
    the input and the codebook (fixtures: tests/golden/make_golden_learnable.py)."""
    import os
    from oracle import vq_oracle as O
    fx = torch.load(os.path.join(gu.GOLDEN_DIR, "learnable", name + ".pt"), weights_only=False)
    cfg = fx["cfg"]
    emb = fx["init_embeddings"].clone().requires_grad_(True)
    x = fx["x"].clone().requires_grad_(True)
    dense = dict(use_cosine_sim=cfg.get("cosine", False), ce_commit=bool(cfg.get("ce")),
                 diversity_weight=cfg.get("dw", 0.0), diversity_temperature=cfg.get("temp", 100.0))
    if cfg.get("indices"):   # forward(x, indices=...) -> (quantize, ce): gradients to the input AND the codebook
        q, ce = O.vq_forward_learnable(emb, x, targets=fx["targets"], **dense)
        (q.sum() * 0.01 + ce * 1.3).backward()
        assert torch.equal(q.detach(), fx["quantize"]) and torch.equal(ce.detach(), fx["ce"])
        assert torch.allclose(x.grad, fx["grad_x"], rtol=1e-6, atol=1e-9)
        assert torch.allclose(emb.grad, fx["grad_embeddings"], rtol=1e-6, atol=1e-9)
        return
    if "lr" in cfg:      # in_place_codebook_optimizer: an SGD step on the codebook inside the forward
        q, ind, loss, inplace = O.vq_forward_learnable(emb, x, commitment_weight=cfg["cw"], sync_update_v=cfg["v"],
                                                       mask=fx["mask"],
                                                       inplace_optimizer=torch.optim.SGD([emb], lr=cfg["lr"]))
        assert torch.equal(inplace, fx["inplace_optimize"])
        assert torch.equal(emb.detach(), fx["after_embeddings"])
        assert not torch.equal(emb.detach(), fx["init_embeddings"])
    else:
        q, ind, loss = O.vq_forward_learnable(emb, x, commitment_weight=cfg["cw"], sync_update_v=cfg["v"],
                                              mask=fx["mask"], **dense)
    (q * fx["w"]).sum().add(loss.sum() * 1.7).backward()
    assert torch.equal(ind, fx["indices"])
    assert torch.equal(q.detach(), fx["quantize"])
    assert torch.equal(loss.detach(), fx["loss"])
    assert torch.allclose(x.grad, fx["grad_x"], rtol=1e-6, atol=1e-9)
    assert torch.allclose(emb.grad, fx["grad_embeddings"], rtol=1e-6, atol=1e-9)


@pytest.mark.parametrize("name", gu.dense_fixture_names())
def test_oracle_dense_consumers_match_reference_fixture(name):
    """Cross-entropy to given indices, CE commitment and the diversity loss (the consumers of the dense N x K
    similarities): outputs, post-step buffers and the input gradient of the live reference, bit for bit
    (fixtures: tests/golden/make_golden_dense.py).  The gradient check pins the reference's stale-codebook behaviour:
    pre-update distances combined with post-update code vectors."""
    fx = gu.load_dense(name)
    cfg = fx["cfg"]
    opts = gu.dense_oracle_opts(cfg)
    st = O.CodebookState(fx["init"]["embeddings"].clone(), fx["init"]["embed_avg"].clone(),
                         fx["init"]["cluster_size"].clone())
    x = fx["x"].clone().requires_grad_(True)
    if cfg["kind"] == "indices":
        q, ce = O.vq_forward_dense(st, x, opts, training=cfg["training"], targets=fx["targets"])
        (q.sum() * 0.01 + ce * 1.3 if cfg["training"] else ce * 1.3).backward()
        assert torch.equal(ce.detach(), fx["ce"])
    else:
        q, ind, loss, parts = O.vq_forward_dense(st, x, opts, training=True, mask=fx["mask"], **gu.dense_kwargs(cfg))
        ((q * fx["w"]).sum() + loss.sum() * 1.7).backward()
        assert torch.equal(ind, fx["indices"])
        assert torch.equal(loss.detach(), fx["loss"])
        assert torch.equal(parts["commitment"].detach(), fx["commitment"])
        assert torch.equal(parts["codebook_diversity"].detach(), fx["codebook_diversity"])
    assert torch.equal(q.detach(), fx["quantize"])
    assert torch.allclose(x.grad, fx["grad_x"], rtol=1e-6, atol=1e-9)
    assert torch.equal(st.cluster_size, fx["after"]["cluster_size"])
    assert gu.rel_err(st.embeddings, fx["after"]["embeddings"]) <= 1e-6


@pytest.mark.parametrize("name", gu.grad_fixture_names())
def test_oracle_input_gradient_on_the_ema_path_matches_reference_fixture(name):
    """One training forward under autograd with the EMA step moving the codebook inside it (fixtures:
    tests/golden/make_golden_grads.py): outputs, buffers and the input gradient of the live reference."""
    fx = gu.load_grad(name)
    cfg = fx["cfg"]
    opts = gu.oracle_opts(cfg)
    states = gu.oracle_states(fx)
    x = fx["x"].clone().requires_grad_(True)
    if cfg["kind"] == "vq":
        q, ind, loss, _ = O.vq_forward_dense(states[0], x, opts, training=True, mask=fx["mask"])
    else:
        q, ind, loss = O.rvq_forward_autograd(states, x, opts, mask=fx["mask"])
    ((q * fx["w"]).sum() + loss.sum() * 1.7).backward()
    assert torch.equal(ind, fx["indices"])
    assert torch.equal(q.detach(), fx["quantize"])
    assert torch.equal(loss.detach(), fx["loss"])
    assert torch.allclose(x.grad, fx["grad_x"], rtol=1e-6, atol=1e-9)
    for st, after in zip(states, fx["after"]):
        assert torch.equal(st.cluster_size, after["cluster_size"])
        assert gu.rel_err(st.embeddings, after["embeddings"]) <= 1e-6
