"""Pins the oracle's restatement of the f4 variants (oracle/vq_oracle.py: gumbel_sample, AffineState,
codebook_forward_variants, vq_forward_variants, orthogonal_loss) against the outputs, losses, gradients and buffers
recorded from the live reference (tests/golden/f4, written by tests/golden/make_golden_f4.py)."""
import pytest
import torch

import f4_util as F
import golden_util as gu
from oracle import vq_oracle as O


def _state(init, learnable):
    st = O.CodebookState(init["embeddings"].clone(), init["embed_avg"].clone(), init["cluster_size"].clone())
    if learnable:
        st.embeddings.requires_grad_(True)
    return st


@pytest.mark.parametrize("name", [n for n in F.names() if n != "g_rvq_stochastic_shared"])
def test_oracle_variants_match_reference(name):
    fx = F.load(name)
    cfg = fx["cfg"]
    orth = cfg.get("orth", {})
    param = bool(cfg.get("learnable")) or bool(orth)
    st = _state(fx["init"][0], param)
    copts = O.CodebookOpts(threshold_ema_dead_code=cfg["thr"], use_cosine_sim=cfg.get("cosine", False),
                           weights_l2norm=bool(cfg.get("l2")), ema_update=not cfg.get("learnable", False))
    aff = O.AffineState(**cfg["affine"]) if cfg.get("affine") else None
    for step in fx["steps"]:
        x = step["x"].clone()
        if cfg.get("grads"):
            x.requires_grad_(True)
        u = step["draws"][0] if step["draws"] else None
        torch.manual_seed(step["rng_seed"])
        if u is not None:
            torch.zeros_like(u).uniform_(0, 1)          # the reference's draw precedes the expiry's on the generator
        if cfg["kind"] == "cb":
            q, ind, sim = O.codebook_forward_variants(st, x, copts, gumbel=cfg.get("gumbel"), affine=aff,
                                                      training=step["training"], mask=step["mask"], uniforms=u)
            assert torch.equal(sim.detach(), step["similarities"])
            scalar = (q * step["w"]).sum()
        else:
            vo = O.VQOpts(heads=cfg.get("heads", 1), commitment_weight=cfg.get("cw", 1.0), input_l2norm=bool(cfg.get("l2")),
                          codebook=copts)
            if step["perms"]:
                perm = step["perms"][0]
                real = torch.randperm
                torch.randperm = lambda n, *a, **k: perm
            try:
                q, ind, loss, (commit, orthl) = O.vq_forward_variants(
                    st, x, vo, gumbel=cfg.get("gumbel"), affine=aff, training=step["training"], mask=step["mask"],
                    learnable=bool(cfg.get("learnable")), codebook_is_parameter=param, uniforms=u, **orth)
            finally:
                if step["perms"]:
                    torch.randperm = real
            assert torch.allclose(loss.detach(), step["loss"], rtol=1e-6, atol=1e-8)
            assert torch.allclose(orthl.detach(), step["breakdown"][2], rtol=1e-6, atol=1e-8)
            scalar = (q * step["w"]).sum() + loss.sum() * 1.7
        assert torch.equal(ind, step["indices"])
        assert torch.equal(q.detach(), step["quantize"])
        if cfg.get("grads"):
            scalar.backward()
            assert gu.rel_err(x.grad, step["grad_x"]) <= 1e-6
            if step.get("grad_embeddings") is not None:
                assert gu.rel_err(st.embeddings.grad, step["grad_embeddings"]) <= 1e-6
                st.embeddings.grad = None
        after = step["after"][0]
        assert torch.equal(st.cluster_size, after["cluster_size"])
        assert gu.rel_err(st.embeddings.detach(), after["embeddings"]) <= 1e-6
        assert gu.rel_err(st.embed_avg, after["embed_avg"]) <= 1e-6
        if aff is not None:
            for k in ("batch_mean", "batch_variance", "codebook_mean", "codebook_variance"):
                assert gu.rel_err(getattr(aff, k), after[k]) <= 1e-6, k


def test_oracle_rvq_stochastic_shared_matches_reference():
    """The reference's own test configuration (tests/test_residual_vq.py:39-73): three levels over ONE stochastic
    codebook; residual_vq.py:212-243 around the variant forward."""
    fx = F.load("g_rvq_stochastic_shared")
    cfg, step = fx["cfg"], fx["steps"][0]
    st = _state(fx["init"][0], False)
    copts = O.CodebookOpts(threshold_ema_dead_code=cfg["thr"])
    vo = O.VQOpts(codebook=copts)
    torch.manual_seed(step["rng_seed"])
    residual, out, inds, losses = step["x"].clone(), 0.0, [], []
    for u in step["draws"]:
        torch.zeros_like(u).uniform_(0, 1)
        q, ind, loss, _ = O.vq_forward_variants(st, residual, vo, gumbel=cfg["gumbel"], uniforms=u)
        residual = residual - q.detach()
        out = out + q
        inds.append(ind); losses.append(loss)
    assert torch.equal(torch.stack(inds, -1), step["indices"])
    assert torch.equal(out, step["quantize"])
    assert torch.allclose(torch.stack(losses, -1), step["loss"], rtol=1e-6)
    assert gu.rel_err(st.embeddings, step["after"][0]["embeddings"]) <= 1e-6
