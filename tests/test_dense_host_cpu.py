"""Host logic of the dense-similarity consumers (vqb200/vq.py, vqb200/ops.py: target layouts for one / shared /
separate heads, masks and the -1 fill of the returned indices, the pre-/post-update codebook pair, autograd wiring)
on CPU: the five C entry points are replaced by their plain-torch statements (tests/dense_ref.py) and the search /
gather / EMA step by the CPU restatement, and the module is compared with the live-reference fixtures.  The kernels
themselves are checked on the GPU (tests/test_gpu_dense.py)."""
import pytest
import torch

import dense_ref as R
import golden_util as gu
from oracle import vq_oracle as O


def _patch(monkeypatch):
    from vqb200 import ops
    from vqb200.codebook import Codebook
    f = lambda t: t.detach().float()  # noqa: E731
    monkeypatch.setattr(ops, "dense_row_norms", lambda t: (f(t) * f(t)).sum(-1))
    monkeypatch.setattr(ops, "dense_rowstats", lambda x, xn2, e, cn2, cos, a, t: R.rowstats(f(x), e, cos, a, t))
    monkeypatch.setattr(ops, "dense_avgprob", lambda x, xn2, e, cn2, cos, a, lse, n: R.avgprob(f(x), e, cos, a, lse, n))
    monkeypatch.setattr(ops, "dense_rowdot",
                        lambda x, xn2, e, cn2, cos, a, lse, tab, n: R.rowdot(f(x), e, cos, a, lse, tab, n))
    monkeypatch.setattr(ops, "dense_backward",
                        lambda x, xn2, ed, cn2, ec, cos, a, lse, coef, target=None, table=None, rdot=None, n_pos=1:
                        R.backward(f(x), ed, ec, cos, a, lse, coef, target, table, rdot, n_pos))

    def cpu_run(self, x, mask, freeze_codebook, fuse_st, want_commit, normalize_input=False, keep_dense=False):
        H, d = x.shape[0], x.shape[-1]
        lead = tuple(x.shape[1:-1])
        flat = x.reshape(H, -1, d)
        if normalize_input:
            flat = torch.nn.functional.normalize(flat.float(), dim=-1)
        emb = self.embeddings.detach()
        update = self.training and self.ema_update and not freeze_codebook
        if keep_dense:
            self.dense_ctx = ops._DenseCtx(flat, emb.clone() if update else emb, self.embeddings, self.use_cosine_sim)
        st = O.CodebookState(self.embeddings.data, self.embed_avg.data, self.cluster_size.data)
        opts = O.CodebookOpts(threshold_ema_dead_code=0, use_cosine_sim=self.use_cosine_sim,
                              weights_l2norm=self.weights_l2norm)
        xf = flat.reshape(H, *lead, d)
        with torch.no_grad():
            q, ind, _ = O.codebook_forward(st, xf.detach(), opts, training=self.training, mask=mask,
                                           freeze_codebook=freeze_codebook)
        commit = None
        if want_commit and mask is not None:      # vqb_gather_st_loss: mean over the rows with mask != 0
            keep = self._expand_mask(mask, flat.shape[1]).bool()
            commit = torch.nn.functional.mse_loss(q.reshape(H, -1, d), flat.float(), reduction="none")[:, keep].mean()
        elif want_commit:
            commit = torch.nn.functional.mse_loss(q, xf.float())
        if fuse_st and self.training:
            q = xf + (q - xf).detach()
        return q, ind, commit

    monkeypatch.setattr(Codebook, "_run", cpu_run)


def build_module(cfg):
    from vqb200 import CodebookParams, VectorQuantize
    cb_dim = cfg.get("cb_dim", cfg["dim"])
    cp = CodebookParams(dim=cb_dim, codebook_size=cfg["K"], threshold_ema_dead_code=0,
                        use_cosine_sim=cfg.get("cosine", False),
                        transform_input="l2norm" if cfg.get("l2in") else "identity",
                        weights_regularization="l2norm" if cfg.get("l2w") else "identity")
    kw = dict(dim=cfg["dim"], codebook_params=cp, sync_codebook=False, heads=cfg.get("heads", 1),
              separate_codebook_per_head=cfg.get("separate", False), channel_last=cfg.get("channel_last", True),
              commitment_weight=cfg.get("cw", 1.0),
              commitment_use_cross_entropy_loss=cfg["kind"] == "commit" or cfg.get("ce_commit", False),
              codebook_diversity_loss_weight=cfg.get("dw", 0.0),
              codebook_diversity_temperature=cfg.get("temp", 100.0))
    if cfg.get("heads", 1) > 1:
        kw["codebook_dim"] = cb_dim
    return VectorQuantize(**kw)


def load_state(vq, fx):
    cb = vq._codebook
    with torch.no_grad():
        dev = cb.embeddings.device
        cb.embeddings.copy_(fx["init"]["embeddings"].to(dev))
        cb.embed_avg.copy_(fx["init"]["embed_avg"].to(dev))
        cb.cluster_size.copy_(fx["init"]["cluster_size"].to(dev))
    cb.invalidate_cache()


@pytest.mark.parametrize("name", gu.dense_fixture_names())
def test_dense_host_logic_matches_reference_fixture(name, monkeypatch):
    _patch(monkeypatch)
    fx = gu.load_dense(name)
    cfg = fx["cfg"]
    vq = build_module(cfg)
    vq.train(cfg["training"])
    load_state(vq, fx)
    x = fx["x"].clone().requires_grad_(True)
    if cfg["kind"] == "indices":
        out = vq(x, indices=fx["targets"])
        assert isinstance(out, tuple) and len(out) == 2
        q, ce = out
        (q.sum() * 0.01 + ce * 1.3 if cfg["training"] else ce * 1.3).backward()
        assert torch.allclose(ce.detach(), fx["ce"], rtol=1e-6)
    else:
        q, ind, loss, bd = vq(x, mask=fx["mask"], return_loss_breakdown=True)
        ((q * fx["w"]).sum() + loss.sum() * 1.7).backward()
        assert torch.equal(ind, fx["indices"])          # incl. the -1 the reference writes into masked positions
        assert loss.shape == fx["loss"].shape
        assert torch.allclose(loss.detach(), fx["loss"], rtol=1e-6, atol=1e-7)
        assert torch.allclose(bd.commitment.detach(), fx["commitment"], rtol=1e-6)
        assert torch.allclose(bd.codebook_diversity.detach(), fx["codebook_diversity"], rtol=1e-6)
    assert torch.equal(q.detach(), fx["quantize"])
    assert gu.rel_err(x.grad, fx["grad_x"]) <= 2e-6
    assert gu.rel_err(vq._codebook.embeddings, fx["after"]["embeddings"]) <= 1e-6
    assert vq._codebook.dense_ctx is None                # nothing of the step stays referenced by the module
