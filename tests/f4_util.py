"""Shared by the CPU orchestration test and the GPU test of the f4 variants: build the vqb200 module a fixture of
tests/golden/f4 describes, load its initial state, run its steps with the recorded random draws injected, compare."""
import glob
import os

import torch

import golden_util as gu

F4_DIR = os.path.join(gu.GOLDEN_DIR, "f4")


def names():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(F4_DIR, "*.pt")))


def load(name):
    return torch.load(os.path.join(F4_DIR, name + ".pt"), weights_only=False)


def build(cfg, dev):
    from vqb200 import Codebook, CodebookParams, GumbelParams, ResidualVQ, VectorQuantize
    from vqb200.params import AffineParameters
    cpk = dict(dim=cfg.get("cb_dim", cfg["dim"]), codebook_size=cfg["K"], threshold_ema_dead_code=cfg["thr"])
    if cfg.get("gumbel"):
        cpk["gumbel_params"] = GumbelParams(**cfg["gumbel"])
    if cfg.get("affine"):
        cpk.update(use_affine=True, affine_params=AffineParameters(**cfg["affine"]))
    if cfg.get("learnable"):
        cpk.update(learnable_codebook=True, ema_update=False)
    if cfg.get("cosine"):
        cpk.update(use_cosine_sim=True)
    if cfg.get("l2"):
        cpk.update(transform_input="l2norm", weights_regularization="l2norm")
    torch.manual_seed(0)
    if cfg["kind"] == "cb":
        mod = Codebook(**cpk)
        books = [mod]
    elif cfg["kind"] == "rvq":
        mod = ResidualVQ(dim=cfg["dim"], num_quantizers=cfg["Q"], shared_codebook=cfg.get("shared", False),
                         codebook_params=CodebookParams(**cpk), sync_codebook=False)
        books = [l._codebook for l in mod.layers]
    else:
        kw = dict(cfg.get("orth", {}))
        if "heads" in cfg:
            kw.update(heads=cfg["heads"], codebook_dim=cfg["cb_dim"])
        mod = VectorQuantize(dim=cfg["dim"], codebook_params=CodebookParams(**cpk), sync_codebook=False,
                             commitment_weight=cfg.get("cw", 1.0), **kw)
        books = [mod._codebook]
    mod = mod.to(dev)
    return mod, books


def load_init(books, fx):
    with torch.no_grad():
        for b, st in zip(books, fx["init"]):
            for k in ("embeddings", "embed_avg", "cluster_size"):
                getattr(b, k).copy_(st[k])
            b.invalidate_cache()


def run_and_check(fx, dev, monkeypatch, tol_buf=1e-5, tol_grad=2e-5):
    """Returns nothing; asserts.  Index mismatches are allowed only where the reference's own top-2 gap of the
    sampling logits is below 1e-6 relative (never the case in these fixtures: asserted equal)."""
    from vqb200 import codebook as CB, vq as VQM
    cfg = fx["cfg"]
    mod, books = build(cfg, dev)
    load_init(books, fx)
    for si, st in enumerate(fx["steps"]):
        mod.train(st["training"])
        x = st["x"].to(dev).clone()
        if cfg.get("grads"):
            x.requires_grad_(True)
        mask = st["mask"].to(dev) if st["mask"] is not None else None
        w = st["w"].to(dev)
        # the gumbel uniforms: the reference draws them from the CPU generator, interleaved with the draws of the
        # dead-code replacement; the sampling op is wrapped so that it makes the reference's own draw on the CPU
        # generator (same seed -> same numbers, checked against the recorded ones) and hands it to the kernel
        from vqb200 import ops as OPS
        real_sample = OPS.dense_gumbel_sample
        recorded = [d for d in st["draws"]]

        def sample_with_cpu_draw(xx, emb, cos, tau, uniforms=None, generator=None, _real=real_sample):
            u = torch.zeros(xx.shape[0], xx.shape[1], emb.shape[1]).uniform_(0, 1)      # reference general.py:108
            assert recorded and torch.equal(u, recorded.pop(0)), "the CPU draw is not the recorded one"
            return _real(xx, emb, cos, tau, uniforms=u.to(xx.device))
        monkeypatch.setattr(OPS, "dense_gumbel_sample", sample_with_cpu_draw)
        # ... the orthogonal loss' randperm ...
        perms = [p.to(dev) for p in st["perms"]]
        if perms:
            monkeypatch.setattr(VQM.VectorQuantize, "_draw_perm", staticmethod(lambda n, device, _p=perms: _p.pop(0)))
        # ... and the dead-code replacement rows (drawn by the reference's calls on the CPU generator)

        def cpu_draw(num_rows, m, device):
            r = torch.randperm(num_rows)[:m] if num_rows >= m else torch.randint(0, num_rows, (m,))
            return r.to(device)
        monkeypatch.setattr(CB.Codebook, "_draw_rows", staticmethod(cpu_draw))
        torch.manual_seed(st["rng_seed"])
        if cfg["kind"] == "cb":
            mod.return_similarities = True
            q, ind, sim = mod(x, mask=mask)
            assert sim is not None and tuple(sim.shape) == tuple(st["similarities"].shape)
            assert gu.rel_err(sim.detach().cpu(), st["similarities"]) <= 1e-5
            scalar = (q * w).sum()
        elif cfg["kind"] == "rvq":
            q, ind, loss = mod(x)
            scalar = None
        else:
            q, ind, loss, bd = mod(x, mask=mask, return_loss_breakdown=True)
            scalar = (q * w).sum() + loss.sum() * 1.7
        monkeypatch.setattr(OPS, "dense_gumbel_sample", real_sample)
        assert not recorded, "fewer gumbel draws than the reference made"
        assert torch.equal(ind.cpu(), st["indices"]), f"step {si}: indices differ"
        assert gu.rel_err(q.detach().cpu(), st["quantize"]) <= 1e-6, f"step {si}: quantize"
        if "loss" in st:
            assert torch.allclose(loss.detach().cpu().reshape(-1), st["loss"].reshape(-1), rtol=1e-5, atol=1e-7), \
                (si, loss, st["loss"])
        if "breakdown" in st:
            for a, b in zip(bd, st["breakdown"]):
                assert torch.allclose(a.detach().cpu().reshape(-1), b.reshape(-1), rtol=2e-5, atol=1e-7), (si, a, b)
        if cfg.get("grads") and scalar is not None:
            scalar.backward()
            assert gu.rel_err(x.grad.cpu(), st["grad_x"]) <= tol_grad, f"step {si}: grad_x"
            if st.get("grad_embeddings") is not None:
                e = books[0].embeddings
                assert e.grad is not None, "the codebook received no gradient"
                assert gu.rel_err(e.grad.cpu(), st["grad_embeddings"]) <= tol_grad, f"step {si}: grad_embeddings"
                e.grad = None
        for b, after in zip(books, st["after"]):
            sd = b.state_dict()
            for k, ref in after.items():
                if k.endswith("_needs_init"):
                    continue
                got = sd[k].detach().cpu()
                if k == "cluster_size" and not cfg.get("gumbel", {}).get("straight_through"):
                    assert gu.rel_err(got, ref) <= 1e-6, (si, k)
                else:
                    assert gu.rel_err(got, ref) <= tol_buf, (si, k, gu.rel_err(got, ref))
