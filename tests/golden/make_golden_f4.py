"""Golden fixtures for the f4 variants: gumbel sampling (stochastic / straight-through / reinmax,
reference utils/general.py:107-151 at codebooks.py:388), the affine re-parametrisation (codebooks.py:274-348,373-384,
400-403) and the orthogonal regularisation of the codebook (vector_quantize_pytorch.py:366-390, utils/losses.py:22-27).

    python tests/golden/make_golden_f4.py       # writes tests/golden/f4/*.pt

Recorded from the reference running in this container.  Two recording shims, neither changes what the reference
computes:
  * `utils.general.gumbel_noise` is wrapped by a function that executes the reference's own two statements and keeps a
    copy of the uniform draw (the CPU generator's numbers cannot be re-created on a GPU; tests inject them);
  * `Codebook.embed` is defined as an alias of `Codebook.embeddings`: the reference's orthogonal-loss block reads
    `self._codebook.embed` (vector_quantize_pytorch.py:367), an attribute that does not exist -- unpatched it raises
    AttributeError before computing anything.  SURVEY 8(b) prescribes `embeddings`.
`torch.randperm` is wrapped the same way for `orthogonal_reg_max_codes`.
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as MG  # noqa: E402

DRAWS, PERMS = [], []


def _shims():
    import vector_quantization.utils.general as G
    from vector_quantization.codebooks import Codebook

    def gumbel_noise(t):
        noise = torch.zeros_like(t).uniform_(0, 1)          # reference utils/general.py:108
        DRAWS.append(noise.clone())
        return -G.log(-G.log(noise))                         # :109

    G.gumbel_noise = gumbel_noise
    Codebook.embed = property(lambda self: self.embeddings)
    orig = torch.randperm

    def randperm(n, *a, **k):
        p = orig(n, *a, **k)
        PERMS.append(p.clone())
        return p

    return orig, randperm


def _snap(cb):
    out = {k: v.detach().clone() for k, v in cb.state_dict().items()}
    return out


CASES = {
    # ---- stochastic sampling (EMA codebook): forward values + EMA buffers
    "g_stochastic": dict(kind="vq", dim=24, K=40, shape=(3, 50, 24), thr=0, gumbel=dict(stochastic=True, temperature=0.7),
                         steps=2),
    "g_stochastic_cos_eval": dict(kind="vq", dim=16, K=32, shape=(2, 40, 16), thr=0, cosine=True, l2=True, train=False,
                                  gumbel=dict(stochastic=True, temperature=1.0), steps=1),
    "g_stochastic_masked_expire": dict(kind="vq", dim=16, K=48, shape=(2, 60, 16), thr=2, mask=True,
                                       gumbel=dict(stochastic=True, temperature=0.5), steps=2),
    # ---- straight-through / reinmax with an EMA codebook through VectorQuantize: the gradient through the one-hot is
    # dropped by the reference (commit_quantize is detached), values as the hard one-hot
    "g_st_ema": dict(kind="vq", dim=16, K=24, shape=(2, 30, 16), thr=0, grads=True,
                     gumbel=dict(straight_through=True, temperature=0.9), steps=1),
    # ---- learnable codebook: the gradient flows through the soft one-hot
    "g_st_learnable": dict(kind="vq", dim=16, K=24, shape=(2, 30, 16), thr=0, learnable=True, grads=True, cw=0.7,
                           gumbel=dict(straight_through=True, temperature=0.9), steps=1),
    "g_st_stochastic_learnable_masked": dict(kind="vq", dim=12, K=20, shape=(3, 21, 12), thr=0, learnable=True,
                                             grads=True, mask=True, cw=1.0,
                                             gumbel=dict(straight_through=True, stochastic=True, temperature=0.6), steps=1),
    "g_reinmax_learnable": dict(kind="vq", dim=16, K=24, shape=(2, 30, 16), thr=0, learnable=True, grads=True, cw=1.0,
                                gumbel=dict(straight_through=True, reinmax=True, temperature=0.8), steps=1),
    # ---- Codebook.forward used directly under autograd (quantize keeps the graph through the one-hot)
    "g_st_codebook_direct": dict(kind="cb", dim=16, K=24, shape=(2, 30, 16), thr=0, grads=True,
                                 gumbel=dict(straight_through=True, temperature=0.9), steps=1),
    # ---- the reference's own test configuration (tests/test_residual_vq.py:39-73)
    "g_rvq_stochastic_shared": dict(kind="rvq", dim=8, K=32, shape=(1, 100, 8), thr=2, shared=True, Q=3,
                                    gumbel=dict(stochastic=True), steps=1),
    # ---- affine re-parametrisation
    "affine_cb": dict(kind="cb", dim=16, K=32, shape=(2, 60, 16), thr=0, affine=dict(sync=False), steps=3),
    "affine_cb_masked_eval": dict(kind="cb", dim=12, K=20, shape=(2, 31, 12), thr=0, mask=True, affine=dict(sync=False),
                                  steps=2, eval_last=True),
    # (through VectorQuantize the reference passes `asdict(codebook_params)`, which turns AffineParameters into a dict, and
    # then raises AttributeError at codebooks.py:288 -- only a directly constructed Codebook runs, hence kind="cb")
    "affine_cb_expire": dict(kind="cb", dim=16, K=40, shape=(2, 50, 16), thr=2,
                             affine=dict(sync=False, batch_decay=0.9, codebook_decay=0.8), steps=2),
    # ---- orthogonal regularisation
    "orth_plain": dict(kind="vq", dim=16, K=24, shape=(2, 30, 16), thr=0, orth=dict(orthogonal_reg_weight=0.7), grads=True,
                       steps=1),
    "orth_active_max": dict(kind="vq", dim=12, K=40, shape=(2, 25, 12), thr=0, grads=True, steps=1,
                            orth=dict(orthogonal_reg_weight=1.3, orthogonal_reg_active_codes_only=True,
                                      orthogonal_reg_max_codes=9)),
    "orth_learnable_heads": dict(kind="vq", dim=16, K=20, shape=(2, 18, 16), thr=0, grads=True, steps=1, learnable=True,
                                 heads=2, cb_dim=8, orth=dict(orthogonal_reg_weight=0.5)),
}


def main():
    MG._import_reference()
    import vector_quantization as VQ
    from vector_quantization.codebooks import AffineParameters, Codebook, CodebookParams, GumbelParams
    from vector_quantization.residual_vq import ResidualVQ
    orig_randperm, rec_randperm = _shims()
    os.makedirs(os.path.join(HERE, "f4"), exist_ok=True)
    for name, cfg in CASES.items():
        torch.manual_seed(0)
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
            mod = VQ.VectorQuantize(dim=cfg["dim"], codebook_params=CodebookParams(**cpk), sync_codebook=False,
                                    commitment_weight=cfg.get("cw", 1.0), **kw)
            books = [mod._codebook]
        g = torch.Generator().manual_seed(9)
        with torch.no_grad():
            for b in {id(b): b for b in books}.values():
                e = torch.randn(b.embeddings.shape, generator=g) * 0.6
                if cfg.get("l2"):
                    e = torch.nn.functional.normalize(e, dim=-1)
                b.embeddings.copy_(e); b.embed_avg.copy_(e); b.cluster_size.fill_(1.0)
        init = [_snap(b) for b in books]
        steps = []
        for si in range(cfg["steps"]):
            training = cfg.get("train", True) and not (cfg.get("eval_last") and si == cfg["steps"] - 1)
            mod.train(training)
            x = torch.randn(*cfg["shape"], generator=g)
            mask = None
            if cfg.get("mask"):
                mask = torch.rand(cfg["shape"][0], cfg["shape"][1], generator=g) > 0.3
            w = torch.randn(*cfg["shape"], generator=g)
            if cfg.get("grads"):
                x.requires_grad_(True)
            DRAWS.clear(); PERMS.clear()
            torch.manual_seed(100 + si)
            torch.randperm = rec_randperm if cfg.get("orth") else orig_randperm
            rec = {"x": x.detach().clone(), "mask": mask, "w": w, "training": training, "rng_seed": 100 + si}
            if cfg["kind"] == "cb":
                q, ind, sim = mod(x[None] if False else x, mask=mask)
                rec.update(quantize=q.detach().clone(), indices=ind.clone(), similarities=sim.detach().clone())
                if cfg.get("grads"):
                    (q * w).sum().backward()
                    rec["grad_x"] = x.grad.clone()
            elif cfg["kind"] == "rvq":
                q, ind, loss = mod(x)
                rec.update(quantize=q.detach().clone(), indices=ind.clone(), loss=loss.detach().clone())
            else:
                q, ind, loss, bd = mod(x, mask=mask, return_loss_breakdown=True)
                rec.update(quantize=q.detach().clone(), indices=ind.clone(), loss=loss.detach().clone(),
                           breakdown=[t.detach().clone() for t in bd])
                if cfg.get("grads"):
                    (q * w).sum().add(loss.sum() * 1.7).backward()
                    rec["grad_x"] = x.grad.clone()
                    e = books[0].embeddings
                    rec["grad_embeddings"] = e.grad.clone() if getattr(e, "grad", None) is not None else None
                    if e.grad is not None:
                        e.grad = None
            torch.randperm = orig_randperm
            rec["draws"] = [d.clone() for d in DRAWS]
            rec["perms"] = [p.clone() for p in PERMS]
            rec["after"] = [_snap(b) for b in books]
            steps.append(rec)
        torch.save({"cfg": cfg, "init": init, "steps": steps}, os.path.join(HERE, "f4", name + ".pt"))
        last = steps[-1]
        print(name, "ok: draws", [tuple(d.shape) for d in last["draws"]], "perms", len(last["perms"]),
              "loss", float(last["loss"].sum()) if "loss" in last else None)


if __name__ == "__main__":
    main()
