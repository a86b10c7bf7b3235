"""Golden fixtures for the consumers of the dense N x K similarities (reference vector_quantize_pytorch.py:284-299
cross-entropy to given indices, :338-346 cross-entropy commitment loss, :324-333 codebook diversity loss): forward
outputs, the buffers after the step and the gradient of a fixed scalar with respect to the input, recorded from the
UNMODIFIED reference.

    python tests/golden/make_golden_dense.py       # writes tests/golden/dense/*.pt
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as MG  # noqa: E402  (einx stand-in + reference import)

# kind: "indices" = forward(x, indices=...) -> (quantize, ce);  "commit" = CE commitment;  "diversity" = diversity loss
CASES = {
    "ce_indices_train": dict(kind="indices", dim=16, K=24, shape=(2, 9, 16), training=True, ignore=True),
    "ce_indices_eval": dict(kind="indices", dim=32, K=80, shape=(3, 50, 32), training=False, ignore=True),
    "ce_indices_cosine": dict(kind="indices", dim=24, K=70, shape=(2, 37, 24), training=True, cosine=True, l2in=True,
                              l2w=True),
    "ce_indices_heads": dict(kind="indices", dim=32, K=40, shape=(2, 21, 32), training=True, heads=4, cb_dim=8,
                             separate=True, ignore=True),
    "ce_commit": dict(kind="commit", dim=32, K=100, shape=(3, 41, 32), training=True, cw=0.7),
    "ce_commit_masked": dict(kind="commit", dim=16, K=33, shape=(2, 30, 16), training=True, mask=True, cw=1.0),
    "ce_commit_cosine_heads": dict(kind="commit", dim=32, K=48, shape=(2, 25, 32), training=True, cosine=True,
                                   l2in=True, l2w=True, heads=2, cb_dim=16, cw=0.5),
    "diversity_t2": dict(kind="diversity", dim=16, K=40, shape=(5, 19, 16), training=True, dw=0.6, temp=2.0),
    "diversity_default_t": dict(kind="diversity", dim=32, K=72, shape=(4, 33, 32), training=True, dw=1.0, temp=100.0,
                                cb_scale=0.05, x_scale=0.05),
    "diversity_cosine_heads": dict(kind="diversity", dim=32, K=36, shape=(3, 17, 32), training=True, dw=0.8, temp=5.0,
                                   cosine=True, l2in=True, l2w=True, heads=2, cb_dim=16),
    # (CE commitment on image / video input raises inside the reference: the indices are already un-flattened at :346)
    "all_three_chfirst": dict(kind="diversity", dim=16, K=30, shape=(2, 16, 25), training=True, dw=0.3, temp=3.0,
                              ce_commit=True, cw=0.4, channel_last=False),
    "diversity_img": dict(kind="diversity", dim=16, K=30, shape=(2, 16, 5, 5), training=True, dw=0.3, temp=3.0,
                          channel_last=False),
}


def build(cfg, VectorQuantize, CodebookParams):
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
    if cb_dim * cfg.get("heads", 1) != cfg["dim"]:
        kw["codebook_dim"] = cb_dim
    elif cfg.get("heads", 1) > 1:
        kw["codebook_dim"] = cb_dim
    return VectorQuantize(**kw)


def main():
    VectorQuantize, _, CodebookParams, _ = MG._import_reference()
    os.makedirs(os.path.join(HERE, "dense"), exist_ok=True)
    for name, cfg in CASES.items():
        torch.manual_seed(0)
        vq = build(cfg, VectorQuantize, CodebookParams)
        vq.train(cfg["training"])
        cb = vq._codebook
        g = torch.Generator().manual_seed(11)
        with torch.no_grad():
            e = torch.randn(cb.embeddings.shape, generator=g) * cfg.get("cb_scale", 0.6)
            if cfg.get("l2w"):
                e = torch.nn.functional.normalize(e, dim=-1)
            cb.embeddings.copy_(e)
            cb.embed_avg.copy_(e)
            cb.cluster_size.fill_(1.0)
        proj = {k: v.detach().clone() for k, v in vq.state_dict().items() if k.startswith("project")}
        x = (torch.randn(*cfg["shape"], generator=g) * cfg.get("x_scale", 1.0)).requires_grad_(True)
        lead = (cfg["shape"][0],) + tuple(cfg["shape"][2:] if not cfg.get("channel_last", True) else cfg["shape"][1:-1])
        mask = None
        if cfg.get("mask"):
            mask = torch.rand(lead, generator=g) > 0.3
        w = torch.randn(*cfg["shape"], generator=g)
        init = MG._snap(cb)
        fx = {"cfg": cfg, "x": x.detach().clone(), "mask": mask, "w": w, "init": init, "proj": proj}
        if cfg["kind"] == "indices":
            tshape = lead + ((cfg["heads"],) if cfg.get("heads", 1) > 1 else ())
            tgt = torch.randint(0, cfg["K"], tshape, generator=g)
            if cfg.get("ignore"):
                tgt.view(-1)[::5] = -1
            fx["targets"] = tgt.clone()
            q, ce = vq(x, indices=tgt)
            # the returned quantize is the codebook-side tensor (before head merge / projection): weight it with ones
            scalar = q.sum() * 0.01 + ce * 1.3 if cfg["training"] else ce * 1.3
            scalar.backward()
            fx.update(quantize=q.detach().clone(), ce=ce.detach().clone())
        else:
            q, ind, loss, bd = vq(x, mask=mask, return_loss_breakdown=True)
            ((q * w).sum() + loss.sum() * 1.7).backward()
            fx.update(quantize=q.detach().clone(), indices=ind.clone(), loss=loss.detach().clone(),
                      commitment=bd.commitment.detach().clone(),
                      codebook_diversity=bd.codebook_diversity.detach().clone())
        fx["grad_x"] = x.grad.clone()
        fx["after"] = MG._snap(cb)
        torch.save(fx, os.path.join(HERE, "dense", name + ".pt"))
        print(name, {k: (float(v) if v.numel() == 1 else tuple(v.shape)) for k, v in fx.items()
                     if k in ("ce", "loss", "commitment", "codebook_diversity")},
              "|grad_x|", float(fx["grad_x"].abs().max()))


if __name__ == "__main__":
    main()
