"""Golden fixtures for ResidualVQ over LEARNABLE codebooks (learnable_codebook=True, ema_update=False) where the input
carries NO gradient (frozen encoder / precomputed features): the commitment losses must still reach every level's
codebook Parameter (reference vector_quantize_pytorch.py:263-269 through residual_vq.py:212-243).  Recorded from the
UNMODIFIED reference.  (the fused level loop ran under no_grad here and silently dropped those gradients.)

    python tests/golden/make_golden_rvq_learnable.py       # writes tests/golden/rvq_learnable/*.pt
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as MG  # noqa: E402  (einx stand-in + reference import)

CASES = {
    "rvq_learnable_nograd": {"dim": 32, "K": 40, "Q": 3, "shape": (2, 60, 32), "mask": False, "cw": 1.0,
                             "x_grad": False},
    "rvq_learnable_nograd_masked": {"dim": 16, "K": 24, "Q": 4, "shape": (3, 25, 16), "mask": True, "cw": 0.5,
                                    "x_grad": False},
    "rvq_learnable_xgrad": {"dim": 16, "K": 24, "Q": 2, "shape": (2, 30, 16), "mask": False, "cw": 1.0,
                            "x_grad": True},
}


def main():
    _, ResidualVQ, CodebookParams, _ = MG._import_reference()
    os.makedirs(os.path.join(HERE, "rvq_learnable"), exist_ok=True)
    for name, cfg in CASES.items():
        torch.manual_seed(0)
        cp = CodebookParams(dim=cfg["dim"], codebook_size=cfg["K"], learnable_codebook=True, ema_update=False,
                            threshold_ema_dead_code=0)
        rvq = ResidualVQ(dim=cfg["dim"], num_quantizers=cfg["Q"], codebook_params=cp, commitment_weight=cfg["cw"],
                         sync_codebook=False).train()
        g = torch.Generator().manual_seed(3)
        init = []
        with torch.no_grad():
            for li, layer in enumerate(rvq.layers):
                e = layer._codebook.embeddings
                e.copy_(torch.randn(e.shape, generator=g) * (0.7 / 1.5 ** li))
                init.append(e.detach().clone())
        x = torch.randn(*cfg["shape"], generator=g)
        if cfg["x_grad"]:
            x.requires_grad_(True)
        mask = None
        if cfg["mask"]:
            mask = torch.rand(*cfg["shape"][:2], generator=g) > 0.3
        w = torch.randn(*cfg["shape"], generator=g)
        q, ind, losses = rvq(x, mask=mask)
        ((q * w).sum() * (1.0 if cfg["x_grad"] else 0.0) + (losses * torch.arange(1, cfg["Q"] + 1)).sum()).backward()
        fx = {"cfg": cfg, "x": x.detach().clone(), "mask": mask, "w": w, "init_embeddings": init,
              "quantize": q.detach().clone(), "indices": ind.clone(), "losses": losses.detach().clone(),
              "grad_x": x.grad.clone() if cfg["x_grad"] else None,
              "grad_embeddings": [l._codebook.embeddings.grad.clone() for l in rvq.layers]}
        torch.save(fx, os.path.join(HERE, "rvq_learnable", name + ".pt"))
        print(name, "losses", losses.tolist(), "|grad_emb|", [float(t.abs().max()) for t in fx["grad_embeddings"]])


if __name__ == "__main__":
    main()
