"""Golden fixtures for the learnable-codebook path (reference vector_quantize_pytorch.py:261-273,362 with
learnable_codebook=True, ema_update=False; optionally sync_update_v > 0): forward outputs and the gradients of a fixed
scalar with respect to the input and the codebook, recorded from the UNMODIFIED reference.

    python tests/golden/make_golden_learnable.py       # writes tests/golden/learnable/*.pt
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as MG  # noqa: E402  (einx stand-in + reference import)

CASES = {
    "learnable_plain": {"dim": 32, "K": 48, "shape": (3, 70, 32), "mask": False, "v": 0.0, "cw": 1.0},
    "learnable_masked_v": {"dim": 16, "K": 20, "shape": (2, 33, 16), "mask": True, "v": 0.3, "cw": 0.25},
    # in_place_codebook_optimizer (reference vector_quantize_pytorch.py:233-256): an SGD step on the codebook inside
    # forward, then a second codebook pass; "lr" = learning rate of that SGD
    "inplace_sgd": {"dim": 24, "K": 40, "shape": (3, 50, 24), "mask": False, "v": 0.0, "cw": 1.0, "lr": 0.5},
    "inplace_sgd_masked": {"dim": 16, "K": 24, "shape": (2, 41, 16), "mask": True, "v": 0.2, "cw": 0.5, "lr": 2.0},
    # learnable codebook + the losses on the dense similarities: the codebook receives their gradient through
    # `similarities` (codebooks.py:375-377 leaves `embeddings` attached)
    "learnable_ce_commit": {"dim": 24, "K": 70, "shape": (3, 45, 24), "mask": True, "v": 0.0, "cw": 0.8, "ce": True},
    "learnable_ce_commit_dot": {"dim": 16, "K": 40, "shape": (2, 50, 16), "mask": False, "v": 0.3, "cw": 1.0,
                                "ce": True, "cosine": True},
    "learnable_diversity": {"dim": 32, "K": 48, "shape": (4, 30, 32), "mask": False, "v": 0.0, "cw": 0.6, "dw": 0.5,
                            "temp": 2.0},
    "learnable_ce_indices": {"dim": 16, "K": 36, "shape": (2, 40, 16), "mask": False, "v": 0.0, "cw": 1.0,
                             "indices": True},
}


def main():
    VectorQuantize, _, CodebookParams, _ = MG._import_reference()
    for name, cfg in CASES.items():
        torch.manual_seed(0)
        cp = CodebookParams(dim=cfg["dim"], codebook_size=cfg["K"], learnable_codebook=True, ema_update=False,
                            threshold_ema_dead_code=0, use_cosine_sim=cfg.get("cosine", False))
        extra = {}
        if cfg.get("ce"):
            extra["commitment_use_cross_entropy_loss"] = True
        if cfg.get("dw"):
            extra.update(codebook_diversity_loss_weight=cfg["dw"], codebook_diversity_temperature=cfg["temp"])
        if "lr" in cfg:
            extra["in_place_codebook_optimizer"] = lambda params, lr=cfg["lr"]: torch.optim.SGD(params, lr=lr)
        vq = VectorQuantize(dim=cfg["dim"], codebook_params=cp, commitment_weight=cfg["cw"], sync_update_v=cfg["v"],
                            sync_codebook=False, **extra).train()
        g = torch.Generator().manual_seed(7)
        with torch.no_grad():
            vq._codebook.embeddings.copy_(torch.randn(vq._codebook.embeddings.shape, generator=g) * 0.7)
        x = torch.randn(*cfg["shape"], generator=g).requires_grad_(True)
        mask = None
        if cfg["mask"]:
            b, n = cfg["shape"][:2]
            mask = torch.rand(b, n, generator=g) > 0.3
        w = torch.randn(*cfg["shape"], generator=g)
        init = vq._codebook.embeddings.detach().clone()
        if cfg.get("indices"):
            tgt = torch.randint(0, cfg["K"], cfg["shape"][:2], generator=g)
            tgt.view(-1)[::6] = -1
            q, ce = vq(x, indices=tgt)
            (q.sum() * 0.01 + ce * 1.3).backward()
            fx = {"cfg": cfg, "x": x.detach().clone(), "mask": None, "w": w, "init_embeddings": init, "targets": tgt,
                  "quantize": q.detach().clone(), "ce": ce.detach().clone(), "grad_x": x.grad.clone(),
                  "grad_embeddings": vq._codebook.embeddings.grad.clone()}
            torch.save(fx, os.path.join(HERE, "learnable", name + ".pt"))
            print(name, "ce", float(ce), "|grad_emb|", float(fx["grad_embeddings"].abs().max()))
            continue
        q, ind, loss, bd = vq(x, mask=mask, return_loss_breakdown=True)
        (q * w).sum().add(loss.sum() * 1.7).backward()
        fx = {"cfg": cfg, "x": x.detach().clone(), "mask": mask, "w": w, "init_embeddings": init,
              "quantize": q.detach().clone(), "indices": ind.clone(), "loss": loss.detach().clone(),
              "grad_x": x.grad.clone(), "grad_embeddings": vq._codebook.embeddings.grad.clone(),
              "inplace_optimize": bd.inplace_optimize.detach().clone(),
              "after_embeddings": vq._codebook.embeddings.detach().clone()}
        torch.save(fx, os.path.join(HERE, "learnable", name + ".pt"))
        print(name, "loss", float(loss), "|grad_emb|", float(fx["grad_embeddings"].abs().max()))


if __name__ == "__main__":
    main()
