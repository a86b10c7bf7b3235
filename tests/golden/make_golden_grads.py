"""Golden fixtures WITH input gradients on the EMA path: one training forward of the UNMODIFIED reference under
autograd and the gradient of a fixed scalar, (quantize * w).sum() + 1.7 * loss.sum(), with respect to the input.
The EMA step moves the codebook inside that forward, so these pin which codebook (pre- / post-update) every term of
the backward pass reads.

    python tests/golden/make_golden_grads.py       # writes tests/golden/grads/*.pt
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as MG  # noqa: E402

CASES = {
    "g_euclid": dict(kind="vq", dim=32, K=64, shape=(4, 128, 32), thr=0, cb_scale=0.5),
    "g_default_init": dict(kind="vq", dim=32, K=48, shape=(2, 100, 32), thr=0),
    "g_masked": dict(kind="vq", dim=32, K=64, shape=(3, 50, 32), thr=0, cb_scale=0.5, mask=True),
    "g_heads_shared": dict(kind="vq", dim=32, K=64, shape=(2, 40, 32), thr=0, heads=4, cb_dim=8, cb_scale=0.5),
    "g_heads_separate_masked": dict(kind="vq", dim=32, K=64, shape=(2, 40, 32), thr=0, heads=4, cb_dim=8,
                                    separate=True, cb_scale=0.5, mask=True),
    "g_cosine_l2": dict(kind="vq", dim=64, K=128, shape=(2, 96, 64), thr=0, cosine=True, l2in=True, l2w=True),
    "g_channel_first_img": dict(kind="vq", dim=16, K=32, shape=(2, 16, 6, 6), thr=0, channel_last=False,
                                cb_scale=0.6),
    "g_one_vector": dict(kind="vq", dim=8, K=16, shape=(5, 8), thr=0, cb_scale=0.8),
    "g_rvq3": dict(kind="rvq", dim=32, K=64, Q=3, shape=(2, 64, 32), thr=0, cb_scale=0.5),
    "g_rvq_shared_masked": dict(kind="rvq", dim=16, K=32, Q=3, shape=(2, 48, 16), thr=0, shared=True, cb_scale=0.5,
                                mask=True),
}


def main():
    api = MG._import_reference()
    os.makedirs(os.path.join(HERE, "grads"), exist_ok=True)
    for name, cfg in CASES.items():
        mod, books = MG.build(cfg, *api)
        if cfg["kind"] == "rvq" and "cb_scale" in cfg and not cfg.get("shared"):
            pass    # MG.build already gave every level its own scaled codebook
        mod.train()
        g = torch.Generator().manual_seed(99)
        x = torch.randn(*cfg["shape"], generator=g).requires_grad_(True)
        w = torch.randn(*cfg["shape"], generator=g)
        mask = None
        if cfg.get("mask"):
            b, n = cfg["shape"][:2]
            mask = torch.rand(b, n, generator=g) > 0.35
        init = [MG._snap(b) for b in books]
        q, ind, loss = mod(x, mask=mask)
        ((q * w).sum() + loss.sum() * 1.7).backward()
        fx = {"cfg": cfg, "x": x.detach().clone(), "w": w, "mask": mask, "init": init,
              "quantize": q.detach().clone(), "indices": ind.clone(), "loss": loss.detach().clone(),
              "grad_x": x.grad.clone(), "after": [MG._snap(b) for b in books]}
        torch.save(fx, os.path.join(HERE, "grads", name + ".pt"))
        moved = max(float((a["embeddings"] - i["embeddings"]).abs().max()) for a, i in zip(fx["after"], init))
        print(f"{name:26s} loss {loss.detach().flatten().tolist()}  |grad - w| "
              f"{float((x.grad - w).abs().max()):.3e}  codebook moved {moved:.3f}")


if __name__ == "__main__":
    main()
