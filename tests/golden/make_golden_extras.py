"""Golden fixtures for the API variants the other fixture families do not pin by VALUE: projections (+ LayerNorm),
freeze_codebook, video input, bf16 input, eval with a mask, ResidualVQ with projections / quantize-dropout /
return_all_codes, GroupedResidualVQ.  Recorded from the UNMODIFIED reference; each fixture carries the constructor
description (so the test builds vqb200's module from the same arguments) and the reference module's state_dict
(so the test also proves that checkpoints load across).

    python tests/golden/make_golden_extras.py       # writes tests/golden/extras/*.pt
"""
import os
import random
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as MG  # noqa: E402

# cls: VectorQuantize | ResidualVQ | GroupedResidualVQ;  cp: CodebookParams kwargs;  kw: module kwargs;
# fwd: forward kwargs;  train: module mode;  dtype: input dtype;  py_seed: random.seed before forward (GroupedResidualVQ
# draws the quantize-dropout seed from python's RNG)
CASES = {
    "x_proj": dict(cls="VectorQuantize", cp=dict(dim=16, codebook_size=40, threshold_ema_dead_code=0),
                   kw=dict(dim=48, codebook_dim=16), shape=(3, 30, 48)),
    "x_proj_ln_heads": dict(cls="VectorQuantize", cp=dict(dim=8, codebook_size=24, threshold_ema_dead_code=0),
                            kw=dict(dim=40, codebook_dim=8, heads=4, separate_codebook_per_head=True,
                                    layernorm_after_project_in=True), shape=(2, 25, 40)),
    "x_freeze": dict(cls="VectorQuantize", cp=dict(dim=24, codebook_size=50, threshold_ema_dead_code=0),
                     kw=dict(dim=24, commitment_weight=0.3), shape=(2, 60, 24), fwd=dict(freeze_codebook=True)),
    "x_video": dict(cls="VectorQuantize", cp=dict(dim=16, codebook_size=32, threshold_ema_dead_code=0),
                    kw=dict(dim=16), shape=(2, 3, 4, 5, 16)),
    "x_video_channel_first_heads": dict(cls="VectorQuantize",
                                        cp=dict(dim=8, codebook_size=20, threshold_ema_dead_code=0),
                                        kw=dict(dim=16, codebook_dim=8, heads=2, channel_last=False),
                                        shape=(2, 16, 3, 4, 5)),
    "x_bf16": dict(cls="VectorQuantize", cp=dict(dim=32, codebook_size=64, threshold_ema_dead_code=0),
                   kw=dict(dim=32), shape=(2, 70, 32), dtype="bfloat16"),
    "x_eval_mask": dict(cls="VectorQuantize", cp=dict(dim=16, codebook_size=32, threshold_ema_dead_code=0),
                        kw=dict(dim=16), shape=(3, 20, 16), train=False, mask=True),
    "x_rvq_proj": dict(cls="ResidualVQ", cp=dict(dim=12, codebook_size=30, threshold_ema_dead_code=0),
                       kw=dict(dim=32, codebook_dim=12, num_quantizers=3), shape=(2, 40, 32)),
    "x_rvq_dropout_all_codes": dict(cls="ResidualVQ", cp=dict(dim=16, codebook_size=24, threshold_ema_dead_code=0),
                                    kw=dict(dim=16, num_quantizers=4, quantize_dropout=True,
                                            quantize_dropout_cutoff_index=1),
                                    shape=(2, 30, 16),
                                    fwd=dict(rand_quantize_dropout_fixed_seed=3, return_all_codes=True)),
    "x_rvq_img_channel_first": dict(cls="ResidualVQ", cp=dict(dim=16, codebook_size=24, threshold_ema_dead_code=0),
                                    kw=dict(dim=16, num_quantizers=2, channel_last=False), shape=(2, 16, 5, 6)),
    "x_grouped": dict(cls="GroupedResidualVQ", cp=dict(dim=8, codebook_size=20, threshold_ema_dead_code=0),
                      kw=dict(dim=16, groups=2, num_quantizers=2), shape=(2, 35, 16), py_seed=11,
                      fwd=dict(return_all_codes=True)),
}


def main():
    MG._import_reference()
    import vector_quantization as VQ
    from vector_quantization.codebooks import CodebookParams
    os.makedirs(os.path.join(HERE, "extras"), exist_ok=True)
    for name, cfg in CASES.items():
        torch.manual_seed(0)
        cls = getattr(VQ, cfg["cls"], None) or getattr(__import__("vector_quantization.residual_vq", fromlist=["x"]),
                                                       cfg["cls"])
        mod = cls(codebook_params=CodebookParams(**cfg["cp"]), sync_codebook=False, **cfg["kw"])
        g = torch.Generator().manual_seed(5)
        with torch.no_grad():          # trained-like codebooks instead of the tiny default init
            for n_, b in mod.named_buffers():
                if n_.endswith("embeddings"):
                    b.copy_(torch.randn(b.shape, generator=g) * 0.6)
            sd0 = mod.state_dict()
            for n_ in list(sd0):
                if n_.endswith("embed_avg"):
                    sd0[n_].copy_(sd0[n_.replace("embed_avg", "embeddings")])
                if n_.endswith("cluster_size"):
                    sd0[n_].fill_(1.0)
        mod.train(cfg.get("train", True))
        x = torch.randn(*cfg["shape"], generator=g)
        if cfg.get("dtype") == "bfloat16":
            x = x.to(torch.bfloat16)
        mask = None
        if cfg.get("mask"):
            mask = torch.rand(cfg["shape"][0], cfg["shape"][1], generator=g) > 0.3
        state0 = {k: v.detach().clone() for k, v in mod.state_dict().items()}
        fwd = dict(cfg.get("fwd", {}))
        if mask is not None:
            fwd["mask"] = mask
        if "py_seed" in cfg:
            random.seed(cfg["py_seed"])
        with torch.no_grad():
            out = mod(x, **fwd)
        fx = {"cfg": cfg, "x": x, "mask": mask, "state_dict": state0,
              "out": [(o if torch.is_tensor(o) else torch.stack(list(o))).detach().clone() for o in out],
              "state_dict_after": {k: v.detach().clone() for k, v in mod.state_dict().items()}}
        torch.save(fx, os.path.join(HERE, "extras", name + ".pt"))
        print(f"{name:30s}", [(tuple(o.shape), str(o.dtype).replace('torch.', '')) for o in fx["out"]])


if __name__ == "__main__":
    main()
