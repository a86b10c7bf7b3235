"""Generate golden fixtures by running the UNMODIFIED reference in the build container.

    python tests/golden/make_golden.py            # writes tests/golden/*.pt

The reference (/root/reference, read-only, pure Python) cannot travel to the GPU
box, so its outputs on seeded inputs are committed here as small fixtures, with
this script.  ``einx`` is not installed in the image; the reference uses exactly
one einx call (``get_at("q [c] d, b n q -> q b n d", codebooks, indices)``,
reference residual_vq.py:117), so a stand-in module providing only that pattern
is put on ``sys.path`` first.  Nothing else of the reference is altered.

Each fixture is a dict of plain tensors / python scalars:
  cfg        : kwargs describing the case (our own schema, see CASES)
  x          : input latents
  mask       : optional bool mask
  init       : list (one per level) of {embeddings, embed_avg, cluster_size} BEFORE the steps
  steps      : list (one per forward call) of {quantize, indices, loss,
               after: [{embeddings, embed_avg, cluster_size} per level], top2_rel_gap}
  rng_seed   : torch.manual_seed value set right before every forward call (expiry RNG)
"""
import os
import sys
import tempfile
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("VQB_REFERENCE", "/root/reference")


def _install_einx_standin():
    m = types.ModuleType("einx")

    def get_at(pattern, codebooks, indices):
        assert pattern.replace(" ", "") == "q[c]d,bnq->qbnd", pattern
        q = codebooks.shape[0]
        return torch.stack([codebooks[i][indices[..., i]] for i in range(q)], 0)

    m.get_at = get_at
    sys.modules["einx"] = m


def _import_reference():
    _install_einx_standin()
    sys.path.insert(0, REF)
    from vector_quantization import ResidualVQ, VectorQuantize  # noqa
    from vector_quantization.codebooks import CodebookParams, KmeansParameters  # noqa
    return VectorQuantize, ResidualVQ, CodebookParams, KmeansParameters


def _snap(cb):
    return {"embeddings": cb.embeddings.detach().clone(), "embed_avg": cb.embed_avg.clone(),
            "cluster_size": cb.cluster_size.clone()}


def _gap(sim):
    top2 = sim.topk(2, dim=-1).values
    return ((top2[..., 0] - top2[..., 1]).abs() / top2[..., 0].abs().clamp_min(1e-30))


# our own case schema -------------------------------------------------------
CASES = {
    # config 1 of BASELINE.json: README shape, default ctor (thr=2 -> every code expires on step 1)
    "c1_default": dict(kind="vq", dim=256, K=512, shape=(1, 1024, 256), steps=1, thr=2),
    "c1_noexpire": dict(kind="vq", dim=256, K=512, shape=(1, 1024, 256), steps=2, thr=0),
    "eval_small": dict(kind="vq", dim=64, K=128, shape=(2, 128, 64), steps=1, thr=0, training=False),
    # randn-scale codebook (trained-like), small
    "euclid_small": dict(kind="vq", dim=32, K=64, shape=(4, 128, 32), steps=3, thr=0, cb_scale=0.5),
    "euclid_expire": dict(kind="vq", dim=32, K=256, shape=(2, 96, 32), steps=3, thr=2, cb_scale=0.5),
    "cosine_l2": dict(kind="vq", dim=64, K=128, shape=(2, 256, 64), steps=3, thr=2, cosine=True,
                      l2in=True, l2w=True),
    "cosine_raw": dict(kind="vq", dim=16, K=32, shape=(1, 100, 16), steps=2, thr=0, cosine=True),
    "masked": dict(kind="vq", dim=32, K=64, shape=(3, 50, 32), steps=2, thr=0, cb_scale=0.5, mask=True),
    "heads_shared": dict(kind="vq", dim=32, K=64, shape=(2, 40, 32), steps=2, thr=0, heads=4, cb_dim=8),
    "heads_separate": dict(kind="vq", dim=32, K=64, shape=(2, 40, 32), steps=2, thr=0, heads=4, cb_dim=8,
                           separate=True),
    "channel_first_img": dict(kind="vq", dim=16, K=32, shape=(2, 16, 6, 6), steps=2, thr=0,
                              channel_last=False),
    "ragged_tail": dict(kind="vq", dim=24, K=40, shape=(1, 77, 24), steps=2, thr=2, cb_scale=0.7),
    "one_vector": dict(kind="vq", dim=8, K=16, shape=(5, 8), steps=1, thr=0),
    "rvq3": dict(kind="rvq", dim=32, K=64, Q=3, shape=(2, 128, 32), steps=2, thr=0),
    "rvq4_expire": dict(kind="rvq", dim=16, K=32, Q=4, shape=(2, 64, 16), steps=2, thr=2),
    "rvq_shared": dict(kind="rvq", dim=16, K=32, Q=3, shape=(2, 64, 16), steps=2, thr=0, shared=True),
    "rvq_eval": dict(kind="rvq", dim=32, K=64, Q=3, shape=(2, 128, 32), steps=1, thr=0, training=False),
    "kmeans_vq": dict(kind="vq", dim=16, K=32, shape=(2, 200, 16), steps=2, thr=0, kmeans=True),
}


def build(cfg, VectorQuantize, ResidualVQ, CodebookParams, KmeansParameters):
    torch.manual_seed(0)
    cb_dim = cfg.get("cb_dim", cfg["dim"])
    cp = CodebookParams(
        dim=cb_dim, codebook_size=cfg["K"], threshold_ema_dead_code=cfg["thr"],
        use_cosine_sim=cfg.get("cosine", False),
        transform_input="l2norm" if cfg.get("l2in") else "identity",
        weights_regularization="l2norm" if cfg.get("l2w") else "identity",
        initialization_by_kmeans=cfg.get("kmeans", False),
        kmeans_params=KmeansParameters() if cfg.get("kmeans") else None,
    )
    if cfg["kind"] == "vq":
        mod = VectorQuantize(dim=cfg["dim"], codebook_params=cp, codebook_dim=cfg.get("cb_dim"),
                             heads=cfg.get("heads", 1),
                             separate_codebook_per_head=cfg.get("separate", False),
                             channel_last=cfg.get("channel_last", True), sync_codebook=False)
        books = [mod._codebook]
    else:
        mod = ResidualVQ(dim=cfg["dim"], num_quantizers=cfg["Q"], codebook_params=cp,
                         shared_codebook=cfg.get("shared", False), sync_codebook=False)
        books = [l._codebook for l in mod.layers]
    if "cb_scale" in cfg:
        g = torch.Generator().manual_seed(7)
        for cb in ([books[0]] if cfg.get("shared") else books):
            c = torch.randn(cb.embeddings.shape, generator=g) * cfg["cb_scale"]
            cb.embeddings.copy_(c)
            cb.embed_avg.copy_(c)
            cb.cluster_size.fill_(1.0)
    return mod, books


def run_case(name, cfg, api):
    mod, books = build(cfg, *api)
    training = cfg.get("training", True)
    mod.train(training)
    g = torch.Generator().manual_seed(1234)
    x = torch.randn(*cfg["shape"], generator=g)
    mask = None
    if cfg.get("mask"):
        b, n = cfg["shape"][:2]
        lens = torch.tensor([n, n // 2, max(1, n // 5)][:b])
        mask = torch.arange(n)[None, :] < lens[:, None]
    fx = {"cfg": cfg, "x": x, "mask": mask, "init": [_snap(b) for b in books], "steps": [],
          "rng_seed": 4321}

    # capture the similarities the reference computed (third return of Codebook.forward) for gaps
    gaps = []

    def hook(_m, _inp, out):
        sim = out[2]
        gaps.append(_gap(sim.detach()))

    handles = [b.register_forward_hook(hook) for b in ({id(b): b for b in books}.values())]
    for s in range(cfg["steps"]):
        gaps.clear()
        torch.manual_seed(fx["rng_seed"] + s)
        with torch.no_grad():
            if cfg["kind"] == "vq":
                q, ind, loss = mod(x + 0.01 * s, mask=mask)
            else:
                q, ind, loss = mod(x + 0.01 * s, mask=mask)
        fx["steps"].append({"quantize": q.clone(), "indices": ind.clone(), "loss": loss.clone(),
                            "after": [_snap(b) for b in books],
                            "top2_rel_gap": [g_.clone() for g_ in gaps]})
    for h in handles:
        h.remove()
    return fx


def main():
    api = _import_reference()
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    for name, cfg in CASES.items():
        fx = run_case(name, cfg, api)
        path = os.path.join(HERE, name + ".pt")
        torch.save(fx, path)
        print(f"{name:20s} -> {os.path.getsize(path) / 1e6:.2f} MB")


if __name__ == "__main__":
    main()
