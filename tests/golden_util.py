"""Shared helpers: load a golden fixture and map its cfg onto oracle opts / vqb200 modules."""
import glob
import os

import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def fixture_names():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.pt")))


def load(name):
    return torch.load(os.path.join(GOLDEN_DIR, name + ".pt"), weights_only=False)


def oracle_opts(cfg):
    from oracle.vq_oracle import CodebookOpts, VQOpts
    cb = CodebookOpts(threshold_ema_dead_code=cfg["thr"], use_cosine_sim=cfg.get("cosine", False),
                      weights_l2norm=bool(cfg.get("l2w")))
    return VQOpts(heads=cfg.get("heads", 1), separate_codebook_per_head=cfg.get("separate", False),
                  channel_last=cfg.get("channel_last", True), input_l2norm=bool(cfg.get("l2in")), codebook=cb)


def oracle_states(fx):
    from oracle.vq_oracle import CodebookState
    cfg = fx["cfg"]
    sts = [CodebookState(s["embeddings"].clone(), s["embed_avg"].clone(), s["cluster_size"].clone(),
                         is_initialized=not cfg.get("kmeans", False)) for s in fx["init"]]
    if cfg.get("shared"):
        sts = [sts[0]] * len(sts)
    return sts


def rel_err(a, b):
    """max |a-b| relative to max |b| (buffer-level relative error)."""
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def dense_fixture_names():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "dense", "*.pt")))


def load_dense(name):
    return torch.load(os.path.join(GOLDEN_DIR, "dense", name + ".pt"), weights_only=False)


def dense_oracle_opts(cfg):
    """cfg schema of tests/golden/make_golden_dense.py -> oracle VQOpts."""
    from oracle.vq_oracle import CodebookOpts, VQOpts
    cb = CodebookOpts(threshold_ema_dead_code=0, use_cosine_sim=cfg.get("cosine", False),
                      weights_l2norm=bool(cfg.get("l2w")))
    return VQOpts(heads=cfg.get("heads", 1), separate_codebook_per_head=cfg.get("separate", False),
                  channel_last=cfg.get("channel_last", True), commitment_weight=cfg.get("cw", 1.0),
                  input_l2norm=bool(cfg.get("l2in")), codebook=cb)


def dense_kwargs(cfg):
    """the dense-consumer switches of a make_golden_dense.py case (oracle keyword names)."""
    return dict(ce_commit=cfg["kind"] == "commit" or cfg.get("ce_commit", False),
                diversity_weight=cfg.get("dw", 0.0), diversity_temperature=cfg.get("temp", 100.0))


def grad_fixture_names():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "grads", "*.pt")))


def load_grad(name):
    return torch.load(os.path.join(GOLDEN_DIR, "grads", name + ".pt"), weights_only=False)


def extras_fixture_names():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "extras", "*.pt")))


def load_extras(name):
    return torch.load(os.path.join(GOLDEN_DIR, "extras", name + ".pt"), weights_only=False)


def build_from_description(cfg):
    """vqb200 module from the constructor description stored in a tests/golden/make_golden_extras.py fixture."""
    import vqb200
    return getattr(vqb200, cfg["cls"])(codebook_params=vqb200.CodebookParams(**cfg["cp"]), sync_codebook=False,
                                       **cfg["kw"])


def rvq_learnable_fixture_names():
    return sorted(os.path.splitext(os.path.basename(p))[0]
                  for p in glob.glob(os.path.join(GOLDEN_DIR, "rvq_learnable", "*.pt")))


def check_rvq_learnable(name, device):
    """ResidualVQ over learnable codebooks (tests/golden/make_golden_rvq_learnable.py): outputs, per-level losses and
    the gradient every level's codebook Parameter receives -- also when the input carries no gradient."""
    from vqb200 import CodebookParams, ResidualVQ
    fx = torch.load(os.path.join(GOLDEN_DIR, "rvq_learnable", name + ".pt"), weights_only=False)
    cfg = fx["cfg"]
    cp = CodebookParams(dim=cfg["dim"], codebook_size=cfg["K"], learnable_codebook=True, ema_update=False,
                        threshold_ema_dead_code=0)
    rvq = ResidualVQ(dim=cfg["dim"], num_quantizers=cfg["Q"], codebook_params=cp, commitment_weight=cfg["cw"],
                     sync_codebook=False).to(device).train()
    with torch.no_grad():
        for layer, e in zip(rvq.layers, fx["init_embeddings"]):
            layer._codebook.embeddings.copy_(e)
            layer._codebook.invalidate_cache()
    x = fx["x"].to(device)
    if cfg["x_grad"]:
        x.requires_grad_(True)
    mask = fx["mask"].to(device) if fx["mask"] is not None else None
    q, ind, losses = rvq(x, mask=mask)
    scale = torch.arange(1, cfg["Q"] + 1, device=device)
    ((q * fx["w"].to(device)).sum() * (1.0 if cfg["x_grad"] else 0.0) + (losses * scale).sum()).backward()
    assert torch.equal(ind.cpu(), fx["indices"])
    assert rel_err(q.detach().cpu(), fx["quantize"]) <= 1e-6
    assert torch.allclose(losses.detach().cpu(), fx["losses"], rtol=1e-5)
    for li, (layer, ge) in enumerate(zip(rvq.layers, fx["grad_embeddings"])):
        got = layer._codebook.embeddings.grad
        assert got is not None, f"{name}: level {li} codebook received no gradient"
        assert rel_err(got.cpu(), ge) <= 1e-5, f"{name}: level {li} codebook gradient"
    if cfg["x_grad"]:
        assert rel_err(x.grad.cpu(), fx["grad_x"]) <= 1e-5
