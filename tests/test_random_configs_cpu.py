"""Randomised differential tests on the CPU (seeded, reproducible):

A. the CPU restatement (oracle/) against the LIVE reference on random option combinations -- heads, shared / separate
   codebooks, masks, cosine + l2norm, series / image / channel-first / 2-D inputs, train / eval / freeze, and the
   dense-similarity losses -- outputs bit for bit, input gradients to 1e-6.  Needs /root/reference (present in the
   build container, absent on the GPU box: skipped there; nothing under `-m gpu` reads it).
B. the PRODUCT's host logic (vqb200/vq.py + ops.py autograd wiring) against the restatement on the same combinations,
   with the C entry points replaced by their plain-torch statements (tests/dense_ref.py) and search / gather / EMA
   by the restatement: layouts, masks, loss assembly, return conventions.
"""
import os
import random
import sys

import pytest
import torch

import golden_util as gu
from oracle import vq_oracle as O
from test_dense_host_cpu import _patch

REF = os.environ.get("VQB_REFERENCE", "/root/reference")
SEEDS = list(range(28))


def make_case(seed):
    r = random.Random(seed)
    heads = r.choice([1, 1, 2, 4])
    cb_dim = r.choice([4, 8])
    kind = r.choice(["series", "series", "image", "channel_first", "one"])
    dense = r.choice(["none", "ce_commit", "diversity", "both", "indices"])
    if kind in ("image", "channel_first") and dense in ("ce_commit", "both"):
        dense = "diversity"          # CE commitment on un-flattened indices raises inside the reference
    if kind == "one" and heads > 1:
        heads = 1
    if kind == "one" and dense in ("ce_commit", "both"):
        dense = "diversity"          # ... and so does CE commitment on a 2-D input (indices already squeezed)
    cosine = r.random() < 0.3
    training = dense in ("ce_commit", "diversity", "both") or r.random() < 0.75
    cfg = dict(heads=heads, separate=heads > 1 and r.random() < 0.5, cb_dim=cb_dim, dim=cb_dim * heads,
               K=r.choice([7, 16, 33]), kind=kind, dense=dense, cosine=cosine, l2=cosine and r.random() < 0.7,
               training=training, freeze=training and r.random() < 0.2, cw=r.choice([1.0, 0.5, 0.0]),
               dw=r.choice([0.3, 1.0]), temp=r.choice([1.0, 3.0, 100.0]),
               mask=kind == "series" and dense != "indices" and r.random() < 0.4, seed=seed)
    if dense == "ce_commit" and cfg["cw"] == 0.0:
        cfg["cw"] = 0.7
    b, n = r.choice([1, 2, 3]), r.choice([1, 5, 9])
    d = cfg["dim"]
    cfg["shape"] = {"series": (b, n, d), "image": (b, 2, 3, d), "channel_first": (b, d, 3, 2), "one": (b + 1, d)}[kind]
    return cfg


def module_kwargs(cfg, CodebookParams):
    cp = CodebookParams(dim=cfg["cb_dim"], codebook_size=cfg["K"], threshold_ema_dead_code=0,
                        use_cosine_sim=cfg["cosine"], transform_input="l2norm" if cfg["l2"] else "identity",
                        weights_regularization="l2norm" if cfg["l2"] else "identity")
    kw = dict(dim=cfg["dim"], codebook_params=cp, sync_codebook=False, heads=cfg["heads"],
              separate_codebook_per_head=cfg["separate"], channel_last=cfg["kind"] != "channel_first",
              commitment_weight=cfg["cw"], commitment_use_cross_entropy_loss=cfg["dense"] in ("ce_commit", "both"),
              codebook_diversity_loss_weight=cfg["dw"] if cfg["dense"] in ("diversity", "both") else 0.0,
              codebook_diversity_temperature=cfg["temp"])
    if cfg["heads"] > 1:
        kw["codebook_dim"] = cfg["cb_dim"]
    return kw


def inputs(cfg):
    g = torch.Generator().manual_seed(1000 + cfg["seed"])
    H = cfg["heads"] if cfg["separate"] else 1
    emb = torch.randn(H, cfg["K"], cfg["cb_dim"], generator=g) * 0.6
    if cfg["l2"]:
        emb = torch.nn.functional.normalize(emb, dim=-1)
    x = torch.randn(*cfg["shape"], generator=g)
    w = torch.randn(*cfg["shape"], generator=g)
    mask = (torch.rand(cfg["shape"][0], cfg["shape"][1], generator=g) > 0.3) if cfg["mask"] else None
    targets = None
    if cfg["dense"] == "indices":
        lead = {"series": cfg["shape"][:2], "image": (cfg["shape"][0], 6), "channel_first": (cfg["shape"][0], 6),
                "one": (cfg["shape"][0], 1)}[cfg["kind"]]
        tshape = tuple(lead) + ((cfg["heads"],) if cfg["heads"] > 1 else ())
        targets = torch.randint(0, cfg["K"], tshape, generator=g)
        targets.view(-1)[::4] = -1
    return emb, x, w, mask, targets


def oracle_opts(cfg):
    cb = O.CodebookOpts(threshold_ema_dead_code=0, use_cosine_sim=cfg["cosine"], weights_l2norm=cfg["l2"])
    return O.VQOpts(heads=cfg["heads"], separate_codebook_per_head=cfg["separate"],
                    channel_last=cfg["kind"] != "channel_first", commitment_weight=cfg["cw"], input_l2norm=cfg["l2"],
                    codebook=cb)


def run_oracle(cfg):
    emb, x, w, mask, targets = inputs(cfg)
    st = O.CodebookState(emb.clone(), emb.clone(), torch.ones(emb.shape[:2]))
    x = x.clone().requires_grad_(True)
    out = O.vq_forward_dense(st, x, oracle_opts(cfg), training=cfg["training"], mask=mask, targets=targets,
                             freeze_codebook=cfg["freeze"],
                             ce_commit=cfg["dense"] in ("ce_commit", "both"),
                             diversity_weight=cfg["dw"] if cfg["dense"] in ("diversity", "both") else 0.0,
                             diversity_temperature=cfg["temp"])
    return finish(cfg, out, x, w, st.embeddings, st.cluster_size)


def finish(cfg, out, x, w, emb_after, cs_after):
    """common scalar + backward; returns a dict of comparable tensors."""
    res = {}
    if cfg["dense"] == "indices":
        q, ce = out
        scalar = ce * 1.3 + (q.sum() * 0.01 if q.requires_grad else 0.0)
        res.update(q=q.detach(), ce=ce.detach())
    else:
        q, ind, loss = out[:3]
        scalar = (q * w).sum() + loss.sum() * 1.7
        res.update(q=q.detach(), ind=ind, loss=loss.detach())
    if torch.is_tensor(scalar) and scalar.requires_grad:
        scalar.backward()
        res["gx"] = x.grad.clone()
    res["emb"], res["cs"] = emb_after.detach().clone(), cs_after.detach().clone()
    return res


def run_module(cfg, VectorQuantize, CodebookParams):
    emb, x, w, mask, targets = inputs(cfg)
    vq = VectorQuantize(**module_kwargs(cfg, CodebookParams))
    vq.train(cfg["training"])
    cb = vq._codebook
    with torch.no_grad():
        cb.embeddings.copy_(emb); cb.embed_avg.copy_(emb); cb.cluster_size.fill_(1.0)
    x = x.clone().requires_grad_(True)
    kw = dict(freeze_codebook=cfg["freeze"])
    if targets is not None:
        out = vq(x, indices=targets, **kw)
    else:
        out = vq(x, mask=mask, **kw)
    return finish(cfg, out, x, w, cb.embeddings, cb.cluster_size)


def compare(a, b, exact, cfg):
    assert set(a) == set(b), (set(a), set(b))
    # the diversity loss multiplies the similarities by its temperature before the softmax: fp32 evaluation orders
    # (restatement vs plain-torch statement of the kernels) differ by that factor more
    scale = max(1.0, cfg["temp"]) if cfg["dense"] in ("diversity", "both") else 1.0
    for k in a:
        assert a[k].shape == b[k].shape and a[k].dtype == b[k].dtype, k
        if k in ("ind", "cs") or (exact and k in ("q", "loss", "ce")):
            assert torch.equal(a[k].nan_to_num(nan=12345.0), b[k].nan_to_num(nan=12345.0)), k     # NaN where the
            # reference gives NaN (e.g. the mse commitment of a batch whose mask selects nothing)
        else:
            assert torch.allclose(a[k], b[k], rtol=2e-5, atol=2e-6 * scale, equal_nan=True), \
                (k, float((a[k] - b[k]).abs().max()))


_REF_API = None


def _reference_api():
    global _REF_API
    if _REF_API is None:
        sys.path.insert(0, os.path.join(gu.GOLDEN_DIR))
        import make_golden as MG
        VectorQuantize, _, CodebookParams, _ = MG._import_reference()
        _REF_API = (VectorQuantize, CodebookParams)
    return _REF_API


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "vector_quantization")),
                    reason="the live reference is only present in the build container")
@pytest.mark.parametrize("seed", SEEDS)
def test_restatement_equals_live_reference_on_random_options(seed):
    cfg = make_case(seed)
    compare(run_module(cfg, *_reference_api()), run_oracle(cfg), exact=True, cfg=cfg)


@pytest.mark.parametrize("seed", SEEDS)
def test_product_host_logic_equals_restatement_on_random_options(seed, monkeypatch):
    _patch(monkeypatch)
    from vqb200 import CodebookParams, VectorQuantize
    cfg = make_case(seed)
    compare(run_module(cfg, VectorQuantize, CodebookParams), run_oracle(cfg), exact=False, cfg=cfg)


# ---------------------------------------------------------------------------------------------------------------
# learnable codebook (single head, channel-last): restatement vs live reference, incl. the CODEBOOK gradient
# ---------------------------------------------------------------------------------------------------------------
def make_learnable_case(seed):
    r = random.Random(500 + seed)
    dense = r.choice(["none", "ce_commit", "diversity", "both", "indices", "inplace"])
    cfg = dict(dim=r.choice([4, 8, 12]), K=r.choice([9, 20]), shape=(r.choice([1, 2, 3]), r.choice([3, 8])),
               dense=dense, cosine=r.random() < 0.3, mask=dense != "indices" and r.random() < 0.4,
               v=r.choice([0.0, 0.0, 0.4]), cw=r.choice([1.0, 0.3]), dw=r.choice([0.4, 1.0]),
               temp=r.choice([1.0, 4.0]), lr=r.choice([0.3, 1.5]), seed=seed)
    return cfg


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "vector_quantization")),
                    reason="the live reference is only present in the build container")
@pytest.mark.parametrize("seed", list(range(16)))
def test_learnable_restatement_equals_live_reference_on_random_options(seed):
    VectorQuantize, CodebookParams = _reference_api()
    cfg = make_learnable_case(seed)
    g = torch.Generator().manual_seed(2000 + seed)
    b, n = cfg["shape"]
    emb0 = torch.randn(1, cfg["K"], cfg["dim"], generator=g) * 0.7
    x0 = torch.randn(b, n, cfg["dim"], generator=g)
    w = torch.randn(b, n, cfg["dim"], generator=g)
    mask = (torch.rand(b, n, generator=g) > 0.3) if cfg["mask"] else None
    targets = torch.randint(0, cfg["K"], (b, n), generator=g) if cfg["dense"] == "indices" else None
    ce = cfg["dense"] in ("ce_commit", "both")
    dw = cfg["dw"] if cfg["dense"] in ("diversity", "both") else 0.0

    cp = CodebookParams(dim=cfg["dim"], codebook_size=cfg["K"], learnable_codebook=True, ema_update=False,
                        threshold_ema_dead_code=0, use_cosine_sim=cfg["cosine"])
    extra = dict(commitment_use_cross_entropy_loss=ce, codebook_diversity_loss_weight=dw,
                 codebook_diversity_temperature=cfg["temp"])
    if cfg["dense"] == "inplace":
        extra["in_place_codebook_optimizer"] = lambda p: torch.optim.SGD(p, lr=cfg["lr"])
    vq = VectorQuantize(dim=cfg["dim"], codebook_params=cp, commitment_weight=cfg["cw"], sync_update_v=cfg["v"],
                        sync_codebook=False, **extra).train()
    with torch.no_grad():
        vq._codebook.embeddings.copy_(emb0)
    x = x0.clone().requires_grad_(True)
    emb = emb0.clone().requires_grad_(True)
    xo = x0.clone().requires_grad_(True)
    okw = dict(commitment_weight=cfg["cw"], sync_update_v=cfg["v"], mask=mask, use_cosine_sim=cfg["cosine"],
               ce_commit=ce, diversity_weight=dw, diversity_temperature=cfg["temp"])
    if targets is not None:
        q, c = vq(x, indices=targets)
        (q.sum() * 0.01 + c * 1.3).backward()
        qo, co = O.vq_forward_learnable(emb, xo, targets=targets, **okw)
        (qo.sum() * 0.01 + co * 1.3).backward()
        assert torch.equal(c.detach(), co.detach())
    else:
        q, ind, loss = vq(x, mask=mask)
        ((q * w).sum() + loss.sum() * 1.7).backward()
        if cfg["dense"] == "inplace":
            qo, io, lo, _ = O.vq_forward_learnable(emb, xo, inplace_optimizer=torch.optim.SGD([emb], lr=cfg["lr"]),
                                                   **okw)
        else:
            qo, io, lo = O.vq_forward_learnable(emb, xo, **okw)
        ((qo * w).sum() + lo.sum() * 1.7).backward()
        assert torch.equal(ind, io)
        assert torch.equal(loss.detach().nan_to_num(nan=1e9), lo.detach().nan_to_num(nan=1e9))
    assert torch.equal(q.detach(), qo.detach())
    assert torch.allclose(x.grad, xo.grad, rtol=1e-6, atol=1e-9, equal_nan=True)
    assert torch.allclose(vq._codebook.embeddings.grad, emb.grad, rtol=1e-6, atol=1e-9, equal_nan=True)
    assert torch.equal(vq._codebook.embeddings.detach(), emb.detach())
