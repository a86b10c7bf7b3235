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


# ---------------------------------------------------------------------------------------------------------------
# ResidualVQ / GroupedResidualVQ host logic (vqb200/rvq.py generic level loop: quantize-dropout, shared codebook,
# projections, masks, return_all_codes, channel-first, codes-from-indices) against the LIVE reference, the kernels
# replaced as above.  The reference's state_dict is loaded into the product module (strict).
# ---------------------------------------------------------------------------------------------------------------
def make_rvq_case(seed):
    r = random.Random(900 + seed)
    grouped = r.random() < 0.3
    groups = r.choice([2, 3]) if grouped else 1
    cb_dim = r.choice([4, 8])
    proj = (not grouped) and r.random() < 0.3
    dim = (cb_dim + 3 if proj else cb_dim) * groups
    dropout = r.random() < 0.4
    Q = r.choice([2, 3, 4])
    channel_first = (not proj) and r.random() < 0.25
    cfg = dict(grouped=grouped, groups=groups, cb_dim=cb_dim, dim=dim, proj=proj, Q=Q, K=r.choice([6, 15]),
               shared=r.random() < 0.3, dropout=dropout, cutoff=r.randrange(0, Q) if dropout else 0,
               mult=r.choice([1, 1, 2]) if dropout else 1, training=r.random() < 0.8,
               all_codes=r.random() < 0.4, channel_first=channel_first,
               mask=(not channel_first) and r.random() < 0.3, b=r.choice([1, 2]), n=r.choice([4, 7]),
               drop_seed=r.randrange(100), seed=seed)
    return cfg


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "vector_quantization")),
                    reason="the live reference is only present in the build container")
@pytest.mark.parametrize("seed", list(range(24)))
def test_rvq_host_logic_equals_live_reference_on_random_options(seed, monkeypatch):
    _reference_api()
    from vector_quantization.codebooks import CodebookParams as RefParams
    from vector_quantization.residual_vq import GroupedResidualVQ as RefGrouped, ResidualVQ as RefRVQ
    import vqb200
    _patch(monkeypatch)
    cfg = make_rvq_case(seed)

    def build(RVQ, Grouped, Params):
        cp = Params(dim=cfg["cb_dim"], codebook_size=cfg["K"], threshold_ema_dead_code=0)
        kw = dict(num_quantizers=cfg["Q"], codebook_params=cp, shared_codebook=cfg["shared"], sync_codebook=False,
                  quantize_dropout=cfg["dropout"], quantize_dropout_cutoff_index=cfg["cutoff"],
                  quantize_dropout_multiple_of=cfg["mult"])
        if cfg["grouped"]:
            return Grouped(dim=cfg["dim"], groups=cfg["groups"], channel_last=not cfg["channel_first"], **kw) \
                if not cfg["channel_first"] else \
                Grouped(dim=cfg["dim"], groups=cfg["groups"], channel_last=False, **kw)
        if cfg["proj"]:
            kw["codebook_dim"] = cfg["cb_dim"]
        if cfg["channel_first"]:
            kw["channel_last"] = False
        return RVQ(dim=cfg["dim"], **kw)

    torch.manual_seed(seed)
    ref = build(RefRVQ, RefGrouped, RefParams)
    g = torch.Generator().manual_seed(3000 + seed)
    with torch.no_grad():
        for name, buf in ref.named_buffers():
            if name.endswith("embeddings"):
                buf.copy_(torch.randn(buf.shape, generator=g) * 0.6)
        sd = ref.state_dict()
        for name in sd:
            if name.endswith("embed_avg"):
                sd[name].copy_(sd[name.replace("embed_avg", "embeddings")])
            if name.endswith("cluster_size"):
                sd[name].fill_(1.0)
    ours = build(vqb200.ResidualVQ, vqb200.GroupedResidualVQ, vqb200.CodebookParams)
    if cfg["grouped"] and cfg["channel_first"]:
        pytest.skip("GroupedResidualVQ(channel_last=False) does not forward channel_last to its ResidualVQs in the "
                    "reference: each group would see a channel-last view of channel-first data")
    ours.load_state_dict(ref.state_dict(), strict=True)
    ref.train(cfg["training"]); ours.train(cfg["training"])
    shape = (cfg["b"], cfg["dim"], cfg["n"]) if cfg["channel_first"] else (cfg["b"], cfg["n"], cfg["dim"])
    x0 = torch.randn(*shape, generator=g)
    w = torch.randn(*shape, generator=g)
    mask = (torch.rand(cfg["b"], cfg["n"], generator=g) > 0.3) if cfg["mask"] else None
    outs = []
    for mod in (ref, ours):
        x = x0.clone().requires_grad_(True)
        kw = dict(mask=mask, return_all_codes=cfg["all_codes"])
        if cfg["grouped"]:
            random.seed(cfg["drop_seed"])             # GroupedResidualVQ draws the dropout seed from python's RNG
        else:
            kw["rand_quantize_dropout_fixed_seed"] = cfg["drop_seed"]
        out = mod(x, **kw)
        out = [o if torch.is_tensor(o) else torch.stack(list(o)) for o in out]
        scalar = (out[0] * w).sum() + out[2].sum() * 1.7
        if scalar.requires_grad:
            scalar.backward()
        outs.append((out, x.grad, {k: v.detach().clone() for k, v in mod.state_dict().items()}))
    (ro, rg, rs), (oo, og, os_) = outs
    assert len(ro) == len(oo)
    for a, b in zip(ro, oo):
        assert a.shape == b.shape and a.dtype == b.dtype
    assert torch.equal(ro[1], oo[1]), "indices"
    assert torch.allclose(ro[0], oo[0], rtol=1e-6, atol=1e-7, equal_nan=True)
    assert torch.allclose(ro[2], oo[2], rtol=1e-6, atol=1e-8, equal_nan=True)
    if len(ro) > 3:
        assert torch.allclose(ro[3], oo[3], rtol=1e-6, atol=1e-7)
    assert (rg is None) == (og is None)
    if rg is not None:
        assert torch.allclose(rg, og, rtol=1e-5, atol=1e-7, equal_nan=True)
    for k in rs:
        assert torch.allclose(rs[k], os_[k], rtol=1e-6, atol=1e-7), k
    # codes / output from indices (the reference uses einx there; the stand-in of tests/golden/make_golden.py)
    if not cfg["dropout"] or not cfg["training"]:
        ind = ro[1]
        assert torch.allclose(ref.get_codes_from_indices(ind), ours.get_codes_from_indices(ind), rtol=1e-6, atol=1e-7)
        if not cfg["grouped"]:      # the reference's grouped variant reads an undefined `self.split_dim` here
            assert torch.allclose(ref.get_output_from_indices(ind), ours.get_output_from_indices(ind), rtol=1e-6,
                                  atol=1e-7)


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "vector_quantization")),
                    reason="the live reference is only present in the build container")
@pytest.mark.parametrize("seed", list(range(16)))
def test_projected_vq_host_logic_equals_live_reference_on_random_options(seed, monkeypatch):
    """VectorQuantize with project_in / project_out (+ LayerNorm) on top of the random option combinations above:
    the product (kernels replaced, reference state_dict loaded) against the live reference, incl. input gradients."""
    RefVQ, RefParams = _reference_api()
    import vqb200
    _patch(monkeypatch)
    cfg = make_case(100 + seed)
    r = random.Random(7000 + seed)
    cfg["dim"] = cfg["dim"] + r.choice([1, 3])           # codebook_input_dim != dim -> projections
    ln = r.random() < 0.5
    shape = {"series": lambda s: (s[0], s[1], cfg["dim"]), "image": lambda s: (s[0], 2, 3, cfg["dim"]),
             "channel_first": lambda s: (s[0], cfg["dim"], 3, 2), "one": lambda s: (s[0], cfg["dim"])}[cfg["kind"]]
    cfg["shape"] = shape(cfg["shape"])
    mods = []
    for VQ, Params in ((RefVQ, RefParams), (vqb200.VectorQuantize, vqb200.CodebookParams)):
        kw = module_kwargs(cfg, Params)
        kw["codebook_dim"] = cfg["cb_dim"]
        kw["layernorm_after_project_in"] = ln
        torch.manual_seed(seed)
        mods.append(VQ(**kw))
    ref, ours = mods
    emb, x0, w, mask, targets = inputs(cfg)
    with torch.no_grad():
        ref._codebook.embeddings.copy_(emb); ref._codebook.embed_avg.copy_(emb); ref._codebook.cluster_size.fill_(1.0)
    ours.load_state_dict(ref.state_dict(), strict=True)
    res = []
    for mod in (ref, ours):
        mod.train(cfg["training"])
        x = x0.clone().requires_grad_(True)
        if targets is not None:
            out = mod(x, indices=targets, freeze_codebook=cfg["freeze"])
        else:
            out = mod(x, mask=mask, freeze_codebook=cfg["freeze"])
        res.append(finish(cfg, out, x, w, mod._codebook.embeddings, mod._codebook.cluster_size))
        res[-1]["proj_grad"] = mod.project_in[0].weight.grad.clone() if ln and mod.project_in[0].weight.grad is not None \
            else (mod.project_in.weight.grad.clone() if not ln and mod.project_in.weight.grad is not None
                  else torch.zeros(1))
    compare(res[0], res[1], exact=False, cfg=cfg)
