"""GPU value-parity for the API variants pinned by tests/golden/extras/ (recorded from the live reference): projections
(+ LayerNorm), freeze_codebook, video input, bf16 input, eval with a mask, ResidualVQ with projections /
quantize-dropout / return_all_codes / channel-first images, GroupedResidualVQ.  The module is built from the same
constructor arguments, loads the REFERENCE's state_dict, runs the same forward."""
import random

import pytest
import torch

import golden_util as gu

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", gu.extras_fixture_names())
def test_api_variant_matches_reference_fixture(name):
    dev = torch.device("cuda:0")
    fx = gu.load_extras(name)
    cfg = fx["cfg"]
    mod = gu.build_from_description(cfg)
    mod.load_state_dict(fx["state_dict"], strict=True)
    mod = mod.to(dev)
    mod.train(cfg.get("train", True))
    fwd = dict(cfg.get("fwd", {}))
    if fx["mask"] is not None:
        fwd["mask"] = fx["mask"].to(dev)
    if "py_seed" in cfg:
        random.seed(cfg["py_seed"])
    with torch.no_grad():
        out = mod(fx["x"].to(dev), **fwd)
    out = [(o if torch.is_tensor(o) else torch.stack(list(o))).cpu() for o in out]
    ref = fx["out"]
    assert len(out) == len(ref)
    for i, (o, r) in enumerate(zip(out, ref)):
        assert o.shape == r.shape and o.dtype == r.dtype, (i, o.shape, r.shape, o.dtype, r.dtype)
    assert torch.equal(out[1], ref[1]), f"{name}: indices differ in {int((out[1] != ref[1]).sum())} places"
    # projections / LayerNorm run in torch on the device: agreement with the CPU run is to rounding, not bitwise
    assert gu.rel_err(out[0], ref[0]) <= 1e-5, gu.rel_err(out[0], ref[0])
    assert torch.allclose(out[2], ref[2], rtol=1e-5, atol=1e-7), (out[2], ref[2])
    if len(ref) > 3:
        assert gu.rel_err(out[3], ref[3]) <= 1e-5
    after = {k: v.cpu() for k, v in mod.state_dict().items()}
    assert set(after) == set(fx["state_dict_after"])
    for k, r in fx["state_dict_after"].items():
        if k.endswith("cluster_size"):
            assert torch.equal(after[k], r), k
        else:
            assert gu.rel_err(after[k], r) <= 1e-5, (k, gu.rel_err(after[k], r))
    if cfg.get("fwd", {}).get("freeze_codebook") or not cfg.get("train", True):
        for k, r in fx["state_dict"].items():
            assert torch.equal(after[k], r), f"{k} moved although the codebook was frozen / in eval mode"
