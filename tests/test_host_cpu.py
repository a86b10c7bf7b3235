"""CPU-only checks: the C-ABI library loads and exports every symbol include/vqb.h declares (no compute calls),
the module API mirrors the reference (signatures, buffers, state_dict keys), unsupported options and CPU tensors
fail loudly (there is no fallback)."""
import inspect
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "vqb.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vqb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_header_symbol():
    from vqb200 import _lib
    handle = _lib.lib()
    names = _header_symbols()
    assert len(names) >= 18
    for n in names:
        assert hasattr(handle, n), f"libvqb200.so does not export {n}"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature"
    assert sorted(_lib.SIGNATURES) == names
    assert handle.vqb_version() == 100


def test_size_helpers_and_argument_errors_without_gpu():
    from vqb200 import _lib
    h = _lib.lib()
    assert h.vqb_codebook_cache_bytes(1, 8192, 256) >= 8192 * 256 * 2
    assert h.vqb_search_workspace_bytes(1, 1024, 512, 256) > 1024 * 256 * 2
    assert h.vqb_ema_workspace_bytes(1, 1024, 512, 256) >= 512 * 256 * 8
    assert h.vqb_gather_workspace_bytes(1, 1024, 256) > 0
    # argument validation happens before any CUDA call
    rc = h.vqb_search(None, 0, None, None, 0, 1, 1, 1, 1, 0, None, None, 0, None, 0, None)
    assert rc == -1 and b"null" in h.vqb_last_error()
    with pytest.raises(ValueError):
        _lib.check(rc, "vqb_search")


def test_signatures_match_reference_api():
    from vqb200 import Codebook, ResidualVQ, VectorQuantize
    cb = list(inspect.signature(Codebook.__init__).parameters)
    assert cb[1:] == ["dim", "codebook_size", "num_codebooks", "initialization_by_kmeans", "kmeans_params", "decay",
                      "eps_for_smoothing", "threshold_ema_dead_code", "reset_cluster_size", "use_ddp",
                      "distributed_replace_codes", "learnable_codebook", "gumbel_params", "ema_update", "use_affine",
                      "affine_params", "transform_input", "use_cosine_sim", "weights_regularization"]
    vq = list(inspect.signature(VectorQuantize.__init__).parameters)
    assert vq[1:] == ["dim", "codebook_params", "codebook_dim", "heads", "separate_codebook_per_head",
                      "layernorm_after_project_in", "channel_last", "commitment_weight",
                      "commitment_use_cross_entropy_loss", "orthogonal_reg_weight", "orthogonal_reg_active_codes_only",
                      "orthogonal_reg_max_codes", "codebook_diversity_loss_weight", "codebook_diversity_temperature",
                      "sync_codebook", "in_place_codebook_optimizer", "sync_update_v"]
    assert list(inspect.signature(VectorQuantize.forward).parameters)[1:] == \
        ["x", "indices", "mask", "freeze_codebook", "return_loss_breakdown"]
    assert list(inspect.signature(ResidualVQ.forward).parameters)[1:] == \
        ["x", "mask", "indices", "return_all_codes", "freeze_codebook", "rand_quantize_dropout_fixed_seed"]
    assert list(inspect.signature(Codebook.forward).parameters)[1:] == ["x", "mask", "freeze_codebook"]


def test_buffers_and_state_dict_keys_match_reference():
    from vqb200 import CodebookParams, ResidualVQ, VectorQuantize
    vq = VectorQuantize(dim=256, codebook_params=CodebookParams(dim=256, codebook_size=512))
    sd = vq.state_dict()
    assert list(sd) == ["_codebook.cluster_size", "_codebook.embed_avg", "_codebook.embeddings"]
    assert sd["_codebook.embeddings"].shape == (1, 512, 256) and sd["_codebook.embeddings"].dtype == torch.float32
    assert sd["_codebook.cluster_size"].shape == (1, 512) and float(sd["_codebook.cluster_size"].abs().sum()) == 0.0
    assert torch.equal(sd["_codebook.embed_avg"], sd["_codebook.embeddings"])
    bound = (6.0 / (512 * 256)) ** 0.5            # kaiming-uniform on (1,K,d): fan_in = K*d
    assert float(sd["_codebook.embeddings"].abs().max()) <= bound
    rvq = ResidualVQ(dim=32, num_quantizers=3, codebook_params=CodebookParams(dim=32, codebook_size=64))
    assert rvq.codebooks.shape == (3, 64, 32)
    assert "layers.2._codebook.embed_avg" in rvq.state_dict()
    shared = ResidualVQ(dim=32, num_quantizers=3, shared_codebook=True,
                        codebook_params=CodebookParams(dim=32, codebook_size=64))
    assert shared.layers[0]._codebook is shared.layers[2]._codebook
    mh = VectorQuantize(dim=32, heads=4, codebook_dim=8, separate_codebook_per_head=True,
                        codebook_params=CodebookParams(dim=8, codebook_size=16))
    assert mh._codebook.embeddings.shape == (4, 16, 8)
    cos = VectorQuantize(dim=16, codebook_params=CodebookParams(dim=16, codebook_size=8, use_cosine_sim=True,
                                                                weights_regularization="l2norm"))
    assert torch.allclose(cos._codebook.embeddings.norm(dim=-1), torch.ones(1, 8), atol=1e-6)


def test_no_cpu_fallback_and_unsupported_options_fail_loudly():
    from vqb200 import Codebook, CodebookParams, GumbelParams, VectorQuantize
    vq = VectorQuantize(dim=8, codebook_params=CodebookParams(dim=8, codebook_size=4))
    with pytest.raises(RuntimeError, match="CUDA"):
        vq(torch.randn(1, 3, 8))
    cb = Codebook(8, 4, learnable_codebook=True, ema_update=False)       # supported: embeddings is a Parameter
    assert isinstance(cb.embeddings, torch.nn.Parameter) and "embeddings" in cb.state_dict()
    with pytest.raises(ValueError):
        Codebook(8, 4, use_affine=True)            # needs affine_params (the reference dereferences None later)
    from vqb200.params import AffineParameters
    aff = Codebook(8, 4, use_affine=True, affine_params=AffineParameters(sync=False))
    assert {"codebook_mean", "codebook_variance", "codebook_mean_needs_init"} <= set(aff.state_dict())
    cbg = Codebook(8, 4, gumbel_params=GumbelParams(stochastic=True, temperature=0.5))
    assert cbg._variants_active() and cbg._variant_sampling() == (True, False)
    assert not Codebook(8, 4, gumbel_params=GumbelParams(stochastic=True, temperature=0.0))._variants_active()
    with pytest.raises(RuntimeError, match="CUDA"):  # the variants have no CPU implementation either
        cbg(torch.randn(1, 3, 8))
    with pytest.raises(ValueError):
        Codebook(8, 4, transform_input="tanh")
    vqorth = VectorQuantize(dim=8, codebook_params=CodebookParams(dim=8, codebook_size=4), orthogonal_reg_weight=1.0)
    # reference :95-102: the orthogonal loss makes the codebook a Parameter; the EMA update stays on
    assert isinstance(vqorth._codebook.embeddings, torch.nn.Parameter) and vqorth._codebook.ema_update
    assert not vqorth._codebook.commit_grad_to_codebook
    # in_place_codebook_optimizer: a factory over the codebook's parameters, so it needs a learnable codebook
    # (with the default EMA codebook torch raises on the empty parameter list, as in the reference)
    with pytest.raises(ValueError):
        VectorQuantize(dim=8, codebook_params=CodebookParams(dim=8, codebook_size=4),
                       in_place_codebook_optimizer=lambda p: torch.optim.SGD(p, lr=0.1))
    vqo = VectorQuantize(dim=8, codebook_params=CodebookParams(dim=8, codebook_size=4, learnable_codebook=True,
                                                               ema_update=False),
                         in_place_codebook_optimizer=lambda p: torch.optim.SGD(p, lr=0.1))
    assert isinstance(vqo.in_place_codebook_optimizer, torch.optim.SGD)
    # the consumers of the dense similarities are built (csrc/dense.cu), with a learnable codebook too
    for kw in (dict(codebook_diversity_loss_weight=0.1), dict(commitment_use_cross_entropy_loss=True)):
        VectorQuantize(dim=8, codebook_params=CodebookParams(dim=8, codebook_size=4), **kw)
        VectorQuantize(dim=8, codebook_params=CodebookParams(dim=8, codebook_size=4, learnable_codebook=True,
                                                             ema_update=False), **kw)
    with pytest.raises(AssertionError):          # reference: sync_update_v needs a learnable codebook
        VectorQuantize(dim=8, codebook_params=CodebookParams(dim=8, codebook_size=4), sync_update_v=0.5)
    with pytest.raises(AssertionError):          # reference: learnable codebook is not compatible with the EMA update
        VectorQuantize(dim=8, codebook_params=CodebookParams(dim=8, codebook_size=4, learnable_codebook=True))


def test_product_code_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "vector-quantization-by-ml_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text, \
                    f"{f} mentions the oracle; the product path must not depend on it"


@pytest.mark.parametrize("name", __import__("golden_util").extras_fixture_names())
def test_reference_checkpoints_load_strictly(name):
    """state_dict keys / shapes of the reference's modules (projections, LayerNorm, per-level and per-group codebooks)
    load into the vqb200 modules built from the same constructor arguments, strict=True, and round-trip."""
    import golden_util as gu
    fx = gu.load_extras(name)
    mod = gu.build_from_description(fx["cfg"])
    missing = mod.load_state_dict(fx["state_dict"], strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    sd = mod.state_dict()
    assert set(sd) == set(fx["state_dict"])
    for k, v in fx["state_dict"].items():
        assert torch.equal(sd[k], v), k


def test_c_abi_rejects_bad_arguments_before_touching_a_device():
    """Error convention of include/vqb.h: negative VQB_ERR_* code, message via vqb_last_error(), no exception, no exit --
    and argument validation comes before any CUDA call, so it can be exercised here without a GPU."""
    from vqb200 import _lib as L
    lib = L.lib()
    P = 1 << 12     # a non-null dummy pointer: validation must fail before anything dereferences it

    def err():
        return lib.vqb_last_error().decode()

    assert lib.vqb_dense_rowstats(0, 0, 0, 0, 0, 0, 1.0, 0, 0, 0, 0, 10, 4, 4, 0) == -1 and "null" in err()
    assert lib.vqb_dense_rowstats(P, 0, 0, P, 0, 7, 1.0, 0, P, 0, 1, 10, 4, 4, 0) == -1 and "metric" in err()
    assert lib.vqb_dense_rowstats(P, 9, P, P, P, 0, 1.0, 0, P, 0, 1, 10, 4, 4, 0) == -1 and "dtype" in err()
    assert lib.vqb_dense_rowstats(P, 0, 0, P, 0, 0, 1.0, 0, P, 0, 1, 10, 4, 4, 0) == -1 and "norms" in err()
    assert lib.vqb_dense_backward(P, 0, P, P, P, P, 0, 1.0, P, P, P, P, P, 1, P, 1, 10, 4, 4, 0) == -1 \
        and "exactly one" in err()
    assert lib.vqb_dense_avgprob(P, 0, P, P, P, 0, 1.0, P, P, 3, 1, 10, 4, 4, 0) == -1 and "multiple" in err()
    assert lib.vqb_dense_backward_codes(P, 0, P, P, P, 0, 1.0, P, P, P, 0, 0, 1, P, 99, 1, 10, 4, 4, 0) == -1 \
        and "n_splits" in err()
    assert lib.vqb_search(0, 0, 0, 0, 0, 1, 10, 4, 4, 0, 0, 0, 0, 0, 0, 0) == -1 and "vqb_search" in err()
    assert lib.vqb_ema_reduce(0, 0, 0, 0, 0, 1, 10, 4, 4, 0, 0, 0, 0) == -1
    assert lib.vqb_gather_st_loss(0, 0, 0, 0, 0, 1, 1, 0, 0, 1, 10, 4, 4, 0, 0, 0) == -1
    assert lib.vqb_prepare_codebook(0, 1, 4, 4, 0, 0, 0, 0) == -1
    assert lib.vqb_minkey_pack(0, 0, 5, 0, 0) == -1
    # the Python wrapper maps the codes to exceptions
    with pytest.raises(ValueError, match="metric"):
        L.check(lib.vqb_dense_rowstats(P, 0, 0, P, 0, 7, 1.0, 0, P, 0, 1, 10, 4, 4, 0), "vqb_dense_rowstats")
    # size helpers are pure functions of the shape
    assert lib.vqb_search_workspace_bytes(1, 1 << 20, 8192, 256) > (1 << 20) * 256 * 2
    assert lib.vqb_dense_backward_codes_splits(1, 1 << 20, 8192, 256) >= 1
    assert lib.vqb_dense_backward_codes_splits(1, 10, 64, 64) == 1          # never more splits than latent tiles
