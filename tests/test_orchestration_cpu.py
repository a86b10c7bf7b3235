"""The product's orchestration -- `Codebook._run` (branch selection, EMA step order, pre-update snapshot), dead-code
expiry, kmeans init, the VectorQuantize / ResidualVQ glue around them -- on CPU, with the tensor-level wrappers of
vqb200/ops.py replaced by plain-torch statements of what each C entry point computes (tests/cpu_kernels.py), against
the fixtures recorded from the live reference.  The kernels behind those wrappers are tested on the GPU."""
import pytest
import torch

import cpu_kernels
import golden_util as gu


def _cpu_draw(num_rows, m, device):
    if num_rows >= m:
        return torch.randperm(num_rows)[:m]
    return torch.randint(0, num_rows, (m,))


@pytest.fixture()
def cpu_ops(monkeypatch):
    from vqb200 import _lib, codebook, ops
    saved = {n: getattr(ops, n) for n in ("prepare_codebook", "search", "l2norm_rows", "gather_st_loss", "ema_reduce",
                                          "ema_apply", "expire_scatter", "l2norm_prepare_supported",
                                          "quantize_ema_supported")}
    saved_guard = _lib.require_device
    cpu_kernels.install(ops, _lib)
    monkeypatch.setattr(codebook.Codebook, "_draw_rows", staticmethod(_cpu_draw))
    yield
    for n, f in saved.items():
        setattr(ops, n, f)
    _lib.require_device = saved_guard


@pytest.mark.parametrize("name", gu.fixture_names())
def test_orchestration_matches_reference_fixture(name, cpu_ops):
    from test_gpu_parity import build_module, load_state
    fx = gu.load(name)
    cfg = fx["cfg"]
    mod, books = build_module(cfg)
    with torch.no_grad():
        load_state(books, fx)
    mod.train(cfg.get("training", True))
    for s, step in enumerate(fx["steps"]):
        torch.manual_seed(fx["rng_seed"] + s)
        with torch.no_grad():
            q, ind, loss = mod(fx["x"] + 0.01 * s, mask=fx["mask"])
        assert torch.equal(ind, step["indices"]), f"{name} step {s}: indices"
        assert ind.dtype == torch.int64 and q.dtype == torch.float32 and loss.shape == step["loss"].shape
        assert gu.rel_err(q, step["quantize"]) <= 1e-6
        assert torch.allclose(loss, step["loss"], rtol=1e-5, atol=1e-12)
        for cb, after in zip(books, step["after"]):
            assert torch.equal(cb.cluster_size, after["cluster_size"])
            assert gu.rel_err(cb.embed_avg, after["embed_avg"]) <= 1e-5
            assert gu.rel_err(cb.embeddings, after["embeddings"]) <= 1e-5


def test_device_guard_is_back_after_the_stand_ins():
    from vqb200 import CodebookParams, VectorQuantize
    vq = VectorQuantize(dim=8, codebook_params=CodebookParams(dim=8, codebook_size=4))
    with pytest.raises(RuntimeError, match="CUDA"):
        vq(torch.randn(1, 3, 8))
