"""f4 variants on the GPU: gumbel sampling (vqb_dense_gumbel_sample), affine re-parametrisation (vqb_column_moments +
the usual kernels), orthogonal regularisation, materialised similarities (vqb_dense_scores) -- the product against the
15 fixtures recorded from the live reference (tests/golden/make_golden_f4.py), and the in-kernel Philox stream against
torch's own `uniform_` on the same generator state."""
import pytest
import torch

import f4_util as F
import golden_util as gu

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", F.names())
def test_f4_matches_reference_fixture(name, monkeypatch):
    F.run_and_check(F.load(name), torch.device("cuda:0"), monkeypatch)


def _torch_sample(x, emb, cos, tau, u):
    sim = torch.einsum("hnd,hcd->hnc", x, emb) if cos else -torch.cdist(x, emb)
    g = -torch.log((-torch.log(u.clamp(min=1e-5))).clamp(min=1e-5))
    logits = sim / tau + g
    top2 = logits.topk(2, dim=-1).values
    return logits.argmax(-1), (top2[..., 0] - top2[..., 1]).abs() / top2[..., 0].abs().clamp_min(1e-30)


@pytest.mark.parametrize("H,N,K,d,cos", [(1, 100, 32, 8, False), (2, 777, 129, 24, True), (1, 5000, 512, 64, False),
                                         (1, 40000, 1024, 32, False)])
def test_in_kernel_philox_is_torchs_uniform_stream(H, N, K, d, cos):
    """Same generator state -> the kernel's noise is bit for bit what `torch.zeros(H,N,K).uniform_(0, 1)` draws (sizes
    below and above the one-wave grid cap of ATen's launch), and the generator ends at the same offset."""
    from vqb200 import ops
    dev = torch.device("cuda:0")
    g0 = torch.Generator(device=dev).manual_seed(1234)
    x = torch.randn(H, N, d, generator=g0, device=dev)
    emb = torch.randn(H, K, d, generator=g0, device=dev) * 0.5
    if cos:
        x, emb = torch.nn.functional.normalize(x, dim=-1), torch.nn.functional.normalize(emb, dim=-1)
    ga = torch.Generator(device=dev).manual_seed(77)
    gb = torch.Generator(device=dev).manual_seed(77)
    torch.empty(1000, device=dev).uniform_(0, 1, generator=ga)          # a non-zero starting offset
    torch.empty(1000, device=dev).uniform_(0, 1, generator=gb)
    u = torch.zeros(H, N, K, device=dev).uniform_(0, 1, generator=ga)
    want, gap = _torch_sample(x, emb, cos, 0.7, u)
    got = ops.dense_gumbel_sample(x, emb, cos, 0.7, generator=gb)
    assert gb.get_offset() == ga.get_offset(), "the generator was not advanced like torch's uniform_"
    bad = (got != want) & (gap >= 1e-5)
    assert not bool(bad.any()), f"{int(bad.sum())} of {H * N} samples differ from torch's stream"
    assert float((got == want).float().mean()) > 0.999
    # with an explicit draw the same numbers come back
    again = ops.dense_gumbel_sample(x, emb, cos, 0.7, uniforms=u)
    assert torch.equal(again, got)


def test_gumbel_sampling_follows_softmax_probabilities():
    """One latent repeated many times: code frequencies of the stochastic sampler approach softmax(sim / tau)."""
    from vqb200 import ops
    dev = torch.device("cuda:0")
    g0 = torch.Generator(device=dev).manual_seed(5)
    K, d, n, tau = 16, 8, 200000, 0.5
    emb = torch.randn(1, K, d, generator=g0, device=dev) * 0.4
    x = (torch.randn(1, 1, d, generator=g0, device=dev) * 0.4).expand(1, n, d).contiguous()
    idx = ops.dense_gumbel_sample(x, emb, False, tau, generator=torch.Generator(device=dev).manual_seed(9))
    freq = torch.bincount(idx.reshape(-1), minlength=K).double() / n
    p = torch.softmax(-torch.cdist(x[:, :1], emb)[0, 0].double() / tau, dim=-1)
    assert float((freq - p).abs().max()) < 5e-3, (freq, p)


def test_similarities_and_column_moments():
    from vqb200 import ops
    dev = torch.device("cuda:0")
    g0 = torch.Generator(device=dev).manual_seed(3)
    x = torch.randn(2, 1000, 48, generator=g0, device=dev)
    emb = torch.randn(2, 70, 48, generator=g0, device=dev)
    for cos in (False, True):
        want = torch.einsum("hnd,hcd->hnc", x, emb) if cos else -torch.cdist(x, emb)
        assert gu.rel_err(ops.dense_scores(x, emb, cos).cpu(), want.cpu()) <= 2e-6
    mask = (torch.rand(1000, generator=g0, device=dev) > 0.4)
    for m, xx in ((None, x), (mask.to(torch.uint8), x), (None, x.bfloat16())):
        sums, rows = ops.column_moments(xx, m)
        sel = xx.double() if m is None else xx.double()[:, mask]
        assert int(rows[0]) == sel.shape[1]
        assert gu.rel_err(sums[..., 0].cpu(), sel.sum(1).cpu()) <= 1e-12
        assert gu.rel_err(sums[..., 1].cpu(), (sel * sel).sum(1).cpu()) <= 1e-12
