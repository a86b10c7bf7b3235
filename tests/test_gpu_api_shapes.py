"""API-compat suite on CUDA tensors, modelled on the scenarios of the reference's own shape tests
(reference tests/test_vector_quantize_pytorch.py, tests/test_residual_vq.py, fixtures in tests/conftest.py:
series (1,100,d), image (1,8,8,d), video (1,10,8,8,d), channel-last and channel-first)."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
DIM = 4


def vectors(dim, channel_last=True):
    shapes = {"series": (1, 100), "image": (1, 8, 8), "video": (1, 10, 8, 8)}
    out = {}
    for k, s in shapes.items():
        t = torch.randn(*s, dim, device=DEV)
        if not channel_last:
            t = t.movedim(-1, 1).contiguous()
        out[k] = (t, s)
    return out


def check(q, ind, loss, feats, idx_shape, heads=1):
    assert q.shape == feats.shape and q.dtype == torch.float32
    assert tuple(ind.shape) == (tuple(idx_shape) + ((heads,) if heads > 1 else ()))
    assert ind.dtype == torch.int64 and torch.isfinite(q).all() and torch.isfinite(loss).all()


def make_vq(**kw):
    from vqb200 import CodebookParams, KmeansParameters, VectorQuantize
    cb = dict(dim=kw.pop("cb_dim", DIM), codebook_size=32)
    for k in ("use_cosine_sim", "initialization_by_kmeans", "kmeans_params", "transform_input", "weights_regularization"):
        if k in kw:
            cb[k] = kw.pop(k)
    return VectorQuantize(dim=kw.pop("dim", DIM), codebook_params=CodebookParams(**cb), **kw).to(DEV)


@pytest.mark.parametrize("kind", ["series", "image", "video"])
@pytest.mark.parametrize("variant", ["default", "channel_first", "cosine", "cosine_l2", "heads_separate", "heads_shared",
                                     "lower_codebook_dim", "kmeans", "kmeans_cosine", "kmeans_heads"])
def test_vector_quantize_variants(kind, variant):
    from vqb200 import KmeansParameters
    torch.manual_seed(0)
    heads, channel_last = 1, True
    if variant == "default":
        vq = make_vq()
    elif variant == "channel_first":
        vq, channel_last = make_vq(channel_last=False), False
    elif variant == "cosine":
        vq = make_vq(use_cosine_sim=True)
    elif variant == "cosine_l2":
        vq = make_vq(use_cosine_sim=True, transform_input="l2norm", weights_regularization="l2norm")
    elif variant == "heads_separate":
        heads = 2
        vq = make_vq(dim=8, cb_dim=4, codebook_dim=4, heads=2, separate_codebook_per_head=True)
    elif variant == "heads_shared":
        heads = 2
        vq = make_vq(dim=8, cb_dim=4, codebook_dim=4, heads=2)
    elif variant == "lower_codebook_dim":
        vq = make_vq(dim=8, cb_dim=4, codebook_dim=4)
    elif variant == "kmeans":
        vq = make_vq(initialization_by_kmeans=True, kmeans_params=KmeansParameters())
    elif variant == "kmeans_cosine":
        vq = make_vq(initialization_by_kmeans=True, kmeans_params=KmeansParameters(), use_cosine_sim=True)
    else:
        heads = 2
        vq = make_vq(dim=8, cb_dim=4, codebook_dim=4, heads=2, separate_codebook_per_head=True,
                     initialization_by_kmeans=True, kmeans_params=KmeansParameters())
    feats, idx_shape = vectors(vq.dim, channel_last)[kind]
    for mode in (True, True, False):                 # two training steps (EMA, expiry), one eval
        vq.train(mode)
        q, ind, loss = vq(feats)
        check(q, ind, loss, feats, idx_shape, heads)
        assert int(ind.min()) >= 0 and int(ind.max()) < 32
    q, ind, loss, bd = vq(feats, return_loss_breakdown=True)
    assert hasattr(bd, "commitment")


def test_kmeans_with_fewer_samples_than_codes():
    from vqb200 import CodebookParams, KmeansParameters, VectorQuantize
    vq = VectorQuantize(dim=DIM, codebook_params=CodebookParams(dim=DIM, codebook_size=256, initialization_by_kmeans=True,
                                                                kmeans_params=KmeansParameters())).to(DEV)
    x = torch.randn(1, 100, DIM, device=DEV)
    q, ind, loss = vq(x)
    check(q, ind, loss, x, (1, 100))


def test_backward_through_straight_through_and_commitment():
    vq = make_vq()
    x = torch.randn(2, 50, DIM, device=DEV, requires_grad=True)
    vq.train()
    before = vq.codebook.clone()
    q, ind, loss = vq(x)
    (q.sum() + loss.sum()).backward()
    assert x.grad is not None and x.grad.shape == x.shape and torch.isfinite(x.grad).all()
    assert not torch.equal(vq.codebook, before)          # the EMA step moved the codebook before backward ran
    # d(sum q)/dx = 1 (straight-through) plus the commitment term 2 (x - c) / numel with the codes the loss was
    # computed on: the PRE-update codebook (reference: commit_quantize is materialised before the EMA step)
    c = before[ind]
    ref = torch.ones_like(x) + 2 * (x.detach() - c) / x.numel()
    assert torch.allclose(x.grad, ref, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("variant", ["rvq3", "shared", "kmeans", "grouped", "rvq4_kmeans_long", "all_codes", "dropout"])
def test_residual_vq_variants(variant):
    from vqb200 import CodebookParams, GroupedResidualVQ, KmeansParameters, ResidualVQ
    torch.manual_seed(0)
    cp = CodebookParams(dim=DIM, codebook_size=32)
    cpk = CodebookParams(dim=DIM, codebook_size=32, initialization_by_kmeans=True, kmeans_params=KmeansParameters())
    x = torch.randn(1, 100, DIM, device=DEV)
    if variant == "rvq3":
        m, Q = ResidualVQ(dim=DIM, num_quantizers=3, codebook_params=cp), 3
    elif variant == "shared":
        m, Q = ResidualVQ(dim=DIM, num_quantizers=3, shared_codebook=True, codebook_params=cp), 3
    elif variant == "kmeans":
        m, Q = ResidualVQ(dim=DIM, num_quantizers=3, codebook_params=cpk), 3
    elif variant == "rvq4_kmeans_long":
        m, Q = ResidualVQ(dim=DIM, num_quantizers=4, codebook_params=cpk), 4
        x = torch.randn(1, 1024, DIM, device=DEV)
    elif variant == "all_codes":
        m, Q = ResidualVQ(dim=DIM, num_quantizers=3, codebook_params=cp), 3
    elif variant == "dropout":
        m, Q = ResidualVQ(dim=DIM, num_quantizers=4, quantize_dropout=True, codebook_params=cp), 4
    else:
        g = GroupedResidualVQ(dim=8, groups=2, num_quantizers=3, codebook_params=cp).to(DEV)
        xg = torch.randn(1, 100, 8, device=DEV)
        q, ind, loss = g(xg)
        assert q.shape == xg.shape and tuple(ind.shape) == (2, 1, 100, 3) and tuple(loss.shape) == (2, 1, 3)
        return
    m = m.to(DEV)
    for mode in (True, False):
        m.train(mode)
        if variant == "all_codes":
            q, ind, loss, codes = m(x, return_all_codes=True)
            assert tuple(codes.shape) == (Q, *x.shape)
            if not mode:
                assert torch.allclose(codes.sum(0), q, atol=1e-5)
        else:
            q, ind, loss = m(x)
        assert q.shape == x.shape and tuple(ind.shape) == (*x.shape[:2], Q) and tuple(loss.shape) == (1, Q)
        assert ind.dtype == torch.int64
    assert m.codebooks.shape == (Q, 32, DIM)
    out = m.get_output_from_indices(ind)
    assert out.shape == x.shape
