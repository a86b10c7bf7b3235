"""Host layout helpers against the einops patterns of the reference (property tests, CPU): the (H, N) row layout the
kernels work in must be the one the reference's rearranges produce."""
import torch
from einops import rearrange, repeat
from hypothesis import given, settings, strategies as st


def _vq(heads, separate):
    from vqb200 import CodebookParams, VectorQuantize
    return VectorQuantize(dim=4 * heads, codebook_dim=4, heads=heads, separate_codebook_per_head=separate,
                          codebook_params=CodebookParams(dim=4, codebook_size=8))


@settings(max_examples=40, deadline=None)
@given(b=st.integers(1, 4), n=st.integers(1, 7), heads=st.integers(1, 4), separate=st.booleans())
def test_target_rows_follow_the_reference_distance_layout(b, n, heads, separate):
    """reference vector_quantize_pytorch.py:285-290: similarities "c b n l" / "1 (b h) n l" are scored against codes
    "b n c" / "b n h"; `_rows_of` must put code (b, n, h) on the similarity row of (h, b, n)."""
    vq = _vq(heads, separate)
    multi = heads > 1
    codes = torch.arange(b * n * heads).reshape(b, n, heads) if multi else torch.arange(b * n).reshape(b, n)
    rows = vq._rows_of(codes, b, multi)
    if not multi:
        want = rearrange(codes, "b n -> 1 (b n)")
    elif separate:
        want = rearrange(codes, "b n h -> h (b n)")
    else:
        want = rearrange(codes, "b n h -> 1 (b h n)")
    assert torch.equal(rows, want)


@settings(max_examples=40, deadline=None)
@given(b=st.integers(1, 4), n=st.integers(1, 9), rep=st.integers(1, 4), seed=st.integers(0, 100))
def test_mask_expansion_matches_reference_repeat(b, n, rep, seed):
    """reference codebooks.py:361-367: repeat(mask, "b n -> c (b h n)"); the kernels take one row of it."""
    from vqb200 import Codebook
    cb = Codebook(4, 8)
    mask = torch.rand(b, n, generator=torch.Generator().manual_seed(seed)) > 0.4
    got = cb._expand_mask(mask, b * rep * n)
    want = repeat(mask, "b n -> c (b h n)", c=1, h=rep)[0]
    assert got.dtype == torch.uint8 and torch.equal(got.bool(), want)
    assert cb._expand_mask(None, 5) is None
