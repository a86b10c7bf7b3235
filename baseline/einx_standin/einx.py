"""Stand-in for the one `einx` call the reference makes (residual_vq.py:117:
`get_at("q [c] d, b n q -> q b n d", codebooks, indices)`); einx is not installable in this image (no network)."""
import torch


def get_at(pattern, codebooks, indices):
    assert pattern.replace(" ", "") == "q[c]d,bnq->qbnd", pattern
    return torch.stack([codebooks[i][indices[..., i]] for i in range(codebooks.shape[0])], 0)
