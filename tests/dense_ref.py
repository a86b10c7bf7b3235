"""Plain-torch statement of what each entry point of csrc/dense.cu computes (include/vqb.h, "consumers of the dense
N x K similarities"), materialising N x K.  Test infrastructure: the fp64 / fp32 reference of the kernels' numerics
tests, and the stand-in for the C entry points in the CPU host-logic test."""
import torch
import torch.nn.functional as F


def sims(x, c, cosine):
    dot = torch.einsum("hnd,hkd->hnk", x, c)
    if cosine:
        return dot
    xn2, cn2 = (x * x).sum(-1), (c * c).sum(-1)
    return -(xn2[..., None] + cn2[:, None, :] - 2 * dot).clamp_min(0).sqrt()


def rowstats(x, c, cosine, alpha, target=None):
    s = sims(x, c, cosine)
    lse = torch.logsumexp(alpha * s, -1)
    st = None
    if target is not None:
        st = torch.gather(s, -1, target.clamp_min(0)[..., None])[..., 0] * (target >= 0)
    return lse, st


def _table_rows(table, H, N, n_pos):
    return table[None].expand(H * N // n_pos, n_pos, -1).reshape(H, N, -1)


def avgprob(x, c, cosine, alpha, lse, n_pos):
    p = torch.exp(alpha * sims(x, c, cosine) - lse[..., None])
    return p.reshape(-1, n_pos, p.shape[-1]).mean(0)


def rowdot(x, c, cosine, alpha, lse, table, n_pos):
    H, N, _ = x.shape
    p = torch.exp(alpha * sims(x, c, cosine) - lse[..., None])
    return (p * _table_rows(table, H, N, n_pos)).sum(-1)


def backward(x, c_dist, c_comb, cosine, alpha, lse, coef, target=None, table=None, rdot=None, n_pos=1):
    H, N, _ = x.shape
    s = sims(x, c_dist, cosine)
    p = torch.exp(alpha * s - lse[..., None])
    if table is not None:
        w = coef[..., None] * p * (_table_rows(table, H, N, n_pos) - rdot[..., None])
    else:
        onehot = F.one_hot(target.clamp_min(0), s.shape[-1]).to(s.dtype) * (target >= 0)[..., None]
        w = coef[..., None] * (p - onehot)
    if cosine:
        rho, xcoef = -w, torch.zeros(H, N, dtype=s.dtype, device=s.device)
    else:
        D = -s
        rho = torch.where(D > 0, -w / D, torch.zeros_like(D))
        xcoef = rho.sum(-1)
    return xcoef[..., None] * x - torch.einsum("hnk,hkd->hnd", rho, c_comb)


def backward_codes(x, c, cosine, alpha, lse, coef, target=None, table=None, rdot=None, n_pos=1):
    """codebook side of `backward` (ATen _euclidean_dist_backward, x2 branch / bmm): (H,K,d)."""
    H, N, _ = x.shape
    s = sims(x, c, cosine)
    p = torch.exp(alpha * s - lse[..., None])
    if table is not None:
        w = coef[..., None] * p * (_table_rows(table, H, N, n_pos) - rdot[..., None])
    else:
        onehot = F.one_hot(target.clamp_min(0), s.shape[-1]).to(s.dtype) * (target >= 0)[..., None]
        w = coef[..., None] * (p - onehot)
    if cosine:
        return torch.einsum("hnk,hnd->hkd", w, x)
    D = -s
    rho = torch.where(D > 0, -w / D, torch.zeros_like(D))
    return c * rho.sum(1)[..., None] - torch.einsum("hnk,hnd->hkd", rho, x)
