"""Plain-torch CPU stand-ins for the tensor-level wrappers of vqb200/ops.py (each states what the C entry point behind
it computes, include/vqb.h).  TEST INFRASTRUCTURE: they let the orchestration in vqb200/codebook.py -- branch selection
in `_run`, the packed statistics all_reduce, dead-code expiry and kmeans init under data parallel -- run in the
world_size-2 Gloo tests on CPU.  The kernels themselves are tested on the GPU."""
import torch
import torch.nn.functional as F


def _sims(x, emb, cosine):
    x = x.float()
    return torch.einsum("hnd,hkd->hnk", x, emb) if cosine else -torch.cdist(x, emb)


def prepare_codebook(embeddings, use_cosine_sim, out=None):
    return torch.zeros(8, dtype=torch.uint8)


def search(x, embeddings, cache, use_cosine_sim, *, idx_offset=0, want_score=False, latents_prepared=False,
           force_exact=False):
    s = _sims(x, embeddings, use_cosine_sim)
    best, idx = s.max(-1)
    return idx + idx_offset, (-best if want_score else None), torch.zeros(8, dtype=torch.uint8)


def l2norm_rows(x):
    return F.normalize(x.float(), p=2, dim=-1)


def gather_st_loss(x, embeddings, idx, mask_u8, training, want_loss):
    H, N, d = x.shape
    xf = x.float()
    c = torch.stack([embeddings[h][idx[h]] for h in range(H)], 0)
    q = xf + (c - xf) if training else c
    loss = None
    if want_loss:
        keep = torch.ones(N, dtype=torch.bool) if mask_u8 is None else mask_u8.bool()
        rows = int(keep.sum()) * H
        err = ((c - xf) ** 2)[:, keep]
        loss = torch.stack([err.mean() if rows else torch.tensor(float("nan")), torch.tensor(float(rows))])
    return q, loss


def st_commit_backward(grad_q, g, x, embeddings, idx, mask_u8):
    H, N, d = x.shape
    c = torch.stack([embeddings[h][idx[h]] for h in range(H)], 0)
    keep = torch.ones(N, dtype=torch.bool) if mask_u8 is None else mask_u8.bool()
    return torch.where(keep[None, :, None], grad_q + g[0] * (x.float() - c), grad_q)     # masked rows: grad_q only


def ema_reduce(x, idx, mask_u8, K, bound_ws=None):
    H, N, d = x.shape
    onehot = F.one_hot(idx, K).float()
    if mask_u8 is not None:
        onehot = onehot * mask_u8.bool()[None, :, None]
    sums = torch.einsum("hnd,hnk->hkd", x.float(), onehot)
    return torch.cat([sums, onehot.sum(1)[..., None]], -1).contiguous()


def ema_apply(stats, cluster_size, embed_avg, embeddings, weight, eps, weights_l2norm):
    K = stats.shape[1]
    cluster_size.lerp_(stats[..., -1], weight)
    embed_avg.lerp_(stats[..., :-1], weight)
    total = cluster_size.sum(-1, keepdim=True)
    smoothed = (cluster_size + eps) / (total + K * eps) * total
    new = embed_avg / smoothed[..., None]
    embeddings.copy_(F.normalize(new, dim=-1) if weights_l2norm else new)


def expire_scatter(x_rows, sample_rows, threshold, reset, weights_l2norm, cluster_size, embed_avg, embeddings):
    dead = cluster_size < threshold
    picked = x_rows[sample_rows.to(torch.int64)].float()
    if weights_l2norm:
        picked = F.normalize(picked, dim=-1)
    embeddings[dead] = picked
    cluster_size[dead] = reset
    embed_avg[dead] = picked * reset


def rvq_level(residual, residual_next, embeddings, idx, mask_u8, training, first_level, quantized_out, next_cache,
              q_out=None):
    """One ResidualVQ level on (N,d) fp32 residuals (vqb_rvq_level): gather + straight-through + loss +
    residual_next = residual - q + quantized_out (+)= q; masked-out rows pass the residual through (q = residual)."""
    r = residual
    c = embeddings[idx]
    q = r + (c - r) if training else c
    keep = torch.ones(r.shape[0], dtype=torch.bool) if mask_u8 is None else mask_u8.bool()
    q = torch.where(keep[:, None], q, r)
    residual_next.copy_(r - q)
    if quantized_out is not None:
        quantized_out.copy_(0.0 + q if first_level else quantized_out + q)
    if q_out is not None:
        q_out.copy_(q)
    rows = int(keep.sum())
    err = ((c - r) ** 2)[keep]
    return torch.stack([err.mean() if rows else torch.tensor(float("nan")), torch.tensor(float(rows))])


def rvq_level_ema(residual, residual_next, embeddings, idx, training, first_level, quantized_out, next_cache,
                  bound_ws=None, q_out=None):
    stats = ema_reduce(residual[None], idx[None], None, embeddings.shape[0])
    loss = rvq_level(residual, residual_next, embeddings, idx, None, training, first_level, quantized_out, next_cache,
                     q_out)
    return loss, stats


def rvq_replay_out(x, codebooks, idxs, training, mask_u8, out=None):
    """vqb_rvq_replay_out: the running sum of the levels' outputs, replayed from the input and the indices."""
    keep = torch.ones(x.shape[0], dtype=torch.bool) if mask_u8 is None else mask_u8.bool()
    r, acc = x, torch.zeros_like(x)
    for c, i, tr in zip(codebooks, idxs, training):
        cq = c[i]
        q = r + (cq - r) if tr else cq
        q = torch.where(keep[:, None], q, r)
        acc = acc + q
        r = r - q
    if out is not None:
        out.copy_(acc)
        return out
    return acc


def dense_gumbel_sample(x, emb, use_cosine_sim, temperature, uniforms=None, generator=None):
    """vqb_dense_gumbel_sample with the reference's own torch calls (utils/general.py:107-129)."""
    sim = torch.einsum("hnd,hcd->hnc", x.float(), emb) if use_cosine_sim else -torch.cdist(x.float(), emb)
    u = uniforms.reshape(sim.shape) if uniforms is not None else torch.zeros_like(sim).uniform_(0, 1)
    g = -torch.log((-torch.log(u.clamp(min=1e-5))).clamp(min=1e-5))
    return (sim / temperature + g).argmax(dim=-1)


def dense_scores(x, emb, use_cosine_sim):
    return torch.einsum("hnd,hcd->hnc", x.float(), emb) if use_cosine_sim else -torch.cdist(x.float(), emb)


def column_moments(x, mask_u8):
    xd = x.double()
    if mask_u8 is not None:
        xd = xd * mask_u8.bool()[None, :, None]
        rows = torch.full((x.shape[0],), int(mask_u8.bool().sum()), dtype=torch.int64)
    else:
        rows = torch.full((x.shape[0],), x.shape[1], dtype=torch.int64)
    return torch.stack([xd.sum(dim=1), (xd * xd).sum(dim=1)], dim=-1), rows


def rvq_backward(x, codebooks, idxs, training, coef, g_out, mask_u8):
    """vqb_rvq_backward: Q * g_out + sum_l live * coef[l] * (r_l - c_l), residuals replayed from x."""
    keep = torch.ones(x.shape[0], dtype=torch.bool) if mask_u8 is None else mask_u8.bool()
    r = x
    acc = len(codebooks) * (g_out.float() if g_out is not None else torch.zeros_like(x))
    for l, (c, i, tr) in enumerate(zip(codebooks, idxs, training)):
        cq = c[i]
        acc = acc + torch.where(keep[:, None], coef[l] * (r - cq), torch.zeros_like(r))
        q = r + (cq - r) if tr else cq
        q = torch.where(keep[:, None], q, r)
        r = r - q
    return acc


def minkey_pack(score, idx):
    """(orderable fp32 score << 32) | index as int64: smaller score first, lowest index on ties (vqb_minkey_pack)."""
    u = score.contiguous().view(torch.int32).to(torch.int64) & 0xFFFFFFFF
    neg = (u >> 31) == 1
    u = torch.where(neg, (~u) & 0xFFFFFFFF, u | 0x80000000)
    u = u ^ 0x80000000                                   # keep the sign bit of the int64 key clear
    key = (u << 32) | (idx & 0xFFFFFFFF)
    return key - ((key >> 63) << 64)


def minkey_unpack(keys, want_score=False):
    return keys & 0xFFFFFFFF, None


def ema_apply_sharded(stats, cluster_size, embed_avg, embeddings, weight, eps, weights_l2norm, k_total, all_reduce):
    cluster_size.lerp_(stats[..., -1], weight)
    totals = cluster_size.sum(-1)
    all_reduce(totals)                                   # sum(cluster_size) over the shards
    embed_avg.lerp_(stats[..., :-1], weight)
    total = totals[..., None]
    smoothed = (cluster_size + eps) / (total + k_total * eps) * total
    new = embed_avg / smoothed[..., None]
    embeddings.copy_(F.normalize(new, dim=-1) if weights_l2norm else new)


REPLACED = ("prepare_codebook", "search", "l2norm_rows", "gather_st_loss", "st_commit_backward", "ema_reduce",
            "ema_apply", "expire_scatter", "minkey_pack", "minkey_unpack", "ema_apply_sharded", "rvq_level",
            "rvq_level_ema", "rvq_replay_out", "rvq_backward", "dense_gumbel_sample", "dense_scores", "column_moments",
            "l2norm_prepare_supported", "quantize_ema_supported", "rvq_level_ema_supported",
            "rvq_replay_out_supported")


def install(ops, lib):
    """Replace the wrappers on the `vqb200.ops` module object and the device guard of `vqb200._lib`."""
    for name in ("prepare_codebook", "search", "l2norm_rows", "gather_st_loss", "st_commit_backward", "ema_reduce",
                 "ema_apply", "expire_scatter", "minkey_pack", "minkey_unpack", "ema_apply_sharded", "rvq_level", "rvq_level_ema",
                 "rvq_replay_out", "rvq_backward", "dense_gumbel_sample", "dense_scores", "column_moments"):
        setattr(ops, name, globals()[name])
    ops.l2norm_prepare_supported = lambda d: False
    ops.quantize_ema_supported = lambda d: False
    ops.rvq_level_ema_supported = lambda d: False
    ops.rvq_replay_out_supported = lambda d, q: True
    lib.require_device = lambda x: None
