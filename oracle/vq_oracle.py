"""CPU oracle for the codebook hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import this module.  The product
path (``vqb200``) never does; it fails loudly when the CUDA library is missing.

What this is
------------
A restatement, in plain torch-on-CPU, of the arithmetic the reference
(`MisterBourbaki/vector-quantization-by-ml`, a pure-Python torch library)
performs on the codebook path.  The reference's arithmetic lives in third-party
torch ops (``torch.cdist`` -> ATen ``_euclidean_dist`` -> SGEMM, ``einsum`` ->
``bmm``, ``argmax``, ``F.one_hot``, ``lerp_``, ``mse_loss``, ``F.normalize``;
reference lock pins torch 2.4.0, this image has torch 2.11), so the oracle calls
the *same* torch ops in the *same* order -- that is what makes it bit-faithful
and also what makes it a fair CPU timing baseline.

Pinning
-------
The reference's own tests pin no numeric value on this path (shape assertions
only: reference ``tests/test_vector_quantize_pytorch.py:30-31``).  The oracle is
therefore pinned against outputs of the reference itself, run in the build
container: ``tests/golden/make_golden.py`` imports ``/root/reference`` and
writes fixtures under ``tests/golden/*.pt``; ``tests/test_oracle_golden.py``
checks this file against them bit-for-bit (indices, quantize, loss,
cluster_size) / to 1e-6 (embed_avg, embeddings).

Reference lines followed (relative to /root/reference/vector_quantization):
  codebooks.py:350-435   Codebook.forward
  codebooks.py:230-255   replace_codes / expire_codes_
  utils/general.py:41-89 sample_vectors / batched_sample_vectors
  utils/general.py:92-98 ema_inplace (lerp_)
  utils/general.py:112-136 gumbel_sample deterministic branch (argmax + one_hot)
  utils/general.py:154-163 laplace_smoothing / batched_embedding
  utils/losses.py:5-19   l2norm
  vector_quantize_pytorch.py:182-430 VectorQuantize.forward (layout glue, ST, commit loss)
  residual_vq.py:134-269 ResidualVQ.forward
  utils/kmeans.py:38-120 kmeans (first-forward init)
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, List, Optional, Tuple

import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------- #
# small helpers
# --------------------------------------------------------------------------- #

def l2norm(t: torch.Tensor) -> torch.Tensor:
    """utils/losses.py:19 -- F.normalize(p=2, dim=-1), eps 1e-12."""
    return F.normalize(t, p=2, dim=-1)


def default_init(num_codebooks: int, codebook_size: int, dim: int) -> torch.Tensor:
    """utils/general.py:101-104 -- kaiming_uniform_ on a fresh (H,K,d) tensor."""
    t = torch.empty(num_codebooks, codebook_size, dim)
    torch.nn.init.kaiming_uniform_(t)
    return t


def similarities(flat: torch.Tensor, emb: torch.Tensor, use_cosine_sim: bool) -> torch.Tensor:
    """codebooks.py:122-123 (dot) / :128-129 (-cdist).  (H,N,d),(H,K,d)->(H,N,K)."""
    if use_cosine_sim:
        return torch.einsum("hnd,hcd->hnc", flat, emb)
    return -torch.cdist(flat, emb)


def pick_rows(n_rows: int, m: int, device) -> torch.Tensor:
    """utils/general.py:62-66 -- which batch rows replace m dead codes.

    Consumes the global torch generator exactly as the reference does."""
    if n_rows >= m:
        return torch.randperm(n_rows, device=device)[:m]
    return torch.randint(0, n_rows, (m,), device=device)


# --------------------------------------------------------------------------- #
# codebook state + parameters
# --------------------------------------------------------------------------- #

@dataclass
class CodebookState:
    """The three persistent buffers of reference Codebook (codebooks.py:183-190)."""
    embeddings: torch.Tensor     # (H,K,d) fp32
    embed_avg: torch.Tensor      # (H,K,d) fp32
    cluster_size: torch.Tensor   # (H,K)   fp32
    is_initialized: bool = True

    @staticmethod
    def fresh(num_codebooks: int, codebook_size: int, dim: int,
              weights_l2norm: bool = False, kmeans_init: bool = False) -> "CodebookState":
        """codebooks.py:136-139,182-190."""
        emb = (torch.zeros(num_codebooks, codebook_size, dim) if kmeans_init
               else default_init(num_codebooks, codebook_size, dim))
        if weights_l2norm:
            emb = l2norm(emb)
        return CodebookState(emb, emb.clone(), torch.zeros(num_codebooks, codebook_size),
                             is_initialized=not kmeans_init)

    def clone(self) -> "CodebookState":
        return CodebookState(self.embeddings.clone(), self.embed_avg.clone(),
                             self.cluster_size.clone(), self.is_initialized)


@dataclass
class CodebookOpts:
    """Subset of reference CodebookParams (codebooks.py:58-78) on the hot path."""
    decay: float = 0.8
    eps: float = 1e-5
    threshold_ema_dead_code: int = 2
    reset_cluster_size: Optional[int] = None
    use_cosine_sim: bool = False
    weights_l2norm: bool = False          # weights_regularization == "l2norm"
    ema_update: bool = True
    kmeans_iters: int = 10

    @property
    def reset(self) -> float:
        return float(self.threshold_ema_dead_code if self.reset_cluster_size is None
                     else self.reset_cluster_size)


# --------------------------------------------------------------------------- #
# kmeans init (utils/kmeans.py:38-120)
# --------------------------------------------------------------------------- #

def kmeans_init(vectors: torch.Tensor, num_clusters: int, iters: int, use_cosine_sim: bool,
                all_reduce: Callable[[torch.Tensor], None] = lambda t: None
                ) -> Tuple[torch.Tensor, torch.Tensor]:
    H, N, d = vectors.shape
    cents = torch.stack([v[pick_rows(N, num_clusters, v.device)] for v in vectors.unbind(0)], 0)
    counts = None
    for _ in range(iters):
        if use_cosine_sim:
            sim = vectors @ cents.transpose(1, 2)
        else:
            sim = -torch.cdist(vectors, cents)
        labels = sim.argmax(-1)
        counts = torch.zeros(H, num_clusters, dtype=labels.dtype)
        counts.scatter_add_(-1, labels, torch.ones_like(labels))
        all_reduce(counts)
        empty = counts == 0
        denom = counts.masked_fill(empty, 1)
        sums = torch.zeros(H, num_clusters, d, dtype=vectors.dtype)
        sums.scatter_add_(1, labels[..., None].expand(-1, -1, d), vectors)
        means = sums / denom[..., None]
        all_reduce(means)
        if use_cosine_sim:
            means = l2norm(means)
        cents = torch.where(empty[..., None], cents, means)
    return cents, counts


# --------------------------------------------------------------------------- #
# Codebook.forward  (codebooks.py:350-435)
# --------------------------------------------------------------------------- #

def codebook_forward(state: CodebookState, x: torch.Tensor, opts: CodebookOpts, *,
                     training: bool = True, mask: Optional[torch.Tensor] = None,
                     freeze_codebook: bool = False, row_chunk: Optional[int] = None,
                     all_reduce: Callable[[torch.Tensor], None] = lambda t: None,
                     want_gap: bool = False):
    """One forward of the reference Codebook on CPU.

    x: (B,n,d) or (H,B,n,d), any float dtype.  Mutates ``state`` in place when
    training with EMA.  Returns (quantize fp32 like x, indices int64, extras) where
    extras has 'top2_rel_gap' (H,N) when ``want_gap`` -- the reference's own fp32
    top-2 distance gap, relative, used for the north-star index exemption.

    ``row_chunk``: evaluate search/gather/statistics on row chunks (the reference
    cannot hold N x K at the big configs; cdist on row chunks is bitwise identical
    to the unchunked call).  EMA statistics are accumulated across
    chunks with the same einsum per chunk, so embed_sum differs from the unchunked
    SGEMM only by fp32 summation order.
    """
    needs_h = x.ndim < 4
    x = x.float()                                          # :354
    if needs_h:
        x = x.unsqueeze(0)                                 # :356-357
    H, d = x.shape[0], x.shape[-1]
    lead = x.shape[1:-1]
    flat = x.reshape(H, -1, d)                             # :359
    N = flat.shape[1]
    K = state.embeddings.shape[1]

    flat_mask = None
    if mask is not None:                                   # :361-367
        rep = N // (mask.shape[0] * mask.shape[1])
        flat_mask = mask[:, None, :].expand(mask.shape[0], rep, mask.shape[1]).reshape(1, -1).expand(H, -1)

    if not state.is_initialized:                           # :368-370, :208-228
        data = flat
        if flat_mask is not None:
            data = flat[flat_mask].reshape(H, -1, d)
        cents, counts = kmeans_init(data, K, opts.kmeans_iters, opts.use_cosine_sim, all_reduce)
        state.embeddings.copy_(cents)
        state.embed_avg.copy_(cents * counts[..., None])
        state.cluster_size.copy_(counts)
        state.is_initialized = True

    emb = state.embeddings
    do_ema = training and opts.ema_update and not freeze_codebook

    chunk = N if row_chunk is None else max(1, int(row_chunk))
    ind = torch.empty(H, N, dtype=torch.long)
    quant = torch.empty(H, N, d)
    counts = torch.zeros(H, K)
    sums = torch.zeros(H, K, d)
    gap = torch.empty(H, N) if want_gap else None

    for lo in range(0, N, chunk):
        hi = min(N, lo + chunk)
        part = flat[:, lo:hi]
        sim = similarities(part, emb, opts.use_cosine_sim)            # :386
        idx = sim.argmax(-1)                                         # general.py:128
        onehot = F.one_hot(idx, K).type(sim.dtype)                   # general.py:129
        ind[:, lo:hi] = idx
        if training:
            quant[:, lo:hi] = torch.einsum("hnc,hcd->hnd", onehot, emb)   # :395
        else:
            quant[:, lo:hi] = emb.gather(1, idx[..., None].expand(-1, -1, d))  # :397
        if want_gap and K > 1:
            top2 = sim.topk(2, dim=-1).values
            denom = top2[..., 0].abs().clamp_min(1e-30)
            gap[:, lo:hi] = (top2[..., 0] - top2[..., 1]).abs() / denom
        elif want_gap:
            gap[:, lo:hi] = float("inf")
        if do_ema:
            if flat_mask is not None:
                onehot[~flat_mask[:, lo:hi]] = 0.0                   # :405-406
            counts += onehot.sum(dim=1)                              # :408
            sums += torch.einsum("hnd,hnc->hcd", part, onehot)       # :413

    if do_ema:
        w = 1 - opts.decay
        all_reduce(counts)                                           # :410
        state.cluster_size.lerp_(counts, w)                          # :411
        sums = sums.contiguous()
        all_reduce(sums)                                             # :415
        state.embed_avg.lerp_(sums, w)                               # :417
        total = state.cluster_size.sum(dim=-1, keepdim=True)
        smoothed = (state.cluster_size + opts.eps) / (total + K * opts.eps) * total   # :419-421
        new_emb = state.embed_avg / smoothed[..., None]              # :423
        if opts.weights_l2norm:
            new_emb = l2norm(new_emb)                                # :424
        state.embeddings.copy_(new_emb)                              # :425
        expire_codes(state, x, opts)                                 # :426

    quant = quant.reshape(H, *lead, d)
    ind = ind.reshape(H, *lead)
    if needs_h:
        quant, ind = quant[0], ind[0]
    extras = {}
    if want_gap:
        extras["top2_rel_gap"] = gap
    return quant, ind, extras


def expire_codes(state: CodebookState, x: torch.Tensor, opts: CodebookOpts,
                 sample_fn: Optional[Callable[[torch.Tensor, int], torch.Tensor]] = None) -> None:
    """codebooks.py:245-255 + :230-243.  x is the (H,...,d) input handed to forward."""
    if opts.threshold_ema_dead_code == 0:
        return
    dead = state.cluster_size < opts.threshold_ema_dead_code
    if not bool(torch.any(dead)):
        return
    H, d = x.shape[0], x.shape[-1]
    rows = x.reshape(H, -1, d)
    if opts.weights_l2norm:
        rows = l2norm(rows)                                 # :231
    for h in range(H):
        m = int(dead[h].sum().item())                       # :234
        if sample_fn is None:
            picked = rows[h][pick_rows(rows.shape[1], m, rows.device)]
        else:
            picked = sample_fn(rows[h], m)
        state.embeddings[h][dead[h]] = picked               # :241
        state.cluster_size[h][dead[h]] = opts.reset         # :242
        state.embed_avg[h][dead[h]] = picked * opts.reset   # :243


# --------------------------------------------------------------------------- #
# VectorQuantize.forward (vector_quantize_pytorch.py:182-430), EMA path only
# --------------------------------------------------------------------------- #

@dataclass
class VQOpts:
    heads: int = 1
    separate_codebook_per_head: bool = False
    channel_last: bool = True
    commitment_weight: float = 1.0
    input_l2norm: bool = False            # transform_input == "l2norm"
    codebook: CodebookOpts = field(default_factory=CodebookOpts)


def vq_forward(state: CodebookState, x: torch.Tensor, opts: VQOpts, *, training: bool = True,
               mask: Optional[torch.Tensor] = None, freeze_codebook: bool = False,
               row_chunk: Optional[int] = None, want_gap: bool = False,
               all_reduce: Callable[[torch.Tensor], None] = lambda t: None):
    """Returns (quantize, indices, loss[1], extras) like the reference forward."""
    orig = x
    only_one = x.ndim == 2
    if only_one:
        x = x[:, None, :]                                   # :194-196
    heads = opts.heads
    multi = heads > 1
    B = x.shape[0]
    if not opts.channel_last:                               # :210-211
        x = x.movedim(1, -1)
    spatial = None
    if x.ndim >= 4:                                         # :212-213
        spatial = x.shape[1:-1]
        x = x.reshape(B, -1, x.shape[-1])
    if multi:                                               # :217-219
        n = x.shape[1]
        dh = x.shape[-1] // heads
        xh = x.reshape(B, n, heads, dh)
        if opts.separate_codebook_per_head:
            x = xh.permute(2, 0, 1, 3)                      # h b n d
        else:
            x = xh.permute(0, 2, 1, 3).reshape(1, B * heads, n, dh)   # 1 (b h) n d
    if opts.input_l2norm:                                   # :221
        x = l2norm(x)

    quant, ind, extras = codebook_forward(state, x, opts.codebook, training=training, mask=mask,
                                          freeze_codebook=freeze_codebook, row_chunk=row_chunk,
                                          want_gap=want_gap, all_reduce=all_reduce)
    commit_q = quant
    if training:
        quant = x + (quant - x)                             # :273 (values; detach is autograd-only)

    if multi:                                               # :303-307
        if opts.separate_codebook_per_head:
            ind = ind.permute(1, 2, 0)
        else:
            ind = ind.reshape(B, heads, -1).permute(0, 2, 1)
    if spatial is not None:                                 # :309-312
        ind = ind.reshape(B, *spatial, *ind.shape[2:])
    if only_one:
        ind = ind[:, 0]                                     # :314-315

    loss = torch.tensor([0.0])                              # :319
    if training and opts.commitment_weight > 0:
        if mask is not None:                                # :347-360
            per = F.mse_loss(commit_q, x.float(), reduction="none")
            lm = mask
            if multi:
                lm = mask[None, :, None, :].expand(per.shape[0], mask.shape[0],
                                                   per.shape[1] // mask.shape[0], mask.shape[1]
                                                   ).reshape(per.shape[0], per.shape[1], mask.shape[1])
            commit = per[lm].mean()
        else:
            commit = F.mse_loss(commit_q, x.float())        # :362
        loss = loss + commit * opts.commitment_weight       # :364

    if multi:                                               # :394-398
        if opts.separate_codebook_per_head:
            quant = quant.permute(1, 2, 0, 3).reshape(B, quant.shape[2], -1)
        else:
            n = quant.shape[2]
            quant = quant.reshape(B, heads, n, -1).permute(0, 2, 1, 3).reshape(B, n, -1)
    if spatial is not None:                                 # :406-407
        quant = quant.reshape(B, *spatial, quant.shape[-1])
    if not opts.channel_last:                               # :408-409
        quant = quant.movedim(-1, 1)
    if only_one:
        quant = quant[:, 0]                                 # :410-411
    if mask is not None:                                    # :415-418
        quant = torch.where(mask[..., None], quant, orig)
    return quant, ind, loss, extras


# --------------------------------------------------------------------------- #
# Learnable codebook (codebooks.py:186-190, 375-377; vector_quantize_pytorch.py:261-279, 362): single head,
# channel-last, no EMA.  Autograd-carrying restatement: gradients come from torch autograd over the same ops.
# --------------------------------------------------------------------------- #

def vq_forward_learnable(embeddings: torch.Tensor, x: torch.Tensor, *, commitment_weight: float = 1.0,
                         sync_update_v: float = 0.0, mask: Optional[torch.Tensor] = None, inplace_optimizer=None,
                         use_cosine_sim: bool = False, ce_commit: bool = False, diversity_weight: float = 0.0,
                         diversity_temperature: float = 100.0, targets: Optional[torch.Tensor] = None):
    """`embeddings` (1,K,d) and `x` (B,n,d) may require grad.  Returns (quantize, indices, loss[1]); with
    `inplace_optimizer` (a torch optimizer over [embeddings]; vector_quantize_pytorch.py:233-256) the codebook is
    first stepped on mse(codes, x.detach()) inside the forward, and that loss is returned as a fourth value.
    `ce_commit` / `diversity_weight` / `targets`: the losses on the dense similarities, which stay attached to
    `embeddings` here (codebooks.py:375-377); with `targets` the return is (quantize, ce) (:298-299)."""
    B, n, d = x.shape
    inplace_loss = None
    if inplace_optimizer is not None:
        xd = x.detach().float()
        sim0 = similarities(xd.reshape(1, -1, d), embeddings.detach(), use_cosine_sim)
        oh0 = F.one_hot(sim0.argmax(-1).reshape(1, B, n), embeddings.shape[1]).type(xd.dtype)
        q0 = torch.einsum("h b n c, h c d -> h b n d", oh0, embeddings)[0]
        if mask is not None:
            inplace_loss = F.mse_loss(q0, xd, reduction="none")[mask].mean()   # :234-246
        else:
            inplace_loss = F.mse_loss(q0, xd)                            # :249
        inplace_loss.backward()                                         # :251-253
        inplace_optimizer.step()
        inplace_optimizer.zero_grad()
        inplace_loss = inplace_loss.detach()
    flat = x.float()[None]                                              # codebooks.py:354-357
    sim = similarities(flat.reshape(1, -1, d), embeddings, use_cosine_sim)   # :386, attached to x and the codebook
    ind = sim.detach().argmax(-1).reshape(1, B, n)                      # utils/general.py:128
    onehot = F.one_hot(ind, embeddings.shape[1]).type(flat.dtype)       # :129
    quant = torch.einsum("h b n c, h c d -> h b n d", onehot, embeddings)[0]   # codebooks.py:393-395 (differentiable)
    commit_q = quant                                                    # vector_quantize_pytorch.py:262-268
    out = x + (quant - x).detach()                                      # :273
    if sync_update_v > 0.0:
        out = out + sync_update_v * (out - out.detach())                # :275-279
    logits = sim.reshape(B, n, -1).permute(0, 2, 1)                     # "1 b n l -> b l n" (:286)
    if targets is not None:
        return out, F.cross_entropy(logits, targets, ignore_index=-1)   # :298-299
    ind_out = ind[0].clone()
    loss = torch.tensor([0.0])
    if diversity_weight > 0:                                            # :324-333
        prob = (-sim.reshape(1, B, n, -1) * diversity_temperature).softmax(dim=-1)
        avg_prob = prob.reshape(-1, n, prob.shape[-1]).mean(0)
        div = -((-avg_prob * _log_eps(avg_prob)).sum(dim=-1)).mean()
        loss = loss + div * diversity_weight
    if commitment_weight > 0:
        if ce_commit:                                                   # :338-346
            if mask is not None:
                ind_out.masked_fill_(~mask, -1)
            commit = F.cross_entropy(logits, ind_out, ignore_index=-1)
        elif mask is not None:
            commit = F.mse_loss(commit_q, x, reduction="none")[mask].mean()   # :347-360
        else:
            commit = F.mse_loss(commit_q, x)                            # :362
        loss = loss + commit * commitment_weight
    if mask is not None:
        out = torch.where(mask[..., None], out, x)                      # :415-418
    if inplace_optimizer is not None:
        return out, ind_out, loss, inplace_loss
    return out, ind_out, loss


# --------------------------------------------------------------------------- #
# Consumers of the dense N x K similarities (vector_quantize_pytorch.py:284-299 CE to given indices,
# :338-346 CE commitment, :324-333 codebook diversity loss).  Autograd-carrying restatement on the same torch ops.
# --------------------------------------------------------------------------- #

def _log_eps(t: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    """utils/general.py:25-26."""
    return t.clamp(min=eps).log()


def vq_forward_dense(state: CodebookState, x: torch.Tensor, opts: VQOpts, *, training: bool = True,
                     mask: Optional[torch.Tensor] = None, targets: Optional[torch.Tensor] = None,
                     ce_commit: bool = False, diversity_weight: float = 0.0, diversity_temperature: float = 100.0,
                     freeze_codebook: bool = False):
    """``x`` may require grad.  With ``targets``: returns (quantize, ce) like the reference's ``return_loss`` branch
    (:298-299).  Otherwise (quantize, indices, loss[1], {"commitment", "codebook_diversity"}).

    The similarities are built on ``embeddings.detach()`` (codebooks.py:375-377, 386) and the EMA step then overwrites
    the SAME storage through ``.data`` (:425), which autograd does not see: the reference's input gradient therefore
    combines the PRE-update distances with the POST-update code vectors.  This restatement keeps that (the update
    below goes through ``.data`` aliases of the state tensors)."""
    orig = x
    only_one = x.ndim == 2
    if only_one:
        x = x[:, None, :]
    heads, multi = opts.heads, opts.heads > 1
    B = x.shape[0]
    if not opts.channel_last:
        x = x.movedim(1, -1)
    spatial = None
    if x.ndim >= 4:
        spatial = x.shape[1:-1]
        x = x.reshape(B, -1, x.shape[-1])
    if multi:
        n, dh = x.shape[1], x.shape[-1] // heads
        xh = x.reshape(B, n, heads, dh)
        x = xh.permute(2, 0, 1, 3) if opts.separate_codebook_per_head else \
            xh.permute(0, 2, 1, 3).reshape(1, B * heads, n, dh)
    if opts.input_l2norm:
        x = l2norm(x)                                                   # :221
    x4 = (x if x.ndim == 4 else x[None]).float()                        # codebooks.py:354-357
    H, Bp, n, d = x4.shape
    K = state.embeddings.shape[1]
    sim = similarities(x4.reshape(H, -1, d), state.embeddings.detach(), opts.codebook.use_cosine_sim)   # :386
    sim = sim.reshape(H, Bp, n, K)                                      # :433

    alias = CodebookState(state.embeddings.data, state.embed_avg.data, state.cluster_size.data, state.is_initialized)
    with torch.no_grad():
        quant, ind, _ = codebook_forward(alias, x.detach(), opts.codebook, training=training, mask=mask,
                                         freeze_codebook=freeze_codebook)
    codes_q = quant                                                     # :262-268 (detached: no learnable codebook here)
    if training:
        quant = x + (quant - x).detach()                                # :273

    def ce_loss(codes):                                                 # :284-296
        if not multi:
            logits = sim.permute(0, 1, 3, 2)[0]                         # 1 b n l -> b l n
        elif opts.separate_codebook_per_head:
            logits = sim.permute(1, 3, 2, 0)                            # c b n l -> b l n c
        else:
            logits = sim.reshape(B, heads, n, K).permute(0, 3, 2, 1)    # 1 (b h) n l -> b l n h
        return F.cross_entropy(logits, codes, ignore_index=-1)

    if targets is not None:
        return quant, ce_loss(targets)                                  # :298-299

    if multi:                                                           # :303-307
        ind = ind.permute(1, 2, 0) if opts.separate_codebook_per_head else ind.reshape(B, heads, -1).permute(0, 2, 1)
    if spatial is not None:
        ind = ind.reshape(B, *spatial, *ind.shape[2:])
    if only_one:
        ind = ind[:, 0]
    ind = ind.contiguous()

    loss = torch.tensor([0.0])
    parts = {"commitment": torch.tensor(0.0), "codebook_diversity": torch.tensor(0.0)}
    if training:
        if diversity_weight > 0:                                        # :324-333
            prob = (-sim * diversity_temperature).softmax(dim=-1)
            avg_prob = prob.reshape(-1, n, K).mean(0)                   # "... n l -> n l"
            div = -((-avg_prob * _log_eps(avg_prob)).sum(dim=-1)).mean()
            parts["codebook_diversity"] = div
            loss = loss + div * diversity_weight
        if opts.commitment_weight > 0:
            if ce_commit:                                               # :338-346
                if mask is not None:
                    m = mask[..., None].expand(*mask.shape, heads) if multi else mask
                    ind.masked_fill_(~m, -1)                            # in place: the RETURNED indices carry the -1
                commit = ce_loss(ind)
            elif mask is not None:                                      # :347-360
                per = F.mse_loss(codes_q, x, reduction="none")
                lm = mask
                if multi:
                    lm = mask[None, :, None, :].expand(per.shape[0], mask.shape[0], per.shape[1] // mask.shape[0],
                                                       mask.shape[1]).reshape(per.shape[0], per.shape[1], mask.shape[1])
                commit = per[lm].mean()
            else:
                commit = F.mse_loss(codes_q, x)                         # :362
            parts["commitment"] = commit
            loss = loss + commit * opts.commitment_weight

    if multi:                                                           # :394-398
        if opts.separate_codebook_per_head:
            quant = quant.permute(1, 2, 0, 3).reshape(B, quant.shape[2], -1)
        else:
            quant = quant.reshape(B, heads, n, -1).permute(0, 2, 1, 3).reshape(B, n, -1)
    if spatial is not None:
        quant = quant.reshape(B, *spatial, quant.shape[-1])
    if not opts.channel_last:
        quant = quant.movedim(-1, 1)
    if only_one:
        quant = quant[:, 0]
    if mask is not None:
        quant = torch.where(mask[..., None], quant, orig)
    return quant, ind, loss, parts


def rvq_forward_autograd(states: List[CodebookState], x: torch.Tensor, opts: VQOpts, *,
                         mask: Optional[torch.Tensor] = None, **dense_kw):
    """ResidualVQ training forward under autograd (residual_vq.py:212-243): every level is `vq_forward_dense`
    (plain mse commitment unless `dense_kw` says otherwise); `residual -= quantized.detach()`.
    Returns (quantized_out, indices (...,Q), losses (1,Q))."""
    out, residual = 0.0, x
    inds, losses = [], []
    for st in states:
        q, ind, loss, _ = vq_forward_dense(st, residual, opts, training=True, mask=mask, **dense_kw)
        residual = residual - q.detach()                    # :232
        out = out + q                                       # :233
        inds.append(ind)
        losses.append(loss)
    return out, torch.stack(inds, -1), torch.stack(losses, -1)


# --------------------------------------------------------------------------- #
# ResidualVQ.forward (residual_vq.py:134-269), no quantize-dropout
# --------------------------------------------------------------------------- #

def rvq_forward(states: List[CodebookState], x: torch.Tensor, opts: VQOpts, *, training: bool = True,
                mask: Optional[torch.Tensor] = None, freeze_codebook: bool = False,
                row_chunk: Optional[int] = None, want_gap: bool = False,
                all_reduce: Callable[[torch.Tensor], None] = lambda t: None):
    """states: one CodebookState per level (the same object repeated for shared_codebook).

    Returns (quantized_out, indices (...,Q), losses (1,Q), extras-per-level)."""
    out = 0.0                                               # :154
    residual = x                                            # :155
    all_ind, all_loss, all_extras = [], [], []
    for st in states:                                       # :212
        q, ind, loss, ex = vq_forward(st, residual, opts, training=training, mask=mask,
                                      freeze_codebook=freeze_codebook, row_chunk=row_chunk,
                                      want_gap=want_gap, all_reduce=all_reduce)
        residual = residual - q                             # :232
        out = out + q                                       # :233
        all_ind.append(ind)
        all_loss.append(loss)
        all_extras.append(ex)
    return out, torch.stack(all_ind, -1), torch.stack(all_loss, -1), all_extras


# --------------------------------------------------------------------------- #
# f4 variants: gumbel sampling, affine re-parametrisation, orthogonal regularisation.
# Autograd-carrying restatement on the reference's own torch ops; pinned by tests/golden/f4 (recorded from the live
# reference by tests/golden/make_golden_f4.py) in tests/test_oracle_f4.py.
# --------------------------------------------------------------------------- #

def gumbel_noise(u: torch.Tensor) -> torch.Tensor:
    """utils/general.py:107-109 for a given uniform draw u (the reference draws `zeros_like(t).uniform_(0, 1)`)."""
    return -_log_eps(-_log_eps(u))


def gumbel_sample(logits: torch.Tensor, *, temperature: float = 1.0, stochastic: bool = False,
                  straight_through: bool = False, reinmax: bool = False, training: bool = True,
                  uniforms: Optional[torch.Tensor] = None):
    """utils/general.py:112-151 (dim = -1).  Returns (ind, one_hot); the one-hot carries the straight-through / reinmax
    graph when those are on.  `uniforms`: the (h,n,c) draw to use instead of a fresh one."""
    size = logits.shape[-1]
    if training and stochastic and temperature > 0:                       # :123-126
        u = uniforms if uniforms is not None else torch.zeros_like(logits).uniform_(0, 1)
        sampling_logits = (logits / temperature) + gumbel_noise(u)
    else:
        sampling_logits = logits
    ind = sampling_logits.argmax(dim=-1)                                  # :128
    one_hot = F.one_hot(ind, size).type(logits.dtype)                     # :129
    assert not (reinmax and not straight_through)                         # :131-133
    if not straight_through or temperature <= 0.0 or not training:        # :135-136
        return ind, one_hot
    if reinmax:                                                           # :141-146 (softmax over dim=1 as written there)
        prob0 = logits.softmax(dim=-1)
        prob1 = (one_hot + (logits / temperature).softmax(dim=-1)) / 2
        prob1 = ((_log_eps(prob1) - logits).detach() + logits).softmax(dim=1)
        prob2 = 2 * prob1 - 0.5 * prob0
        one_hot = prob2 - prob2.detach() + one_hot
    else:                                                                 # :147-149
        prob1 = (logits / temperature).softmax(dim=-1)
        one_hot = one_hot + prob1 - prob1.detach()
    return ind, one_hot


def orthogonal_loss(t: torch.Tensor) -> torch.Tensor:
    """utils/losses.py:22-27."""
    h, n = t.shape[:2]
    normed = l2norm(t)
    cosine_sim = torch.einsum("h i d, h j d -> h i j", normed, normed)
    return (cosine_sim ** 2).sum() / (h * n ** 2) - (1 / n)


@dataclass
class AffineState:
    """The affine buffers of reference Codebook (codebooks.py:194-206) + AffineParameters (:31-37)."""
    sync: bool = False
    batch_decay: float = 0.99
    codebook_decay: float = 0.9
    batch_mean: Optional[torch.Tensor] = None
    batch_variance: Optional[torch.Tensor] = None
    codebook_mean: Optional[torch.Tensor] = None
    codebook_variance: Optional[torch.Tensor] = None

    def _decay(self, name: str, new: torch.Tensor, decay: float) -> None:      # :258-272
        old = getattr(self, name)
        setattr(self, name, new.detach() if old is None else old * decay + new.detach() * (1 - decay))

    def update(self, data: torch.Tensor, embeddings: torch.Tensor, training: bool,
               flat_mask: Optional[torch.Tensor]) -> None:                     # :274-348, sync off / single process
        if training:
            self._decay("codebook_mean", embeddings.mean(dim=1, keepdim=True), self.codebook_decay)
            self._decay("codebook_variance", embeddings.var(dim=1, unbiased=False, keepdim=True), self.codebook_decay)
        if flat_mask is not None:
            data = data[flat_mask].reshape(data.shape[0], -1, data.shape[-1])
        self._decay("batch_mean", data.mean(dim=1, keepdim=True), self.batch_decay)
        self._decay("batch_variance", data.var(dim=1, unbiased=False, keepdim=True), self.batch_decay)


def codebook_forward_variants(state: CodebookState, x: torch.Tensor, opts: CodebookOpts, *, gumbel: Optional[dict] = None,
                              affine: Optional[AffineState] = None, training: bool = True,
                              mask: Optional[torch.Tensor] = None, freeze_codebook: bool = False,
                              learnable: bool = False, uniforms: Optional[torch.Tensor] = None):
    """reference Codebook.forward (codebooks.py:350-435) with gumbel sampling and the affine re-parametrisation, under
    autograd.  `state.embeddings` may require grad (`learnable`).  Returns (quantize, ind, similarities) shaped as the
    reference returns them; mutates `state` (EMA step, expiry) and `affine`."""
    gumbel = {"training": True, **(gumbel or {})}
    needs_h = x.ndim < 4
    x = x.float()
    if needs_h:
        x = x.unsqueeze(0)
    H, d = x.shape[0], x.shape[-1]
    lead = x.shape[1:-1]
    flat = x.reshape(H, -1, d)
    N, K = flat.shape[1], state.embeddings.shape[1]
    flat_mask = None
    if mask is not None:
        rep = N // (mask.shape[0] * mask.shape[1])
        flat_mask = mask[:, None, :].expand(mask.shape[0], rep, mask.shape[1]).reshape(1, -1).expand(H, -1)
    if affine is not None:
        affine.update(flat, state.embeddings.detach(), training, flat_mask)   # :372-373
    emb = state.embeddings if learnable else state.embeddings.detach()    # :375-377 (an alias of the buffer)
    codebook_std = batch_std = None
    if affine is not None:                                                # :379-384
        codebook_std = affine.codebook_variance.clamp(min=1e-5).sqrt()
        batch_std = affine.batch_variance.clamp(min=1e-5).sqrt()
        emb = (emb - affine.codebook_mean) * (batch_std / codebook_std) + affine.batch_mean
    sim = similarities(flat, emb, opts.use_cosine_sim)                    # :386
    ind, onehot = gumbel_sample(sim, uniforms=uniforms, **{k: v for k, v in gumbel.items() if k != "dim"})   # :388-390
    if training:
        quant = torch.einsum("h n c, h c d -> h n d", onehot, emb)        # :393-395
    else:
        quant = emb.gather(1, ind[..., None].expand(-1, -1, d)) if not emb.requires_grad else \
            torch.stack([emb[h][ind[h]] for h in range(H)], 0)            # :397
    if training and opts.ema_update and not freeze_codebook:              # :399-426
        with torch.no_grad():
            rows = flat.detach()
            if affine is not None:
                rows = (rows - affine.batch_mean) * (codebook_std / batch_std) + affine.codebook_mean
            oh = onehot.detach().clone()
            if flat_mask is not None:
                oh[~flat_mask] = 0.0
            w = 1 - opts.decay
            state.cluster_size.data.lerp_(oh.sum(dim=1), w)
            state.embed_avg.data.lerp_(torch.einsum("h n d, h n c -> h c d", rows, oh).contiguous(), w)
            total = state.cluster_size.sum(dim=-1, keepdim=True)
            smoothed = (state.cluster_size + opts.eps) / (total + K * opts.eps) * total
            new_emb = state.embed_avg / smoothed[..., None]
            if opts.weights_l2norm:
                new_emb = l2norm(new_emb)
            state.embeddings.data.copy_(new_emb)                          # in place: graph aliases see the new values
            expire_codes(state, x.detach(), opts)
    quant = quant.reshape(H, *lead, d)
    ind = ind.reshape(H, *lead)
    if needs_h:
        quant, ind = quant[0], ind[0]
    return quant, ind, sim.reshape(H, *lead, K)                           # :431-435


def vq_forward_variants(state: CodebookState, x: torch.Tensor, opts: VQOpts, *, gumbel: Optional[dict] = None,
                        affine: Optional[AffineState] = None, training: bool = True, mask: Optional[torch.Tensor] = None,
                        learnable: bool = False, codebook_is_parameter: bool = False,
                        uniforms: Optional[torch.Tensor] = None, orthogonal_reg_weight: float = 0.0,
                        orthogonal_reg_active_codes_only: bool = False,
                        orthogonal_reg_max_codes: Optional[int] = None):
    """VectorQuantize.forward (vector_quantize_pytorch.py:182-430) around `codebook_forward_variants`: channel-last,
    no projections, shared-codebook heads.  `learnable` is VectorQuantize's own flag (commit_quantize stays attached,
    :262-268); `codebook_is_parameter` is the Codebook's (the similarities stay attached, codebooks.py:375-377 -- also
    set by the orthogonal loss alone, :95-102).  Returns (quantize, ind, loss[1], (commit, orthogonal))."""
    B, n, D = x.shape
    heads = opts.heads
    xin = x
    if heads > 1:                                                         # :213-219 "b n (h d) -> 1 (b h) n d"
        xin = x.reshape(B, n, heads, D // heads).permute(0, 2, 1, 3).reshape(1, B * heads, n, D // heads)
    if opts.input_l2norm:
        xin = l2norm(xin)                                                 # :221
    quant, ind, _ = codebook_forward_variants(state, xin, opts.codebook, gumbel=gumbel, affine=affine,
                                              training=training, mask=mask, learnable=codebook_is_parameter,
                                              uniforms=uniforms)
    loss = torch.tensor([0.0])
    commit = orth = torch.tensor(0.0)
    if training:
        commit_q = quant if learnable else quant.detach()                 # :262-268
        quant = xin + (quant - xin).detach()                              # :273
    if heads > 1:                                                         # :302-308 "1 (b h) n -> b n h"
        ind = ind.reshape(B, heads, n).permute(0, 2, 1)
    if training:
        if opts.commitment_weight > 0:                                    # :335-364
            if mask is not None:
                per = F.mse_loss(commit_q, xin.float(), reduction="none")
                lm = mask if heads == 1 else mask[None, :, None, :].expand(per.shape[0], B, per.shape[1] // B, n) \
                    .reshape(per.shape[0], per.shape[1], n)
                commit = per[lm].mean()
            else:
                commit = F.mse_loss(commit_q, xin.float())
            loss = loss + commit * opts.commitment_weight
        if orthogonal_reg_weight > 0:                                     # :366-390 (on `embeddings`)
            codebook = state.embeddings
            if orthogonal_reg_active_codes_only:
                codebook = codebook[:, torch.unique(ind)]
            num_codes = codebook.shape[-2]
            if orthogonal_reg_max_codes is not None and num_codes > orthogonal_reg_max_codes:
                codebook = codebook[:, torch.randperm(num_codes)[:orthogonal_reg_max_codes]]
            orth = orthogonal_loss(codebook)
            loss = loss + orth * orthogonal_reg_weight
    if heads > 1:                                                         # :394-404 "1 (b h) n d -> b n (h d)"
        quant = quant.reshape(B, heads, n, D // heads).permute(0, 2, 1, 3).reshape(B, n, D)
    if mask is not None:
        quant = torch.where(mask[..., None], quant, x)                    # :415-418
    return quant, ind, loss, (commit, orth)
