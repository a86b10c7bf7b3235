"""Multi-GPU host logic (one process per GPU, the user's torch.distributed group: NCCL on GPUs, Gloo in CPU tests).

Data parallel (reference utils/distributed.py + codebooks.py:410,415): every rank quantises its own latents against
a replicated codebook; the only exchange per step is ONE all_reduce(SUM) of the packed (H,K,d+1) statistics (done in
Codebook._run).  Dead-code replacement must pick the SAME vectors on every rank or the replicas drift apart; the
reference does that with `sample_vectors_distributed` (utils/distributed.py:55-75), restated here.

Sharded codebook (no reference counterpart; BASELINE config 5): rank r owns codes [r*K/W, (r+1)*K/W); every rank
searches its shard for all latents and an all_reduce(MIN) over packed (score, global index) int64 keys picks the
winner -- smallest score, lowest index on ties, i.e. torch.argmax semantics on the un-sharded codebook.
"""
from __future__ import annotations

from typing import Callable, Optional

import torch
import torch.distributed as dist


def is_distributed() -> bool:
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def all_reduce_sum(t: torch.Tensor, group=None) -> None:
    if is_distributed():
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)


def maybe_distributed_mean(t: torch.Tensor, group=None) -> torch.Tensor:
    """reference utils/distributed.py:86-92: mean over the ranks when a process group is up (used for the locally
    sampled replacement codes when `distributed_replace_codes=False`, codebooks.py:238-239)."""
    if not is_distributed():
        return t
    t = t.contiguous()
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t / dist.get_world_size(group)


def merge_min_keys(keys: torch.Tensor, group=None) -> torch.Tensor:
    """all_reduce(MIN) of int64 (score, index) keys; in place."""
    if is_distributed():
        dist.all_reduce(keys, op=dist.ReduceOp.MIN, group=group)
    return keys


# ---- replica-consistent sampling ----------------------------------------------------------------------------------
def sample_vectors_distributed(local: torch.Tensor, num: int, draw_rows: Callable, group=None) -> torch.Tensor:
    """`num` vectors drawn from the union of all ranks' `local` (N_r, d) rows; identical result on every rank.

    Same purpose as reference utils/distributed.py:55-75 (all_gather sizes -> rank-0 multinomial on the CPU ->
    broadcast -> per-rank sampling -> W variable-size broadcasts), with three collectives and no host round trip
    except rank 0's row count: rank 0 draws GLOBAL row ids with the very calls a single process would make on the
    rank-concatenated batch (utils/general.py:62-66), broadcasts them, and one all_reduce(SUM) assembles the rows
    (every rank contributes the rows it owns, zeros elsewhere).  So W ranks replace dead codes exactly like one
    process on the concatenated batch with the same generator state.
    """
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    dev = local.device
    n_local = local.shape[0]
    sizes = [torch.empty(1, dtype=torch.long, device=dev) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([n_local], dtype=torch.long, device=dev), group=group)
    sizes = torch.cat(sizes)
    ends = torch.cumsum(sizes, 0)
    start = ends[rank] - sizes[rank]                       # device scalar: no host sync on ranks != 0
    if rank == 0:
        rows = draw_rows(int(ends[-1].item()), num, dev).to(torch.long)
    else:
        rows = torch.empty(num, dtype=torch.long, device=dev)
    dist.broadcast(rows, src=0, group=group)
    mine = (rows >= start) & (rows < start + n_local)
    picked = local[(rows - start).clamp_(0, max(n_local - 1, 0))].float()
    out = torch.where(mine[:, None], picked, torch.zeros((), dtype=torch.float32, device=dev))
    dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
    return out


# ---- sharded codebook -----------------------------------------------------------------------------------------
class ShardedCodebook(torch.nn.Module):
    """Codebook of `codebook_size` codes whose rows are split contiguously over the ranks of `group`.

    forward(x): x (N,d) or (B,n,d) replicated on every rank (all_gather it first if it is sharded).  Returns
    (quantize fp32, indices int64 -- global code ids, commit loss) like a single-device codebook in eval/training
    mode; in training mode each rank applies the EMA update to ITS shard from the rows it won.
    """

    def __init__(self, dim: int, codebook_size: int, decay: float = 0.8, eps_for_smoothing: float = 1e-5,
                 use_cosine_sim: bool = False, group=None, rank: Optional[int] = None, world: Optional[int] = None):
        super().__init__()
        self.group = group
        self.world = world if world is not None else (dist.get_world_size(group) if is_distributed() else 1)
        self.rank = rank if rank is not None else (dist.get_rank(group) if is_distributed() else 0)
        assert codebook_size % self.world == 0, "codebook_size must divide evenly over the ranks"
        self.codebook_size, self.shard_size, self.dim = codebook_size, codebook_size // self.world, dim
        self.decay, self.eps, self.use_cosine_sim = decay, eps_for_smoothing, use_cosine_sim
        self.register_buffer("embeddings", torch.zeros(1, self.shard_size, dim))
        self.register_buffer("embed_avg", torch.zeros(1, self.shard_size, dim))
        self.register_buffer("cluster_size", torch.zeros(1, self.shard_size))
        self._cache = None
        self._dirty = True

    @property
    def offset(self) -> int:
        return self.rank * self.shard_size

    def load_full_codebook(self, full: torch.Tensor) -> None:
        """Take this rank's rows of a (K,d) codebook; embed_avg = embeddings, cluster_size = 1."""
        rows = full[self.offset:self.offset + self.shard_size].to(self.embeddings.device, torch.float32)
        self.embeddings.copy_(rows[None]); self.embed_avg.copy_(rows[None]); self.cluster_size.fill_(1.0)
        self._dirty = True

    def gather_full_codebook(self) -> torch.Tensor:
        if self.world == 1:
            return self.embeddings[0]
        parts = [torch.empty_like(self.embeddings[0]) for _ in range(self.world)]
        dist.all_gather(parts, self.embeddings[0].contiguous(), group=self.group)
        return torch.cat(parts, 0)

    @torch.no_grad()
    def forward(self, x: torch.Tensor):
        from . import ops
        shape = x.shape
        flat = x.reshape(1, -1, shape[-1]).contiguous()
        if self._dirty or self._cache is None:
            self._cache = ops.prepare_codebook(self.embeddings, self.use_cosine_sim, self._cache)
            self._dirty = False
        # local shard search -> (exact score, global index) -> cross-GPU min
        idx, score, ws = ops.search(flat, self.embeddings, self._cache, self.use_cosine_sim, idx_offset=self.offset,
                                    want_score=True)
        keys = ops.minkey_pack(score.reshape(-1), idx.reshape(-1))
        merge_min_keys(keys, self.group)
        gidx, _ = ops.minkey_unpack(keys)
        gidx = gidx.reshape(1, -1)
        full = self.gather_full_codebook()[None].contiguous()            # (1,K,d) replica for the gather
        quant, loss = ops.gather_st_loss(flat, full, gidx, None, self.training, self.training)
        if self.training:
            mine = ((gidx >= self.offset) & (gidx < self.offset + self.shard_size)).reshape(-1)
            local_idx = (gidx - self.offset).clamp_(0, self.shard_size - 1)
            stats = ops.ema_reduce(flat, local_idx, mine.to(torch.uint8), self.shard_size, bound_ws=ws)
            ops.ema_apply_sharded(stats, self.cluster_size, self.embed_avg, self.embeddings, 1 - self.decay, self.eps,
                                  False, self.codebook_size, lambda t: all_reduce_sum(t, self.group))
            self._dirty = True
        commit = loss[0] if loss is not None else torch.zeros((), device=x.device)
        return quant.reshape(shape), gidx.reshape(shape[:-1]), commit
