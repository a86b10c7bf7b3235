"""Multi-GPU host logic (one process per GPU, the user's torch.distributed group: NCCL on GPUs, Gloo in CPU tests).

Data parallel (reference utils/distributed.py + codebooks.py:410,415): every rank quantises its own latents against
a replicated codebook; the only exchange per step is ONE all_reduce(SUM) of the packed (H,K,d+1) statistics (done in
Codebook._run).  Dead-code replacement must pick the SAME vectors on every rank or the replicas drift apart; the
reference does that with `sample_vectors_distributed` (utils/distributed.py:55-75), restated here.

Sharded codebook (no reference counterpart; BASELINE config 5): rank r owns codes [r*K/W, (r+1)*K/W); every rank
searches its shard for all latents and an all_reduce(MIN) over packed (score, global index) int64 keys picks the
winner -- smallest score, lowest index on ties, i.e. torch.argmax semantics on the un-sharded codebook.  The path
lives in `Codebook` itself (`Codebook.sharded`, chosen at construction when `codebook_size >= SHARD_MIN_CODES`, a
process group is up and `use_ddp` is set), so `VectorQuantize` / `ResidualVQ` reach it unchanged; this module holds
the collectives it uses.
"""
from __future__ import annotations

from typing import Callable, Optional

import torch
import torch.distributed as dist


SHARD_MIN_CODES = 65536      # north star: "codebooks of 64K entries or more are sharded across GPUs"


def is_distributed() -> bool:
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def world_size() -> int:
    return dist.get_world_size() if is_distributed() else 1


def rank() -> int:
    return dist.get_rank() if is_distributed() else 0


def all_gather_codes(shard: torch.Tensor, dim: int = 1, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Concatenate the ranks' shards along `dim` (rank order) -- the replica of a row-sharded codebook buffer."""
    if not is_distributed():
        return shard
    W = dist.get_world_size()
    shard = shard.contiguous()
    stacked = torch.empty((W,) + tuple(shard.shape), dtype=shard.dtype, device=shard.device)
    dist.all_gather_into_tensor(stacked.view(-1), shard.view(-1))
    full = stacked.movedim(0, dim).reshape(*shard.shape[:dim], W * shard.shape[dim], *shard.shape[dim + 1:])
    if out is not None and out.shape == full.shape and out.dtype == full.dtype:
        out.copy_(full)
        return out
    return full.contiguous()


def all_gather_rows(local: torch.Tensor, dim: int) -> torch.Tensor:
    """Concatenate equally sized per-rank batches along `dim` (rank order)."""
    return all_gather_codes(local, dim=dim)


def all_gather_small(t: torch.Tensor) -> torch.Tensor:
    """(W, *t.shape) on the HOST: the ranks' copies of a tiny tensor (one collective + one host sync)."""
    if not is_distributed():
        return t[None].cpu()
    W = dist.get_world_size()
    out = torch.empty((W,) + tuple(t.shape), dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(out.view(-1), t.contiguous().view(-1))
    return out.cpu()


def broadcast_from_rank0(t: torch.Tensor) -> torch.Tensor:
    if is_distributed():
        dist.broadcast(t, src=0)
    return t


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
