"""``Codebook``: drop-in for reference vector_quantization/codebooks.py:81-435 on the EMA path.

Same constructor, same persistent buffers (``embeddings`` (H,K,d), ``embed_avg`` (H,K,d),
``cluster_size`` (H,K), fp32) and the same ``forward(x, mask, freeze_codebook)`` contract;
the numeric work (search, gather, EMA statistics, refresh, expiry scatter) runs in
libvqb200.so.  ``learnable_codebook=True`` makes ``embeddings`` a Parameter that receives the gradient
of the commitment loss / of the gathered codes (a segmented sum over the rows of each code).

Variants off the named hot path take `_run_variants` (same kernels, no fused passes):
  * gumbel sampling (reference utils/general.py:107-151, codebooks.py:388): `stochastic` = argmax of
    similarities / temperature + gumbel noise as an epilogue of the tiled fp32 score pass (vqb_dense_gumbel_sample;
    the noise is generated in the kernel from torch's Philox stream, no N x K tensor exists);
    `straight_through` / `reinmax` change only the GRADIENT through the one-hot (their forward value is the hard
    one-hot to 2^-24), which the reference drops unless the codebook is learnable or `Codebook.forward` is used
    directly under autograd -- that corner runs the reference's formulas as row-chunked torch ops (library code, not
    kernels of this package);
  * affine re-parametrisation (codebooks.py:274-348,373-384,400-403): batch moments from one fp64 pass over the latents
    (vqb_column_moments), search / gather on the transformed codebook, EMA sums transformed per code instead of per row.
"""
from __future__ import annotations

from dataclasses import asdict, is_dataclass
from typing import Optional

import torch
import torch.distributed as dist
from torch import nn

from . import _lib, ops
from .params import GumbelParams, KmeansParameters


def _uniform_init(*shape):
    # reference utils/general.py:101-104
    t = torch.empty(shape)
    nn.init.kaiming_uniform_(t)
    return t


def _l2norm_cpu_or_cuda(t: torch.Tensor) -> torch.Tensor:
    # construction-time only (tiny, may be on CPU); the forward path uses the CUDA kernel
    return torch.nn.functional.normalize(t, p=2, dim=-1)


class Codebook(nn.Module):
    def __init__(self, dim, codebook_size, num_codebooks=1, initialization_by_kmeans: bool = False,
                 kmeans_params: KmeansParameters = None, decay: float = 0.8, eps_for_smoothing: float = 1e-5,
                 threshold_ema_dead_code: int = 2, reset_cluster_size: int = None, use_ddp: bool = False,
                 distributed_replace_codes: bool = True, learnable_codebook: bool = False,
                 gumbel_params: GumbelParams = GumbelParams(), ema_update: bool = True, use_affine: bool = False,
                 affine_params=None, transform_input: str = "identity", use_cosine_sim: bool = False,
                 weights_regularization: str = "identity"):
        super().__init__()
        for name, val in (("transform_input", transform_input), ("weights_regularization", weights_regularization)):
            if val not in ("identity", "l2norm"):
                # the reference does `raise f"..."` here (a TypeError); keep the intent, not the bug
                raise ValueError(f"The option {val} for {name} is not implemented")
        gp = asdict(gumbel_params) if is_dataclass(gumbel_params) else dict(gumbel_params)
        # reference codebooks.py:154-158: `sample_fn_training` (the only one forward() calls) = gumbel_sample(**params)
        self.gumbel = {"temperature": 1.0, "stochastic": False, "reinmax": False, "straight_through": False, "dim": -1,
                       "training": True, **gp}
        self._uniform_queue: list = []       # tests: the (H,N,K) draws a fixture was recorded with, in call order

        self.input_l2norm = transform_input == "l2norm"
        self.weights_l2norm = weights_regularization == "l2norm"
        self.use_cosine_sim = use_cosine_sim
        self.decay = decay
        self.ema_update = ema_update
        self.codebook_size = codebook_size
        self.num_codebooks = num_codebooks
        self.dim = dim
        self.kmeans_params = asdict(kmeans_params) if is_dataclass(kmeans_params) else kmeans_params
        self.eps_for_smoothing = eps_for_smoothing
        self.threshold_ema_dead_code = threshold_ema_dead_code
        self.reset_cluster_size = reset_cluster_size if reset_cluster_size is not None else threshold_ema_dead_code
        assert not (use_ddp and num_codebooks > 1 and initialization_by_kmeans), \
            "kmeans init is not compatible with multiple codebooks in distributed environment for now"
        self.use_ddp = use_ddp
        # the reference indexes kmeans_params["sync"] unconditionally and crashes when it is None
        # (codebooks.py:164-166); treat a missing KmeansParameters as its defaults instead
        self._kmeans_sync = bool((self.kmeans_params or {"sync": True})["sync"])
        self.distributed_replace_codes = distributed_replace_codes
        self.learnable_codebook = bool(learnable_codebook)
        # VectorQuantize clears this when ITS learnable_codebook is off while the codebook is a Parameter only because of
        # the orthogonal regularisation (reference vector_quantize_pytorch.py:99-104,262-268: commit_quantize is detached)
        self.commit_grad_to_codebook = True
        self.return_similarities = False       # opt in to the dense (H, ..., K) third return value of forward()
        self._sim: Optional[torch.Tensor] = None
        self.use_affine = bool(use_affine)
        if self.use_affine:
            if affine_params is None:
                raise ValueError("vqb200.Codebook: use_affine=True needs affine_params")
            self.affine_params = asdict(affine_params) if is_dataclass(affine_params) else dict(affine_params)

        init = torch.zeros(num_codebooks, codebook_size, dim) if initialization_by_kmeans else \
            _uniform_init(num_codebooks, codebook_size, dim)
        if self.weights_l2norm:
            init = _l2norm_cpu_or_cuda(init)

        # Sharded codebook (north star config 5; no reference counterpart): with >= SHARD_MIN_CODES codes, a process
        # group up and `use_ddp` set, rank r owns rows [r K/W, (r+1) K/W) of the three buffers.  Every rank draws the
        # same full initialisation (the ranks share the module seed, as replicated data parallel requires too) and
        # keeps its rows; `state_dict()` gathers the full (H,K,d) tensors under the reference's keys and
        # `load_state_dict()` takes this rank's rows of them (both collective-free on load, one all_gather on save).
        from . import distributed as D
        self.sharded = bool(use_ddp and D.is_distributed() and codebook_size >= D.SHARD_MIN_CODES
                            and codebook_size % D.world_size() == 0)
        self.shard_world = D.world_size() if self.sharded else 1
        self.shard_rank = D.rank() if self.sharded else 0
        self.shard_size = codebook_size // self.shard_world
        # "replicated": every rank passes the SAME latents (BASELINE config 5); "all_gather": every rank passes its own
        # batch and the ranks' batches are concatenated before the search (data parallel over a sharded codebook)
        self.sharded_input = "all_gather"
        self._replica: Optional[torch.Tensor] = None
        self._replica_dirty = True
        if self.sharded:
            if learnable_codebook or self.use_affine or any(self._variant_sampling()):
                raise NotImplementedError("vqb200: a sharded codebook (>= %d codes under a process group) is EMA-only "
                                          "with deterministic sampling and no affine re-parametrisation"
                                          % D.SHARD_MIN_CODES)
            lo = self.shard_rank * self.shard_size
            init = init[:, lo:lo + self.shard_size].contiguous()
            self._register_state_dict_hook(Codebook._gather_state_hook)
            self._register_load_state_dict_pre_hook(self._slice_state_pre_hook)
        self.is_initialized = not initialization_by_kmeans
        self.register_buffer("cluster_size", torch.zeros(num_codebooks, self.shard_size))
        self.register_buffer("embed_avg", init.clone())
        if self.learnable_codebook:          # reference codebooks.py:186-190: a Parameter, trained by the user's optimizer
            self.embeddings = nn.Parameter(init)
        else:
            self.register_buffer("embeddings", init)

        if self.use_affine:                    # reference codebooks.py:194-206, same buffer names
            self.register_buffer("batch_mean", None)
            self.register_buffer("batch_variance", None)
            self.register_buffer("codebook_mean_needs_init", torch.Tensor([True]))
            self.register_buffer("codebook_mean", torch.empty(num_codebooks, 1, dim))
            self.register_buffer("codebook_variance_needs_init", torch.Tensor([True]))
            self.register_buffer("codebook_variance", torch.empty(num_codebooks, 1, dim))

        # derived, non-persistent: scaled fp16 copy + norms for the tensor-core search
        self._cache: Optional[torch.Tensor] = None
        self._cache_key = None
        self._dirty = True
        self.dense_ctx = None
        self.last_search_ws: Optional[torch.Tensor] = None
        self.fused_quantize_ema = True     # False: separate gather and EMA-reduce passes (same results)

    # ------------------------------------------------------------------ transforms
    def transform_input(self, x: torch.Tensor) -> torch.Tensor:
        """reference: identity | l2norm (utils/losses.py:19), applied by VectorQuantize before the codebook."""
        if not self.input_l2norm:
            return x
        return ops.l2norm_rows(x.contiguous())

    # ------------------------------------------------------------------ cache
    def _codebook_cache(self) -> torch.Tensor:
        e = self.embeddings
        key = (e.data_ptr(), e._version, tuple(e.shape), self.use_cosine_sim)
        if self._dirty or self._cache is None or key != self._cache_key:
            self._cache = ops.prepare_codebook(e, self.use_cosine_sim, self._cache)
            self._cache_key = key
            self._dirty = False
        return self._cache

    def invalidate_cache(self) -> None:
        self._dirty = True

    # ------------------------------------------------------------------ helpers
    def _all_reduce(self, t: torch.Tensor) -> None:
        if self.use_ddp and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(t)

    @staticmethod
    def _draw_rows(num_rows: int, m: int, device) -> torch.Tensor:
        # reference utils/general.py:62-66 -- same calls on the same (global) generator
        if num_rows >= m:
            return torch.randperm(num_rows, device=device)[:m]
        return torch.randint(0, num_rows, (m,), device=device)

    def _flatten(self, x: torch.Tensor):
        """(H, ..., d) -> contiguous (H, N, d) in a kernel dtype; also returns the leading shape."""
        if x.dtype not in (torch.float32, torch.bfloat16, torch.float16):
            x = x.float()
        H, d = x.shape[0], x.shape[-1]
        lead = tuple(x.shape[1:-1])
        return _lib.aligned(x.reshape(H, -1, d)), lead

    def _expand_mask(self, mask: Optional[torch.Tensor], n_rows: int) -> Optional[torch.Tensor]:
        # reference codebooks.py:361-367: repeat(mask, "b n -> c (b h n)")
        if mask is None:
            return None
        b, n = mask.shape
        rep = n_rows // (b * n)
        m = mask[:, None, :].expand(b, rep, n).reshape(-1)
        return m.to(torch.uint8).contiguous()

    # ------------------------------------------------------------------ kmeans init (reference utils/kmeans.py)
    @torch.no_grad()
    def _kmeans_init(self, flat: torch.Tensor, mask_u8: Optional[torch.Tensor]) -> None:
        H, N, d = flat.shape
        K = self.codebook_size
        data = flat.float()
        if mask_u8 is not None:
            keep = mask_u8.bool()
            data = data[:, keep].contiguous()
            N = data.shape[1]
        iters = int((self.kmeans_params or {"iter": 10})["iter"])
        sync = self.use_ddp and self._kmeans_sync
        from . import distributed as D
        if sync and D.is_distributed():      # reference codebooks.py:164-168: identical centroids on every rank
            cents = torch.stack([D.sample_vectors_distributed(data[h], K, self._draw_rows) for h in range(H)], 0)
        else:
            cents = torch.stack([data[h][self._draw_rows(N, K, data.device)] for h in range(H)], 0).contiguous()
        counts = torch.zeros(H, K, device=data.device)
        for _ in range(iters):
            cache = ops.prepare_codebook(cents, self.use_cosine_sim)
            idx, _, ws = ops.search(data, cents, cache, self.use_cosine_sim)
            stats = ops.ema_reduce(data, idx, None, K, bound_ws=ws)
            counts = stats[..., d].clone()
            if sync:
                self._all_reduce(counts)
            empty = counts == 0
            means = stats[..., :d] / counts.masked_fill(empty, 1.0)[..., None]
            if sync:
                means = means.contiguous()
                self._all_reduce(means)
            if self.use_cosine_sim:
                means = ops.l2norm_rows(means.contiguous())
            cents = torch.where(empty[..., None], cents, means).contiguous()
        self.embeddings.data.copy_(cents)
        self.embed_avg.data.copy_(cents * counts[..., None])
        self.cluster_size.data.copy_(counts)
        self._dirty = True

    # ------------------------------------------------------------------ expiry (reference codebooks.py:230-255)
    @torch.no_grad()
    def expire_codes_(self, flat: torch.Tensor) -> None:
        if self.threshold_ema_dead_code == 0:
            return
        dead = self.cluster_size < self.threshold_ema_dead_code
        # ONE host sync: the per-codebook counts (the reference syncs twice: torch.any at :251, .item() at :234)
        counts = dead.sum(dim=-1).tolist()
        if not any(counts):
            return
        H, N, d = flat.shape
        from . import distributed as D
        synced = self.use_ddp and self._kmeans_sync and self.distributed_replace_codes and D.is_distributed()
        for h in range(H):
            m = int(counts[h])
            if synced:
                # reference codebooks.py:171-175: sample_vectors_distributed -> identical replacements on every rank
                src = D.sample_vectors_distributed(flat[h], m, self._draw_rows)
                rows = torch.arange(m, device=flat.device)
            elif not self.distributed_replace_codes and D.is_distributed():
                # reference codebooks.py:231-239: every rank samples locally (after weights_regularization) and the
                # replacements are the mean over the ranks (not re-normalised)
                src = flat[h][self._draw_rows(N, m, flat.device)].float().contiguous()
                if self.weights_l2norm:
                    src = ops.l2norm_rows(src)
                src = D.maybe_distributed_mean(src)
                ops.expire_scatter(src, torch.arange(m, device=flat.device), float(self.threshold_ema_dead_code),
                                   float(self.reset_cluster_size), False, self.cluster_size.data[h],
                                   self.embed_avg.data[h], self.embeddings.data[h])
                continue
            else:
                src, rows = flat[h], self._draw_rows(N, m, flat.device)
            ops.expire_scatter(src, rows, float(self.threshold_ema_dead_code), float(self.reset_cluster_size),
                               self.weights_l2norm, self.cluster_size.data[h], self.embed_avg.data[h],
                               self.embeddings.data[h])
        self._dirty = True

    # ------------------------------------------------------------------ sharded codebook (config 5)
    @property
    def shard_offset(self) -> int:
        return self.shard_rank * self.shard_size

    @staticmethod
    def _gather_state_hook(module, state_dict, prefix, local_metadata):
        from . import distributed as D
        for name in ("embeddings", "embed_avg", "cluster_size"):
            state_dict[prefix + name] = D.all_gather_codes(getattr(module, name).detach(), dim=1)

    def _slice_state_pre_hook(self, state_dict, prefix, *args):
        lo = self.shard_offset
        for name in ("embeddings", "embed_avg", "cluster_size"):
            t = state_dict.get(prefix + name)
            if t is not None and t.shape[1] == self.codebook_size:
                state_dict[prefix + name] = t[:, lo:lo + self.shard_size]
        self._dirty = self._replica_dirty = True

    def load_full_codebook(self, full: torch.Tensor, cluster_size: float = 1.0) -> None:
        """(H,K,d) or (K,d) codebook -> this rank's rows (or all of them when not sharded); embed_avg = embeddings."""
        full = full if full.ndim == 3 else full[None]
        lo = self.shard_offset
        rows = full[:, lo:lo + self.shard_size].to(self.embeddings.device, torch.float32)
        with torch.no_grad():
            self.embeddings.copy_(rows); self.embed_avg.copy_(rows); self.cluster_size.fill_(cluster_size)
        self._dirty = self._replica_dirty = True

    def full_codebook(self) -> torch.Tensor:
        """(H,K,d) replica of the sharded `embeddings` (one all_gather when a shard changed since the last call)."""
        if not self.sharded:
            return self.embeddings.detach()
        if self._replica_dirty or self._replica is None:
            from . import distributed as D
            self._replica = D.all_gather_codes(self.embeddings.detach(), dim=1, out=self._replica)
            self._replica_dirty = False
        return self._replica

    def _shard_search(self, flat: torch.Tensor, emb: torch.Tensor, cache) -> torch.Tensor:
        """Global nearest code of every row: local shard search -> (fp32 score of the reference recipe, global index)
        packed into int64 keys -> all_reduce(MIN): smallest score, lowest index on ties = torch.argmax on the
        un-sharded similarities.  Every rank evaluates the score with the same fp64-accumulated recipe, so the
        cross-shard comparison is consistent."""
        from . import distributed as D
        H, N, _ = flat.shape
        idx, score, ws = ops.search(flat, emb, cache, self.use_cosine_sim, idx_offset=self.shard_offset,
                                    want_score=True)
        self.last_search_ws = ws
        keys = ops.minkey_pack(score.reshape(-1), idx.reshape(-1))
        D.merge_min_keys(keys)
        gidx, _ = ops.minkey_unpack(keys)
        return gidx.reshape(H, N)

    @torch.no_grad()
    def _kmeans_init_sharded(self, flat: torch.Tensor, mask_u8: Optional[torch.Tensor]) -> None:
        """reference utils/kmeans.py:38-120 on a row-sharded set of K centroids: rank 0's draw picks the K start rows
        (all ranks hold the same latents), every iteration is a sharded search + segment means on the owned rows."""
        from . import distributed as D
        H, N, d = flat.shape
        data = flat.float()
        if mask_u8 is not None:
            data = data[:, mask_u8.bool()].contiguous()
            N = data.shape[1]
        iters = int((self.kmeans_params or {"iter": 10})["iter"])
        lo, Ks = self.shard_offset, self.shard_size
        cents = []
        for h in range(H):
            rows = D.broadcast_from_rank0(self._draw_rows(N, self.codebook_size, data.device).to(torch.long))
            cents.append(data[h][rows[lo:lo + Ks]])
        cents = torch.stack(cents, 0).contiguous()
        counts = torch.zeros(H, Ks, device=data.device)
        for _ in range(iters):
            cache = ops.prepare_codebook(cents, self.use_cosine_sim)
            gidx = self._shard_search(data, cents, cache)
            mine = (gidx >= lo) & (gidx < lo + Ks)
            stats = torch.stack([ops.ema_reduce(data[h:h + 1], (gidx[h:h + 1] - lo).clamp_(0, Ks - 1),
                                                mine[h].to(torch.uint8), Ks)[0] for h in range(H)], 0)
            counts = stats[..., d].clone()
            empty = counts == 0
            means = stats[..., :d] / counts.masked_fill(empty, 1.0)[..., None]
            if self.use_cosine_sim:
                means = ops.l2norm_rows(means.contiguous())
            cents = torch.where(empty[..., None], cents, means).contiguous()
        self.embeddings.data.copy_(cents)
        self.embed_avg.data.copy_(cents * counts[..., None])
        self.cluster_size.data.copy_(counts)
        self._dirty = self._replica_dirty = True

    @torch.no_grad()
    def _expire_codes_sharded(self, flat: torch.Tensor) -> None:
        """reference codebooks.py:230-255 on shards: the dead codes of all shards, in ascending GLOBAL code order, take
        the rows rank 0 draws with the reference's calls (utils/general.py:62-66) -- exactly what one process does on
        the un-sharded codebook with rank 0's generator.  One all_gather of the per-shard dead counts (the host sync
        the reference has too) and one broadcast of the drawn row ids."""
        if self.threshold_ema_dead_code == 0:
            return
        from . import distributed as D
        H, N, d = flat.shape
        dead = self.cluster_size < self.threshold_ema_dead_code
        counts = D.all_gather_small(dead.sum(dim=-1))                  # (W, H) on the host
        if int(counts.sum()) == 0:
            return
        for h in range(H):
            m_total = int(counts[:, h].sum())
            if m_total == 0:
                continue
            rows = D.broadcast_from_rank0(self._draw_rows(N, m_total, flat.device).to(torch.long))
            first = int(counts[:self.shard_rank, h].sum())
            m = int(counts[self.shard_rank, h])
            if m:
                ops.expire_scatter(flat[h], rows[first:first + m], float(self.threshold_ema_dead_code),
                                   float(self.reset_cluster_size), self.weights_l2norm, self.cluster_size.data[h],
                                   self.embed_avg.data[h], self.embeddings.data[h])
        self._dirty = self._replica_dirty = True

    def _run_sharded(self, x, mask, freeze_codebook, fuse_st, want_commit, normalize_input):
        """`_run` for a row-sharded codebook.  Every rank ends up with the quantised vectors / indices of ITS latents
        (all of them when the input is replicated) and updates ITS rows of the EMA buffers from the rows they won --
        no statistics all_reduce; the only exchanges are the min-key all_reduce, one scalar all_reduce for the Laplace
        normaliser and the all_gather that refreshes the codebook replica the gather reads."""
        from . import distributed as D
        _lib.require_device(x)
        flat, lead = self._flatten(x)
        H, n_local, d = flat.shape
        if H != self.num_codebooks or d != self.embeddings.shape[-1]:
            raise ValueError(f"vqb200.Codebook: input {tuple(x.shape)} does not match codebook "
                             f"(H={self.num_codebooks}, d={self.embeddings.shape[-1]})")
        if torch.is_grad_enabled() and flat.requires_grad and normalize_input:
            flat = ops.l2norm_rows_autograd(flat)
        elif normalize_input:
            flat = ops.l2norm_rows(flat)
        mask_u8 = self._expand_mask(mask, n_local)
        gathered = self.sharded_input == "all_gather"
        rows_all, mask_all = flat.detach(), mask_u8
        if gathered:
            rows_all = D.all_gather_rows(flat.detach(), dim=1)
            if mask_u8 is not None:
                mask_all = D.all_gather_rows(mask_u8, dim=0)
        N = rows_all.shape[1]
        if not self.is_initialized:
            self._kmeans_init_sharded(rows_all, mask_all)
            self.is_initialized = True

        gidx_all = self._shard_search(rows_all, self.embeddings.detach(), self._codebook_cache())
        replica = self.full_codebook()
        training = self.training
        update = training and self.ema_update and not freeze_codebook
        if update and torch.is_grad_enabled() and flat.requires_grad:
            replica = replica.clone()          # backward reads the PRE-update codes (see _run)
        lo_r = self.shard_rank * n_local if gathered else 0
        gidx = gidx_all[:, lo_r:lo_r + n_local] if gathered else gidx_all
        commit = None
        stats_full = None
        whole = update and not gathered and mask_u8 is None and self.fused_quantize_ema and ops.quantize_ema_supported(d)
        if whole and fuse_st:
            # replicated input: gather/ST/loss over all rows and the EMA sums of ALL codes in one pass; this rank
            # keeps the statistics of its own rows of the codebook
            quant, commit, stats_full = ops.quantize_training(flat, replica, gidx, None, want_commit, ema=True,
                                                              bound_ws=self.last_search_ws)
        elif whole:
            with torch.no_grad():
                quant, _, stats_full = ops.quantize_ema(flat, replica, gidx, False, False, bound_ws=self.last_search_ws)
        elif fuse_st and training:
            quant, commit, _ = ops.quantize_training(flat, replica, gidx.contiguous(), mask_u8, want_commit)
        else:
            with torch.no_grad():
                quant, _ = ops.gather_st_loss(flat, replica, gidx.contiguous(), None, False, False)
        if update:
            with torch.no_grad():
                lo, Ks = self.shard_offset, self.shard_size
                if stats_full is not None:
                    stats = stats_full[:, lo:lo + Ks].contiguous()
                else:
                    mine = (gidx_all >= lo) & (gidx_all < lo + Ks)
                    if mask_all is not None:
                        mine = mine & mask_all.bool()[None]
                    stats = torch.stack([ops.ema_reduce(rows_all[h:h + 1], (gidx_all[h:h + 1] - lo).clamp_(0, Ks - 1),
                                                        mine[h].to(torch.uint8), Ks)[0] for h in range(H)], 0)
                ops.ema_apply_sharded(stats, self.cluster_size.data, self.embed_avg.data, self.embeddings.data,
                                      1 - self.decay, self.eps_for_smoothing, self.weights_l2norm, self.codebook_size,
                                      D.all_reduce_sum)
                self._dirty = self._replica_dirty = True
                self._expire_codes_sharded(rows_all)
        return quant.reshape(H, *lead, d), gidx.reshape(H, *lead), commit

    # ------------------------------------------------------------------ variants: gumbel sampling, affine
    def _variant_sampling(self):
        """(stochastic, straight_through) as reference gumbel_sample applies them (utils/general.py:121-137)."""
        g = self.gumbel
        active = bool(g["training"]) and g["temperature"] > 0
        return (active and bool(g["stochastic"])), (active and bool(g["straight_through"]))

    def _variants_active(self) -> bool:
        stochastic, st = self._variant_sampling()
        return self.use_affine or stochastic or st

    def _update_with_decay(self, name: str, new_value: torch.Tensor, decay: float) -> None:
        # reference codebooks.py:258-272
        old = getattr(self, name)
        needs_init = getattr(self, name + "_needs_init", None)
        first = needs_init is not None and bool(needs_init.item())
        if first:
            needs_init.fill_(0.0)
        if old is None or first:
            setattr(self, name, new_value.detach().clone())
            return
        setattr(self, name, old * decay + new_value.detach() * (1 - decay))

    @torch.no_grad()
    def _update_affine(self, flat: torch.Tensor, mask_u8: Optional[torch.Tensor]) -> None:
        """reference codebooks.py:274-348.  The batch moments come from ONE fp64 pass over the latents; with
        `affine_params.sync` the per-rank sums are all-reduced (mean and variance of the union of the batches, as the
        reference's three all_reduce calls compute)."""
        ap = self.affine_params
        emb = self.embeddings.detach()
        if self.training:
            self._update_with_decay("codebook_mean", emb.mean(dim=1, keepdim=True), ap["codebook_decay"])
            self._update_with_decay("codebook_variance", emb.var(dim=1, unbiased=False, keepdim=True),
                                    ap["codebook_decay"])
        sums, rows = ops.column_moments(flat.detach(), mask_u8)
        if ap["sync"]:
            from . import distributed as D
            packed = torch.cat([sums.reshape(-1), rows.double()])
            D.all_reduce_sum(packed)
            sums, rows = packed[:sums.numel()].reshape(sums.shape), packed[sums.numel():]
        n = rows.double()[:, None]
        mean = sums[..., 0] / n
        var = (sums[..., 1] / n - mean * mean).clamp_min(0.0)
        self._update_with_decay("batch_mean", mean.float()[:, None, :], ap["batch_decay"])
        self._update_with_decay("batch_variance", var.float()[:, None, :], ap["batch_decay"])

    def _gumbel_st_quantize(self, flat: torch.Tensor, emb_g: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
        """quantize = one_hot' @ embeddings with the straight-through / reinmax one-hot of reference
        utils/general.py:139-149, as torch ops under autograd (library code): only the gradient through the one-hot
        differs from a gather.  Row-chunked (bounds the size of each temporary, not the total the graph keeps);
        `reinmax` cannot be chunked because the reference normalises `prob1` over dim=1 -- the ROW axis of (h, n, c) --
        which couples all rows.  No recomputation in backward: with an EMA codebook `emb_g` aliases the live buffer,
        which this forward's EMA step overwrites in place before backward runs -- the reference's graph has exactly that
        aliasing (its saved `embeddings.detach()` shares the buffer's storage, codebooks.py:375-377,425), so the same
        torch ops on the same alias reproduce its gradient: pre-update distances and probabilities, post-update code
        vectors."""
        import torch.nn.functional as F
        g = self.gumbel
        tau, K, cos = float(g["temperature"]), emb_g.shape[1], self.use_cosine_sim

        def log_eps(t):
            return t.clamp(min=1e-5).log()

        def piece(xc, e, ic):
            xc = xc.float()
            logits = torch.einsum("hnd,hcd->hnc", xc, e) if cos else -torch.cdist(xc, e)
            one_hot = F.one_hot(ic, K).type(logits.dtype)
            if g["reinmax"]:
                prob0 = logits.softmax(dim=-1)
                prob1 = (one_hot + (logits / tau).softmax(dim=-1)) / 2
                prob1 = ((log_eps(prob1) - logits).detach() + logits).softmax(dim=1)
                prob2 = 2 * prob1 - 0.5 * prob0
                one_hot = prob2 - prob2.detach() + one_hot
            else:
                prob1 = (logits / tau).softmax(dim=-1)
                one_hot = one_hot + prob1 - prob1.detach()
            return torch.einsum("hnc,hcd->hnd", one_hot, e)

        N = flat.shape[1]
        rows = N if g["reinmax"] else max(1, (1 << 24) // max(K, 1))
        if rows >= N:
            return piece(flat, emb_g, idx)
        outs = [piece(flat[:, lo:lo + rows], emb_g, idx[:, lo:lo + rows]) for lo in range(0, N, rows)]
        return torch.cat(outs, dim=1)

    def _run_variants(self, x, mask, freeze_codebook, fuse_st, want_commit, normalize_input, keep_dense):
        """`_run` with the affine re-parametrisation and / or gumbel sampling (see the module docstring): the same
        kernels, un-fused; returns what `_run` returns."""
        g = self.gumbel
        assert not (g["reinmax"] and not g["straight_through"]), \
            "reinmax can only be turned on if using straight through gumbel softmax"
        _lib.require_device(x)
        flat, lead = self._flatten(x)
        H, N, d = flat.shape
        K = self.codebook_size
        if H != self.num_codebooks or d != self.embeddings.shape[-1]:
            raise ValueError(f"vqb200.Codebook: input {tuple(x.shape)} does not match codebook "
                             f"{tuple(self.embeddings.shape)}")
        mask_u8 = self._expand_mask(mask, N)
        if normalize_input:
            flat = ops.l2norm_rows_autograd(flat) if (torch.is_grad_enabled() and flat.requires_grad) \
                else ops.l2norm_rows(flat)
        if not self.is_initialized:
            self._kmeans_init(flat, mask_u8)
            self.is_initialized = True

        training = self.training
        update = training and self.ema_update and not freeze_codebook
        grad_on = torch.is_grad_enabled()
        emb_param = self.embeddings if (self.learnable_codebook and self.commit_grad_to_codebook and training
                                        and not freeze_codebook and grad_on and self.embeddings.requires_grad) else None
        base = emb_param if emb_param is not None else self.embeddings.detach()
        inv_scale = cm = bm = None
        if self.use_affine:
            self._update_affine(flat, mask_u8)
            codebook_std = self.codebook_variance.clamp(min=1e-5).sqrt()
            batch_std = self.batch_variance.clamp(min=1e-5).sqrt()
            cm, bm, inv_scale = self.codebook_mean, self.batch_mean, codebook_std / batch_std
            emb_eff = (base - cm) * (batch_std / codebook_std) + bm          # reference :381-384
        else:
            emb_eff = base
        emb_val = emb_eff.detach().contiguous()
        if update and emb_val.data_ptr() == self.embeddings.data_ptr() and (keep_dense or (grad_on and flat.requires_grad)):
            emb_val = emb_val.clone()          # backward reads the PRE-update codes (see _run)
        if keep_dense:
            self.dense_ctx = ops._DenseCtx(flat, emb_val, emb_val if self.use_affine else self.embeddings,
                                           self.use_cosine_sim)

        stochastic, st = self._variant_sampling()
        x_det = flat.detach()
        if stochastic:
            u = self._uniform_queue.pop(0) if self._uniform_queue else None
            idx = ops.dense_gumbel_sample(x_det, emb_val, self.use_cosine_sim, float(g["temperature"]), uniforms=u)
            self.last_search_ws = None
        else:
            cache = ops.prepare_codebook(emb_val, self.use_cosine_sim) if self.use_affine else self._codebook_cache()
            idx, _, ws = ops.search(x_det, emb_val, cache, self.use_cosine_sim)
            self.last_search_ws = ws
        if self.return_similarities and not fuse_st:
            self._sim = ops.dense_scores(x_det, emb_val, self.use_cosine_sim)

        commit = None
        wants_x = grad_on and flat.requires_grad
        need_graph = st and training and grad_on and (emb_param is not None if fuse_st
                                                      else (wants_x or emb_param is not None))
        if need_graph:
            emb_graph = emb_eff if (emb_param is not None or self.use_affine) else self.embeddings.detach()
            q_graph = self._gumbel_st_quantize(flat, emb_graph, idx)
            if fuse_st:     # reference vector_quantize_pytorch.py:262-273,335-362 with commit_quantize attached
                xf = flat.float()
                if want_commit:
                    per = (q_graph - xf) ** 2
                    commit = per[:, mask_u8.bool()].mean() if mask_u8 is not None else per.mean()
                quant = xf + (q_graph - xf).detach()
            else:
                quant = q_graph
        elif fuse_st and training:
            quant, commit, _ = ops.quantize_training(flat, emb_val if emb_param is None else emb_eff.contiguous(), idx,
                                                     mask_u8, want_commit)
        elif emb_param is not None:
            quant = ops.gather_codes(emb_eff.contiguous(), idx)
        else:
            with torch.no_grad():
                quant, _ = ops.gather_st_loss(x_det, emb_val, idx, None, False, False)

        if update:
            with torch.no_grad():
                stats = ops.ema_reduce(x_det, idx, mask_u8, K)
                if self.use_affine:
                    # reference :400-403 transforms every latent, (x - batch_mean) * (codebook_std / batch_std) +
                    # codebook_mean, before the sums; the map is affine, so the per-code sums transform the same way
                    cnt = stats[..., d:]
                    stats[..., :d] = (stats[..., :d] - cnt * bm) * inv_scale + cnt * cm
                self._all_reduce(stats)
                ops.ema_apply(stats, self.cluster_size.data, self.embed_avg.data, self.embeddings.data,
                              1 - self.decay, self.eps_for_smoothing, self.weights_l2norm)
                self._dirty = True
                self.expire_codes_(x_det)
        return quant.reshape(H, *lead, d), idx.reshape(H, *lead), commit

    # ------------------------------------------------------------------ core
    def _run(self, x: torch.Tensor, mask: Optional[torch.Tensor], freeze_codebook: bool, fuse_st: bool,
             want_commit: bool, normalize_input: bool = False, keep_dense: bool = False):
        """Shared by forward() and VectorQuantize.  x: (H, ..., d).  Returns (quantize (H,...,d), idx (H,...),
        commit scalar | None).  `keep_dense`: leave in `self.dense_ctx` what the consumers of the dense similarities
        (cross-entropy to indices, CE commitment, diversity loss) need: the (H,N,d) latents the search saw and the
        codebook as it was BEFORE this forward's EMA step (reference codebooks.py:386 runs before :425)."""
        if self.sharded:
            if keep_dense:
                raise NotImplementedError("vqb200: the dense-similarity losses are not available on a sharded codebook")
            return self._run_sharded(x, mask, freeze_codebook, fuse_st, want_commit, normalize_input)
        if self._variants_active():
            return self._run_variants(x, mask, freeze_codebook, fuse_st, want_commit, normalize_input, keep_dense)
        _lib.require_device(x)        # raises for a CPU tensor: there is no CPU implementation
        flat, lead = self._flatten(x)
        H, N, d = flat.shape
        if H != self.num_codebooks or d != self.embeddings.shape[-1]:
            raise ValueError(f"vqb200.Codebook: input {tuple(x.shape)} does not match codebook "
                             f"{tuple(self.embeddings.shape)}")
        mask_u8 = self._expand_mask(mask, N)

        prepared = False
        if normalize_input:
            # transform_input="l2norm" (reference vector_quantize_pytorch.py:221): normalisation and the search's operand
            # preparation share one pass over x when the codebook cache is already final
            if torch.is_grad_enabled() and flat.requires_grad:
                # the encoder's gradient passes through the normalisation: same kernel forward (so the values do not
                # depend on whether a gradient is wanted), F.normalize's Jacobian on the way back
                flat = ops.l2norm_rows_autograd(flat)
            elif self.is_initialized and ops.l2norm_prepare_supported(d):
                flat = ops.l2norm_prepare(flat, self.codebook_size, self._codebook_cache())
                prepared = True
            else:
                flat = ops.l2norm_rows(flat)

        if not self.is_initialized:
            self._kmeans_init(flat, mask_u8)
            self.is_initialized = True

        emb = self.embeddings.detach()
        update = self.training and self.ema_update and not freeze_codebook
        if update and (keep_dense or (torch.is_grad_enabled() and flat.requires_grad)):
            # this forward's EMA step overwrites `embeddings` in place before backward runs; the reference's
            # commitment loss holds the gathered PRE-update codes (vector_quantize_pytorch.py:262-268, a tensor
            # materialised at codebooks.py:395 before :425), so everything saved for backward reads a snapshot
            emb = emb.clone()
        if keep_dense:
            self.dense_ctx = ops._DenseCtx(flat, emb, self.embeddings, self.use_cosine_sim)
        idx, _, ws = ops.search(flat, emb, self._codebook_cache(), self.use_cosine_sim, latents_prepared=prepared)
        self.last_search_ws = ws
        if self.return_similarities and not fuse_st:
            self._sim = ops.dense_scores(flat.detach(), emb, self.use_cosine_sim)   # before the EMA step, as :386

        training = self.training
        commit = None
        update = training and self.ema_update and not freeze_codebook
        # learnable codebook (reference codebooks.py:375-377, vector_quantize_pytorch.py:262-268): the commitment loss
        # and the gathered codes are differentiable with respect to `embeddings`
        emb_param = self.embeddings if (self.learnable_codebook and self.commit_grad_to_codebook and training
                                        and not freeze_codebook and torch.is_grad_enabled()
                                        and self.embeddings.requires_grad) else None
        # un-masked training step: gather/ST/loss and the EMA sums share ONE pass over the latents
        fused = update and mask_u8 is None and self.fused_quantize_ema and ops.quantize_ema_supported(d) \
            and emb_param is None
        stats = None
        if fused and fuse_st:
            quant, commit, stats = ops.quantize_training(flat, emb, idx, None, want_commit, ema=True, bound_ws=ws)
        elif fused:
            with torch.no_grad():
                quant, _, stats = ops.quantize_ema(flat, emb, idx, False, False, bound_ws=ws)
        elif fuse_st and training:
            quant, commit, _ = ops.quantize_training(flat, emb if emb_param is None else emb_param, idx, mask_u8,
                                                     want_commit)
        elif emb_param is not None:
            quant = ops.gather_codes(emb_param, idx)
        else:
            with torch.no_grad():
                quant, _ = ops.gather_st_loss(flat, emb, idx, None, False, False)

        if update:
            with torch.no_grad():
                if stats is None:
                    stats = ops.ema_reduce(flat, idx, mask_u8, self.codebook_size, bound_ws=ws)
                self._all_reduce(stats)                       # one packed (H,K,d+1) allreduce (reference: two)
                ops.ema_apply(stats, self.cluster_size.data, self.embed_avg.data, self.embeddings.data,
                              1 - self.decay, self.eps_for_smoothing, self.weights_l2norm)
                self._dirty = True
                self.expire_codes_(flat)

        return quant.reshape(H, *lead, d), idx.reshape(H, *lead), commit

    @torch.amp.autocast(device_type="cuda", enabled=False)
    def forward(self, x, mask=None, freeze_codebook=False):
        """reference codebooks.py:350-435.  Returns (quantize, embed_ind, similarities).  `similarities` is None unless
        `self.return_similarities` is set: the (H, ..., K) matrix is not needed by anything on this path and is only
        materialised on request (one extra fp32 score pass, vqb_dense_scores); like the reference's it keeps the leading
        codebook axis even when the input had none (:431-433)."""
        needs_codebook_dim = x.ndim < 4
        if needs_codebook_dim:
            x = x[None]
        self._sim = None
        quant, ind, _ = self._run(x, mask, freeze_codebook, fuse_st=False, want_commit=False)
        sim, self._sim = self._sim, None
        if sim is not None:
            sim = sim.reshape(*ind.shape, sim.shape[-1])
        if needs_codebook_dim:
            quant, ind = quant[0], ind[0]
        return quant, ind, sim


class EuclideanCodebook(Codebook):
    """Upstream name for ``Codebook(use_cosine_sim=False)`` (reference Changelog.md:11-12)."""

    def __init__(self, dim, codebook_size, **kw):
        kw.setdefault("use_cosine_sim", False)
        super().__init__(dim, codebook_size, **kw)


class CosineSimCodebook(Codebook):
    """Upstream name for the cosine codebook: dot similarity on l2-normalised inputs and codes."""

    def __init__(self, dim, codebook_size, **kw):
        kw.setdefault("use_cosine_sim", True)
        kw.setdefault("transform_input", "l2norm")
        kw.setdefault("weights_regularization", "l2norm")
        super().__init__(dim, codebook_size, **kw)
