"""``Codebook``: drop-in for reference vector_quantization/codebooks.py:81-435 on the EMA path.

Same constructor, same persistent buffers (``embeddings`` (H,K,d), ``embed_avg`` (H,K,d),
``cluster_size`` (H,K), fp32) and the same ``forward(x, mask, freeze_codebook)`` contract;
the numeric work (search, gather, EMA statistics, refresh, expiry scatter) runs in
libvqb200.so.  Options that are off the named hot path (affine re-parametrisation,
stochastic gumbel sampling) raise NotImplementedError -- there is
no silent fallback.  ``learnable_codebook=True`` makes ``embeddings`` a Parameter that receives the gradient
of the commitment loss / of the gathered codes (a segmented sum over the rows of each code).
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
        if use_affine:
            raise NotImplementedError("vqb200: use_affine is outside the accelerated EMA path")
        gp = asdict(gumbel_params) if is_dataclass(gumbel_params) else dict(gumbel_params)
        if gp.get("stochastic") or gp.get("straight_through") or gp.get("reinmax"):
            raise NotImplementedError("vqb200: stochastic / straight-through gumbel sampling is outside the "
                                      "accelerated path (deterministic argmax only)")

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
        self.use_affine = False

        init = torch.zeros(num_codebooks, codebook_size, dim) if initialization_by_kmeans else \
            _uniform_init(num_codebooks, codebook_size, dim)
        if self.weights_l2norm:
            init = _l2norm_cpu_or_cuda(init)
        self.is_initialized = not initialization_by_kmeans
        self.register_buffer("cluster_size", torch.zeros(num_codebooks, codebook_size))
        self.register_buffer("embed_avg", init.clone())
        if self.learnable_codebook:          # reference codebooks.py:186-190: a Parameter, trained by the user's optimizer
            self.embeddings = nn.Parameter(init)
        else:
            self.register_buffer("embeddings", init)

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

    # ------------------------------------------------------------------ core
    def _run(self, x: torch.Tensor, mask: Optional[torch.Tensor], freeze_codebook: bool, fuse_st: bool,
             want_commit: bool, normalize_input: bool = False, keep_dense: bool = False):
        """Shared by forward() and VectorQuantize.  x: (H, ..., d).  Returns (quantize (H,...,d), idx (H,...),
        commit scalar | None).  `keep_dense`: leave in `self.dense_ctx` what the consumers of the dense similarities
        (cross-entropy to indices, CE commitment, diversity loss) need: the (H,N,d) latents the search saw and the
        codebook as it was BEFORE this forward's EMA step (reference codebooks.py:386 runs before :425)."""
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

        training = self.training
        commit = None
        update = training and self.ema_update and not freeze_codebook
        # learnable codebook (reference codebooks.py:375-377, vector_quantize_pytorch.py:262-268): the commitment loss
        # and the gathered codes are differentiable with respect to `embeddings`
        emb_param = self.embeddings if (self.learnable_codebook and training and not freeze_codebook
                                        and torch.is_grad_enabled() and self.embeddings.requires_grad) else None
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
        """reference codebooks.py:350-435.  Returns (quantize, embed_ind, similarities) with
        similarities=None: the N x K matrix is never materialised on this path."""
        needs_codebook_dim = x.ndim < 4
        if needs_codebook_dim:
            x = x[None]
        quant, ind, _ = self._run(x, mask, freeze_codebook, fuse_st=False, want_commit=False)
        if needs_codebook_dim:
            quant, ind = quant[0], ind[0]
        return quant, ind, None


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
