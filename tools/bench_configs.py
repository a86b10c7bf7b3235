"""Per-config measurements for BASELINE.json's five configs on ONE GPU (C5 = the per-GPU shard of the 8-GPU layout plus
the un-sharded 1-GPU case): step time, lookups/s, search-kernel TFLOP/s and the achieved HBM GB/s of the bandwidth
kernels against their algorithmic bytes.  Writes one JSON object per config to stdout."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "vector-quantization-by-ml_b200"))
import torch

from vqb200 import CodebookParams, ResidualVQ, VectorQuantize, ops

dev = torch.device("cuda:0")
PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) \
    else {"bf16_tflops": 1590.0, "hbm_gbs": 6650.0}
g = torch.Generator(device=dev).manual_seed(1)


def timed(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def seed(cb, scale=0.5, l2=False):
    c = torch.randn(cb.embeddings.shape, generator=g, device=dev) * scale
    if l2:
        c = torch.nn.functional.normalize(c, dim=-1)
    cb.embeddings.copy_(c); cb.embed_avg.copy_(c); cb.cluster_size.fill_(1.0); cb.invalidate_cache()


def op_level(x, c, cos, K):
    """search kernel TFLOP/s + gather / EMA achieved GB/s on (1,N,d) latents."""
    H, N, d = x.shape
    esz = x.element_size()
    cache = ops.prepare_codebook(c, cos)
    ops.TIME_SEARCH_KERNEL = True
    ops.search_kernel_times_ms()
    t_search = timed(lambda: ops.search(x, c, cache, cos), iters=5, warm=2)
    tc = ops.search_kernel_times_ms()[2:]
    ops.TIME_SEARCH_KERNEL = False
    tc_ms = sum(tc) / len(tc)
    idx, _, ws = ops.search(x, c, cache, cos)
    st = ops.search_stats(ws)
    t_gather = timed(lambda: ops.gather_st_loss(x, c, idx, None, True, True), iters=5, warm=2)
    t_ema = timed(lambda: ops.ema_reduce(x, idx, None, K, bound_ws=ws), iters=5, warm=2)
    t_fused = None
    if ops.quantize_ema_supported(d):
        t_fused = timed(lambda: ops.quantize_ema(x, c, idx, True, True, bound_ws=ws), iters=5, warm=2)
    gather_bytes = N * (esz * d + 4 * d + 8)
    ema_bytes = N * (esz * d + 8) + K * (d + 1) * 4
    flops = 2.0 * N * K * d
    return {"search_pipeline_ms": t_search, "search_kernel_ms": tc_ms, "search_tflops": flops / tc_ms / 1e9,
            "search_frac_of_measured_peak": flops / tc_ms / 1e9 / PEAK["bf16_tflops"],
            "search_lookups_per_s": N / t_search * 1e3, "reranked_rows": st["reranked_rows"],
            "rescanned_rows": st["rescanned_rows"],
            "gather_ms": t_gather, "gather_GBs": gather_bytes / t_gather / 1e6,
            "gather_frac_hbm": gather_bytes / t_gather / 1e6 / PEAK["hbm_gbs"],
            "fused_gather_ema_ms": t_fused,
            "fused_GBs": None if t_fused is None else (N * (esz * d + 4 * d + 8) + K * (d + 1) * 4) / t_fused / 1e6,
            "fused_frac_hbm": None if t_fused is None else (N * (esz * d + 4 * d + 8) + K * (d + 1) * 4) / t_fused / 1e6 / PEAK["hbm_gbs"],
            "fused_vs_separate_ops_frac_hbm": None if t_fused is None else (gather_bytes + ema_bytes) / t_fused / 1e6 / PEAK["hbm_gbs"],
            "ema_reduce_ms": t_ema, "ema_GBs": ema_bytes / t_ema / 1e6, "ema_frac_hbm": ema_bytes / t_ema / 1e6 / PEAK["hbm_gbs"]}


def run(name):
    out = {"config": name}
    if name == "C1":
        vq = VectorQuantize(dim=256, codebook_params=CodebookParams(dim=256, codebook_size=512)).to(dev).train()
        x = torch.randn(1, 1024, 256, generator=g, device=dev)
        with torch.no_grad():
            ms = timed(lambda: vq(x), iters=50, warm=5)
        out.update(step_ms=ms, lookups_per_s=1024 / ms * 1e3, note="default ctor: every code expires on step 1 (host syncs + randperm)")
    elif name == "C2":
        vq = VectorQuantize(dim=256, codebook_params=CodebookParams(dim=256, codebook_size=8192)).to(dev).train()
        seed(vq._codebook)
        x = torch.randn(1024, 1024, 256, generator=g, device=dev).bfloat16()
        with torch.no_grad():
            ms = timed(lambda: vq(x))
        out.update(step_ms=ms, lookups_per_s=x.shape[0] * x.shape[1] / ms * 1e3)
        out.update(op_level(x.reshape(1, -1, 256), vq._codebook.embeddings.clone(), False, 8192))
    elif name == "C3":
        vq = VectorQuantize(dim=512, codebook_params=CodebookParams(dim=512, codebook_size=16384, use_cosine_sim=True,
                            transform_input="l2norm", weights_regularization="l2norm")).to(dev).train()
        seed(vq._codebook, l2=True)
        x = torch.randn(512, 1024, 512, generator=g, device=dev)
        with torch.no_grad():
            ms = timed(lambda: vq(x), iters=5, warm=3)
        out.update(step_ms=ms, lookups_per_s=x.shape[0] * x.shape[1] / ms * 1e3)
        xn = ops.l2norm_rows(x.reshape(1, -1, 512))
        out.update(op_level(xn, vq._codebook.embeddings.clone(), True, 16384))
    elif name == "C4":
        rvq = ResidualVQ(dim=512, num_quantizers=8, codebook_params=CodebookParams(dim=512, codebook_size=1024)).to(dev).train()
        for i, l in enumerate(rvq.layers):
            seed(l._codebook, scale=0.5 / 1.4 ** i)
        x = torch.randn(64, 4096, 512, generator=g, device=dev)
        with torch.no_grad():
            ms = timed(lambda: rvq(x), iters=5, warm=3)
            rvq.use_fused_levels = False
            ms_generic = timed(lambda: rvq(x), iters=3, warm=2)
        out.update(step_ms=ms, step_ms_generic_loop=ms_generic, lookups_per_s=8 * 64 * 4096 / ms * 1e3,
                   note="lookups = tokens x 8 levels")
        out.update(op_level(x.reshape(1, -1, 512), rvq.layers[0]._codebook.embeddings.clone(), False, 1024))
    elif name in ("C5_shard", "C5_full"):
        K = 8192 if name == "C5_shard" else 65536
        x = torch.randn(1, 1 << 22, 64, generator=g, device=dev)
        c = torch.randn(1, K, 64, generator=g, device=dev) * 0.5
        out.update(op_level(x, c, False, K))
        out["note"] = "K/8 = 8192 codes per GPU (8-GPU sharded layout)" if name == "C5_shard" else "whole 65536-code codebook on one GPU"
    print(json.dumps(out), flush=True)


for n in (sys.argv[1:] or ["C1", "C2", "C3", "C4", "C5_shard", "C5_full"]):
    run(n)
    torch.cuda.empty_cache()
