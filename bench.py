#!/usr/bin/env python
"""bench.py -- VQ lookups/s on BASELINE.json's headline workload.

A "step" is one pass of the codebook hot path over one batch of synthetic latents: the training-mode
`VectorQuantize.forward` = nearest-code search + gather/straight-through/commitment loss + EMA statistics,
codebook refresh and dead-code check.  Workload (configs[1] of BASELINE.json): 1,048,576 latents x d=256,
codebook K=8192, bf16 latents, Euclidean, EMA decay 0.8; config 2 is "search+EMA" (dead-code expiry belongs to
config 3), so threshold_ema_dead_code=0 here -- stated in the JSON line -- in both arms.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

N>1 is launched by the driver through torch.distributed.run (one rank per GPU, NCCL): data parallel, every rank
quantises its own 1M latents ("weak" scaling) and the EMA statistics are summed with one packed all_reduce.
`--impl reference` times the CPU oracle port of the reference path (the reference is pure Python + torch and
cannot travel; oracle/ restates it op for op) on a bounded sample of the same workload, on rank 0 only.
Prints ONE JSON line.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "vector-quantization-by-ml_b200"))

import torch  # noqa: E402

N_ROWS, K_CODES, DIM = 1 << 20, 8192, 256
SHAPE = (1024, 1024, DIM)               # (batch, tokens, d) -> N = 2^20 latents
CPU_SAMPLE_ROWS = 16384                 # bounded sample for the CPU legs (N x K fp32 must fit the host)
WORKLOAD = "C2: EuclideanCodebook search+EMA, N=1048576 latents x d=256, K=8192, bf16 latents"


def ncu_traffic_bytes():
    """dram__bytes_read.sum + dram__bytes_write.sum of one search-kernel launch, from the committed ncu --set full
    capture of this same command (profiles/rNN_ncu_summary.json); None if there is no capture."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_summary.json")))
    if not files:
        return None
    try:
        for k in json.load(open(files[-1]))["ncu_full"]:
            if "search_aug_kernel" in k["kernel"] or "search_tc_kernel" in k["kernel"]:
                return (k["dram_rd_MB"] + k["dram_wr_MB"]) * 1e6
    except Exception:
        return None
    return None


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"tflops": float(p["bf16_tflops"]), "tflops_sustained": float(p.get("bf16_tflops_sustained", p["bf16_tflops"])),
                "hbm_gbs": float(p["hbm_gbs"]), "source": "measured"}
    return {"tflops": 1590.0, "tflops_sustained": 1400.0, "hbm_gbs": 6650.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (profiling recipe's clocks line)."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        import datetime
        for line in self.proc.stdout:
            line = line.strip()
            t = time.time()
            try:    # nvidia-smi's own timestamp (local time, ms): immune to pipe buffering
                t = datetime.datetime.strptime(line.split(",")[0].strip(), "%Y/%m/%d %H:%M:%S.%f").timestamp()
            except Exception:
                pass
            self.rows.append((t, line))

    def stop(self, t_begin=None, t_end=None):
        """Summary of the samples taken in [t_begin, t_end] (host clock around the timed region, which ends with a
        device sync).  nvidia-smi needs ~0.2 s to deliver its first line, so the sampler is started before the
        warm-up; if the window is shorter than the sampling period the samples nearest to it (all under the same
        load: the warm-up runs the same step) are used."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        rows = self.rows
        if t_begin is not None:
            inside = [r for (t, r) in rows if t_begin <= t <= t_end + 0.03]
            if len(inside) < 3:
                mid = 0.5 * (t_begin + t_end)
                inside = [r for (_, r) in sorted(rows, key=lambda tr: abs(tr[0] - mid))[:5]]
        else:
            inside = [r for (_, r) in rows]
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in inside:
            f = [c.strip() for c in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_port_step(rows, threads):
    """One training forward of the reference path (oracle port) on `rows` latents of the workload, on the CPU."""
    from oracle import vq_oracle as O
    torch.set_num_threads(threads)
    g = torch.Generator().manual_seed(0)
    c = torch.randn(1, K_CODES, DIM, generator=g) * 0.5
    st = O.CodebookState(c.clone(), c.clone(), torch.ones(1, K_CODES))
    x = torch.randn(1, rows, DIM, generator=g).bfloat16()
    opts = O.VQOpts(codebook=O.CodebookOpts(threshold_ema_dead_code=0))

    def step():
        O.vq_forward(st, x, opts, training=True)
    return step


def time_cpu_port(budget_s, threads, rows=CPU_SAMPLE_ROWS, warmup=1, max_steps=None):
    step = cpu_port_step(rows, threads)
    for _ in range(warmup):
        step()
    times = []
    t_end = time.time() + budget_s
    while (time.time() < t_end and (max_steps is None or len(times) < max_steps)) or not times:
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
    return rows / (sum(times) / len(times)), len(times)


def run_reference(args, rank, world):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    rows = CPU_SAMPLE_ROWS
    step = cpu_port_step(rows, threads)
    for _ in range(max(args.warmup, 1)):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    val = rows / dt
    sample = (f"{rows} of {N_ROWS} latents per step (the reference materialises N x K fp32/int64: the full batch needs "
              f">=160 GiB), K={K_CODES}, d={DIM}, all host threads")
    out = {"impl": "reference", "metric": "vq_lookups_per_sec", "value": val, "unit": "lookups/s", "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": WORKLOAD, "sample": sample},
           "cpu_baseline": {"value": val, "unit": "lookups/s", "cores": threads, "kind": "port", "sample": sample},
           "e2e": {"value": val, "unit": "lookups/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline leg")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch.distributed as dist
    from vqb200 import CodebookParams, VectorQuantize, ops

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU implementation")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    # ---- module + synthetic data (seeds: module 0 on every rank, latents 1234 + rank) ----
    torch.manual_seed(0)
    vq = VectorQuantize(dim=DIM, codebook_params=CodebookParams(dim=DIM, codebook_size=K_CODES,
                                                                threshold_ema_dead_code=0),
                        sync_codebook=world > 1).to(dev)
    g = torch.Generator().manual_seed(0)
    c = (torch.randn(1, K_CODES, DIM, generator=g) * 0.5).to(dev)      # trained-like codebook
    cb = vq._codebook
    cb.embeddings.copy_(c); cb.embed_avg.copy_(c); cb.cluster_size.fill_(1.0)
    cb.invalidate_cache()
    vq.train()
    gx = torch.Generator(device=dev).manual_seed(1234 + rank)
    n_bufs = 2                                                          # 512 MiB each, > 126 MB L2: never L2-hot
    xs = [torch.randn(SHAPE, generator=gx, device=dev, dtype=torch.float32).bfloat16() for _ in range(n_bufs)]
    x_host = torch.empty(SHAPE, dtype=torch.bfloat16).pin_memory()
    x_host.copy_(xs[0])
    idx_host = torch.empty(SHAPE[:2], dtype=torch.int64).pin_memory()
    loss_host = torch.empty(1, dtype=torch.float32).pin_memory()

    def step(i):
        with torch.no_grad():
            return vq(xs[i % n_bufs])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)                   # nvidia-smi is streaming by the time the warm-up starts
    t_warm = time.time()                  # the warm-up runs the same step: its samples are under the same load
    for i in range(max(args.warmup, 3)):
        step(i)
    barrier()

    # ---- timed region: device-resident inputs ----
    ops.TIME_SEARCH_KERNEL = True
    ops.search_kernel_times_ms()
    launches0 = ops.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_begin = time.time()
    e0.record()
    for i in range(args.steps):
        step(i)
    e1.record()
    barrier()
    t_end = time.time()
    ms = e0.elapsed_time(e1)
    launches = ops.launch_count() - launches0
    tc_ms = ops.search_kernel_times_ms()
    ops.TIME_SEARCH_KERNEL = False
    clocks = sampler.stop(t_warm, t_end) if rank == 0 else None
    stats = ops.search_stats(cb.last_search_ws)

    # ---- e2e: host buffers in, indices + loss out, copies inside the timed region ----
    # Every step's latents come from pinned host memory and its indices + loss go back to pinned host memory.  The
    # copies run on a side stream, double buffered, so the H2D of step i+1 overlaps the kernels of step i (the way a
    # data loader feeds a training step); nothing is cached between steps.
    copy_stream = torch.cuda.Stream(device=dev)
    main = torch.cuda.current_stream(dev)
    dbuf = [torch.empty(SHAPE, dtype=torch.bfloat16, device=dev) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]

    def prefetch(i):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[i % 2])
            dbuf[i % 2].copy_(x_host, non_blocking=True)
            ready[i % 2].record(copy_stream)

    def e2e_loop(n):
        for b in range(2):
            consumed[b].record(main)
        prefetch(0)
        for i in range(n):
            if i + 1 < n:
                prefetch(i + 1)
            main.wait_event(ready[i % 2])
            with torch.no_grad():
                _, ind, loss = vq(dbuf[i % 2])
            consumed[i % 2].record(main)
            idx_host.copy_(ind, non_blocking=True)
            loss_host.copy_(loss, non_blocking=True)

    e2e_steps = max(3, min(args.steps, 10))
    e2e_loop(2)
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    e2e_loop(e2e_steps)
    f1.record()
    barrier()
    e2e_ms = f0.elapsed_time(f1)

    t = torch.tensor([ms, e2e_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_ms = float(t[0]), float(t[1])

    if rank == 0:
        pk = peaks()
        total_rows = N_ROWS * world
        value = total_rows * args.steps / (ms / 1e3)
        e2e_val = total_rows * e2e_steps / (e2e_ms / 1e3)
        flops = 2.0 * N_ROWS * K_CODES * DIM                     # algorithmic flops per launch of the search kernel
        tc_avg = sum(tc_ms) / len(tc_ms) if tc_ms else float("nan")
        achieved = flops / (tc_avg / 1e3) / 1e12
        roofline = {"kernel": "search_aug_kernel (tcgen05 GEMM, bias as an extra MMA k-step, fused packed top-k epilogue)", "bound": "tensor",
                    "achieved": achieved, "peak": pk["tflops"], "unit": "TFLOP/s", "frac": achieved / pk["tflops"],
                    "peak_source": pk["source"] + " bf16 burst (cuBLAS 8192^3); sustained %.1f" % pk["tflops_sustained"],
                    "frac_of_sustained": achieved / pk["tflops_sustained"],
                    "kernel_ms": tc_avg, "kernel_share_of_step": tc_avg / (ms / args.steps),
                    "algorithmic_flops_per_launch": flops, "traffic": ncu_traffic_bytes(),
                    "traffic_unit": "bytes of DRAM read+write per launch (ncu --set full, profiles/)"}
        threads = os.cpu_count() or 1
        cpu_val, cpu_n = time_cpu_port(args.cpu_seconds, threads)
        sample = (f"{CPU_SAMPLE_ROWS} of {N_ROWS} latents per step x {cpu_n} steps (the reference materialises N x K: "
                  f"the full batch needs >=160 GiB), K={K_CODES}, d={DIM}")
        out = {"metric": "vq_lookups_per_sec", "value": value, "unit": "lookups/s", "n_gpus": world,
               "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
               "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16",
               "data": "synthetic",
               "config": {"workload": WORKLOAD, "step": "VectorQuantize training forward: search + gather/ST/commit loss "
                          "+ EMA reduce/refresh (config 2 = search+EMA: threshold_ema_dead_code=0; expiry is measured "
                          "with config 3 in tools/bench_configs.py)", "rows_per_gpu": N_ROWS, "codebook_size": K_CODES,
                          "dim": DIM, "latent_dtype": "bf16", "parallelism": f"dp{world}",
                          "l2": "inputs (512 MiB per batch, 2 rotating buffers) exceed the 126 MB L2; no flush",
                          "search": stats},
               "roofline": roofline,
               "cpu_baseline": {"value": cpu_val, "unit": "lookups/s", "cores": threads, "kind": "port", "sample": sample},
               "e2e": {"value": e2e_val, "unit": "lookups/s", "steps": e2e_steps,
                       "h2d_bytes_per_step": x_host.numel() * 2 * world,
                       "d2h_bytes_per_step": (idx_host.numel() * 8 + 4) * world,
                       "note": "pinned host latents -> device (side stream, double buffered), forward, indices+loss -> "
                               "pinned host; quantized fp32 stays on the device for the consumer"},
               "gpu_launches": int(launches), "clocks": clocks}
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
