"""Where does a step go?  Wraps every top-level call into `vqb200.ops` / `vqb200.distributed` (and the collectives of
torch.distributed) in CUDA events on the current stream and prints, per named call, the device time per step and what
is left over (gaps between launches, host work).  There is no nsys in this image; this is the timeline substitute.

    python tools/phase_timing.py --config C5 [--steps 5]                      # one GPU (un-sharded C5)
    python -m torch.distributed.run --nproc-per-node W --master-addr 127.0.0.1 tools/phase_timing.py --config C5

Configs: C2 | C3 | C4 | C5 (the shapes of BASELINE.json; C5 is sharded over the ranks when W > 1).
Writes one JSON line per rank-0 run (append to --out when given)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "vector-quantization-by-ml_b200"))
import torch
import torch.distributed as dist

from vqb200 import CodebookParams, ResidualVQ, VectorQuantize, distributed as D, ops  # noqa: E402


class Phases:
    def __init__(self):
        self.depth = 0
        self.records = []      # (name, ev0, ev1)
        self.on = False

    def wrap(self, mod, name, label):
        fn = getattr(mod, name)
        me = self

        def timed(*a, **k):
            if not me.on or me.depth > 0:
                return fn(*a, **k)
            me.depth += 1
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            try:
                return fn(*a, **k)
            finally:
                e1.record()
                me.records.append((label, e0, e1))
                me.depth -= 1
        setattr(mod, name, timed)

    def install(self):
        import types
        for name, fn in list(vars(ops).items()):
            if isinstance(fn, types.FunctionType) and not name.startswith("_") and name not in (
                    "launch_count", "search_kernel_times_ms", "release_workspaces", "workspace", "search_stats"):
                self.wrap(ops, name, "ops." + name)
        for name, fn in list(vars(D).items()):
            if isinstance(fn, types.FunctionType) and not name.startswith("_") and name not in (
                    "is_distributed", "world_size", "rank"):
                self.wrap(D, name, "dist." + name)
        for name in ("all_reduce", "all_gather_into_tensor", "broadcast"):
            self.wrap(dist, name, "nccl." + name)

    def summary(self, steps):
        torch.cuda.synchronize()
        tot = {}
        cnt = {}
        for label, e0, e1 in self.records:
            tot[label] = tot.get(label, 0.0) + e0.elapsed_time(e1)
            cnt[label] = cnt.get(label, 0) + 1
        return {k: {"ms_per_step": v / steps, "calls_per_step": cnt[k] / steps} for k, v in
                sorted(tot.items(), key=lambda kv: -kv[1])}


def build(config, dev, world, rank):
    g = torch.Generator().manual_seed(0)
    if config == "C2":
        K, d, shape, dt = 8192, 256, (1024, 1024, 256), torch.bfloat16
        m = VectorQuantize(dim=d, codebook_params=CodebookParams(dim=d, codebook_size=K, threshold_ema_dead_code=0),
                           sync_codebook=world > 1).to(dev).train()
        c = (torch.randn(1, K, d, generator=g) * 0.5).to(dev)
        cb = m._codebook
        cb.embeddings.copy_(c); cb.embed_avg.copy_(c); cb.cluster_size.fill_(1.0); cb.invalidate_cache()
    elif config == "C3":
        K, d, shape, dt = 16384, 512, (512, 1024, 512), torch.float32
        m = VectorQuantize(dim=d, codebook_params=CodebookParams(dim=d, codebook_size=K, threshold_ema_dead_code=2,
                                                                 use_cosine_sim=True, transform_input="l2norm",
                                                                 weights_regularization="l2norm"),
                           sync_codebook=world > 1).to(dev).train()
        c = torch.nn.functional.normalize(torch.randn(1, K, d, generator=g), dim=-1).to(dev)
        cb = m._codebook
        cb.embeddings.copy_(c); cb.embed_avg.copy_(c); cb.cluster_size.fill_(1.0); cb.invalidate_cache()
    elif config == "C4":
        K, d, shape, dt = 1024, 512, (64, 4096, 512), torch.float32
        m = ResidualVQ(dim=d, num_quantizers=8, codebook_params=CodebookParams(dim=d, codebook_size=K),
                       sync_codebook=world > 1).to(dev).train()
        for li, layer in enumerate(m.layers):
            c = (torch.randn(1, K, d, generator=g) * (0.5 / 1.4 ** li)).to(dev)
            cb = layer._codebook
            cb.embeddings.copy_(c); cb.embed_avg.copy_(c); cb.cluster_size.fill_(8.0); cb.invalidate_cache()
    elif config == "C5":
        K, d, shape, dt = 65536, 64, (4096, 1024, 64), torch.float32
        m = VectorQuantize(dim=d, codebook_params=CodebookParams(dim=d, codebook_size=K, threshold_ema_dead_code=0),
                           sync_codebook=world > 1).to(dev).train()
        m._codebook.load_full_codebook(torch.randn(1, K, d, generator=g) * 0.5)
        m._codebook.sharded_input = "replicated"
    else:
        raise SystemExit("unknown config " + config)
    gx = torch.Generator(device=dev).manual_seed(4321 if config == "C5" else 1234 + rank)
    xs = [torch.randn(shape, generator=gx, device=dev).to(dt) for _ in range(2)]
    return m, xs


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="C5")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ph = Phases()
    ph.install()
    m, xs = build(args.config, dev, world, rank)
    with torch.no_grad():
        for i in range(args.warmup):
            m(xs[i % 2])
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ph.on = True
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.steps):
            m(xs[i % 2])
        e1.record()
        torch.cuda.synchronize()
        ph.on = False
    step_ms = e0.elapsed_time(e1) / args.steps
    phases = ph.summary(args.steps)
    named = sum(v["ms_per_step"] for v in phases.values())
    line = {"config": args.config, "world": world, "rank": rank, "ms_per_step": step_ms,
            "named_ms_per_step": named, "unattributed_ms_per_step": step_ms - named, "phases": phases}
    if rank == 0:
        s = json.dumps(line)
        print(s)
        if args.out:
            os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
            with open(args.out, "a") as f:
                f.write(s + "\n")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
