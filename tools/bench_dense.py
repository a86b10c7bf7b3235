"""Times the fp32 dense-similarity passes (csrc/dense.cu) with CUDA events: one JSON line per pass.
    python tools/bench_dense.py [N K d [passes [iters]]]      # passes: comma-separated subset, e.g. rowstats,backward_ce
Algorithmic work: 2 N K d flops per score pass; the backward pass does two contractions (scores + combination)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vector-quantization-by-ml_b200"))
from vqb200 import ops  # noqa: E402


def timed(fn, iters):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def main():
    N, K, d = (int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (65536, 8192, 256)
    only = sys.argv[4].split(",") if len(sys.argv) >= 5 else None
    iters = int(sys.argv[5]) if len(sys.argv) >= 6 else 5
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(0)
    x = torch.randn(1, N, d, generator=g).to(dev)
    c = (torch.randn(1, K, d, generator=g) * 0.5).to(dev)
    tgt = torch.randint(0, K, (1, N), generator=g).to(dev)
    xn2, cn2 = ops.dense_row_norms(x), ops.dense_row_norms(c)
    lse, _ = ops.dense_rowstats(x, xn2, c, cn2, False, 1.0, tgt)
    coef = torch.full((1, N), 1.0 / N, device=dev)
    n_pos = 1024 if N % 1024 == 0 else 1
    table = torch.randn(n_pos, K, generator=g).to(dev)
    flops = 2.0 * N * K * d
    runs = {
        "rowstats": (lambda: ops.dense_rowstats(x, xn2, c, cn2, False, 1.0, tgt), 1),
        "avgprob": (lambda: ops.dense_avgprob(x, xn2, c, cn2, False, 1.0, lse, n_pos), 1),
        "rowdot": (lambda: ops.dense_rowdot(x, xn2, c, cn2, False, 1.0, lse, table, n_pos), 1),
        "backward_ce": (lambda: ops.dense_backward(x, xn2, c, cn2, c, False, 1.0, lse, coef, target=tgt), 2),
        "backward_codes_ce": (lambda: ops.dense_backward_codes(x, xn2, c, cn2, False, 1.0, lse, coef, target=tgt), 2),
    }
    for name, (fn, passes) in runs.items():
        if only is not None and name not in only:
            continue
        ms = timed(fn, iters)
        print(json.dumps({"pass": name, "bk": int(os.environ.get("VQB_DENSE_BK", "0")) or "default", "N": N, "K": K, "d": d,
                          "ms": round(ms, 3),
                          "fp32_tflops": round(passes * flops / ms / 1e9, 2)}))


if __name__ == "__main__":
    main()
