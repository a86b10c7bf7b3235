"""Bring-up: per-tile clock stamps of cluster 0 of the d = 64 search kernel (library built with `make -C vector-quantization-by-ml_b200/csrc trace`
-> tools/micro/libvqb200_trace.so).  Prints the intervals of the accumulator hand-off chain in SM cycles, per epilogue warp."""
import os, sys, ctypes as C, statistics as st
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "vector-quantization-by-ml_b200"))
import torch
from vqb200 import ops, _lib
_lib.LIB_PATH = os.path.join(ROOT, "tools", "micro", os.environ.get("VQB_TRACE_LIB", "libvqb200_trace.so"))
dev = torch.device("cuda:0")
N, K, d = (int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (1 << 22, 8192, 64)
g = torch.Generator(device=dev).manual_seed(0)
x = torch.randn(1, N, d, generator=g, device=dev)
c = torch.randn(1, K, d, generator=g, device=dev) * 0.5
cache = ops.prepare_codebook(c, False)
ops.TIME_SEARCH_KERNEL = True
for _ in range(3):
    ops.search(x, c, cache, False)
torch.cuda.synchronize(); ops.search_kernel_times_ms()
for _ in range(3):
    ops.search(x, c, cache, False)
torch.cuda.synchronize()
t = ops.search_kernel_times_ms()
print(f"kernel {sum(t)/len(t):.3f} ms ({_lib.LIB_PATH.split('/')[-1]})")
T = 256
bm = (C.c_longlong * (3 * T))(); be = (C.c_longlong * (96 * T))()
raw = _lib.lib()
raw.vqb_debug_trace.argtypes = [C.c_void_p, C.c_void_p]
assert raw.vqb_debug_trace(C.cast(bm, C.c_void_p), C.cast(be, C.c_void_p)) == 0
M = [[bm[r * T + i] for i in range(T)] for r in range(3)]
E = [[[be[(r * 32 + w) * T + i] for i in range(T)] for w in range(32)] for r in range(3)]
rng = range(64, 250)
f = lambda xs: f"{st.mean(xs):7.0f}"
print("MMA warp: empty_seen->issued", f([M[1][i] - M[0][i] for i in rng]), " issued->full seen by itself (_lat build)",
      f([M[2][i] - M[1][i] for i in rng]), " period", f([M[0][i + 1] - M[0][i] for i in rng]))
print("warp  issued->full_seen  full_seen->arrive  arrive->ranked  ranked->next full_seen  arrive(t)->mma empty_seen(t+2)   (leader CTA: same clock as the MMA warp)")
for w in range(16):
    print(f"{w:4d} {f([E[0][w][i] - M[1][i] for i in rng])}          {f([E[1][w][i] - E[0][w][i] for i in rng])}          "
          f"{f([E[2][w][i] - E[1][w][i] for i in rng])}         {f([E[0][w][i + 1] - E[2][w][i] for i in rng])}              "
          f"{f([M[0][i + 2] - E[1][w][i] for i in rng])}")
print("last leader arrive(t) -> mma empty_seen(t+2):", f([M[0][i + 2] - max(E[1][w][i] for w in range(16)) for i in rng]),
      "  first:", f([M[0][i + 2] - min(E[1][w][i] for w in range(16)) for i in rng]))
print("first leader full_seen(t) - issued(t):", f([min(E[0][w][i] for w in range(16)) - M[1][i] for i in rng]))
print("peer CTA (own clock): warp  full_seen->arrive  arrive->ranked  ranked->next full_seen   arrive spread (last-first)")
for w in range(16, 32):
    print(f"{w - 16:4d} {f([E[1][w][i] - E[0][w][i] for i in rng])}   {f([E[2][w][i] - E[1][w][i] for i in rng])}   "
          f"{f([E[0][w][i + 1] - E[2][w][i] for i in rng])}")
print("arrive spread leader", f([max(E[1][w][i] for w in range(16)) - min(E[1][w][i] for w in range(16)) for i in rng]),
      " peer", f([max(E[1][w][i] for w in range(16, 32)) - min(E[1][w][i] for w in range(16, 32)) for i in rng]))
print("full_seen spread leader", f([max(E[0][w][i] for w in range(16)) - min(E[0][w][i] for w in range(16)) for i in rng]),
      " peer", f([max(E[0][w][i] for w in range(16, 32)) - min(E[0][w][i] for w in range(16, 32)) for i in rng]))
print("one tile in detail (tile 130), leader warps: full_seen / arrive / ranked relative to issued(130)")
i = 130
for w in range(16):
    print(f"  w{w:2d} {E[0][w][i] - M[1][i]:6d} {E[1][w][i] - M[1][i]:6d} {E[2][w][i] - M[1][i]:6d}")
print(f"  mma: empty_seen(130) {M[0][i] - M[1][i]}, empty_seen(131) {M[0][i+1] - M[1][i]}, empty_seen(132) {M[0][i+2] - M[1][i]}, issued(131) {M[1][i+1] - M[1][i]}")
# per-position-in-row-tile period of the MMA warp (where a row tile spends its time)
NT = (K + 255) // 256
if NT <= 64 and M[0][NT] != 0:      # the MMA warp's stamps exist for one-stage tiles (d_pad = 64) only
    print(f"MMA period by N tile within a row tile (NT = {NT}; mean over the traced row tiles, cycles):")
    per = [[] for _ in range(NT)]
    for i in range(NT, min(T - 1, (T // NT) * NT - 1)):
        per[i % NT].append(M[0][i + 1] - M[0][i])
    print("  " + " ".join(f"{int(st.mean(p)):5d}" for p in per if p))
    tot = sum(st.mean(p) for p in per if p)
    print(f"  row tile total {tot:.0f} cycles; steady-state tile (median position) {sorted(st.mean(p) for p in per if p)[NT // 2]:.0f}")
