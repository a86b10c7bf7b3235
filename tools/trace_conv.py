"""Bring-up: clock stamps of the converter warps and the producer of cluster 0 (`make -C vector-quantization-by-ml_b200/csrc trace` -> tools/micro/libvqb200_trace.so)."""
import os, sys, ctypes as C, statistics as st
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "vector-quantization-by-ml_b200"))
import torch
from vqb200 import ops, _lib
_lib.LIB_PATH = os.path.join(ROOT, "tools", "micro", "libvqb200_trace.so")
dev = torch.device("cuda:0")
N, K, d = 1 << 20, 8192, 256
g = torch.Generator(device=dev).manual_seed(0)
x = torch.randn(1, N, d, generator=g, device=dev).bfloat16()
c = torch.randn(1, K, d, generator=g, device=dev) * 0.5
cache = ops.prepare_codebook(c, False)
ops.TIME_SEARCH_KERNEL = True
for _ in range(3):
    ops.search(x, c, cache, False, fused_prep=True)
torch.cuda.synchronize()
print("kernel ms", ops.search_kernel_times_ms())
T = 256
bm = (C.c_longlong * (3 * T))(); be = (C.c_longlong * (96 * T))()
raw = _lib.lib(); raw.vqb_debug_trace.argtypes = [C.c_void_p, C.c_void_p]
assert raw.vqb_debug_trace(C.cast(bm, C.c_void_p), C.cast(be, C.c_void_p)) == 0
E = [[[be[(r * 32 + w) * T + i] for i in range(T)] for w in range(32)] for r in range(3)]
t0 = E[0][18][0]
print("row tile: conv0 start / got slot / done | conv1 start / got slot / done | producer asks / gets   (cycles since the converter's first stamp)")
for i in range(0, 40):
    print(f"{i:3d}  {E[0][18][i]-t0:9d} {E[1][18][i]-t0:9d} {E[2][18][i]-t0:9d} | {E[0][19][i]-t0:9d} {E[1][19][i]-t0:9d} {E[2][19][i]-t0:9d} | {E[0][20][i]-t0:9d} {E[1][20][i]-t0:9d}")
conv = [E[2][18][i] - E[1][18][i] for i in range(2, 50)]
print("conversion of 64 rows by one warp: mean", st.mean(conv), "cycles; producer wait mean", st.mean([E[1][20][i] - E[0][20][i] for i in range(2, 50)]))
