"""Randomised differential check of vqb_search: the tensor-core path (with and without the in-kernel latent
conversion, with and without scores, with an index offset) against the exact scan of the same library, over random
shapes / dtypes / metrics / value scales, incl. duplicated codes, zero rows and outlier rows.  Indices and scores must be
bit-identical.  usage: python tools/fuzz_search.py [cases] [seed]"""
import os, sys, random
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "vector-quantization-by-ml_b200"))
import torch
from vqb200 import ops

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 150
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rnd = random.Random(seed)
dev = torch.device("cuda:0")
bad = 0
tc_used = 0
n_fused = 0
for it in range(cases):
    H = rnd.choice([1, 1, 1, 2, 3])
    d = rnd.choice([8, 16, 24, 32, 40, 64, 72, 96, 128, 136, 192, 256, 264, 320, 384, 512])
    K = rnd.choice([1, 2, 7, 64, 255, 256, 257, 1000, 1024, 4096, 5000, 8192, 12288])
    N = rnd.choice([1, 5, 127, 128, 129, 255, 256, 257, 1000, 4096, 8191, 8192, 20000, 40000])
    if H * N * K * d > 3e11:
        N = 4096
    dt = rnd.choice([torch.float32, torch.bfloat16, torch.float16])
    cos = rnd.random() < 0.3
    if rnd.random() < 0.15:      # a shape that qualifies for the in-kernel latent conversion
        H, d, K = rnd.choice([1, 2]), rnd.choice([192, 248, 256]), rnd.choice([8192, 8200, 12288])
        N, dt = rnd.choice([8192, 9000, 20000, 33000]), rnd.choice([torch.bfloat16, torch.float16])
    scale = 10.0 ** rnd.uniform(-3, 2)
    g = torch.Generator(device=dev).manual_seed(seed * 100003 + it)
    x = torch.randn(H, N, d, generator=g, device=dev) * scale
    c = torch.randn(H, K, d, generator=g, device=dev) * scale * rnd.choice([0.1, 0.5, 1.0, 3.0])
    if K > 4 and rnd.random() < 0.3:
        c[:, K // 2] = c[:, 1]                                   # duplicated code: the lower index must win
    if N > 64 and rnd.random() < 0.5:
        x[:, 3:9] = 0.0
        x[:, 17] *= 50.0
        x[:, N // 2] = c[:, 0] if d == c.shape[-1] else x[:, N // 2]   # a latent exactly on a code
    if cos:
        c = torch.nn.functional.normalize(c, dim=-1)
    x = x.to(dt)
    want = rnd.random() < 0.5
    off = rnd.choice([0, 0, 12345])
    fused = rnd.random() < 0.5
    n_fused += int(fused and dt != torch.float32 and d <= 256 and d % 8 == 0 and N >= 8192 and ((K + 255) // 256) * ((d + 63) // 64) >= 96)
    try:
        cache = ops.prepare_codebook(c, cos)
        idx, sc, ws = ops.search(x, c, cache, cos, want_score=want, idx_offset=off, fused_prep=fused)
        st = ops.search_stats(ws)
        ex, es, _ = ops.search(x, c, None, cos, force_exact=True, want_score=want, idx_offset=off)
        ok = torch.equal(idx, ex) and (not want or torch.equal(sc, es))
    except Exception as e:                                       # noqa: BLE001
        ok = False
        st = {"error": repr(e)[:200]}
    tc_used += int(st.get("tensor_core_pass", 0) == 1)
    if not ok:
        bad += 1
        nd = int((idx != ex).sum()) if "error" not in st else -1
        print(f"MISMATCH case {it}: H={H} N={N} K={K} d={d} {dt} cos={cos} scale={scale:.3g} score={want} off={off} "
              f"fused={fused} rows_differ={nd} {st}", flush=True)
print(f"fuzz_search: {cases} cases, seed {seed}, {tc_used} through the tensor-core path, {n_fused} with the in-kernel conversion, {bad} mismatches")
sys.exit(1 if bad else 0)
