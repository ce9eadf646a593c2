"""Device-time micro-benchmark of dsb_gemm_ex variants (CUDA-graph replay, CUDA events): separates the fixed per-tile cost
(prologue + epilogue, K=64) from the mainloop slope, and compares epilogue flavours / tile widths."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _pkg

_pkg.load()
from diffsound_b200 import ops


def timeit(fn, reps=20):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(3):
        g.replay()
    e.record(); e.synchronize()
    return s.elapsed_time(e) / (3 * reps) * 1e3  # us


M = 4240
print(f"{'case':58s} {'us':>8s} {'TFLOP/s':>9s}")
for N, K in [(1024, 64), (1024, 256), (1024, 1024), (1024, 4096), (3072, 1024), (4096, 1024), (4096, 64)]:
    a = torch.randn(M, K, device="cuda").half()
    ws = [(torch.randn(N, K, device="cuda") * 0.05).half() for _ in range(8)]  # rotate weights: no L2-warm W
    bias = torch.randn(N, device="cuda")
    res = torch.randn(M, N, device="cuda")
    out32 = torch.empty(M, N, device="cuda")
    out16 = torch.empty(M, N, device="cuda", dtype=torch.float16)
    it = [0]

    def w():
        it[0] += 1
        return ws[it[0] % 8]

    cases = [("1cta f32 out, bias, residual (in place)", lambda: ops.gemm(a, w(), bias, out32, out32, dtype=ops.F16, cta_pair=-1)),
             ("pair f32 out, bias, residual (in place)", lambda: ops.gemm(a, w(), bias, out32, out32, dtype=ops.F16, cta_pair=1)),
             ("1cta f16 out, bias", lambda: ops.gemm(a, w(), bias, None, out16, dtype=ops.F16, cta_pair=-1)),
             ("pair f16 out, bias", lambda: ops.gemm(a, w(), bias, None, out16, dtype=ops.F16, cta_pair=1)),
             ("1cta f16 out, bias, gelu", lambda: ops.gemm(a, w(), bias, None, out16, dtype=ops.F16, gelu=True, cta_pair=-1)),
             ("pair f16 out, bias, gelu", lambda: ops.gemm(a, w(), bias, None, out16, dtype=ops.F16, gelu=True, cta_pair=1))]
    for name, fn in cases:
        us = timeit(fn)
        print(f"N={N:5d} K={K:5d} {name:42s} {us:8.1f} {2.0 * M * N * K / us / 1e6:9.1f}")
# memcpy-style reference: how fast can 17 MB be written / 17+17 read+written
x = torch.empty(M, 1024, device="cuda"); y = torch.empty(M, 1024, device="cuda")
print("torch copy_ 17MB->17MB us:", timeit(lambda: y.copy_(x)))
