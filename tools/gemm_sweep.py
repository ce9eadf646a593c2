"""One-call sweep of dsb_gemm_ex configurations for the shapes that matter (round-2 starting point; no numbers are assumed here):

  * inference, M = B*265 for B in {16, 20, 64}: the six Linear shapes of a denoiser layer, tile width 128 / 256, 1-CTA / CTA-pair, fp16 operands;
  * training, M = 20*265: the same shapes in bf16 plus their data-gradient (W as stored, MN-major B) and weight-gradient
    (both operands token-major, MN-major A and B, K = M) forms.

Timing: CUDA-graph replay of `reps` launches rotating over 8 weight buffers, CUDA events (tools/gemm_microbench.py's method).  Prints one
table; `--json out.json` also writes it.  Usage on the GPU box:  python tools/gemm_sweep.py [--batches 16 20 64] [--json gpurun_out/gemm_sweep.json]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _pkg  # noqa: E402

_pkg.load()
from diffsound_b200 import ops  # noqa: E402


def timeit(fn, reps=16):
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


LAYER = [("qkv", 3072, 1024, "f16out"), ("proj", 1024, 1024, "res"), ("mlp1", 4096, 1024, "gelu"), ("mlp2", 1024, 4096, "res"), ("logits", 256, 1024, "f32")]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batches", type=int, nargs="+", default=[16, 20, 64])
    ap.add_argument("--json", default=None)
    a = ap.parse_args()
    rows = []

    def rec(**kw):
        rows.append(kw)
        print(" ".join(f"{k}={v}" if not isinstance(v, float) else f"{k}={v:.1f}" for k, v in kw.items()), flush=True)

    for B in a.batches:
        M = B * 265
        for name, N, K, epi in LAYER:
            for dt, code in ((torch.float16, ops.F16), (torch.bfloat16, ops.BF16)):
                if dt == torch.bfloat16 and B != 20:
                    continue
                x = torch.randn(M, K, device="cuda").to(dt)
                ws = [(torch.randn(N, K, device="cuda") * 0.05).to(dt) for _ in range(8)]
                bias = torch.randn(N, device="cuda")
                o32 = torch.empty(M, N, device="cuda")
                o16 = torch.empty(M, N, device="cuda", dtype=dt)
                it = [0]

                def w():
                    it[0] += 1
                    return ws[it[0] % 8]
                for bn in (128, 256):
                    for pair in (-1, 1):
                        if pair == 1 and bn == 128:
                            continue
                        if epi == "res":
                            fn = lambda: ops.gemm(x, w(), bias, o32, o32, dtype=code, block_n=bn, cta_pair=pair)
                        elif epi == "gelu":
                            fn = lambda: ops.gemm(x, w(), bias, None, o16, dtype=code, gelu=True, block_n=bn, cta_pair=pair)
                        elif epi == "f16out":
                            fn = lambda: ops.gemm(x, w(), bias, None, o16, dtype=code, block_n=bn, cta_pair=pair)
                        else:
                            fn = lambda: ops.gemm(x, w(), bias, None, o32, dtype=code, block_n=bn, cta_pair=pair)
                        us = timeit(fn)
                        rec(kind="fwd", B=B, shape=name, M=M, N=N, K=K, dtype=str(dt).split(".")[-1], block_n=bn, pair=pair, us=us, tflops=2.0 * M * N * K / us / 1e6)
                if dt == torch.bfloat16:  # training-only forms
                    dy = torch.randn(M, N, device="cuda").to(dt)
                    dx = torch.empty(M, K, device="cuda", dtype=dt)
                    dw = torch.empty(N, K, device="cuda")
                    for bn in (128, 256):
                        us = timeit(lambda: ops.gemm(dy, w(), None, None, dx, dtype=code, w_mn=True, block_n=bn))
                        rec(kind="dgrad", B=B, shape=name, M=M, N=K, K=N, dtype="bfloat16", block_n=bn, pair=-1, us=us, tflops=2.0 * M * N * K / us / 1e6)
                        us = timeit(lambda: ops.gemm(dy, x, None, None, dw, dtype=code, a_mn=True, w_mn=True, block_n=bn))
                        rec(kind="wgrad", B=B, shape=name, M=N, N=K, K=M, dtype="bfloat16", block_n=bn, pair=-1, us=us, tflops=2.0 * M * N * K / us / 1e6)
    if a.json:
        with open(a.json, "w") as f:
            json.dump(rows, f)


if __name__ == "__main__":
    main()
