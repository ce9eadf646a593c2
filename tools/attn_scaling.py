import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _pkg; _pkg.load()
from diffsound_b200 import ops
from tools.attn_microbench import timeit
H, D = 16, 1024
for Lq, Lk in ((265, 265), (256, 265), (265, 77), (128, 265)):
    for B in (4, 8, 16, 32, 64):
        qkv = torch.randn(B * Lq, 3 * D, device="cuda").half(); kv = torch.randn(B * Lk, 2 * D, device="cuda").half()
        out = torch.empty(B * Lq, D, device="cuda", dtype=torch.float16)
        us = timeit(lambda: ops.attention_tc(qkv[:, :D], kv[:, :D], kv[:, D:], out, B=B, H=H, Lq=Lq, Lk=Lk, scale=0.125))
        nqt = Lq // 128 if (Lq > 128 and 0 < Lq % 128 <= 16) else (Lq + 127) // 128
        print(f"Lq={Lq} Lk={Lk} B={B:3d}: {us:7.1f} us   tiles/CTA={B * H * nqt / min(148, B * H * nqt):5.2f}")
