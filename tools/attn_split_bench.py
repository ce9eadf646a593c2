"""Device-time micro-benchmark of the split-fp16 attention kernel at the bench shapes (B=16, 16 heads, L=265 / Lc=77), next to the
single-pass tcgen05 kernel and the split GEMMs of one layer."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _pkg  # noqa: E402

_pkg.load()
from diffsound_b200 import ops  # noqa: E402
from tools.attn_microbench import timeit  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
H, L, Lc, D, NL = 16, 265, 77, 1024, 19
qkv = (torch.randn(B * L, 6 * D, device="cuda") * 0.5).half()
kv = (torch.randn(B * Lc, 2 * NL * 2 * D, device="cuda") * 0.5).half()
q2 = (torch.randn(B * L, 2 * D, device="cuda") * 0.5).half()
att = torch.empty(B * L, 2 * D, device="cuda", dtype=torch.float16)
us = timeit(lambda: ops.attention_tc_split(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:3 * D], att[:, :D], q_lo=3 * D, k_lo=3 * D, v_lo=3 * D, o_lo=D, B=B, H=H, Lq=L, Lk=L,
                                           scale=0.125))
print(f"B={B} self-attention  split: {us:7.1f} us  ({3 * 4 * B * H * L * L * 64 / us / 1e6:7.1f} executed TFLOP/s)")
us = timeit(lambda: ops.attention_tc_split(q2[:, :D], kv[:, :D], kv[:, D:2 * D], att[:, :D], q_lo=D, k_lo=NL * 2 * D, v_lo=NL * 2 * D, o_lo=D, B=B, H=H, Lq=L, Lk=Lc,
                                           scale=0.125))
print(f"B={B} cross-attention split: {us:7.1f} us  ({3 * 4 * B * H * L * Lc * 64 / us / 1e6:7.1f} executed TFLOP/s)")
us = timeit(lambda: ops.attention_tc(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:3 * D], att[:, :D], B=B, H=H, Lq=L, Lk=L, scale=0.125))
print(f"B={B} self-attention  f16 tc2: {us:7.1f} us")
us = timeit(lambda: ops.attention_tc(q2[:, :D], kv[:, :D], kv[:, D:2 * D], att[:, :D], B=B, H=H, Lq=L, Lk=Lc, scale=0.125))
print(f"B={B} cross-attention f16 tc2: {us:7.1f} us")
x = torch.randn(B, L, D, device="cuda")
tab = torch.randn(100, 2 * D, device="cuda")
t = torch.full((B,), 5, device="cuda", dtype=torch.long)
h = torch.empty(B, L, 2 * D, device="cuda", dtype=torch.float16)
print(f"ada_layernorm split out (L2-warm x): {timeit(lambda: ops.ada_layernorm(x, tab, t, out=h, split=True)):7.1f} us")
