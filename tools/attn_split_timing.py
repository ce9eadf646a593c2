"""Phase timestamps of CTA 0 of the split-fp16 attention kernel (self-attention shape, B=16): where does a 128-row tile's chain spend its cycles?"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _pkg  # noqa: E402

_pkg.load()
from diffsound_b200 import _lib, ops  # noqa: E402

B, H, L, D = 16, 16, 265, 1024
Lk = int(sys.argv[1]) if len(sys.argv) > 1 else L
qkv = (torch.randn(B * L, 6 * D, device="cuda") * 0.5).half()
kvb = (torch.randn(B * Lk, 6 * D, device="cuda") * 0.5).half()
att = torch.empty(B * L, 2 * D, device="cuda", dtype=torch.float16)
run = lambda: ops.attention_tc_split(qkv[:, :D], kvb[:, D:2 * D], kvb[:, 2 * D:3 * D], att[:, :D], q_lo=3 * D, k_lo=3 * D, v_lo=3 * D, o_lo=D, B=B, H=H, Lq=L, Lk=Lk,
                                     scale=0.125)
run(); torch.cuda.synchronize()
_lib.check(_lib.lib().dsb_attention_split_timing(1, None), "timing on")
run(); torch.cuda.synchronize()
buf = (ctypes.c_longlong * 128)()
_lib.check(_lib.lib().dsb_attention_split_timing(0, ctypes.cast(buf, ctypes.c_void_p)), "timing read")
t = [list(buf[i * 16:(i + 1) * 16]) for i in range(8)]
base = min(v for row in t for v in row if v > 0)
names = {0: "S ready (softmax w0)", 1: "pass1 max done", 2: "after bar", 3: "P written", 4: "O ready", 5: "epilogue done", 8: "ctrl: S issued", 9: "ctrl: S done/Q,K refill",
         10: "ctrl: P ready", 11: "ctrl: PV issued", 12: "ctrl: loop top", 13: "ctrl: Q ready", 14: "ctrl: fenced", 6: "epiW: O ready", 7: "epiW: stored"}
for i, row in enumerate(t[:4]):
    print(f"tile {i}: " + "  ".join(f"{names[k]}={row[k] - base}" for k in sorted(names) if row[k] > 0))
