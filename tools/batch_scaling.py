"""Per-clip cost of a diffusion step vs batch size on ONE GPU (K=512 codebook, 19 layers): is the sampler's time linear in B?
Used to read the strong-scaling numbers of bench.py's configs[4] extra (512 / N clips per rank).

    python tools/batch_scaling.py --batches 16,64,128,256,512
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import _pkg  # noqa: E402

_pkg.load()
from bench import ClockSampler, synthetic_cond  # noqa: E402
from diffsound_b200.utils import builders  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batches", default="16,64,128,256,512")
ap.add_argument("--K", type=int, default=512)
ap.add_argument("--steps", type=int, default=6)
ap.add_argument("--precision", default="f16x3")
a = ap.parse_args()
dev = torch.device("cuda:0")
dalle = builders.build_dalle(K=a.K, NL=19, precision=a.precision, seed=0)
tr = dalle.transformer
out = []
for B in [int(b) for b in a.batches.split(",")]:
    cond = synthetic_cond(B, 7).to(dev)
    skip = max(0, 100 // a.steps - 1)
    kw = dict(condition_token=None, condition_mask=None, condition_embed=cond, content_token=None, filter_ratio=0, temperature=1.0, return_att_weight=False,
              return_logits=False, print_log=False, sample_type="top0.85r", batch_size=B)
    tr.truncation, tr.resample_rate = "top0.85r", 0.0
    tr.sample_fast(skip_step=skip, **kw)  # warm-up at this batch: graph capture, workspaces
    torch.cuda.synchronize()
    clk = ClockSampler(0)
    clk.start()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    n = 3
    for _ in range(n):
        tr.sample_fast(skip_step=skip, **kw)
    e.record()
    torch.cuda.synchronize()
    st = list(range(99, -1, -1 - skip))
    calls = len(st) + (1 if st[-1] != 0 else 0)  # sample_fast appends t = 0
    ms = s.elapsed_time(e) / n
    c = clk.stop()
    out.append({"B": B, "ms_per_pass": round(ms, 2), "denoiser_calls": calls, "ms_per_call": round(ms / calls, 3), "us_per_clip_call": round(ms / calls / B * 1e3, 2),
                "sm_mhz": c.get("sm_mhz"), "reasons": c.get("reasons")})
    print(out[-1], file=sys.stderr)
print(json.dumps(out))
