"""Run a few eager (no CUDA graph) p_sample steps of the full-size denoiser -- the target command for ncu captures.

    ncu --metrics gpu__time_duration.sum --clock-control none -s <launches of step 0> -c <launches of one step> --csv \
        --log-file gpurun_out/launches.csv python tools/profile_step.py --steps 2
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import synthetic_cond  # noqa: E402
import _pkg  # noqa: E402
_pkg.load()
from diffsound_b200.utils.builders import build_diffusion_transformer  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--precision", default="f16x3")
ap.add_argument("--layers", type=int, default=19)
args = ap.parse_args()

torch.manual_seed(0)
m = build_diffusion_transformer(256, 1024, args.layers, 16, 512, precision=args.precision)
m.truncation = "top0.85r"
m.use_cuda_graph = False
cond = synthetic_cond(args.batch, 1).cuda()
torch.manual_seed(1234)
steps = list(range(99, 99 - args.steps, -1))
tok = m._run_steps(cond, args.batch, steps, steps)
torch.cuda.synchronize()
print("launches per step:", m.transformer.engine.launches_per_forward + 1, "(+ torch.rand, copy_, 2 fill_)", "tokens", tok[0, :6].tolist())
