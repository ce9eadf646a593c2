"""Target of the `ncu --set full` captures (profiles/r2_*): one eager pass of MelGAN, the SpecVQGAN decoder and the first denoiser layers at the
bench shape (B=16), bracketed by cudaProfilerStart/Stop so that set-up work (weight packing, calibration) is not profiled.

    ncu --set full --clock-control none --import-source on --profile-from-start off -c 300 -o gpurun_out/r2_full python tools/profile_kernels.py
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import _pkg  # noqa: E402

_pkg.load()
from bench import synthetic_cond  # noqa: E402
from diffsound_b200.utils import builders  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--layers", type=int, default=2, help="denoiser layers to build (each layer repeats the same 11 launches)")
ap.add_argument("--precision", default="f16x3")
ap.add_argument("--parts", default="vocoder,decoder,denoiser")
a = ap.parse_args()
B = a.batch
dalle = builders.build_dalle(K=256, NL=a.layers, precision=a.precision, seed=0)
voc = builders.build_vocoder(os.path.join(ROOT, "oracle", "_ref", "best_netG.pt"))
dalle.content_codec.engine.use_cuda_graph = False
voc.engine.use_cuda_graph = False
tok = torch.randint(0, 256, (B, 265), device="cuda")
mel = dalle.decode_to_img(tok, (B, 256, 5, 53))     # warm-up: packs weights, sizes workspaces
wav = voc((mel[:, 0] + 1) / 2)
tr = dalle.transformer
eng = tr.transformer.engine
cond = synthetic_cond(B, 1).cuda()
kv = eng.encode_condition(cond)
x = torch.full((B, 265), 256, dtype=torch.long, device="cuda")
t = torch.full((B,), 99, dtype=torch.long, device="cuda")
u = torch.rand(B, 257, 265, device="cuda")
from diffsound_b200 import ops  # noqa: E402
logits = eng.forward(x, kv, t, 77)
ops.posterior_sample(logits, x, t, u, tr._sched(), T=100, trunc_mode=1, trunc_r=0.85)
torch.cuda.synchronize()
torch.cuda.profiler.start()
if "vocoder" in a.parts:
    voc((mel[:, 0] + 1) / 2)
if "decoder" in a.parts:
    dalle.decode_to_img(tok, (B, 256, 5, 53))
if "denoiser" in a.parts:
    logits = eng.forward(x, kv, t, 77)
    ops.posterior_sample(logits, x, t, u, tr._sched(), T=100, trunc_mode=1, trunc_r=0.85)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("launches: vocoder", voc.engine.launches, "decoder", dalle.content_codec.engine.launches, "denoiser", eng.launches_per_forward + 1)
