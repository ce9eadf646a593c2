"""Target of the ncu captures (profiles/r2_*): one eager pass of ONE part of the pipeline at the bench shape (B=16), bracketed by
cudaProfilerStart/Stop so that set-up work (weight packing, calibration, warm-up) is not profiled.

    ncu --set full --clock-control none --profile-from-start off -c 30 -o gpurun_out/r2_full_denoiser python tools/profile_kernels.py --part denoiser
    ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_launches_decoder.csv \
        python tools/profile_kernels.py --part decoder
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
from diffsound_b200 import ops  # noqa: E402
from diffsound_b200.utils import builders  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--layers", type=int, default=2, help="denoiser layers to build (each layer repeats the same 11 launches)")
ap.add_argument("--precision", default="f16x3")
ap.add_argument("--part", default="denoiser", choices=["vocoder", "decoder", "denoiser"])
a = ap.parse_args()
B = a.batch
if a.part == "vocoder":
    voc = builders.build_vocoder(os.path.join(ROOT, "oracle", "_ref", "best_netG.pt"))
    voc.engine.use_cuda_graph = False
    mel = torch.rand(B, 80, 848, device="cuda")
    run = lambda: voc(mel)
elif a.part == "decoder":
    dalle = builders.build_dalle(K=256, NL=1, precision=a.precision, seed=0)
    dalle.content_codec.engine.use_cuda_graph = False
    tok = torch.randint(0, 256, (B, 265), device="cuda")
    run = lambda: dalle.decode_to_img(tok, (B, 256, 5, 53))
else:
    tr = builders.build_diffusion_transformer(256, 1024, a.layers, 16, 512, precision=a.precision)
    eng = tr.transformer.engine
    kv = eng.encode_condition(synthetic_cond(B, 1).cuda())
    x = torch.full((B, 265), 256, dtype=torch.long, device="cuda")
    t = torch.full((B,), 99, dtype=torch.long, device="cuda")
    u = torch.rand(B, 257, 265, device="cuda")

    def run():
        logits = eng.forward(x, kv, t, 77)
        ops.posterior_sample(logits, x, t, u, tr._sched(), T=100, trunc_mode=1, trunc_r=0.85)
run()  # warm-up: packs weights, sizes workspaces
torch.cuda.synchronize()
torch.cuda.profiler.start()
run()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled one pass of", a.part)
