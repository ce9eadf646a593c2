"""DDP training check (BASELINE config 4 shape): launched with torchrun, one rank per GPU over NCCL.
  1. gradients through torch DistributedDataParallel == the all-reduced average of each rank's stand-alone gradients;
  2. a few timed AdamW steps -> aggregate samples/s (max over ranks, CUDA events)."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import _pkg  # noqa: E402

_pkg.load()
from diffsound_b200.utils.builders import build_diffusion_transformer as build_dt  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=20)
    ap.add_argument("--layers", type=int, default=19)
    ap.add_argument("--steps", type=int, default=5)
    a = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
    dist.init_process_group("nccl")
    K, D, NH, CD, L = 256, 1024, 16, 512, 265
    torch.manual_seed(0)
    m = build_dt(K, D, a.layers, NH, CD).train()
    for p in m.parameters():
        p.requires_grad_(True)
    g = torch.Generator().manual_seed(100 + rank)
    batch = {"content_token": torch.randint(0, K, (a.batch, L), generator=g).cuda(),
             "condition_embed_token": torch.nn.functional.normalize(torch.randn(a.batch, 77, CD, generator=g), dim=-1).cuda()}
    # 1. stand-alone gradients, averaged by hand
    torch.manual_seed(1000 + rank)
    m(batch, return_loss=True, return_logits=False)["loss"].backward()
    expected = []
    for p in m.transformer.parameters():
        e = p.grad.detach().clone()
        dist.all_reduce(e)
        expected.append(e / world)
        p.grad = None
    ddp = torch.nn.parallel.DistributedDataParallel(m, device_ids=[torch.cuda.current_device()])
    torch.manual_seed(1000 + rank)
    ddp(batch, return_loss=True, return_logits=False)["loss"].backward()
    worst = 0.0
    for p, e in zip(m.transformer.parameters(), expected):
        worst = max(worst, float((p.grad - e).abs().max() / e.abs().max().clamp_min(1e-20)))
    # 2. timed steps
    opt = torch.optim.AdamW(m.parameters(name="transformer"), lr=3e-6, betas=(0.9, 0.96), fused=True)
    times = []
    for it in range(a.steps + 2):
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = ddp(batch, return_loss=True, return_logits=False)
        opt.zero_grad(set_to_none=True)
        out["loss"].backward()
        opt.step()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        if it >= 2:
            times.append(float(ms))
    if rank == 0:
        ms = sum(times) / len(times)
        print(json.dumps({"ddp_world": world, "ddp_grad_vs_manual_allreduce_max_rel_err": worst, "train_step_ms": ms,
                          "samples_per_s": world * a.batch / ms * 1e3, "per_gpu_batch": a.batch, "layers": a.layers}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
