"""Time one full-size training step (BASELINE config 4 shape: per-GPU batch 20, K=256, 19 layers, bf16 operands) through the drop-in
DiffusionTransformer.forward(return_loss=True) + loss.backward() + AdamW.step().  Prints samples/s, peak memory, launch counts."""
import argparse
import json
import os
import sys
import time

import torch

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
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--cpu-baseline", type=int, default=0, metavar="B", help="also time the oracle (torch autograd on the host cores) on B samples")
    a = ap.parse_args()
    K, D, NH, CD, L = 256, 1024, 16, 512, 265
    m = build_dt(K, D, a.layers, NH, CD)
    m.transformer.train_engine.__init__(m.transformer, precision=a.precision)
    m.transformer.train_engine.use_cuda_graph = not a.no_graph
    m.train()
    for p in m.parameters():
        p.requires_grad_(True)
    opt = torch.optim.AdamW(m.parameters(name="transformer"), lr=3e-6, betas=(0.9, 0.96), fused=True)
    g = torch.Generator().manual_seed(0)
    batch = {"content_token": torch.randint(0, K, (a.batch, L), generator=g).cuda(),
             "condition_embed_token": torch.nn.functional.normalize(torch.randn(a.batch, 77, CD, generator=g), dim=-1).cuda()}
    times, parts = [], []
    for it in range(a.steps + 2):
        torch.cuda.synchronize()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        t0 = time.perf_counter()
        e[0].record()
        out = m(batch, return_loss=True, return_logits=False)
        e[1].record()
        opt.zero_grad(set_to_none=True)
        out["loss"].backward()
        e[2].record()
        opt.step()
        e[3].record()
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) * 1e3
        if it >= 2:
            times.append(e[0].elapsed_time(e[3]))
            parts.append((e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2]), e[2].elapsed_time(e[3]), wall))
        print(f"step {it}: loss {float(out["loss"].detach()):.4f} gpu {e[0].elapsed_time(e[3]):.1f} ms (fwd {e[0].elapsed_time(e[1]):.1f} bwd {e[1].elapsed_time(e[2]):.1f} "
              f"opt {e[2].elapsed_time(e[3]):.1f}) wall {wall:.1f} ms", flush=True)
    ms = sum(times) / len(times)
    cpu = None
    if a.cpu_baseline:
        # the reference's training step is torch autograd through the same math: time the oracle restatement on the host cores (bounded sample)
        from oracle import diffsound_oracle as O
        torch.set_num_threads(min(16, os.cpu_count() or 1))
        sd = {k: v.detach().float().cpu() for k, v in m.state_dict().items()}
        leaf = {k: (v.clone().requires_grad_(True) if k.startswith("transformer.") and "attn2.mask" not in k else v) for k, v in sd.items()}
        sched = {k: sd[k] for k in sd if k.startswith("log_")}
        Bc = a.cpu_baseline
        x0 = batch["content_token"][:Bc].cpu()
        cond = batch["condition_embed_token"][:Bc].cpu()
        t = torch.randint(0, 100, (Bc,))
        pt = torch.full((Bc,), 0.01)
        u = torch.rand(Bc, K + 1, L)
        best = None
        for _ in range(2):
            t0 = time.perf_counter()
            out = O.train_loss(leaf, sched, x0, cond, t, pt, u, n_layer=a.layers, n_head=NH, spatial=(5, 53), T=100, aux_weight=5e-4, adaptive_aux=True)
            out["loss"].backward()
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        cpu = {"samples_per_s": Bc / best, "cores": torch.get_num_threads(), "kind": "port", "sample": f"{Bc} samples, forward + autograd backward, fp32"}
    print(json.dumps({"cpu_baseline": cpu, "train_step_ms": ms, "samples_per_s": a.batch / ms * 1e3, "batch": a.batch, "layers": a.layers, "precision": a.precision, "cuda_graph": not a.no_graph,
                      "fwd_ms": sum(p[0] for p in parts) / len(parts), "bwd_ms": sum(p[1] for p in parts) / len(parts),
                      "opt_ms": sum(p[2] for p in parts) / len(parts), "wall_ms": sum(p[3] for p in parts) / len(parts),
                      "peak_mem_gb": torch.cuda.max_memory_allocated() / 2**30}))


if __name__ == "__main__":
    main()
