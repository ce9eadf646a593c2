#!/usr/bin/env python
"""bench.py -- Diffsound hot path on B200: clips/s for 10 s clips at 100 diffusion steps (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch 16] [--codebook 256]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...
    python bench.py --impl reference ...      # the reference algorithm's CPU path (oracle port) on the host cores

One "step" = one DiffusionTransformer.sample() (100 sequential p_sample steps, top0.85r truncation) over a batch of
synthetic caption embeddings = `batch` clips per GPU.  Weak scaling: every rank samples its own batch; finished token
grids are all-gathered over NCCL inside the timed region.  One JSON line on stdout (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_CLIP_STEP = {256: 158.25e9, 512: 158.39e9}  # denoiser FLOPs per clip per diffusion step


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons while the timed region runs (B200_PROFILING.md clocks line)."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.rows, self._halt = index, [], threading.Event()

    def run(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self._halt.wait(0.2)

    def stop(self):
        self._halt.set()
        self.join(timeout=3)
        sm = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows for i in range(4) if len(r) >= 7 and r[3 + i].lower().startswith("active")})
        mx = [int(float(r[1])) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(self.rows)}


def synthetic_cond(B, seed, cond_dim=512):
    g = torch.Generator().manual_seed(seed)
    c = torch.randn(B, 77, cond_dim, generator=g)
    return c / c.norm(dim=-1, keepdim=True)  # normalize: True (clip_text_embedding.py:78-79)


# ------------------------------------------------------------------------------------------------ CPU arm (oracle port)
def cpu_port_clips_per_s(K, B, steps_sample, n_layer=19, threads=None):
    """Time the oracle port of sample() on the host cores for `steps_sample` of the 100 steps and extrapolate linearly."""
    from oracle import diffsound_oracle as O
    sd = O.make_transformer_state_dict(K=K, D=1024, n_layer=n_layer, n_head=16, cond_dim=512, seed=0)
    cond = synthetic_cond(B, 1)
    if threads is None:
        # "all the host threads it can use": torch's intra-op pool stops scaling (and then regresses) long before 100+
        # threads on M = B*265 row GEMMs, so probe a few pool sizes on one layer and keep the fastest.
        ncpu = os.cpu_count() or 1
        x = torch.randn(B, 265, 1024)
        tt = torch.full((B,), 50, dtype=torch.long)
        best = (float("inf"), 1)
        for n in sorted({min(n, ncpu) for n in (8, 16, 32, 64, ncpu)}):
            torch.set_num_threads(n)
            with torch.no_grad():
                O.transformer_block(sd, "transformer.blocks.0.", x, cond, tt, 16)
                t0 = time.perf_counter()
                for _ in range(2):
                    O.transformer_block(sd, "transformer.blocks.0.", x, cond, tt, 16)
                dt = time.perf_counter() - t0
            best = min(best, (dt, n))
        threads = best[1]
    torch.set_num_threads(threads)
    gen = torch.Generator().manual_seed(1234)
    steps = list(range(99, 99 - steps_sample, -1))
    with torch.no_grad():
        O.sample(sd, cond, gen, n_layer=n_layer, n_head=16, spatial=(5, 53), steps=steps[:1])  # warm-up step
        t0 = time.perf_counter()
        O.sample(sd, cond, gen, n_layer=n_layer, n_head=16, spatial=(5, 53), steps=steps)
        dt = time.perf_counter() - t0
    per_step = dt / len(steps)
    return B / (per_step * 100.0), threads, f"B={B}, {len(steps)} of 100 p_sample steps (t=99..{steps[-1]}), x{100 / len(steps):g} linear extrapolation, {dt:.1f} s of CPU work"


def run_reference_arm(args):
    """CPU arm: the reference algorithm (oracle port -- the reference itself is Python under /root/reference and cannot travel to the
    GPU box) on the host cores.  One bounded sample (B=4, 5 x steps diffusion steps, linearly extrapolated to 100) keeps the run short."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    v, cores, sample = cpu_port_clips_per_s(args.codebook, min(args.batch, 4), max(5, min(50, 5 * args.steps)), n_layer=args.layers)
    line = {"impl": "reference", "metric": "clips/sec (10s audio, 100 diffusion steps)", "value": v, "unit": "clips/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * args.batch / v, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"Diffsound AudioCaps inference: batch {args.batch}, 100 steps, K={args.codebook}, 265-token grid (CPU port of the reference algorithm)"},
            "cpu_baseline": {"value": v, "unit": "clips/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ GPU arm
def build_model(K, n_layer=19, precision="f16"):
    import _pkg
    _pkg.load()
    from tests.test_gpu_transformer import build_dt
    torch.manual_seed(0)
    m = build_dt(K, 1024, n_layer, 16, 512, precision=precision)
    m.truncation = "top0.85r"
    return m


def time_events(fn, iters, stream):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(stream)
    for _ in range(iters):
        fn()
    e.record(stream)
    e.synchronize()
    return s.elapsed_time(e) / iters  # ms


def gemm_roofline(model, B, peaks, peaks_src):
    """Average device time of the dominant kernel (gemm_tcgen05_kernel) over the launches of one denoiser pass, CUDA events on
    the launching stream; every launch uses a different layer's weights so nothing is L2-warm."""
    eng = model.transformer.engine
    L, D = 265, eng.D
    ws = eng.workspace(B, L)
    M = B * L
    st = torch.cuda.current_stream()
    h2, x2 = ws["h"].view(M, D), ws["x"].view(M, D)
    lin = eng._linear
    shapes = [("qkv", lambda l: lin(h2, l["wqkv"], l["bqkv"], out=ws["qkv"]), 3 * D, D),
              ("proj1", lambda l: lin(ws["att"], l["wo1"], l["bo1"], residual=x2, out=x2), D, D),
              ("q2", lambda l: lin(h2, l["wq2"], l["bq2"], out=ws["q2"]), D, D),
              ("proj2", lambda l: lin(ws["att"], l["wo2"], l["bo2"], residual=x2, out=x2), D, D),
              ("mlp1", lambda l: lin(h2, l["w1"], l["b1"], out=ws["hid"], gelu=True, round_out=True), 4 * D, D),
              ("mlp2", lambda l: lin(ws["hid"], l["w2"], l["bm2"], residual=x2, out=x2), D, 4 * D)]
    ws["x"].normal_(); ws["h"].normal_(); ws["att"].normal_(); ws["hid"].normal_()
    per = {}
    tot_ms, tot_flop, launches = 0.0, 0.0, 0
    for name, fn, N, Kd in shapes:
        def run_all():
            for lay in eng.layers:
                fn(lay)
        run_all()
        torch.cuda.synchronize()
        # replay the 19 launches from a CUDA graph so the measurement is device time, not Python launch overhead
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            run_all()
        g.replay()
        torch.cuda.synchronize()
        ms = time_events(g.replay, 5, torch.cuda.current_stream()) / len(eng.layers)
        fl = 2.0 * M * N * Kd
        per[name] = {"us": round(ms * 1e3, 2), "tflops": round(fl / (ms * 1e-3) / 1e12, 1)}
        tot_ms += ms
        tot_flop += fl
        launches += 1
    achieved = tot_flop / (tot_ms * 1e-3) / 1e12
    peak_bf16 = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops")))
    tf32 = eng.precision == "tf32"
    peak = peak_bf16 / 2.0 if tf32 else peak_bf16  # kind::tf32 issues at half the kind::f16 rate; fp16 == bf16 rate
    return {"bound": "tensor", "kernel": f"gemm_tcgen05_kernel<{eng.precision}>", "achieved": round(achieved, 1), "peak": round(peak, 1), "unit": "TFLOP/s",
            "frac": round(achieved / peak, 4),
            # dram__bytes_read.sum + dram__bytes_write.sum per launch, mean over the six layer GEMMs of one `ncu --set full` capture
            # (profiles/r1_d_ncu_full_layer.md; B=16 f16 run; outputs stay L2-resident, so this is below the 46.6 MB algorithmic bytes)
            "traffic": 27.0e6 if (B == 16 and eng.precision == "f16") else None, "traffic_unit": "bytes/launch",
            "peak_source": f"{peaks_src} MEASURED_PEAKS.json bf16_tflops_sustained={peak_bf16}" + (" / 2 (tf32 dense rate is half of bf16)" if tf32 else
                           " (cuBLAS bf16 GEMM inside a long loop; kind::f16 fp16 operands issue at the same rate)"),
            "per_gemm": per,
            "flop_per_launch_avg": tot_flop / launches, "us_per_launch_avg": round(tot_ms * 1e3 / launches, 2)}


def full_pipeline_probe(model, cond_dev, B):
    """Secondary figure (BASELINE configs[2] shape at this batch): tokens -> SpecVQGAN decoder -> MelGAN on top of the sampler."""
    import _pkg
    _pkg.load()
    from diffsound_b200.modeling.codecs.spec_codec.vqgan import VQModel
    from diffsound_b200.vocoder.modules import Generator
    dd = dict(double_z=False, z_channels=256, resolution=848, in_channels=1, out_ch=1, ch=128, ch_mult=[1, 1, 2, 2, 4], num_res_blocks=2,
              attn_resolutions=[53], dropout=0.0)
    torch.manual_seed(0)
    vq = VQModel(dd, None, n_embed=256, embed_dim=256).cuda().eval()
    ck = os.path.join(ROOT, "oracle", "_ref", "best_netG.pt")
    voc = Generator(80, 32, 3)  # the shipped MelGAN weights when build() staged them; otherwise the module's own random initialisation
    if os.path.exists(ck):
        voc.load_state_dict(torch.load(ck, map_location="cpu"), strict=True)
    voc = voc.cuda().eval()

    def run():
        tok = model.sample(condition_token=None, condition_mask=None, condition_embed=cond_dev, filter_ratio=0, batch_size=B)["content_token"]
        tok = tok % 256
        mel = vq.decode_tokens(tok, (5, 53))
        return voc((mel[:, 0] + 1) / 2)

    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    run(); torch.cuda.synchronize()
    st = torch.cuda.current_stream()
    ev[0].record(st)
    tok = model.sample(condition_token=None, condition_mask=None, condition_embed=cond_dev, filter_ratio=0, batch_size=B)["content_token"] % 256
    ev[1].record(st)
    mel = vq.decode_tokens(tok, (5, 53))
    ev[2].record(st)
    wav = voc((mel[:, 0] + 1) / 2)
    ev[3].record(st)
    torch.cuda.synchronize()
    t = [ev[i].elapsed_time(ev[i + 1]) for i in range(3)]
    return {"clips_per_s": B / (sum(t) * 1e-3), "batch": B, "ms": {"sampler_100_steps": round(t[0], 2), "decoder": round(t[1], 2), "vocoder": round(t[2], 2)},
            "precision": "sampler f16 operands; decoder/vocoder split-TF32 (3-pass)", "wav_shape": list(wav.shape),
            "launches": {"decoder": vq.engine.launches, "vocoder": voc.engine.launches}}


def run_gpu_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    torch.cuda.set_device(local)
    if world > 1:
        os.environ["NCCL_DEBUG"] = os.environ.get("DSB_NCCL_DEBUG", "WARN")  # keep stdout to the one JSON line
        import torch.distributed as dist
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    B, K = args.batch, args.codebook
    model = build_model(K, args.layers, args.precision)
    cond_host = synthetic_cond(B, 1 + rank).pin_memory()
    cond_dev = cond_host.to(dev)
    gathered = [torch.empty(B, 265, dtype=torch.int64, device=dev) for _ in range(world)] if world > 1 else None
    tok_host = torch.empty(B, 265, dtype=torch.int64).pin_memory()

    def one_clip_batch(e2e: bool):
        c = cond_host.to(dev, non_blocking=True) if e2e else cond_dev
        tok = model.sample(condition_token=None, condition_mask=None, condition_embed=c, filter_ratio=0, batch_size=B)["content_token"]
        if world > 1:
            dist.all_gather(gathered, tok)  # the only collective of the path: gather finished clips
        if e2e:
            tok_host.copy_(tok, non_blocking=True)
        return tok

    def timed(e2e: bool, steps: int):
        torch.manual_seed(1234 + rank)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        st = torch.cuda.current_stream()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(st)
        for _ in range(steps):
            one_clip_batch(e2e)
        e.record(st)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = torch.tensor([s.elapsed_time(e)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for _ in range(max(args.warmup, 3)):
        one_clip_batch(False)
    torch.cuda.synchronize()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    ms_dev = timed(False, args.steps)
    ms_e2e = timed(True, args.steps)
    clk = clocks.stop() if rank == 0 else None
    launches = model.last_gpu_launches * args.steps
    clips = B * world * args.steps
    value = clips / (ms_dev * 1e-3)
    e2e_v = clips / (ms_e2e * 1e-3)
    full = full_pipeline_probe(model, cond_dev, B) if (rank == 0 and world == 1 and not args.no_full_pipeline) else None
    if rank == 0:
        peaks, src = load_peaks()
        roof = gemm_roofline(model, B, peaks, src)
        flops = FLOP_PER_CLIP_STEP.get(K, 158.3e9) * 100 * (args.layers / 19.0)
        roof["pipeline_tflops"] = round(value * flops / 1e12 / world, 1)
        if args.no_cpu_baseline or world > 1:  # the CPU arm is timed at N=1 only
            cpu = None
        else:
            v, cores, sample = cpu_port_clips_per_s(K, 4, 20, n_layer=args.layers)  # ~10-20 s of CPU work
            cpu = {"value": v, "unit": "clips/s", "cores": cores, "kind": "port", "sample": sample}
        line = {"metric": "clips/sec (10s audio, 100 diffusion steps)", "value": value, "unit": "clips/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": {"f16x3": "f16x3", "f16": "f16", "tf32": "tf32", "fp32": "f32"}[args.precision], "data": "synthetic",
                "config": {"workload": f"Diffsound AudioCaps inference: batch {B}/GPU, 100 steps, K={K} codebook, 265-token grid, top0.85r "
                                       f"(BASELINE.json configs[1]); {args.layers}-layer D=1024 denoiser, random-init weights, synthetic caption embeddings",
                           "arithmetic": "GEMM operands fp16 (11-bit significand, = TF32), fp32 accumulation; residual stream / LayerNorm / softmax fp32; log_softmax fp64",
                           "global_batch": B * world, "parallelism": f"dp{world} (independent captions per rank, all_gather of tokens)",
                           "l2_policy": "working set per diffusion step (1.53 GB fp32 weights) exceeds the 126 MB L2; no explicit flush"},
                "e2e": {"value": e2e_v, "unit": "clips/s", "h2d_bytes_per_step": cond_host.numel() * 4, "d2h_bytes_per_step": tok_host.numel() * 8,
                        "api": "DiffusionTransformer.sample(condition_embed=<pinned host tensor -> device>) -> tokens copied to pinned host"},
                "gpu_launches": launches, "roofline": roof, "cpu_baseline": cpu, "clocks": clk, "full_pipeline": full}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=16, help="clips per GPU per step (BASELINE configs[1]: 16)")
    ap.add_argument("--codebook", type=int, default=256)
    ap.add_argument("--layers", type=int, default=19)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-full-pipeline", action="store_true")
    ap.add_argument("--precision", default="f16", choices=["f16x3", "f16", "tf32", "fp32"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback (use --impl reference for the CPU arm)")
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
