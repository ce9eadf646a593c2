"""Ladder (iii) of SURVEY.md 7.2: free-running 100-step sample() on the GPU vs the fp32 CPU oracle, same weights, same uniforms.
Prints the per-step token agreement and the final agreement for the f16 / fp32-GEMM precisions."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _pkg; _pkg.load()
from diffsound_b200 import ops
from oracle import diffsound_oracle as O
from diffsound_b200.utils.builders import build_diffusion_transformer as build_dt
NL = int(sys.argv[1]) if len(sys.argv) > 1 else 19
K, D, NH, CD, B, L = 256, 1024, 16, 512, 1, 265
torch.set_num_threads(16)
sd = O.make_transformer_state_dict(K=K, D=D, n_layer=NL, n_head=NH, cond_dim=CD, seed=0)
g = torch.Generator().manual_seed(5)
cond = torch.randn(B, 77, CD, generator=g); cond = cond / cond.norm(dim=-1, keepdim=True)
us = [torch.rand(B, K + 1, L, generator=g) for _ in range(100)]
t0 = time.time()
ref, trace = O.sample(sd, cond, lambda i: us[i], n_layer=NL, n_head=NH, spatial=(5, 53), return_trace=True)
print(f"oracle 100 steps: {time.time() - t0:.1f} s")
for prec in ("f16", "fp32"):
    m = build_dt(K, D, NL, NH, CD, sd, precision=prec)
    eng = m.transformer.engine
    kv = eng.encode_condition(cond.cuda())
    x = torch.full((B, L), K, dtype=torch.long, device="cuda")
    agree = []
    for i, ti in enumerate(range(99, -1, -1)):
        t = torch.full((B,), ti, dtype=torch.long, device="cuda")
        logits = eng.forward(x, kv, t, 77)
        x = ops.posterior_sample(logits, x, t, us[i].cuda(), m._sched(), T=100)
        agree.append(float((x.cpu() == trace[i]["x_out"]).float().mean()))
    print(f"[{prec}] free-running token agreement: after 10 steps {agree[9]:.4f}, 50 steps {agree[49]:.4f}, final {agree[-1]:.4f}; "
          f"first divergence at step {next((i for i, a in enumerate(agree) if a < 1.0), None)}")
    del m
