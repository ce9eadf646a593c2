"""Device-time micro-benchmark of the attention kernels at the bench shapes (B=16, 16 heads, L=265 / Lc=77)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _pkg; _pkg.load()
from diffsound_b200 import ops

def timeit(fn, reps=20):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(3):
        g.replay()
    e.record(); e.synchronize()
    return s.elapsed_time(e) / (3 * reps) * 1e3  # us

if __name__ != "__main__":
    B = None
else:
  B, H, L, Lc, D = 16, 16, 265, 77, 1024
  qkv = torch.randn(B * L, 3 * D, device="cuda").half()
  kv = torch.randn(B * Lc, 38912, device="cuda").half()
  q2 = torch.randn(B * L, D, device="cuda").half()
  out = torch.empty(B * L, D, device="cuda", dtype=torch.float16)
  us = timeit(lambda: ops.attention(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], out, B=B, H=H, Lq=L, Lk=L, scale=0.125))
  print(f"self-attention  f16: {us:7.1f} us  {4 * B * H * L * L * 64 / us / 1e6:7.1f} TFLOP/s")
  us = timeit(lambda: ops.attention(q2, kv[:, :D], kv[:, D:2 * D], out, B=B, H=H, Lq=L, Lk=Lc, scale=0.125))
  print(f"cross-attention f16: {us:7.1f} us  {4 * B * H * L * Lc * 64 / us / 1e6:7.1f} TFLOP/s")
  us = timeit(lambda: ops.attention_tc(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], out, B=B, H=H, Lq=L, Lk=L, scale=0.125))
  print(f"self-attention  tcgen05: {us:7.1f} us  {4 * B * H * L * L * 64 / us / 1e6:7.1f} TFLOP/s")
  us = timeit(lambda: ops.attention_tc(q2, kv[:, :D], kv[:, D:2 * D], out, B=B, H=H, Lq=L, Lk=Lc, scale=0.125))
  print(f"cross-attention tcgen05: {us:7.1f} us  {4 * B * H * L * Lc * 64 / us / 1e6:7.1f} TFLOP/s")
  if len(sys.argv) > 1 and __name__ == "__main__": sys.exit(0)
  x = torch.randn(B, L, D, device="cuda"); tab = torch.randn(100, 2 * D, device="cuda"); t = torch.full((B,), 5, device="cuda", dtype=torch.long)
  h = torch.empty(B, L, D, device="cuda", dtype=torch.float16)
  print(f"ada_layernorm (L2-warm x): {timeit(lambda: ops.ada_layernorm(x, tab, t, out=h)):7.1f} us")
  logits = torch.randn(B, L, 256, device="cuda"); u = torch.rand(B, 257, L, device="cuda"); xt = torch.full((B, L), 256, device="cuda", dtype=torch.long)
  from oracle import diffsound_oracle as O
  sb = O.schedule_buffers(100, 257); sched = torch.zeros(8, 101)
  for i, n in enumerate(["log_at", "log_bt", "log_ct", "log_1_min_ct", "log_cumprod_at", "log_cumprod_bt", "log_cumprod_ct", "log_1_min_cumprod_ct"]):
      sched[i, :sb[n].numel()] = sb[n]
  sched = sched.cuda(); xn = torch.empty_like(xt)
  print(f"posterior_sample top0.85r: {timeit(lambda: ops.posterior_sample(logits, xt, t, u, sched, T=100, x_next=xn)):7.1f} us")
  print(f"posterior_sample no trunc: {timeit(lambda: ops.posterior_sample(logits, xt, t, u, sched, T=100, trunc_mode=0, x_next=xn)):7.1f} us")
