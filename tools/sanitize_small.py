"""Small invocation of every hot kernel for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _pkg; _pkg.load()
from diffsound_b200 import ops
from oracle import diffsound_oracle as O
torch.manual_seed(0)
B, H, L, Lc, D, K = 2, 2, 265, 77, 128, 32
qkv = torch.randn(B * L, 3 * D, device="cuda").half(); kv = torch.randn(B * Lc, 2 * D, device="cuda").half()
o = torch.empty(B * L, D, device="cuda", dtype=torch.float16)
ops.attention(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], o, B=B, H=H, Lq=L, Lk=L, scale=0.125)
ops.attention(qkv[:, :D], kv[:, :D], kv[:, D:], o, B=B, H=H, Lq=L, Lk=Lc, scale=0.125)
q32 = qkv.float(); o32 = torch.empty(B * L, D, device="cuda")
ops.attention(q32[:, :D], q32[:, D:2 * D], q32[:, 2 * D:], o32, B=B, H=H, Lq=L, Lk=L, scale=0.125)
a = torch.randn(B * L, D, device="cuda").half(); w = (torch.randn(300, D, device="cuda") * 0.1).half(); bias = torch.randn(300, device="cuda")
res = torch.randn(B * L, 300, device="cuda")
for pair in (-1, 1):
    ops.gemm(a, w, bias, res, None, dtype=ops.F16, gelu=True, cta_pair=pair)
    ops.gemm(a, w, bias, None, None, dtype=ops.F16, out_f16=True, cta_pair=pair)
at = ops.round_tf32(torch.randn(7 * 30, 64, device="cuda")); wt = ops.round_tf32(torch.randn(33, 3 * 64, device="cuda"))
ops.gemm(at, wt, taps=[-7, 0, 7], geo=(70, 7, 1, 9, 1, 6))
ops.gemm_split(ops.split_tf32(at), ops.pack_split_weight(wt, 3), taps=[-7, 0, 7])
x = torch.randn(B, L, D, device="cuda"); tab = torch.randn(100, 2 * D, device="cuda"); t = torch.tensor([5, 99], device="cuda")
ops.ada_layernorm(x, tab, t); ops.layernorm(x, tab[0, :D].contiguous(), tab[0, D:].contiguous())
sd = O.make_transformer_state_dict(K=K, D=D, n_layer=1, n_head=2, cond_dim=64)
p = "transformer.content_emb."
ids = torch.randint(0, K + 1, (B, L), device="cuda")
ops.embed_tokens(ids, sd[p + "emb.weight"].cuda(), sd[p + "height_emb.weight"].cuda(), sd[p + "width_emb.weight"].cuda())
sb = O.schedule_buffers(100, K + 1); sched = torch.zeros(8, 101)
for i, n in enumerate(["log_at", "log_bt", "log_ct", "log_1_min_ct", "log_cumprod_at", "log_cumprod_bt", "log_cumprod_ct", "log_1_min_cumprod_ct"]):
    sched[i, :sb[n].numel()] = sb[n]
logits = torch.randn(B, L, K, device="cuda"); u = torch.rand(B, K + 1, L, device="cuda"); lpo = torch.empty(B, K + 1, L, device="cuda")
ops.posterior_sample(logits, ids, t, u, sched.cuda(), T=100, log_prob_out=lpo)
z = ops.codebook_gather_padded(torch.randint(0, K, (B, 14), device="cuda"), torch.randn(K, 64, device="cuda"), 2, 7, split=True)
xp = torch.zeros(B, 4, 9, 64, device="cuda"); xp[:, 1:-1, 1:-1] = torch.randn(B, 2, 7, 64, device="cuda")
st = ops.groupnorm_stats(xp); ops.groupnorm_apply(xp, st, torch.ones(64, device="cuda"), torch.zeros(64, device="cuda"), split=True)
ops.upsample2x_padded(xp, split=True); ops.lrelu_pad(torch.randn(B, 40, 80, device="cuda"), 3, split=True)
torch.cuda.synchronize(); print("sanitize_small done")
