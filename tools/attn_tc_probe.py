import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _pkg; _pkg.load()
from diffsound_b200 import ops
torch.manual_seed(0)
def relerr(a, b): return float((a.double() - b.double()).abs().max() / b.double().abs().max())
for (B, H, Lq, Lk) in [(1, 1, 128, 64), (1, 1, 128, 272), (2, 2, 265, 265), (2, 16, 265, 77), (1, 1, 16, 5), (16, 16, 265, 265), (16, 16, 265, 77)]:
    D = H * 64
    qkv = torch.randn(B * Lq, 3 * D, device="cuda").half(); kv = torch.randn(B * Lk, 2 * D, device="cuda").half()
    q, k, v = qkv[:, :D], kv[:, :D], kv[:, D:]
    qh = q.double().view(B, Lq, H, 64).transpose(1, 2); kh = k.double().view(B, Lk, H, 64).transpose(1, 2); vh = v.double().view(B, Lk, H, 64).transpose(1, 2)
    ref = (torch.softmax(qh @ kh.transpose(-1, -2) / 8.0, -1) @ vh).transpose(1, 2).reshape(B * Lq, D)
    for pipelined in (False, True):
        out = torch.full((B * Lq, D), float("nan"), device="cuda", dtype=torch.float16)
        try:
            for _ in range(3):
                ops.attention_tc(q, k, v, out, B=B, H=H, Lq=Lq, Lk=Lk, scale=0.125, pipelined=pipelined)
            torch.cuda.synchronize()
            e = relerr(out.float().cpu(), ref.cpu())
            print(f"pipelined={pipelined} B={B} H={H} Lq={Lq} Lk={Lk}: rel err {e:.3e} nan {int(torch.isnan(out).sum())}", "OK" if e < 2e-3 else "MISMATCH")
        except RuntimeError as ex:
            print("ERROR", str(ex)[:300]); sys.exit(1)
