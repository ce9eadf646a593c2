import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _pkg; _pkg.load()
from diffsound_b200 import ops
M, N, K = 4240, 1024, int(sys.argv[1]) if len(sys.argv) > 1 else 64
a = torch.randn(M, K, device="cuda").half(); w = (torch.randn(N, K, device="cuda") * 0.05).half()
bias = torch.randn(N, device="cuda"); out = torch.empty(M, N, device="cuda", dtype=torch.float16)
for _ in range(4):
    ops.gemm(a, w, bias, None, out, dtype=ops.F16, gelu=True)
torch.cuda.synchronize()
