"""Debug aid: run the MelGAN engine on the GPU and on the CPU emulation of the kernel contracts (tests/cpu_state_gemm_emulation.py) call by call,
and report the first launch after which any state buffer differs."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _pkg  # noqa: E402

_pkg.load()
from diffsound_b200 import ops, packing, vocoder_engine  # noqa: E402
from diffsound_b200.vocoder.modules import Generator  # noqa: E402
from tests import cpu_state_gemm_emulation as E  # noqa: E402

ngf, T = int(sys.argv[1]), int(sys.argv[2])
torch.manual_seed(0)
mel = torch.rand(2, 80, T)
NAMES = ("gemm_desc", "mel_pack_f16", "edge_pad_f16")


def run(gen, device):
    eng = gen.engine
    eng.use_cuda_graph = False
    snaps = []
    real = {n: getattr(ops, n) for n in NAMES}

    def wrap(n):
        def f(*a, **k):
            r = real[n](*a, **k)
            if device == "cuda":
                torch.cuda.synchronize()
                print("ok", n, {q: k.get(q) for q in ("M", "N", "K", "batch", "flags")}, flush=True) if os.environ.get("VERBOSE") else None
            bufs = eng._bufs.get((2, T))
            if bufs is not None:
                if device == "cuda":
                    torch.cuda.synchronize()
                snaps.append((n, k.get("N"), k.get("M"), [t.detach().float().cpu().clone() for t in bufs[0] + [y for y in bufs[1] if y is not None]]))
            return r
        return f
    for n in NAMES:
        setattr(ops, n, wrap(n))
    try:
        eng.repack()
        wav = eng._forward(mel.to(device))
    finally:
        for n in NAMES:
            setattr(ops, n, real[n])
    return snaps, wav.float().cpu()


g_gpu = Generator(80, ngf, 3).cuda().eval()
sd = {k: v.detach().cpu() for k, v in g_gpu.state_dict().items()}
s_gpu, w_gpu = run(g_gpu, "cuda")

# CPU emulation
for n in ("gemm_desc", "mel_pack_f16", "edge_pad_f16", "split_f16"):
    setattr(ops, n, getattr(E, n))
rz, re_ = torch.zeros, torch.empty
vocoder_engine.torch.zeros = lambda *a, **k: E.track(rz(*a, **k))
vocoder_engine.torch.empty = lambda *a, **k: E.track(re_(*a, **k))
g_cpu = Generator(80, ngf, 3).eval()
g_cpu.load_state_dict(sd)
_init = packing.PackedConv.__init__


def _tracked_init(self, blocks, bias):
    _init(self, blocks, bias)
    E.track(self.w)


packing.PackedConv.__init__ = _tracked_init
packing.torch.zeros = vocoder_engine.torch.zeros
s_cpu, w_cpu = run(g_cpu, "cpu")
print("launches", len(s_gpu), len(s_cpu))
for i, ((n, N, M, a), (_, _, _, b)) in enumerate(zip(s_gpu, s_cpu)):
    errs = [float((x - y).abs().max()) for x, y in zip(a, b)]
    bad = [j for j, e in enumerate(errs) if not (e < 1e-2)]
    print(i, n, "N", N, "M", M, "max abs diff per buffer", ["%.2e" % e for e in errs], "<-- DIFF" if bad else "")
    if bad:
        j = bad[0]
        d = (a[j] - b[j]).abs()
        idx = (d > 1e-2).nonzero()
        print("   first differing buffer", j, "shape", tuple(a[j].shape), "count", idx.shape[0], "first", idx[:5].tolist(), "last", idx[-5:].tolist())
        break
print("wav max abs diff", float((w_gpu - w_cpu).abs().max()))
