"""MelGAN generator on sm_100a: every Conv1d / ConvTranspose1d is a tcgen05 GEMM over channels-last (B, T, C) activations.

* weight_norm (w = g * v / ||v||, reference vocoder/modules.py:18-23) is folded once at pack time instead of every forward;
* Conv1d(k, dilation d) = GEMM with k taps (row shifts 0, d, 2d, ...) over a reflection-padded copy of the input;
* ConvTranspose1d(stride r, kernel 2r, padding r/2) is evaluated in polyphase form: output time q*r + ph only touches
  inputs q-1, q (ph < r/2) or q, q+1 (ph >= r/2), so it is two GEMMs with 2 taps each whose (T, r*Cout) row-major output IS
  the (r*T, Cout) channels-last result -- no zero-stuffing, no scatter;
* LeakyReLU / tanh / bias / residual live in the GEMM epilogue; only the reflection pad of a ResnetBlock input needs its own
  (HBM-bound) pass.
Reference: vocoder/modules.py:72-85 (ResnetBlock), :88-130 (Generator).
"""
from __future__ import annotations

import torch

from . import ops
from .graphs import GraphCache


def _fold(m) -> torch.Tensor:
    v, g = m.weight_v.detach().float(), m.weight_g.detach().float()
    return g * v / v.flatten(1).norm(dim=1).view(-1, *([1] * (v.dim() - 1)))


class VocoderEngine:
    def __init__(self, gen, precision: str = "tf32x3"):
        """precision 'tf32x3' (split-TF32, fp32-class accuracy: the shipped checkpoint's weights span a wide dynamic range and
        single-pass TF32 gives 2.6e-2 waveform error) or 'tf32' (single pass)."""
        if precision not in ("tf32x3", "tf32"):
            raise ValueError("precision must be 'tf32x3' or 'tf32'")
        self.gen = gen
        self.precision = precision
        self.packed = False
        self.launches = 0
        self.use_cuda_graph = True
        self.max_batch = 32  # clips per pass (27.8 MB fp32 per clip for the widest activation)
        self._graphs = GraphCache()

    def _pack(self, w2d: torch.Tensor, ntaps: int) -> torch.Tensor:
        """(N, ntaps*Cin) tap-major fp32 -> packed GEMM weight for the selected precision."""
        w2d = w2d.contiguous().float()
        return ops.pack_split_weight(w2d, ntaps) if self.precision == "tf32x3" else ops.round_tf32(w2d)

    def _pack_conv1d(self, w) -> torch.Tensor:
        return self._pack(w.permute(0, 2, 1).reshape(w.shape[0], -1), w.shape[2])  # (Cout, k*Cin), tap-major

    def _mm(self, a, w, bias=None, residual=None, out=None, presplit=False, **kw):
        if self.precision == "tf32x3":
            kw.pop("round_out", None)
            return ops.gemm_split(a if presplit else ops.split_tf32(a), w, bias, residual, out, **kw)
        return ops.gemm(a, w, bias, residual, out, **kw)

    @torch.no_grad()
    def repack(self):
        mods = list(self.gen.model)
        if mods[1].bias.device.type != "cuda":
            raise RuntimeError("VocoderEngine needs the module on a CUDA device (no CPU fallback)")
        f = lambda p: p.detach().float().contiguous()
        self.first = (self._pack_conv1d(_fold(mods[1])), f(mods[1].bias))
        self.stages = []
        i = 2
        for r in self.gen.ratios:
            ct = mods[i + 1]
            w = _fold(ct)  # (Cin, Cout, 2r); weight_norm dim=0 -> norm over (Cout, k) per input channel
            p = r // 2 + r % 2
            assert r % 2 == 0, "polyphase split assumes even stride (the Diffsound ratios 8,8,2,2)"
            half = r - p
            # phases [0, half): taps (q-1 -> k=ph+p+r, q -> k=ph+p); phases [half, r): taps (q -> k=ph+p, q+1 -> k=ph+p-r)
            wa = torch.stack([torch.cat([w[:, :, ph + p + r].t(), w[:, :, ph + p].t()], dim=1) for ph in range(half)], 0)
            wb = torch.stack([torch.cat([w[:, :, ph + p].t(), w[:, :, ph + p - r].t()], dim=1) for ph in range(half, r)], 0)
            cout = w.shape[1]
            st = dict(r=r, cout=cout, half=half, wa=self._pack(wa.reshape(half * cout, -1), 2),
                      wb=self._pack(wb.reshape((r - half) * cout, -1), 2),
                      ba=f(ct.bias).repeat(half), bb=f(ct.bias).repeat(r - half), res=[])
            i += 2
            for _ in range(self.gen.n_residual_layers):
                rb = mods[i]
                st["res"].append(dict(d=rb.dilation, wd=self._pack_conv1d(_fold(rb.block[2])), bd=f(rb.block[2].bias),
                                      w1=self._pack_conv1d(_fold(rb.block[4])), b1=f(rb.block[4].bias),
                                      ws=self._pack_conv1d(_fold(rb.shortcut)), bs=f(rb.shortcut.bias)))
                i += 1
            self.stages.append(st)
        last = mods[i + 2]
        self.last = (self._pack_conv1d(_fold(last)), f(last.bias))
        self.packed = True
        self._graphs.clear()

    @torch.no_grad()
    def forward(self, mel: torch.Tensor) -> torch.Tensor:
        if not self.packed:
            self.repack()
        if not mel.is_cuda:
            raise RuntimeError("VocoderEngine.forward needs a CUDA tensor (no CPU fallback)")
        mel = mel.detach().float().contiguous()
        if mel.shape[0] > self.max_batch:
            return torch.cat([self.forward(mel[i:i + self.max_batch]) for i in range(0, mel.shape[0], self.max_batch)], 0)
        if self.use_cuda_graph:
            return self._graphs.run(tuple(mel.shape), self._forward, mel)
        return self._forward(mel)

    def _forward(self, mel: torch.Tensor) -> torch.Tensor:
        B, Cm, T = mel.shape
        n = 0
        rnd = self.precision == "tf32"
        sp = not rnd
        xl = ops.lrelu_pad(mel, 3, slope=1.0, reflect=True, channel_major=True, round_out=rnd, split=sp)          # ReflectionPad1d(3) of the mel, channels-last
        w0, b0 = self.first
        # conv k=7 -> LeakyReLU (the activation in front of the first ConvTranspose1d) fused in the epilogue
        x = self._mm(xl, w0, b0, taps=list(range(7)), out_rows=T, lrelu=True, round_out=rnd, presplit=sp)   # (B, T, 16*ngf)
        n += 2
        for si, st in enumerate(self.stages):
            r, cout, half = st["r"], st["cout"], st["half"]
            y = torch.empty(B, T, r * cout, dtype=torch.float32, device=mel.device)
            self._mm(x, st["wa"], st["ba"], out=y[:, :, : half * cout], taps=[-1, 0], round_out=rnd)
            self._mm(x, st["wb"], st["bb"], out=y[:, :, half * cout:], taps=[0, 1], round_out=rnd)
            T = T * r
            x = y.view(B, T, cout)
            n += 2
            for ri, rb in enumerate(st["res"]):
                d = rb["d"]
                xl = ops.lrelu_pad(x, d, slope=0.2, reflect=True, round_out=rnd, split=sp)                          # LeakyReLU + ReflectionPad1d(d)
                y1 = self._mm(xl, rb["wd"], rb["bd"], taps=[0, d, 2 * d], out_rows=T, lrelu=True, round_out=rnd, presplit=sp)
                tmp = self._mm(y1, rb["w1"], rb["b1"])
                last_of_stage = ri == len(st["res"]) - 1
                # shortcut(x) + block(x); after the last block of a stage the next consumer is LeakyReLU -> ConvT / final conv
                x = self._mm(x, rb["ws"], rb["bs"], residual=tmp, round_out=rnd, lrelu=last_of_stage, res_before_act=last_of_stage)
                n += 4
        xl = ops.lrelu_pad(x, 3, slope=1.0, reflect=True, round_out=rnd, split=sp)                                  # x already went through LeakyReLU
        wl, bl = self.last
        wav = self._mm(xl, wl, bl, taps=list(range(7)), out_rows=T, tanh=True, presplit=sp)             # (B, T, 1)
        self.launches = n + 2
        return wav.view(B, 1, T)
