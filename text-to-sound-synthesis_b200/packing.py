"""Weight packing for the split-fp16 ("f16x3") conv GEMMs of the SpecVQGAN decoder and the MelGAN generator."""
from __future__ import annotations

import math

import torch

from . import ops


def k64(c: int) -> int:
    """Channels rounded up to the GEMM's 64-element fp16 k-block."""
    return (c + 63) // 64 * 64


class PackedConv:
    """(N, n_blocks * 2 * Kp) fp16: per K-block j (one spatial tap of a conv, or one of several 1x1 operands), columns [j*2Kp, +Cin) hold the hi
    half and [j*2Kp + Kp, +Cin) the lo half of 2^s * W_j; everything else is zero, so an A box that reads Kp columns where only Cin exist multiplies
    the overhang by zeros.  alpha = 2^-s undoes the scale in the GEMM epilogue."""

    def __init__(self, blocks, bias, fold=False):
        """fold (Cin <= 32, operand rows laid out [hi(32) | lo(32)]): per block two 64-deep k-blocks [Wh | Wh] and [Wl | 0], so that ONE A box
        [hi | lo] yields hi*Wh + lo*Wh and hi*Wl -- the three products of the split-fp16 scheme from one staged copy of the operand."""
        N, Cin = blocks[0].shape
        Kp = k64(Cin)
        self.fold = bool(fold)
        if fold and Cin > 32:
            raise ValueError("folded packing needs Cin <= 32")
        amax = max(float(b.abs().max()) for b in blocks)
        s = 0 if amax == 0.0 or not math.isfinite(amax) else 13 - math.frexp(amax)[1]
        w = torch.zeros(N, len(blocks), 2, Kp, dtype=torch.float16, device=blocks[0].device)
        for j, b in enumerate(blocks):
            pr = ops.split_f16(b.detach().contiguous().float(), 2.0 ** s)  # (N, 2*Cin)
            w[:, j, 0, :Cin] = pr[:, :Cin]
            if fold:
                w[:, j, 0, 32:32 + Cin] = pr[:, :Cin]
            w[:, j, 1, :Cin] = pr[:, Cin:]
        self.w = w.reshape(N, -1).contiguous()
        self.alpha, self.Kp, self.N, self.nblk = 2.0 ** (-s), Kp, N, len(blocks)
        self.bias = bias.detach().float().contiguous()

    def taps(self, spatial):
        """spatial: per K-block (row_shift, a_col_hi, a_col_lo, use_a2) -> the 3-pass tap list of dsb_gemm_ex (lo*hi, hi*lo, hi*hi)."""
        out = []
        for j, (sh, ah, al, a2) in enumerate(spatial):
            wh, wl = j * 2 * self.Kp, j * 2 * self.Kp + self.Kp
            if self.fold:
                if al != ah + 32:
                    raise ValueError("folded packing: the lo half must follow the hi half at +32 columns")
                out += [(sh, ah, wh, a2), (sh, ah, wl, a2)]
            else:
                out += [(sh, al, wh, a2), (sh, ah, wl, a2), (sh, ah, wh, a2)]
        return out

    def taps64(self, spatial):
        """The same products as taps(), cut into 64-deep k-blocks (K = 64 per tap: the form dsb_gemm_ex's resident_w kernel takes); taps that
        read the same A box are adjacent so that the kernel stages it once."""
        out = []
        for sh, ac, wc, a2 in self.taps(spatial):
            out += [(sh, ac + 64 * i, wc + 64 * i, a2) for i in range(self.Kp // 64)]
        if self.Kp > 64:  # regroup: (al_i, wh_i), (ah_i, wl_i), (ah_i, wh_i) per 64-column slice i
            n = self.Kp // 64
            trip = [out[k:k + 3 * n] for k in range(0, len(out), 3 * n)]
            out = [tp[p * n + i] for tp in trip for i in range(n) for p in range(3)]
        return out

    def resident_ok(self, n_taps):
        """Do n_taps W boxes (N rounded to 16 rows x 128 bytes each) fit dsb_gemm_ex's resident-W kernel?"""
        return self.N <= 128 and n_taps <= 32 and n_taps * ((self.N + 15) // 16 * 16) * 128 <= 96 * 1024
