"""torch-tensor front end of the training-side C-ABI kernels (csrc/train.cu).

Same rules as ops.py: PyTorch supplies memory and streams only, every byte of arithmetic happens in libdiffsound_b200.so, no fallback.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib
from .ops import BF16, TF32, _need_cuda, _ptr, _stream


def act_code(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return TF32
    if t.dtype == torch.bfloat16:
        return BF16
    raise RuntimeError(f"training activations are fp32 (tf32) or bf16, got {t.dtype}")


def q_sample(x0, t, uniform, sched, T: int, out=None):
    """x_t ~ q(x_t | x_0) with the uniforms supplied (diffusion_transformer.py:370-377).  uniform: (B, K+1, L)."""
    _need_cuda(x0, t, uniform, sched)
    B, L = x0.shape
    K = uniform.shape[1] - 1
    out = torch.empty_like(x0) if out is None else out
    _lib.check(_lib.lib().dsb_q_sample(x0.data_ptr(), t.data_ptr(), uniform.data_ptr(), sched.data_ptr(), out.data_ptr(), B, K, L, T, _stream()), "dsb_q_sample")
    return out


def train_loss(logits, x0, x_t, t, pt, sched, T: int, *, aux_weight: float, adaptive: bool, mask_weight, dlogits=None, log_model_prob=None,
               hits=None, lt_history=None, lt_count=None, prob_as_exp: bool = False, bufs: Optional[dict] = None):
    """Fused _train_loss (:408-476).  logits (B, L, K) fp32.  Returns dict(loss (1,), vb_loss (B,), kl_loss (B,), col (B,L,2))."""
    _need_cuda(logits, x0, x_t, t, pt, sched)
    B, L, K = logits.shape
    dev = logits.device
    bufs = {} if bufs is None else bufs
    def buf(name, shape, dtype=torch.float32):
        v = bufs.get(name)
        if v is None or tuple(v.shape) != tuple(shape):
            v = bufs[name] = torch.empty(shape, dtype=dtype, device=dev)
        return v
    col, kl, vb, loss, scratch = buf("col", (B, L, 2)), buf("kl", (B,)), buf("vb", (B,)), buf("loss", (1,)), buf("scratch", (B,))
    _lib.check(_lib.lib().dsb_train_loss(logits.data_ptr(), x0.data_ptr(), x_t.data_ptr(), t.data_ptr(), pt.data_ptr(), sched.data_ptr(),
                                         _ptr(dlogits), _ptr(log_model_prob), col.data_ptr(), _ptr(hits), kl.data_ptr(), vb.data_ptr(),
                                         loss.data_ptr(), _ptr(lt_history), _ptr(lt_count), scratch.data_ptr(), B, K, L, T, float(aux_weight),
                                         1 if adaptive else 0, float(mask_weight[0]), float(mask_weight[1]), 1 if prob_as_exp else 0, _stream()), "dsb_train_loss")
    return dict(loss=loss, vb_loss=vb, kl_loss=kl, col=col)


def transpose(x: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
    """out[..., c, r] = x[..., r, c].  x: (rows, cols) or (batch, rows, cols) row-strided; out: (cols, >= rows) / (batch, cols, >= rows)."""
    _need_cuda(x, out)
    if x.stride(-1) != 1 or out.stride(-1) != 1 or x.element_size() != out.element_size():
        raise RuntimeError("transpose: operands must be unit-stride in the last dim and of the same element size")
    batched = x.dim() == 3
    rows, cols = x.shape[-2], x.shape[-1]
    if out.shape[-2] != cols or out.shape[-1] < rows:
        raise RuntimeError(f"transpose: out {tuple(out.shape)} cannot hold the transpose of {tuple(x.shape)}")
    _lib.check(_lib.lib().dsb_transpose(x.data_ptr(), x.stride(-2), x.stride(0) if batched else 0, out.data_ptr(), out.stride(-2),
                                        out.stride(0) if batched else 0, rows, cols, x.shape[0] if batched else 1, x.element_size(), _stream()),
               "dsb_transpose")
    return out


def heads_split(tok: torch.Tensor, heads: torch.Tensor, B: int, H: int, Lx: int) -> torch.Tensor:
    """tok: (B*Lx, >= H*64) row-strided view; heads: contiguous (B*H, Lx, 64)."""
    _need_cuda(tok, heads)
    _lib.check(_lib.lib().dsb_heads_split(tok.data_ptr(), tok.stride(0), heads.data_ptr(), B, H, Lx, tok.element_size(), _stream()), "dsb_heads_split")
    return heads


def heads_merge(heads: torch.Tensor, tok: torch.Tensor, B: int, H: int, Lx: int) -> torch.Tensor:
    _need_cuda(tok, heads)
    _lib.check(_lib.lib().dsb_heads_merge(heads.data_ptr(), tok.data_ptr(), tok.stride(0), B, H, Lx, tok.element_size(), _stream()), "dsb_heads_merge")
    return tok


def cast_scale(x: torch.Tensor, out: torch.Tensor, scale: Optional[torch.Tensor] = None) -> torch.Tensor:
    _need_cuda(x, out, scale)
    _lib.check(_lib.lib().dsb_cast_scale(x.data_ptr(), out.data_ptr(), x.numel(), _ptr(scale), act_code(out), _stream()), "dsb_cast_scale")
    return out


def colsum(x: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
    """out[n] = sum_r x[r, n]  (x row-strided 2-D, fp32 or bf16; out fp32, overwritten)."""
    _need_cuda(x, out)
    _lib.check(_lib.lib().dsb_colsum(x.data_ptr(), x.stride(0), out.data_ptr(), x.shape[0], x.shape[1], act_code(x), _stream()), "dsb_colsum")
    return out


def gelu2_fwd(u, a):
    _lib.check(_lib.lib().dsb_gelu2_fwd(u.data_ptr(), a.data_ptr(), u.numel(), act_code(u), _stream()), "dsb_gelu2_fwd")
    return a


def gelu2_bwd(u, da, du):
    _lib.check(_lib.lib().dsb_gelu2_bwd(u.data_ptr(), da.data_ptr(), du.data_ptr(), u.numel(), act_code(u), _stream()), "dsb_gelu2_bwd")
    return du


def silu_bwd(x, dy, dx):
    _lib.check(_lib.lib().dsb_silu_bwd(x.data_ptr(), dy.data_ptr(), dx.data_ptr(), x.numel(), _stream()), "dsb_silu_bwd")
    return dx


def gather_rows(table, idx, out):
    _lib.check(_lib.lib().dsb_gather_rows(table.data_ptr(), idx.data_ptr(), out.data_ptr(), idx.numel(), table.shape[1], _stream()), "dsb_gather_rows")
    return out


def scatter_add_rows(table, idx, src):
    _lib.check(_lib.lib().dsb_scatter_add_rows(table.data_ptr(), idx.data_ptr(), src.data_ptr(), idx.numel(), table.shape[1], _stream()), "dsb_scatter_add_rows")
    return table


def layernorm_bwd(x, dy, dx_io, gamma, dgamma, dbeta, eps=1e-5, dx_act=None):
    """dx_act (optional, fp32 or bf16, same shape): also receives the updated dx_io in the activation dtype."""
    D = x.shape[-1]
    _lib.check(_lib.lib().dsb_layernorm_bwd(x.data_ptr(), dy.data_ptr(), dx_io.data_ptr(), gamma.data_ptr(), dgamma.data_ptr(), dbeta.data_ptr(),
                                            x.numel() // D, D, eps, _ptr(dx_act), act_code(dx_act) if dx_act is not None else 0, _stream()),
               "dsb_layernorm_bwd")
    return dx_io


def ada_layernorm_bwd(x, dy, dx_io, table, idx, dtable, eps=1e-5, dx_act=None):
    B, L, D = x.shape
    _lib.check(_lib.lib().dsb_ada_layernorm_bwd(x.data_ptr(), dy.data_ptr(), dx_io.data_ptr(), table.data_ptr(), idx.data_ptr(), dtable.data_ptr(),
                                                B, L, D, eps, _ptr(dx_act), act_code(dx_act) if dx_act is not None else 0, _stream()),
               "dsb_ada_layernorm_bwd")
    return dx_io


def softmax_fwd(S, P, n: int):
    """S: (rows, ld_s) fp32, P: (rows, ld_p); softmax over the first n columns of each row."""
    _lib.check(_lib.lib().dsb_softmax_fwd(S.data_ptr(), S.stride(-2), P.data_ptr(), P.stride(-2), S.numel() // S.shape[-1], n, act_code(P), _stream()),
               "dsb_softmax_fwd")
    return P


def softmax_bwd(P, dP, dS, n: int, alpha: float):
    _lib.check(_lib.lib().dsb_softmax_bwd(P.data_ptr(), P.stride(-2), dP.data_ptr(), dP.stride(-2), dS.data_ptr(), dS.stride(-2),
                                          P.numel() // P.shape[-1], n, alpha, act_code(P), _stream()), "dsb_softmax_bwd")
    return dS


def embed_bwd(ids, dx, demb, dheight, dwidth):
    B, L = ids.shape
    D = dx.shape[-1]
    _lib.check(_lib.lib().dsb_embed_bwd(ids.data_ptr(), dx.data_ptr(), demb.data_ptr(), dheight.data_ptr(), dwidth.data_ptr(), B, L, D,
                                        dheight.shape[0], dwidth.shape[0], demb.shape[0], _stream()), "dsb_embed_bwd")


def attention_train_fwd(q, k, v, o, lse, B: int, H: int, Lq: int, Lk: int, scale: float):
    """Fused bf16 attention forward on token-major row-strided views (head h = columns [64h, 64h+64)); writes o and lse (B*H, Lq)."""
    _need_cuda(q, k, v, o, lse)
    _lib.check(_lib.lib().dsb_attention_train_fwd(q.data_ptr(), q.stride(0), k.data_ptr(), k.stride(0), v.data_ptr(), v.stride(0), o.data_ptr(), o.stride(0),
                                                  lse.data_ptr(), B, H, Lq, Lk, scale, _stream()), "dsb_attention_train_fwd")
    return o


def attention_train_bwd(q, k, v, o, dout, lse, delta, dq, dk, dv, B: int, H: int, Lq: int, Lk: int, scale: float):
    _need_cuda(q, k, v, o, dout, lse, delta, dq, dk, dv)
    _lib.check(_lib.lib().dsb_attention_train_bwd(q.data_ptr(), q.stride(0), k.data_ptr(), k.stride(0), v.data_ptr(), v.stride(0), o.data_ptr(), o.stride(0),
                                                  dout.data_ptr(), dout.stride(0), lse.data_ptr(), delta.data_ptr(), dq.data_ptr(), dq.stride(0),
                                                  dk.data_ptr(), dk.stride(0), dv.data_ptr(), dv.stride(0), B, H, Lq, Lk, scale, _stream()),
               "dsb_attention_train_bwd")
