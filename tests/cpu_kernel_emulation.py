"""CPU stand-ins for the C-ABI kernels, used ONLY by tests/test_cpu_train_engine.py to exercise the ORCHESTRATION of DenoiserTrainEngine (which
buffer feeds which launch, in what order, with which strides / flags) without a GPU.  Each function restates the documented contract of the
kernel it replaces (include/diffsound_b200.h) in plain torch; numerics are fp32 (storage dtype of the output buffer is honoured).
TEST INFRASTRUCTURE -- never imported by the product package."""
import math

import torch
import torch.nn.functional as F

TF32, BF16, F16 = 0, 1, 2


def _f(x):
    return x.float()


# ------------------------------------------------------------------------------------------------ ops.*
def gemm(a, w, bias=None, residual=None, out=None, *, dtype=TF32, gelu=False, round_out=False, alpha=1.0, a_mn=False, w_mn=False, **kw):
    A = _f(a).transpose(-1, -2) if a_mn else _f(a)          # (.., M, K)
    W = _f(w) if w_mn else _f(w).transpose(-1, -2)           # (.., K, N)
    y = alpha * (A @ W)
    if bias is not None:
        y = y + bias
    if gelu:
        y = y * torch.sigmoid(1.702 * y)
    if residual is not None:
        y = y + residual
    if out is None:
        return y
    out[..., : y.shape[-1]].copy_(y) if out.shape[-1] != y.shape[-1] else out.copy_(y)
    return out


def gemm_f32(a, w, bias=None, residual=None, out=None, *, gelu=False, round_out=False):
    return gemm(a, w, bias, residual, out, gelu=gelu)


def silu(x, out=None):
    y = F.silu(x)
    return y if out is None else out.copy_(y)


def embed_tokens(ids, emb, height_emb, width_emb, out=None, err_flag=None):
    B, L = ids.shape
    W = width_emb.shape[0]
    l = torch.arange(L)
    y = emb[ids.clamp(min=0)] + (height_emb[l // W] + width_emb[l % W])
    return y if out is None else out.copy_(y)


def layernorm(x, gamma, beta, out=None, *, eps=1e-5, **kw):
    y = F.layer_norm(x, (x.shape[-1],), gamma, beta, eps)
    return y if out is None else out.copy_(y)


def ada_layernorm(x, table, t, out=None, *, eps=1e-5, **kw):
    D = x.shape[-1]
    sel = table[t]
    y = F.layer_norm(x, (D,), None, None, eps) * (1 + sel[:, None, :D]) + sel[:, None, D:]
    return y if out is None else out.copy_(y)


# ------------------------------------------------------------------------------------------------ train_ops.*
def cast_scale(x, out, scale=None):
    return out.copy_((x * (scale if scale is not None else 1.0)).reshape(out.shape))


def transpose(x, out):
    r = x.shape[-2]
    out[..., :r].copy_(x.transpose(-1, -2))
    return out


def heads_split(tok, heads, B, H, Lx):
    return heads.copy_(tok[:, : H * 64].reshape(B, Lx, H, 64).permute(0, 2, 1, 3).reshape(B * H, Lx, 64))


def heads_merge(heads, tok, B, H, Lx):
    tok[:, : H * 64].copy_(heads.reshape(B, H, Lx, 64).permute(0, 2, 1, 3).reshape(B * Lx, H * 64))
    return tok


def colsum(x, out):
    return out.copy_(_f(x).sum(0))


def gelu2_fwd(u, a):
    x = _f(u)
    return a.copy_(x * torch.sigmoid(1.702 * x))


def gelu2_bwd(u, da, du):
    x = _f(u)
    s = torch.sigmoid(1.702 * x)
    return du.copy_(_f(da) * (s + 1.702 * x * s * (1 - s)))


def silu_bwd(x, dy, dx):
    s = torch.sigmoid(x)
    return dx.copy_(dy * (s + x * s * (1 - s)))


def gather_rows(table, idx, out):
    return out.copy_(table[idx])


def scatter_add_rows(table, idx, src):
    return table.index_add_(0, idx, src)


def _ln_bwd(x, dy, g, eps):
    D = x.shape[-1]
    mu = x.mean(-1, keepdim=True)
    rstd = torch.rsqrt(x.var(-1, unbiased=False, keepdim=True) + eps)
    xhat = (x - mu) * rstd
    dxh = dy * g
    dx = rstd * (dxh - dxh.mean(-1, keepdim=True) - xhat * (dxh * xhat).mean(-1, keepdim=True))
    return dx, xhat


def layernorm_bwd(x, dy, dx_io, gamma, dgamma, dbeta, eps=1e-5, dx_act=None):
    dx, xhat = _ln_bwd(x, dy.reshape(x.shape), gamma, eps)
    dx_io.add_(dx)
    dgamma.add_((dy.reshape(x.shape) * xhat).reshape(-1, x.shape[-1]).sum(0))
    dbeta.add_(dy.reshape(-1, x.shape[-1]).sum(0))
    if dx_act is not None:
        dx_act.copy_(dx_io.reshape(dx_act.shape))
    return dx_io


def ada_layernorm_bwd(x, dy, dx_io, table, idx, dtable, eps=1e-5, dx_act=None):
    B, L, D = x.shape
    dy = dy.reshape(x.shape)
    g = 1 + table[idx][:, None, :D]
    dx, xhat = _ln_bwd(x, dy, g, eps)
    dx_io.add_(dx)
    dtable.index_add_(0, idx, torch.cat(((dy * xhat).sum(1), dy.sum(1)), dim=1))
    if dx_act is not None:
        dx_act.copy_(dx_io.reshape(dx_act.shape))
    return dx_io


def softmax_fwd(S, P, n):
    P[..., :n].copy_(torch.softmax(S[..., :n], -1))
    return P


def softmax_bwd(P, dP, dS, n, alpha):
    p = _f(P[..., :n])
    d = dP[..., :n]
    dS[..., :n].copy_(alpha * p * (d - (d * p).sum(-1, keepdim=True)))
    return dS


def embed_bwd(ids, dx, demb, dheight, dwidth):
    B, L = ids.shape
    D = dx.shape[-1]
    W = dwidth.shape[0]
    demb.index_add_(0, ids.reshape(-1).clamp(min=0), dx.reshape(-1, D))
    l = torch.arange(L)
    per_pos = dx.reshape(B, L, D).sum(0)
    dheight.index_add_(0, l // W, per_pos)
    dwidth.index_add_(0, l % W, per_pos)


def _heads(x, B, H, L):
    return _f(x[:, : H * 64]).reshape(B, L, H, 64).permute(0, 2, 1, 3)


def _unheads(x, B, H, L):
    return x.permute(0, 2, 1, 3).reshape(B * L, H * 64)


def attention_train_fwd(q, k, v, o, lse, B, H, Lq, Lk, scale):
    Q, K, V = _heads(q, B, H, Lq), _heads(k, B, H, Lk), _heads(v, B, H, Lk)
    s = (Q @ K.transpose(-1, -2)) * scale
    o[:, : H * 64].copy_(_unheads(torch.softmax(s, -1) @ V, B, H, Lq))
    lse.copy_((torch.logsumexp(s, -1) * 1.4426950408889634).reshape(B * H, Lq))
    return o


def attention_train_bwd(q, k, v, o, dout, lse, delta, dq, dk, dv, B, H, Lq, Lk, scale):
    Q, K, V, dO = _heads(q, B, H, Lq), _heads(k, B, H, Lk), _heads(v, B, H, Lk), _heads(dout, B, H, Lq)
    s = (Q @ K.transpose(-1, -2)) * scale
    P = torch.exp2(s * 1.4426950408889634 - lse.reshape(B, H, Lq, 1))          # rebuilt from the saved log-sum-exp, as the kernel does
    dV = P.transpose(-1, -2) @ dO
    dP = dO @ V.transpose(-1, -2)
    dl = (dO * _heads(o, B, H, Lq)).sum(-1, keepdim=True)
    delta.copy_(dl.reshape(B * H, Lq))
    dS = scale * P * (dP - dl)
    dq[:, : H * 64].copy_(_unheads(dS @ K, B, H, Lq))
    dk[:, : H * 64].copy_(_unheads(dS.transpose(-1, -2) @ Q, B, H, Lk))
    dv[:, : H * 64].copy_(_unheads(dV, B, H, Lk))


# ------------------------------------------------------------------------------------------------ inference-side stand-ins
def to_f16(x):
    return x.float().clone()      # storage stays fp32 on the CPU: the orchestration is under test, not the rounding


def round_tf32(x, out=None):
    return x.clone() if out is None else out.copy_(x)


def attention(q, k, v, out, *, B, H, Lq, Lk, scale, round_out=False, causal=False):
    Q, K, V = _heads(q, B, H, Lq), _heads(k, B, H, Lk), _heads(v, B, H, Lk)
    s = (Q @ K.transpose(-1, -2)) * scale
    if causal:
        s = s + torch.full((Lq, Lk), float("-inf")).triu_(1)
    out[:, : H * 64].copy_(_unheads(torch.softmax(s, -1) @ V, B, H, Lq))
    return out


def attention_tc(q, k, v, out, *, B, H, Lq, Lk, scale, pipelined=True):
    return attention(q, k, v, out, B=B, H=H, Lq=Lq, Lk=Lk, scale=scale)


def l2_normalize_rows_(x):
    return x.div_(x.norm(dim=-1, keepdim=True))
