"""CPU oracle for the Diffsound hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product package
(``text-to-sound-synthesis_b200/``) never imports anything under ``oracle/``.

This is a *restatement* of the reference algorithm in plain functional torch (CPU, fp32, with
fp64 exactly where the reference uses fp64).  It takes a ``state_dict`` with the reference's own
key names, so the same weights drive the oracle, the reference (when importable) and the CUDA
path.  Every function cites the reference file:line it follows; paths are relative to
``/root/reference/Diffsound``.

Parity pinning: the reference ships no tests / golden vectors, so this
oracle is pinned against *outputs of the reference itself*, produced in the build container by
``oracle/gen_golden.py`` (which imports the unmodified reference classes) and committed under
``tests/golden/``; ``tests/test_oracle_golden.py`` replays them.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor
LOG_1E30 = float(np.log(np.float32(1e-30)))  # index_to_log_onehot floor, diffusion_transformer.py:54


# --------------------------------------------------------------------------------------------
# A9: schedule  (sound_synthesis/modeling/transformers/diffusion_transformer.py:122-151, :193-231)
# --------------------------------------------------------------------------------------------
def alpha_schedule(time_step: int, N: int, att_1=0.99999, att_T=0.000009, ctt_1=0.000009, ctt_T=0.9):
    """fp64 numpy mask-and-uniform schedule (diffusion_transformer.py:122-151)."""
    att = np.arange(0, time_step) / (time_step - 1) * (att_T - att_1) + att_1
    att = np.concatenate(([1], att))
    at = att[1:] / att[:-1]
    ctt = np.arange(0, time_step) / (time_step - 1) * (ctt_T - ctt_1) + ctt_1
    ctt = np.concatenate(([0], ctt))
    one_minus_ctt = 1 - ctt
    one_minus_ct = one_minus_ctt[1:] / one_minus_ctt[:-1]
    ct = 1 - one_minus_ct
    bt = (1 - at - ct) / N
    att = np.concatenate((att[1:], [1]))
    ctt = np.concatenate((ctt[1:], [0]))
    btt = (1 - att - ctt) / N
    return at, bt, ct, att, btt, ctt


def schedule_buffers(num_timesteps: int, num_classes: int) -> Dict[str, Tensor]:
    """The eight fp32 log-space buffers registered at diffusion_transformer.py:224-231.

    ``num_classes`` is K+1 (the mask token included), as at :194.
    """
    at, bt, ct, att, btt, ctt = alpha_schedule(num_timesteps, N=num_classes)
    t64 = lambda a: torch.tensor(a.astype("float64"))
    log_at, log_bt, log_ct = torch.log(t64(at)), torch.log(t64(bt)), torch.log(t64(ct))
    log_cumprod_at, log_cumprod_bt, log_cumprod_ct = torch.log(t64(att)), torch.log(t64(btt)), torch.log(t64(ctt))
    log_1_min_a = lambda a: torch.log(1 - a.exp() + 1e-40)  # :25-26
    return {
        "log_at": log_at.float(), "log_bt": log_bt.float(), "log_ct": log_ct.float(),
        "log_cumprod_at": log_cumprod_at.float(), "log_cumprod_bt": log_cumprod_bt.float(),
        "log_cumprod_ct": log_cumprod_ct.float(),
        "log_1_min_ct": log_1_min_a(log_ct).float(),
        "log_1_min_cumprod_ct": log_1_min_a(log_cumprod_ct).float(),
    }


# --------------------------------------------------------------------------------------------
# A5 + A4: denoiser  (modeling/embeddings/dalle_mask_image_embedding.py:36-58,
#                     modeling/transformers/transformer_utils.py)
# --------------------------------------------------------------------------------------------
def content_embedding(sd: Dict[str, Tensor], prefix: str, index: Tensor, spatial: Tuple[int, int]) -> Tensor:
    """emb[x] + (height_emb[h] + width_emb[w]) over a row-major (H,W) grid (dalle_mask_image_embedding.py:36-58)."""
    H, W = spatial
    emb = F.embedding(index.clamp_min(0), sd[prefix + "emb.weight"])
    pos = (sd[prefix + "height_emb.weight"][:, None, :] + sd[prefix + "width_emb.weight"][None, :, :]).reshape(1, H * W, -1)
    return emb + pos[:, : emb.shape[1], :]


def ada_layer_norm(sd, prefix: str, x: Tensor, t: Tensor) -> Tensor:
    """AdaLayerNorm with a learned timestep table (transformer_utils.py:134-149, branch :140)."""
    e = F.linear(F.silu(F.embedding(t, sd[prefix + "emb.weight"])), sd[prefix + "linear.weight"], sd[prefix + "linear.bias"]).unsqueeze(1)
    scale, shift = torch.chunk(e, 2, dim=2)
    return F.layer_norm(x, (x.shape[-1],)) * (1 + scale) + shift


def attention(sd, prefix: str, x: Tensor, kv_src: Tensor, n_head: int) -> Tensor:
    """FullAttention / CrossAttention forward (transformer_utils.py:43-58, :91-109); dropout p=0."""
    B, T, C = x.shape
    Te = kv_src.shape[1]
    hs = C // n_head
    k = F.linear(kv_src, sd[prefix + "key.weight"], sd[prefix + "key.bias"]).view(B, Te, n_head, hs).transpose(1, 2)
    q = F.linear(x, sd[prefix + "query.weight"], sd[prefix + "query.bias"]).view(B, T, n_head, hs).transpose(1, 2)
    v = F.linear(kv_src, sd[prefix + "value.weight"], sd[prefix + "value.bias"]).view(B, Te, n_head, hs).transpose(1, 2)
    att = (q @ k.transpose(-2, -1)) * (1.0 / math.sqrt(hs))
    att = F.softmax(att, dim=-1)
    y = (att @ v).transpose(1, 2).contiguous().view(B, T, C)
    return F.linear(y, sd[prefix + "proj.weight"], sd[prefix + "proj.bias"])


def gelu2(x: Tensor) -> Tensor:
    return x * torch.sigmoid(1.702 * x)  # transformer_utils.py:111-115


def transformer_block(sd, prefix: str, x: Tensor, cond: Tensor, t: Tensor, n_head: int) -> Tensor:
    """Block.forward, 'selfcross' branch (transformer_utils.py:255-272)."""
    x = x + _self_attn(sd, prefix, x, t, n_head)
    x = x + attention(sd, prefix + "attn2.", ada_layer_norm(sd, prefix + "ln1_1.", x, t), cond, n_head)
    h = F.layer_norm(x, (x.shape[-1],), sd[prefix + "ln2.weight"], sd[prefix + "ln2.bias"])
    h = F.linear(gelu2(F.linear(h, sd[prefix + "mlp.0.weight"], sd[prefix + "mlp.0.bias"])), sd[prefix + "mlp.2.weight"], sd[prefix + "mlp.2.bias"])
    return x + h


def _self_attn(sd, prefix, x, t, n_head):
    h = ada_layer_norm(sd, prefix + "ln1.", x, t)
    return attention(sd, prefix + "attn1.", h, h, n_head)


def transformer_forward(sd: Dict[str, Tensor], x_t: Tensor, cond_emb: Tensor, t: Tensor, *, n_layer: int, n_head: int,
                        spatial: Tuple[int, int], prefix: str = "transformer.") -> Tensor:
    """Text2ImageTransformer.forward -> logits (B, K, L)  (transformer_utils.py:421-443)."""
    emb = content_embedding(sd, prefix + "content_emb.", x_t, spatial)
    for n in range(n_layer):
        emb = transformer_block(sd, f"{prefix}blocks.{n}.", emb, cond_emb, t, n_head)
    h = F.layer_norm(emb, (emb.shape[-1],), sd[prefix + "to_logits.0.weight"], sd[prefix + "to_logits.0.bias"])
    logits = F.linear(h, sd[prefix + "to_logits.1.weight"], sd[prefix + "to_logits.1.bias"])
    return logits.permute(0, 2, 1)  # 'b l c -> b c l' (:442)


# --------------------------------------------------------------------------------------------
# A3 tail, A6, A7, A8: the per-column posterior / sampler math
# --------------------------------------------------------------------------------------------
def predict_start_tail(out: Tensor) -> Tensor:
    """fp64 log_softmax over K, append the -70 mask row, clamp (diffusion_transformer.py:285-289)."""
    B, K, L = out.shape
    log_pred = F.log_softmax(out.double(), dim=1).float()
    log_pred = torch.cat((log_pred, torch.zeros(B, 1, L) - 70), dim=1)
    return torch.clamp(log_pred, -70, 0)


def nucleus_filter(lp: Tensor, r: float) -> Tensor:
    """'top{r}r' truncation (modeling/models/dalle_spec.py:158-174).  Note torch's CPU cumsum
    accumulates fp32 inputs in fp64 (ATen acc_type<float,false>) and rounds each prefix to fp32."""
    temp, indices = torch.sort(lp, 1, descending=True)
    temp2 = torch.exp(temp).cumsum(dim=1)
    temp3 = temp2 < r
    keep_sorted = torch.cat((torch.full_like(temp3[:, 0:1, :], True), temp3), dim=1)[:, :-1, :]
    keep = keep_sorted.gather(1, indices.argsort(1))
    return keep.float() * lp + (1 - keep.float()) * (-70)


def topk_filter(lp: Tensor, k: int) -> Tensor:
    """'top{k}p' truncation (dalle_spec.py:147-157)."""
    val, ind = lp.topk(k=k, dim=1)
    return torch.full_like(lp, -70).scatter_(1, ind, val)


def log_add_exp(a: Tensor, b: Tensor) -> Tensor:
    m = torch.max(a, b)  # diffusion_transformer.py:28-30
    return m + torch.log(torch.exp(a - m) + torch.exp(b - m))


def index_to_log_onehot(x: Tensor, num_classes: int) -> Tensor:
    oh = F.one_hot(x, num_classes).permute(0, 2, 1)  # diffusion_transformer.py:45-56
    return torch.log(oh.float().clamp(min=1e-30))


def q_pred(sched, log_x_start: Tensor, t: Tensor, T: int) -> Tensor:
    """q(x_t | x_0) in log space (diffusion_transformer.py:253-267); t wraps mod T+1."""
    t = (t + (T + 1)) % (T + 1)
    g = lambda name: sched[name].gather(-1, t).reshape(-1, 1, 1)
    out = torch.zeros_like(log_x_start)
    out[:, :-1, :] = log_add_exp(log_x_start[:, :-1, :] + g("log_cumprod_at"), g("log_cumprod_bt"))
    out[:, -1:, :] = log_add_exp(log_x_start[:, -1:, :] + g("log_1_min_cumprod_ct"), g("log_cumprod_ct"))
    return out


def q_pred_one_timestep(sched, log_x_t: Tensor, t: Tensor) -> Tensor:
    g = lambda name: sched[name].gather(-1, t).reshape(-1, 1, 1)  # :241-251
    out = torch.zeros_like(log_x_t)
    out[:, :-1, :] = log_add_exp(log_x_t[:, :-1, :] + g("log_at"), g("log_bt"))
    out[:, -1:, :] = log_add_exp(log_x_t[:, -1:, :] + g("log_1_min_ct"), g("log_ct"))
    return out


def q_posterior(sched, log_x_start: Tensor, log_x_t: Tensor, t: Tensor, T: int) -> Tensor:
    """log p_theta(x_{t-1} | x_t) closed form (diffusion_transformer.py:293-339)."""
    B, C, L = log_x_start.shape
    x_t = log_x_t.argmax(1)
    mask = (x_t == C - 1).unsqueeze(1)
    log_one = torch.zeros(B, 1, 1)
    log_zero = torch.log(log_one + 1.0e-30).expand(-1, -1, L)
    log_qt = q_pred(sched, log_x_t, t, T)
    log_qt = torch.cat((log_qt[:, :-1, :], log_zero), dim=1)
    ct_cum = sched["log_cumprod_ct"].gather(-1, t).reshape(-1, 1, 1).expand(-1, C - 1, -1)
    ct_cum = torch.cat((ct_cum, log_one), dim=1)
    log_qt = (~mask) * log_qt + mask * ct_cum
    log_q1 = q_pred_one_timestep(sched, log_x_t, t)
    log_q1 = torch.cat((log_q1[:, :-1, :], log_zero), dim=1)
    ct = sched["log_ct"].gather(-1, t).reshape(-1, 1, 1).expand(-1, C - 1, -1)
    ct = torch.cat((ct, log_one), dim=1)
    log_q1 = (~mask) * log_q1 + mask * ct
    q = log_x_start - log_qt
    lse = torch.logsumexp(q, dim=1, keepdim=True)
    q = q - lse
    out = q_pred(sched, q, t - 1, T) + log_q1 + lse
    return torch.clamp(out, -70, 0)


def gumbel_argmax(logits: Tensor, uniform: Tensor) -> Tensor:
    """log_sample_categorical with the uniforms supplied (diffusion_transformer.py:359-365); returns ids."""
    g = -torch.log(-torch.log(uniform + 1e-30) + 1e-30)
    return (g + logits).argmax(dim=1)


def posterior_sample_step(sched, out: Tensor, x_t: Tensor, t: Tensor, uniform: Tensor, *, T: int,
                          truncation: Optional[str] = "top0.85r", first_step_carrier: bool = False,
                          t_posterior: Optional[Tensor] = None):
    """One p_sample minus the denoiser: logits (B,K,L) + ids (B,L) + t + u (B,K+1,L) -> next ids.

    Returns (next_ids, model_log_prob, log_pred_filtered).  ``first_step_carrier`` reproduces the
    -inf carrier of the all-[MASK] start state (diffusion_transformer.py:633-636); it is
    numerically irrelevant because every position is masked there (checked in tests).
    ``t_posterior`` lets sample_fast use t-skip_step in q_posterior (:799-802).
    """
    C = out.shape[1] + 1
    lp = predict_start_tail(out)
    if truncation is not None:
        if truncation.endswith("r"):
            lp = nucleus_filter(lp, float(truncation[:-1].replace("top", "")))
        elif truncation.endswith("p"):
            lp = topk_filter(lp, int(truncation[:-1].replace("top", "")))
        else:
            raise ValueError(truncation)
    if first_step_carrier:
        log_x_t = torch.log(F.one_hot(x_t, C).permute(0, 2, 1).float())
    else:
        log_x_t = index_to_log_onehot(x_t, C)
    post = q_posterior(sched, lp, log_x_t, t if t_posterior is None else t_posterior, T)
    return gumbel_argmax(post, uniform), post, lp


# --------------------------------------------------------------------------------------------
# A13: training loss  (diffusion_transformer.py:370-377 q_sample, :408-476 _train_loss, :566-584 forward)
# --------------------------------------------------------------------------------------------
def multinomial_kl(log_prob1: Tensor, log_prob2: Tensor) -> Tensor:
    return (log_prob1.exp() * (log_prob1 - log_prob2)).sum(dim=1)  # diffusion_transformer.py:236-238


def q_sample_ids(sched, x_start: Tensor, t: Tensor, uniform: Tensor, *, T: int, num_classes: int) -> Tensor:
    """q_sample with the uniforms supplied (:370-377): x_t ~ q(x_t | x_0) by Gumbel argmax; returns ids (B, L)."""
    return gumbel_argmax(q_pred(sched, index_to_log_onehot(x_start, num_classes), t, T), uniform)


def train_loss_from_logits(sched, out: Tensor, x_start: Tensor, x_t: Tensor, t: Tensor, pt: Tensor, *, T: int = 100,
                            aux_weight: float = 0.0, adaptive_aux: bool = False, mask_weight=(1.0, 1.0), is_train: bool = True):
    """_train_loss after the denoiser call (:420-474): out = transformer logits (B, K, L).  Differentiable in ``out``."""
    B, L = x_start.shape
    C = out.shape[1] + 1
    log_x_start = index_to_log_onehot(x_start, C)
    log_xt = index_to_log_onehot(x_t, C)
    log_x0_recon = predict_start_tail(out)                                        # :269-291 (no truncation in training)
    log_model_prob = q_posterior(sched, log_x0_recon, log_xt, t, T)                 # :421
    log_true_prob = q_posterior(sched, log_x_start, log_xt, t, T)                   # :439
    kl = multinomial_kl(log_true_prob, log_model_prob)
    mask_region = (x_t == C - 1).float()
    mw = mask_region * mask_weight[0] + (1.0 - mask_region) * mask_weight[1]
    kl = (kl * mw).sum(-1)
    decoder_nll = -(log_x_start.exp() * log_model_prob).sum(dim=1).sum(-1)          # :24 log_categorical, :445-446
    is0 = (t == 0).float()
    kl_loss = is0 * decoder_nll + (1.0 - is0) * kl
    vb = kl_loss / pt
    kl_aux_loss = None
    if aux_weight != 0 and is_train:
        kl_aux = (multinomial_kl(log_x_start[:, :-1, :], log_x0_recon[:, :-1, :]) * mw).sum(-1)
        kl_aux_loss = is0 * decoder_nll + (1.0 - is0) * kl_aux
        w = (t / T + 1.0) if adaptive_aux else 1.0
        vb = vb + w * aux_weight * kl_aux_loss / pt
    return dict(loss=vb.sum() / (B * L), vb_loss=vb, kl_loss=kl_loss, kl_aux_loss=kl_aux_loss, log_model_prob=log_model_prob, x_t=x_t,
                x0_recon=log_x0_recon.argmax(1), xt_1_recon=log_model_prob.argmax(1))


def train_loss(sd, sched, x_start: Tensor, cond_emb: Tensor, t: Tensor, pt: Tensor, uniform: Tensor, *, n_layer: int, n_head: int,
               spatial, T: int = 100, aux_weight: float = 0.0, adaptive_aux: bool = False, mask_weight=(1.0, 1.0),
               is_train: bool = True, prefix: str = "transformer."):
    """DiffusionTransformer._train_loss + the normalisation of forward() (:408-476, :568-569) with (t, pt) = sample_time(...) and the
    q_sample uniforms supplied by the caller.  Differentiable in ``sd`` (torch autograd) -- that is how the gradient goldens and
    the GPU backward parity references are produced.  Returns a dict; 'loss' is the scalar forward() puts in out['loss']."""
    C = sd[prefix + "to_logits.1.weight"].shape[0] + 1
    x_t = q_sample_ids(sched, x_start, t, uniform, T=T, num_classes=C)
    out = transformer_forward(sd, x_t, cond_emb, t, n_layer=n_layer, n_head=n_head, spatial=spatial, prefix=prefix)
    return train_loss_from_logits(sched, out, x_start, x_t, t, pt, T=T, aux_weight=aux_weight, adaptive_aux=adaptive_aux,
                                  mask_weight=mask_weight, is_train=is_train)


def sample(sd, cond_emb: Tensor, uniforms, *, n_layer: int, n_head: int, spatial, num_timesteps: int = 100,
           truncation: Optional[str] = "top0.85r", steps: Optional[Sequence[int]] = None, x_init: Optional[Tensor] = None,
           return_trace: bool = False, post_steps: Optional[Sequence[int]] = None):
    """DiffusionTransformer.sample, filter_ratio=0 branch (diffusion_transformer.py:628-654).

    ``uniforms`` is either a callable step_index -> (B,K+1,L) tensor or a torch.Generator (then
    ``torch.rand`` is drawn per step exactly as rand_like would on CPU).
    """
    B = cond_emb.shape[0]
    L = spatial[0] * spatial[1]
    K = sd["transformer.to_logits.1.weight"].shape[0]
    sched = {k: sd[k] for k in ("log_at", "log_bt", "log_ct", "log_cumprod_at", "log_cumprod_bt", "log_cumprod_ct",
                                "log_1_min_ct", "log_1_min_cumprod_ct")}
    x = torch.full((B, L), K, dtype=torch.long) if x_init is None else x_init.clone()
    trace = []
    steps = list(range(num_timesteps - 1, -1, -1)) if steps is None else list(steps)
    for i, ti in enumerate(steps):
        t = torch.full((B,), ti, dtype=torch.long)
        out = transformer_forward(sd, x, cond_emb, t, n_layer=n_layer, n_head=n_head, spatial=spatial)
        u = uniforms(i) if callable(uniforms) else torch.rand((B, K + 1, L), generator=uniforms)
        tp = None if post_steps is None else torch.full((B,), post_steps[i], dtype=torch.long)   # sample_fast: q_posterior at t - skip_step (:799-802)
        x_new, post, lp = posterior_sample_step(sched, out, x, t, u, T=num_timesteps, truncation=truncation, t_posterior=tp)
        if return_trace:
            trace.append({"t": ti, "x_in": x.clone(), "logits": out.clone(), "x_out": x_new.clone()})
        x = x_new
    return (x, trace) if return_trace else x


# --------------------------------------------------------------------------------------------
# A10 + A11: ids -> mel   (dalle_spec.py:80-91, permuter.py:21-55, quantize.py:88-103,
#                          spec_codec/vqgan.py:62-65, specvqgan/modules/diffusionmodules/model.py)
# --------------------------------------------------------------------------------------------
def fast_schedule(num_timesteps: int, skip_step: int):
    """(denoiser steps, posterior steps) of sample_fast (diffusion_transformer.py:792-802)."""
    steps = list(range(num_timesteps - 1, -1, -1 - skip_step))
    if steps[-1] != 0:
        steps.append(0)
    return steps, [s - skip_step if s > skip_step else s for s in steps]


def column_major_reverse(ids: Tensor, H: int, W: int) -> Tensor:
    idx = torch.arange(H * W).reshape(H, W).T.reshape(-1)  # permuter.py:51-55
    return ids[:, torch.argsort(idx)]                      # :46-49 reverse=True


def codebook_lookup(sd, ids: Tensor, bhwc, prefix="content_codec.") -> Tensor:
    z = F.embedding(ids.reshape(-1), sd[prefix + "quantize.embedding.weight"])  # one-hot matmul == gather (quantize.py:91-95)
    return z.view(bhwc).permute(0, 3, 1, 2).contiguous()


def _gn(sd, p, x):
    return F.group_norm(x, 32, sd[p + "weight"], sd[p + "bias"], eps=1e-6)  # model.py:34-35


def _swish(x):
    return x * torch.sigmoid(x)  # model.py:29-31


def _conv(sd, p, x, pad):
    return F.conv2d(x, sd[p + "weight"], sd[p + "bias"], padding=pad)


def dec_resnet_block(sd, p, x):
    """ResnetBlock.forward with temb=None, dropout 0 (model.py:131-151)."""
    h = _conv(sd, p + "conv1.", _swish(_gn(sd, p + "norm1.", x)), 1)
    h = _conv(sd, p + "conv2.", _swish(_gn(sd, p + "norm2.", h)), 1)
    if (p + "nin_shortcut.weight") in sd:
        x = _conv(sd, p + "nin_shortcut.", x, 0)
    return x + h


def dec_attn_block(sd, p, x):
    """AttnBlock.forward (model.py:202-226): single head, scale c^-0.5."""
    h = _gn(sd, p + "norm.", x)
    q, k, v = _conv(sd, p + "q.", h, 0), _conv(sd, p + "k.", h, 0), _conv(sd, p + "v.", h, 0)
    b, c, hh, ww = q.shape
    w_ = torch.bmm(q.reshape(b, c, -1).permute(0, 2, 1), k.reshape(b, c, -1)) * (int(c) ** (-0.5))
    w_ = F.softmax(w_, dim=2)
    h = torch.bmm(v.reshape(b, c, -1), w_.permute(0, 2, 1)).reshape(b, c, hh, ww)
    return x + _conv(sd, p + "proj_out.", h, 0)


def decoder_forward(sd, z: Tensor, *, ch_mult=(1, 1, 2, 2, 4), num_res_blocks=2, prefix="content_codec.decoder.") -> Tensor:
    """Decoder.forward (model.py:640-671); attention wherever the state_dict has attn blocks."""
    p = prefix
    h = _conv(sd, p + "conv_in.", z, 1)
    h = dec_resnet_block(sd, p + "mid.block_1.", h)
    h = dec_attn_block(sd, p + "mid.attn_1.", h)
    h = dec_resnet_block(sd, p + "mid.block_2.", h)
    for lvl in reversed(range(len(ch_mult))):
        for blk in range(num_res_blocks + 1):
            h = dec_resnet_block(sd, f"{p}up.{lvl}.block.{blk}.", h)
            if f"{p}up.{lvl}.attn.{blk}.norm.weight" in sd:
                h = dec_attn_block(sd, f"{p}up.{lvl}.attn.{blk}.", h)
        if lvl != 0:
            h = F.interpolate(h, scale_factor=2.0, mode="nearest")  # model.py:48-52
            h = _conv(sd, f"{p}up.{lvl}.upsample.conv.", h, 1)
    return _conv(sd, p + "conv_out.", _swish(_gn(sd, p + "norm_out.", h)), 1)


# --------------------------------------------------------------------------------------------
# N4: SpecVQGAN encoder + nearest-codebook quantiser = DALLE.get_tokens (training-side tokeniser)
# --------------------------------------------------------------------------------------------
def encoder_forward(sd, x: Tensor, *, ch_mult=(1, 1, 2, 2, 4), num_res_blocks=2, prefix="content_codec.encoder.") -> Tensor:
    """Encoder.forward (specvqgan/modules/diffusionmodules/model.py:476-500); Downsample = zero pad (0,1,0,1) + 3x3 stride-2 conv (:55-75)."""
    p = prefix
    h = _conv(sd, p + "conv_in.", x, 1)
    for lvl in range(len(ch_mult)):
        for blk in range(num_res_blocks):
            h = dec_resnet_block(sd, f"{p}down.{lvl}.block.{blk}.", h)
            if f"{p}down.{lvl}.attn.{blk}.norm.weight" in sd:
                h = dec_attn_block(sd, f"{p}down.{lvl}.attn.{blk}.", h)
        if lvl != len(ch_mult) - 1:
            h = F.conv2d(F.pad(h, (0, 1, 0, 1)), sd[f"{p}down.{lvl}.downsample.conv.weight"], sd[f"{p}down.{lvl}.downsample.conv.bias"], stride=2)
    h = dec_resnet_block(sd, p + "mid.block_1.", h)
    h = dec_attn_block(sd, p + "mid.attn_1.", h)
    h = dec_resnet_block(sd, p + "mid.block_2.", h)
    return _conv(sd, p + "conv_out.", _swish(_gn(sd, p + "norm_out.", h)), 1)


def encode_to_tokens(sd, mel: Tensor, *, ch_mult=(1, 1, 2, 2, 4), num_res_blocks=2, prefix="content_codec."):
    """VQModel.encode (spec_codec/vqgan.py:48-54) + VectorQuantizer.forward's argmin (quantize.py:56-63) + DALLE.get_tokens' ColumnMajor
    permutation (dalle_spec.py:71-78, permuter.py:41-49).  Returns (z after quant_conv (B,E,H,W), token ids (B, H*W) in the transformer's order)."""
    h = encoder_forward(sd, mel, ch_mult=ch_mult, num_res_blocks=num_res_blocks, prefix=prefix + "encoder.")
    z = _conv(sd, prefix + "quant_conv.", h, 0)
    B, E, H, W = z.shape
    zf = z.permute(0, 2, 3, 1).reshape(-1, E)
    emb = sd[prefix + "quantize.embedding.weight"]
    d = torch.sum(zf ** 2, dim=1, keepdim=True) + torch.sum(emb ** 2, dim=1) - 2 * torch.matmul(zf, emb.t())
    idx = torch.argmin(d, dim=1).view(B, H * W)                         # row-major (h * W + w)
    col_major = torch.arange(H * W).reshape(H, W).t().reshape(-1)       # ColumnMajor.forward: x[:, idx]
    return z, idx[:, col_major]


def decode_to_img(sd, ids: Tensor, *, grid=(5, 53), embed_dim=256, ch_mult=(1, 1, 2, 2, 4), num_res_blocks=2,
                  prefix="content_codec.") -> Tensor:
    """DALLE.decode_to_img (dalle_spec.py:80-91): ids (B,L) -> mel (B,1,16H,16W)."""
    H, W = grid
    B = ids.shape[0]
    ids = column_major_reverse(ids, H, W)
    z = codebook_lookup(sd, ids, (B, H, W, embed_dim), prefix)
    z = _conv(sd, prefix + "post_quant_conv.", z, 0)  # vqgan.py:62-65
    return decoder_forward(sd, z, ch_mult=ch_mult, num_res_blocks=num_res_blocks, prefix=prefix + "decoder.")


# --------------------------------------------------------------------------------------------
# A12: MelGAN generator (vocoder/modules.py:72-130) on weight_g / weight_v checkpoints
# --------------------------------------------------------------------------------------------
def _wn(sd, p):
    v, g = sd[p + "weight_v"], sd[p + "weight_g"]  # torch.nn.utils.weight_norm, dim=0 (modules.py:18-23)
    return g * v / v.flatten(1).norm(dim=1).view(-1, 1, 1)


def melgan_forward(sd, mel: Tensor, *, ratios=(8, 8, 2, 2), n_residual_layers=3, prefix="model.") -> Tensor:
    """Generator.forward (modules.py:95-130): mel01 (B,80,T) -> wav (B,1,T*256)."""
    i = 0
    x = F.conv1d(F.pad(mel, (3, 3), mode="reflect"), _wn(sd, f"{prefix}1."), sd[f"{prefix}1.bias"])
    i = 2
    for r in ratios:
        x = F.leaky_relu(x, 0.2)
        i += 1
        x = F.conv_transpose1d(x, _wn(sd, f"{prefix}{i}."), sd[f"{prefix}{i}.bias"], stride=r,
                               padding=r // 2 + r % 2, output_padding=r % 2)
        i += 1
        for j in range(n_residual_layers):
            d = 3 ** j
            p = f"{prefix}{i}."
            h = F.pad(F.leaky_relu(x, 0.2), (d, d), mode="reflect")
            h = F.conv1d(h, _wn(sd, p + "block.2."), sd[p + "block.2.bias"], dilation=d)
            h = F.conv1d(F.leaky_relu(h, 0.2), _wn(sd, p + "block.4."), sd[p + "block.4.bias"])
            x = F.conv1d(x, _wn(sd, p + "shortcut."), sd[p + "shortcut.bias"]) + h
            i += 1
    x = F.pad(F.leaky_relu(x, 0.2), (3, 3), mode="reflect")
    i += 2
    return torch.tanh(F.conv1d(x, _wn(sd, f"{prefix}{i}."), sd[f"{prefix}{i}.bias"]))


# --------------------------------------------------------------------------------------------
# Deterministic synthetic weights with the reference's key names/shapes
# --------------------------------------------------------------------------------------------
# --------------------------------------------------------------------------------------------
# N2: CLIP text tower as used by Diffsound (CLIPTextEmbedding with pick_last_embedding=False, embed_dim=512, normalize=True)
# --------------------------------------------------------------------------------------------
def clip_text_forward(sd, tokens: Tensor, *, n_layer: int, n_head: int = 8, normalize: bool = True, prefix: str = "") -> Tensor:
    """CLIPTextEmbedding.forward for the Diffsound config (embeddings/clip_text_embedding.py:46-88 with modules/clip/model.py:166-199):
    token + positional embedding, n_layer pre-LN residual blocks (causal multi-head attention, QuickGELU MLP), ln_final, per-token L2 norm."""
    p = prefix
    tok = tokens.clamp_min(0)                                                       # text[text < 0] = 0
    x = F.embedding(tok, sd[p + "token_embedding.weight"]) + sd[p + "positional_embedding"]
    B, L, D = x.shape
    hs = D // n_head
    mask = torch.full((L, L), float("-inf")).triu_(1)
    for i in range(n_layer):
        b = f"{p}transformer.resblocks.{i}."
        h = F.layer_norm(x, (D,), sd[b + "ln_1.weight"], sd[b + "ln_1.bias"])
        qkv = F.linear(h, sd[b + "attn.in_proj_weight"], sd[b + "attn.in_proj_bias"])
        q, k, v = [t.view(B, L, n_head, hs).transpose(1, 2) for t in qkv.chunk(3, dim=-1)]
        att = F.softmax((q @ k.transpose(-2, -1)) * (1.0 / math.sqrt(hs)) + mask, dim=-1)
        y = (att @ v).transpose(1, 2).reshape(B, L, D)
        x = x + F.linear(y, sd[b + "attn.out_proj.weight"], sd[b + "attn.out_proj.bias"])
        h = F.layer_norm(x, (D,), sd[b + "ln_2.weight"], sd[b + "ln_2.bias"])
        x = x + F.linear(gelu2(F.linear(h, sd[b + "mlp.c_fc.weight"], sd[b + "mlp.c_fc.bias"])), sd[b + "mlp.c_proj.weight"], sd[b + "mlp.c_proj.bias"])
    x = F.layer_norm(x, (D,), sd[p + "ln_final.weight"], sd[p + "ln_final.bias"])
    return x / x.norm(dim=-1, keepdim=True) if normalize else x


def make_clip_text_state_dict(*, n_layer=12, width=512, vocab=49408, ctx=77, seed=0) -> Dict[str, Tensor]:
    """Random text-tower weights under CLIP's parameter names (CLIP.initialize_parameters scales, modules/clip/model.py:300-321)."""
    g = torch.Generator().manual_seed(seed)
    n = lambda *s, std: torch.randn(*s, generator=g) * std
    sd = {"token_embedding.weight": n(vocab, width, std=0.02), "positional_embedding": n(ctx, width, std=0.01),
          "ln_final.weight": 1 + n(width, std=0.05), "ln_final.bias": n(width, std=0.05), "text_projection": n(width, width, std=width ** -0.5)}
    proj_std, attn_std, fc_std = (width ** -0.5) * ((2 * n_layer) ** -0.5), width ** -0.5, (2 * width) ** -0.5
    for i in range(n_layer):
        b = f"transformer.resblocks.{i}."
        sd[b + "attn.in_proj_weight"], sd[b + "attn.in_proj_bias"] = n(3 * width, width, std=attn_std), n(3 * width, std=0.02)
        sd[b + "attn.out_proj.weight"], sd[b + "attn.out_proj.bias"] = n(width, width, std=proj_std), n(width, std=0.02)
        sd[b + "ln_1.weight"], sd[b + "ln_1.bias"] = 1 + n(width, std=0.05), n(width, std=0.05)
        sd[b + "ln_2.weight"], sd[b + "ln_2.bias"] = 1 + n(width, std=0.05), n(width, std=0.05)
        sd[b + "mlp.c_fc.weight"], sd[b + "mlp.c_fc.bias"] = n(4 * width, width, std=fc_std), n(4 * width, std=0.02)
        sd[b + "mlp.c_proj.weight"], sd[b + "mlp.c_proj.bias"] = n(width, 4 * width, std=proj_std), n(width, std=0.02)
    return sd


def make_transformer_state_dict(*, K=256, D=1024, n_layer=19, n_head=16, spatial=(5, 53), cond_dim=512, T=100,
                                mlp_times=4, seed=0, std=0.02) -> Dict[str, Tensor]:
    """Random DiffusionTransformer state_dict: N(0,std) Linear/Embedding weights, zero biases, unit
    LayerNorm (the distribution of transformer_utils.py:355-363; not the reference's RNG stream --
    the same dict feeds both the oracle and the CUDA path)."""
    g = torch.Generator().manual_seed(seed)
    n = lambda *s: torch.randn(*s, generator=g) * std
    sd = dict(schedule_buffers(T, K + 1))
    sd["Lt_history"], sd["Lt_count"] = torch.zeros(T), torch.zeros(T)
    p = "transformer."
    sd[p + "content_emb.emb.weight"] = n(K + 1, D)
    sd[p + "content_emb.height_emb.weight"] = n(spatial[0], D)
    sd[p + "content_emb.width_emb.weight"] = n(spatial[1], D)
    L = spatial[0] * spatial[1]
    for i in range(n_layer):
        b = f"{p}blocks.{i}."
        for ln in ("ln1.", "ln1_1."):
            sd[b + ln + "emb.weight"] = n(T, D)
            sd[b + ln + "linear.weight"] = n(2 * D, D)
            sd[b + ln + "linear.bias"] = n(2 * D) * 0.5
        sd[b + "ln2.weight"], sd[b + "ln2.bias"] = 1 + n(D), n(D)
        for a, kd in (("attn1.", D), ("attn2.", cond_dim)):
            sd[b + a + "key.weight"], sd[b + a + "key.bias"] = n(D, kd), n(D)
            sd[b + a + "value.weight"], sd[b + a + "value.bias"] = n(D, kd), n(D)
            sd[b + a + "query.weight"], sd[b + a + "query.bias"] = n(D, D), n(D)
            sd[b + a + "proj.weight"], sd[b + a + "proj.bias"] = n(D, D), n(D)
        sd[b + "attn2.mask"] = torch.tril(torch.ones(L, L)).view(1, 1, L, L)
        sd[b + "mlp.0.weight"], sd[b + "mlp.0.bias"] = n(mlp_times * D, D), n(mlp_times * D)
        sd[b + "mlp.2.weight"], sd[b + "mlp.2.bias"] = n(D, mlp_times * D), n(D)
    sd[p + "to_logits.0.weight"], sd[p + "to_logits.0.bias"] = 1 + n(D), n(D)
    sd[p + "to_logits.1.weight"], sd[p + "to_logits.1.bias"] = n(K, D), n(K)
    return sd


def make_decoder_state_dict(*, n_embed=256, embed_dim=256, z_channels=256, ch=128, ch_mult=(1, 1, 2, 2, 4),
                            num_res_blocks=2, attn_levels=(4,), out_ch=1, seed=0, prefix="content_codec.") -> Dict[str, Tensor]:
    """Random VQModel decoder-side state_dict (quantize.embedding, post_quant_conv, decoder.*)."""
    g = torch.Generator().manual_seed(seed)

    def conv(cin, cout, k):
        bound = 1.0 / math.sqrt(cin * k * k)
        w = (torch.rand(cout, cin, k, k, generator=g) * 2 - 1) * bound
        b = (torch.rand(cout, generator=g) * 2 - 1) * bound
        return w, b

    sd: Dict[str, Tensor] = {}

    def put_conv(name, cin, cout, k):
        sd[name + "weight"], sd[name + "bias"] = conv(cin, cout, k)

    def put_gn(name, c):
        sd[name + "weight"] = 1 + 0.1 * torch.randn(c, generator=g)
        sd[name + "bias"] = 0.1 * torch.randn(c, generator=g)

    def put_res(name, cin, cout):
        put_gn(name + "norm1.", cin); put_conv(name + "conv1.", cin, cout, 3)
        put_gn(name + "norm2.", cout); put_conv(name + "conv2.", cout, cout, 3)
        if cin != cout:
            put_conv(name + "nin_shortcut.", cin, cout, 1)

    def put_attn(name, c):
        put_gn(name + "norm.", c)
        for n_ in ("q.", "k.", "v.", "proj_out."):
            put_conv(name + n_, c, c, 1)

    sd[prefix + "quantize.embedding.weight"] = (torch.rand(n_embed, embed_dim, generator=g) * 2 - 1) / n_embed * 64
    put_conv(prefix + "post_quant_conv.", embed_dim, z_channels, 1)
    p = prefix + "decoder."
    block_in = ch * ch_mult[-1]
    put_conv(p + "conv_in.", z_channels, block_in, 3)
    put_res(p + "mid.block_1.", block_in, block_in); put_attn(p + "mid.attn_1.", block_in); put_res(p + "mid.block_2.", block_in, block_in)
    for lvl in reversed(range(len(ch_mult))):
        block_out = ch * ch_mult[lvl]
        for blk in range(num_res_blocks + 1):
            put_res(f"{p}up.{lvl}.block.{blk}.", block_in, block_out)
            block_in = block_out
            if lvl in attn_levels:
                put_attn(f"{p}up.{lvl}.attn.{blk}.", block_in)
        if lvl != 0:
            put_conv(f"{p}up.{lvl}.upsample.conv.", block_in, block_in, 3)
    put_gn(p + "norm_out.", block_in)
    put_conv(p + "conv_out.", block_in, out_ch, 3)
    return sd


def make_melgan_state_dict(*, input_size=80, ngf=32, n_residual_layers=3, ratios=(8, 8, 2, 2), seed=0) -> Dict[str, Tensor]:
    """Random Generator state_dict in weight_g/weight_v form (126 keys for the shipped config)."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, Tensor] = {}

    def put(name, shape):
        v = torch.randn(*shape, generator=g) * 0.05
        sd[name + "bias"] = torch.randn(shape[1] if name.endswith("T.") else shape[0], generator=g) * 0.02
        sd[name + "weight_g"] = v.flatten(1).norm(dim=1).view(-1, 1, 1) * (1 + 0.1 * torch.randn(shape[0], 1, 1, generator=g))
        sd[name + "weight_v"] = v

    mult = 2 ** len(ratios)
    put("model.1.", (mult * ngf, input_size, 7))
    i = 2
    for r in ratios:
        i += 1
        cin, cout = mult * ngf, mult * ngf // 2
        v = torch.randn(cin, cout, 2 * r, generator=g) * 0.05  # ConvTranspose1d weight is (Cin, Cout, k); weight_norm dim=0
        sd[f"model.{i}.bias"] = torch.randn(cout, generator=g) * 0.02
        sd[f"model.{i}.weight_g"] = v.flatten(1).norm(dim=1).view(-1, 1, 1)
        sd[f"model.{i}.weight_v"] = v
        i += 1
        for _ in range(n_residual_layers):
            put(f"model.{i}.block.2.", (cout, cout, 3))
            put(f"model.{i}.block.4.", (cout, cout, 1))
            put(f"model.{i}.shortcut.", (cout, cout, 1))
            i += 1
        mult //= 2
    i += 2
    put(f"model.{i}.", (1, ngf, 7))
    return sd
