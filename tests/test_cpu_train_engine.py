"""DenoiserTrainEngine's ORCHESTRATION on the CPU: the C-ABI kernels are replaced by plain-torch stand-ins (tests/cpu_kernel_emulation.py) so that
the launch sequence itself -- which buffer feeds which launch, the gradient layout, the fused-QKV / all-layer-KV bookkeeping, MN-major operand
choices, the fused-attention vs composed-attention branches -- is checked against torch autograd through the oracle without a GPU.
(The kernels themselves are checked on the B200 by tests/test_gpu_train.py.)"""
import pytest
import torch

from oracle import diffsound_oracle as O
from tests import cpu_kernel_emulation as emu


@pytest.fixture
def engine_env(monkeypatch):
    import _pkg
    _pkg.load()
    from diffsound_b200 import ops, train_ops
    for name in ("gemm", "gemm_f32", "silu", "embed_tokens", "layernorm", "ada_layernorm"):
        monkeypatch.setattr(ops, name, getattr(emu, name))
    for name in ("cast_scale", "transpose", "heads_split", "heads_merge", "colsum", "gelu2_fwd", "gelu2_bwd", "silu_bwd", "gather_rows", "scatter_add_rows",
                 "layernorm_bwd", "ada_layernorm_bwd", "softmax_fwd", "softmax_bwd", "embed_bwd", "attention_train_fwd", "attention_train_bwd"):
        monkeypatch.setattr(train_ops, name, getattr(emu, name))
    return ops


def _model(K, D, NL, NH, CD, sd):
    from diffsound_b200.modeling.transformers.diffusion_transformer import DiffusionTransformer
    m = DiffusionTransformer(
        content_emb_config=dict(target="diffsound_b200.modeling.embeddings.dalle_mask_image_embedding.DalleMaskImageEmbedding",
                                params=dict(num_embed=K, spatial_size=(5, 53), embed_dim=D, trainable=True, pos_emb_type="embedding")),
        condition_emb_config=None,
        transformer_config=dict(target="diffsound_b200.modeling.transformers.transformer_utils.Text2ImageTransformer",
                                params=dict(attn_type="selfcross", n_layer=NL, condition_seq_len=77, content_seq_len=265, content_spatial_size=[5, 53],
                                            n_embd=D, condition_dim=CD, n_head=NH, attn_pdrop=0.0, resid_pdrop=0.0, block_activate="GELU2",
                                            timestep_type="adalayernorm", mlp_hidden_times=4)),
        diffusion_step=100, alpha_init_type="alpha1", auxiliary_loss_weight=5.0e-4, adaptive_auxiliary_loss=True, mask_weight=[1, 1])
    m.load_state_dict(sd, strict=False)
    return m


@pytest.mark.parametrize("mode", ["composed_transposes", "fused_attention_mn_major"])
def test_engine_orchestration_matches_oracle_autograd(engine_env, mode):
    K, D, NL, NH, CD, B, L = 32, 128, 2, 2, 64, 2, 265
    sd = O.make_transformer_state_dict(K=K, D=D, n_layer=NL, n_head=NH, cond_dim=CD, seed=4)
    m = _model(K, D, NL, NH, CD, sd)
    eng = m.transformer.train_engine
    eng.__init__(m.transformer, precision="tf32" if mode == "composed_transposes" else "bf16")
    eng.adt = torch.float32            # exact storage: what is under test is the launch sequence, not the rounding of the kernels
    eng._kernel_device, eng.use_cuda_graph = "cpu", False
    g = torch.Generator().manual_seed(1)
    cond = torch.nn.functional.normalize(torch.randn(B, 77, CD, generator=g), dim=-1)
    x0 = torch.randint(0, K, (B, L), generator=g)
    t, pt = torch.tensor([40, 0]), torch.tensor([0.01, 0.02])
    u = torch.rand(B, K + 1, L, generator=g)
    names = [n for n in sd if n.startswith("transformer.") and "attn2.mask" not in n]
    leaf = {k: (v.clone().requires_grad_(True) if k in names else v) for k, v in sd.items()}
    sched = {k: sd[k] for k in sd if k.startswith("log_")}
    x_t = O.q_sample_ids(sched, x0, t, u, T=100, num_classes=K + 1)
    out_ref = O.transformer_forward(leaf, x_t, cond, t, n_layer=NL, n_head=NH, spatial=(5, 53))
    logits = eng.forward(x_t, cond, t)                                             # (B, L, K)
    assert torch.allclose(logits.permute(0, 2, 1), out_ref.detach(), rtol=1e-4, atol=1e-5)
    lg = logits.detach().permute(0, 2, 1).clone().requires_grad_(True)
    O.train_loss_from_logits(sched, lg, x0, x_t, t, pt, T=100, aux_weight=5e-4, adaptive_aux=True)["loss"].backward()
    grads = eng.backward(lg.grad.permute(0, 2, 1).contiguous(), scale=torch.tensor([2.0]))   # upstream d loss = 2 (e.g. a GradScaler)
    ref = O.train_loss_from_logits(sched, out_ref, x0, x_t, t, pt, T=100, aux_weight=5e-4, adaptive_aux=True)
    (2.0 * ref["loss"]).backward()
    gmax = max(float(leaf["transformer." + n].grad.abs().max()) for n in grads)
    for n, gr in grads.items():
        r = leaf["transformer." + n].grad
        assert gr.shape == r.shape, n
        err = float((gr - r).abs().max()) / max(float(r.abs().max()), 1e-4 * gmax)
        assert err < 2e-3, (mode, n, err)
    assert set(grads) == {n for n, _ in m.transformer.named_parameters()}
    # the segmented form (one autograd node per segment, so DDP can all-reduce a layer's gradients while earlier layers still run) produces the same
    # numbers as the monolithic pass, covers every parameter exactly once, and returns buffers the engine does not overwrite afterwards
    eng.backward_begin(lg.grad.permute(0, 2, 1).contiguous(), scale=torch.tensor([2.0]))
    seg_grads, order = {}, ["head"] + [("layer", li) for li in range(NL - 1, -1, -1)] + ["tail"]
    for seg in order:
        part = eng.backward_segment(seg)
        assert set(part) == set(eng.segment_names(seg)) and not (set(part) & set(seg_grads)), seg
        seg_grads.update(part)
    assert set(seg_grads) == set(grads)
    for n in grads:
        assert torch.equal(seg_grads[n], grads[n]), n


def test_inference_engine_and_text_tower_orchestration(engine_env, monkeypatch):
    """DenoiserEngine.forward (hoisted AdaLN tables, fused QKV, all-layer cross K/V GEMM, in-place residual GEMMs) and TextTowerEngine.forward
    (embedding-as-grid trick, causal attention, final L2 norm) with the kernel stand-ins, against the oracle."""
    ops = engine_env
    for name in ("to_f16", "round_tf32", "attention", "attention_tc", "l2_normalize_rows_"):
        monkeypatch.setattr(ops, name, getattr(emu, name))
    K, D, NL, NH, CD, B, L = 32, 128, 2, 2, 64, 2, 265
    sd = O.make_transformer_state_dict(K=K, D=D, n_layer=NL, n_head=NH, cond_dim=CD, seed=8)
    m = _model(K, D, NL, NH, CD, sd)
    eng = m.transformer.engine
    monkeypatch.setattr(type(eng), "device", property(lambda self: torch.device("cuda")), raising=False)   # repack()'s device gate
    g = torch.Generator().manual_seed(2)
    cond = torch.randn(B, 77, CD, generator=g)
    x_t = torch.randint(0, K + 1, (B, L), generator=g)
    t = torch.tensor([99, 3])
    import functools
    real_empty, real_zeros = torch.empty, torch.zeros
    strip = lambda fn: functools.wraps(fn)(lambda *a, **k: fn(*a, **{kk: vv for kk, vv in k.items() if kk != "device"}))
    monkeypatch.setattr(torch, "empty", strip(real_empty))
    monkeypatch.setattr(torch, "zeros", strip(real_zeros))
    for prec in ("f16", "tf32", "fp32"):
        eng.__init__(m.transformer, precision=prec)
        kv = eng.encode_condition(cond)
        logits = eng.forward(x_t, kv.float(), t, 77)
        ref = O.transformer_forward(sd, x_t, cond, t, n_layer=NL, n_head=NH, spatial=(5, 53))
        err = float((logits.permute(0, 2, 1) - ref).abs().max() / ref.abs().max())
        assert err < (3e-3 if prec == "f16" else 5e-5), (prec, err)   # 'f16' keeps real fp16 activation buffers (storage rounding), the others are exact here
        assert eng.launches_per_forward == 2 + 11 * NL + 1
    # CLIP text tower
    from diffsound_b200.modeling.embeddings.clip_text_embedding import CLIPTextEmbedding
    tsd = O.make_clip_text_state_dict(n_layer=2, vocab=500, seed=3)
    clip = CLIPTextEmbedding(num_embed=500, text_layers=2, pick_last_embedding=False, embed_dim=512)
    clip.load_state_dict(tsd, strict=True)
    teng = clip.engine
    tok = torch.randint(1, 500, (2, 77), generator=g)
    tok[1, 50:] = 0
    monkeypatch.setattr(type(clip.token_embedding.weight), "device", property(lambda self: torch.device("cuda")), raising=False)
    out = teng.forward(tok)
    ref = O.clip_text_forward(tsd, tok, n_layer=2)
    assert float((out - ref).abs().max() / ref.abs().max()) < 3e-3   # real fp16 activation buffers on the way (storage rounding only)
