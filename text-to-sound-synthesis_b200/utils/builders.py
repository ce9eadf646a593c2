"""Model builders shared by bench.py, __graft_entry__.smoke(), tools/ and the tests: reference-style `{target, params}` configs
(the layout of Diffsound/configs/caps.yaml:2-87) instantiated through this package's drop-in classes.  Weights are the modules'
own seeded random initialisation (the reference ships no Diffsound / SpecVQGAN checkpoint)."""
from __future__ import annotations

import os

import torch

from .misc import instantiate_from_config, retarget_config

DDCONFIG = dict(double_z=False, z_channels=256, resolution=848, in_channels=1, out_ch=1, ch=128, ch_mult=[1, 1, 2, 2, 4], num_res_blocks=2,
                attn_resolutions=[53], dropout=0.0)  # Diffsound/configs/caps.yaml:13-23


def diffusion_config(K, D, NL, NH, CD, spatial=(5, 53), T=100, precision=None, train_precision=None):
    """`diffusion_config` block of caps.yaml (reference class paths; retarget_config maps them to the drop-ins)."""
    tp = dict(attn_type="selfcross", n_layer=NL, condition_seq_len=77, content_seq_len=spatial[0] * spatial[1], content_spatial_size=list(spatial),
              n_embd=D, condition_dim=CD, n_head=NH, attn_pdrop=0.0, resid_pdrop=0.0, block_activate="GELU2", timestep_type="adalayernorm",
              mlp_hidden_times=4)
    if precision is not None:
        tp["precision"] = precision
    if train_precision is not None:
        tp["train_precision"] = train_precision
    return {"target": "sound_synthesis.modeling.transformers.diffusion_transformer.DiffusionTransformer", "params": {
        "diffusion_step": T, "alpha_init_type": "alpha1", "auxiliary_loss_weight": 5.0e-4, "adaptive_auxiliary_loss": True, "mask_weight": [1, 1],
        "condition_emb_config": None,
        "transformer_config": {"target": "sound_synthesis.modeling.transformers.transformer_utils.Text2ImageTransformer", "params": tp},
        "content_emb_config": {"target": "sound_synthesis.modeling.embeddings.dalle_mask_image_embedding.DalleMaskImageEmbedding", "params": dict(
            num_embed=K, spatial_size=tuple(spatial), embed_dim=D, trainable=True, pos_emb_type="embedding")}}}


def dalle_config(K, D, NL, NH, CD, precision=None, train_precision=None, ddconfig=None):
    """`model` block of caps.yaml with synthetic-embedding conditioning (condition_codec_config None, the reference's own bypass)."""
    return {"target": "sound_synthesis.modeling.models.dalle_spec.DALLE", "params": {
        "content_info": {"key": "image"}, "condition_info": {"key": "text"},
        "content_codec_config": {"target": "sound_synthesis.modeling.codecs.spec_codec.vqgan.VQModel", "params": {
            "ckpt_path": None, "embed_dim": 256, "n_embed": K, "lossconfig": {"target": "specvqgan.modules.losses.DummyLoss"},
            "ddconfig": dict(ddconfig or DDCONFIG)}},
        "condition_codec_config": None,
        "first_stage_permuter_config": {"target": "specvqgan.modules.transformer.permuter.ColumnMajor", "params": {"H": 5, "W": 53}},
        "diffusion_config": diffusion_config(K, D, NL, NH, CD, precision=precision, train_precision=train_precision)}}


def build_diffusion_transformer(K, D, NL, NH, CD, sd=None, spatial=(5, 53), T=100, precision=None):
    """DiffusionTransformer drop-in on cuda:current, eval mode; `sd` (reference key names) is loaded with strict=False."""
    m = instantiate_from_config(retarget_config(diffusion_config(K, D, NL, NH, CD, spatial=spatial, T=T, precision=precision)))
    if sd is not None:
        missing, unexpected = m.load_state_dict(sd, strict=False)
        assert not unexpected, unexpected
        assert all("attn2.mask" in k for k in missing), missing  # reference goldens drop the dead causal-mask buffers
    return m.cuda().eval()


def build_dalle(K=256, D=1024, NL=19, NH=16, CD=512, precision=None, seed=0):
    """The full caps.yaml model (DALLE = SpecVQGAN codec + ColumnMajor + DiffusionTransformer), seeded random init, on the GPU."""
    torch.manual_seed(seed)
    return instantiate_from_config(retarget_config(dalle_config(K, D, NL, NH, CD, precision=precision))).cuda().eval()


def build_vocoder(ckpt=None):
    """MelGAN Generator(80, 32, 3) (vocoder/modules.py:88-130) with the reference's shipped weights when `ckpt` exists."""
    from ..vocoder.modules import Generator
    voc = Generator(80, 32, 3)
    if ckpt and os.path.exists(ckpt):
        voc.load_state_dict(torch.load(ckpt, map_location="cpu"), strict=True)
    return voc.cuda().eval()
