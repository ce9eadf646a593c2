"""Import the UNMODIFIED reference (``/root/reference/Diffsound``) on CPU -- test infrastructure.

Used only in the build container (``/root/reference`` does not exist on the GPU box) by
``oracle/gen_golden.py``, by ``tests/test_oracle_vs_reference.py`` (skipped when the tree is absent)
and by ``bench.py --impl reference`` when available.  Nothing here is copied from the reference; it
only installs the three import shims SURVEY.md section 8(c) lists and builds the reference's own classes
from the reference's own YAML.
"""
from __future__ import annotations

import copy
import os
import sys
import types

import torch

REF_ROOT = os.environ.get("DIFFSOUND_REFERENCE", "/root/reference")
REF_DIFFSOUND = os.path.join(REF_ROOT, "Diffsound")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_DIFFSOUND, "sound_synthesis"))


def install_shims() -> None:
    """(1) Tensor.cuda -> identity (transformer_utils.py:434 calls t.cuda()); (2) pytorch_lightning
    stub (spec_codec/vqgan.py:3,11); (3) librosa stub (vocoder/modules.py:4)."""
    if getattr(install_shims, "_done", False):
        return
    torch.Tensor.cuda = lambda self, *a, **k: self
    pl = types.ModuleType("pytorch_lightning")
    pl.LightningModule = torch.nn.Module
    sys.modules.setdefault("pytorch_lightning", pl)
    lib = types.ModuleType("librosa")
    filt = types.ModuleType("librosa.filters")
    filt.mel = None
    lib.filters = filt
    sys.modules.setdefault("librosa", lib)
    sys.modules.setdefault("librosa.filters", filt)
    if REF_DIFFSOUND not in sys.path:
        sys.path.insert(0, REF_DIFFSOUND)
    install_shims._done = True


def load_config(name: str = "evaluation/caps_text.yaml") -> dict:
    import yaml
    with open(os.path.join(REF_DIFFSOUND, name)) as f:
        return yaml.full_load(f)


def build_dalle(*, K: int = 256, overrides: dict | None = None, seed: int = 0):
    """Reference DALLE with no checkpoints / CLIP; ``overrides`` patches
    transformer params (n_layer, n_embd, n_head, ...) and decoder ddconfig for small golden cases."""
    install_shims()
    from sound_synthesis.utils.misc import instantiate_from_config
    cfg = copy.deepcopy(load_config())
    mp = cfg["model"]["params"]
    mp["content_codec_config"]["params"]["ckpt_path"] = None
    mp["condition_codec_config"] = None
    dp = mp["diffusion_config"]["params"]
    dp["condition_emb_config"] = None
    tp = dp["transformer_config"]["params"]
    ce = dp["content_emb_config"]["params"]
    ov = dict(overrides or {})
    for k in ("n_layer", "n_embd", "n_head", "condition_dim", "mlp_hidden_times"):
        if k in ov:
            tp[k] = ov[k]
    if "n_embd" in ov:
        ce["embed_dim"] = ov["n_embd"]
    if K != 256:
        ce["num_embed"] = K
        mp["content_codec_config"]["params"]["n_embed"] = K
    dd = mp["content_codec_config"]["params"]["ddconfig"]
    for k in ("ch", "ch_mult", "num_res_blocks", "attn_resolutions", "z_channels", "resolution"):
        if ("dec_" + k) in ov:
            dd[k] = ov["dec_" + k]
    if "embed_dim" in ov:
        mp["content_codec_config"]["params"]["embed_dim"] = ov["embed_dim"]
    if "grid" in ov:
        H, W = ov["grid"]
        tp["content_seq_len"] = H * W
        tp["content_spatial_size"] = [H, W]
        ce["spatial_size"] = (H, W)
        mp["first_stage_permuter_config"]["params"].update(H=H, W=W)
    torch.manual_seed(seed)
    model = instantiate_from_config(cfg["model"]).eval()
    for p in model.parameters():
        p.requires_grad_(False)
    return model, cfg


def build_vocoder(real_weights: bool = True, seed: int = 0):
    install_shims()
    from vocoder.modules import Generator
    torch.manual_seed(seed)
    g = Generator(80, 32, 3)
    if real_weights:
        sd = torch.load(os.path.join(REF_DIFFSOUND, "vocoder/logs/vggsound/best_netG.pt"), map_location="cpu")
        g.load_state_dict(sd)
    return g.eval()
