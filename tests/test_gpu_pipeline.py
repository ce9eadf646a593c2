"""End-to-end through the drop-in DALLE + Generator: tokens -> mel -> wav on the GPU, each stage against the CPU oracle."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import diffsound_oracle as O  # noqa: E402
from tests.helpers import ROOT, rel_err  # noqa: E402


def _config(K, D, NL, NH, CD):
    return {"target": "sound_synthesis.modeling.models.dalle_spec.DALLE", "params": {
        "content_info": {"key": "image"}, "condition_info": {"key": "text"},
        "content_codec_config": {"target": "sound_synthesis.modeling.codecs.spec_codec.vqgan.VQModel", "params": {
            "ckpt_path": None, "embed_dim": 256, "n_embed": K, "lossconfig": {"target": "specvqgan.modules.losses.DummyLoss"},
            "ddconfig": dict(double_z=False, z_channels=256, resolution=848, in_channels=1, out_ch=1, ch=128, ch_mult=[1, 1, 2, 2, 4],
                             num_res_blocks=2, attn_resolutions=[53], dropout=0.0)}},
        "condition_codec_config": None,
        "first_stage_permuter_config": {"target": "specvqgan.modules.transformer.permuter.ColumnMajor", "params": {"H": 5, "W": 53}},
        "diffusion_config": {"target": "sound_synthesis.modeling.transformers.diffusion_transformer.DiffusionTransformer", "params": {
            "diffusion_step": 100, "alpha_init_type": "alpha1", "auxiliary_loss_weight": 5.0e-4, "adaptive_auxiliary_loss": True, "mask_weight": [1, 1],
            "condition_emb_config": None,
            "transformer_config": {"target": "sound_synthesis.modeling.transformers.transformer_utils.Text2ImageTransformer", "params": dict(
                attn_type="selfcross", n_layer=NL, condition_seq_len=77, content_seq_len=265, content_spatial_size=[5, 53], n_embd=D, condition_dim=CD,
                n_head=NH, attn_pdrop=0.0, resid_pdrop=0.0, block_activate="GELU2", timestep_type="adalayernorm", mlp_hidden_times=4)},
            "content_emb_config": {"target": "sound_synthesis.modeling.embeddings.dalle_mask_image_embedding.DalleMaskImageEmbedding", "params": dict(
                num_embed=K, spatial_size=(5, 53), embed_dim=D, trainable=True, pos_emb_type="embedding")}}}}}


def test_text_to_wav_pipeline_matches_oracle_stagewise():
    """Reference-style YAML config -> retarget -> DALLE.generate_content -> Generator.  The GPU's own tokens are pushed through the
    oracle decoder and vocoder: mel and waveform within 1e-3 relative (north_star tolerance), tokens are valid codebook ids, and the
    whole call is deterministic under a fixed seed."""
    import _pkg
    _pkg.load()
    from diffsound_b200.utils.misc import instantiate_from_config, retarget_config
    from diffsound_b200.vocoder.modules import Generator
    from diffsound_b200 import pipeline
    K, D, NL, NH, CD, B = 256, 128, 2, 2, 512, 2
    torch.manual_seed(0)
    dalle = instantiate_from_config(retarget_config(_config(K, D, NL, NH, CD)))
    dsd = O.make_decoder_state_dict(seed=4)
    dalle.load_state_dict(dsd, strict=False)
    dalle = dalle.cuda().eval()
    ck = os.path.join(ROOT, "oracle", "_ref", "best_netG.pt")
    vsd = torch.load(ck, map_location="cpu") if os.path.exists(ck) else O.make_melgan_state_dict(seed=1)
    voc = Generator(80, 32, 3)
    voc.load_state_dict(vsd, strict=True)
    voc = voc.cuda().eval()
    g = torch.Generator().manual_seed(3)
    cond = torch.randn(B, 77, CD, generator=g)
    cond = (cond / cond.norm(dim=-1, keepdim=True)).cuda()
    out = pipeline.synthesize(dalle, voc, cond, sample_type="top0.85r", seed=1234)
    tok, mel, wav = out["tokens"].cpu(), out["mel"].cpu(), out["wav"].cpu()
    assert tok.shape == (B, 265) and tok.dtype == torch.int64 and int(tok.min()) >= 0 and int(tok.max()) < K
    assert mel.shape == (B, 1, 80, 848) and wav.shape == (B, 1, 217088)
    ref_mel = O.decode_to_img(dsd, tok)
    e_mel = rel_err(mel, ref_mel)
    ref_wav = O.melgan_forward(vsd, (mel[:, 0] + 1) / 2)
    e_wav = rel_err(wav, ref_wav)
    print("pipeline mel rel err", e_mel, "mel MSE", float(((mel - ref_mel) ** 2).mean()), "wav rel err", e_wav)
    assert e_mel < 1e-3 and e_wav < 1e-3
    again = pipeline.synthesize(dalle, voc, cond, sample_type="top0.85r", seed=1234)
    assert torch.equal(again["tokens"].cpu(), tok) and torch.equal(again["wav"].cpu(), wav)
    # the reference's skip-step sampler ('top0.85r,fast3': 25 denoiser calls) through the same entry point
    fast = dalle.generate_content(batch={"condition_embed": cond}, filter_ratio=0, sample_type="top0.85r,fast3")
    assert fast["content"].shape == (B, 1, 80, 848) and int(fast["content_token"].max()) < K


def test_dalle_training_step_and_q_resampling_through_reference_config():
    """Solver.step's call (solver_spec.py:286-320) on the drop-in DALLE built from a reference-style config: model(batch=..., return_loss=True)
    -> loss.backward() -> optimizer over model.parameters(name='transformer'); plus the 'q' sample type (repeated p_sample at the same t)."""
    import random
    import _pkg
    _pkg.load()
    from diffsound_b200.utils.misc import instantiate_from_config, retarget_config
    K, D, NL, NH, CD, B = 64, 128, 2, 2, 64, 3
    torch.manual_seed(0)
    dalle = instantiate_from_config(retarget_config(_config(K, D, NL, NH, CD))).cuda().train()
    assert not dalle.content_codec.training  # the codec's .train is disabled, as in the reference (:17-20, :48)
    groups = dalle.parameters(name="transformer")
    opt = torch.optim.AdamW(groups, lr=1e-3, betas=(0.9, 0.96))
    g = torch.Generator().manual_seed(1)
    batch = {"content_token": torch.randint(0, K, (B, 265), generator=g), "condition_embed": torch.randn(B, 77, CD, generator=g)}
    losses = []
    for _ in range(6):
        torch.manual_seed(5)
        out = dalle(batch=batch, name="transformer", return_loss=True, step=0)
        opt.zero_grad()
        out["loss"].backward()
        opt.step()
        losses.append(float(out["loss"].detach()))
    assert losses[-1] < losses[0], losses
    # the reference's real training batch: mels under content_info['key'], tokenised on the fly by the frozen SpecVQGAN encoder (get_tokens)
    mel_batch = {"image": torch.rand(2, 1, 80, 848, generator=g) * 2 - 1, "condition_embed": torch.randn(2, 77, CD, generator=g)}
    quant_z, toks = dalle.get_tokens(mel_batch["image"].cuda())
    assert quant_z.shape == (2, 256, 5, 53) and toks.shape == (2, 265) and int(toks.max()) < K and dalle.zshape == quant_z.shape
    assert dalle.decode_to_img(toks, quant_z.shape).shape == (2, 1, 80, 848)          # tokens round-trip through the decoder entry
    out = dalle(batch=mel_batch, return_loss=True)
    out["loss"].backward()
    assert torch.isfinite(out["loss"]) and out["logits"].shape == (2, K + 1, 265)
    stale = dalle(batch=batch, return_loss=True)["loss"]
    dalle(batch=batch, return_loss=True)
    with pytest.raises(RuntimeError, match="overwritten by a later forward"):
        stale.backward()
    # 'q' re-sampling: with rate 1 every step runs twice -> 200 denoiser calls; same seeds -> same tokens as calling it again
    random.seed(3); torch.manual_seed(9)
    a = dalle.generate_content(batch=batch, filter_ratio=0, sample_type="top0.85r,q1.0")["content_token"]
    assert dalle.transformer.resample_rate == 1.0 and dalle.transformer.last_gpu_launches == 200 * (dalle.transformer.transformer.engine.launches_per_forward + 1)
    random.seed(3); torch.manual_seed(9)
    b = dalle.generate_content(batch=batch, filter_ratio=0, sample_type="top0.85r,q1.0")["content_token"]
    assert torch.equal(a, b) and int(a.max()) < K
    dalle.generate_content(batch=batch, filter_ratio=0, sample_type="top0.85r")
    assert dalle.transformer.resample_rate == 0.0
