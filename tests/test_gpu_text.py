"""N2 on the B200: the CLIP text tower (CLIPTextEmbedding drop-in -> TextTowerEngine) against the reference-generated golden and the oracle."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import diffsound_oracle as O  # noqa: E402
from tests.helpers import load_golden, rel_err  # noqa: E402


def _build(NL, V, sd, **kw):
    from tests import gpu_common  # noqa: F401
    from diffsound_b200.modeling.embeddings.clip_text_embedding import CLIPTextEmbedding
    m = CLIPTextEmbedding(num_embed=V, text_layers=NL, **kw)
    m.load_state_dict(sd, strict=True)
    return m.cuda()


def test_causal_attention_flag_matches_masked_softmax():
    from tests import gpu_common as G
    B, H, L = 2, 8, 77
    g = torch.Generator().manual_seed(0)
    qkv = torch.randn(B * L, 3 * H * 64, generator=g).half().cuda()
    D = H * 64
    out = torch.empty(B * L, D, dtype=torch.float16, device="cuda")
    G.ops.attention(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], out, B=B, H=H, Lq=L, Lk=L, scale=0.125, causal=True)
    hd = lambda x: x.float().reshape(B, L, H, 64).permute(0, 2, 1, 3)
    q, k, v = hd(qkv[:, :D]), hd(qkv[:, D:2 * D]), hd(qkv[:, 2 * D:])
    mask = torch.full((L, L), float("-inf"), device="cuda").triu_(1)
    ref = (torch.softmax(q @ k.transpose(-1, -2) * 0.125 + mask, -1) @ v).permute(0, 2, 1, 3).reshape(B * L, D)
    assert G.relerr(out.float(), ref) < 2e-3
    assert G.relerr(out.float()[0::L], v.permute(0, 2, 1, 3).reshape(B * L, D)[0::L]) < 2e-3  # position 0 sees only itself


def test_clip_text_tower_matches_reference_golden():
    """3 layers, seeded weights: the reference's CLIPTextEmbedding.forward (Diffsound flags) vs the CUDA tower; incl. a negative pad id."""
    _, g = load_golden("clip_text.npz")
    NL, V, seed = [int(v) for v in g["__cfg"]]
    m = _build(NL, V, O.make_clip_text_state_dict(n_layer=NL, vocab=V, seed=seed), pick_last_embedding=False, embed_dim=512, normalize=True)
    tok = torch.from_numpy(g["in_tokens"]).cuda()
    out = m(tok).cpu()
    ref = torch.from_numpy(g["out_features"])
    err = rel_err(out, ref)
    print("clip text tower (3 layers) rel err", err, "launches", m.engine.launches)
    assert out.shape == (3, 77, 512) and err < 3e-3
    assert float((out.norm(dim=-1) - 1).abs().max()) < 1e-5
    assert int(tok[2, 40]) == -100  # the caller's tensor is not modified


def test_clip_text_tower_full_depth_and_other_modes_match_oracle():
    """12 layers / 49408-token vocabulary (the ViT-B/32 text shape) vs the oracle; pick_last_embedding and the 1024-wide duplicate mode."""
    NL, V = 12, 49408
    sd = O.make_clip_text_state_dict(n_layer=NL, vocab=V, seed=1)
    gen = torch.Generator().manual_seed(2)
    tok = torch.zeros(4, 77, dtype=torch.long)
    for i, n in enumerate((5, 20, 77, 41)):
        tok[i, :n] = torch.randint(1, V - 2, (n,), generator=gen)
        tok[i, n - 1] = V - 1  # <|endoftext|> is the largest id: pick_last_embedding finds it by argmax
    ref = O.clip_text_forward(sd, tok, n_layer=NL)
    m = _build(NL, V, sd, pick_last_embedding=False, embed_dim=512, normalize=True)
    err = rel_err(m(tok.cuda()).cpu(), ref)
    print("clip text tower (12 layers) rel err", err)
    assert err < 5e-3
    m2 = _build(NL, V, sd, pick_last_embedding=False, embed_dim=1024, normalize=True)
    wide = m2(tok.cuda()).cpu()
    assert wide.shape == (4, 77, 1024) and rel_err(wide, torch.cat((ref, ref), 2) / (2 ** 0.5)) < 5e-3
    m3 = _build(NL, V, sd, pick_last_embedding=True, embed_dim=512, normalize=True)
    last = m3(tok.cuda()).cpu()
    x = O.clip_text_forward(sd, tok, n_layer=NL, normalize=False)
    ref_last = x[torch.arange(4), tok.argmax(-1)] @ sd["text_projection"]
    ref_last = ref_last / ref_last.norm(dim=-1, keepdim=True)
    assert last.shape == (4, 512) and rel_err(last, ref_last) < 5e-3


def test_captions_condition_the_sampler_through_the_reference_config_path():
    """condition_emb_config -> CLIPTextEmbedding inside DiffusionTransformer: sample(condition_token=ids) runs the text tower, then the fused loop."""
    from tests.test_gpu_transformer import build_dt  # noqa: F401
    from tests import gpu_common  # noqa: F401
    from diffsound_b200.modeling.transformers.diffusion_transformer import DiffusionTransformer
    K, D, NL, NH = 64, 128, 2, 2
    cfg = dict(
        content_emb_config=dict(target="diffsound_b200.modeling.embeddings.dalle_mask_image_embedding.DalleMaskImageEmbedding",
                                params=dict(num_embed=K, spatial_size=(5, 53), embed_dim=D, trainable=True, pos_emb_type="embedding")),
        condition_emb_config=dict(target="diffsound_b200.modeling.embeddings.clip_text_embedding.CLIPTextEmbedding",
                                  params=dict(clip_name="ViT-B/32", num_embed=3000, normalize=True, pick_last_embedding=False, keep_seq_len_dim=False,
                                              additional_last_embedding=False, embed_dim=512, text_layers=2)),
        transformer_config=dict(target="diffsound_b200.modeling.transformers.transformer_utils.Text2ImageTransformer",
                                params=dict(attn_type="selfcross", n_layer=NL, condition_seq_len=77, content_seq_len=265, content_spatial_size=[5, 53],
                                            n_embd=D, condition_dim=512, n_head=NH, attn_pdrop=0.0, resid_pdrop=0.0, block_activate="GELU2",
                                            timestep_type="adalayernorm", mlp_hidden_times=4)),
        diffusion_step=100, alpha_init_type="alpha1", auxiliary_loss_weight=5.0e-4, adaptive_auxiliary_loss=True, mask_weight=[1, 1])
    m = DiffusionTransformer(**cfg).cuda().eval()
    assert not any(p.requires_grad for p in m.condition_emb.parameters())
    tok = torch.randint(1, 3000, (2, 77)).cuda()
    m.truncation = "top0.85r"
    torch.manual_seed(3)
    a = m.sample(condition_token=tok, condition_mask=None, condition_embed=None, filter_ratio=0)["content_token"]
    emb = m.condition_emb(tok)
    tower, m.condition_emb = m.condition_emb, None      # without a condition_emb the module takes pre-computed embeddings (:623-627)
    torch.manual_seed(3)
    b = m.sample(condition_token=None, condition_mask=None, condition_embed=emb, filter_ratio=0, batch_size=2)["content_token"]
    m.condition_emb = tower
    assert a.shape == (2, 265) and torch.equal(a, b)
