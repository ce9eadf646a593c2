"""SpecVQGAN decoder + MelGAN vocoder parity on the B200 (reference-generated goldens + the CPU oracle)."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import diffsound_oracle as O  # noqa: E402
from tests.helpers import ROOT, load_golden, rel_err  # noqa: E402


@pytest.fixture(scope="module")
def G():
    from tests import gpu_common
    return gpu_common


def build_vq(K, E, ch, ch_mult, sd=None, z_channels=None, prefix="content_codec.", precision="tf32x3"):
    from diffsound_b200.modeling.codecs.spec_codec.vqgan import VQModel
    dd = dict(double_z=False, z_channels=z_channels or E, resolution=848, in_channels=1, out_ch=1, ch=ch, ch_mult=list(ch_mult), num_res_blocks=2,
              attn_resolutions=[53], dropout=0.0)
    m = VQModel(dd, None, n_embed=K, embed_dim=E, precision=precision)
    if sd is not None:
        m.load_state_dict({k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}, strict=True)
    return m.cuda().eval()


@pytest.mark.parametrize("precision,tol", [("tf32x3", 1e-3), ("tf32", 1.5e-2)])
def test_decoder_tiny_matches_reference_golden(G, precision, tol):
    """north_star tolerance (1e-3 relative) holds in the default split-TF32 mode; single-pass TF32 is the fast, looser mode."""
    sd, g = load_golden("decoder_tiny.npz")
    K, E, ch, H, W = [int(v) for v in g["__cfg"]]
    m = build_vq(K, E, ch, (1, 1, 1, 1, 2), sd, precision=precision)
    mel = m.decode_tokens(torch.from_numpy(g["in_ids"]).long().cuda(), (H, W)).cpu()
    ref = torch.from_numpy(g["out_mel"])
    assert mel.shape == ref.shape
    err = rel_err(mel, ref)
    print(f"decoder tiny [{precision}] rel err", err)
    assert err < tol
    # the reference-shaped entry point (NCHW latents) agrees with the token fast path
    ids_rm = O.column_major_reverse(torch.from_numpy(g["in_ids"]).long(), H, W)
    z = O.codebook_lookup(sd, ids_rm, (ids_rm.shape[0], H, W, E))
    mel2 = m.decode(z.cuda()).cpu()
    assert rel_err(mel2, mel) < 1e-5


def test_decoder_full_config_matches_oracle(G):
    """The shipped ddconfig (ch=128, ch_mult 1,1,2,2,4, 256-d codebook) on the 5x53 grid, B=1: mel (1,1,80,848)."""
    sd = O.make_decoder_state_dict(seed=2)
    m = build_vq(256, 256, 128, (1, 1, 2, 2, 4), sd)
    ids = torch.randint(0, 256, (1, 265), generator=torch.Generator().manual_seed(9))
    ref = O.decode_to_img(sd, ids)
    mel = m.decode_tokens(ids.cuda(), (5, 53)).cpu()
    assert mel.shape == (1, 1, 80, 848)
    err = rel_err(mel, ref)
    mse = float(((mel - ref) ** 2).mean())
    print("decoder full [tf32x3] rel err", err, "mel MSE", mse, "ref rms", float(ref.pow(2).mean().sqrt()))
    assert err < 1e-3


def test_melgan_tiny_matches_reference_golden(G):
    from diffsound_b200.vocoder.modules import Generator
    sd, g = load_golden("melgan_tiny.npz")
    m = Generator(80, 4, 3)
    m.load_state_dict(sd, strict=True)
    m = m.cuda().eval()
    wav = m(torch.from_numpy(g["in_mel"]).cuda()).cpu()
    ref = torch.from_numpy(g["out_wav"])
    assert wav.shape == ref.shape
    err = rel_err(wav, ref)
    print("melgan tiny [tf32x3] rel err", err)
    assert err < 1e-3


def test_melgan_real_checkpoint(G):
    """The reference's shipped generator weights (staged in oracle/_ref by build()): reference-generated I/O golden, then a
    full 848-frame clip against the oracle."""
    from diffsound_b200.vocoder.modules import Generator
    ck = os.path.join(ROOT, "oracle", "_ref", "best_netG.pt")
    if not os.path.exists(ck):
        pytest.skip("oracle/_ref/best_netG.pt not staged")
    sd = torch.load(ck, map_location="cpu")
    m = Generator(80, 32, 3)
    m.load_state_dict(sd, strict=True)
    m = m.cuda().eval()
    _, g = load_golden("melgan_real_io.npz")
    wav = m(torch.from_numpy(g["in_mel"]).cuda()).cpu()
    ref = torch.from_numpy(g["out_wav"])
    err = rel_err(wav, ref)
    print("melgan real (40 frames) [tf32x3] rel err", err)
    assert err < 1e-3
    mel = torch.rand(2, 80, 848, generator=torch.Generator().manual_seed(21))
    ref = O.melgan_forward(sd, mel)
    wav = m(mel.cuda()).cpu()
    assert wav.shape == (2, 1, 217088)
    err = rel_err(wav, ref)
    print("melgan real (848 frames, B=2) [tf32x3] rel err", err, "rms ref", float(ref.pow(2).mean().sqrt()))
    assert err < 1e-3
