"""Eager decoder + vocoder pass at B=4 for ncu launch lists (tools/summarize via --marker decoder)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _pkg; _pkg.load()
from diffsound_b200.modeling.codecs.spec_codec.vqgan import VQModel
from diffsound_b200.vocoder.modules import Generator
from oracle import diffsound_oracle as O
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
dd = dict(double_z=False, z_channels=256, resolution=848, in_channels=1, out_ch=1, ch=128, ch_mult=[1, 1, 2, 2, 4], num_res_blocks=2, attn_resolutions=[53], dropout=0.0)
torch.manual_seed(0)
vq = VQModel(dd, None, n_embed=256, embed_dim=256).cuda().eval(); vq.engine.use_cuda_graph = False
voc = Generator(80, 32, 3); voc.load_state_dict(O.make_melgan_state_dict(seed=1)); voc = voc.cuda().eval(); voc.engine.use_cuda_graph = False
ids = torch.randint(0, 256, (B, 265), device="cuda")
for _ in range(2):
    mel = vq.decode_tokens(ids, (5, 53)); wav = voc((mel[:, 0] + 1) / 2)
torch.cuda.synchronize()
print("done", mel.shape, wav.shape, vq.engine.launches, voc.engine.launches)
