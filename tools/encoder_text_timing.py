"""Timing of the two once-per-batch front ends: SpecVQGAN encoder + tokeniser (training, N4) and the CLIP text tower (N2)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import _pkg  # noqa: E402

_pkg.load()
from tests.test_gpu_decoder import build_vq  # noqa: E402
from diffsound_b200.modeling.embeddings.clip_text_embedding import CLIPTextEmbedding  # noqa: E402


def timed(fn, n=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


B = 20
vq = build_vq(256, 256, 128, (1, 1, 2, 2, 4))
mel = (torch.rand(B, 1, 80, 848) * 2 - 1).cuda()
enc_ms = timed(lambda: vq.encode(mel))
clip = CLIPTextEmbedding(num_embed=49408, pick_last_embedding=False, embed_dim=512).cuda()
tok = torch.randint(1, 49407, (16, 77)).cuda()
txt_ms = timed(lambda: clip(tok), n=5)
print(json.dumps({"encoder_tokeniser_ms_per_batch20": enc_ms, "mels_per_s": B / enc_ms * 1e3, "encoder_launches": vq.enc_engine.launches,
                  "clip_text_tower_ms_per_batch16": txt_ms, "captions_per_s": 16 / txt_ms * 1e3, "text_launches": clip.engine.launches}))
