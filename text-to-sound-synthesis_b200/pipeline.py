"""caption embeddings -> token grid -> mel -> waveform on one GPU, and the sharded multi-GPU driver.

Mirrors what Diffsound/evaluation/generate_samples_batch.py:143-187 does per caption batch, minus disk I/O: the reference
decodes the mel, copies it to the host, and runs the vocoder at batch 1 through another host round trip (:178-185); here
tokens, mel and waveform stay on the device and the vocoder is batched.
"""
from __future__ import annotations

from typing import Optional

import torch


@torch.no_grad()
def synthesize(dalle, vocoder, cond_emb: torch.Tensor, *, sample_type: str = "top0.85r", seed: Optional[int] = None, codec_batch: Optional[int] = None):
    """cond_emb (B,77,512) on the device -> dict(tokens (B,265) int64, mel (B,1,80,848) in ~[-1,1], wav (B,1,217088)).
    codec_batch: clips per SpecVQGAN-decoder / MelGAN pass (default: the engines' max_batch); the sampler always runs the whole batch."""
    if seed is not None:
        torch.manual_seed(seed)
    if codec_batch:
        dalle.content_codec.engine.max_batch = int(codec_batch)
        if vocoder is not None:
            vocoder.engine.max_batch = int(codec_batch)
    out = dalle.generate_content(batch={"condition_embed": cond_emb}, filter_ratio=0, replicate=1, sample_type=sample_type)
    mel = out["content"]
    spec01 = (mel[:, 0] + 1) / 2  # the script's (spec + 1) / 2 before saving / vocoding (generate_samples_batch.py:181)
    wav = vocoder(spec01) if vocoder is not None else None
    return {"tokens": out["content_token"], "mel": mel, "wav": wav}


@torch.no_grad()
def synthesize_sharded(dalle, vocoder, cond_emb_all: torch.Tensor, *, sample_type="top0.85r", base_seed=1234, gather="wav"):
    """Shard captions over the ranks of the default process group (contiguous blocks), sample independently with seed
    base_seed + rank, and all_gather the finished clips (the only collective of the path)."""
    import torch.distributed as dist
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    B = cond_emb_all.shape[0]
    assert B % world == 0, "caption count must divide the world size"
    per = B // world
    local = synthesize(dalle, vocoder, cond_emb_all[rank * per:(rank + 1) * per].to(dalle.device), sample_type=sample_type, seed=base_seed + rank)
    if world == 1:
        return local
    out = {}
    for k in ("tokens",) + (("wav",) if gather == "wav" and local["wav"] is not None else ()):
        parts = [torch.empty_like(local[k]) for _ in range(world)]
        dist.all_gather(parts, local[k].contiguous())
        out[k] = torch.cat(parts, 0)
    return out


@torch.no_grad()
def synthesize_captions(dalle, vocoder, captions, *, sample_type: str = "top0.85r", replicate: int = 1, seed: Optional[int] = None, shard: bool = False):
    """Captions (list of str) -> dict(tokens, mel, wav), through the model's own text front end (Tokenize + CLIPTextEmbedding, i.e. a DALLE built with
    a condition_codec).  shard=True splits the caption list in contiguous blocks over the ranks of the default process group (seed + rank per
    rank, as synthesize_sharded) and returns this rank's clips together with the caption indices they belong to."""
    import torch.distributed as dist
    world = dist.get_world_size() if (shard and dist.is_initialized()) else 1
    rank = dist.get_rank() if (shard and dist.is_initialized()) else 0
    per = (len(captions) + world - 1) // world
    mine = list(range(rank * per, min(len(captions), (rank + 1) * per)))
    if seed is not None:
        torch.manual_seed(seed + rank)
    if not mine:
        return {"tokens": None, "mel": None, "wav": None, "caption_index": []}
    out = dalle.generate_content(batch={"text": [captions[i] for i in mine], "image": None}, filter_ratio=0, replicate=replicate, sample_type=sample_type)
    mel = out["content"]
    wav = vocoder((mel[:, 0] + 1) / 2) if vocoder is not None else None
    return {"tokens": out["content_token"], "mel": mel, "wav": wav, "caption_index": [i for _ in range(replicate) for i in mine]}


def save_clip(save_root: str, base_name: str, index: int, mel: torch.Tensor, wav: Optional[torch.Tensor], sample_rate: int = 22050):
    """Write one clip the way Diffsound/evaluation/generate_samples_batch.py:173-187 does:
    ``{base}_mel_sample_{n}.npy`` = the (80, 848) mel scaled to [0,1] with (spec + 1) / 2 (the layout/range Codebook/evaluate.py and
    Codebook/evaluation/datasets/fakes.py read back), and ``{base}_mel_sample_{n}.wav`` = mono 24-bit PCM at 22 050 Hz (what
    ``soundfile.write(..., 'PCM_24')`` produces; written here with the standard library, soundfile is not a dependency)."""
    import os
    import wave

    import numpy as np
    os.makedirs(save_root, exist_ok=True)
    stem = os.path.join(save_root, f"{base_name}_mel_sample_{index}")
    spec = mel.detach().float().cpu().numpy()
    spec = spec.reshape(spec.shape[-2], spec.shape[-1])
    np.save(stem + ".npy", (spec + 1) / 2)
    if wav is not None:
        x = np.clip(wav.detach().float().cpu().numpy().reshape(-1), -1.0, 1.0)
        q = np.round(x * 8388607.0).astype(np.int32)  # 2^23 - 1
        b = np.empty((q.size, 3), dtype=np.uint8)
        b[:, 0], b[:, 1], b[:, 2] = q & 0xFF, (q >> 8) & 0xFF, (q >> 16) & 0xFF
        with wave.open(stem + ".wav", "wb") as f:
            f.setnchannels(1)
            f.setsampwidth(3)
            f.setframerate(sample_rate)
            f.writeframes(b.tobytes())
    return stem
