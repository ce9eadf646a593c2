"""Drop-in for sound_synthesis/modeling/codecs/text_codec/tokenize.py::Tokenize (captions -> CLIP token ids + mask; host side of N2).

get_tokens(text) -> {'token': (B, context_length) int64, 'mask': (B, context_length) bool}; with clip_embedding=True the CLIP text tower
(CLIPTextEmbedding, CUDA kernels) is run here as in the reference and its output added as 'embed_token'."""
from __future__ import annotations

from typing import List, Union

import torch
from torch import nn

from ....utils.misc import instantiate_from_config


def tokenize(texts: Union[str, List[str]], context_length: int = 77, add_start_and_end: bool = True, with_mask: bool = True, pad_value: int = 0,
             tokenizer=None, just_token: bool = False):
    """clip.tokenize (reference modules/clip/clip.py:164-216): optional <|startoftext|> / <|endoftext|>, right padding with pad_value, and
    over-long captions truncated to context_length keeping the final token."""
    if isinstance(texts, str):
        texts = [texts]
    sot = [tokenizer.encoder["<|startoftext|>"]] if add_start_and_end else []
    eot = [tokenizer.encoder["<|endoftext|>"]] if add_start_and_end else []
    rows = [sot + tokenizer.encode(t.lower()) + eot for t in texts]
    if just_token:
        return rows
    token = torch.full((len(rows), context_length), pad_value, dtype=torch.long)
    mask = torch.zeros(len(rows), context_length, dtype=torch.bool)
    for i, ids in enumerate(rows):
        if len(ids) > context_length:
            ids = ids[:context_length - 1] + [ids[-1]]
        token[i, :len(ids)] = torch.tensor(ids, dtype=torch.long)
        mask[i, :len(ids)] = True
    return {"token": token, "mask": mask} if with_mask else {"token": token}


class Tokenize(nn.Module):
    def __init__(self, context_length: int = 256, add_start_and_end: bool = False, just_token=False, with_mask: bool = True, pad_value: int = 0,
                 clip_embedding=False, condition_emb_config=None,
                 tokenizer_config={"target": "diffsound_b200.modeling.modules.clip.simple_tokenizer.SimpleTokenizer", "params": {"end_idx": 49152}}):
        super().__init__()
        self.context_length, self.add_start_and_end, self.with_mask = context_length, add_start_and_end, with_mask
        self.pad_value, self.just_token, self.trainable = pad_value, just_token, False
        self.clip_embedding = clip_embedding
        self.condition_emb = None
        if clip_embedding:
            assert condition_emb_config is not None
            self.condition_emb = instantiate_from_config(condition_emb_config)
        self.tokenizer = instantiate_from_config(tokenizer_config)

    def __repr__(self):
        return f"Tokenize for text\n\tcontent_length: {self.context_length}\n\tadd_start_and_end: {self.add_start_and_end}\n\twith_mask: {self.with_mask}"

    def check_length(self, token):
        return len(token) <= self.context_length

    def get_tokens(self, text, **kwargs):
        out = tokenize(text, context_length=self.context_length, add_start_and_end=self.add_start_and_end, with_mask=self.with_mask,
                       pad_value=self.pad_value, tokenizer=self.tokenizer, just_token=self.just_token)
        if self.clip_embedding:
            with torch.no_grad():
                res = self.condition_emb(out["token"].cuda())
            if getattr(self.condition_emb, "additional_last_embedding", False):
                out["embed_token"], out["last_embed"] = res[0].detach(), res[1]
            else:
                out["embed_token"] = res.detach()
        return out
