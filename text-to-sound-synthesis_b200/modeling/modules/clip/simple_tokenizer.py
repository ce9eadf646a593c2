"""Drop-in for sound_synthesis/modeling/modules/clip/simple_tokenizer.py::SimpleTokenizer (host side).

CLIP's byte-level BPE, re-implemented from its published description: text is cleaned and lower-cased, split by CLIP's pre-tokenisation
pattern, each piece is mapped byte-by-byte onto a printable-unicode alphabet, and adjacent symbols are merged greedily in the order of the
merge table (lowest rank first) until no listed pair is left; ids are positions in  [256 byte symbols, 256 word-final byte symbols,
one entry per merge, <|startoftext|>, <|endoftext|>].  `end_idx` trims the merge table exactly like the reference (49152 for CLIP).

The merge table itself (`bpe_simple_vocab_16e6.txt.gz`, OpenAI CLIP) is DATA that this repository does not redistribute: pass `bpe_path`, or set
$DIFFSOUND_BPE_VOCAB, or run from a checkout of the reference (its copy is found on sys.path).
"""
from __future__ import annotations

import gzip
import html
import os
import sys
from typing import Dict, List, Tuple

import regex

_VOCAB_NAME = "bpe_simple_vocab_16e6.txt.gz"
_SPLIT = regex.compile(r"<\|startoftext\|>|<\|endoftext\|>|'s|'t|'re|'ve|'m|'ll|'d|[\p{L}]+|[\p{N}]|[^\s\p{L}\p{N}]+", regex.IGNORECASE)


def find_vocab(bpe_path: str | None = None) -> str:
    cands = [bpe_path, os.environ.get("DIFFSOUND_BPE_VOCAB")]
    rel = os.path.join("sound_synthesis", "modeling", "modules", "clip", _VOCAB_NAME)
    cands += [os.path.join(p, rel) for p in sys.path if p] + [os.path.join("/root/reference/Diffsound", rel)]
    for c in cands:
        if c and os.path.isfile(c):
            return c
    raise RuntimeError(f"CLIP BPE merge table {_VOCAB_NAME} not found: pass bpe_path=..., set $DIFFSOUND_BPE_VOCAB, or put the reference checkout on sys.path")


def byte_alphabet() -> Dict[int, str]:
    """byte value -> one printable unicode character: the 188 printable latin-1 bytes map to themselves, the other 68 to U+0100.. in byte order."""
    keep = list(range(0x21, 0x7F)) + list(range(0xA1, 0xAD)) + list(range(0xAE, 0x100))
    table, extra = {}, 0
    for b in range(256):
        if b in keep:
            table[b] = chr(b)
        else:
            table[b] = chr(256 + extra)
            extra += 1
    # the id order of the vocabulary is: kept bytes in ascending order first, then the remapped ones
    return {b: table[b] for b in keep + [b for b in range(256) if b not in keep]}


class SimpleTokenizer:
    def __init__(self, end_idx: int = 49152, bpe_path: str | None = None):
        self.byte_encoder = byte_alphabet()
        self.byte_decoder = {c: b for b, c in self.byte_encoder.items()}
        with gzip.open(find_vocab(bpe_path)) as f:
            lines = f.read().decode("utf-8").split("\n")
        merges: List[Tuple[str, str]] = [tuple(ln.split()) for ln in lines[1:end_idx - 256 - 2 + 1]]
        symbols = list(self.byte_encoder.values())
        vocab = symbols + [s + "</w>" for s in symbols] + ["".join(m) for m in merges] + ["<|startoftext|>", "<|endoftext|>"]
        self.encoder = {tok: i for i, tok in enumerate(vocab)}
        self.decoder = {i: tok for tok, i in self.encoder.items()}
        self.rank = {m: i for i, m in enumerate(merges)}
        self._cache: Dict[str, List[str]] = {"<|startoftext|>": ["<|startoftext|>"], "<|endoftext|>": ["<|endoftext|>"]}

    def _merge(self, piece: str) -> List[str]:
        """Greedy lowest-rank-first pair merging of one pre-token (already in the byte alphabet)."""
        hit = self._cache.get(piece)
        if hit is not None:
            return hit
        word = list(piece[:-1]) + [piece[-1] + "</w>"]
        while len(word) > 1:
            best, best_rank = None, None
            for pair in zip(word, word[1:]):
                r = self.rank.get(pair)
                if r is not None and (best_rank is None or r < best_rank):
                    best, best_rank = pair, r
            if best is None:
                break
            merged, i = [], 0
            while i < len(word):
                if i + 1 < len(word) and word[i] == best[0] and word[i + 1] == best[1]:
                    merged.append(best[0] + best[1])
                    i += 2
                else:
                    merged.append(word[i])
                    i += 1
            word = merged
        self._cache[piece] = word
        return word

    @staticmethod
    def clean(text: str) -> str:
        """html-unescape twice, collapse whitespace, strip, lower.  (The reference also runs ftfy.fix_text first -- a no-op on well-formed text;
        ftfy is used when it is installed.)"""
        try:
            import ftfy
            text = ftfy.fix_text(text)
        except ImportError:
            pass
        text = html.unescape(html.unescape(text)).strip()
        return regex.sub(r"\s+", " ", text).strip().lower()

    def encode(self, text: str) -> List[int]:
        ids: List[int] = []
        for piece in _SPLIT.findall(self.clean(text)):
            mapped = "".join(self.byte_encoder[b] for b in piece.encode("utf-8"))
            ids.extend(self.encoder[s] for s in self._merge(mapped))
        return ids

    def decode(self, tokens) -> str:
        text = "".join(self.decoder[int(t)] for t in tokens).replace("</w>", " ")
        return bytearray(self.byte_decoder[c] if c in self.byte_decoder else ord(" ") for c in text).decode("utf-8", errors="replace")
