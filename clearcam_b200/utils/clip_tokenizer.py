"""CLIP byte-level BPE tokenizer (host side of `OpenCLIP._encode_text`, models/objects.py:135-141).

Own implementation of the public OpenAI-CLIP / open_clip `SimpleTokenizer` algorithm that the reference ships as
utils/clip_tokenizer.py (same vocabulary file `bpe_simple_vocab_16e6.txt.gz`, same cleaning = html-unescape,
whitespace-collapse, lower-case; same split pattern).  tests/test_tokenizer_cpu.py pins it against token ids
produced by the reference's own tokenizer (tests/golden/clip_tokens.json, made by oracle/make_golden.py).
"""
from __future__ import annotations

import gzip
import html
import os
from functools import lru_cache
from typing import Dict, List, Tuple

import regex

_VOCAB = os.path.join(os.path.dirname(os.path.abspath(__file__)), "bpe_simple_vocab_16e6.txt.gz")
SOT, EOT = "<start_of_text>", "<end_of_text>"
_SPLIT = regex.compile(r"<start_of_text>|<end_of_text>|'s|'t|'re|'ve|'m|'ll|'d|[\p{L}]+|[\p{N}]|[^\s\p{L}\p{N}]+",
                       regex.IGNORECASE)


@lru_cache()
def _byte_alphabet() -> Dict[int, str]:
    """Printable stand-ins for all 256 byte values (GPT-2 convention): visible latin-1 bytes map to themselves,
    the remaining ones to code points 256+."""
    keep = list(range(33, 127)) + list(range(161, 173)) + list(range(174, 256))
    table, extra = {}, 0
    for b in range(256):
        if b in keep:
            table[b] = chr(b)
        else:
            table[b] = chr(256 + extra)
            extra += 1
    return table


class SimpleTokenizer:
    def __init__(self, bpe_path: str = _VOCAB):
        lines = gzip.open(bpe_path).read().decode("utf-8").split("\n")
        merges: List[Tuple[str, str]] = [tuple(l.split()) for l in lines[1:49152 - 256 - 2 + 1]]
        # vocabulary order: 256 byte symbols in *keep-first* order, the same with the end-of-word mark, merges, specials
        alpha = _byte_alphabet()
        keep = list(range(33, 127)) + list(range(161, 173)) + list(range(174, 256))
        order = keep + [b for b in range(256) if b not in keep]
        symbols = [alpha[b] for b in order]
        vocab = symbols + [s + "</w>" for s in symbols] + ["".join(m) for m in merges] + [SOT, EOT]
        self.encoder = {tok: i for i, tok in enumerate(vocab)}
        self.decoder = {i: tok for tok, i in self.encoder.items()}
        self.rank = {m: i for i, m in enumerate(merges)}
        self.sot_token_id, self.eot_token_id = self.encoder[SOT], self.encoder[EOT]
        self._memo: Dict[str, List[str]] = {SOT: [SOT], EOT: [EOT]}

    @staticmethod
    def clean(text: str) -> str:
        text = html.unescape(html.unescape(text)).strip()
        return " ".join(text.split()).strip().lower()

    def _merge(self, token: str) -> List[str]:
        """Greedy lowest-rank-first pair merging of one pre-token."""
        hit = self._memo.get(token)
        if hit is not None:
            return hit
        parts = list(token[:-1]) + [token[-1] + "</w>"]
        while len(parts) > 1:
            best, best_rank = None, None
            for a, b in zip(parts, parts[1:]):
                r = self.rank.get((a, b))
                if r is not None and (best_rank is None or r < best_rank):
                    best, best_rank = (a, b), r
            if best is None:
                break
            merged, i = [], 0
            while i < len(parts):
                if i + 1 < len(parts) and parts[i] == best[0] and parts[i + 1] == best[1]:
                    merged.append(parts[i] + parts[i + 1])
                    i += 2
                else:
                    merged.append(parts[i])
                    i += 1
            parts = merged
        self._memo[token] = parts
        return parts

    def encode(self, text: str) -> List[int]:
        alpha = _byte_alphabet()
        ids: List[int] = []
        for piece in _SPLIT.findall(self.clean(text)):
            mapped = "".join(alpha[b] for b in piece.encode("utf-8"))
            ids.extend(self.encoder[p] for p in self._merge(mapped))
        return ids

    def decode(self, ids) -> str:
        inv = {v: k for k, v in _byte_alphabet().items()}
        text = "".join(self.decoder[int(i)] for i in ids)
        return bytearray(inv[c] for c in text).decode("utf-8", errors="replace").replace("</w>", " ")
