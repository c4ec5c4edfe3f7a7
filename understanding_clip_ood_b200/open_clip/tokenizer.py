"""Byte-pair-encoding tokenizer producing the same ids as the reference's `SimpleTokenizer`
(deps/open_clip/src/open_clip/tokenizer.py:133-265) — the CPU-side producer of `encode_text`'s input.

Tokenisation is outside the accelerated path ("boundary input producer"); this is an
independent implementation of the published CLIP BPE scheme (lower-cased, whitespace-collapsed text; byte ->
printable-unicode alphabet; ranked merges; `</w>` word-end marker; <start_of_text>/<end_of_text> = last two
ids) so that callers of `get_tokenizer()` keep working.  The merge table itself is DATA that is not shipped
here: point `B200CLIP_BPE_VOCAB` (or `bpe_path=`) at a `bpe_simple_vocab_16e6.txt.gz` (e.g. the one inside an
open_clip checkout).
"""
from __future__ import annotations

import gzip
import html
import os
from functools import lru_cache
from pathlib import Path
from typing import List, Optional, Union

import torch

DEFAULT_CONTEXT_LENGTH = 77
_NUM_MERGES = 49152 - 256 - 2

_SEARCH = (
    Path(__file__).resolve().parent / "bpe_simple_vocab_16e6.txt.gz",
    Path("/root/reference/deps/open_clip/src/open_clip/bpe_simple_vocab_16e6.txt.gz"),
)


def default_bpe() -> str:
    env = os.environ.get("B200CLIP_BPE_VOCAB")
    if env:
        return env
    for cand in _SEARCH:
        if cand.exists():
            return str(cand)
    raise FileNotFoundError("CLIP BPE merge table not found: set B200CLIP_BPE_VOCAB to a bpe_simple_vocab_16e6.txt.gz")


@lru_cache()
def _byte_alphabet() -> dict:
    """byte value -> printable unicode character (printable latin-1 bytes map to themselves, the rest to 256+)."""
    printable = [*range(33, 127), *range(161, 173), *range(174, 256)]
    table, extra = {}, 0
    for b in range(256):
        if b in printable:
            table[b] = chr(b)
        else:
            table[b] = chr(256 + extra)
            extra += 1
    # vocabulary order of the reference: printable bytes first, then the remapped ones
    ordered = [table[b] for b in printable] + [table[b] for b in range(256) if b not in printable]
    return {"map": table, "ordered": ordered}


def _fix_text(text: str) -> str:
    try:
        import ftfy  # optional: identical to the reference when installed
        return ftfy.fix_text(text)
    except ImportError:
        return text


def _clean_lower(text: str) -> str:
    text = html.unescape(html.unescape(_fix_text(text))).strip()
    return " ".join(text.split()).strip().lower()


class SimpleTokenizer:
    def __init__(self, bpe_path: Optional[str] = None, additional_special_tokens: Optional[List[str]] = None,
                 context_length: Optional[int] = DEFAULT_CONTEXT_LENGTH, clean: str = "lower", reduction_mask: str = ""):
        import regex
        if clean != "lower" or reduction_mask:
            raise NotImplementedError("only clean='lower' without a reduction mask is supported")
        alpha = _byte_alphabet()
        self._byte_map = alpha["map"]
        lines = gzip.open(bpe_path or default_bpe()).read().decode("utf-8").split("\n")
        merges = [tuple(ln.split()) for ln in lines[1:_NUM_MERGES + 1]]
        vocab = list(alpha["ordered"]) + [c + "</w>" for c in alpha["ordered"]] + ["".join(m) for m in merges]
        specials = ["<start_of_text>", "<end_of_text>"] + list(additional_special_tokens or [])
        vocab += specials
        self.encoder = {tok: i for i, tok in enumerate(vocab)}
        self.decoder = {i: tok for tok, i in self.encoder.items()}
        self._rank = {m: i for i, m in enumerate(merges)}
        self._cache = {t: [self.encoder[t]] for t in specials}
        self._pat = regex.compile("|".join(regex.escape(s) for s in specials) +
                                  r"""|'s|'t|'re|'ve|'m|'ll|'d|[\p{L}]+|[\p{N}]|[^\s\p{L}\p{N}]+""", regex.IGNORECASE)
        self.vocab_size = len(self.encoder)
        self.all_special_ids = [self.encoder[t] for t in specials]
        self.sot_token_id, self.eot_token_id = self.all_special_ids[0], self.all_special_ids[1]
        self.context_length = context_length

    def _word_ids(self, word: str) -> List[int]:
        """ids of one pre-token after greedy lowest-rank pair merging."""
        hit = self._cache.get(word)
        if hit is not None:
            return hit
        parts = list(word[:-1]) + [word[-1] + "</w>"]
        while len(parts) > 1:
            best, best_rank = -1, None
            for i in range(len(parts) - 1):
                r = self._rank.get((parts[i], parts[i + 1]))
                if r is not None and (best_rank is None or r < best_rank):
                    best, best_rank = i, r
            if best_rank is None:
                break
            a, b = parts[best], parts[best + 1]
            merged, i = [], 0
            while i < len(parts):           # merge every occurrence of the winning pair, left to right
                if i < len(parts) - 1 and parts[i] == a and parts[i + 1] == b:
                    merged.append(a + b)
                    i += 2
                else:
                    merged.append(parts[i])
                    i += 1
            parts = merged
        ids = [self.encoder[p] for p in parts]
        self._cache[word] = ids
        return ids

    def encode(self, text: str) -> List[int]:
        out: List[int] = []
        for tok in self._pat.findall(_clean_lower(text)):
            out.extend(self._word_ids("".join(self._byte_map[b] for b in tok.encode("utf-8"))))
        return out

    def decode(self, tokens) -> str:
        inv = {v: k for k, v in self._byte_map.items()}
        text = "".join(self.decoder[int(t)] for t in tokens)
        return bytearray(inv[c] for c in text if c in inv).decode("utf-8", errors="replace").replace("</w>", " ")

    def __call__(self, texts: Union[str, List[str]], context_length: Optional[int] = None) -> torch.LongTensor:
        if isinstance(texts, str):
            texts = [texts]
        context_length = context_length or self.context_length
        assert context_length, "Please set a valid context length"
        result = torch.zeros(len(texts), context_length, dtype=torch.long)
        for i, text in enumerate(texts):
            ids = [self.sot_token_id, *self.encode(text), self.eot_token_id]
            if len(ids) > context_length:
                ids = ids[:context_length]
                ids[-1] = self.eot_token_id
            result[i, :len(ids)] = torch.tensor(ids)
        return result


_default = None


def tokenize(texts: Union[str, List[str]], context_length: int = DEFAULT_CONTEXT_LENGTH) -> torch.LongTensor:
    global _default
    if _default is None:
        _default = SimpleTokenizer()
    return _default(texts, context_length=context_length)
