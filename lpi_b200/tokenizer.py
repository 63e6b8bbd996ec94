"""Host-side CLIP byte-pair tokenizer (the text front end feeding PromptLearner).

Behavioural mirror of the reference's `clip.tokenize` / `SimpleTokenizer`
(retrieval/models/clip/clip.py:185-221, retrieval/models/clip/simple_tokenizer.py:62-132): lower-cased, whitespace-collapsed
text -> byte-level BPE with the 48 894 OpenAI merges -> [SOT] ids [EOT] zero-padded to 77, RuntimeError if longer.
Tokenisation is host work, not a GPU job; captions are tokenised once and the ids cached by the callers.

The merge table is the data file `bpe_simple_vocab_16e6.txt.gz` shipped with CLIP (a constant table, byte-identical to the
reference's copy); it ships inside the package at lpi_b200/data/ and can be overridden with $LPI_BPE_VOCAB.
`ftfy.fix_text` is applied when ftfy is installed.  It is the identity on ASCII captions; without ftfy a non-ASCII caption could
tokenise differently from the reference, so that case warns once instead of diverging silently.
"""
from __future__ import annotations

import gzip
import html
import os
from functools import lru_cache
from typing import Dict, List, Sequence, Tuple, Union

import regex as re
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
VOCAB_NAME = "bpe_simple_vocab_16e6.txt.gz"


def find_vocab() -> str:
    cands = [os.environ.get("LPI_BPE_VOCAB"), os.path.join(_HERE, "data", VOCAB_NAME)]
    for c in cands:
        if c and os.path.isfile(c):
            return c
    raise FileNotFoundError(f"{VOCAB_NAME} not found under lpi_b200/data/ (or $LPI_BPE_VOCAB)")


@lru_cache()
def _byte_table() -> Dict[int, str]:
    """Reversible byte -> printable unicode map of GPT-2 style BPE: printable latin-1 bytes map to themselves,
    the remaining 68 bytes to code points 256, 257, ..."""
    keep = list(range(ord("!"), ord("~") + 1)) + list(range(0xA1, 0xAC + 1)) + list(range(0xAE, 0xFF + 1))
    table, extra = {}, 0
    for b in range(256):
        if b in keep:
            table[b] = chr(b)
        else:
            table[b] = chr(256 + extra)
            extra += 1
    return table


_WARNED_FTFY = False


def _clean(text: str) -> str:
    try:
        import ftfy

        text = ftfy.fix_text(text)
    except ImportError:
        global _WARNED_FTFY
        if not _WARNED_FTFY and not text.isascii():
            import warnings

            warnings.warn("lpi_b200.tokenizer: ftfy is not installed; non-ASCII captions skip ftfy.fix_text and may tokenise "
                          "differently from the reference (simple_tokenizer.py:51)", RuntimeWarning, stacklevel=3)
            _WARNED_FTFY = True
    text = html.unescape(html.unescape(text)).strip()
    return re.sub(r"\s+", " ", text).strip()


class Tokenizer:
    def __init__(self, vocab_path: str = None):
        path = vocab_path or find_vocab()
        lines = gzip.open(path).read().decode("utf-8").split("\n")
        merges = [tuple(m.split()) for m in lines[1:49152 - 256 - 2 + 1]]
        byte_syms = list(_byte_table().values())
        # vocabulary order: the 256 byte symbols in the order of bytes_to_unicode()'s value list, their </w> forms, merges, specials
        keep = list(range(ord("!"), ord("~") + 1)) + list(range(0xA1, 0xAC + 1)) + list(range(0xAE, 0xFF + 1))
        ordered = [chr(b) for b in keep] + [_byte_table()[b] for b in range(256) if b not in keep]
        vocab = ordered + [s + "</w>" for s in ordered] + ["".join(m) for m in merges] + ["<|startoftext|>", "<|endoftext|>"]
        assert len(byte_syms) == 256
        self.encoder = {tok: i for i, tok in enumerate(vocab)}
        self.decoder = {i: tok for tok, i in self.encoder.items()}
        self.ranks = {m: i for i, m in enumerate(merges)}
        self.cache = {"<|startoftext|>": "<|startoftext|>", "<|endoftext|>": "<|endoftext|>"}
        self.pat = re.compile(r"""<\|startoftext\|>|<\|endoftext\|>|'s|'t|'re|'ve|'m|'ll|'d|[\p{L}]+|[\p{N}]|[^\s\p{L}\p{N}]+""",
                              re.IGNORECASE)
        self.sot, self.eot = self.encoder["<|startoftext|>"], self.encoder["<|endoftext|>"]

    def _bpe(self, token: str) -> str:
        if token in self.cache:
            return self.cache[token]
        word: Tuple[str, ...] = tuple(token[:-1]) + (token[-1] + "</w>",)
        while len(word) > 1:
            pairs = {(word[i], word[i + 1]) for i in range(len(word) - 1)}
            best = min(pairs, key=lambda p: self.ranks.get(p, float("inf")))
            if best not in self.ranks:
                break
            a, b = best
            merged, i = [], 0
            while i < len(word):
                if i < len(word) - 1 and word[i] == a and word[i + 1] == b:
                    merged.append(a + b)
                    i += 2
                else:
                    merged.append(word[i])
                    i += 1
            word = tuple(merged)
        out = " ".join(word)
        self.cache[token] = out
        return out

    def encode(self, text: str) -> List[int]:
        ids: List[int] = []
        bt = _byte_table()
        for tok in re.findall(self.pat, _clean(text).lower()):
            sym = "".join(bt[b] for b in tok.encode("utf-8"))
            ids.extend(self.encoder[p] for p in self._bpe(sym).split(" "))
        return ids

    def decode(self, ids: Sequence[int]) -> str:
        inv = {v: k for k, v in _byte_table().items()}
        text = "".join(self.decoder[int(i)] for i in ids)
        return bytearray(inv[c] for c in text).decode("utf-8", errors="replace").replace("</w>", " ")


@lru_cache()
def default_tokenizer() -> Tokenizer:
    return Tokenizer()


def tokenize(texts: Union[str, Sequence[str]], context_length: int = 77, truncate: bool = False) -> torch.Tensor:
    """[N, context_length] int64: SOT, ids, EOT, zero padding; raises RuntimeError when a text is too long (clip.py:213-218)."""
    if isinstance(texts, str):
        texts = [texts]
    tk = default_tokenizer()
    out = torch.zeros(len(texts), context_length, dtype=torch.long)
    for i, t in enumerate(texts):
        ids = [tk.sot] + tk.encode(t) + [tk.eot]
        if len(ids) > context_length:
            if not truncate:
                raise RuntimeError(f"Input {t} is too long for context length {context_length}")
            ids = ids[:context_length]
            ids[-1] = tk.eot
        out[i, :len(ids)] = torch.tensor(ids)
    return out
