"""CLIP byte-pair-encoding tokenizer for the text encoder.

The reference tokenizes with transformers' CLIPTokenizer (diffusert/lcm/lcm_controlnet.py:144-150: padding="max_length",
max_length=77, truncation=True). The vocabulary (vocab.json + merges.txt of openai/clip-vit-large-patch14) is a download
and is not present offline, so:
  * `ClipTokenizer(path)` loads those two files from a checkpoint's `tokenizer/` directory and reproduces the published
    algorithm (lower-case, whitespace clean-up, the CLIP regex split, byte -> unicode map, greedy lowest-rank merges, "</w>"
    word ends, <|startoftext|> ... <|endoftext|>, pad with <|endoftext|> to 77);
  * `HashTokenizer()` is the stand-in when no vocabulary is available: deterministic word -> id hashing into the same id
    range with the same BOS/EOS/pad framing. It exercises the whole GPU path; ids are NOT CLIP's.
tests/test_host.py checks ClipTokenizer against transformers.CLIPTokenizer on a small synthetic vocabulary.
"""
import hashlib
import json
import os
import re

BOS, EOS, MAX_LEN = 49406, 49407, 77
# transformers CLIPTokenizer's pattern written for the stdlib `re` (no \p classes): letters / single digits / other symbols
_PAT = re.compile(r"<\|startoftext\|>|<\|endoftext\|>|'s|'t|'re|'ve|'m|'ll|'d|[^\W\d_]+|\d|[^\s\w]+|_+", re.IGNORECASE | re.UNICODE)


def _bytes_to_unicode():
    bs = list(range(ord("!"), ord("~") + 1)) + list(range(ord("\xa1"), ord("\xac") + 1)) + list(range(ord("\xae"), ord("\xff") + 1))
    cs = bs[:]
    n = 0
    for b in range(256):
        if b not in bs:
            bs.append(b)
            cs.append(256 + n)
            n += 1
    return dict(zip(bs, (chr(c) for c in cs)))


def _frame(ids):
    ids = [BOS] + ids[:MAX_LEN - 2] + [EOS]
    return ids + [EOS] * (MAX_LEN - len(ids))


class ClipTokenizer:
    def __init__(self, path):
        with open(os.path.join(path, "vocab.json"), encoding="utf-8") as f:
            self.vocab = json.load(f)
        with open(os.path.join(path, "merges.txt"), encoding="utf-8") as f:
            lines = f.read().strip().split("\n")
        if lines and lines[0].startswith("#"):
            lines = lines[1:]
        self.ranks = {tuple(m.split()): i for i, m in enumerate(lines) if len(m.split()) == 2}
        self.byte_map = _bytes_to_unicode()
        self.unk = self.vocab.get("<|endoftext|>", EOS)
        self.bos = self.vocab.get("<|startoftext|>", BOS)
        self.eos = self.vocab.get("<|endoftext|>", EOS)
        self._cache = {}

    def _bpe(self, token):
        if token in self._cache:
            return self._cache[token]
        word = tuple(token[:-1]) + (token[-1] + "</w>",)
        while len(word) > 1:
            pairs = {(word[i], word[i + 1]) for i in range(len(word) - 1)}
            best = min(pairs, key=lambda p: self.ranks.get(p, float("inf")))
            if best not in self.ranks:
                break
            a, b = best
            out, i = [], 0
            while i < len(word):
                if i < len(word) - 1 and word[i] == a and word[i + 1] == b:
                    out.append(a + b)
                    i += 2
                else:
                    out.append(word[i])
                    i += 1
            word = tuple(out)
        self._cache[token] = word
        return word

    def __call__(self, text):
        text = re.sub(r"\s+", " ", text).strip().lower()
        ids = []
        for tok in _PAT.findall(text):
            tok = "".join(self.byte_map[b] for b in tok.encode("utf-8"))
            ids.extend(self.vocab.get(t, self.unk) for t in self._bpe(tok))
        ids = [self.bos] + ids[:MAX_LEN - 2] + [self.eos]
        return ids + [self.eos] * (MAX_LEN - len(ids))


class HashTokenizer:
    """Stand-in without a vocabulary: one id per word, sha256-derived, in [0, 49406)."""

    def __call__(self, text):
        words = _PAT.findall(re.sub(r"\s+", " ", text).strip().lower())
        return _frame([int.from_bytes(hashlib.sha256(w.encode()).digest()[:4], "little") % BOS for w in words])


def load(path=None, allow_hash=False):
    """The checkpoint's CLIP BPE tokenizer. Without vocab.json + merges.txt this raises: a real text encoder fed with ids
    of another vocabulary conditions on garbage. allow_hash=True (random-init text towers in tests / benchmarks, where ids
    carry no meaning) returns the vocabulary-free HashTokenizer instead."""
    if path and os.path.exists(os.path.join(path, "vocab.json")) and os.path.exists(os.path.join(path, "merges.txt")):
        return ClipTokenizer(path)
    if allow_hash:
        return HashTokenizer()
    raise FileNotFoundError(f"CLIP tokenizer files (vocab.json, merges.txt) not found under {path!r}")
