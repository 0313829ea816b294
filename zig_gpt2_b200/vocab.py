"""Vocabulary files for the tokenizer (reference: download_weights.py:69-90, main.zig:316-320).

`byte_encoder.json` is reproduced exactly (it is an algorithm, not data).  The real GPT-2
`encoder.json` cannot be downloaded offline, so `synth_encoder` builds a deterministic stand-in
of the same size and shape: ids 0..255 are the 256 single mapped-byte characters (so every byte
string is encodable), then unique multi-character tokens, then `<|endoftext|>` as the last id.
"""
from __future__ import annotations

import json
import os
from typing import Dict

import numpy as np


def unicode_to_bytes() -> "Dict[str, int]":
    """unicode character -> byte, the inverse of GPT-2's bytes_to_unicode (download_weights.py:69-90):
    printable latin-1 bytes map to themselves, the other 68 bytes to U+0100.. in byte order."""
    keep = [*range(ord("!"), ord("~") + 1), *range(0xA1, 0xAC + 1), *range(0xAE, 0xFF + 1)]
    mapping, extra = {}, 0
    for b in keep:
        mapping[chr(b)] = b
    for b in range(256):
        if b not in keep:
            mapping[chr(256 + extra)] = b
            extra += 1
    return mapping


def synth_encoder(vocab_size: int = 50257, seed: int = 7) -> "Dict[str, int]":
    u2b = unicode_to_bytes()
    b2u = {b: u for u, b in u2b.items()}
    enc: "Dict[str, int]" = {}
    for b in range(256):
        enc[b2u[b]] = len(enc)
    rng = np.random.Generator(np.random.PCG64(seed))
    letters = "etaoinshrdlucmfwypvbgkqjxz"
    space = b2u[ord(" ")]
    specials = ["'s", "'t", "'re", "'ve", "'m", "'ll", "'d", space + space, b2u[ord("\n")] + b2u[ord("\n")], "!!", "..."]
    for s in specials:
        if s not in enc and len(enc) < vocab_size - 1:
            enc[s] = len(enc)
    while len(enc) < vocab_size - 1:
        n = int(rng.integers(2, 9))
        kind = rng.random()
        if kind < 0.85:
            idx = np.minimum(rng.geometric(0.12, n) - 1, len(letters) - 1)
            w = "".join(letters[i] for i in idx)
            if rng.random() < 0.3:
                w = w.capitalize()
        elif kind < 0.95:
            w = "".join(str(int(d)) for d in rng.integers(0, 10, min(n, 4)))
        else:
            w = "".join(b2u[int(b)] for b in rng.integers(0x80, 0x100, min(n, 3)))
        if rng.random() < 0.5:
            w = space + w
        if w not in enc:
            enc[w] = len(enc)
    enc["<|endoftext|>"] = len(enc)
    return enc


def write_vocab(model_dir: str, vocab_size: int = 50257, seed: int = 7) -> None:
    os.makedirs(model_dir, exist_ok=True)
    with open(os.path.join(model_dir, "encoder.json"), "w") as f:
        json.dump(synth_encoder(vocab_size, seed), f)
    with open(os.path.join(model_dir, "byte_encoder.json"), "w") as f:
        json.dump(unicode_to_bytes(), f)
