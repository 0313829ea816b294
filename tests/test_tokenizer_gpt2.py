"""The opt-in tokenizer fidelity mode (csrc/host/bpe_gpt2.cpp: GPT-2's merges + pre-tokenizer pattern, no 20-byte limits)
pinned against `transformers`' GPT2Tokenizer (the `tokenizers` byte-level BPE) built from the SAME synthetic vocab.json
/ merges.txt -- no real GPT-2 files exist offline.  The default tokenizer stays the bit-exact mirror of src/bpe.zig
(tests/test_cabi.py, tests/test_oracle_bpe.py); this file also states where the two differ."""
import collections
import json
import os

import numpy as np
import pytest

transformers = pytest.importorskip("transformers")

from zig_gpt2_b200.tokenizer import Gpt2Tokenizer, gpt2_pretokenize  # noqa: E402
from zig_gpt2_b200.vocab import unicode_to_bytes  # noqa: E402

CORPUS = ("the quick brown fox jumps over the lazy dog. it's the dog's day, isn't it? they'll say we've won 1234 times "
          "in 2024!  double  spaces\nand\n\nnewlines\tand tabs. naïve café señor 你好 世界 こんにちは Привет мир "
          "email@example.com costs $9.99 (approx.) -- that's all, folks... THE END ") * 3


def train_merges(text: bytes, n_merges: int):
    """A tiny BPE trainer over the pre-tokenized corpus, so that vocab and merges are mutually consistent."""
    b2u = {b: u for u, b in unicode_to_bytes().items()}
    words = collections.Counter(tuple(b2u[c] for c in piece) for piece in gpt2_pretokenize(text))
    merges = []
    for _ in range(n_merges):
        pairs = collections.Counter()
        for w, c in words.items():
            for a, b in zip(w, w[1:]):
                pairs[(a, b)] += c
        if not pairs:
            break
        (a, b), _ = max(pairs.items(), key=lambda kv: (kv[1], kv[0]))
        merges.append((a, b))
        new = collections.Counter()
        for w, c in words.items():
            out, i = [], 0
            while i < len(w):
                if i + 1 < len(w) and w[i] == a and w[i + 1] == b:
                    out.append(a + b)
                    i += 2
                else:
                    out.append(w[i])
                    i += 1
            new[tuple(out)] += c
        words = new
    return merges


@pytest.fixture(scope="module")
def toks(tmp_path_factory):
    d = tmp_path_factory.mktemp("gpt2tok")
    u2b = unicode_to_bytes()
    b2u = {b: u for u, b in u2b.items()}
    merges = train_merges(CORPUS.encode(), 400)
    vocab = {b2u[b]: b for b in range(256)}
    for a, b in merges:
        vocab.setdefault(a + b, len(vocab))
    vocab["<|endoftext|>"] = len(vocab)
    json.dump(vocab, open(d / "vocab.json", "w"))
    json.dump(u2b, open(d / "byte_encoder.json", "w"))
    open(d / "merges.txt", "w", encoding="utf-8").write("#version: 0.2\n" + "".join(f"{a} {b}\n" for a, b in merges))
    ours = Gpt2Tokenizer(str(d / "vocab.json"), str(d / "merges.txt"), str(d / "byte_encoder.json"))
    hf = transformers.GPT2Tokenizer(str(d / "vocab.json"), str(d / "merges.txt"))
    return ours, hf, len(merges)


TEXTS = [
    "the quick brown fox", "it's the dog's day, isn't it?", "they'll say we've won 1234 times in 2024!",
    "  leading spaces", "trailing spaces   ", "double  spaces   triple", "tabs\tand\nnewlines\n\n  mixed \n x",
    "naïve café señor", "你好 世界", "こんにちは", "Привет мир", "email@example.com costs $9.99 (approx.)",
    "'s 't 're 've 'm 'll 'd 'x '", "I'M SHOUTING: DON'T", "a" * 100, "word" * 30 + " " + "x" * 50, " ", "", "\n", "  ",
    " nbsp emspace　ideographic", "mixed123abc456 7x8", "....!!!???", "🙂 emoji 🙂🙂", "tab\t", "x \n",
]


@pytest.mark.parametrize("text", TEXTS)
def test_ids_equal_transformers_gpt2_tokenizer(toks, text):
    ours, hf, n_merges = toks
    assert n_merges >= 150
    ids = ours.encode(text.encode("utf-8"))
    assert ids == hf.encode(text)
    assert ours.decode(ids) == text.encode("utf-8")


def test_random_strings_equal_transformers(toks):
    ours, hf, _ = toks
    rs = np.random.RandomState(0)
    alphabet = list("abcdefghijklmnopqrstuvwxyzTHE   \n\t'.,!?0123456789-éñ你世界") + ["'s", "'ll", " the", "ing"]
    for _ in range(300):
        text = "".join(alphabet[i] for i in rs.randint(0, len(alphabet), rs.randint(0, 80)))
        ids = ours.encode(text.encode("utf-8"))
        assert ids == hf.encode(text), repr(text)
        assert ours.decode(ids) == text.encode("utf-8")


def test_invalid_utf8_round_trips(toks):
    ours, _, _ = toks
    data = bytes([0xff, 0xfe, 0x41, 0xc3, 0x28, 0x80, 0x20, 0xe2, 0x82]) + "ok".encode()
    assert ours.decode(ours.encode(data)) == data  # byte-level BPE: every byte string is encodable


def test_pretokenizer_pattern_cases():
    """GPT-2's pattern, case by case: 's|'t|'re|'ve|'m|'ll|'d| ?\\p{L}+| ?\\p{N}+| ?[^\\s\\p{L}\\p{N}]+|\\s+(?!\\S)|\\s+"""
    P = lambda s: [p.decode() for p in gpt2_pretokenize(s.encode())]  # noqa: E731
    assert P("it's") == ["it", "'s"]
    assert P("a  b") == ["a", " ", " b"]            # \s+(?!\S) leaves the last space to the word
    assert P("a   b") == ["a", "  ", " b"]
    assert P("a \nb") == ["a", " ", "\n", "b"]      # only a literal space can prefix a word
    assert P("x  ") == ["x", "  "]                  # a trailing run is one piece
    assert P("abc123") == ["abc", "123"]
    assert P(" 12ab") == [" 12", "ab"]
    assert P("hi!!! there") == ["hi", "!!!", " there"]
    assert P("'x") == ["'", "x"]                    # not a contraction
    assert P("we'LL") == ["we", "'", "LL"]          # contractions are case-sensitive in GPT-2


def test_where_the_reference_tokenizer_differs(toks):
    """bpe.zig is not byte-pair encoding: a POSIX split without the space-prefix / lookahead rules (bpe.zig:34-40) and a
    greedy longest-prefix match over the vocabulary (bpe.zig:80-92).  On the same vocabulary the two modes cut
    " the" / double spaces differently -- which is why the fidelity mode is a separate, opt-in class and the default
    tokenizer stays bit-exact to the reference (tests/test_cabi.py)."""
    import zg_oracle as zo

    ours, hf, _ = toks
    u2b = unicode_to_bytes()
    ref = zo.Encoder(hf.get_vocab(), u2b)
    text = b"the  dog's day"
    a, b = ours.encode(text), ref.encode(text)
    assert ours.decode(a) == ref.decode(b) == text  # both are lossless ...
    assert a != b                                    # ... but they are different tokenizations
