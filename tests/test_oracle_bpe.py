"""Oracle restatement of src/bpe.zig.  The reference has no tokenizer tests; these pin the
restatement to the split behaviour SURVEY.md 3.4 verified against glibc, and to a pure-Python
restatement of the same algorithm."""
import re

import pytest

import zg_oracle as zo
from zig_gpt2_b200.vocab import synth_encoder, unicode_to_bytes


@pytest.fixture(scope="module")
def enc():
    return zo.Encoder(synth_encoder(), unicode_to_bytes())


@pytest.fixture(scope="module")
def tables():
    e = synth_encoder()
    u2b = unicode_to_bytes()
    return e, u2b, {b: u for u, b in u2b.items()}


def py_split(data: bytes):
    """The POSIX ERE of bpe.zig:34-40 under leftmost-LONGEST semantics (Python's `re` is
    first-alternative, so take the longest alternative explicitly)."""
    alts = [rb"'s", rb"'t", rb"'re", rb"'ve", rb"'m", rb"'ll", rb"'d",
            rb"[ \t\n\r\f\v]?[A-Za-z]+", rb"[ \t\n\r\f\v]?[0-9]+",
            rb"[ \t\n\r\f\v]?[^ \t\n\r\f\vA-Za-z0-9]+", rb"[ \t\n\r\f\v]+"]
    out, off = [], 0
    while off < len(data):
        best = max((m.end() for m in (re.compile(a).match(data, off) for a in alts) if m), default=off)
        assert best > off
        out.append(data[off:best])
        off = best
    return out


def py_encode(data: bytes, e, b2u):
    ids = []
    for word in py_split(data):
        w = "".join(b2u[b] for b in word).encode("utf-8")
        so, eo = 0, len(w)
        while so < eo:
            key = w[so:eo]
            try:
                k = key.decode("utf-8")
            except UnicodeDecodeError:
                k = None
            if k is not None and k in e:
                ids.append(e[k])
                so, eo = eo, len(w)
            else:
                eo -= 1
    return ids


SPLIT_CASES = [  # SURVEY.md 3.4, probed against glibc
    (b"Marcus Aurelius said thus: ", [b"Marcus", b" Aurelius", b" said", b" thus", b":", b" "]),
    (b"it's 42!!  two  spaces\nnew", [b"it", b"'s", b" 42", b"!!", b"  ", b"two", b"  ", b"spaces", b"\nnew"]),
    (b"'sx 'llama", [b"'s", b"x", b" '", b"llama"]),
    ("café naïve".encode(), [b"caf", "é".encode(), b" na", "ï".encode(), b"ve"]),
    (b"a\tb", [b"a", b"\tb"]),
]


@pytest.mark.parametrize("text,pieces", SPLIT_CASES)
def test_python_restatement_matches_probed_split(text, pieces):
    assert py_split(text) == pieces


@pytest.mark.parametrize("text", [c[0] for c in SPLIT_CASES] + [b"", b" ", b"x", b"hello world 123 !!! it's", b"\n\n\ttabs and  double  spaces "])
def test_encode_matches_python_restatement_and_round_trips(enc, tables, text):
    e, _, b2u = tables
    ids = enc.encode(text)
    assert ids == py_encode(text, e, b2u)
    assert enc.decode(ids) == text  # ids 0..255 cover every byte, so nothing is dropped


def test_decode_every_token(enc, tables):
    e, u2b, _ = tables
    for tok, idx in list(e.items())[:2000]:
        assert enc.decode([idx]) == bytes(u2b[ch] for ch in tok)


def test_word_longer_than_20_bytes_is_flagged_not_overflowed(enc):
    # bpe.zig:71 has a fixed [20]u8 word buffer; the reference would write past it.
    with pytest.raises(OverflowError):
        enc.encode(b"a" * 21)
    assert len(enc.encode(b"a" * 20)) >= 1
