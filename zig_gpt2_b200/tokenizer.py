"""ctypes handles on the C++ host tokenizers (csrc/host/bpe.cpp, bpe_gpt2.cpp; built into libzg_host.so).

`Encoder` is the mirror of the reference's src/bpe.zig (bit-exact: POSIX-regex split + greedy longest-prefix match) and
what `zig_gpt2` uses by default.  `Gpt2Tokenizer` is the opt-in fidelity mode: GPT-2's real merges and pre-tokenizer
pattern, no 20-byte limits.  Both are CPU code in the reference too (bpe.zig never touches BLAS)."""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Sequence

HERE = os.path.dirname(os.path.abspath(__file__))
_H = None


def host() -> C.CDLL:
    global _H
    if _H is None:
        from . import build

        H = C.CDLL(build.build_host())
        sz, szp = C.c_size_t, C.POINTER(C.c_size_t)
        H.zgh_encoder_create_from_files.restype = C.c_void_p
        H.zgh_encoder_create_from_files.argtypes = [C.c_char_p, C.c_char_p]
        H.zgh_encoder_destroy.argtypes = [C.c_void_p]
        H.zgh_encoder_encode.restype = sz
        H.zgh_encoder_encode.argtypes = [C.c_void_p, C.c_char_p, sz, szp, sz]
        H.zgh_encoder_decode.restype = sz
        H.zgh_encoder_decode.argtypes = [C.c_void_p, szp, sz, C.POINTER(C.c_ubyte), sz]
        H.zgh_gpt2_create_from_files.restype = C.c_void_p
        H.zgh_gpt2_create_from_files.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p]
        H.zgh_gpt2_destroy.argtypes = [C.c_void_p]
        H.zgh_gpt2_encode.restype = sz
        H.zgh_gpt2_encode.argtypes = [C.c_void_p, C.c_char_p, sz, szp, sz]
        H.zgh_gpt2_decode.restype = sz
        H.zgh_gpt2_decode.argtypes = [C.c_void_p, szp, sz, C.POINTER(C.c_ubyte), sz]
        H.zgh_gpt2_pretokenize.restype = sz
        H.zgh_gpt2_pretokenize.argtypes = [C.c_char_p, sz, szp, sz]
        _H = H
    return _H


class _Tok:
    _enc = _dec = _free = None

    def __init__(self, handle):
        if not handle:
            raise RuntimeError("tokenizer files missing or malformed")
        self._h = handle

    def encode(self, text: bytes) -> List[int]:
        cap = 4 * len(text) + 16
        out = (C.c_size_t * cap)()
        n = getattr(host(), self._enc)(self._h, text, len(text), out, cap)
        if n == C.c_size_t(-1).value:
            raise ValueError("text does not tokenize with this vocabulary")
        return list(out[:n])

    def decode(self, ids: Sequence[int]) -> bytes:
        arr = (C.c_size_t * len(ids))(*ids)
        cap = 64 * len(ids) + 16
        buf = (C.c_ubyte * cap)()
        n = getattr(host(), self._dec)(self._h, arr, len(ids), buf, cap)
        if n == C.c_size_t(-1).value:
            raise ValueError("unknown token id")
        return bytes(buf[:n])

    def close(self):
        if self._h:
            getattr(host(), self._free)(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Encoder(_Tok):  # bpe.zig:4-119
    _enc, _dec, _free = "zgh_encoder_encode", "zgh_encoder_decode", "zgh_encoder_destroy"

    def __init__(self, model_dir: str):  # load_encoder, main.zig:316-320
        super().__init__(host().zgh_encoder_create_from_files(os.path.join(model_dir, "encoder.json").encode(),
                                                              os.path.join(model_dir, "byte_encoder.json").encode()))


class Gpt2Tokenizer(_Tok):
    _enc, _dec, _free = "zgh_gpt2_encode", "zgh_gpt2_decode", "zgh_gpt2_destroy"

    def __init__(self, encoder_json: str, merges_txt: str, byte_encoder_json: str):
        super().__init__(host().zgh_gpt2_create_from_files(encoder_json.encode(), merges_txt.encode(), byte_encoder_json.encode()))


def gpt2_pretokenize(text: bytes) -> List[bytes]:
    """The pieces GPT-2's pattern cuts `text` into."""
    cap = len(text) + 1
    ends = (C.c_size_t * cap)()
    n = host().zgh_gpt2_pretokenize(text, len(text), ends, cap)
    out, start = [], 0
    for e in ends[:n]:
        out.append(text[start:e])
        start = e
    return out
