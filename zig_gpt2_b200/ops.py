"""Host-side mirror of the reference's src/ops.zig operator surface over the CUDA C-ABI.

Same names, fields and argument meaning as the Zig structs; a "slice" is a `DeviceBuffer`
(device pointer + length) or a `(ptr, len)` view of one.  Every `forward` is exactly one C-ABI call
(include/zg_b200.h), like the Zig host shown in INTEGRATION.md.  Nothing here computes on the CPU.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple, Union

import numpy as np

from . import lib as _lib
from .lib import DeviceBuffer, ZgAttention, ZgEmbedding, ZgLayerNorm, ZgLinear

Slice = Union[DeviceBuffer, Tuple[int, int]]  # (device pointer, element count)


def _pl(s: Slice) -> Tuple[int, int]:
    return (s.ptr, s.len) if isinstance(s, DeviceBuffer) else (int(s[0]), int(s[1]))


def view(buf: DeviceBuffer, start: int, stop: int) -> Tuple[int, int]:
    """buf[start..stop] as the Zig host would slice it."""
    assert 0 <= start <= stop <= buf.len
    return (buf.at(start), stop - start)


class Linear:  # ops.zig:4-47
    def __init__(self, in_features: int, out_features: int, weight: Slice, bias: Optional[Slice]):
        self.in_features, self.out_features, self.weight, self.bias = in_features, out_features, weight, bias
        self.c = ZgLinear(in_features, out_features, _pl(weight)[0], _pl(bias)[0] if bias is not None else None)

    def forward(self, inputs: Slice, outputs: Slice) -> None:
        ip, il = _pl(inputs)
        _lib.load().zg_linear_forward(C.byref(self.c), ip, il, _pl(outputs)[0])
        _lib.check()


class Embedding:  # ops.zig:49-68
    def __init__(self, emb_dim: int, weight: Slice):
        self.emb_dim, self.weight = emb_dim, weight
        self.c = ZgEmbedding(emb_dim, _pl(weight)[0])

    def forward(self, idxs: Sequence[int], embeddings: Slice) -> None:
        idx = np.ascontiguousarray(idxs, np.uint64)  # `[]const usize` on the host, 64-bit (tests.zig:93-98)
        _lib.load().zg_embedding_forward(C.byref(self.c), idx.ctypes.data_as(_lib.c_size_p), idx.size, _pl(embeddings)[0])
        _lib.check()


class LayerNorm:  # ops.zig:70-105
    def __init__(self, n_features: int, weight: Slice, bias: Slice, eps: float = 1e-5):
        self.n_features, self.weight, self.bias, self.eps = n_features, weight, bias, eps
        self.c = ZgLayerNorm(n_features, _pl(weight)[0], _pl(bias)[0], eps)

    def forward(self, inputs: Slice) -> None:  # in place
        ip, il = _pl(inputs)
        _lib.load().zg_layer_norm_forward(C.byref(self.c), ip, il)
        _lib.check()


class CausalSelfAttention:  # ops.zig:107-217
    def __init__(self, n_heads: int, n_embed: int, c_attn: Optional[Linear], c_proj: Optional[Linear]):
        self.n_heads, self.n_embed, self.head_dim = n_heads, n_embed, n_embed // n_heads
        self.c_attn, self.c_proj = c_attn, c_proj
        self.c = ZgAttention()
        self.c.n_heads, self.c.n_embed, self.c.head_dim = n_heads, n_embed, self.head_dim
        if c_attn is not None:
            self.c.c_attn = c_attn.c
        if c_proj is not None:
            self.c.c_proj = c_proj.c

    def forward(self, seq_len: int, inputs: Slice, k_cache: Slice, v_cache: Slice, outputs: Slice,
                _qkv: Slice, _q: Slice, _k: Optional[Slice], _v: Optional[Slice], _attn: Optional[Slice]) -> None:
        g = lambda s: _pl(s)[0] if s is not None else None  # noqa: E731
        _lib.load().zg_attention_forward(C.byref(self.c), seq_len, g(inputs), g(k_cache), g(v_cache), g(outputs),
                                         g(_qkv), g(_q), g(_k), g(_v), g(_attn))
        _lib.check()

    def split_qkv(self, seq_len: int, inputs: Slice, split_idx: int, outputs: Slice) -> None:
        ip, il = _pl(inputs)
        _lib.load().zg_split_qkv(C.byref(self.c), seq_len, ip, il, split_idx, _pl(outputs)[0])
        _lib.check()

    @staticmethod
    def transpose(shape: Sequence[int], inputs: Slice, outputs: Slice) -> None:
        ip, il = _pl(inputs)
        sh = (C.c_size_t * 3)(*shape)
        _lib.load().zg_transpose(sh, ip, il, _pl(outputs)[0])
        _lib.check()


def gelu(inputs: Slice) -> None:  # ops.zig:221-228
    ip, il = _pl(inputs)
    _lib.load().zg_gelu(ip, il)
    _lib.check()


def softmax(inputs: Slice) -> None:  # ops.zig:231-241
    ip, il = _pl(inputs)
    _lib.load().zg_softmax(ip, il)
    _lib.check()


def scaled_dot_product_attention(q: Slice, k: Slice, v: Slice, n_heads: int, seq_len: int, head_dim: int,
                                 outputs: Slice, _attn: Optional[Slice] = None) -> None:  # ops.zig:249-307
    kp, kl = _pl(k)
    _lib.load().zg_sdpa(_pl(q)[0], kp, kl, _pl(v)[0], n_heads, seq_len, head_dim, _pl(outputs)[0],
                        _pl(_attn)[0] if _attn is not None else None)
    _lib.check()


def load_tensor(path: str, shape: Sequence[int], dtype=np.float32) -> DeviceBuffer:
    """ops.zig:309-320: headerless little-endian raw file -> (device) slice.  Unlike the reference,
    a short read is an error."""
    n = int(np.prod(shape))
    a = np.fromfile(path, dtype=np.dtype(dtype).newbyteorder("<"), count=n)
    if a.size != n:
        raise ValueError(f"{path}: expected {n} elements, file has {a.size}")
    return DeviceBuffer.from_numpy(a.astype(dtype, copy=False))
