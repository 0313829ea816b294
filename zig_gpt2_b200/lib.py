"""ctypes binding of include/zg_b200.h (libzg_b200.so).  No CPU fallback: if the shared library is
missing, or no sm_100 device is present, every entry point raises."""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.environ.get("ZG_B200_LIB") or os.path.join(HERE, "libzg_b200.so")  # override: kernel experiments only

c_float_p = C.POINTER(C.c_float)
c_size_p = C.POINTER(C.c_size_t)


class ZgLinear(C.Structure):  # ops.zig:4-19
    _fields_ = [("in_features", C.c_size_t), ("out_features", C.c_size_t), ("weight", C.c_void_p), ("bias", C.c_void_p)]


class ZgEmbedding(C.Structure):  # ops.zig:49-57
    _fields_ = [("emb_dim", C.c_size_t), ("weight", C.c_void_p)]


class ZgLayerNorm(C.Structure):  # ops.zig:70-80
    _fields_ = [("n_features", C.c_size_t), ("weight", C.c_void_p), ("bias", C.c_void_p), ("eps", C.c_float)]


class ZgAttention(C.Structure):  # ops.zig:107-124
    _fields_ = [("n_heads", C.c_size_t), ("n_embed", C.c_size_t), ("head_dim", C.c_size_t), ("c_attn", ZgLinear), ("c_proj", ZgLinear)]


class ZgConfig(C.Structure):  # main.zig:5-23
    _fields_ = [(n, C.c_size_t) for n in ("vocab_size", "context_size", "n_layer", "n_heads", "n_embed")]


class ZgState(C.Structure):  # main.zig:26-65
    _fields_ = [("pos_emb", C.c_void_p), ("x", C.c_void_p), ("o", C.c_void_p), ("logits", C.c_void_p),
                ("decoded", C.c_void_p), ("_h", C.c_void_p), ("_4xh", C.c_void_p), ("_qkv", C.c_void_p),
                ("_q", C.c_void_p), ("_k", C.c_void_p), ("_v", C.c_void_p), ("_attn", C.c_void_p)]


class ZgMLP(C.Structure):  # main.zig:67-83
    _fields_ = [("c_fc", ZgLinear), ("c_proj", ZgLinear)]


class ZgBlock(C.Structure):  # main.zig:85-117
    _fields_ = [("n_embed", C.c_size_t), ("ln_1", ZgLayerNorm), ("attn", ZgAttention), ("ln_2", ZgLayerNorm),
                ("mlp", ZgMLP), ("k_cache", C.c_void_p), ("v_cache", C.c_void_p)]


class ZgGPT(C.Structure):  # main.zig:149-176
    _fields_ = [("config", ZgConfig), ("wte", ZgEmbedding), ("wpe", ZgEmbedding), ("h", C.POINTER(ZgBlock)),
                ("ln_f", ZgLayerNorm), ("lm_head", ZgLinear)]


# every symbol include/zg_b200.h declares: (restype, argtypes)
V, I, Z, P = None, C.c_int, C.c_size_t, C.c_void_p
SIGNATURES = {
    "zg_init": (I, [I]), "zg_shutdown": (I, []), "zg_device_count": (I, []), "zg_sm_count": (I, []),
    "zg_alloc": (P, [Z]), "zg_free": (I, [P]), "zg_memset": (I, [P, I, Z]),
    "zg_upload": (I, [P, P, Z]), "zg_download": (I, [P, P, Z]), "zg_sync": (I, []),
    "zg_last_error": (I, []), "zg_last_error_string": (C.c_char_p, []), "zg_clear_error": (V, []),
    "zg_set_stream": (I, [P]), "zg_launch_count": (C.c_ulonglong, []), "zg_alloc_count": (C.c_ulonglong, []),
    "zg_philox_uniform": (C.c_double, [C.c_ulonglong, C.c_ulonglong, C.c_ulonglong]),
    "zg_philox4x32_10": (V, [C.POINTER(C.c_uint), C.POINTER(C.c_uint), C.POINTER(C.c_uint)]),
    "zg_timer_begin": (I, []), "zg_timer_end_ms": (C.c_float, []),
    "zg_linear_forward": (V, [C.POINTER(ZgLinear), P, Z, P]),
    "zg_linear_forward_tc": (V, [C.POINTER(ZgLinear), P, Z, P, I, P, I, P, I]),
    "zg_linear_forward_skinny": (V, [C.POINTER(ZgLinear), P, Z, P, I, I, P]),
    "zg_linear_argmax_skinny": (V, [C.POINTER(ZgLinear), P, Z, I, P, P]),
    "zg_to_f16": (V, [P, P, Z]), "zg_tc_error": (I, []), "zg_tc_pair_launch_count": (C.c_ulonglong, []), "zg_tc_set_direct_epilogue": (V, [I]),
    "zg_embedding_forward": (V, [C.POINTER(ZgEmbedding), c_size_p, Z, P]),
    "zg_layer_norm_forward": (V, [C.POINTER(ZgLayerNorm), P, Z]),
    "zg_attention_forward": (V, [C.POINTER(ZgAttention), Z] + [P] * 9),
    "zg_split_qkv": (V, [C.POINTER(ZgAttention), Z, P, Z, Z, P]),
    "zg_transpose": (V, [c_size_p, P, Z, P]),
    "zg_gelu": (V, [P, Z]), "zg_softmax": (V, [P, Z]),
    "zg_sdpa": (V, [P, P, Z, P, Z, Z, Z, P, P]),
    "zg_state_init": (I, [C.POINTER(ZgState), C.POINTER(ZgConfig), I]), "zg_state_free": (V, [C.POINTER(ZgState)]),
    "zg_mlp_forward": (V, [C.POINTER(ZgMLP), P, Z, C.POINTER(ZgState)]),
    "zg_block_forward": (V, [C.POINTER(ZgBlock), Z, P, C.POINTER(ZgState)]),
    "zg_gpt_forward": (V, [C.POINTER(ZgGPT), Z, Z, I, C.POINTER(ZgState)]),
    "zg_gpt_sample": (Z, [C.POINTER(ZgGPT), Z, C.c_float, Z, C.POINTER(ZgState), C.c_double]),
    "zg_gpt_sample_greedy": (Z, [C.POINTER(ZgGPT), Z, Z, C.POINTER(ZgState)]),
    "zg_weight_count": (Z, [C.POINTER(ZgConfig)]), "zg_weight_elems": (Z, [C.POINTER(ZgConfig), Z]),
    "zg_gpt_init": (I, [C.POINTER(ZgGPT), C.POINTER(ZgConfig), C.POINTER(P)]), "zg_gpt_free": (V, [C.POINTER(ZgGPT)]),
    "zg_load_gpt": (I, [C.POINTER(ZgGPT), C.POINTER(ZgConfig), C.c_char_p]),
    "zg_engine_create": (P, [C.POINTER(ZgGPT), C.POINTER(ZgState)]), "zg_engine_destroy": (V, [P]),
    "zg_engine_forward": (V, [P, Z, Z, I]), "zg_engine_sample_greedy": (Z, [P, Z, Z]),
    "zg_engine_sample": (Z, [P, Z, C.c_float, Z, C.c_double]),
    "zg_engine_generate_greedy": (I, [P, c_size_p, Z, Z, c_size_p]),
    "zg_engine_generate_sample": (I, [P, c_size_p, Z, Z, C.c_float, C.c_ulonglong, C.c_ulonglong, c_size_p]),
    "zg_engine_set_prompt": (I, [P, c_size_p, Z]), "zg_engine_run_steps": (V, [P, Z, Z]),
    "zg_engine_read_tokens": (I, [P, Z, Z, c_size_p]),
    "zg_engine_read_profile": (Z, [P, C.POINTER(C.c_ulonglong), Z]),
    "zg_batch_create": (P, [C.POINTER(ZgGPT), Z, Z, Z, I]), "zg_batch_destroy": (V, [P]),
    "zg_batch_forward": (V, [P, Z, c_size_p, I]), "zg_batch_logits": (P, [P]), "zg_batch_logits_pitch": (Z, [P]),
    "zg_batch_prefill": (I, [P, c_size_p, Z, I]), "zg_batch_prefill_resident": (I, [P, Z, I]),
    "zg_batch_generate_greedy": (I, [P, c_size_p, Z, Z, c_size_p, I]),
    "zg_batch_generate_sample": (I, [P, c_size_p, Z, Z, C.c_float, C.c_ulonglong, C.c_ulonglong, c_size_p, I]),
    "zg_batch_set_position": (V, [P, Z]), "zg_batch_run_steps": (V, [P, Z]),
    "zg_batch_read_tokens": (I, [P, c_size_p]), "zg_batch_fused_argmax": (I, [P]), "zg_batch_storage_bits": (I, [P]),
    "zg_batch_k_cache": (P, [P, Z]), "zg_batch_v_cache": (P, [P, Z]),
    "zg_attention_prefill": (V, [P, P, Z, Z, Z, Z]),
    "zg_attention_decode_batch": (V, [P, P, P, Z, Z, Z, Z, Z, P]),
}

_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    """dlopen libzg_b200.so and bind every declared symbol (no device needed for this step)."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO):
            raise RuntimeError(f"{SO} is missing: build it with `python -m zig_gpt2_b200.build` (there is no CPU fallback)")
        L = C.CDLL(SO)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


class ZgError(RuntimeError):
    pass


def check() -> None:
    L = load()
    if L.zg_last_error():
        msg = L.zg_last_error_string().decode()
        L.zg_clear_error()
        raise ZgError(msg)


_inited_device: Optional[int] = None


def init(device: int = 0) -> C.CDLL:
    global _inited_device
    L = load()
    if _inited_device != device:
        rc = L.zg_init(device)
        if rc != 0:
            msg = L.zg_last_error_string().decode()
            L.zg_clear_error()
            raise ZgError(f"zg_init({device}) failed ({rc}): {msg or 'no usable sm_100 CUDA device'}; there is no CPU fallback")
        _inited_device = device
    return L


class DeviceBuffer:
    """A device allocation standing in for a Zig `[]f32` slice: (ptr, len) and nothing else."""

    def __init__(self, n: int, dtype=np.float32, zero: bool = True):
        L = init(_inited_device if _inited_device is not None else 0)
        self.dtype = np.dtype(dtype)
        self.len = int(n)
        self.ptr = L.zg_alloc(max(1, self.len) * self.dtype.itemsize)
        check()
        if zero and self.len:
            L.zg_memset(self.ptr, 0, self.len * self.dtype.itemsize)

    @classmethod
    def from_numpy(cls, a: np.ndarray) -> "DeviceBuffer":
        a = np.ascontiguousarray(a)
        b = cls(a.size, a.dtype, zero=False)
        b.upload(a)
        return b

    def upload(self, a: np.ndarray, offset: int = 0) -> None:
        a = np.ascontiguousarray(a, self.dtype)
        assert offset + a.size <= self.len
        load().zg_upload(self.ptr + offset * self.dtype.itemsize, a.ctypes.data, a.nbytes)
        check()

    def download(self, n: Optional[int] = None, offset: int = 0) -> np.ndarray:
        n = self.len - offset if n is None else n
        out = np.empty(n, self.dtype)
        load().zg_download(out.ctypes.data, self.ptr + offset * self.dtype.itemsize, out.nbytes)
        check()
        return out

    def at(self, offset: int) -> int:
        return self.ptr + offset * self.dtype.itemsize

    def free(self) -> None:
        if self.ptr:
            load().zg_free(self.ptr)
            self.ptr = 0

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass
