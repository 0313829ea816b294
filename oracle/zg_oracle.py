"""ctypes wrapper over the CPU ORACLE (oracle/libzg_oracle.so).  TEST INFRASTRUCTURE ONLY.

Importable only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product package (zig_gpt2_b200/) never imports this.
"""
from __future__ import annotations

import ctypes as C
import glob
import os
import subprocess
from typing import Dict, List, Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libzg_oracle.so")

c_size_p = C.POINTER(C.c_size_t)
c_float_p = C.POINTER(C.c_float)


class Linear(C.Structure):
    _fields_ = [("in_features", C.c_size_t), ("out_features", C.c_size_t), ("weight", c_float_p), ("bias", c_float_p)]


class Embedding(C.Structure):
    _fields_ = [("emb_dim", C.c_size_t), ("weight", c_float_p)]


class LayerNorm(C.Structure):
    _fields_ = [("n_features", C.c_size_t), ("weight", c_float_p), ("bias", c_float_p), ("eps", C.c_float)]


class Attention(C.Structure):
    _fields_ = [("n_heads", C.c_size_t), ("n_embed", C.c_size_t), ("head_dim", C.c_size_t), ("c_attn", Linear), ("c_proj", Linear)]


class Config(C.Structure):
    _fields_ = [(n, C.c_size_t) for n in ("vocab_size", "context_size", "n_layer", "n_heads", "n_embed")]


class State(C.Structure):
    _fields_ = [("pos_emb", c_float_p), ("x", c_float_p), ("o", c_float_p), ("logits", c_float_p),
                ("decoded", C.POINTER(C.c_ubyte)), ("_h", c_float_p), ("_4xh", c_float_p), ("_qkv", c_float_p),
                ("_q", c_float_p), ("_k", c_float_p), ("_v", c_float_p), ("_attn", c_float_p)]


class MLP(C.Structure):
    _fields_ = [("c_fc", Linear), ("c_proj", Linear)]


class Block(C.Structure):
    _fields_ = [("n_embed", C.c_size_t), ("ln_1", LayerNorm), ("attn", Attention), ("ln_2", LayerNorm),
                ("mlp", MLP), ("k_cache", c_float_p), ("v_cache", c_float_p)]


class GPT(C.Structure):
    _fields_ = [("config", Config), ("wte", Embedding), ("wpe", Embedding), ("h", C.POINTER(Block)),
                ("ln_f", LayerNorm), ("lm_head", Linear)]


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (oracle/Makefile)."""
    srcs = [os.path.join(_HERE, f) for f in ("zg_ops.c", "zg_model.c", "zg_bpe.c", "zg_oracle.h")]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs):
        subprocess.run(["make", "-C", _HERE, "-B"], check=True, capture_output=True)
    return _SO


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.zo_blas_load.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p]
        L.zo_blas_load.restype = C.c_int
        L.zo_blas_set_threads.argtypes = [C.c_int]
        L.zo_linear_forward.argtypes = [C.POINTER(Linear), c_float_p, C.c_size_t, c_float_p]
        L.zo_embedding_forward.argtypes = [C.POINTER(Embedding), c_size_p, C.c_size_t, c_float_p]
        L.zo_layer_norm_forward.argtypes = [C.POINTER(LayerNorm), c_float_p, C.c_size_t]
        L.zo_attention_forward.argtypes = [C.POINTER(Attention), C.c_size_t] + [c_float_p] * 9
        L.zo_split_qkv.argtypes = [C.POINTER(Attention), C.c_size_t, c_float_p, C.c_size_t, C.c_size_t, c_float_p]
        L.zo_transpose.argtypes = [c_size_p, c_float_p, C.c_size_t, c_float_p]
        L.zo_gelu.argtypes = [c_float_p, C.c_size_t]
        L.zo_softmax.argtypes = [c_float_p, C.c_size_t]
        L.zo_sdpa.argtypes = [c_float_p, c_float_p, C.c_size_t, c_float_p, C.c_size_t, C.c_size_t, C.c_size_t, c_float_p, c_float_p]
        L.zo_state_init.argtypes = [C.POINTER(State), C.POINTER(Config)]
        L.zo_state_free.argtypes = [C.POINTER(State)]
        L.zo_mlp_forward.argtypes = [C.POINTER(MLP), c_float_p, C.c_size_t, C.POINTER(State)]
        L.zo_block_forward.argtypes = [C.POINTER(Block), C.c_size_t, c_float_p, C.POINTER(State)]
        L.zo_gpt_forward.argtypes = [C.POINTER(GPT), C.c_size_t, C.c_size_t, C.c_int, C.POINTER(State)]
        L.zo_gpt_sample.argtypes = [C.POINTER(GPT), C.c_size_t, C.c_float, C.c_size_t, C.POINTER(State), C.c_double]
        L.zo_gpt_sample.restype = C.c_size_t
        L.zo_gpt_sample_greedy.argtypes = [C.POINTER(GPT), C.c_size_t, C.c_size_t, C.POINTER(State)]
        L.zo_gpt_sample_greedy.restype = C.c_size_t
        L.zo_weight_count.argtypes = [C.POINTER(Config)]
        L.zo_weight_count.restype = C.c_size_t
        L.zo_gpt_init.argtypes = [C.POINTER(GPT), C.POINTER(Config), C.POINTER(c_float_p)]
        L.zo_gpt_free.argtypes = [C.POINTER(GPT)]
        L.zo_load_gpt.argtypes = [C.POINTER(GPT), C.POINTER(Config), C.c_char_p, C.POINTER(C.POINTER(c_float_p))]
        L.zo_generate_greedy.argtypes = [C.POINTER(GPT), c_size_p, C.c_size_t, C.c_size_t, C.POINTER(State), c_size_p, c_float_p]
        L.zo_encoder_init.argtypes = [C.POINTER(C.c_char_p), c_size_p, c_size_p, C.c_size_t, C.POINTER(C.c_char_p), c_size_p, C.POINTER(C.c_ubyte), C.c_size_t]
        L.zo_encoder_init.restype = C.c_void_p
        L.zo_encoder_deinit.argtypes = [C.c_void_p]
        L.zo_encoder_encode.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, c_size_p, C.c_size_t]
        L.zo_encoder_encode.restype = C.c_size_t
        L.zo_encoder_decode.argtypes = [C.c_void_p, c_size_p, C.c_size_t, C.POINTER(C.c_ubyte), C.c_size_t]
        L.zo_encoder_decode.restype = C.c_size_t
        _lib = L
    return _lib


def find_openblas() -> Optional[str]:
    """The reference links OpenBLAS on Linux (build.zig:31-32).  A real OpenBLAS ships inside the
    scipy wheel: libscipy_openblas (0.3.31.dev, LP64, symbols prefixed scipy_)."""
    try:
        import scipy
    except Exception:
        return None
    hits = glob.glob(os.path.join(os.path.dirname(scipy.__file__), "..", "scipy.libs", "libscipy_openblas-*.so"))
    return os.path.realpath(hits[0]) if hits else None


def use_openblas(threads: Optional[int] = None) -> bool:
    p = find_openblas()
    if p is None:
        return False
    rc = lib().zo_blas_load(p.encode(), b"scipy_cblas_sgemm", b"scipy_openblas_set_num_threads")
    if rc != 0:
        return False
    if threads is not None:
        lib().zo_blas_set_threads(int(threads))
    return True


def use_scalar_blas() -> None:
    lib().zo_blas_load(None, None, None)


def fp(a: np.ndarray):
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(c_float_p)


def sp(a: np.ndarray):
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(c_size_p)


def mk_linear(w: np.ndarray, b: Optional[np.ndarray]) -> Linear:
    return Linear(w.shape[1], w.shape[0], fp(w), fp(b) if b is not None else None)


# ---- per-op helpers (numpy in, numpy out) ------------------------------------------------
def linear(x: np.ndarray, w: np.ndarray, b: Optional[np.ndarray]) -> np.ndarray:
    x = np.ascontiguousarray(x, np.float32)
    out = np.empty((x.size // w.shape[1], w.shape[0]), np.float32)
    lin = mk_linear(w, b)
    lib().zo_linear_forward(C.byref(lin), fp(x), x.size, fp(out))
    return out


def embedding(w: np.ndarray, idxs: Sequence[int]) -> np.ndarray:
    idx = np.ascontiguousarray(idxs, np.uint64)
    out = np.empty((idx.size, w.shape[1]), np.float32)
    e = Embedding(w.shape[1], fp(w))
    lib().zo_embedding_forward(C.byref(e), sp(idx), idx.size, fp(out))
    return out


def layer_norm(x: np.ndarray, g: np.ndarray, b: np.ndarray, eps: float = 1e-5) -> np.ndarray:
    out = np.array(x, np.float32, order="C")
    ln = LayerNorm(g.size, fp(g), fp(b), eps)
    lib().zo_layer_norm_forward(C.byref(ln), fp(out), out.size)
    return out


def gelu(x: np.ndarray) -> np.ndarray:
    out = np.array(x, np.float32, order="C")
    lib().zo_gelu(fp(out), out.size)
    return out


def softmax(x: np.ndarray) -> np.ndarray:
    out = np.array(x, np.float32, order="C")
    lib().zo_softmax(fp(out), out.size)
    return out


def _attn_struct(n_heads, n_embed, c_attn: Optional[Linear] = None, c_proj: Optional[Linear] = None) -> Attention:
    a = Attention()
    a.n_heads, a.n_embed, a.head_dim = n_heads, n_embed, n_embed // n_heads
    if c_attn is not None:
        a.c_attn = c_attn
    if c_proj is not None:
        a.c_proj = c_proj
    return a


def split_qkv(x: np.ndarray, seq_len: int, n_heads: int, n_embed: int, split_idx: int) -> np.ndarray:
    x = np.ascontiguousarray(x, np.float32)
    out = np.zeros(x.size // 3, np.float32)
    a = _attn_struct(n_heads, n_embed)
    lib().zo_split_qkv(C.byref(a), seq_len, fp(x), x.size, split_idx, fp(out))
    return out


def transpose(x: np.ndarray, shape) -> np.ndarray:
    x = np.ascontiguousarray(x, np.float32)
    out = np.zeros(x.size, np.float32)
    sh = (C.c_size_t * 3)(*shape)
    lib().zo_transpose(sh, fp(x), x.size, fp(out))
    return out


def sdpa(q: np.ndarray, k: np.ndarray, v: np.ndarray, n_heads: int, seq_len: int, head_dim: int) -> np.ndarray:
    q, k, v = (np.ascontiguousarray(t, np.float32) for t in (q, k, v))
    out = np.zeros(q.size, np.float32)
    attn = np.zeros(seq_len, np.float32)
    lib().zo_sdpa(fp(q), fp(k), k.size, fp(v), n_heads, seq_len, head_dim, fp(out), fp(attn))
    return out


class AttentionRunner:
    """Drives CausalSelfAttention.forward token by token the way tests.zig:245-334 does."""

    def __init__(self, n_heads, n_embed, w_attn, b_attn, w_proj, b_proj, context=1024):
        self.keep = [np.ascontiguousarray(t, np.float32) for t in (w_attn, b_attn, w_proj, b_proj)]
        self.a = _attn_struct(n_heads, n_embed, mk_linear(self.keep[0], self.keep[1]), mk_linear(self.keep[2], self.keep[3]))
        E = n_embed
        self.E = E
        self.k_cache = np.zeros(context * E, np.float32)
        self.v_cache = np.zeros(context * E, np.float32)
        self.bufs = [np.zeros(n, np.float32) for n in (3 * E, E, context * E, context * E, context)]

    def step(self, seq_len: int, x: np.ndarray) -> np.ndarray:
        x = np.ascontiguousarray(x, np.float32)
        out = np.zeros(self.E, np.float32)
        qkv, q, k, v, attn = self.bufs
        lib().zo_attention_forward(C.byref(self.a), seq_len, fp(x), fp(self.k_cache), fp(self.v_cache), fp(out),
                                   fp(qkv), fp(q), fp(k), fp(v), fp(attn))
        return out


class Model:
    """GPT + State over numpy weights (dict in zig_gpt2_b200.weights naming)."""

    def __init__(self, cfg, weights: "Dict[str, np.ndarray]"):
        from zig_gpt2_b200.weights import ordered

        self.cfg = cfg
        self.c = Config(cfg.vocab_size, cfg.context_size, cfg.n_layer, cfg.n_heads, cfg.n_embed)
        self._w = ordered(weights, cfg)
        arr = (c_float_p * len(self._w))(*[fp(t) for t in self._w])
        self.gpt = GPT()
        if lib().zo_gpt_init(C.byref(self.gpt), C.byref(self.c), arr) != 0:
            raise MemoryError("zo_gpt_init")
        self.state = State()
        if lib().zo_state_init(C.byref(self.state), C.byref(self.c)) != 0:
            raise MemoryError("zo_state_init")

    def close(self):
        if self.gpt is not None:
            lib().zo_gpt_free(C.byref(self.gpt))
            lib().zo_state_free(C.byref(self.state))
            self.gpt = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _view(self, ptr, n):
        return np.ctypeslib.as_array(ptr, shape=(n,))

    def forward(self, seq_len: int, token: int, compute_logits: bool = True) -> Optional[np.ndarray]:
        lib().zo_gpt_forward(C.byref(self.gpt), seq_len, token, int(compute_logits), C.byref(self.state))
        return self._view(self.state.logits, self.cfg.vocab_size).copy() if compute_logits else None

    def x(self) -> np.ndarray:
        return self._view(self.state.x, self.cfg.n_embed).copy()

    def kv(self, layer: int, seq_len: int):
        E = self.cfg.n_embed
        b = self.gpt.h[layer]
        return (self._view(b.k_cache, seq_len * E).reshape(seq_len, E).copy(),
                self._view(b.v_cache, seq_len * E).reshape(seq_len, E).copy())

    def sample(self, seq_len: int, temp: float, token: int, u: float) -> int:
        return int(lib().zo_gpt_sample(C.byref(self.gpt), seq_len, temp, token, C.byref(self.state), u))

    def generate_greedy(self, prompt: Sequence[int], n_total: int, want_logits: bool = False):
        p = np.ascontiguousarray(prompt, np.uint64)
        out = np.zeros(n_total, np.uint64)
        n_gen = max(0, n_total - p.size)
        logits = np.zeros((n_gen, self.cfg.vocab_size), np.float32) if want_logits else None
        lib().zo_generate_greedy(C.byref(self.gpt), sp(p), p.size, n_total, C.byref(self.state), sp(out),
                                 fp(logits) if want_logits else None)
        return (out.astype(np.int64), logits) if want_logits else out.astype(np.int64)


class Encoder:
    def __init__(self, token_to_idx: "Dict[str, int]", unicode_to_byte: "Dict[str, int]"):
        toks = [k.encode("utf-8") for k in token_to_idx]
        ids = np.ascontiguousarray(list(token_to_idx.values()), np.uint64)
        tl = np.ascontiguousarray([len(t) for t in toks], np.uint64)
        unis = [k.encode("utf-8") for k in unicode_to_byte]
        ul = np.ascontiguousarray([len(t) for t in unis], np.uint64)
        ub = (C.c_ubyte * len(unis))(*unicode_to_byte.values())
        self._h = lib().zo_encoder_init((C.c_char_p * len(toks))(*toks), sp(tl), sp(ids), len(toks),
                                        (C.c_char_p * len(unis))(*unis), sp(ul), ub, len(unis))
        if not self._h:
            raise RuntimeError("zo_encoder_init failed")

    def encode(self, text: bytes, max_out: int = 1024) -> "List[int]":
        out = np.zeros(max_out, np.uint64)
        n = lib().zo_encoder_encode(self._h, text, len(text), sp(out), max_out)
        if n == C.c_size_t(-1).value:
            raise OverflowError("word longer than the reference's 20-byte buffer, or output overflow")
        return [int(t) for t in out[:n]]

    def decode(self, ids: Sequence[int], max_out: int = 1 << 16) -> bytes:
        a = np.ascontiguousarray(ids, np.uint64)
        buf = (C.c_ubyte * max_out)()
        n = lib().zo_encoder_decode(self._h, sp(a), a.size, buf, max_out)
        if n == C.c_size_t(-1).value:
            raise OverflowError("decode: unknown id or output overflow")
        return bytes(buf[:n])

    def __del__(self):
        try:
            lib().zo_encoder_deinit(self._h)
        except Exception:
            pass
