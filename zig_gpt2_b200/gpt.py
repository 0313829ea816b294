"""Host-side mirror of the reference's src/main.zig over the CUDA C-ABI: GPTConfig, State (the
preallocated buffer set), MLP, Block (owns its KV cache), GPT.forward / GPT.sample, the loaders and
generate().  `GPT.forward` runs the fused persistent decode engine (csrc/zg_decode.cu);
`GPT.forward_unfused` composes the per-op kernels the way main.zig composes ops.zig.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Callable, Dict, List, Optional, Sequence

import numpy as np

from . import lib as _lib
from .config import GPTConfig, SIZES
from .lib import DeviceBuffer, ZgBlock, ZgConfig, ZgGPT, ZgMLP, ZgState
from .ops import CausalSelfAttention, Embedding, LayerNorm, Linear, load_tensor
from .weights import tensor_shapes


class State:  # main.zig:26-65
    FIELDS = ("pos_emb", "x", "o", "logits", "_h", "_4xh", "_qkv", "_q", "_k", "_v", "_attn")

    def __init__(self, config: GPTConfig, want_transpose_scratch: bool = False):
        E, C_ = config.n_embed, config.context_size
        sizes = {"pos_emb": E, "x": E, "o": E, "logits": config.vocab_size, "_h": E, "_4xh": 4 * E, "_qkv": 3 * E,
                 "_q": E, "_attn": C_}
        if want_transpose_scratch:  # the reference's [n,T,hd] copies of the whole cache; the CUDA path never needs them
            sizes["_k"] = sizes["_v"] = C_ * E
        self.bufs: Dict[str, DeviceBuffer] = {k: DeviceBuffer(n) for k, n in sizes.items()}
        self.decoded = bytearray(20)  # main.zig:52
        self.c = ZgState()
        for k in self.FIELDS:
            setattr(self.c, k, self.bufs[k].ptr if k in self.bufs else None)

    def __getattr__(self, name):
        bufs = self.__dict__.get("bufs", {})
        if name in bufs:
            return bufs[name]
        raise AttributeError(name)


class MLP:  # main.zig:67-83
    def __init__(self, c_fc: Linear, c_proj: Linear):
        self.c_fc, self.c_proj = c_fc, c_proj
        self.c = ZgMLP(c_fc.c, c_proj.c)

    def forward(self, inputs, state: State) -> None:
        from .ops import _pl

        ip, il = _pl(inputs)
        _lib.load().zg_mlp_forward(C.byref(self.c), ip, il, C.byref(state.c))
        _lib.check()


class Block:  # main.zig:85-147
    def __init__(self, n_embed: int, ln_1: LayerNorm, attn: CausalSelfAttention, ln_2: LayerNorm, mlp: MLP,
                 k_cache: DeviceBuffer, v_cache: DeviceBuffer):
        self.n_embed, self.ln_1, self.attn, self.ln_2, self.mlp = n_embed, ln_1, attn, ln_2, mlp
        self.k_cache, self.v_cache = k_cache, v_cache
        self.c = ZgBlock(n_embed, ln_1.c, attn.c, ln_2.c, mlp.c, k_cache.ptr, v_cache.ptr)

    def forward(self, seq_len: int, inputs, state: State) -> None:
        from .ops import _pl

        _lib.load().zg_block_forward(C.byref(self.c), seq_len, _pl(inputs)[0], C.byref(state.c))
        _lib.check()


class GPT:  # main.zig:149-208
    def __init__(self, config: GPTConfig, wte: Embedding, wpe: Embedding, h: List[Block], ln_f: LayerNorm, lm_head: Linear):
        self.config, self.wte, self.wpe, self.h, self.ln_f, self.lm_head = config, wte, wpe, h, ln_f, lm_head
        self._blocks = (ZgBlock * len(h))(*[b.c for b in h])
        self.c = ZgGPT(ZgConfig(config.vocab_size, config.context_size, config.n_layer, config.n_heads, config.n_embed),
                       wte.c, wpe.c, self._blocks, ln_f.c, lm_head.c)
        self._engine = None
        self._engine_state = None

    # -- fused persistent engine ----------------------------------------------------------------
    def engine(self, state: State) -> int:
        if self._engine is None or self._engine_state is not state:
            if self._engine is not None:
                _lib.load().zg_engine_destroy(self._engine)
            self._engine = _lib.load().zg_engine_create(C.byref(self.c), C.byref(state.c))
            _lib.check()
            if not self._engine:
                raise _lib.ZgError("zg_engine_create failed")
            self._engine_state = state
        return self._engine

    def forward(self, seq_len: int, token: int, compute_logits: bool, state: State) -> None:
        """GPT.forward (main.zig:178-195): logits land in state.logits."""
        _lib.load().zg_engine_forward(self.engine(state), seq_len, token, int(compute_logits))
        _lib.check()

    def forward_unfused(self, seq_len: int, token: int, compute_logits: bool, state: State) -> None:
        _lib.load().zg_gpt_forward(C.byref(self.c), seq_len, token, int(compute_logits), C.byref(state.c))
        _lib.check()

    def sample(self, seq_len: int, temp: float, token: int, state: State, u: Optional[float] = None) -> int:
        """GPT.sample (main.zig:198-207).  The reference re-seeds its PRNG from the wall clock on every
        call; here the uniform draw is an argument (default: numpy's global generator)."""
        if u is None:
            u = float(np.random.random())
        t = _lib.load().zg_engine_sample(self.engine(state), seq_len, temp, token, u)
        _lib.check()
        return int(t)

    def sample_greedy(self, seq_len: int, token: int, state: State, fused: bool = True) -> int:
        if fused:
            t = _lib.load().zg_engine_sample_greedy(self.engine(state), seq_len, token)
        else:
            t = _lib.load().zg_gpt_sample_greedy(C.byref(self.c), seq_len, token, C.byref(state.c))
        _lib.check()
        return int(t)

    def generate_greedy(self, inputs: Sequence[int], n_total: int, state: State) -> np.ndarray:
        """generate() (main.zig:322-342) with greedy sampling, as one persistent-kernel launch."""
        p = np.ascontiguousarray(inputs, np.uint64)
        out = np.zeros(n_total, np.uint64)
        rc = _lib.load().zg_engine_generate_greedy(self.engine(state), p.ctypes.data_as(_lib.c_size_p), p.size, n_total,
                                                   out.ctypes.data_as(_lib.c_size_p))
        _lib.check()
        if rc:
            raise _lib.ZgError(f"zg_engine_generate_greedy -> {rc}")
        return out.astype(np.int64)

    def generate_sample(self, inputs: Sequence[int], n_total: int, state: State, temp: float, seed: int,
                        sequence: int = 0) -> np.ndarray:
        """generate() with GPT.sample (main.zig:198-207): temperature softmax + inverse-CDF draw on the device, the
        uniform of step s being zg_philox_uniform(seed, s, sequence); no host round trip per token."""
        p = np.ascontiguousarray(inputs, np.uint64)
        out = np.zeros(n_total, np.uint64)
        rc = _lib.load().zg_engine_generate_sample(self.engine(state), p.ctypes.data_as(_lib.c_size_p), p.size, n_total,
                                                   temp, seed, sequence, out.ctypes.data_as(_lib.c_size_p))
        _lib.check()
        if rc:
            raise _lib.ZgError(f"zg_engine_generate_sample -> {rc}")
        return out.astype(np.int64)

    def close(self):
        if self._engine is not None:
            _lib.load().zg_engine_destroy(self._engine)
            self._engine = None


# ---- loaders, main.zig:210-320 ---------------------------------------------------------------------
def _name(model_dir: str, name: str, suffix: str = "") -> str:
    return os.path.join(model_dir, "raw", f"model-{name}{suffix}")


def load_linear(name: str, in_features: int, out_features: int, model_dir: str) -> Linear:
    return Linear(in_features, out_features, load_tensor(_name(model_dir, name, "-w"), (in_features, out_features)),
                  load_tensor(_name(model_dir, name, "-b"), (out_features,)))


def load_layer_norm(name: str, n_features: int, model_dir: str) -> LayerNorm:
    return LayerNorm(n_features, load_tensor(_name(model_dir, name, "-g"), (n_features,)),
                     load_tensor(_name(model_dir, name, "-b"), (n_features,)))


def load_embedding(name: str, vocab_size: int, emb_dim: int, model_dir: str) -> Embedding:
    return Embedding(emb_dim, load_tensor(_name(model_dir, name), (vocab_size, emb_dim)))


def _assemble(config: GPTConfig, get: Callable[[str], DeviceBuffer]) -> GPT:
    E = config.n_embed
    wte, wpe = Embedding(E, get("wte")), Embedding(E, get("wpe"))
    h = []
    for l in range(config.n_layer):
        ln_1 = LayerNorm(E, get(f"h{l}-ln_1-g"), get(f"h{l}-ln_1-b"))
        c_attn = Linear(E, 3 * E, get(f"h{l}-attn-c_attn-w"), get(f"h{l}-attn-c_attn-b"))
        c_proj = Linear(E, E, get(f"h{l}-attn-c_proj-w"), get(f"h{l}-attn-c_proj-b"))
        ln_2 = LayerNorm(E, get(f"h{l}-ln_2-g"), get(f"h{l}-ln_2-b"))
        c_fc = Linear(E, 4 * E, get(f"h{l}-mlp-c_fc-w"), get(f"h{l}-mlp-c_fc-b"))
        c_proj2 = Linear(4 * E, E, get(f"h{l}-mlp-c_proj-w"), get(f"h{l}-mlp-c_proj-b"))
        attn = CausalSelfAttention(config.n_heads, E, c_attn, c_proj)
        k_cache = DeviceBuffer(config.context_size * E)  # main.zig:298-299
        v_cache = DeviceBuffer(config.context_size * E)
        h.append(Block(E, ln_1, attn, ln_2, MLP(c_fc, c_proj2), k_cache, v_cache))
    ln_f = LayerNorm(E, get("ln_f-g"), get("ln_f-b"))
    lm_head = Linear(E, config.vocab_size, wte.weight, None)  # main.zig:312: tied to wte, no bias
    return GPT(config, wte, wpe, h, ln_f, lm_head)


def load_gpt(config: GPTConfig, model_dir: str) -> GPT:
    """load_gpt (main.zig:304-314) from `<model_dir>/raw/model-*` files in the reference's format."""
    shapes = tensor_shapes(config)
    return _assemble(config, lambda n: load_tensor(_name(model_dir, n), shapes[n]))


def gpt_from_numpy(config: GPTConfig, weights: "Dict[str, np.ndarray]") -> GPT:
    """Same assembly from in-memory tensors (synthetic weights: no checkpoint is available offline)."""
    return _assemble(config, lambda n: DeviceBuffer.from_numpy(np.ascontiguousarray(weights[n], np.float32)))


def generate(gpt: GPT, encoder, temp: float, inputs: Sequence[int], state: State, n_total: Optional[int] = None,
             greedy: bool = False, emit: Callable[[bytes], None] = lambda b: None, rng=None, seed: Optional[int] = None) -> List[int]:
    """generate (main.zig:322-342): prompt tokens are forwarded one at a time without logits, then tokens are
    sampled up to context_size; the LAST PROMPT TOKEN IS FORWARDED TWICE (main.zig:329-338), as in the
    reference.  Every token (prompt included) is decoded and emitted (main.zig:339-340)."""
    n_total = gpt.config.context_size if n_total is None else n_total
    out: List[int] = []
    if greedy:
        toks = gpt.generate_greedy(inputs, n_total, state)
        for t in toks:
            out.append(int(t))
            if encoder is not None:
                emit(encoder.decode([int(t)]))
        return out
    if seed is not None:  # device-resident sampling loop, reproducible: u(step) = Philox(seed, step, 0)
        toks = gpt.generate_sample(inputs, n_total, state, temp, seed)
        for t in toks:
            out.append(int(t))
            if encoder is not None:
                emit(encoder.decode([int(t)]))
        return out
    rng = rng or np.random.default_rng()
    token = 0
    for s in range(n_total):
        if s < len(inputs):
            token = int(inputs[s])
            gpt.forward(s + 1, token, False, state)
        else:
            token = gpt.sample(s + 1, temp, token, state, float(rng.random()))
        out.append(token)
        if encoder is not None:
            emit(encoder.decode([token]))
    return out
