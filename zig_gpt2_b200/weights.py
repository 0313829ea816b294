"""Weights in the reference's on-disk format, and the synthetic initialiser.

Format (download_weights.py:57-65, main.zig:210-314): one headerless little-endian fp32 file
per tensor, `<dir>/model-<name>`; Linear weights are stored [out_features, in_features]
(the transpose of the TF checkpoint, download_weights.py:60-61); `wte` is [V, E], `wpe` [C, E].

No checkpoint can be downloaded offline, so `synth_weights` defines the random-init model the
BASELINE configs are measured on (scales in `synth_weights`).
"""
from __future__ import annotations

import hashlib
import os
from typing import Dict, List

import numpy as np

from .config import GPTConfig, SIZES, SIZE_INDEX

BLOCK_KINDS = (
    "ln_1-g", "ln_1-b", "attn-c_attn-w", "attn-c_attn-b", "attn-c_proj-w", "attn-c_proj-b",
    "ln_2-g", "ln_2-b", "mlp-c_fc-w", "mlp-c_fc-b", "mlp-c_proj-w", "mlp-c_proj-b",
)


def tensor_shapes(cfg: GPTConfig) -> "Dict[str, tuple]":
    """name -> shape in the canonical order (wte, wpe, blocks..., ln_f-g, ln_f-b)."""
    E, V, C = cfg.n_embed, cfg.vocab_size, cfg.context_size
    shapes = {"wte": (V, E), "wpe": (C, E)}
    per_block = {
        "ln_1-g": (E,), "ln_1-b": (E,), "attn-c_attn-w": (3 * E, E), "attn-c_attn-b": (3 * E,),
        "attn-c_proj-w": (E, E), "attn-c_proj-b": (E,), "ln_2-g": (E,), "ln_2-b": (E,),
        "mlp-c_fc-w": (4 * E, E), "mlp-c_fc-b": (4 * E,), "mlp-c_proj-w": (E, 4 * E), "mlp-c_proj-b": (E,),
    }
    for l in range(cfg.n_layer):
        for k in BLOCK_KINDS:
            shapes[f"h{l}-{k}"] = per_block[k]
    shapes["ln_f-g"] = (E,)
    shapes["ln_f-b"] = (E,)
    return shapes


def synth_weights(cfg: GPTConfig, seed: int = 1234, std_linear: float = 0.1, std_embed: float = 0.05,
                  std_bias: float = 0.02) -> "Dict[str, np.ndarray]":
    """Random-init GPT-2.  Scales were chosen (and are recorded in DESIGN.md) so that greedy decode
    of the random model is input-dependent with usable top-1/top-2 logit margins: SURVEY.md 8d's
    first proposal (every std 0.02, residual projections / sqrt(2L)) collapses to one repeated
    token.  Linear weights ~ N(0, (0.1*sqrt(768/E))^2) so pre-activations have the same scale at
    every model size; embeddings ~ N(0, 0.05^2); biases and LayerNorm shifts ~ N(0, 0.02^2);
    LayerNorm gains ~ 1 + N(0, 0.02^2)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    out = {}
    lin = std_linear * float(np.sqrt(768.0 / cfg.n_embed))
    for name, shape in tensor_shapes(cfg).items():
        n = int(np.prod(shape))
        t = rng.standard_normal(n, dtype=np.float32)
        if name in ("wte", "wpe"):
            t *= np.float32(std_embed)
        elif name.endswith("-w"):
            t *= np.float32(lin)
        else:
            t *= np.float32(std_bias)
            if name.endswith("-g"):
                t += np.float32(1.0)
        out[name] = t.reshape(shape)
    return out


def synth_for_size(size: str) -> "Dict[str, np.ndarray]":
    return synth_weights(SIZES[size], seed=1234 + SIZE_INDEX[size])


def ordered(weights: "Dict[str, np.ndarray]", cfg: GPTConfig) -> "List[np.ndarray]":
    return [np.ascontiguousarray(weights[n], dtype=np.float32) for n in tensor_shapes(cfg)]


def save_raw(weights: "Dict[str, np.ndarray]", raw_dir: str) -> None:
    os.makedirs(raw_dir, exist_ok=True)
    for name, t in weights.items():
        with open(os.path.join(raw_dir, f"model-{name}"), "wb") as f:
            f.write(np.ascontiguousarray(t, dtype="<f4").tobytes())


def load_raw(cfg: GPTConfig, raw_dir: str) -> "Dict[str, np.ndarray]":
    out = {}
    for name, shape in tensor_shapes(cfg).items():
        t = np.fromfile(os.path.join(raw_dir, f"model-{name}"), dtype="<f4")
        n = int(np.prod(shape))
        if t.size < n:  # the reference accepts short reads (ops.zig:318); we do not
            raise ValueError(f"model-{name}: expected {n} floats, file has {t.size}")
        out[name] = t[:n].reshape(shape)
    return out


def fingerprint(weights: "Dict[str, np.ndarray]") -> str:
    """sha256 over a fixed sample of every tensor: cheap check that two processes (or a
    committed golden fixture) are talking about the same synthetic model."""
    h = hashlib.sha256()
    for name in sorted(weights):
        flat = weights[name].reshape(-1)
        h.update(name.encode())
        h.update(np.ascontiguousarray(flat[:: max(1, flat.size // 4096)]).tobytes())
    return h.hexdigest()
