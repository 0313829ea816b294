"""Deterministic inputs for the golden fixtures (shared by make_golden.py and the tests).

Shapes and distributions follow the reference's generate_test_data.py (torch.randn inputs,
nn.Linear's U(-1/sqrt(in), 1/sqrt(in)) init, nn.Embedding's N(0,1), nn.LayerNorm's ones/zeros),
drawn from numpy's RandomState -- a stream numpy guarantees never to change -- so that only the
torch-computed OUTPUTS need to be committed.
"""
from __future__ import annotations

import hashlib

import numpy as np


def _lin(rs, out_f, in_f):
    bound = 1.0 / np.sqrt(in_f)
    return (rs.uniform(-bound, bound, (out_f, in_f)).astype(np.float32),
            rs.uniform(-bound, bound, (out_f,)).astype(np.float32))


def ops_inputs(seed: int = 20231017):
    rs = np.random.RandomState(seed)
    r = lambda *s: rs.standard_normal(s).astype(np.float32)  # noqa: E731
    i = {}
    i["linear_inputs"] = r(3, 768)
    i["linear_weight"], i["linear_bias"] = _lin(rs, 4 * 768, 768)
    i["gelu_inputs"] = r(3, 768)
    i["softmax_inputs"] = r(3, 768)
    i["embedding_weight"] = r(10, 768)
    i["embedding_inputs"] = rs.randint(0, 10, (3,)).astype(np.int64)
    i["layer_norm_inputs"] = r(3, 768)
    i["layer_norm_weight"] = np.ones(768, np.float32)
    i["layer_norm_bias"] = np.zeros(768, np.float32)
    i["layer_norm_affine_weight"] = (1.0 + 0.1 * r(768)).astype(np.float32)
    i["layer_norm_affine_bias"] = (0.1 * r(768)).astype(np.float32)
    for b in (1, 3):
        i[f"transpose_inputs_b{b}"] = r(b, 5, 12, 64)
        i[f"split_inputs_b{b}"] = r(b, 5, 3 * 768)
    i["attn_inputs"] = r(1, 5, 768)
    i["attn_c_attn_weight"], i["attn_c_attn_bias"] = _lin(rs, 3 * 768, 768)
    i["attn_c_proj_weight"], i["attn_c_proj_bias"] = _lin(rs, 768, 768)
    return i


def inputs_digest(i) -> str:
    h = hashlib.sha256()
    for k in sorted(i):
        h.update(k.encode())
        h.update(np.ascontiguousarray(i[k]).tobytes())
    return h.hexdigest()


def gpt_prompt(vocab_size: int, n: int = 16, seed: int = 1235) -> np.ndarray:
    return np.random.RandomState(seed).randint(0, vocab_size, (n,)).astype(np.int64)
