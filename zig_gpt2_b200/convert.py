"""Converter from a Hugging Face `transformers` GPT-2 state dict to the reference's on-disk format.

The reference's downloader (download_weights.py:57-65) walks a TensorFlow checkpoint and writes one headerless fp32 file
per tensor, transposing every 2-D kernel so that Linear weights are [out_features, in_features].  HF's GPT-2 stores the
same kernels in `Conv1D` modules as [in_features, out_features] -- the TensorFlow orientation -- so the same transpose
applies (download_weights.py:60-61).  Names map as

    transformer.wte.weight                 -> wte            [V, E]        (no transpose: an embedding table)
    transformer.wpe.weight                 -> wpe            [C, E]
    transformer.h.<i>.ln_1.{weight,bias}   -> h<i>-ln_1-{g,b}
    transformer.h.<i>.attn.c_attn.*        -> h<i>-attn-c_attn-{w,b}       w: [E, 3E] -> [3E, E]
    transformer.h.<i>.attn.c_proj.*        -> h<i>-attn-c_proj-{w,b}
    transformer.h.<i>.ln_2.*               -> h<i>-ln_2-{g,b}
    transformer.h.<i>.mlp.c_fc.*           -> h<i>-mlp-c_fc-{w,b}          w: [E, 4E] -> [4E, E]
    transformer.h.<i>.mlp.c_proj.*         -> h<i>-mlp-c_proj-{w,b}        w: [4E, E] -> [E, 4E]
    transformer.ln_f.{weight,bias}         -> ln_f-{g,b}

`lm_head.weight` is tied to `wte` (main.zig:312) and `attn.bias` / `attn.masked_bias` are the causal-mask buffers: all
three are dropped.  No checkpoint can be downloaded offline; tests/test_hf_converter.py builds a random-init
`GPT2LMHeadModel`, converts it and compares logits, which also makes HF an independent third oracle.

    python -m zig_gpt2_b200.convert <hf_model_dir> <out_model_dir>     # writes <out>/raw/model-* and the vocab JSONs
"""
from __future__ import annotations

import json
import os
import sys
from typing import Dict, Mapping

import numpy as np

from .config import GPTConfig
from .weights import save_raw, tensor_shapes

_BLOCK = {
    "ln_1.weight": ("ln_1-g", False), "ln_1.bias": ("ln_1-b", False),
    "attn.c_attn.weight": ("attn-c_attn-w", True), "attn.c_attn.bias": ("attn-c_attn-b", False),
    "attn.c_proj.weight": ("attn-c_proj-w", True), "attn.c_proj.bias": ("attn-c_proj-b", False),
    "ln_2.weight": ("ln_2-g", False), "ln_2.bias": ("ln_2-b", False),
    "mlp.c_fc.weight": ("mlp-c_fc-w", True), "mlp.c_fc.bias": ("mlp-c_fc-b", False),
    "mlp.c_proj.weight": ("mlp-c_proj-w", True), "mlp.c_proj.bias": ("mlp-c_proj-b", False),
}


def config_from_hf(hf_config) -> GPTConfig:
    """GPTConfig (main.zig:5-23) from a transformers GPT2Config (or its dict)."""
    g = (lambda k: hf_config[k]) if isinstance(hf_config, Mapping) else (lambda k: getattr(hf_config, k))
    return GPTConfig(vocab_size=int(g("vocab_size")), context_size=int(g("n_positions")), n_layer=int(g("n_layer")),
                     n_heads=int(g("n_head")), n_embed=int(g("n_embd")))


def _np(t) -> np.ndarray:
    if hasattr(t, "detach"):
        t = t.detach().cpu().float().numpy()
    return np.ascontiguousarray(t, dtype=np.float32)


def from_hf_state_dict(state_dict: Mapping[str, object], cfg: GPTConfig) -> "Dict[str, np.ndarray]":
    """name -> fp32 array in the reference's layout (every tensor `tensor_shapes(cfg)` lists, nothing else)."""
    sd = {k[len("transformer."):] if k.startswith("transformer.") else k: v for k, v in state_dict.items()}
    out: "Dict[str, np.ndarray]" = {"wte": _np(sd["wte.weight"]), "wpe": _np(sd["wpe.weight"])}
    for i in range(cfg.n_layer):
        for hf_name, (ours, transpose) in _BLOCK.items():
            t = _np(sd[f"h.{i}.{hf_name}"])
            out[f"h{i}-{ours}"] = np.ascontiguousarray(t.T) if transpose else t  # download_weights.py:60-61
    out["ln_f-g"], out["ln_f-b"] = _np(sd["ln_f.weight"]), _np(sd["ln_f.bias"])
    shapes = tensor_shapes(cfg)
    for name, shape in shapes.items():
        if name not in out:
            raise KeyError(f"state dict has no tensor for {name}")
        if tuple(out[name].shape) != tuple(shape):
            raise ValueError(f"{name}: expected shape {tuple(shape)}, state dict gives {tuple(out[name].shape)}")
    return {n: out[n] for n in shapes}


def convert_hf_dir(hf_dir: str, out_dir: str) -> GPTConfig:
    """A local HF model directory (config.json + weights, optionally vocab.json) -> `<out_dir>/raw/model-*` plus
    `encoder.json` / `byte_encoder.json` (main.zig:316-320)."""
    from transformers import GPT2LMHeadModel

    from .vocab import unicode_to_bytes

    model = GPT2LMHeadModel.from_pretrained(hf_dir)
    cfg = config_from_hf(model.config)
    save_raw(from_hf_state_dict(model.state_dict(), cfg), os.path.join(out_dir, "raw"))
    vocab = os.path.join(hf_dir, "vocab.json")
    if os.path.exists(vocab):  # GPT-2's vocab.json IS the reference's encoder.json (token string -> id)
        with open(vocab) as f, open(os.path.join(out_dir, "encoder.json"), "w") as g:
            json.dump(json.load(f), g)
    with open(os.path.join(out_dir, "byte_encoder.json"), "w") as f:
        json.dump(unicode_to_bytes(), f)
    return cfg


if __name__ == "__main__":
    if len(sys.argv) != 3:
        sys.exit(__doc__)
    print(convert_hf_dir(sys.argv[1], sys.argv[2]))
