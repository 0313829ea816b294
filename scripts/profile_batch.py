"""Short workloads for `ncu` launch lists: one batched decode step (cfg5 shape) and one prefill (cfg3 shape, 2 layers)."""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from zig_gpt2_b200 import gpt as G, lib
from zig_gpt2_b200.batch import BatchEngine
from zig_gpt2_b200.config import SIZES, GPTConfig
from zig_gpt2_b200.weights import synth_weights

which = sys.argv[1]
L = lib.init(0)
if which == "decode":
    cfg = GPTConfig(50257, 1024, 2, 12, 768)
    model = G.gpt_from_numpy(cfg, synth_weights(cfg, seed=1))
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
    eng = BatchEngine(model, B, cache_rows=288, graph=False)
    for _ in range(3):
        eng.set_position(160)
        eng.run_steps(1)
    L.zg_sync()
elif which == "decode_xl":
    cfg = GPTConfig(50257, 1024, 2, 25, 1600)
    model = G.gpt_from_numpy(cfg, synth_weights(cfg, seed=1))
    eng = BatchEngine(model, 64, cache_rows=1024, graph=False, tf32_single_pass=os.environ.get("ZG_TF32") == "1")
    for _ in range(3):
        eng.set_position(1023)
        eng.run_steps(1)
    L.zg_sync()
else:
    cfg = GPTConfig(50257, 1024, 2, 16, 1024)
    model = G.gpt_from_numpy(cfg, synth_weights(cfg, seed=1))
    eng = BatchEngine(model, 16, cache_rows=1024, max_prompt=1024)
    toks = np.random.default_rng(0).integers(0, cfg.vocab_size, (16, 1024))
    for _ in range(3):
        eng.prefill(toks, True)
    L.zg_sync()
lib.check()
