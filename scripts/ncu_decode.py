"""Target for `ncu -k regex:decode_persistent`: a few 24-token launches of the persistent decode kernel (124M by default)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from zig_gpt2_b200 import gpt as G, lib  # noqa: E402
from zig_gpt2_b200.config import SIZES  # noqa: E402
from zig_gpt2_b200.weights import synth_for_size  # noqa: E402

size = sys.argv[1] if len(sys.argv) > 1 else "124M"
n_steps = int(sys.argv[2]) if len(sys.argv) > 2 else 24
L = lib.init(0)
cfg = SIZES[size]
model = G.gpt_from_numpy(cfg, synth_for_size(size))
state = G.State(cfg)
eng = model.engine(state)
prompt = np.random.Generator(np.random.PCG64(1235)).integers(0, cfg.vocab_size, 16).astype(np.uint64)
L.zg_engine_set_prompt(eng, prompt.ctypes.data_as(lib.c_size_p), 16)
L.zg_engine_run_steps(eng, 0, 24)
L.zg_sync()
for _ in range(4):
    L.zg_engine_run_steps(eng, 24, n_steps)
    L.zg_sync()
lib.check()
print("done")
