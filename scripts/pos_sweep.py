"""us/token of the persistent decode kernel as a function of the position (context length)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from zig_gpt2_b200 import gpt as G, lib  # noqa: E402
from zig_gpt2_b200.config import SIZES  # noqa: E402
from zig_gpt2_b200.weights import synth_for_size  # noqa: E402

size = sys.argv[1] if len(sys.argv) > 1 else "124M"
L = lib.init(0)
cfg = SIZES[size]
model = G.gpt_from_numpy(cfg, synth_for_size(size))
state = G.State(cfg)
eng = model.engine(state)
prompt = np.random.Generator(np.random.PCG64(1235)).integers(0, cfg.vocab_size, 16).astype(np.uint64)
L.zg_engine_set_prompt(eng, prompt.ctypes.data_as(lib.c_size_p), 16)
L.zg_engine_run_steps(eng, 0, 1016)
L.zg_sync()
n = 16
for first in (24, 60, 90, 120, 200, 230, 330, 460, 700, 1000):
    for _ in range(5):
        L.zg_engine_run_steps(eng, first, n)
    L.zg_sync()
    ts = []
    for _ in range(5):
        L.zg_timer_begin()
        L.zg_engine_run_steps(eng, first, n)
        ts.append(L.zg_timer_end_ms())
    print(f"{size} positions {first}..{first+n-1}: {np.median(ts)*1e3/n:.1f} us/token")
lib.check()
