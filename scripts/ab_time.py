"""A/B timing of one build of the decode engine (select it with ZG_B200_LIB): us/token at a few positions plus a
checksum of 64 greedy tokens, so that variants can be compared for speed AND identical output on the same box.
Usage: ZG_B200_LIB=... python scripts/ab_time.py [size] [positions,comma,separated]"""
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from zig_gpt2_b200 import gpt as G, lib  # noqa: E402
from zig_gpt2_b200.config import SIZES  # noqa: E402
from zig_gpt2_b200.weights import synth_for_size  # noqa: E402

size = sys.argv[1] if len(sys.argv) > 1 else "124M"
positions = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "24,200,460").split(",")]
L = lib.init(0)
cfg = SIZES[size]
model = G.gpt_from_numpy(cfg, synth_for_size(size))
state = G.State(cfg)
eng = model.engine(state)
prompt = np.random.Generator(np.random.PCG64(1235)).integers(0, cfg.vocab_size, 16).astype(np.uint64)
pp = prompt.ctypes.data_as(lib.c_size_p)
out = np.zeros(16 + 8 + 296, np.uint64)
rc = L.zg_engine_generate_greedy(eng, pp, 16, len(out), out.ctypes.data_as(lib.c_size_p))
lib.check()
crc = zlib.crc32(out.tobytes())
L.zg_engine_set_prompt(eng, pp, 16)
L.zg_engine_run_steps(eng, 0, max(positions) + 16)
L.zg_sync()
n = 16
res = []
for first in positions:
    for _ in range(8):
        L.zg_engine_run_steps(eng, first, n)
    L.zg_sync()
    ts = []
    for _ in range(9):
        L.zg_timer_begin()
        L.zg_engine_run_steps(eng, first, n)
        ts.append(L.zg_timer_end_ms())
    res.append(f"T{first}: {np.median(ts)*1e3/n:6.1f}")
lib.check()
name = os.path.basename(os.environ.get("ZG_B200_LIB", "libzg_b200.so"))
print(f"{name:28s} {size}  " + "  ".join(res) + f"  us/token   tokens crc {crc:08x} (rc {rc})")
