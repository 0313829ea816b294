"""Per-phase device timestamps of the persistent decode kernel (CTA 0, %globaltimer at every grid barrier).
Usage: python scripts/phase_profile.py [size] [n_steps]   -> prints mean us per phase kind."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from zig_gpt2_b200 import gpt as G, lib  # noqa: E402
from zig_gpt2_b200.config import SIZES  # noqa: E402
from zig_gpt2_b200.weights import synth_for_size  # noqa: E402

size = sys.argv[1] if len(sys.argv) > 1 else "124M"
n_steps = int(sys.argv[2]) if len(sys.argv) > 2 else 32
L = lib.init(0)
cfg = SIZES[size]
model = G.gpt_from_numpy(cfg, synth_for_size(size))
state = G.State(cfg)
eng = model.engine(state)
prompt = np.random.Generator(np.random.PCG64(1235)).integers(0, cfg.vocab_size, 16).astype(np.uint64)
L.zg_engine_set_prompt(eng, prompt.ctypes.data_as(lib.c_size_p), 16)
L.zg_engine_run_steps(eng, 0, 24)
L.zg_sync()
for _ in range(20):
    L.zg_engine_run_steps(eng, 24, n_steps)
L.zg_sync()
L.zg_engine_read_profile(eng, None, 1)  # enable
L.zg_engine_run_steps(eng, 24, n_steps)
L.zg_sync()
buf = (C.c_ulonglong * 8192)()
n = L.zg_engine_read_profile(eng, buf, 8192)
lib.check()
ts = np.array(buf[:n], dtype=np.int64)
d = np.diff(ts) / 1e3  # us; entries: start, barrier..., end
per_tok = 5 * cfg.n_layer + 1
nb = n - 2
tok = nb // per_tok
print(f"{size}: {n_steps} steps, {nb} barriers, total {(ts[-1]-ts[0])/1e3:.1f} us, {(ts[-1]-ts[0])/1e3/n_steps:.2f} us/token")
if tok >= 2:
    body = d[: tok * per_tok].reshape(tok, per_tok)[1:]  # drop the first token (ring fill)
    names = ["P1 ln1+qkv", "P2 attn", "P3 proj", "P4 ln2+fc", "P5 proj2"]
    layers = body[:, : 5 * cfg.n_layer].reshape(-1, cfg.n_layer, 5)
    for i, nm in enumerate(names):
        print(f"  {nm:12s} mean {layers[:, :, i].mean():6.2f} us  (layer0 {layers[:, 0, i].mean():6.2f}, last {layers[:, -1, i].mean():6.2f})")
    print(f"  lm_head+amax mean {body[:, -1].mean():6.2f} us")
    print(f"  per token: layers {layers.sum(axis=(1, 2)).mean():.1f} us + lm_head {body[:, -1].mean():.1f} us")
