"""Fine-grained timeline of the persistent decode kernel (CTA 0, thread 0: %globaltimer at tagged points).
Usage: python scripts/phase_profile.py [size] [n_steps]
Tags: phase*16 + point; phases 1..5 = P1..P5, 6 = lm_head; points: 1 = activation vector ready (load + LayerNorm),
2 = all ring units of the batch consumed, 3 = reduction + epilogue done, 4 = grid barrier passed."""
import ctypes as C
import os
import sys
from collections import defaultdict

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
if len(sys.argv) > 4:  # optional overrides: n_layer vocab  (e.g. a model small enough to live in L2)
    from zig_gpt2_b200.config import GPTConfig
    from zig_gpt2_b200.weights import synth_weights
    cfg = GPTConfig(int(sys.argv[4]), cfg.context_size, int(sys.argv[3]), cfg.n_heads, cfg.n_embed)
    model = G.gpt_from_numpy(cfg, synth_weights(cfg, seed=5))
else:
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
L.zg_timer_begin()
L.zg_engine_run_steps(eng, 24, n_steps)
ms = L.zg_timer_end_ms()
print(f"{size}: unprofiled launch {ms*1e3/n_steps:.2f} us/token")
L.zg_engine_read_profile(eng, None, 1)  # enable
L.zg_engine_run_steps(eng, 24, n_steps)
L.zg_sync()
buf = (C.c_ulonglong * (2 * 16384))()
n = L.zg_engine_read_profile(eng, buf, 2 * 16384)
lib.check()
a = np.array(buf[: 2 * n], dtype=np.int64).reshape(n, 2)
tags, ts = a[:, 0], a[:, 1]
print(f"{n} marks, total {(ts[-1]-ts[0])/1e3:.1f} us, {(ts[-1]-ts[0])/1e3/n_steps:.2f} us/token (profiled)")
# skip the first token (ring fill): find the first lm_head barrier (tag 100)
first = int(np.argmax(tags == 100)) + 1 if (tags == 100).any() else 0
seg = defaultdict(list)
for i in range(max(first, 1), n):
    seg[(int(tags[i - 1]), int(tags[i]))].append((ts[i] - ts[i - 1]) / 1e3)
names = {1: "P1 qkv", 2: "P2 attn", 3: "P3 proj", 4: "P4 fc", 5: "P5 proj2", 6: "lm_head"}
pts = {0: "start", 1: "vec gathered", 3: "phase done", 4: "token reduced"}
fine = {260: "V:prefetch issued", 261: "V:gather done", 262: "V:gather sync", 263: "G:wait done", 264: "G:dot done",
        265: "G:arrive done", 266: "G:shuffles done", 267: "G:before wait"}
def nm(t):
    if t in fine: return fine[t]
    ph, pt = t // 16, t % 16
    return f"{names.get(ph, ph)}:{pts.get(pt, pt)}"
if any(t >= 256 for t in tags):
    agg = defaultdict(list)
    for (a_, b_), v in seg.items():
        agg[(nm(a_) if a_ >= 256 else "phase-mark", nm(b_) if b_ >= 256 else "phase-mark:" + str(pts.get(b_ % 16, b_ % 16)))].extend(v)
    print("fine-grained segments (mean us, count):")
    for k, v in sorted(agg.items(), key=lambda kv: -np.sum(kv[1])):
        print(f"  {k[0]:22s} -> {k[1]:28s} mean {np.mean(v):6.3f}  n={len(v):5d}  total/token {np.sum(v)/max(1,n_steps-1):7.2f}")
tot = 0.0
for (a_, b_), v in sorted(seg.items(), key=lambda kv: (kv[0][1], kv[0][0])):
    ph, pt = b_ // 16, b_ % 16
    per_tok = np.sum(v) / max(1, n_steps - 1)
    tot += per_tok
    print(f"  {str(names.get(ph, ph)):8s} -> {str(pts.get(pt, pt)):16s} (from tag {a_:3d}): mean {np.mean(v):7.3f} us x {len(v)/max(1,n_steps-1):5.1f}/token = {per_tok:7.2f} us/token")
print(f"  sum {tot:.1f} us/token")
