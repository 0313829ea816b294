"""Cycle-level breakdown of the persistent decode kernel's phases (thread 0 of one CTA, %clock at tagged points).
Usage: python scripts/clock_profile.py [size] [n_steps] [cta] [first position]"""
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
n_steps = int(sys.argv[2]) if len(sys.argv) > 2 else 16
cta = int(sys.argv[3]) if len(sys.argv) > 3 else 100
first = int(sys.argv[4]) if len(sys.argv) > 4 else 24
L = lib.init(0)
cfg = SIZES[size]
model = G.gpt_from_numpy(cfg, synth_for_size(size))
state = G.State(cfg)
eng = model.engine(state)
prompt = np.random.Generator(np.random.PCG64(1235)).integers(0, cfg.vocab_size, 16).astype(np.uint64)
L.zg_engine_set_prompt(eng, prompt.ctypes.data_as(lib.c_size_p), 16)
L.zg_engine_run_steps(eng, 0, first)
for _ in range(10):
    L.zg_engine_run_steps(eng, first, n_steps)
L.zg_sync()
L.zg_timer_begin()
L.zg_engine_run_steps(eng, first, n_steps)
print(f"{size}: {L.zg_timer_end_ms()*1e3/n_steps:.1f} us/token with the timeline off")
L.zg_engine_read_profile(eng, None, 1 + cta)
L.zg_engine_run_steps(eng, first, n_steps)
L.zg_sync()
buf = (C.c_ulonglong * (2 * 16384))()
n = L.zg_engine_read_profile(eng, buf, 2 * 16384)
lib.check()
a = np.array(buf[: 2 * n], dtype=np.int64).reshape(n, 2)
kinds = {0: "P1 qkv", 1: "P2 attn", 2: "P3 proj", 3: "P4 mlp", 4: "P5 reduce", 5: "lm_head"}
pts = {0: "start", 1: "top issued", 2: "gather done", 3: "sync", 4: "x regs + stats", 5: "weights ready", 6: "dot done",
       7: "reduced", 8: "rows done", 9: "fbuf sync", 10: "pass 2 done", 11: "end"}
seg = defaultdict(list)
i = 0
prev_end = {}
while i + 12 <= n:
    tag0 = int(a[i, 0])
    if tag0 < 512:
        i += 1
        continue
    kind = (tag0 - 512) // 16
    t = [int(a[i + k, 1]) for k in range(12)]
    last = 0
    for k in range(1, 12):
        if t[k] != 0 or k == 11:
            d = (t[k] - t[last]) & 0xFFFFFFFF
            if t[k] != 0:
                seg[(kind, last, k)].append(d)
                last = k
    seg[(kind, 0, 99)].append((t[11] - t[0]) & 0xFFFFFFFF)
    if prev_end.get("t") is not None:
        seg[(kind, -1, 0)].append((t[0] - prev_end["t"]) & 0xFFFFFFFF)
    prev_end["t"] = t[11]
    i += 12
L.zg_timer_begin()
L.zg_engine_run_steps(eng, first, n_steps)
print(f"{size}: {L.zg_timer_end_ms()*1e3/n_steps:.1f} us/token with the timeline on")
print(f"{size}: cycles at the SM clock, thread 0 of CTA {cta}; mean over {n_steps} tokens at positions {first}..{first + n_steps - 1}")
for kind in sorted(kinds):
    tot = seg.get((kind, 0, 99))
    if not tot:
        continue
    print(f"  {kinds[kind]:10s} total {np.mean(tot):7.0f} cycles = {np.mean(tot)/1.965e3:5.2f} us  (n={len(tot)})")
    for (k0, a0, b0), v in sorted(seg.items()):
        if k0 == kind and b0 != 99:
            print(f"      {pts.get(a0, 'prev phase end'):>16s} -> {pts[b0]:<16s} {np.mean(v):7.0f}  (median {np.median(v):6.0f})")
