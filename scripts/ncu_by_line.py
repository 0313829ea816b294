#!/usr/bin/env python
"""Aggregate an ncu source-page CSV (SASS rows with stall samples) by CUDA source line.

    ncu -i X.ncu-rep --page source --csv > sass.csv
    cuobjdump -xelf all zig_gpt2_b200/libzg_b200.so ; nvdisasm -g -c zg_decode.sm_100a.cubin > dec.sass
    python scripts/ncu_by_line.py sass.csv dec.sass <mangled-kernel-substring> zig_gpt2_b200/csrc/zg_decode.cu [top]

The i-th instruction of the kernel in the nvdisasm listing is the i-th row of the ncu table (same cubin)."""
import csv
import re
import sys
from collections import defaultdict

sass_csv, disasm, kern, src_path = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 60
rows = list(csv.reader(open(sass_csv)))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
body = rows[hdr_i + 1:]
col = {n: i for i, n in enumerate(hdr)}
# instruction -> (file, line, inlined-at chain) from nvdisasm -g
lines = []
cur = None
infn = False
for l in open(disasm):
    m = re.match(r"\s*\.text\.(\S+):", l)
    if m:
        infn = kern in m.group(1)
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)), m.group(3))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+\S", l):
        lines.append(cur)
print(f"{len(body)} ncu rows, {len(lines)} disassembled instructions", file=sys.stderr)
n = min(len(body), len(lines))
src = open(src_path).read().split("\n")
agg = defaultdict(lambda: defaultdict(float))
stall_cols = [c for c in hdr if c.startswith("stall_") and "Not Issued" not in c]
tot = 0.0
for i in range(n):
    r = body[i]
    key = lines[i][:2] if lines[i] else ("?", 0)
    s = float(r[col["# Samples"]] or 0)
    agg[key]["samples"] += s
    agg[key]["inst"] += float(r[col["Instructions Executed"]] or 0)
    tot += s
    for c in stall_cols:
        v = float(r[col[c]] or 0)
        if v:
            agg[key][c] += v
print(f"total samples {tot:.0f}")
for key, a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:top]:
    stalls = sorted(((c[6:], v) for c, v in a.items() if c.startswith("stall_")), key=lambda x: -x[1])[:3]
    text = src[key[1] - 1].strip()[:70] if key[0].endswith("zg_decode.cu") and 0 < key[1] <= len(src) else ""
    print(f"{100*a['samples']/tot:5.1f}% {key[0][-14:]:>14}:{key[1]:<4} inst {a['inst']:9.0f}  {' '.join(f'{c}={v:.0f}' for c, v in stalls):44s} | {text}")

# ---- optional region summary: pass ranges as name:lo-hi[,lo-hi] after `top` ----
if len(sys.argv) > 6:
    regs = []
    for spec in sys.argv[6:]:
        name, rs = spec.split(":")
        regs.append((name, [tuple(map(int, x.split("-"))) for x in rs.split(",")]))
    out = defaultdict(lambda: [0.0, 0.0])
    for key, a in agg.items():
        nm = "other:" + key[0][-12:]
        if key[0].endswith("zg_decode.cu"):
            nm = "other"
            for name, rs in regs:
                if any(lo <= key[1] <= hi for lo, hi in rs):
                    nm = name
                    break
        out[nm][0] += a["samples"]
        out[nm][1] += a["inst"]
    print("regions:")
    for nm, (s, i) in sorted(out.items(), key=lambda kv: -kv[1][0]):
        print(f"  {nm:24s} samples {100*s/tot:5.1f}%   warp-instructions {i:12.0f}")
