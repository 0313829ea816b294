#!/usr/bin/env python
"""Blackwell-native evidence, produced here without a GPU (profiles/r02_sass_opcodes.txt, profiles/r02_ptxas_registers.txt):
  * per kernel of zig_gpt2_b200/libzg_b200.so, the count of the SASS mnemonics that prove tcgen05 / TMEM / TMA
    (UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG/UTMAREDG = cp.async.bulk.tensor / cp.reduce,
    UBLKCP = cp.async.bulk, SYNCS = mbarrier, UTCBAR = tcgen05.commit) next to the legacy tensor path (HMMA must be 0);
  * registers / spills / shared memory per kernel from `nvcc -Xptxas -v`."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "zig_gpt2_b200", "libzg_b200.so")
CSRC = os.path.join(ROOT, "zig_gpt2_b200", "csrc")
KEYS = ["UTCHMMA", "UTCQMMA", "UTCIMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UBLKCP", "UTCBAR", "SYNCS", "HMMA", "REDUX", "RED", "LDG.E.ENL2.256"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    return dict(zip(names, out))


def sass():
    txt = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
    per = collections.OrderedDict()
    cur = None
    for ln in txt.split("\n"):
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            per[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
        if m and cur:
            op = m.group(1)
            per[cur]["_total"] += 1
            for k in KEYS:
                if op == k or op.startswith(k + ".") or (k == "LDG.E.ENL2.256" and op.startswith(k)):
                    per[cur][k] += 1
    return per


def main():
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    per = sass()
    names = demangle(list(per))
    with open(os.path.join(ROOT, "profiles", "r02_sass_opcodes.txt"), "w") as f:
        f.write("# cuobjdump -sass zig_gpt2_b200/libzg_b200.so (sm_100a only) -- opcode counts per kernel; scripts/sass_evidence.py\n")
        f.write("# UTC*MMA = tcgen05.mma   LDTM/STTM = tcgen05.ld/st   UTMALDG/UTMASTG/UTMAREDG = TMA tensor load/store/reduce\n")
        f.write("# UBLKCP = cp.async.bulk   UTCBAR = tcgen05.commit   SYNCS = mbarrier   HMMA = legacy mma.sync (must be 0)\n")
        f.write("%-86s %7s " % ("kernel", "instrs") + " ".join("%8s" % k[:8] for k in KEYS) + "\n")
        tot = collections.Counter()
        for fn, c in per.items():
            short = re.sub(r"\(anonymous namespace\)::", "", names[fn])
            short = re.sub(r"\(.*", "", short)[:86]
            f.write("%-86s %7d " % (short, c["_total"]) + " ".join("%8d" % c[k] for k in KEYS) + "\n")
            tot.update(c)
        f.write("%-86s %7d " % ("TOTAL", tot["_total"]) + " ".join("%8d" % tot[k] for k in KEYS) + "\n")
    rows = []
    for cu in sorted(x for x in os.listdir(CSRC) if x.endswith(".cu")):
        r = subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xptxas", "-v",
                            "-I", CSRC, "-c", os.path.join(CSRC, cu), "-o", "/dev/null"], capture_output=True, text=True)
        fn = None
        for ln in r.stderr.split("\n"):
            m = re.search(r"Compiling entry function '(\S+)'", ln)
            if m:
                fn = m.group(1)
            m2 = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", ln)
            if m2 and fn:
                spill = m2.groups()
            m3 = re.search(r"Used (\d+) registers(?:, used (\d+) barriers)?(?:, (\d+) bytes cumulative stack size)?(?:, (\d+) bytes smem)?", ln)
            if m3 and fn:
                rows.append((cu, fn, int(m3.group(1)), spill, m3.group(4) or "0"))
                fn = None
    names = demangle([r[1] for r in rows])
    with open(os.path.join(ROOT, "profiles", "r02_ptxas_registers.txt"), "w") as f:
        f.write("# nvcc -gencode arch=compute_100a,code=sm_100a -O3 -Xptxas -v, per kernel; scripts/sass_evidence.py\n")
        f.write("%-14s %-90s %5s %6s %7s %7s %9s\n" % ("file", "kernel", "regs", "stack", "spill_st", "spill_ld", "static_smem"))
        for cu, fn, regs, spill, smem in rows:
            short = re.sub(r"\(anonymous namespace\)::", "", names[fn])
            short = re.sub(r"\(.*", "", short)[:90]
            f.write("%-14s %-90s %5d %6s %7s %7s %9s\n" % (cu, short, regs, spill[0], spill[1], spill[2], smem))
    print("wrote profiles/r02_sass_opcodes.txt, profiles/r02_ptxas_registers.txt", file=sys.stderr)


if __name__ == "__main__":
    main()
