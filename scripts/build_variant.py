#!/usr/bin/env python
"""Kernel experiments: build zig_gpt2_b200/variants/libzg_<name>.so with extra -D flags on zg_decode.cu only
(the other translation units are compiled once and cached as .o).  Select it with ZG_B200_LIB=<path>.

    python scripts/build_variant.py nowait -DZG_NOWAIT
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "zig_gpt2_b200", "csrc")
VDIR = os.path.join(ROOT, "zig_gpt2_b200", "variants")
os.makedirs(VDIR, exist_ok=True)
BASE = ["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
        "-Xcompiler", "-fPIC"]
name, flags = sys.argv[1], sys.argv[2:]
objs = []
for f in sorted(os.listdir(CSRC)):
    if not f.endswith(".cu"):
        continue
    src = os.path.join(CSRC, f)
    if f == "zg_decode.cu":
        src = os.environ.get("ZG_DECODE_SRC", src)  # e.g. an older revision, for A/B runs on the same box
        obj = os.path.join(VDIR, f"zg_decode_{name}.o")
        subprocess.check_call(BASE + flags + ["-I", CSRC, "-c", src, "-o", obj])
    else:
        obj = os.path.join(VDIR, f[:-3] + ".o")
        if not os.path.exists(obj) or os.path.getmtime(obj) < os.path.getmtime(src):
            subprocess.check_call(BASE + ["-c", src, "-o", obj])
    objs.append(obj)
out = os.path.join(VDIR, f"libzg_{name}.so")
subprocess.check_call(BASE + ["-shared", "-cudart", "static", "-o", out] + objs + ["-lcuda"])
print(out)
