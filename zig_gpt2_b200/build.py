"""Builds zig_gpt2_b200/libzg_b200.so (the C-ABI of include/zg_b200.h) with nvcc for sm_100a, in-tree.

    python -m zig_gpt2_b200.build [--force] [--verbose]

nvcc cross-compiles without a GPU; the built .so travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libzg_b200.so")
HOST_OUT = os.path.join(HERE, "libzg_host.so")
CLI_OUT = os.path.join(HERE, "zig_gpt2.bin")  # *.bin is git-ignored; the binary still travels to the GPU box
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared", "-cudart", "static",
]


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _deps():
    d = _sources() + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h", ".hpp"))]
    d.append(os.path.join(os.path.dirname(HERE), "include", "zg_b200.h"))
    return d


def stale() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(p) > t for p in _deps())


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not stale():
        return OUT
    cmd = [NVCC, *FLAGS, *(["-Xptxas", "-v"] if verbose else []), "-o", OUT, *_sources(), "-lcuda"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed building libzg_b200.so")
    return OUT


def build_host(force: bool = False) -> str:
    """g++ build of the C++ host mirror (csrc/host/*.cpp): tokenizer (bpe.zig) + generate loop driver."""
    host_dir = os.path.join(CSRC, "host")
    if not os.path.isdir(host_dir):
        return ""
    srcs = sorted(os.path.join(host_dir, f) for f in os.listdir(host_dir) if f.endswith(".cpp") and f != "main.cpp")
    if not srcs:
        return ""
    deps = srcs + [os.path.join(host_dir, f) for f in os.listdir(host_dir) if f.endswith((".h", ".hpp"))]
    if not force and os.path.exists(HOST_OUT) and all(os.path.getmtime(p) <= os.path.getmtime(HOST_OUT) for p in deps):
        return HOST_OUT
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-o", HOST_OUT, *srcs]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("g++ failed building libzg_host.so")
    return HOST_OUT


def build_cli(force: bool = False) -> str:
    """g++ build of the `zig_gpt2 "<prompt>"` program (main.zig:344-371): csrc/host/main.cpp + bpe.cpp linked against
    libzg_b200.so (found through $ORIGIN at run time).  The CUDA driver library only exists on the GPU box, hence
    --allow-shlib-undefined."""
    host_dir = os.path.join(CSRC, "host")
    srcs = [os.path.join(host_dir, "main.cpp"), os.path.join(host_dir, "bpe.cpp"), os.path.join(host_dir, "bpe_gpt2.cpp")]
    deps = srcs + [os.path.join(host_dir, f) for f in os.listdir(host_dir) if f.endswith((".h", ".hpp"))] + [OUT]
    if not force and os.path.exists(CLI_OUT) and all(os.path.getmtime(p) <= os.path.getmtime(CLI_OUT) for p in deps):
        return CLI_OUT
    cmd = ["g++", "-O2", "-std=c++17", "-Wall", "-o", CLI_OUT, *srcs, "-L", HERE, "-l:libzg_b200.so",
           "-Wl,-rpath,$ORIGIN", "-Wl,--allow-shlib-undefined"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("g++ failed building zig_gpt2.bin")
    return CLI_OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
