"""GPU check of the tcgen05 GEMM (zg_linear_forward_tc) against float64 numpy on operands that are exactly
representable in the tensor-core input format, plus timing of the prefill shapes.  Run under gpurun."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zig_gpt2_b200 import lib  # noqa: E402
from zig_gpt2_b200.lib import DeviceBuffer, ZgLinear  # noqa: E402


def to_tf32(a):
    u = a.astype(np.float32).view(np.uint32)
    return (u & np.uint32(0xFFFFE000)).view(np.float32)


def to_f16_bits(a):
    return a.astype(np.float16).view(np.uint16)


def f16_to_f32(b):
    return b.view(np.float16).astype(np.float32)


def gelu(x):
    return 0.5 * x * (1.0 + np.tanh(x * 0.7978845608 * (1.0 + 0.044715 * x * x)))


def run_case(L, M, N, K, prec, epi, bn, rng, with_bias=True):
    x = rng.standard_normal((M, K)).astype(np.float32)
    w = (rng.standard_normal((N, K)) * 0.05).astype(np.float32)
    b = rng.standard_normal(N).astype(np.float32) if with_bias else None
    res = rng.standard_normal((M, N)).astype(np.float32) if epi == 2 else None
    if prec in (0, 2, 3):
        if prec == 0:
            x, w = to_tf32(x), to_tf32(w)
        if prec == 3:  # single-pass tf32 on operands that are NOT representable: shows the truncation error
            prec = 0
        dx, dw = DeviceBuffer.from_numpy(x), DeviceBuffer.from_numpy(w)
        lowp = None
        xin = dx.ptr
    else:
        xb, wb = to_f16_bits(x), to_f16_bits(w)
        x, w = f16_to_f32(xb), f16_to_f32(wb)
        dxb, dwb = DeviceBuffer.from_numpy(xb), DeviceBuffer.from_numpy(wb)
        dw = DeviceBuffer.from_numpy(w)
        lowp = dwb.ptr
        xin = dxb.ptr
    db = DeviceBuffer.from_numpy(b) if b is not None else None
    dres = DeviceBuffer.from_numpy(res) if res is not None else None
    out = DeviceBuffer(M * N)
    lin = ZgLinear(K, N, dw.ptr, db.ptr if db else None)
    L.zg_linear_forward_tc(C.byref(lin), xin, M * K, out.ptr, prec, lowp, epi, dres.ptr if dres else None, bn)
    lib.check()
    err = L.zg_tc_error()
    got = out.download().reshape(M, N)
    want = x.astype(np.float64) @ w.astype(np.float64).T
    if b is not None:
        want = want + b
    if epi == 1:
        want = gelu(want)
    if epi == 2:
        want = want + res
    scale = np.abs(want).max()
    e = float(np.abs(got - want).max() / scale)
    return e, err


def main():
    L = lib.init(0)
    rng = np.random.default_rng(0)
    results = []
    ok_all = True
    cases = [
        # M, N, K, prec, epi, bn
        (128, 256, 64, 0, 0, 256), (128, 256, 64, 1, 0, 256),
        (128, 32, 32, 0, 0, 32), (128, 64, 128, 1, 0, 64), (256, 128, 256, 0, 0, 128),
        (300, 500, 200, 0, 0, 0), (300, 504, 200, 1, 0, 0),
        (16, 50257, 768, 0, 0, 0), (64, 2304, 768, 0, 0, 0), (64, 768, 3072, 0, 2, 0),
        (1024, 3072, 768, 0, 1, 0), (1024, 3072, 768, 1, 1, 0), (2048, 768, 3072, 1, 2, 0),
        (4096, 4800, 1600, 0, 0, 0), (4096, 1600, 6400, 1, 0, 0),
        (1000, 1000, 1000, 0, 0, 64), (1000, 1000, 1000, 0, 0, 32), (1000, 1000, 1000, 1, 2, 128),
        (128, 256, 64, 2, 0, 256), (64, 2304, 768, 2, 0, 0), (64, 768, 3072, 2, 2, 0), (1000, 1000, 1000, 2, 1, 64),
        (1000, 1000, 1000, 2, 0, 128), (1024, 50257, 768, 2, 0, 0), (64, 2304, 768, 3, 0, 0),
    ]
    for (M, N, K, prec, epi, bn) in cases:
        t0 = time.time()
        try:
            e, err = run_case(L, M, N, K, prec, epi, bn, rng)
        except Exception as ex:  # noqa: BLE001
            e, err = float("nan"), str(ex)
        good = (err == 0) and (e < (2e-3 if prec == 3 else 5e-5))
        ok_all &= bool(good)
        results.append(dict(M=M, N=N, K=K, prec=prec, epi=epi, bn=bn, rel_err=e, tc_err=err, ok=bool(good), s=round(time.time() - t0, 2)))
        print(results[-1], flush=True)
        if err not in (0,):
            print("aborting: watchdog / error", flush=True)
            break
    # timing: cfg 3 shapes (355M, M = 16384)
    if ok_all:
        for (M, N, K, prec) in [(16384, 3072, 1024, 1), (16384, 1024, 1024, 1), (16384, 4096, 1024, 1), (16384, 1024, 4096, 1),
                                (16384, 3072, 1024, 0), (16384, 4096, 1024, 0), (8192, 8192, 8192, 1), (8192, 8192, 8192, 0), (64, 4800, 1600, 2), (64, 6400, 1600, 2), (64, 1600, 6400, 2), (1024, 2304, 768, 2), (1024, 50257, 768, 2), (64, 50257, 1600, 2)]:
            es = 2 if prec == 1 else 4
            dx = DeviceBuffer(M * K * es // 4)
            dw = DeviceBuffer(N * K)
            dwl = DeviceBuffer(N * K * es // 4)
            db = DeviceBuffer(N)
            out = DeviceBuffer(M * N)
            lin = ZgLinear(K, N, dw.ptr, db.ptr)
            for _ in range(3):
                L.zg_linear_forward_tc(C.byref(lin), dx.ptr, M * K, out.ptr, prec, dwl.ptr, 0, None, 0)
            L.zg_sync()
            L.zg_timer_begin()
            reps = 10
            for _ in range(reps):
                L.zg_linear_forward_tc(C.byref(lin), dx.ptr, M * K, out.ptr, prec, dwl.ptr, 0, None, 0)
            ms = L.zg_timer_end_ms() / reps
            lib.check()
            tf = 2.0 * M * N * K / (ms * 1e-3) / 1e12
            r = dict(timing=True, M=M, N=N, K=K, prec=prec, ms=round(ms, 4), tflops=round(tf, 1), w_gbs=round(N * K * es / (ms * 1e-3) / 1e9, 1))
            results.append(r)
            print(r, flush=True)
            for b in (dx, dw, dwl, db, out):
                b.free()
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(results, open("gpurun_out/tc_check.json", "w"), indent=1)
    print("ALL OK" if ok_all else "FAILURES")


if __name__ == "__main__":
    main()
