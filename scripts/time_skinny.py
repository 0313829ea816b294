"""Times the swapped-operand stream-K GEMM (zg_linear_forward_skinny) against the general tcgen05 GEMM at the decode-step
shapes of BASELINE cfg 4 (1.5B, M = 64) and cfg 5 (124M, M = 128): us per call and achieved weight-streaming GB/s."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zig_gpt2_b200 import lib  # noqa: E402
from zig_gpt2_b200.lib import DeviceBuffer, ZgLinear  # noqa: E402

L = lib.init(0)
flush = DeviceBuffer(64 * 1024 * 1024)  # 256 MB > L2


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(n):
        L.zg_memset(flush.ptr, 0, flush.len * 4)
        L.zg_timer_begin()
        fn()
        ts.append(L.zg_timer_end_ms())
    return float(np.median(ts)) * 1e3


for M, E in ((64, 1600), (128, 768), (32, 1600)):
    for name, N, K in (("c_attn", 3 * E, E), ("c_proj", E, E), ("c_fc", 4 * E, E), ("mlp c_proj", E, 4 * E), ("lm_head", 50257, E)):
        rs = np.random.RandomState(0)
        dx = DeviceBuffer.from_numpy(rs.randn(M, K).astype(np.float32))
        dw = DeviceBuffer.from_numpy((rs.randn(N, K) * 0.05).astype(np.float32))
        db = DeviceBuffer.from_numpy(rs.randn(N).astype(np.float32))
        out = DeviceBuffer(M * N)
        lin = ZgLinear(K, N, dw.ptr, db.ptr)
        row = f"M={M:4d} {name:11s} N={N:6d} K={K:5d} {N*K*4/1e6:7.1f} MB |"
        for prec, pname in ((0, "tf32"), (2, "3xtf32")):
            t_old = timeit(lambda: L.zg_linear_forward_tc(C.byref(lin), dx.ptr, M * K, out.ptr, prec, None, 0, None, 0))
            t_new = timeit(lambda: L.zg_linear_forward_skinny(C.byref(lin), dx.ptr, M * K, out.ptr, prec, 0, None))
            lib.check()
            row += f" {pname}: old {t_old:7.1f} us  skinny {t_new:7.1f} us ({N*K*4/t_new/1e3:6.0f} GB/s) |"
        print(row, flush=True)
        for b in (dx, dw, db, out):
            b.free()
