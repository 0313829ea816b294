"""3xTF32 error of the general and the stream-K GEMM against float64 numpy (A/B of the operand split)."""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zig_gpt2_b200 import lib
from zig_gpt2_b200.lib import DeviceBuffer, ZgLinear
L = lib.init(0)
rng = np.random.default_rng(0)
for M, N, K, which in [(1024, 768, 3072, "tc"), (1024, 3072, 768, "tc"), (64, 1600, 6400, "skinny"), (64, 6400, 1600, "skinny"), (64, 1600, 6400, "skinny_gelu")]:
    x = rng.standard_normal((M, K)).astype(np.float32)
    w = (rng.standard_normal((N, K)) * 0.05).astype(np.float32)
    b = rng.standard_normal(N).astype(np.float32)
    dx, dw, db = DeviceBuffer.from_numpy(x), DeviceBuffer.from_numpy(w), DeviceBuffer.from_numpy(b)
    out = DeviceBuffer(M * N)
    lin = ZgLinear(K, N, dw.ptr, db.ptr)
    xr = x.astype(np.float64)
    if which == "tc":
        L.zg_linear_forward_tc(C.byref(lin), dx.ptr, M * K, out.ptr, 2, None, 0, None, 0)
    else:
        xf = 1 if which.endswith("gelu") else 0
        if xf:
            xr = 0.5 * xr * (1.0 + np.tanh(xr * 0.7978845608028654 * (1.0 + 0.044715 * xr * xr)))
        L.zg_linear_forward_skinny(C.byref(lin), dx.ptr, M * K, out.ptr, 2, xf, None)
    lib.check()
    got = out.download().reshape(M, N).astype(np.float64)
    ref = xr @ w.astype(np.float64).T + b
    print(which, M, N, K, "max err / scale = %.3e" % (np.abs(got - ref).max() / np.abs(ref).max()), flush=True)
