"""-m gpu: the swapped-operand stream-K GEMM of the batched decode step (csrc/zg_skinny.cu) against the oracle's Linear
(ops.zig:21-46): plain Linear into a zeroed output, the in-place residual form x += Linear(h) (main.zig:136-145), the
GELU folded into the operand load (main.zig:80), ragged sizes (N not a multiple of 128, M not a multiple of 32), both
precisions.  Tolerances: 3xTF32 <= 1e-4 (fp32 class), TF32 <= 2e-2 of the output scale (north_star)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SHAPES = [(64, 4800, 1600), (64, 1600, 6400), (5, 300, 96), (128, 2304, 768), (100, 1000, 64), (1, 50257, 768), (33, 129, 32)]


def rel(a, b):
    b = np.asarray(b, np.float64)
    return float(np.abs(np.asarray(a, np.float64) - b).max() / np.abs(b).max())


@pytest.mark.parametrize("M,N,K", SHAPES)
@pytest.mark.parametrize("precision,tol", [(2, 1e-4), (0, 2e-2)])
@pytest.mark.parametrize("variant", ["plain", "residual", "gelu", "plain-atomics"])
def test_skinny_linear_matches_oracle(M, N, K, precision, tol, variant):
    import zg_oracle as zo
    from zig_gpt2_b200 import lib
    from zig_gpt2_b200.lib import DeviceBuffer, ZgLinear

    if N > 10000 and (precision != 2 or not variant.startswith("plain")):
        pytest.skip("lm_head shape: one case is enough")
    L = lib.init(0)
    rs = np.random.RandomState(M * 7 + N + K)
    x = rs.randn(M, K).astype(np.float32)
    w = (rs.randn(N, K) * 0.05).astype(np.float32)
    b = rs.randn(N).astype(np.float32)
    r = rs.randn(M, N).astype(np.float32)
    zo.use_openblas()
    xin = zo.gelu(x) if variant == "gelu" else x
    want = zo.linear(xin, w, b)
    zo.use_scalar_blas()
    if variant == "residual":
        want = want + r
    dx, dw, db = DeviceBuffer.from_numpy(x), DeviceBuffer.from_numpy(w), DeviceBuffer.from_numpy(b)
    out = DeviceBuffer.from_numpy(r if variant == "residual" else np.zeros((M, N), np.float32))
    lin = ZgLinear(K, N, dw.ptr, db.ptr)
    xform = 1 if variant == "gelu" else (2 if variant == "plain-atomics" else 0)  # bit 1: scalar-atomic epilogue
    L.zg_linear_forward_skinny(C.byref(lin), dx.ptr, M * K, out.ptr, precision, xform, None)
    lib.check()
    assert L.zg_tc_error() == 0
    e = rel(out.download().reshape(M, N), want)
    assert e <= tol, f"{variant} precision {precision}: {e:.3e} > {tol}"


def test_skinny_refuses_unsupported_shapes():
    from zig_gpt2_b200 import lib
    from zig_gpt2_b200.lib import DeviceBuffer, ZgLinear

    L = lib.init(0)
    buf = DeviceBuffer(256 * 256)
    for M, N, K in ((129, 64, 64), (4, 64, 48)):  # too many rows; in_features not a multiple of 32
        lin = ZgLinear(K, N, buf.ptr, None)
        L.zg_linear_forward_skinny(C.byref(lin), buf.ptr, M * K, buf.ptr, 2, 0, None)
        with pytest.raises(lib.ZgError):
            lib.check()


@pytest.mark.parametrize("M,N,K", [(64, 50257, 768), (7, 1000, 96), (128, 4099, 256)])
def test_fused_argmax_picks_the_first_maximum(M, N, K):
    """Greedy sampling fused into the lm_head GEMM (main.zig:193 + argmax): the token is the oracle's argmax wherever the
    top-2 margin exceeds the 3xTF32 error, and exact ties resolve to the FIRST maximum (duplicated weight rows)."""
    import zg_oracle as zo
    from zig_gpt2_b200 import lib
    from zig_gpt2_b200.lib import DeviceBuffer, ZgLinear

    L = lib.init(0)
    rs = np.random.RandomState(N)
    x = rs.randn(M, K).astype(np.float32)
    w = (rs.randn(N, K) * 0.05).astype(np.float32)
    w[N // 2 + 3] = w[5]  # rows 5 and N/2+3 give bit-identical logits: a tie wherever row 5 wins
    x[0] = w[5] * 40.0    # make row 5 (and its twin) the maximum for batch row 0
    zo.use_openblas()
    logits = zo.linear(x, w, None)
    zo.use_scalar_blas()
    dx, dw = DeviceBuffer.from_numpy(x), DeviceBuffer.from_numpy(w)
    best, tok = DeviceBuffer(2 * M, np.uint64), DeviceBuffer(M, np.uint64)
    lin = ZgLinear(K, N, dw.ptr, None)
    L.zg_linear_argmax_skinny(C.byref(lin), dx.ptr, M * K, 2, best.ptr, tok.ptr)
    lib.check()
    assert L.zg_tc_error() == 0
    got = tok.download().astype(np.int64)
    srt = np.sort(logits, axis=1)
    margin = srt[:, -1] - srt[:, -2]
    want = logits.argmax(axis=1)
    clear = margin > 1e-4 * np.abs(logits).max()
    assert got[0] == 5  # the tie resolves to the first maximum
    assert np.array_equal(got[clear], want[clear])
    picked = logits[np.arange(M), got]
    assert np.all(srt[:, -1] - picked <= 1e-4 * np.abs(logits).max())


@pytest.mark.parametrize("M,N,K", [(64, 4800, 1600), (128, 768, 3072), (9, 200, 64)])
def test_skinny_f16_operands(M, N, K):
    """16-bit weight storage: f16 copies of inputs and weights, kind::f16 MMAs, fp32 accumulation and fp32 output."""
    import zg_oracle as zo
    from zig_gpt2_b200 import lib
    from zig_gpt2_b200.lib import DeviceBuffer, ZgLinear

    L = lib.init(0)
    rs = np.random.RandomState(K)
    x = rs.randn(M, K).astype(np.float32)
    w = (rs.randn(N, K) * 0.05).astype(np.float32)
    b = rs.randn(N).astype(np.float32)
    zo.use_openblas()
    want = zo.linear(x.astype(np.float16).astype(np.float32), w.astype(np.float16).astype(np.float32), b)
    full = zo.linear(x, w, b)
    zo.use_scalar_blas()
    dx, dw, db = DeviceBuffer.from_numpy(x), DeviceBuffer.from_numpy(w), DeviceBuffer.from_numpy(b)
    x16, w16 = DeviceBuffer(M * K, np.uint16), DeviceBuffer(N * K, np.uint16)
    L.zg_to_f16(dx.ptr, x16.ptr, M * K)
    L.zg_to_f16(dw.ptr, w16.ptr, N * K)
    out = DeviceBuffer.from_numpy(np.zeros((M, N), np.float32))
    lin = ZgLinear(K, N, dw.ptr, db.ptr)
    L.zg_linear_forward_skinny(C.byref(lin), x16.ptr, M * K, out.ptr, 1, 0, w16.ptr)
    lib.check()
    assert L.zg_tc_error() == 0
    got = out.download().reshape(M, N)
    assert rel(got, want) <= 1e-4   # against the same f16-rounded operands: only the accumulation order differs
    assert rel(got, full) <= 2e-2   # against the fp32 Linear: the f16 rounding of the operands
