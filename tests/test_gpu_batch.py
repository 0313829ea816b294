"""-m gpu: the tensor-core paths (BASELINE configs 3-5) against the CPU oracle through the C-ABI.

  * Linear.forward for M >= 16 on tcgen05 (tf32, 3xTF32, fp16 operands) vs the oracle's Linear (ops.zig:21-46);
  * causal prefill attention vs the oracle's incremental CausalSelfAttention.forward -- the reference's own
    KV-cache test (src/tests.zig:245-334: token-at-a-time with a growing cache == causal full-sequence attention);
  * batched single-query attention vs the oracle's scaled_dot_product_attention (ops.zig:249-307);
  * the batch engine: GPT.forward per sequence (logits, KV caches), generate() greedy tokens, batched prefill.

Tolerances (north_star): fp32-class paths <= 1e-4 of the output scale (the 3xTF32 GEMMs accumulate ~1e-4 over 12
layers; stated per test); TF32 / fp16 tensor-core paths <= 2e-2 on logits; greedy tokens identical over 64 tokens.
"""
import ctypes as C

import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from zig_gpt2_b200.config import SIZES, GPTConfig  # noqa: E402

TC_RTOL = 2e-2      # north_star: bf16/TF32 tensor-core paths <= 2e-2 on logits
FP32_RTOL = 1e-4    # north_star: fp32 <= 1e-4 relative


def rel(a, b):
    b = np.asarray(b, np.float64)
    return float(np.abs(np.asarray(a, np.float64) - b).max() / np.abs(b).max())


@pytest.fixture(scope="module")
def L():
    from zig_gpt2_b200 import lib

    return lib.init(0)


@pytest.mark.parametrize("M,N,K", [(16, 3072, 768), (64, 2304, 768), (192, 768, 3072), (300, 1000, 200), (1024, 3072, 768), (1024, 50257, 768)])
@pytest.mark.parametrize("precision,tol", [(2, FP32_RTOL), (0, TC_RTOL), (1, TC_RTOL)])
def test_linear_tensor_core_matches_oracle(L, M, N, K, precision, tol):
    import zg_oracle as zo
    from zig_gpt2_b200 import lib
    from zig_gpt2_b200.lib import DeviceBuffer, ZgLinear

    if N > 10000 and precision != 2:
        pytest.skip("lm_head shape is checked on the decode precision only")
    rs = np.random.RandomState(M + N + K)
    x = rs.randn(M, K).astype(np.float32)
    w = (rs.randn(N, K) * 0.05).astype(np.float32)
    b = rs.randn(N).astype(np.float32)
    zo.use_openblas()
    want = zo.linear(x, w, b)
    zo.use_scalar_blas()
    dx, dw, db, out = DeviceBuffer.from_numpy(x), DeviceBuffer.from_numpy(w), DeviceBuffer.from_numpy(b), DeviceBuffer(M * N)
    lin = ZgLinear(K, N, dw.ptr, db.ptr)
    xin, lowp = dx.ptr, None
    if precision == 1:
        x16, w16 = DeviceBuffer(M * K, np.uint16), DeviceBuffer(N * K, np.uint16)
        L.zg_to_f16(dx.ptr, x16.ptr, M * K)
        L.zg_to_f16(dw.ptr, w16.ptr, N * K)
        xin, lowp = x16.ptr, w16.ptr
    L.zg_linear_forward_tc(C.byref(lin), xin, M * K, out.ptr, precision, lowp, 0, None, 0)
    lib.check()
    assert L.zg_tc_error() == 0
    e = rel(out.download().reshape(M, N), want)
    assert e <= tol, f"precision {precision}: {e:.3e} > {tol}"


@pytest.mark.parametrize("precision", [1, 0])
@pytest.mark.parametrize("M,N,K,epi", [(2000, 2500, 200, 0), (2048, 2560, 1024, 1), (2304, 2304, 136, 2), (16384, 1024, 256, 2)])
def test_linear_cta_pair_kernel_exact_on_representable_operands(L, M, N, K, epi, precision):
    """Wide problems run on CTA pairs (cta_group::2: 256 x 256 tiles, each CTA holds half of the W tile).  Operands that
    are exactly representable in the tensor-core input format (f16 / tf32) make the products exact, so the result must
    match float64 numpy to fp32 accumulation error -- for ragged M, N and K (TMA zero fill, clipped stores), the GELU and
    the in-place residual (TMA reduce-add) epilogues.  The launch counter proves the pair kernel is what ran."""
    from zig_gpt2_b200 import lib
    from zig_gpt2_b200.lib import DeviceBuffer, ZgLinear

    rs = np.random.RandomState(M + N + K + epi)
    x = rs.randn(M, K).astype(np.float32)
    w = (rs.randn(N, K) * 0.05).astype(np.float32)
    b, r = rs.randn(N).astype(np.float32), rs.randn(M, N).astype(np.float32)
    if precision == 1:
        x, w = x.astype(np.float16).astype(np.float32), w.astype(np.float16).astype(np.float32)
    else:
        x, w = ((a.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32) for a in (x, w))
    want = x.astype(np.float64) @ w.astype(np.float64).T + b
    if epi == 1:
        want = 0.5 * want * (1.0 + np.tanh(want * 0.7978845608028654 * (1.0 + 0.044715 * want * want)))
    if epi == 2:
        want = want + r
    dx, dw, db = DeviceBuffer.from_numpy(x), DeviceBuffer.from_numpy(w), DeviceBuffer.from_numpy(b)
    out = DeviceBuffer.from_numpy(r) if epi == 2 else DeviceBuffer(M * N)
    lin = ZgLinear(K, N, dw.ptr, db.ptr)
    xin, lowp = dx.ptr, None
    if precision == 1:
        x16, w16 = DeviceBuffer(M * K, np.uint16), DeviceBuffer(N * K, np.uint16)
        L.zg_to_f16(dx.ptr, x16.ptr, M * K)
        L.zg_to_f16(dw.ptr, w16.ptr, N * K)
        xin, lowp = x16.ptr, w16.ptr
    n0 = L.zg_tc_pair_launch_count()
    L.zg_linear_forward_tc(C.byref(lin), xin, M * K, out.ptr, precision, lowp, epi, out.ptr if epi == 2 else None, 0)
    lib.check()
    assert L.zg_tc_error() == 0
    assert L.zg_tc_pair_launch_count() == n0 + 1, "the CTA-pair kernel did not run"
    assert rel(out.download().reshape(M, N), want) <= 2e-5


def test_linear_forward_routes_large_m_to_tensor_cores(L):
    """The reference's own Linear test shape (tests.zig:22-78: x[3,768], W[3072,768]) scaled to M = 48, through the
    reference-facing zg_linear_forward (ops.Linear.forward)."""
    import zg_oracle as zo
    from zig_gpt2_b200 import lib, ops
    from zig_gpt2_b200.lib import DeviceBuffer

    rs = np.random.RandomState(5)
    x, w, b = rs.randn(48, 768).astype(np.float32), (rs.randn(3072, 768) * 0.05).astype(np.float32), rs.randn(3072).astype(np.float32)
    dw, db = DeviceBuffer.from_numpy(w), DeviceBuffer.from_numpy(b)
    out = DeviceBuffer(48 * 3072)
    n0 = L.zg_launch_count()
    ops.Linear(768, 3072, dw, db).forward(DeviceBuffer.from_numpy(x), out)
    assert L.zg_launch_count() == n0 + 1 and L.zg_tc_error() == 0
    zo.use_openblas()
    assert rel(out.download().reshape(48, 3072), zo.linear(x, w, b)) <= FP32_RTOL
    zo.use_scalar_blas()
    ops.Linear(768, 3072, dw, None).forward(DeviceBuffer.from_numpy(x), out)  # bias == null: beta = 0 (ops.zig:29)
    assert rel(out.download().reshape(48, 3072), zo.linear(x, w, None)) <= FP32_RTOL


@pytest.mark.parametrize("direct", [0, 1])
@pytest.mark.parametrize("epi,inplace,out_cols", [(1, False, 3072), (2, False, 3072), (2, True, 3072), (0, False, 3001)])
def test_linear_fused_epilogues(L, epi, inplace, out_cols, direct):
    """GELU folded into c_fc (main.zig:79-80) and the residual add folded into c_proj (main.zig:136-145), through
    both epilogues: staged TMA stores (an in-place residual becomes a TMA reduce-add) and per-row direct stores.
    N = 3001 exercises the clipped last tile."""
    import zg_oracle as zo
    from zig_gpt2_b200 import lib
    from zig_gpt2_b200.lib import DeviceBuffer, ZgLinear

    rs = np.random.RandomState(epi)
    M, N, K = 130, out_cols, 768
    x, w, b = rs.randn(M, K).astype(np.float32), (rs.randn(N, K) * 0.05).astype(np.float32), rs.randn(N).astype(np.float32)
    r = rs.randn(M, N).astype(np.float32)
    want = zo.linear(x, w, b)
    want = zo.gelu(want) if epi == 1 else (want + r if epi == 2 else want)
    dx, dw, db, dr, out = (DeviceBuffer.from_numpy(a) for a in (x, w, b, r, np.zeros(M * N, np.float32)))
    lin = ZgLinear(K, N, dw.ptr, db.ptr)
    L.zg_tc_set_direct_epilogue(direct)
    try:
        dst = dr if inplace else out
        L.zg_linear_forward_tc(C.byref(lin), dx.ptr, M * K, dst.ptr, 2, None, epi, dr.ptr, 0)
        lib.check()
    finally:
        L.zg_tc_set_direct_epilogue(0)
    assert L.zg_tc_error() == 0
    assert rel(dst.download().reshape(M, N), want) <= FP32_RTOL


@pytest.mark.parametrize("M,N,K,precision,tol", [(64, 1600, 6400, 2, None), (1024, 768, 3072, 2, None), (64, 1600, 1600, 0, 2e-2),
                                                  (200, 768, 768, 2, None)])
def test_linear_split_k_inplace_residual(L, M, N, K, precision, tol):
    """x += Linear(h) (main.zig:136-139,142-145) with few output tiles and a long K: gemm_plan cuts K into slices that
    reduce-add their partial tiles into x (slice 0 adds the bias).  Shapes: 1.5B c_proj / mlp c_proj at batch 64 (cfg 4)
    and 124M mlp c_proj at 1024 rows (cfg 5).  Summation order across slices is not fixed, so fp32-tolerance, not bits."""
    import zg_oracle as zo
    from zig_gpt2_b200 import lib
    from zig_gpt2_b200.lib import DeviceBuffer, ZgLinear

    rs = np.random.RandomState(M + N)
    x = rs.randn(M, K).astype(np.float32)
    w = (rs.randn(N, K) * 0.02).astype(np.float32)
    b, r = rs.randn(N).astype(np.float32), rs.randn(M, N).astype(np.float32)
    want = zo.linear(x, w, b) + r
    dx, dw, db, dr = (DeviceBuffer.from_numpy(a) for a in (x, w, b, r))
    lin = ZgLinear(K, N, dw.ptr, db.ptr)
    L.zg_linear_forward_tc(C.byref(lin), dx.ptr, M * K, dr.ptr, precision, None, 2, dr.ptr, 0)
    lib.check()
    assert L.zg_tc_error() == 0
    assert rel(dr.download().reshape(M, N), want) <= (tol if tol else FP32_RTOL)


@pytest.mark.parametrize("B,T,H", [(1, 5, 12), (2, 128, 2), (2, 200, 3), (1, 1024, 2)])
def test_prefill_attention_equals_incremental_kv_cache_attention(L, B, T, H):
    """tests.zig:245-334: feeding tokens one at a time through the KV cache must equal causal full-sequence attention.
    Here the oracle runs the incremental form (scaled_dot_product_attention over the growing cache, ops.zig:249-307)
    and the GPU runs all positions at once."""
    import zg_oracle as zo
    from zig_gpt2_b200 import lib
    from zig_gpt2_b200.lib import DeviceBuffer

    E = H * 64
    rs = np.random.RandomState(T)
    qkv = rs.randn(B * T, 3 * E).astype(np.float16)
    d_in, d_out = DeviceBuffer.from_numpy(qkv.view(np.uint16)), DeviceBuffer(B * T * E, np.uint16)
    L.zg_attention_prefill(d_in.ptr, d_out.ptr, B, T, H, E)
    lib.check()
    assert L.zg_tc_error() == 0
    got = d_out.download().view(np.float16).astype(np.float32).reshape(B, T, E)
    x = qkv.astype(np.float32).reshape(B, T, 3, H, 64)
    steps = range(T) if T <= 200 else list(range(0, T, 97)) + [T - 1]
    worst = 0.0
    for b in range(B):
        for t in steps:  # the oracle's one-query attention over cache rows [0, t]
            q = np.ascontiguousarray(x[b, t, 0]).reshape(-1)                       # [n, 1, hd]
            k = np.ascontiguousarray(x[b, : t + 1, 1].transpose(1, 0, 2)).reshape(-1)  # [n, T, hd] (ops.zig:153)
            v = np.ascontiguousarray(x[b, : t + 1, 2].transpose(1, 0, 2)).reshape(-1)
            want = zo.sdpa(q, k, v, H, t + 1, 64)
            worst = max(worst, float(np.abs(got[b, t] - want.reshape(-1)).max()))
    scale = float(np.abs(got).max())
    assert worst <= 5e-3 * scale, f"{worst:.3e} vs scale {scale:.3e}"  # fp16 P and fp16 output rounding


def test_batched_decode_attention_matches_oracle_sdpa(L):
    import zg_oracle as zo
    from zig_gpt2_b200 import lib
    from zig_gpt2_b200.lib import DeviceBuffer

    B, Ctx, H, T = 3, 96, 4, 77
    E = H * 64
    rs = np.random.RandomState(2)
    q, k, v = rs.randn(B, E).astype(np.float32), rs.randn(B, Ctx, E).astype(np.float32), rs.randn(B, Ctx, E).astype(np.float32)
    dq, dk, dv, out = DeviceBuffer.from_numpy(q), DeviceBuffer.from_numpy(k), DeviceBuffer.from_numpy(v), DeviceBuffer(B * E)
    L.zg_attention_decode_batch(dq.ptr, dk.ptr, dv.ptr, B, Ctx, H, E, T, out.ptr)
    lib.check()
    got = out.download().reshape(B, E)
    for b in range(B):
        kk = np.ascontiguousarray(k[b, :T].reshape(T, H, 64).transpose(1, 0, 2)).reshape(-1)
        vv = np.ascontiguousarray(v[b, :T].reshape(T, H, 64).transpose(1, 0, 2)).reshape(-1)
        assert rel(got[b], zo.sdpa(q[b], kk, vv, H, T, 64).reshape(-1)) <= 1e-5


def _oracle_runs(cfg, w, prompts, n_total):
    import zg_oracle as zo

    zo.use_openblas()
    toks, logits, kvs = [], [], []
    for p in prompts:
        m = zo.Model(cfg, w)
        t, lg = m.generate_greedy(p, n_total, want_logits=True)
        toks.append(t)
        logits.append(lg)
        kvs.append(m.kv(cfg.n_layer - 1, n_total))
        m.close()
    zo.use_scalar_blas()
    return np.stack(toks), logits, kvs


@pytest.fixture(scope="module")
def small_model():
    from zig_gpt2_b200 import gpt, lib
    from zig_gpt2_b200.weights import synth_weights

    lib.init(0)
    cfg = GPTConfig(vocab_size=4099, context_size=160, n_layer=2, n_heads=4, n_embed=256)
    w = synth_weights(cfg, seed=3)
    model = gpt.gpt_from_numpy(cfg, w)
    yield cfg, w, model
    model.close()


def test_batch_forward_logits_and_caches_match_oracle(small_model):
    """GPT.forward per sequence (main.zig:178-195) for 5 sequences at once, teacher-forced with the oracle's tokens."""
    from zig_gpt2_b200.batch import BatchEngine

    cfg, w, model = small_model
    B, n_in, n_total = 5, 8, 30
    prompts = np.random.RandomState(0).randint(0, cfg.vocab_size, (B, n_in))
    toks, logits, kvs = _oracle_runs(cfg, w, prompts, n_total)
    eng = BatchEngine(model, B, cache_rows=64)
    for s in range(n_total):
        sampling = s >= n_in
        feed = toks[:, s] if not sampling else (toks[:, s - 1] if s > n_in else prompts[:, -1])  # main.zig:329-338
        eng.forward(s + 1, feed, sampling)
        if sampling:
            got = eng.logits()
            for b in range(B):
                assert rel(got[b], logits[b][s - n_in]) <= FP32_RTOL, (s, b)
    k, v = eng.kv(cfg.n_layer - 1, n_total)
    for b in range(B):
        assert rel(k[b], kvs[b][0]) <= FP32_RTOL and rel(v[b], kvs[b][1]) <= FP32_RTOL
    eng.close()


@pytest.mark.parametrize("graph", [True, False])
@pytest.mark.parametrize("general", [False, True])
def test_batch_generate_tokens_identical_to_oracle(small_model, graph, general):
    """generate() per sequence (main.zig:322-342) incl. the duplicated last prompt token; empty-ish and ragged cases:
    prompt of one token, and a batch whose size is not a multiple of anything (B = 3).  Both decode-step
    implementations: the stream-K GEMMs (default for <= 128 sequences) and the general kernel, each with the argmax fused
    into the lm_head epilogue."""
    from zig_gpt2_b200.batch import BatchEngine

    cfg, w, model = small_model
    for B, n_in, n_total in ((3, 1, 20), (5, 8, 72)):
        prompts = np.random.RandomState(B).randint(0, cfg.vocab_size, (B, n_in))
        toks, _, _ = _oracle_runs(cfg, w, prompts, n_total)
        eng = BatchEngine(model, B, cache_rows=n_total, graph=graph, general_gemm_only=general)
        assert eng.fused_argmax  # both kernels carry the argmax in the lm_head epilogue
        got = eng.generate_greedy(prompts, n_total)
        assert np.array_equal(got, toks)
        eng.close()


def test_batch_forward_argmax_only_equals_logits_argmax(small_model):
    """compute_logits = 2 (next-token ids only, argmax in the lm_head epilogue) gives the ids that the logits path gives."""
    from zig_gpt2_b200.batch import BatchEngine

    cfg, w, model = small_model
    B = 9
    toks = np.random.RandomState(4).randint(0, cfg.vocab_size, (6, B))
    a, b = BatchEngine(model, B, cache_rows=16), BatchEngine(model, B, cache_rows=16)
    for s in range(6):
        a.forward(s + 1, toks[s], 1)
        b.forward(s + 1, toks[s], 2)
        want = a.logits().argmax(axis=1)
        assert np.array_equal(a.read_tokens(), want)
        assert np.array_equal(b.read_tokens(), want)
    a.close()
    b.close()


@pytest.mark.parametrize("tf32", [False, True])
def test_batch_argmax_head_full_vocabulary(tf32):
    """The greedy head at the real vocabulary size (50257 rows, not a multiple of any tile): 3xTF32 runs it on the general
    kernel's 128-column tiles, single-pass TF32 on CTA pairs (256 x 256 tiles whose rows past the batch are TMA zero fill),
    both with the argmax in the epilogue.  The ids must be the argmax of the logits the same engine writes."""
    from zig_gpt2_b200 import gpt, lib
    from zig_gpt2_b200.batch import BatchEngine
    from zig_gpt2_b200.weights import synth_weights

    cfg = GPTConfig(SIZES["124M"].vocab_size, 64, 2, 12, 768)
    model = gpt.gpt_from_numpy(cfg, synth_weights(cfg, seed=9))
    B = 5
    L = lib.load()
    toks = np.random.RandomState(8).randint(0, cfg.vocab_size, (4, B))
    a = BatchEngine(model, B, cache_rows=16, tf32_single_pass=tf32)
    b = BatchEngine(model, B, cache_rows=16, tf32_single_pass=tf32)
    n0 = L.zg_tc_pair_launch_count()
    for s in range(4):
        a.forward(s + 1, toks[s], 1)
        b.forward(s + 1, toks[s], 2)
        lg = a.logits()
        want = lg.argmax(axis=1)
        got = b.read_tokens()
        top2 = np.sort(lg, axis=1)[:, -2:]
        clear = (top2[:, 1] - top2[:, 0]) > 1e-3 * np.abs(lg).max()  # the two engines' stream-K layer sums differ in the last bits
        assert np.array_equal(got[clear], want[clear]) and clear.sum() >= B - 1
    assert (L.zg_tc_pair_launch_count() > n0) == tf32
    a.close()
    b.close()
    model.close()


def test_batch_above_128_sequences_uses_the_general_kernel(small_model):
    from zig_gpt2_b200.batch import BatchEngine

    cfg, w, model = small_model
    B, n_in, n_total = 130, 3, 12
    prompts = np.random.RandomState(130).randint(0, cfg.vocab_size, (B, n_in))
    toks, _, _ = _oracle_runs(cfg, w, prompts[:6], n_total)
    eng = BatchEngine(model, B, cache_rows=16)
    got = eng.generate_greedy(prompts, n_total)
    assert np.array_equal(got[:6], toks)
    eng.close()


def test_batch_prefill_matches_oracle(small_model):
    """The batched prefill leaves what T token-at-a-time forwards leave: cache rows [0,T) and last-position logits
    (fp16 tensor-core path: <= 2e-2), and generation continued from it follows the reference loop."""
    from zig_gpt2_b200.batch import BatchEngine

    cfg, w, model = small_model
    B, n_in, n_total = 4, 37, 60
    prompts = np.random.RandomState(7).randint(0, cfg.vocab_size, (B, n_in))
    toks, logits, kvs = _oracle_runs(cfg, w, prompts, n_total)
    eng = BatchEngine(model, B, cache_rows=64, max_prompt=48)
    eng.prefill(prompts, True)
    k, v = eng.kv(cfg.n_layer - 1, n_in)
    import zg_oracle as zo

    zo.use_openblas()
    for b in range(B):
        assert rel(k[b], kvs[b][0][:n_in]) <= TC_RTOL and rel(v[b], kvs[b][1][:n_in]) <= TC_RTOL
        m = zo.Model(cfg, w)  # logits of the last prompt position (what generate_nano_gpt.py:140-141 returns)
        for s in range(n_in):
            lg = m.forward(s + 1, int(prompts[b, s]), s == n_in - 1)
        assert rel(eng.logits()[b], lg) <= TC_RTOL
        m.close()
    zo.use_scalar_blas()
    got = eng.generate_greedy(prompts, n_total, use_prefill=True)
    assert np.array_equal(got[:, :n_in], prompts)
    agree = float((got == toks).mean())
    assert agree >= 0.8, agree  # fp16 prompt pass: tokens may legitimately flip where the top-2 margin is < 2e-2
    eng.close()


_PAIR_PREFILL = r'''
import numpy as np, sys
sys.path.insert(0, ".")
from zig_gpt2_b200 import gpt, lib
from zig_gpt2_b200.batch import BatchEngine
from zig_gpt2_b200.config import GPTConfig
from zig_gpt2_b200.weights import synth_weights
L = lib.init(0)
cfg = GPTConfig(vocab_size=4099, context_size=1024, n_layer=2, n_heads=12, n_embed=768)
model = gpt.gpt_from_numpy(cfg, synth_weights(cfg, seed=5))
B, T = 8, 1024
prompts = np.random.RandomState(3).randint(0, cfg.vocab_size, (B, T))
eng = BatchEngine(model, B, cache_rows=T, max_prompt=T)
n0 = L.zg_tc_pair_launch_count()
eng.prefill(prompts, True)
lib.check()
k, v = eng.kv(cfg.n_layer - 1, T)
np.savez(sys.argv[1], logits=eng.logits(), k=k, v=v, pair_launches=int(L.zg_tc_pair_launch_count() - n0))
print("prefill ok")
'''


def test_prefill_cta_pair_gemms_equal_single_cta_gemms():
    """The f16 prefill at 8 x 1024 rows runs its GEMMs on CTA pairs (c_attn with the fused K/V cache append, c_fc + GELU,
    the two in-place residual projections); ZG_NO_PAIR=1 runs the same plans on the single-CTA kernel.  Both accumulate
    every output element over K in the same order, so the last layer's caches and the logits must agree to fp32 rounding
    of the reduce-add epilogue (<= 1e-5), and the launch counter says which kernel ran."""
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = []
    for tag, env in (("pair", {}), ("single", {"ZG_NO_PAIR": "1"})):
        out = os.path.join(root, "gpurun_out", f"_prefill_{tag}.npz")
        os.makedirs(os.path.dirname(out), exist_ok=True)
        r = subprocess.run([sys.executable, "-c", _PAIR_PREFILL, out], cwd=root, capture_output=True, text=True, timeout=600,
                           env={**os.environ, **env})
        assert r.returncode == 0 and "prefill ok" in r.stdout, r.stderr[-1500:]
        outs.append(dict(np.load(out)))
        os.remove(out)
    a, b = outs
    assert int(a["pair_launches"]) >= 8 and int(b["pair_launches"]) == 0
    assert np.isfinite(a["logits"]).all()
    for key in ("logits", "k", "v"):
        assert rel(a[key], b[key]) <= 1e-5, key


def test_batch_engine_124m_64_greedy_tokens_identical(weights_124m):
    """north_star: identical greedy token sequences over the first 64 tokens (124M, fp32-class 3xTF32 decode path)."""
    from zig_gpt2_b200 import gpt
    from zig_gpt2_b200.batch import BatchEngine

    cfg = SIZES["124M"]
    model = gpt.gpt_from_numpy(cfg, weights_124m)
    B, n_in, n_total = 4, 16, 80
    prompts = np.random.RandomState(11).randint(0, cfg.vocab_size, (B, n_in))
    toks, logits, _ = _oracle_runs(cfg, weights_124m, prompts, n_total)
    eng = BatchEngine(model, B, cache_rows=128, max_prompt=16)
    got = eng.generate_greedy(prompts, n_total)
    assert np.array_equal(got, toks)
    eng.prefill(prompts, True)  # fp16 prefill logits vs the oracle's logits at the first sampling step's predecessor
    eng.forward(n_in + 1, prompts[:, -1], True)  # duplicate-last-token step on top of the prefilled caches
    lg = eng.logits()
    for b in range(B):
        assert rel(lg[b], logits[b][0]) <= TC_RTOL
    eng.close()
    model.close()


@pytest.mark.parametrize("size,n_layer", [("355M", 2), ("1.5B", 2)])
def test_batch_other_widths_truncated_depth(size, n_layer):
    """E = 1024 / 1600 (H = 16 / 25; 1600 is not a multiple of 128) at reduced depth, decode + prefill."""
    from zig_gpt2_b200 import gpt
    from zig_gpt2_b200.batch import BatchEngine
    from zig_gpt2_b200.weights import synth_weights

    full = SIZES[size]
    cfg = GPTConfig(full.vocab_size, 256, n_layer, full.n_heads, full.n_embed)
    w = synth_weights(cfg, seed=77)
    model = gpt.gpt_from_numpy(cfg, w)
    B, n_in, n_total = 3, 9, 24
    prompts = np.random.RandomState(1).randint(0, cfg.vocab_size, (B, n_in))
    toks, logits, kvs = _oracle_runs(cfg, w, prompts, n_total)
    eng = BatchEngine(model, B, cache_rows=32, max_prompt=16)
    assert np.array_equal(eng.generate_greedy(prompts, n_total), toks)
    eng.prefill(prompts, False)
    k, v = eng.kv(n_layer - 1, n_in)
    for b in range(B):
        assert rel(k[b], kvs[b][0][:n_in]) <= TC_RTOL and rel(v[b], kvs[b][1][:n_in]) <= TC_RTOL
    eng.close()
    model.close()


def test_batch_decode_loop_does_not_allocate(small_model):
    """No allocation, graph instantiation or tensor-map encode once the engine has run its first step (plans are
    pre-encoded at zg_batch_create, the step graph is captured on first use)."""
    from zig_gpt2_b200 import lib
    from zig_gpt2_b200.batch import BatchEngine

    cfg, w, model = small_model
    L = lib.load()
    B = 6
    prompts = np.random.RandomState(3).randint(0, cfg.vocab_size, (B, 5))
    eng = BatchEngine(model, B, cache_rows=64)
    eng.generate_greedy(prompts, 12)  # captures the prompt / sampling graphs
    before = L.zg_alloc_count()
    eng.generate_greedy(prompts, 40)
    eng.forward(9, prompts[:, 0], True)
    eng.set_position(9)
    eng.run_steps(6)
    L.zg_sync()
    lib.check()
    assert L.zg_alloc_count() == before
    eng.close()


def test_batch_generate_is_repeatable_and_bit_exact_without_split_k(small_model):
    """Split-K (TMA reduce-adds in arrival order) makes the two in-place residual GEMMs of a decode step
    run-to-run non-deterministic in the last bits: tokens must still repeat, logits agree to fp32 tolerance.  With
    ZG_NO_SPLIT_K=1 (a child process: the switch is read once) the logits are bit-identical run to run."""
    import os
    import subprocess
    import sys

    from zig_gpt2_b200.batch import BatchEngine

    cfg, w, model = small_model
    B, n_in, n_total = 7, 6, 48
    prompts = np.random.RandomState(21).randint(0, cfg.vocab_size, (B, n_in))
    eng = BatchEngine(model, B, cache_rows=64)
    a = eng.generate_greedy(prompts, n_total)
    b = eng.generate_greedy(prompts, n_total)
    assert np.array_equal(a, b)
    eng.forward(n_total, a[:, -1], 1)  # greedy steps fuse the argmax and write no logits: ask for them explicitly
    la = eng.logits().copy()
    eng.forward(n_total, a[:, -1], 1)
    lb = eng.logits().copy()
    eng.close()
    assert rel(la, lb) <= FP32_RTOL

    code = r'''
import numpy as np, sys
sys.path.insert(0, ".")
from zig_gpt2_b200 import gpt, lib
from zig_gpt2_b200.batch import BatchEngine
from zig_gpt2_b200.config import GPTConfig
from zig_gpt2_b200.weights import synth_weights
lib.init(0)
cfg = GPTConfig(vocab_size=4099, context_size=160, n_layer=2, n_heads=4, n_embed=256)
model = gpt.gpt_from_numpy(cfg, synth_weights(cfg, seed=3))
prompts = np.random.RandomState(21).randint(0, cfg.vocab_size, (7, 6))
eng = BatchEngine(model, 7, cache_rows=64)
outs = []  # ZG_NO_SPLIT_K=1: general kernel, no floating-point reduction in arrival order anywhere
for _ in range(3):
    t = eng.generate_greedy(prompts, 48)
    eng.forward(48, t[:, -1], 1)  # greedy steps write no logits: ask for them explicitly
    outs.append((t, eng.logits().copy()))
assert all(np.array_equal(outs[0][0], o[0]) and np.array_equal(outs[0][1], o[1]) for o in outs[1:]), "not bit-exact"
np.save(sys.argv[1], outs[0][0])
print("deterministic ok")
'''
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = os.path.join(root, "gpurun_out", "_nosplit_tokens.npy")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    r = subprocess.run([sys.executable, "-c", code, out], cwd=root, capture_output=True, text=True, timeout=600,
                       env={**os.environ, "ZG_NO_SPLIT_K": "1"})
    assert r.returncode == 0 and "deterministic ok" in r.stdout, r.stderr[-1500:]
    assert np.array_equal(np.load(out), a)  # and the two reduction orders pick the same tokens
    os.remove(out)


def test_exact_prefill_generate_124m_64_tokens_identical(weights_124m):
    """north_star "identical greedy tokens over the first 64" for the PREFILLED product path: 124M, B = 4, the whole prompt
    in one fp32-class prefill pass (3xTF32 GEMMs + fp32 causal attention), then 64 greedy tokens -- identical to the
    oracle's token-at-a-time loop.  The f16 prefill (the fast default) is held to a margin-aware criterion instead: it may
    leave the oracle's token sequence only at a step whose top-2 logit margin is inside the 2e-2 tensor-core tolerance."""
    from zig_gpt2_b200 import gpt
    from zig_gpt2_b200.batch import BatchEngine

    cfg = SIZES["124M"]
    model = gpt.gpt_from_numpy(cfg, weights_124m)
    B, n_in, n_total = 4, 16, 80
    prompts = np.random.RandomState(11).randint(0, cfg.vocab_size, (B, n_in))
    toks, logits, kvs = _oracle_runs(cfg, weights_124m, prompts, n_total)
    eng = BatchEngine(model, B, cache_rows=128, max_prompt=16, exact_prefill=True)
    got = eng.generate_greedy(prompts, n_total, use_prefill=True)
    assert np.array_equal(got, toks)
    eng.prefill(prompts, True)  # caches and last-position logits of the exact prefill, fp32 tolerance
    k, v = eng.kv(cfg.n_layer - 1, n_in)
    for b in range(B):
        assert rel(k[b], kvs[b][0][:n_in]) <= FP32_RTOL and rel(v[b], kvs[b][1][:n_in]) <= FP32_RTOL
    eng.close()

    fast = BatchEngine(model, B, cache_rows=128, max_prompt=16)
    got16 = fast.generate_greedy(prompts, n_total, use_prefill=True)
    fast.close()
    for b in range(B):
        diff = np.nonzero(got16[b] != toks[b])[0]
        if diff.size:  # first departure from the oracle's sequence: only where the oracle itself was nearly tied
            s = int(diff[0])
            assert s >= n_in
            lg = np.sort(logits[b][s - n_in])
            assert lg[-1] - lg[-2] <= TC_RTOL * float(np.abs(lg).max()), (b, s, lg[-1] - lg[-2])
    model.close()


def test_exact_prefill_small_model_ragged(small_model):
    """Exact prefill on the 2-layer model: prompt length that is not a multiple of anything, 5 sequences; tokens identical
    to the oracle and to this engine's own token-at-a-time prompt loop."""
    from zig_gpt2_b200.batch import BatchEngine

    cfg, w, model = small_model
    B, n_in, n_total = 5, 37, 70
    prompts = np.random.RandomState(9).randint(0, cfg.vocab_size, (B, n_in))
    toks, _, _ = _oracle_runs(cfg, w, prompts, n_total)
    eng = BatchEngine(model, B, cache_rows=72, max_prompt=40, exact_prefill=True)
    assert np.array_equal(eng.generate_greedy(prompts, n_total, use_prefill=True), toks)
    assert np.array_equal(eng.generate_greedy(prompts, n_total, use_prefill=False), toks)
    eng.close()


def test_16bit_storage_decode_step(small_model):
    """16-bit weight / KV storage for the batched decode step: logits and caches within the tensor-core
    tolerance (<= 2e-2, north_star) of the fp32 oracle, greedy tokens leave the oracle's sequence only at a near-tie."""
    from zig_gpt2_b200.batch import BatchEngine

    cfg, w, model = small_model
    B, n_in, n_total = 5, 8, 40
    prompts = np.random.RandomState(0).randint(0, cfg.vocab_size, (B, n_in))
    toks, logits, kvs = _oracle_runs(cfg, w, prompts, n_total)
    eng = BatchEngine(model, B, cache_rows=64, storage16=True)
    assert eng.storage_bits == 16 and eng.fused_argmax
    for s in range(n_total):  # teacher-forced with the oracle's tokens
        sampling = s >= n_in
        feed = toks[:, s] if not sampling else (toks[:, s - 1] if s > n_in else prompts[:, -1])
        eng.forward(s + 1, feed, 1 if sampling else 0)
        if sampling:
            got = eng.logits()
            for b in range(B):
                assert rel(got[b], logits[b][s - n_in]) <= TC_RTOL, (s, b)
            assert np.array_equal(eng.read_tokens(), got.argmax(axis=1))
    k, v = eng.kv(cfg.n_layer - 1, n_total)
    for b in range(B):
        assert rel(k[b].astype(np.float32), kvs[b][0]) <= TC_RTOL and rel(v[b].astype(np.float32), kvs[b][1]) <= TC_RTOL
    got = eng.generate_greedy(prompts, n_total)
    eng.close()
    assert np.array_equal(got[:, :n_in], prompts)
    for b in range(B):
        diff = np.nonzero(got[b] != toks[b])[0]
        if diff.size:
            s = int(diff[0])
            lg = np.sort(logits[b][s - n_in])
            assert lg[-1] - lg[-2] <= TC_RTOL * float(np.abs(lg).max()), (b, s)


def test_16bit_storage_rejects_unsupported_shapes(small_model):
    from zig_gpt2_b200 import lib
    from zig_gpt2_b200.batch import BatchEngine

    cfg, w, model = small_model
    with pytest.raises(lib.ZgError):
        BatchEngine(model, 129, cache_rows=16, storage16=True)
    with pytest.raises(lib.ZgError):
        BatchEngine(model, 4, cache_rows=16, max_prompt=8, storage16=True)
