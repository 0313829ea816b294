"""-m gpu: main.zig's MLP / Block / GPT.forward / GPT.sample / generate on the GPU (fused persistent
engine and op-by-op composition) against the CPU oracle on identical synthetic weights and inputs, and
against the committed golden logits produced by the reference's PyTorch model."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from zig_gpt2_b200.config import SIZES, GPTConfig  # noqa: E402

FP32_RTOL = 1e-4  # north_star: fp32 <= 1e-4 relative (to the logit scale)


@pytest.fixture(scope="module")
def gpu_124m(weights_124m):
    from zig_gpt2_b200 import gpt, lib

    lib.init(0)
    cfg = SIZES["124M"]
    model = gpt.gpt_from_numpy(cfg, weights_124m)
    state = gpt.State(cfg)
    yield model, state
    model.close()


@pytest.fixture(scope="module")
def oracle_124m(weights_124m):
    import zg_oracle as zo

    zo.use_openblas()
    m = zo.Model(SIZES["124M"], weights_124m)
    yield m
    m.close()
    zo.use_scalar_blas()


def close(a, b, what, rtol=FP32_RTOL):
    scale = float(np.abs(b).max())
    err = float(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)).max())
    assert err <= rtol * scale, f"{what}: max abs err {err:.3e} > {rtol} * scale {scale:.3e}"


@pytest.mark.parametrize("fused", [False, True])
def test_prompt_logits_match_reference_torch_golden(gpu_124m, gpt_golden, fused):
    model, state = gpu_124m
    p = gpt_golden["prompt"]
    fwd = model.forward if fused else model.forward_unfused
    for s, tok in enumerate(p):
        fwd(s + 1, int(tok), s == len(p) - 1, state)
    close(state.logits.download(), gpt_golden["prompt_logits"], f"prompt logits fused={fused}")


@pytest.mark.parametrize("fused", [False, True])
def test_forward_state_matches_oracle(gpu_124m, oracle_124m, gpt_golden, fused):
    """state.x (ln_f output), state.o (residual stream), logits and the KV cache rows after each step."""
    model, state = gpu_124m
    p = gpt_golden["prompt"][:6]
    fwd = model.forward if fused else model.forward_unfused
    for s, tok in enumerate(p):
        fwd(s + 1, int(tok), True, state)
        ref_logits = oracle_124m.forward(s + 1, int(tok), True)
        close(state.logits.download(), ref_logits, f"logits step {s}")
        close(state.x.download(), oracle_124m.x(), f"state.x step {s}")
    for layer in (0, 5, 11):
        k_ref, v_ref = oracle_124m.kv(layer, len(p))
        E = 768
        close(model.h[layer].k_cache.download(len(p) * E).reshape(len(p), E), k_ref, f"k_cache layer {layer}")
        close(model.h[layer].v_cache.download(len(p) * E).reshape(len(p), E), v_ref, f"v_cache layer {layer}")


def test_generate_greedy_64_tokens_identical_to_oracle(gpu_124m, oracle_124m, gpt_golden):
    """BASELINE cfg 1 vs cfg 2: 16-token prompt, 64 greedy tokens, following generate()'s loop exactly
    (prompt one token at a time without logits, duplicate last prompt token)."""
    model, state = gpu_124m
    p = gpt_golden["prompt"]
    n_total = len(p) + 64
    ref, ref_logits = oracle_124m.generate_greedy(p, n_total, want_logits=True)
    got = model.generate_greedy(p, n_total, state)
    srt = np.sort(ref_logits, axis=1)
    margins = srt[:, -1] - srt[:, -2]
    assert np.array_equal(got[: len(p)], p)
    assert np.array_equal(got, ref), f"first mismatch at {int(np.argmax(got != ref))}; min top-2 margin {margins.min():.3e}"
    assert np.array_equal(got[len(p): len(p) + 12], gpt_golden["greedy_tokens"])  # the reference torch model agrees too


def test_step_by_step_sampling_matches_single_launch(gpu_124m, gpt_golden):
    model, state = gpu_124m
    p = [int(t) for t in gpt_golden["prompt"]]
    one = model.generate_greedy(p, len(p) + 8, state)
    toks, token = [], 0
    for s in range(len(p) + 8):
        if s < len(p):
            token = p[s]
            model.forward(s + 1, token, False, state)
        else:
            token = model.sample_greedy(s + 1, token, state)
        toks.append(token)
    assert toks == [int(t) for t in one]
    toks2, token = [], 0
    for s in range(len(p) + 8):
        if s < len(p):
            token = p[s]
            model.forward_unfused(s + 1, token, False, state)
        else:
            token = model.sample_greedy(s + 1, token, state, fused=False)
        toks2.append(token)
    assert toks2 == toks


def test_temperature_sampling_matches_oracle_inverse_cdf(gpu_124m, oracle_124m, gpt_golden):
    """GPT.sample (main.zig:198-207): temperature softmax + weightedIndex.  The draw `u` is explicit.  The
    device scans the CDF in parallel chunks, the reference sequentially in fp32, so the chosen index must
    bracket u on the float64 CDF of the oracle's probabilities within fp32 summation error."""
    import zg_oracle as zo

    model, state = gpu_124m
    p = [int(t) for t in gpt_golden["prompt"][:5]]
    for s, tok in enumerate(p[:4]):
        model.forward(s + 1, tok, False, state)
        oracle_124m.forward(s + 1, tok, False)
    logits = oracle_124m.forward(5, p[4], True)
    probs = zo.softmax(logits / np.float32(0.8)).astype(np.float64)
    cdf = np.cumsum(probs) / probs.sum()
    tol = 2e-5
    for u in (0.0, 0.3, 0.77, 0.999):
        got = model.sample(5, 0.8, p[4], state, u)
        want = oracle_124m.sample(5, 0.8, p[4], u)
        assert cdf[got] >= u - tol and (got == 0 or cdf[got - 1] <= u + tol), (u, got, want)
        assert abs(got - want) <= 8, (u, got, want)


def test_long_context_attention_splits(weights_124m):
    """Flash-decoding split path (T > 128): a 2-layer slice of the 124M model run to T = 300."""
    import zg_oracle as zo
    from zig_gpt2_b200 import gpt

    cfg = GPTConfig(50257, 1024, 2, 12, 768)
    model = gpt.gpt_from_numpy(cfg, weights_124m)
    state = gpt.State(cfg)
    zo.use_openblas()
    orc = zo.Model(cfg, weights_124m)
    rs = np.random.RandomState(9)
    prompt = rs.randint(0, cfg.vocab_size, 300)
    got = model.generate_greedy(prompt, 310, state)
    ref = orc.generate_greedy(prompt, 310)
    assert np.array_equal(got, ref)
    close(model.h[1].k_cache.download(310 * 768), orc.kv(1, 310)[0].reshape(-1), "k_cache at T=310")
    model.close()
    orc.close()


def test_full_context_1024_tokens_identical_to_oracle(weights_124m):
    """Maximum size: generate() runs to context_size (main.zig:336).  2-layer slice of the 124M model, a 1000-token
    prompt, then greedy tokens up to position 1023: every attention split count (1..8) and the multi-round path
    (more than 112 rows per split) against the oracle, tokens and the last layer's K/V cache."""
    import zg_oracle as zo
    from zig_gpt2_b200 import gpt

    cfg = GPTConfig(50257, 1024, 2, 12, 768)
    model = gpt.gpt_from_numpy(cfg, weights_124m)
    state = gpt.State(cfg)
    zo.use_openblas()
    orc = zo.Model(cfg, weights_124m)
    rs = np.random.RandomState(11)
    prompt = rs.randint(0, cfg.vocab_size, 1000)
    got = model.generate_greedy(prompt, 1024, state)
    ref = orc.generate_greedy(prompt, 1024)
    assert np.array_equal(got, ref)
    k, v = orc.kv(1, 1024)
    close(model.h[1].k_cache.download(1024 * 768), k.reshape(-1), "k_cache at T=1024")
    close(model.h[1].v_cache.download(1024 * 768), v.reshape(-1), "v_cache at T=1024")
    model.close()
    orc.close()


def test_generate_edge_cases(gpu_124m, oracle_124m, gpt_golden):
    """Ragged ends of generate(): a one-token prompt, a prompt that fills the whole request (no token is sampled, the
    output is the prompt), and a request longer than the context (refused, nothing written past the buffer)."""
    model, state = gpu_124m
    p = [int(t) for t in gpt_golden["prompt"]]
    one = model.generate_greedy(p[:1], 6, state)
    assert np.array_equal(one, oracle_124m.generate_greedy(p[:1], 6))
    same = model.generate_greedy(p, len(p), state)
    assert np.array_equal(same, np.asarray(p))
    from zig_gpt2_b200 import lib

    with pytest.raises(lib.ZgError):
        model.generate_greedy(p, model.config.context_size + 1, state)
    lib.load().zg_clear_error()
    again = model.generate_greedy(p, len(p) + 4, state)  # the engine is still usable after a refused call
    assert np.array_equal(again, oracle_124m.generate_greedy(p, len(p) + 4))


@pytest.mark.parametrize("size,n_layer", [("355M", 3), ("774M", 2), ("1.5B", 2)])
def test_other_widths_truncated_depth(size, n_layer):
    """E = 1024 / 1280 / 1600 (H = 16 / 20 / 25) with the layer count cut down so the oracle stays fast."""
    import zg_oracle as zo
    from zig_gpt2_b200 import gpt
    from zig_gpt2_b200.weights import synth_weights

    full = SIZES[size]
    cfg = GPTConfig(full.vocab_size, full.context_size, n_layer, full.n_heads, full.n_embed)
    w = synth_weights(cfg, seed=77)
    model = gpt.gpt_from_numpy(cfg, w)
    state = gpt.State(cfg)
    zo.use_openblas()
    orc = zo.Model(cfg, w)
    prompt = np.random.RandomState(1).randint(0, cfg.vocab_size, 9)
    ref, ref_logits = orc.generate_greedy(prompt, 20, want_logits=True)
    got = model.generate_greedy(prompt, 20, state)
    assert np.array_equal(got, ref)
    model.forward(20 + 1, int(got[-1]), True, state)
    close(state.logits.download(), orc.forward(21, int(ref[-1]), True), f"{size} logits")
    model.close()
    orc.close()


def test_no_allocation_in_the_decode_loop(gpu_124m, gpt_golden):
    """README.md "No memory allocations at runtime": once the engine exists, generate / run_steps / forward / sample
    must not allocate (cudaMalloc, cudaHostAlloc, graph instantiation) or encode a tensor map -- zg_alloc_count is flat."""
    from zig_gpt2_b200 import lib

    model, state = gpu_124m
    L = lib.load()
    p = [int(t) for t in gpt_golden["prompt"]]
    model.generate_greedy(p, len(p) + 2, state)  # engine exists from here on
    before, launches = L.zg_alloc_count(), L.zg_launch_count()
    model.generate_greedy(p, len(p) + 24, state)
    eng = model.engine(state)
    L.zg_engine_run_steps(eng, len(p), 8)
    model.forward(5, 17, True, state)
    model.sample_greedy(6, 17, state)
    model.sample(7, 0.8, 17, state, u=0.25)
    L.zg_sync()
    lib.check()
    assert L.zg_alloc_count() == before
    assert L.zg_launch_count() > launches


def test_bad_positions_and_tokens_are_rejected_on_both_paths(gpu_124m):
    """The op-by-op GPT.forward and the fused engine refuse seq_len outside [1, context_size] and token >= vocab_size
    with the same sticky error instead of indexing wte / wpe / the caches out of bounds."""
    from zig_gpt2_b200 import lib

    model, state = gpu_124m
    cfg = model.config
    for fwd in (model.forward, model.forward_unfused):
        for seq_len, token in ((0, 1), (cfg.context_size + 1, 1), (3, cfg.vocab_size), (3, 2**40)):
            with pytest.raises(lib.ZgError):
                fwd(seq_len, token, True, state)
    model.forward(1, 0, True, state)  # the engine still works afterwards
    assert np.isfinite(state.logits.download()).all()


def test_shutdown_and_reinit_contract():
    """zg_shutdown + zg_init starts clean: per-device lazily created state (tensor-core watchdog word, max-dynamic-smem
    attributes, timer events, the __constant__ layer-table owner) is forgotten, and results are unchanged.  Runs in a
    child process so that the module fixtures of this session keep their context."""
    import subprocess
    import sys

    code = r'''
import ctypes as C, numpy as np, sys
sys.path.insert(0, ".")
from zig_gpt2_b200 import gpt, lib
from zig_gpt2_b200.config import GPTConfig
from zig_gpt2_b200.lib import DeviceBuffer, ZgLinear
from zig_gpt2_b200.weights import synth_weights
cfg = GPTConfig(vocab_size=1031, context_size=64, n_layer=2, n_heads=4, n_embed=256)
w = synth_weights(cfg, seed=5)
rs = np.random.RandomState(0)
x, wm = rs.randn(64, 256).astype(np.float32), rs.randn(384, 256).astype(np.float32)  # M >= 16: tensor-core path
def once():
    L = lib.init(0)
    model, state = gpt.gpt_from_numpy(cfg, w), gpt.State(cfg)
    toks = model.generate_greedy([1, 2, 3], 20, state)
    dx, dw, out = DeviceBuffer.from_numpy(x), DeviceBuffer.from_numpy(wm), DeviceBuffer(64 * 384)
    lin = ZgLinear(256, 384, dw.ptr, None)
    L.zg_linear_forward_tc(C.byref(lin), dx.ptr, x.size, out.ptr, 2, None, 0, None, 0)
    lib.check()
    assert L.zg_tc_error() == 0
    L.zg_timer_begin(); ms = L.zg_timer_end_ms(); assert ms >= 0
    y = out.download()
    model.close()
    for b in (dx, dw, out): b.free()
    return toks, y
t1, y1 = once()
L = lib.load()
assert L.zg_shutdown() == 0
lib._inited_device = None
t2, y2 = once()
assert np.array_equal(t1, t2) and np.array_equal(y1, y2)
assert L.zg_shutdown() == 0 and L.zg_shutdown() == 0   # idempotent
L.zg_gelu(None, 4)
assert L.zg_last_error() != 0                             # not initialised: sticky error, no CPU fallback
print("reinit ok")
'''
    import os

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", code], cwd=root, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "reinit ok" in r.stdout, r.stderr[-1500:]
