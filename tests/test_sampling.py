"""GPT.sample (main.zig:198-207) with a reproducible, counter-based PRNG.

CPU: Philox4x32-10 known-answer vectors (Random123's kat_vectors) and the (seed, step, sequence) -> uniform mapping.
GPU (-m gpu): the device-resident sampling loops of the batch-1 engine and of the batch engine.  A sampled run is checked
by teacher-forcing the oracle with the run's own tokens: at every sampling step the chosen token must bracket that
step's uniform on the oracle's float64 CDF of softmax(logits / temp) (the reference's weightedIndex), within fp32
summation error -- the same criterion tests/test_gpu_model.py uses for a single draw."""
import ctypes as C

import numpy as np
import pytest

from zig_gpt2_b200.config import GPTConfig

KAT = [  # counter[4], key[2] -> output[4]   (philox4x32_10, Random123 kat_vectors)
    ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


def philox_np(counter, key):
    """numpy restatement of the block function (independent of the C code)."""
    c = [np.uint64(x) for x in counter]
    k = [np.uint64(x) for x in key]
    M0, M1, W0, W1, MASK = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), np.uint64(0x9E3779B9), np.uint64(0xBB67AE85), np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        c = [(p1 >> np.uint64(32)) ^ c[1] ^ k[0], p1 & MASK, (p0 >> np.uint64(32)) ^ c[3] ^ k[1], p0 & MASK]
        k = [(k[0] + W0) & MASK, (k[1] + W1) & MASK]
    return tuple(int(x) for x in c)


def test_philox_known_answers():
    from zig_gpt2_b200 import lib

    L = lib.load()
    for ctr, key, want in KAT:
        out = (C.c_uint * 4)()
        L.zg_philox4x32_10((C.c_uint * 4)(*ctr), (C.c_uint * 2)(*key), out)
        assert tuple(out) == want
        assert philox_np(ctr, key) == want


def test_uniform_is_a_pure_function_of_seed_step_sequence():
    from zig_gpt2_b200 import lib

    L = lib.load()
    seed, step, seq = 0x1234567890abcdef, 7, (5 << 32) | 9
    r = philox_np((step & 0xffffffff, step >> 32, seq & 0xffffffff, seq >> 32), (seed & 0xffffffff, seed >> 32))
    assert L.zg_philox_uniform(seed, step, seq) == (r[0] >> 8) / 16777216.0
    us = np.array([L.zg_philox_uniform(1, s, q) for s in range(64) for q in range(64)])
    assert us.min() >= 0.0 and us.max() < 1.0 and len(set(us)) > 4000 and abs(us.mean() - 0.5) < 0.02
    assert L.zg_philox_uniform(1, 3, 0) != L.zg_philox_uniform(2, 3, 0) != L.zg_philox_uniform(1, 4, 0)


# ---- GPU --------------------------------------------------------------------------------------------------------------
CFG = GPTConfig(vocab_size=4099, context_size=128, n_layer=2, n_heads=4, n_embed=256)


def check_against_oracle(cfg, w, prompt, toks, temp, seed, sequence, L):
    """Teacher-force the oracle with `toks`; every sampled token must bracket its step's uniform on the oracle's CDF."""
    import zg_oracle as zo

    zo.use_scalar_blas()
    m = zo.Model(cfg, w)
    n_in = len(prompt)
    assert list(toks[:n_in]) == list(prompt)
    worst = 0.0
    for s in range(len(toks)):
        if s < n_in:
            m.forward(s + 1, int(toks[s]), False)
            continue
        feed = int(toks[s - 1])  # the last prompt token is forwarded twice (main.zig:329-338)
        logits = m.forward(s + 1, feed, True).astype(np.float64) / temp
        p = np.exp(logits - logits.max())
        cdf = np.cumsum(p) / p.sum()
        u = L.zg_philox_uniform(seed, s, sequence)
        t = int(toks[s])
        lo = cdf[t - 1] if t > 0 else 0.0
        slack = 2e-5  # fp32 softmax + fp32 running sum over 4,099 entries
        assert lo - slack <= u <= cdf[t] + slack, (s, t, lo, u, cdf[t])
        worst = max(worst, max(lo - u, u - cdf[t], 0.0))
    m.close()
    return worst


@pytest.fixture(scope="module")
def small():
    from zig_gpt2_b200 import gpt, lib
    from zig_gpt2_b200.weights import synth_weights

    L = lib.init(0)
    w = synth_weights(CFG, seed=3)
    model = gpt.gpt_from_numpy(CFG, w)
    yield L, w, model
    model.close()


@pytest.mark.gpu
def test_engine_sampling_loop_matches_oracle_cdf_and_is_reproducible(small):
    from zig_gpt2_b200 import gpt

    L, w, model = small
    state = gpt.State(CFG)
    prompt = [11, 4000, 7, 123, 9]
    a = model.generate_sample(prompt, 40, state, temp=0.8, seed=42, sequence=3)
    b = model.generate_sample(prompt, 40, state, temp=0.8, seed=42, sequence=3)
    c = model.generate_sample(prompt, 40, state, temp=0.8, seed=43, sequence=3)
    d = model.generate_sample(prompt, 40, state, temp=0.8, seed=42, sequence=4)
    assert np.array_equal(a, b) and not np.array_equal(a, c) and not np.array_equal(a, d)
    assert len(set(a[len(prompt):].tolist())) > 10  # it samples, it does not collapse to the argmax
    check_against_oracle(CFG, w, prompt, a, 0.8, 42, 3, L)
    greedy = model.generate_greedy(prompt, 40, state)
    cold = model.generate_sample(prompt, 40, state, temp=1e-3, seed=1)  # temperature -> 0 is greedy decoding
    assert np.array_equal(cold, greedy)


@pytest.mark.gpu
@pytest.mark.parametrize("general", [False, True])
def test_batch_sampling_per_sequence_draws(small, general):
    """Batched path: sequence b draws Philox(seed, step, seq_base + b); the same sequences in another batch slot / shard
    (different seq_base arithmetic, same global index) draw the same uniforms."""
    from zig_gpt2_b200.batch import BatchEngine

    L, w, model = small
    B, n_in, n_total = 6, 4, 28
    prompts = np.random.RandomState(5).randint(0, CFG.vocab_size, (B, n_in))
    eng = BatchEngine(model, B, cache_rows=32, general_gemm_only=general)
    a = eng.generate_sample(prompts, n_total, temp=0.9, seed=7, seq_base=100)
    b = eng.generate_sample(prompts, n_total, temp=0.9, seed=7, seq_base=100)
    assert np.array_equal(a, b)
    eng.close()
    for row in range(B):
        check_against_oracle(CFG, w, prompts[row], a[row], 0.9, 7, 100 + row, L)
    # the second half of the batch as its own shard: global sequence ids 103..105
    half = BatchEngine(model, 3, cache_rows=32, general_gemm_only=general)
    h = half.generate_sample(prompts[3:], n_total, temp=0.9, seed=7, seq_base=103)
    half.close()
    for row in range(3):
        check_against_oracle(CFG, w, prompts[3 + row], h[row], 0.9, 7, 103 + row, L)
    assert (h == a[3:]).mean() > 0.9  # identical draws; tokens can differ only at a CDF boundary within rounding
