"""Pins the oracle's main.zig restatement (State/MLP/Block/GPT/generate) against the reference's
own PyTorch model (generate_nano_gpt.py:24-152, executed from /root/reference when the golden
was made): an independent formulation (full-sequence masked attention, no KV cache)."""
import numpy as np
import pytest

import zg_oracle as zo
from zig_gpt2_b200.config import SIZES


@pytest.fixture(scope="module")
def model(weights_124m):
    zo.use_openblas()
    m = zo.Model(SIZES["124M"], weights_124m)
    yield m
    m.close()
    zo.use_scalar_blas()


def test_prompt_logits_match_reference_torch_model(model, gpt_golden):
    p = gpt_golden["prompt"]
    logits = None
    for s, tok in enumerate(p):  # main.zig:330-334, one token at a time
        logits = model.forward(s + 1, int(tok), compute_logits=(s == len(p) - 1))
    ref = gpt_golden["prompt_logits"]
    # fp32 <= 1e-4 relative (north_star) -- relative to the logit scale
    assert np.abs(logits - ref).max() <= 1e-4 * np.abs(ref).max()
    assert int(np.argmax(logits)) == int(np.argmax(ref))


def test_generate_greedy_reproduces_duplicate_last_prompt_token(model, gpt_golden):
    """main.zig:329-338: the first sampled step re-forwards the last prompt token at the next
    position.  The golden was produced by running the torch model on prompt + [prompt[-1]] + ..."""
    p = gpt_golden["prompt"]
    n_new = len(gpt_golden["greedy_tokens"])
    toks, logits = model.generate_greedy(p, len(p) + n_new, want_logits=True)
    assert np.array_equal(toks[: len(p)], p)
    assert np.array_equal(toks[len(p):], gpt_golden["greedy_tokens"])
    scale = np.abs(gpt_golden["greedy_logits_first"]).max()
    assert np.abs(logits[0] - gpt_golden["greedy_logits_first"]).max() <= 1e-4 * scale
    assert np.abs(logits[-1] - gpt_golden["greedy_logits_last"]).max() <= 1e-4 * scale
    top = gpt_golden["greedy_top4"]
    for s in range(n_new):
        assert int(top[s, 1, 0]) == int(toks[len(p) + s])
        np.testing.assert_allclose(np.sort(logits[s])[::-1][:4], top[s, 0], rtol=0, atol=1e-4 * scale)


def test_sample_is_inverse_cdf_of_temperature_softmax(model, gpt_golden):
    """main.zig:198-207 with the wall-clock PRNG draw made explicit."""
    p = gpt_golden["prompt"]
    for s, tok in enumerate(p[:4]):
        model.forward(s + 1, int(tok), compute_logits=False)
    base = model.forward(5, int(p[4]))
    probs = zo.softmax(base / np.float32(0.8))
    cdf = np.cumsum(probs.astype(np.float32), dtype=np.float32)
    for u in (0.0, 0.25, 0.5, 0.9, 0.999):
        got = model.sample(5, 0.8, int(p[4]), u)
        want = int(np.searchsorted(cdf, np.float32(u) * cdf[-1], side="right"))
        assert abs(got - want) <= 1  # sequential fp32 running sum vs numpy's pairwise cumsum


def test_kv_cache_layout_is_time_major(model, gpt_golden):
    """main.zig:93-94,127-128 / ops.zig:152: row t of a block's k_cache is the key of token t, heads interleaved."""
    k, v = model.kv(0, 4)
    assert k.shape == (4, 768) and np.abs(k).sum() > 0 and np.abs(v).sum() > 0
