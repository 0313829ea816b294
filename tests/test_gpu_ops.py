"""-m gpu: the CUDA operators (through the C-ABI) against the reference's own unit-test cases
(src/tests.zig:22-388; fixtures by tests/golden/make_golden.py) with its comparator
(src/tests.zig:4-20), and against the CPU oracle on the same inputs."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from conftest import assert_tensors_approx_equal as approx  # noqa: E402


@pytest.fixture(scope="module")
def zg():
    from zig_gpt2_b200 import lib, ops

    lib.init(0)
    return ops


def dev(a):
    from zig_gpt2_b200.lib import DeviceBuffer

    return DeviceBuffer.from_numpy(np.ascontiguousarray(a, np.float32))


def out(n):
    from zig_gpt2_b200.lib import DeviceBuffer

    return DeviceBuffer(n)


def test_linear(zg, ops_golden):  # tests.zig:22-78
    i, g = ops_golden
    x, w, b = dev(i["linear_inputs"]), dev(i["linear_weight"]), dev(i["linear_bias"])
    o = out(3 * 3072)
    zg.Linear(768, 3072, w, b).forward(x, o)
    approx(g["linear_outputs"], o.download(), what="linear")
    zg.Linear(768, 3072, w, None).forward(x, o)
    approx(g["linear_outputs_no_bias"], o.download(), what="linear no bias")


@pytest.mark.parametrize("M,K,N", [(1, 768, 2304), (1, 3072, 768), (5, 1600, 4800), (8, 1024, 1024), (11, 768, 50257), (1, 768, 1)])
def test_linear_shapes_vs_oracle(zg, M, K, N):
    import zg_oracle as zo

    rs = np.random.RandomState(M * 1000 + N)
    x = rs.standard_normal((M, K)).astype(np.float32)
    w = (rs.standard_normal((N, K)) * 0.05).astype(np.float32)
    b = rs.standard_normal(N).astype(np.float32)
    o = out(M * N)
    zg.Linear(K, N, dev(w), dev(b)).forward(dev(x), o)
    zo.use_scalar_blas()
    ref = zo.linear(x, w, b)
    np.testing.assert_allclose(o.download().reshape(M, N), ref, rtol=1e-4, atol=1e-4 * np.abs(ref).max())


def test_embedding(zg, ops_golden):  # tests.zig:80-114, 64-bit indices, exact copy
    i, g = ops_golden
    o = out(3 * 768)
    zg.Embedding(768, dev(i["embedding_weight"])).forward(i["embedding_inputs"], o)
    assert np.array_equal(o.download().reshape(3, 768), g["embedding_outputs"])


def test_embedding_many_indices(zg):
    rs = np.random.RandomState(3)
    w = rs.standard_normal((100, 64)).astype(np.float32)
    idx = rs.randint(0, 100, 1000)
    o = out(1000 * 64)
    zg.Embedding(64, dev(w)).forward(idx, o)
    assert np.array_equal(o.download().reshape(1000, 64), w[idx])


def test_layer_norm(zg, ops_golden):  # tests.zig:116-155
    i, g = ops_golden
    x = dev(i["layer_norm_inputs"])
    zg.LayerNorm(768, dev(i["layer_norm_weight"]), dev(i["layer_norm_bias"])).forward(x)
    approx(g["layer_norm_outputs"], x.download(), what="layer_norm")
    x = dev(i["layer_norm_inputs"])
    zg.LayerNorm(768, dev(i["layer_norm_affine_weight"]), dev(i["layer_norm_affine_bias"])).forward(x)
    approx(g["layer_norm_affine_outputs"], x.download(), what="layer_norm affine")


@pytest.mark.parametrize("b", [1, 3])
def test_split_qkv(zg, ops_golden, b):  # tests.zig:157-209
    i, g = ops_golden
    attn = zg.CausalSelfAttention(12, 768, None, None)
    x = dev(i[f"split_inputs_b{b}"])
    for idx, name in enumerate("qkv"):
        o = out(b * 5 * 768)
        attn.split_qkv(5, x, idx, o)
        assert np.array_equal(o.download(), g[f"split_{name}_b{b}"].reshape(-1))


@pytest.mark.parametrize("b", [1, 3])
def test_transpose(zg, ops_golden, b):  # tests.zig:211-243
    i, g = ops_golden
    o = out(b * 5 * 768)
    zg.CausalSelfAttention.transpose((5, 12, 64), dev(i[f"transpose_inputs_b{b}"]), o)
    assert np.array_equal(o.download(), g[f"transpose_outputs_b{b}"].reshape(-1))


def test_attention_forward_incremental(zg, ops_golden):  # tests.zig:245-334
    i, g = ops_golden
    E, C = 768, 1024
    attn = zg.CausalSelfAttention(12, E, zg.Linear(E, 3 * E, dev(i["attn_c_attn_weight"]), dev(i["attn_c_attn_bias"])),
                                  zg.Linear(E, E, dev(i["attn_c_proj_weight"]), dev(i["attn_c_proj_bias"])))
    k_cache, v_cache = out(C * E), out(C * E)
    o, qkv, q, a = out(E), out(3 * E), out(E), out(C)
    for s in range(5):
        x = dev(i["attn_inputs"][0, s])
        attn.forward(s + 1, x, zg.view(k_cache, 0, (s + 1) * E), zg.view(v_cache, 0, (s + 1) * E), o, qkv, q, None, None,
                     zg.view(a, 0, s + 1))
        approx(g["attn_outputs"][0, s], o.download(), what=f"attention step {s}")


def test_sdpa(zg, ops_golden):  # generate_test_data.py:109-119
    _, g = ops_golden
    T = 5
    o = out(768)
    zg.scaled_dot_product_attention(dev(g["sdpa_q"][:, :, T - 1:T, :]), dev(g["sdpa_k"]), dev(g["sdpa_v"]), 12, T, 64, o)
    approx(g["sdpa_outputs"][0, :, T - 1, :], o.download(), what="sdpa")


def test_sdpa_batched_long_vs_oracle(zg):
    import zg_oracle as zo

    rs = np.random.RandomState(11)
    B, H, T, hd = 3, 25, 333, 64
    q = rs.standard_normal((B, H, 1, hd)).astype(np.float32)
    k = rs.standard_normal((B, H, T, hd)).astype(np.float32)
    v = rs.standard_normal((B, H, T, hd)).astype(np.float32)
    o = out(B * H * hd)
    zg.scaled_dot_product_attention(dev(q), dev(k), dev(v), H, T, hd, o)
    zo.use_scalar_blas()
    np.testing.assert_allclose(o.download(), zo.sdpa(q, k, v, H, T, hd), rtol=1e-4, atol=1e-5)


def test_gelu(zg, ops_golden):  # tests.zig:336-360
    i, g = ops_golden
    x = dev(i["gelu_inputs"])
    zg.gelu(x)
    approx(g["gelu_outputs"], x.download(), what="gelu")


def test_softmax(zg, ops_golden):  # tests.zig:362-388, per row
    i, g = ops_golden
    x = dev(i["softmax_inputs"])
    for r in range(3):
        zg.softmax(zg.view(x, r * 768, (r + 1) * 768))
    approx(g["softmax_outputs"], x.download(), what="softmax")


def test_softmax_vocab_wide(zg):
    import zg_oracle as zo

    x = (np.random.RandomState(5).standard_normal(50257) * 3).astype(np.float32)
    d = dev(x)
    zg.softmax(d)
    np.testing.assert_allclose(d.download(), zo.softmax(x), rtol=1e-4, atol=1e-9)


def test_empty_inputs_are_noops(zg):
    zg.gelu((0, 0))
    zg.softmax((0, 0))
    zg.Linear(768, 8, dev(np.zeros((8, 768))), None).forward((0, 0), (0, 0))
