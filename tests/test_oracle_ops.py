"""Pins the CPU oracle (oracle/, a C restatement of src/ops.zig) against the reference's own
unit tests: the 8 cases of src/tests.zig:22-388 with its comparator (src/tests.zig:4-20), on
fixtures produced by the generate_test_data.py procedure (tests/golden/make_golden.py)."""
import numpy as np
import pytest

import zg_oracle as zo
from conftest import assert_tensors_approx_equal as approx


@pytest.fixture(params=["scalar", "openblas"], autouse=True)
def blas(request):
    if request.param == "openblas":
        if not zo.use_openblas():
            pytest.skip("no scipy-bundled OpenBLAS")
    else:
        zo.use_scalar_blas()
    yield request.param
    zo.use_scalar_blas()


def test_linear(ops_golden):  # tests.zig:22-78
    i, g = ops_golden
    approx(g["linear_outputs"], zo.linear(i["linear_inputs"], i["linear_weight"], i["linear_bias"]), what="linear")
    approx(g["linear_outputs_no_bias"], zo.linear(i["linear_inputs"], i["linear_weight"], None), what="linear no bias")


def test_embedding(ops_golden):  # tests.zig:80-114; indices are 64-bit
    i, g = ops_golden
    out = zo.embedding(i["embedding_weight"], i["embedding_inputs"])
    assert np.array_equal(out, g["embedding_outputs"])


def test_layer_norm(ops_golden):  # tests.zig:116-155
    i, g = ops_golden
    approx(g["layer_norm_outputs"], zo.layer_norm(i["layer_norm_inputs"], i["layer_norm_weight"], i["layer_norm_bias"]), what="layer_norm")
    approx(g["layer_norm_affine_outputs"],
           zo.layer_norm(i["layer_norm_inputs"], i["layer_norm_affine_weight"], i["layer_norm_affine_bias"]), what="layer_norm affine")


@pytest.mark.parametrize("b", [1, 3])
def test_split_qkv(ops_golden, b):  # tests.zig:157-209
    i, g = ops_golden
    for idx, name in enumerate(("q", "k", "v")):
        out = zo.split_qkv(i[f"split_inputs_b{b}"], 5, 12, 768, idx)
        assert np.array_equal(out, g[f"split_{name}_b{b}"].reshape(-1))


@pytest.mark.parametrize("b", [1, 3])
def test_transpose(ops_golden, b):  # tests.zig:211-243
    i, g = ops_golden
    out = zo.transpose(i[f"transpose_inputs_b{b}"], (5, 12, 64))
    assert np.array_equal(out, g[f"transpose_outputs_b{b}"].reshape(-1))


def test_attention_forward_incremental(ops_golden):  # tests.zig:245-334: KV-cache steps == causal full sequence
    i, g = ops_golden
    run = zo.AttentionRunner(12, 768, i["attn_c_attn_weight"], i["attn_c_attn_bias"], i["attn_c_proj_weight"], i["attn_c_proj_bias"])
    for s in range(5):
        out = run.step(s + 1, i["attn_inputs"][0, s])
        approx(g["attn_outputs"][0, s], out, what=f"attention step {s}")


def test_sdpa_last_row(ops_golden):  # generate_test_data.py:109-119 (fixture the reference never consumes)
    _, g = ops_golden
    T = 5
    out = zo.sdpa(g["sdpa_q"][:, :, T - 1 : T, :], g["sdpa_k"], g["sdpa_v"], 12, T, 64)
    approx(g["sdpa_outputs"][0, :, T - 1, :], out, what="sdpa")


def test_gelu(ops_golden):  # tests.zig:336-360
    i, g = ops_golden
    approx(g["gelu_outputs"], zo.gelu(i["gelu_inputs"]), what="gelu")


def test_softmax(ops_golden):  # tests.zig:362-388: called per row
    i, g = ops_golden
    for r in range(3):
        approx(g["softmax_outputs"][r], zo.softmax(i["softmax_inputs"][r]), what=f"softmax row {r}")
