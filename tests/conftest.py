import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests", "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")
    config.addinivalue_line("markers", "slow: takes more than ~30 s on CPU")


def _gpu_box_ready() -> str:
    """'' when the CUDA extension can run here, else why not (a CPU box: skip instead of erroring in every fixture)."""
    try:
        from zig_gpt2_b200 import lib

        if lib.load().zg_device_count() == 0:
            return "no CUDA device"
    except Exception as e:  # missing .so
        return f"libzg_b200.so unavailable: {e}"
    return ""


def pytest_collection_modifyitems(config, items):
    why = None
    for item in items:
        if "gpu" in item.keywords:
            if why is None:
                why = _gpu_box_ready()
            if why:
                item.add_marker(pytest.mark.skip(reason=f"gpu test: {why}"))


def assert_tensors_approx_equal(expected, actual, abs_tol=5e-7, rel_tol=6e-4, what=""):
    """The reference's comparator, src/tests.zig:4-20: per element, if |expected| < 1e-3 the
    absolute tolerance is 5e-7, otherwise the relative tolerance is 6e-4."""
    e = np.asarray(expected, np.float64).reshape(-1)
    a = np.asarray(actual, np.float64).reshape(-1)
    assert e.shape == a.shape, f"{what}: shape {a.shape} != {e.shape}"
    small = np.abs(e) < 1e-3
    err = np.abs(e - a)
    bad = np.where(small, err > abs_tol, err > rel_tol * np.abs(e))
    if bad.any():
        i = int(np.argmax(bad))
        raise AssertionError(f"{what}: {int(bad.sum())}/{e.size} elements out of tolerance; first at {i}: expected {e[i]!r} got {a[i]!r}")


@pytest.fixture(scope="session")
def ops_golden():
    from golden_inputs import inputs_digest, ops_inputs

    g = np.load(os.path.join(ROOT, "tests", "golden", "ops_golden.npz"))
    i = ops_inputs()
    assert str(g["inputs_sha256"]) == inputs_digest(i), "golden inputs drifted: rerun tests/golden/make_golden.py"
    return i, g


@pytest.fixture(scope="session")
def gpt_golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "gpt_golden.npz"))


@pytest.fixture(scope="session")
def weights_124m(gpt_golden):
    from zig_gpt2_b200.weights import fingerprint, synth_for_size

    w = synth_for_size("124M")
    assert fingerprint(w) == str(gpt_golden["weights_fingerprint"]), "synthetic weights drifted from the golden fixture"
    return w
