"""Loaders (reference: main.zig:210-320, ops.zig:309-326, on-disk format download_weights.py:57-65).

CPU part: the raw-file format round trip and the JSON vocabulary loading of the C++ host tokenizer.
GPU part (-m gpu): save_raw -> zg_load_gpt (C) and gpt.load_gpt (Python mirror) give the same logits as the
in-memory assembly, and a truncated tensor file is refused.  The reference reads with `fd.readAll` and ignores the
byte count (ops.zig:318), i.e. a short file silently leaves the tail of the slice uninitialised; both loaders here
refuse it instead (stated divergence)."""
import ctypes as C
import os

import numpy as np
import pytest

from zig_gpt2_b200.config import GPTConfig
from zig_gpt2_b200.weights import load_raw, save_raw, synth_weights, tensor_shapes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CFG = GPTConfig(vocab_size=1031, context_size=64, n_layer=2, n_heads=4, n_embed=256)  # the fused engine needs n_embed >= SMs


def test_raw_format_is_headerless_little_endian_f32(tmp_path):
    w = synth_weights(CFG, seed=11)
    save_raw(w, str(tmp_path))
    names = sorted(os.listdir(tmp_path))
    assert names == sorted(f"model-{n}" for n in tensor_shapes(CFG))  # main.zig:210-314 file names
    for n, shape in tensor_shapes(CFG).items():
        raw = open(tmp_path / f"model-{n}", "rb").read()
        assert len(raw) == 4 * int(np.prod(shape))  # no header
        assert np.array_equal(np.frombuffer(raw, "<f4").reshape(shape), w[n])
    back = load_raw(CFG, str(tmp_path))
    assert all(np.array_equal(back[n], w[n]) for n in w)


def test_raw_short_file_is_refused(tmp_path):
    w = synth_weights(CFG, seed=11)
    save_raw(w, str(tmp_path))
    path = tmp_path / "model-h1-mlp-c_fc-w"
    data = open(path, "rb").read()
    open(path, "wb").write(data[:-64])
    with pytest.raises(ValueError):
        load_raw(CFG, str(tmp_path))


def test_host_tokenizer_loads_json_vocab_files(tmp_path):
    """load_encoder (main.zig:316-320): encoder.json + byte_encoder.json -> the same tokenizer as the in-memory maps."""
    import zg_oracle as zo
    from zig_gpt2_b200 import build
    from zig_gpt2_b200.vocab import synth_encoder, unicode_to_bytes, write_vocab

    build.build_host()
    write_vocab(str(tmp_path), vocab_size=5000)
    H = C.CDLL(os.path.join(ROOT, "zig_gpt2_b200", "libzg_host.so"))
    H.zgh_encoder_create_from_files.restype = C.c_void_p
    H.zgh_encoder_create_from_files.argtypes = [C.c_char_p, C.c_char_p]
    H.zgh_encoder_encode.restype = C.c_size_t
    H.zgh_encoder_encode.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, C.POINTER(C.c_size_t), C.c_size_t]
    H.zgh_encoder_decode.restype = C.c_size_t
    H.zgh_encoder_decode.argtypes = [C.c_void_p, C.POINTER(C.c_size_t), C.c_size_t, C.POINTER(C.c_ubyte), C.c_size_t]
    H.zgh_encoder_destroy.argtypes = [C.c_void_p]
    h = H.zgh_encoder_create_from_files(str(tmp_path / "encoder.json").encode(), str(tmp_path / "byte_encoder.json").encode())
    assert h
    oracle = zo.Encoder(synth_encoder(5000), unicode_to_bytes())
    for text in [b"Marcus Aurelius said thus: ", b"it's 42!!  two  spaces\nnew", "café naïve".encode(), b" ", b"a\tb"]:
        out = (C.c_size_t * 1024)()
        n = H.zgh_encoder_encode(h, text, len(text), out, 1024)
        ids = list(out[:n])
        assert ids == oracle.encode(text)
        arr = (C.c_size_t * len(ids))(*ids)
        buf = (C.c_ubyte * 4096)()
        m = H.zgh_encoder_decode(h, arr, len(ids), buf, 4096)
        assert bytes(buf[:m]) == text
    H.zgh_encoder_destroy(h)
    assert not H.zgh_encoder_create_from_files(b"/nonexistent/encoder.json", b"/nonexistent/byte_encoder.json")


# ---- GPU -----------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_load_gpt_from_raw_files_matches_in_memory_model(tmp_path, monkeypatch):
    from zig_gpt2_b200 import gpt, lib

    L = lib.init(0)
    w = synth_weights(CFG, seed=11)
    save_raw(w, str(tmp_path / "raw"))
    prompt = [5, 900, 17, 3]
    ref_model, ref_state = gpt.gpt_from_numpy(CFG, w), gpt.State(CFG)
    for s, t in enumerate(prompt):
        ref_model.forward(s + 1, t, s == len(prompt) - 1, ref_state)
    want = ref_state.logits.download()

    # (1) the Python mirror of load_gpt / load_tensor
    m2, s2 = gpt.load_gpt(CFG, str(tmp_path)), gpt.State(CFG)
    for s, t in enumerate(prompt):
        m2.forward(s + 1, t, s == len(prompt) - 1, s2)
    assert np.array_equal(s2.logits.download(), want)

    # (2) zg_load_gpt: mmap -> two pinned staging buffers -> async H2D, then the same engine.  A 4 KB chunk forces
    # hundreds of chunks per tensor through the double-buffered path.
    monkeypatch.setenv("ZG_LOAD_CHUNK_KB", "4")
    g = lib.ZgGPT()
    c = lib.ZgConfig(CFG.vocab_size, CFG.context_size, CFG.n_layer, CFG.n_heads, CFG.n_embed)
    assert L.zg_load_gpt(C.byref(g), C.byref(c), str(tmp_path / "raw").encode()) == 0
    lib.check()
    st = lib.ZgState()
    assert L.zg_state_init(C.byref(st), C.byref(c), 0) == 0
    eng = L.zg_engine_create(C.byref(g), C.byref(st))
    assert eng
    for s, t in enumerate(prompt):
        L.zg_engine_forward(eng, s + 1, t, int(s == len(prompt) - 1))
    got = np.empty(CFG.vocab_size, np.float32)
    L.zg_download(got.ctypes.data, st.logits, got.nbytes)
    lib.check()
    assert np.array_equal(got, want)
    L.zg_engine_destroy(eng)
    L.zg_state_free(C.byref(st))
    L.zg_gpt_free(C.byref(g))
    ref_model.close()
    m2.close()


@pytest.mark.gpu
def test_load_gpt_refuses_missing_and_truncated_files(tmp_path):
    from zig_gpt2_b200 import gpt, lib

    L = lib.init(0)
    w = synth_weights(CFG, seed=11)
    save_raw(w, str(tmp_path / "raw"))
    path = tmp_path / "raw" / "model-h0-attn-c_attn-w"
    data = open(path, "rb").read()
    open(path, "wb").write(data[: len(data) // 2])
    g = lib.ZgGPT()
    c = lib.ZgConfig(CFG.vocab_size, CFG.context_size, CFG.n_layer, CFG.n_heads, CFG.n_embed)
    assert L.zg_load_gpt(C.byref(g), C.byref(c), str(tmp_path / "raw").encode()) != 0
    assert b"short read" in L.zg_last_error_string()
    L.zg_clear_error()
    with pytest.raises(ValueError):
        gpt.load_gpt(CFG, str(tmp_path))
    os.remove(path)
    assert L.zg_load_gpt(C.byref(g), C.byref(c), str(tmp_path / "raw").encode()) != 0
    assert b"cannot open" in L.zg_last_error_string()
    L.zg_clear_error()
