"""-m gpu: the `zig_gpt2 "<prompt>"` program (main.zig:344-371) end to end -- C++ host (csrc/host/main.cpp: load_encoder,
load_gpt, State.init, encode, generate, decode + print to stderr) over the CUDA shim, against the oracle's restatement
of the same loop on the same synthetic model files."""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from zig_gpt2_b200.config import GPTConfig  # noqa: E402

CFG = GPTConfig(vocab_size=4099, context_size=128, n_layer=2, n_heads=4, n_embed=256)
PROMPT = b"Marcus Aurelius said thus: it's 42"


@pytest.fixture(scope="module")
def model_dir(tmp_path_factory):
    from zig_gpt2_b200 import build
    from zig_gpt2_b200.vocab import write_vocab
    from zig_gpt2_b200.weights import save_raw, synth_weights

    d = tmp_path_factory.mktemp("model")
    w = synth_weights(CFG, seed=3)
    save_raw(w, str(d / "raw"))
    write_vocab(str(d), vocab_size=CFG.vocab_size)
    exe = build.build_cli()
    return str(d), w, exe


def run_cli(exe, d, *args):
    cfg = f"{CFG.vocab_size},{CFG.context_size},{CFG.n_layer},{CFG.n_heads},{CFG.n_embed}"
    r = subprocess.run([exe, "--config", cfg, "--model-dir", d, *args, PROMPT.decode()], capture_output=True, timeout=300)
    assert r.returncode == 0, r.stderr[-400:]
    return r.stderr  # the reference prints to stderr (std.debug.print, main.zig:340)


def test_cli_greedy_output_is_the_oracles_generate(model_dir):
    import zg_oracle as zo
    from zig_gpt2_b200.vocab import synth_encoder, unicode_to_bytes

    d, w, exe = model_dir
    enc = zo.Encoder(synth_encoder(CFG.vocab_size), unicode_to_bytes())
    ids = enc.encode(PROMPT)
    assert 0 < len(ids) < 40
    zo.use_scalar_blas()
    orc = zo.Model(CFG, w)
    toks = orc.generate_greedy(np.array(ids), len(ids) + 16)
    orc.close()
    want = b"".join(enc.decode([int(t)]) for t in toks) + b"\n"  # every token, prompt included (main.zig:339-340)
    got = run_cli(exe, d, "--greedy", "--max-tokens", "16")
    assert got == want
    assert got.startswith(PROMPT)


def test_cli_sampling_is_reproducible_with_a_seed(model_dir):
    d, _, exe = model_dir
    a = run_cli(exe, d, "--temp", "0.8", "--seed", "7", "--max-tokens", "12")
    b = run_cli(exe, d, "--temp", "0.8", "--seed", "7", "--max-tokens", "12")
    c = run_cli(exe, d, "--temp", "0.8", "--seed", "8", "--max-tokens", "12")
    assert a == b and a.startswith(PROMPT)
    assert a != c


def test_cli_reports_missing_model_files(model_dir, tmp_path):
    _, _, exe = model_dir
    r = subprocess.run([exe, "--config", "4099,128,2,4,256", "--model-dir", str(tmp_path), "hi"], capture_output=True, timeout=60)
    assert r.returncode == 1 and b"cannot load" in r.stderr


def test_cli_fidelity_tokenizer_mode(model_dir, tmp_path):
    """--tokenizer gpt2: the opt-in byte-pair tokenizer drives the same generate loop.  With an empty merge list every
    byte is its own token (ids 0..255 of the synthetic vocabulary), which the oracle loop can follow."""
    import zg_oracle as zo

    d, w, exe = model_dir
    merges = tmp_path / "vocab.bpe"
    merges.write_text("#version: 0.2\n")
    from zig_gpt2_b200.vocab import synth_encoder, unicode_to_bytes

    enc, u2b = synth_encoder(CFG.vocab_size), unicode_to_bytes()
    b2u = {b: u for u, b in u2b.items()}
    ids = [enc[b2u[c]] for c in PROMPT]
    zo.use_scalar_blas()
    orc = zo.Model(CFG, w)
    toks = orc.generate_greedy(np.array(ids), len(ids) + 8)
    orc.close()
    dec = zo.Encoder(enc, u2b)
    want = b"".join(dec.decode([int(t)]) for t in toks) + b"\n"
    got = run_cli(exe, d, "--greedy", "--max-tokens", "8", "--tokenizer", "gpt2", "--merges", str(merges))
    assert got == want
