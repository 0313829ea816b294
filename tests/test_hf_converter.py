"""Converter from a Hugging Face GPT-2 state dict (zig_gpt2_b200/convert.py; reference format: download_weights.py:57-65),
with `transformers`' own GPT2LMHeadModel as a THIRD independent oracle: a random-init HF model is converted, and

  * CPU: the C oracle (restatement of ops.zig / main.zig) on the converted tensors reproduces HF's logits -- full-sequence
    masked attention in PyTorch vs the reference's incremental KV-cache loop;
  * GPU (-m gpu): so do the fused engine and the op-by-op path.

HF's GPT-2 uses the same tanh GELU ("gelu_new"), LayerNorm eps 1e-5 and 1/sqrt(head_dim) scaling as ops.zig:221-228,
:76, :268, so agreement is to fp32 rounding (1e-4 of the logit scale, the north_star tolerance)."""
import numpy as np
import pytest

transformers = pytest.importorskip("transformers")
torch = pytest.importorskip("torch")

from zig_gpt2_b200.config import GPTConfig  # noqa: E402
from zig_gpt2_b200.convert import config_from_hf, from_hf_state_dict  # noqa: E402
from zig_gpt2_b200.weights import load_raw, save_raw, tensor_shapes  # noqa: E402


def hf_model(n_layer=2, n_head=4, n_embd=256, vocab=1031, ctx=64, seed=0):
    torch.manual_seed(seed)
    conf = transformers.GPT2Config(vocab_size=vocab, n_positions=ctx, n_embd=n_embd, n_layer=n_layer, n_head=n_head,
                                   resid_pdrop=0.0, embd_pdrop=0.0, attn_pdrop=0.0)
    m = transformers.GPT2LMHeadModel(conf).eval()
    with torch.no_grad():  # HF initialises biases to 0 and LayerNorm to (1, 0): perturb them so every path is exercised
        for name, p in m.named_parameters():
            if name.endswith(".bias"):
                p.add_(0.02 * torch.randn_like(p))
            elif "ln_" in name and name.endswith(".weight"):
                p.add_(0.02 * torch.randn_like(p))
            elif name.endswith(".weight"):
                p.mul_(4.0)  # std 0.02 -> 0.08: usable logit scale for a 2-layer model
    return m


@pytest.fixture(scope="module")
def converted():
    m = hf_model()
    cfg = config_from_hf(m.config)
    w = from_hf_state_dict(m.state_dict(), cfg)
    ids = np.random.RandomState(1).randint(0, cfg.vocab_size, 12)
    with torch.no_grad():
        logits = m(torch.tensor(ids[None].astype(np.int64))).logits[0].numpy()
    return m, cfg, w, ids, logits


def test_layout_names_shapes_and_transposes(converted):
    m, cfg, w, _, _ = converted
    assert cfg == GPTConfig(1031, 64, 2, 4, 256)
    assert list(w) == list(tensor_shapes(cfg)) and all(w[n].shape == s for n, s in tensor_shapes(cfg).items())
    sd = m.state_dict()
    assert np.array_equal(w["h1-attn-c_attn-w"], sd["transformer.h.1.attn.c_attn.weight"].numpy().T)  # Conv1D is [in, out]
    assert np.array_equal(w["h0-mlp-c_proj-w"], sd["transformer.h.0.mlp.c_proj.weight"].numpy().T)
    assert np.array_equal(w["wte"], sd["lm_head.weight"].numpy())  # tied (main.zig:312)
    assert all(a.dtype == np.float32 and a.flags["C_CONTIGUOUS"] for a in w.values())


def test_round_trip_through_the_raw_files(converted, tmp_path):
    _, cfg, w, _, _ = converted
    save_raw(w, str(tmp_path))
    back = load_raw(cfg, str(tmp_path))
    assert all(np.array_equal(back[n], w[n]) for n in w)


def test_bad_state_dicts_are_refused(converted):
    m, cfg, _, _, _ = converted
    sd = dict(m.state_dict())
    del sd["transformer.h.1.ln_2.bias"]
    with pytest.raises(KeyError):
        from_hf_state_dict(sd, cfg)
    with pytest.raises(ValueError):
        from_hf_state_dict(m.state_dict(), GPTConfig(1031, 64, 2, 4, 128))


def test_oracle_reproduces_hf_logits_at_every_position(converted):
    import zg_oracle as zo

    _, cfg, w, ids, logits = converted
    zo.use_scalar_blas()
    orc = zo.Model(cfg, w)
    scale = float(np.abs(logits).max())
    for s, t in enumerate(ids):
        got = orc.forward(s + 1, int(t), True)
        assert float(np.abs(got - logits[s]).max()) <= 1e-4 * scale, s
    orc.close()


@pytest.mark.gpu
@pytest.mark.parametrize("fused", [True, False])
def test_gpu_paths_reproduce_hf_logits(converted, fused):
    from zig_gpt2_b200 import gpt, lib

    lib.init(0)
    _, cfg, w, ids, logits = converted
    model, state = gpt.gpt_from_numpy(cfg, w), gpt.State(cfg)
    fwd = model.forward if fused else model.forward_unfused
    scale = float(np.abs(logits).max())
    for s, t in enumerate(ids):
        fwd(s + 1, int(t), True, state)
        assert float(np.abs(state.logits.download() - logits[s]).max()) <= 1e-4 * scale, s
    model.close()


def test_convert_a_local_hf_checkpoint_directory(tmp_path):
    """`python -m zig_gpt2_b200.convert <hf_dir> <out_dir>`: save_pretrained() directory -> raw/model-* files + vocab JSONs."""
    import json

    from zig_gpt2_b200.convert import convert_hf_dir

    m = hf_model(n_layer=1, n_head=2, n_embd=128, vocab=300, ctx=32, seed=3)
    hf_dir, out_dir = tmp_path / "hf", tmp_path / "model"
    m.save_pretrained(str(hf_dir))
    json.dump({"a": 0, "b": 1}, open(hf_dir / "vocab.json", "w"))
    cfg = convert_hf_dir(str(hf_dir), str(out_dir))
    assert cfg == GPTConfig(300, 32, 1, 2, 128)
    back = load_raw(cfg, str(out_dir / "raw"))
    want = from_hf_state_dict(m.state_dict(), cfg)
    assert all(np.array_equal(back[n], want[n]) for n in want)
    assert json.load(open(out_dir / "encoder.json")) == {"a": 0, "b": 1}
    assert len(json.load(open(out_dir / "byte_encoder.json"))) == 256
