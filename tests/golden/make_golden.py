"""Regenerates tests/golden/*.npz.  Run in the build container (needs /root/reference + torch):

    python tests/golden/make_golden.py

What it does
------------
1. `ops_golden.npz`: follows the reference's fixture *procedure* (generate_test_data.py:1-135 --
   same ops, same shapes, PyTorch as the arithmetic) for the 8 unit tests of src/tests.zig:22-388.
   The reference calls torch.randn unseeded and ships no bytes (models/ is git-ignored), so inputs
   here come from numpy's frozen RandomState stream (see golden_inputs.py) and only the torch
   OUTPUTS are committed, with a sha256 of the inputs they belong to.
2. `gpt_golden.npz`: executes the reference's own PyTorch model (the class definitions of
   /root/reference/generate_nano_gpt.py:24-152, exec'd from where they lie -- nothing is copied)
   on the synthetic 124M weights and a 16-token prompt, and records last-position logits and a
   greedy continuation that follows generate()'s loop (main.zig:322-342), duplicate last prompt
   token included.

The GPU box has no /root/reference; tests read only the committed .npz files.
"""
from __future__ import annotations

import math
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)

from golden_inputs import ops_inputs, inputs_digest, gpt_prompt  # noqa: E402
from zig_gpt2_b200.config import SIZES  # noqa: E402
from zig_gpt2_b200.weights import synth_for_size, fingerprint  # noqa: E402

REF = "/root/reference"


def t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def make_ops():
    i = ops_inputs()
    o = {}
    with torch.no_grad():
        # Linear, generate_test_data.py:10-28
        o["linear_outputs"] = F.linear(t(i["linear_inputs"]), t(i["linear_weight"]), t(i["linear_bias"])).numpy()
        o["linear_outputs_no_bias"] = F.linear(t(i["linear_inputs"]), t(i["linear_weight"])).numpy()
        # GELU, :31-45 (tanh formula with x + 0.044715 x^3)
        x = t(i["gelu_inputs"])
        o["gelu_outputs"] = (0.5 * x * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * torch.pow(x, 3.0))))).numpy()
        # softmax, :48-51
        o["softmax_outputs"] = F.softmax(t(i["softmax_inputs"]), dim=-1).numpy()
        # Embedding, :54-64
        o["embedding_outputs"] = F.embedding(t(i["embedding_inputs"]), t(i["embedding_weight"])).numpy()
        # LayerNorm, :67-78 (default affine) + a non-trivial affine case
        o["layer_norm_outputs"] = F.layer_norm(t(i["layer_norm_inputs"]), (768,), t(i["layer_norm_weight"]), t(i["layer_norm_bias"])).numpy()
        o["layer_norm_affine_outputs"] = F.layer_norm(t(i["layer_norm_inputs"]), (768,), t(i["layer_norm_affine_weight"]), t(i["layer_norm_affine_bias"])).numpy()
        # transpose / split, :85-95; generated at batch 1 (as the reference does) and batch 3 (as tests.zig reads)
        for b in (1, 3):
            o[f"transpose_outputs_b{b}"] = t(i[f"transpose_inputs_b{b}"]).transpose(1, 2).contiguous().numpy()
            q, k, v = t(i[f"split_inputs_b{b}"]).split(768, dim=2)
            o[f"split_q_b{b}"], o[f"split_k_b{b}"], o[f"split_v_b{b}"] = (z.contiguous().numpy() for z in (q, k, v))
        # causal self attention, :98-130
        B, T, H, hd = 1, 5, 12, 64
        E = H * hd
        y = F.linear(t(i["attn_inputs"]), t(i["attn_c_attn_weight"]), t(i["attn_c_attn_bias"]))
        q, k, v = y.split(E, dim=2)
        q, k, v = (z.view(B, T, H, hd).transpose(1, 2) for z in (q, k, v))
        mask = torch.tril(torch.ones(T, T).view(1, 1, T, T))
        att = q @ k.transpose(-2, -1) / math.sqrt(k.size(-1))
        att = att.masked_fill(mask == 0, float("-inf"))
        att = F.softmax(att, dim=-1)
        out = att @ v
        o["sdpa_q"], o["sdpa_k"], o["sdpa_v"], o["sdpa_outputs"] = (z.contiguous().numpy() for z in (q, k, v, out))
        z = out.transpose(1, 2).contiguous().view(B, T, E)
        o["attn_outputs"] = F.linear(z, t(i["attn_c_proj_weight"]), t(i["attn_c_proj_bias"])).numpy()
    o["inputs_sha256"] = np.array(inputs_digest(i))
    np.savez_compressed(os.path.join(HERE, "ops_golden.npz"), **o)
    print("ops_golden.npz:", {k: v.shape for k, v in o.items()})


def reference_torch_classes():
    """exec the model classes of the reference's generate_nano_gpt.py (lines up to `def load_linear`,
    i.e. :1-152) in a fresh namespace.  The rest of that file needs a checkpoint and network."""
    src = open(os.path.join(REF, "generate_nano_gpt.py")).read()
    head = src[: src.index("def load_linear")]
    import types

    mod = types.ModuleType("reference_generate_nano_gpt")
    sys.modules[mod.__name__] = mod  # dataclasses looks the module up by name
    exec(compile(head, os.path.join(REF, "generate_nano_gpt.py"), "exec"), mod.__dict__)
    return mod.__dict__


def make_gpt(n_new: int = 12):
    ns = reference_torch_classes()
    cfg = SIZES["124M"]
    w = synth_for_size("124M")
    model = ns["GPT"](ns["GPTConfig"]())
    tr = model.transformer
    with torch.no_grad():
        model.lm_head.weight.data = t(w["wte"])  # tied (generate_nano_gpt.py:124-125, :213)
        tr.wte.weight = model.lm_head.weight
        tr.wpe.weight.data = t(w["wpe"])
        for l, blk in enumerate(tr.h):
            blk.ln_1.weight.data, blk.ln_1.bias.data = t(w[f"h{l}-ln_1-g"]), t(w[f"h{l}-ln_1-b"])
            blk.attn.c_attn.weight.data, blk.attn.c_attn.bias.data = t(w[f"h{l}-attn-c_attn-w"]), t(w[f"h{l}-attn-c_attn-b"])
            blk.attn.c_proj.weight.data, blk.attn.c_proj.bias.data = t(w[f"h{l}-attn-c_proj-w"]), t(w[f"h{l}-attn-c_proj-b"])
            blk.ln_2.weight.data, blk.ln_2.bias.data = t(w[f"h{l}-ln_2-g"]), t(w[f"h{l}-ln_2-b"])
            blk.mlp.c_fc.weight.data, blk.mlp.c_fc.bias.data = t(w[f"h{l}-mlp-c_fc-w"]), t(w[f"h{l}-mlp-c_fc-b"])
            blk.mlp.c_proj.weight.data, blk.mlp.c_proj.bias.data = t(w[f"h{l}-mlp-c_proj-w"]), t(w[f"h{l}-mlp-c_proj-b"])
        tr.ln_f.weight.data, tr.ln_f.bias.data = t(w["ln_f-g"]), t(w["ln_f-b"])
    model.eval()
    prompt = gpt_prompt(cfg.vocab_size)
    o = {"weights_fingerprint": np.array(fingerprint(w)), "prompt": prompt}
    with torch.no_grad():
        # last-position logits of the prompt itself == GPT.forward(seq_len=16, prompt[15], true)
        o["prompt_logits"] = model(t(prompt)[None, :])[0, -1].numpy()
        # generate() semantics (main.zig:329-338): the first sampled step forwards prompt[-1] AGAIN
        # at the next position, so the effective sequence is prompt + [prompt[-1]] + generated...
        seq = list(prompt) + [int(prompt[-1])]
        toks, top = [], []
        full = []
        for _ in range(n_new):
            logits = model(torch.tensor(seq, dtype=torch.long)[None, :])[0, -1]
            nxt = int(torch.argmax(logits))
            srt = torch.sort(logits, descending=True)
            toks.append(nxt)
            top.append(np.stack([srt.values[:4].numpy(), srt.indices[:4].numpy().astype(np.float32)]))
            full.append(logits.numpy())
            seq.append(nxt)
        o["greedy_tokens"] = np.array(toks, np.int64)
        o["greedy_top4"] = np.stack(top)  # [n_new, 2, 4]: values, indices
        o["greedy_logits_first"] = full[0]
        o["greedy_logits_last"] = full[-1]
    np.savez_compressed(os.path.join(HERE, "gpt_golden.npz"), **o)
    print("gpt_golden.npz: greedy tokens", toks, "margins", [float(x[0, 0] - x[0, 1]) for x in top])


if __name__ == "__main__":
    torch.manual_seed(0)
    torch.set_num_threads(8)
    make_ops()
    make_gpt()
