"""GPU check of the attention kernels and the batch engine against the CPU oracle.  Run under gpurun."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import zg_oracle as zo  # noqa: E402
from zig_gpt2_b200 import gpt, lib  # noqa: E402
from zig_gpt2_b200.batch import BatchEngine  # noqa: E402
from zig_gpt2_b200.config import SIZES, GPTConfig  # noqa: E402
from zig_gpt2_b200.lib import DeviceBuffer  # noqa: E402
from zig_gpt2_b200.weights import synth_weights, synth_for_size  # noqa: E402

sys.path.insert(0, os.path.join(ROOT, "scripts"))
from gpu_check_tc import f16_to_f32, to_f16_bits  # noqa: E402

R = []


def rec(**kw):
    R.append(kw)
    print(kw, flush=True)


def attn_ref(qkv, B, T, H, E):
    """causal attention in float64 from a [B*T, 3E] matrix"""
    hd = E // H
    x = qkv.astype(np.float64).reshape(B, T, 3, H, hd)
    q, k, v = x[:, :, 0], x[:, :, 1], x[:, :, 2]
    out = np.zeros((B, T, H, hd))
    mask = np.tril(np.ones((T, T), bool))
    for b in range(B):
        for h in range(H):
            s = q[b, :, h] @ k[b, :, h].T / np.sqrt(hd)
            s = np.where(mask, s, -np.inf)
            p = np.exp(s - s.max(1, keepdims=True))
            p /= p.sum(1, keepdims=True)
            out[b, :, h] = p @ v[b, :, h]
    return out.reshape(B * T, E)


def check_attn_prefill(L, B, T, H, rng):
    E = H * 64
    qkv = rng.standard_normal((B * T, 3 * E)).astype(np.float32)
    bits = to_f16_bits(qkv)
    qkv = f16_to_f32(bits)
    d_in = DeviceBuffer.from_numpy(bits)
    d_out = DeviceBuffer(B * T * E, np.uint16)
    L.zg_attention_prefill(d_in.ptr, d_out.ptr, B, T, H, E)
    lib.check()
    err = L.zg_tc_error()
    got = f16_to_f32(d_out.download()).reshape(B * T, E)
    want = attn_ref(qkv, B, T, H, E)
    e = float(np.abs(got - want).max() / np.abs(want).max())
    rec(test="attn_prefill", B=B, T=T, H=H, rel_err=e, tc_err=err, ok=bool(err == 0 and e < 2e-2))
    return err == 0


def check_attn_decode(L, B, C, H, T, rng):
    E = H * 64
    q = rng.standard_normal((B, E)).astype(np.float32)
    k = rng.standard_normal((B, C, E)).astype(np.float32)
    v = rng.standard_normal((B, C, E)).astype(np.float32)
    dq, dk, dv = DeviceBuffer.from_numpy(q), DeviceBuffer.from_numpy(k), DeviceBuffer.from_numpy(v)
    out = DeviceBuffer(B * E)
    L.zg_attention_decode_batch(dq.ptr, dk.ptr, dv.ptr, B, C, H, E, T, out.ptr)
    lib.check()
    got = out.download().reshape(B, H, 64)
    qq = q.astype(np.float64).reshape(B, H, 64)
    kk = k[:, :T].astype(np.float64).reshape(B, T, H, 64)
    vv = v[:, :T].astype(np.float64).reshape(B, T, H, 64)
    s = np.einsum("bhd,bthd->bht", qq, kk) / 8.0
    p = np.exp(s - s.max(-1, keepdims=True))
    p /= p.sum(-1, keepdims=True)
    want = np.einsum("bht,bthd->bhd", p, vv)
    e = float(np.abs(got - want).max() / np.abs(want).max())
    rec(test="attn_decode_batch", B=B, C=C, H=H, T=T, rel_err=e, ok=bool(e < 1e-5))


def check_engine(cfg, w, B, n_inputs, n_total, rng, name):
    model = gpt.gpt_from_numpy(cfg, w)
    zo.use_openblas()
    prompts = rng.integers(0, cfg.vocab_size, (B, n_inputs))
    refs, ref_logits = [], []
    for b in range(B):
        orc = zo.Model(cfg, w)
        t, lg = orc.generate_greedy(prompts[b], n_total, want_logits=True)
        refs.append(np.asarray(t))
        ref_logits.append(lg)
        orc.close()
    refs = np.stack(refs)
    eng = BatchEngine(model, B, cache_rows=min(cfg.context_size, max(n_total, 128)), max_prompt=n_inputs)
    # (1) teacher-forced logits: feed the oracle's own tokens, compare the logits of every sampling step
    worst = 0.0
    for s in range(n_total):
        sampling = s >= n_inputs
        toks = refs[:, s] if not sampling else (refs[:, s - 1] if s > n_inputs else prompts[:, -1])
        eng.forward(s + 1, toks, sampling)
        if sampling:
            got = eng.logits()
            for b in range(B):
                want = ref_logits[b][s - n_inputs]
                worst = max(worst, float(np.abs(got[b] - want).max() / np.abs(want).max()))
    rec(test="batch_teacher_forced_logits", model=name, B=B, rel_err=worst, ok=bool(worst < 2e-2), tc_err=lib.load().zg_tc_error())
    # (2) free-running greedy, token at a time (the reference loop) and with the batched prefill
    for use_prefill in (False, True):
        got = eng.generate_greedy(prompts, n_total, use_prefill=use_prefill)
        match = (got == refs)
        first_bad = [int(np.argmin(m)) if not m.all() else n_total for m in match]
        rec(test="batch_generate_greedy", model=name, prefill=use_prefill, B=B, n_total=n_total,
            identical_sequences=int(match.all(1).sum()), first_mismatch_steps=first_bad[:16], tc_err=lib.load().zg_tc_error())
    # (3) prefill logits + caches vs token-at-a-time engine results
    eng.prefill(prompts, True)
    lg_p = eng.logits().copy()
    kp, vp = eng.kv(cfg.n_layer - 1, n_inputs)
    for s in range(n_inputs):
        eng.forward(s + 1, prompts[:, s], s == n_inputs - 1)
    lg_s = eng.logits()
    ks, vs = eng.kv(cfg.n_layer - 1, n_inputs)
    e1 = float(np.abs(lg_p - lg_s).max() / np.abs(lg_s).max())
    e2 = float(max(np.abs(kp - ks).max() / np.abs(ks).max(), np.abs(vp - vs).max() / np.abs(vs).max()))
    rec(test="prefill_vs_steps", model=name, logits_rel=e1, last_layer_kv_rel=e2, ok=bool(e1 < 2e-2 and e2 < 2e-2), tc_err=lib.load().zg_tc_error())
    eng.close()
    model.close()
    zo.use_scalar_blas()


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which == "all":  # every section in its own process with its own timeout: a hang costs one section, not the call
        import subprocess
        for sec, tmo in (("attn_prefill", 60), ("attn_decode", 60), ("small", 120), ("124M", 240)):
            try:
                r = subprocess.run([sys.executable, __file__, sec], timeout=tmo, capture_output=True, text=True)
                print(r.stdout[-6000:], r.stderr[-2000:], flush=True)
            except subprocess.TimeoutExpired as ex:
                print(f"SECTION {sec} TIMED OUT after {tmo}s", (ex.stdout or b"")[-3000:], flush=True)
        return
    L = lib.init(0)
    rng = np.random.default_rng(1)
    if which == "attn_prefill":
        ok = check_attn_prefill(L, 1, 128, 1, rng)
        if ok:
            check_attn_prefill(L, 2, 256, 2, rng)
            check_attn_prefill(L, 2, 200, 3, rng)
            check_attn_prefill(L, 1, 1024, 2, rng)
    if which == "attn_decode":
        check_attn_decode(L, 3, 64, 2, 1, rng)
        check_attn_decode(L, 3, 64, 2, 37, rng)
        check_attn_decode(L, 2, 1024, 4, 1024, rng)
    if which == "small":
        small = GPTConfig(vocab_size=4099, context_size=128, n_layer=2, n_heads=4, n_embed=256)
        check_engine(small, synth_weights(small, seed=3), 5, 8, 40, rng, "small")
    if which == "124M":
        t0 = time.time()
        check_engine(SIZES["124M"], synth_for_size("124M"), 4, 16, 80, rng, "124M")
        print("124M check took", time.time() - t0)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(R, open(f"gpurun_out/batch_check_{which}.json", "w"), indent=1)


if __name__ == "__main__":
    main()
