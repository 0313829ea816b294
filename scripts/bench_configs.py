#!/usr/bin/env python
"""BASELINE.json configs[2..4] (the batched, tensor-core paths) measured like bench.py measures configs[1].

bench.py imports the measure_* functions and attaches their records to its one JSON line (`configs`); run directly,
this file prints one JSON line per record (rank 0) and appends it to gpurun_out/bench_configs.jsonl:

    python scripts/bench_configs.py cfg3|cfg4|cfg5 [--small]
    torchrun --nproc-per-node N scripts/bench_configs.py cfg5      # sequences sharded over N GPUs, no collective

  cfg3  GPT-2 355M, prefill of 16 x 1024-token synthetic prompts, last-position logits (tensor roofline)
  cfg4  GPT-2 1.5B, batch 64, one decode step at context 1024 / 512 (HBM roofline: weights + KV caches)
  cfg5  GPT-2 124M, 1024 sequences (split over the ranks), 32-token prompts, 256 greedy tokens each
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from zig_gpt2_b200.config import SIZES, GPTConfig  # noqa: E402
from zig_gpt2_b200.weights import synth_for_size, synth_weights  # noqa: E402


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), float(p["bf16_tflops"]), float(p.get("bf16_tflops_sustained", p["bf16_tflops"])), "measured"
    except Exception:
        return 6650.0, 1590.0, 1400.0, "fallback"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz, self.ok = index, [], set(), None, False
        self._stop_evt = threading.Event()
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml_unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def measure_cfg3(L, lib, trials: int = 5, small: bool = False, device_index: int = 0) -> dict:
    """BASELINE configs[2]: one prefill pass over 16 x 1024 tokens, last-position logits (main.zig:192)."""
    from zig_gpt2_b200 import gpt as G
    from zig_gpt2_b200.batch import BatchEngine

    size = "355M"
    cfg = SIZES[size] if not small else GPTConfig(50257, 1024, 4, 16, 1024)
    B, T = 16, 1024
    w = synth_for_size(size) if not small else synth_weights(cfg, seed=5)
    model = G.gpt_from_numpy(cfg, w)
    del w
    eng = BatchEngine(model, B, cache_rows=T, max_prompt=T)
    toks = np.random.default_rng(1236).integers(0, cfg.vocab_size, (B, T))
    for _ in range(3):
        eng.prefill(toks, True)
    L.zg_sync()
    sampler = ClockSampler(device_index)
    sampler.start()
    ts, e2e = [], []
    n0 = L.zg_launch_count()
    for _ in range(trials):
        L.zg_timer_begin()
        eng.prefill_resident(T, True)
        ts.append(L.zg_timer_end_ms())
    launches = int(L.zg_launch_count() - n0) // trials
    for _ in range(trials):
        L.zg_sync()
        t0 = time.perf_counter()
        eng.prefill(toks, True)
        lg = eng.logits()
        e2e.append(time.perf_counter() - t0)
    clocks = sampler.stop()
    lib.check()
    ms = float(np.median(ts))
    flops = cfg.prefill_flops(B, T)
    _, tf_burst, tf_sus, kind = peaks()
    ach = flops / (ms * 1e-3) / 1e12
    rec = {"metric": "prefill_tokens_per_sec", "value": B * T / (ms * 1e-3), "unit": "tok/s", "n_gpus": 1, "steps": 1,
           "ms_per_step": ms, "higher_is_better": True, "dtype": "f16 operands, f32 accumulate", "data": "synthetic",
           "config": {"workload": f"GPT-2 {size}{' (4 layers)' if small else ''} prefill, {B} x {T}-token synthetic prompts, "
                                  "last-position logits (BASELINE configs[2]); tcgen05 GEMMs + causal flash attention",
                      "trials": trials, "l2": "activations + weights per pass (>1 GB) exceed L2"},
           "e2e": {"value": B * T / float(np.median(e2e)), "unit": "tok/s", "h2d_bytes_per_step": B * T * 8,
                   "d2h_bytes_per_step": int(lg.nbytes), "call": "zg_batch_prefill(host tokens) + logits download"},
           "gpu_launches": launches,
           "roofline": {"bound": "tensor", "achieved": ach, "peak": tf_sus, "unit": "TFLOP/s", "frac": ach / tf_sus,
                        "frac_of_burst": ach / tf_burst, "peak_kind": kind + " (sustained bf16 cuBLAS; burst beside it)",
                        "flops_per_launch": flops, "traffic": None},
           "clocks": clocks}
    eng.close()
    model.close()
    return rec


def measure_cfg4(L, lib, trials: int = 5, steps: int = 8, small: bool = False, device_index: int = 0) -> dict:
    """BASELINE configs[3]: GPT-2 1.5B, batch 64, single decode steps at context 1024 and 512, both GEMM modes."""
    from zig_gpt2_b200 import gpt as G
    from zig_gpt2_b200.batch import BatchEngine

    size = "1.5B"
    cfg = SIZES[size] if not small else GPTConfig(50257, 1024, 4, 25, 1600)
    B, T = 64, 1024
    w = synth_for_size(size) if not small else synth_weights(cfg, seed=5)
    model = G.gpt_from_numpy(cfg, w)
    del w
    out = {}
    tok_host = np.random.default_rng(1237).integers(0, cfg.vocab_size, B).astype(np.uint64)
    got = np.zeros(B, np.uint64)
    for mode_name, single in (("3xtf32", False), ("tf32", True), ("f16", None)):
        eng = (BatchEngine(model, B, cache_rows=T, tf32_single_pass=single) if single is not None
               else BatchEngine(model, B, cache_rows=T, storage16=True))
        for Tctx in (1024, 512):
            for _ in range(3):
                eng.set_position(Tctx - 1)
                eng.run_steps(1)
            L.zg_sync()
            sampler = ClockSampler(device_index)
            sampler.start()
            ts, e2e = [], []
            n0 = L.zg_launch_count()
            for _ in range(trials):
                L.zg_timer_begin()
                for _ in range(steps):
                    eng.set_position(Tctx - 1)
                    eng.run_steps(1)
                ts.append(L.zg_timer_end_ms())
            launches = int(L.zg_launch_count() - n0) // (trials * steps)
            for _ in range(trials):  # GPT.forward for 64 sequences with HOST token ids in, HOST next-token ids out
                L.zg_sync()
                t0 = time.perf_counter()
                for _ in range(steps):
                    eng.forward(Tctx, tok_host, 2)
                    L.zg_batch_read_tokens(eng._h, got.ctypes.data_as(lib.c_size_p))
                e2e.append(time.perf_counter() - t0)
            clocks = sampler.stop()
            lib.check()
            ms = float(np.median(ts)) / steps
            bytes_step = eng_step_bytes(cfg, Tctx, B, eng)
            hbm, _, _, kind = peaks()
            ach = bytes_step / (ms * 1e-3) / 1e9
            out[f"cfg4_{mode_name}_t{Tctx}"] = {
                "metric": "decode_tokens_per_sec", "value": B / (ms * 1e-3), "unit": "tok/s", "n_gpus": 1, "steps": steps,
                "ms_per_step": ms, "higher_is_better": True, "data": "synthetic",
                "dtype": (f"f32 storage, {mode_name} tensor-core GEMMs" if single is not None
                          else "f16 weight + KV storage (one-time copies), f16 tensor-core GEMMs, f32 accumulate"),
                "config": {"workload": f"GPT-2 {size}{' (4 layers)' if small else ''} batched decode, batch {B}, context {Tctx} "
                                       "(BASELINE configs[3])", "trials": trials,
                           "l2": "weights + KV read per step (46 GB at context 1024) far exceed L2"},
                "e2e": {"value": B * steps / float(np.median(e2e)), "unit": "tok/s", "h2d_bytes_per_step": B * 8,
                        "d2h_bytes_per_step": B * 8, "call": "zg_batch_forward(host token ids) + zg_batch_read_tokens"},
                "gpu_launches": launches,
                "roofline": {"bound": "hbm", "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm, "peak_kind": kind,
                             "bytes_per_launch": bytes_step, "traffic": None},
                "clocks": clocks}
        eng.close()
    model.close()
    return out


def eng_step_bytes(cfg, seq_len: int, batch: int, eng) -> int:
    """Algorithmic bytes of one batched decode step: weights once, KV rows, and the logits round trip
    unless the engine fuses the argmax into the lm_head epilogue."""
    fused = bool(getattr(eng, "fused_argmax", False))
    elem = 2 if getattr(eng, "storage_bits", 32) == 16 else 4
    b = cfg.decode_bytes(seq_len=seq_len, batch=batch, elem=elem, fused_argmax=fused)
    if not fused:
        b += 4 * batch * cfg.vocab_size  # logits written by the lm_head GEMM, read again by the argmax kernel
    return b


def measure_cfg5(L, lib, rank: int, world: int, dist, trials: int = 3, small: bool = False, device_index: int = 0,
                 both_prompt_modes: bool = True) -> dict:
    """BASELINE configs[4]: 1024 independent sequences sharded contiguously over the ranks (no collective), 32-token
    prompts, 256 greedy tokens each.  Timed through generate() with host prompts in / host tokens out; max over ranks."""
    from zig_gpt2_b200 import gpt as G
    from zig_gpt2_b200.batch import BatchEngine
    from zig_gpt2_b200.sharding import max_over_ranks, shard_range

    size = "124M"
    cfg = SIZES[size]
    S, n_in, n_new = (1024 if not small else 128), 32, 256
    n_total = n_in + n_new
    lo, hi = shard_range(S, world, rank)
    Bl = hi - lo
    model = G.gpt_from_numpy(cfg, synth_for_size(size))
    prompts = np.random.default_rng(1235).integers(0, cfg.vocab_size, (S, n_in))[lo:hi]

    def barrier():
        if dist is not None:
            dist.barrier()
        L.zg_sync()

    out = {}
    modes = (((True, "cfg5", False), (False, "cfg5_reference_prompt_loop", False), (True, "cfg5_tf32", True))
             if both_prompt_modes else ((True, "cfg5", False),))
    for use_prefill, key, tf32 in modes:
        eng = BatchEngine(model, Bl, cache_rows=n_total, max_prompt=n_in, tf32_single_pass=tf32)
        eng.generate_greedy(prompts, n_in + 8, use_prefill=use_prefill)  # warm-up (graph capture, clocks)
        sampler = ClockSampler(device_index)
        sampler.start()
        ts = []
        n0 = L.zg_launch_count()
        for _ in range(trials):
            barrier()
            t0 = time.perf_counter()
            toks = eng.generate_greedy(prompts, n_total, use_prefill=use_prefill)
            ts.append(time.perf_counter() - t0)
            barrier()
        launches = int(L.zg_launch_count() - n0) // trials
        clocks = sampler.stop()
        lib.check()
        sec = float(np.median(ts))
        if dist is not None:
            (sec,) = max_over_ranks(dist, [sec], device="cuda")
        first = n_in if use_prefill else 0
        bytes_total = sum(eng_step_bytes(cfg, s + 1, Bl, eng) for s in range(n_in, n_total))
        bytes_total += sum(cfg.decode_bytes(seq_len=s + 1, batch=Bl, fused_argmax=True) for s in range(first, n_in))
        hbm, _, _, kind = peaks()
        ach = bytes_total / sec / 1e9
        out[key] = {
            "metric": "decode_tokens_per_sec", "value": S * n_new / sec, "unit": "tok/s", "n_gpus": world, "steps": n_new,
            "ms_per_step": sec * 1e3 / n_new, "higher_is_better": True, "scaling": "strong",
            "dtype": ("f32 storage, single-pass TF32 tensor-core GEMMs (tolerance class 2e-2, not token-exact)" if tf32 else
                      "f32 storage, 3xTF32 tensor-core GEMMs") + (" (f16 prompt prefill)" if use_prefill else ""), "data": "synthetic",
            "config": {"workload": f"GPT-2 {size}, {S} independent synthetic sequences sharded over the GPUs, {n_in}-token prompts, "
                                   f"{n_new} greedy tokens each (BASELINE configs[4]); prompt "
                                   f"{'batched prefill' if use_prefill else 'token at a time (reference loop)'}",
                       "sequences_per_gpu": Bl, "parallelism": "replicated weights, contiguous sequence shards, no collective",
                       "trials": trials, "timing": "wall clock around generate() (host prompts in, host tokens out), max over ranks"},
            "e2e": {"value": S * n_new / sec, "unit": "tok/s", "h2d_bytes_per_step": Bl * n_in * 8 / n_new,
                    "d2h_bytes_per_step": Bl * n_total * 8 / n_new, "call": "zg_batch_generate_greedy (the timed call itself)"},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm, "peak_kind": kind,
                         "bytes_per_launch": bytes_total, "traffic": None, "note": "per GPU, whole generate() call"},
            "clocks": clocks, "tokens_tail": [int(t) for t in toks[0, -4:]]}
        eng.close()
    model.close()
    return out


def emit(line):
    print(json.dumps(line), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "bench_configs.jsonl"), "a") as f:
        f.write(json.dumps(line) + "\n")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("which", choices=["cfg3", "cfg4", "cfg5"])
    ap.add_argument("--small", action="store_true", help="reduced depth / sequence count (smoke runs)")
    ap.add_argument("--trials", type=int, default=5)
    ap.add_argument("--steps", type=int, default=8)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist

        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from zig_gpt2_b200 import lib

    L = lib.init(local_rank)
    if args.which == "cfg3":
        recs = {"cfg3": measure_cfg3(L, lib, args.trials, args.small, local_rank)}
    elif args.which == "cfg4":
        recs = measure_cfg4(L, lib, args.trials, args.steps, args.small, local_rank)
    else:
        recs = measure_cfg5(L, lib, rank, world, dist, args.trials, args.small, local_rank)
    if rank == 0:
        for k, r in recs.items():
            emit({"record": k, **r})
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
