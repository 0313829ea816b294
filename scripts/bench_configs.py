#!/usr/bin/env python
"""BASELINE.json configs[2..4] (the batched, tensor-core paths) measured like bench.py measures configs[1].

    python scripts/bench_configs.py cfg3|cfg4|cfg5 [--small]
    torchrun --nproc-per-node N scripts/bench_configs.py cfg5      # sequences sharded over N GPUs, no collective

One JSON line per run (rank 0), appended to gpurun_out/bench_configs.jsonl.
  cfg3  GPT-2 355M, prefill of 16 x 1024-token synthetic prompts, last-position logits (tensor roofline)
  cfg4  GPT-2 1.5B, batch 64, one decode step at context 1024 (HBM roofline: weights + KV caches)
  cfg5  GPT-2 124M, 1024 sequences (split over the ranks), 32-token prompts, 256 greedy tokens each
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import ClockSampler  # noqa: E402
from zig_gpt2_b200.config import SIZES, GPTConfig  # noqa: E402
from zig_gpt2_b200.weights import synth_for_size, synth_weights  # noqa: E402


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), float(p["bf16_tflops"]), float(p.get("bf16_tflops_sustained", p["bf16_tflops"])), "measured"
    except Exception:
        return 6650.0, 1590.0, 1400.0, "fallback"


def emit(line):
    print(json.dumps(line), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "bench_configs.jsonl"), "a") as f:
        f.write(json.dumps(line) + "\n")


def cfg3(args, L, lib):
    from zig_gpt2_b200 import gpt as G
    from zig_gpt2_b200.batch import BatchEngine

    size = "355M"
    cfg = SIZES[size] if not args.small else GPTConfig(50257, 1024, 4, 16, 1024)
    B, T = (16, 1024)
    w = synth_for_size(size) if not args.small else synth_weights(cfg, seed=5)
    model = G.gpt_from_numpy(cfg, w)
    eng = BatchEngine(model, B, cache_rows=T, max_prompt=T)
    toks = np.random.default_rng(1236).integers(0, cfg.vocab_size, (B, T))
    for _ in range(3):
        eng.prefill(toks, True)
    L.zg_sync()
    sampler = ClockSampler(0)
    sampler.start()
    trials, e2e = [], []
    n0 = L.zg_launch_count()
    for _ in range(args.trials):
        L.zg_timer_begin()
        eng.prefill_resident(T, True)
        trials.append(L.zg_timer_end_ms())
    launches = int(L.zg_launch_count() - n0) // args.trials
    for _ in range(args.trials):
        L.zg_sync()
        t0 = time.perf_counter()
        eng.prefill(toks, True)
        lg = eng.logits()
        e2e.append(time.perf_counter() - t0)
    clocks = sampler.stop()
    ms = float(np.median(trials))
    flops = cfg.prefill_flops(B, T)
    hbm, tf_burst, tf_sus, kind = peaks()
    ach = flops / (ms * 1e-3) / 1e12
    emit({"metric": "prefill_tokens_per_sec", "value": B * T / (ms * 1e-3), "unit": "tok/s", "n_gpus": 1, "steps": 1,
          "ms_per_step": ms, "higher_is_better": True, "dtype": "f16 operands, f32 accumulate", "data": "synthetic",
          "config": {"workload": f"GPT-2 {size}{' (4 layers)' if args.small else ''} prefill, {B} x {T}-token synthetic prompts, last-position logits "
                                 "(BASELINE configs[2]); tcgen05 GEMMs + causal flash attention", "trials": args.trials},
          "e2e": {"value": B * T / float(np.median(e2e)), "unit": "tok/s", "h2d_bytes_per_step": B * T * 8,
                  "d2h_bytes_per_step": int(lg.nbytes), "call": "zg_batch_prefill(host tokens) + logits download"},
          "gpu_launches": launches,
          "roofline": {"bound": "tensor", "achieved": ach, "peak": tf_sus, "unit": "TFLOP/s", "frac": ach / tf_sus,
                       "frac_of_burst": ach / tf_burst, "peak_kind": kind, "flops_per_launch": flops, "traffic": None},
          "clocks": clocks})
    eng.close()
    model.close()


def cfg4(args, L, lib):
    from zig_gpt2_b200 import gpt as G
    from zig_gpt2_b200.batch import BatchEngine

    size = "1.5B"
    cfg = SIZES[size] if not args.small else GPTConfig(50257, 1024, 4, 25, 1600)
    B, T = 64, 1024
    t0 = time.time()
    w = synth_for_size(size) if not args.small else synth_weights(cfg, seed=5)
    model = G.gpt_from_numpy(cfg, w)
    del w
    print(f"# weights ready in {time.time() - t0:.1f}s", file=sys.stderr, flush=True)
    out = []
    for mode_name, single in (("3xTF32", False), ("TF32", True)):
        eng = BatchEngine(model, B, cache_rows=T, tf32_single_pass=single)
        for Tctx in (1024, 512):
            for _ in range(3):
                eng.set_position(Tctx - 1)
                eng.run_steps(1)
            L.zg_sync()
            sampler = ClockSampler(0)
            sampler.start()
            K = args.steps
            trials = []
            n0 = L.zg_launch_count()
            for _ in range(args.trials):
                L.zg_timer_begin()
                for _ in range(K):
                    eng.set_position(Tctx - 1)
                    eng.run_steps(1)
                trials.append(L.zg_timer_end_ms())
            launches = int(L.zg_launch_count() - n0) // args.trials
            clocks = sampler.stop()
            lib.check()
            ms = float(np.median(trials)) / K
            bytes_step = cfg.decode_bytes(seq_len=Tctx, batch=B, fused_argmax=False) + 4 * B * cfg.vocab_size  # logits written then read
            hbm, _, _, kind = peaks()
            ach = bytes_step / (ms * 1e-3) / 1e9
            emit({"metric": "decode_tokens_per_sec", "value": B / (ms * 1e-3), "unit": "tok/s", "n_gpus": 1, "steps": K,
                  "ms_per_step": ms, "higher_is_better": True, "dtype": f"f32 storage, {mode_name} tensor-core GEMMs", "data": "synthetic",
                  "config": {"workload": f"GPT-2 {size}{' (4 layers)' if args.small else ''} batched decode, batch {B}, context {Tctx} (BASELINE configs[3])",
                             "trials": args.trials, "l2": "46 GB of weights + KV per step, far larger than L2"},
                  "gpu_launches": launches,
                  "roofline": {"bound": "hbm", "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm, "peak_kind": kind,
                               "bytes_per_launch": bytes_step, "traffic": None},
                  "clocks": clocks})
        eng.close()
    model.close()


def cfg5(args, L, lib, rank, world, dist):
    from zig_gpt2_b200 import gpt as G
    from zig_gpt2_b200.batch import BatchEngine
    from zig_gpt2_b200.sharding import shard_range

    size = "124M"
    cfg = SIZES[size]
    S, n_in, n_new = (1024 if not args.small else 128), 32, 256
    n_total = n_in + n_new
    lo, hi = shard_range(S, world, rank)
    Bl = hi - lo
    model = G.gpt_from_numpy(cfg, synth_for_size(size))
    prompts_all = np.random.default_rng(1235).integers(0, cfg.vocab_size, (S, n_in))
    prompts = prompts_all[lo:hi]

    def barrier():
        if dist is not None:
            dist.barrier()
        L.zg_sync()

    for use_prefill in (False, True):
        eng = BatchEngine(model, Bl, cache_rows=n_total, max_prompt=n_in)
        eng.generate_greedy(prompts, n_in + 8, use_prefill=use_prefill)  # warm-up (graph capture, clocks)
        sampler = ClockSampler(int(os.environ.get("LOCAL_RANK", "0")))
        sampler.start()
        trials = []
        n0 = L.zg_launch_count()
        for _ in range(args.trials):
            barrier()
            t0 = time.perf_counter()
            toks = eng.generate_greedy(prompts, n_total, use_prefill=use_prefill)
            trials.append(time.perf_counter() - t0)
            barrier()
        launches = int(L.zg_launch_count() - n0) // args.trials
        clocks = sampler.stop()
        sec = float(np.median(trials))
        if dist is not None:
            from zig_gpt2_b200.sharding import max_over_ranks

            (sec,) = max_over_ranks(dist, [sec], device="cuda")
        first = n_in if use_prefill else 0
        bytes_total = sum(cfg.decode_bytes(seq_len=s + 1, batch=Bl, fused_argmax=False) + 4 * Bl * cfg.vocab_size for s in range(n_in, n_total))
        bytes_total += sum(cfg.decode_bytes(seq_len=s + 1, batch=Bl, fused_argmax=True) for s in range(first, n_in))
        hbm, _, _, kind = peaks()
        ach = bytes_total / sec / 1e9
        if rank == 0:
            emit({"metric": "decode_tokens_per_sec", "value": S * n_new / sec, "unit": "tok/s", "n_gpus": world, "steps": n_new,
                  "ms_per_step": sec * 1e3 / n_new, "higher_is_better": True, "scaling": "strong",
                  "dtype": "f32 storage, 3xTF32 tensor-core GEMMs" + (" (f16 prompt prefill)" if use_prefill else ""), "data": "synthetic",
                  "config": {"workload": f"GPT-2 {size}, {S} independent synthetic sequences sharded over {world} GPU(s) ({Bl} per GPU), {n_in}-token prompts, "
                                         f"{n_new} greedy tokens each (BASELINE configs[4]); prompt {'batched prefill' if use_prefill else 'token at a time (reference loop)'}",
                             "parallelism": "replicated weights, contiguous sequence shards, no collective", "trials": args.trials,
                             "timing": "wall clock around generate() (host prompts in, host tokens out), max over ranks"},
                  "e2e": {"value": S * n_new / sec, "unit": "tok/s", "h2d_bytes_per_step": Bl * n_in * 8 / n_new, "d2h_bytes_per_step": Bl * n_total * 8 / n_new},
                  "gpu_launches": launches,
                  "roofline": {"bound": "hbm", "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm, "peak_kind": kind,
                               "bytes_per_launch": bytes_total, "traffic": None, "note": "per GPU, whole generate() call"},
                  "clocks": clocks, "tokens_tail": [int(t) for t in toks[0, -4:]]})
        eng.close()
    model.close()


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
        cfg3(args, L, lib)
    elif args.which == "cfg4":
        cfg4(args, L, lib)
    else:
        cfg5(args, L, lib, rank, world, dist)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
