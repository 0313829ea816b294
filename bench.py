#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric on BASELINE.json's config.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (`configs[1]`): GPT-2 124M, random-init fp32 weights, batch-1 greedy decode with KV cache, one
synthetic 16-token prompt per GPU.  A *step* is one decoded token (one pass of GPT.forward + argmax,
main.zig:198-207).  Prompt fill and W warm-up tokens are untimed; exactly K tokens are timed.

  value     decode tokens/s with everything resident in HBM, CUDA events on the launching stream around
            the K-step launch (max over ranks; N ranks decode N independent sequences -> weak scaling).
  e2e       the same metric through the reference-facing per-token call, exactly the loop the reference arm times:
            K calls of GPT.sample (zg_engine_sample_greedy: HOST token in, one launch, D2H of the chosen token, stream
            synchronise -- all inside the timed region), wall clock.  `e2e.generate_call` adds the one-call form
            (zg_engine_generate_greedy, host prompt -> host tokens, prompt fill inside the timed region).
  roofline  achieved = algorithmic bytes per K-step launch / its CUDA-event duration,
            against MEASURED_PEAKS.json's HBM copy bandwidth.  `traffic` is null in the line: DRAM counters cannot be
            read outside a profiler; the `ncu --set full` capture of the same kernel is profiles/r02_decode_persistent_*.
  cpu_baseline  the oracle (C restatement of the reference + OpenBLAS, every host thread) on cfg 1.
  configs   full sub-records for the other BASELINE configs, measured in the same run (scripts/bench_configs.py):
            N = 1: cfg3 (355M prefill 16 x 1024), cfg4 (1.5B batch-64 decode, TF32 and 3xTF32, context 1024 / 512), cfg5;
            N > 1 (torchrun): cfg5 only -- the 1024 sequences split 1024/N per rank, "scaling": "strong".
            `--no-configs` skips them (the headline keys are unaffected either way).

`--impl reference` times the reference's CPU path (the oracle port: no Zig toolchain exists to build the
reference itself) on the host cores for the same K/W.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from zig_gpt2_b200.config import SIZES  # noqa: E402
from zig_gpt2_b200.weights import synth_for_size  # noqa: E402

sys.path.insert(0, os.path.join(ROOT, "scripts"))
from bench_configs import ClockSampler  # noqa: E402

SIZE = "124M"
N_PROMPT = 16
METRIC = "decode_tokens_per_sec"
UNIT = "tok/s"
FALLBACK_HBM_GBS = 6650.0
# the same string on both arms (ours and --impl reference): the driver compares it
WORKLOAD = (f"GPT-2 {SIZE} random-init fp32, batch-1 greedy decode with KV cache, {N_PROMPT}-token synthetic prompt "
            "(BASELINE configs[1]); one step = one decoded token")


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback"


def prompt_for(rank: int, vocab: int) -> np.ndarray:
    return np.random.Generator(np.random.PCG64(1235 + rank)).integers(0, vocab, N_PROMPT).astype(np.uint64)


def cpu_reference(steps: int, warmup: int, threads=None):
    """The reference's CPU path (oracle port + OpenBLAS) on the same workload: prompt fill, `warmup` untimed
    tokens, `steps` timed greedy tokens.  Returns (tok/s, ms/step, cores, kind-of-blas)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import zg_oracle as zo

    cores = threads or len(os.sched_getaffinity(0))
    blas = "openblas" if zo.use_openblas(cores) else "scalar"
    cfg = SIZES[SIZE]
    m = zo.Model(cfg, synth_for_size(SIZE))
    p = prompt_for(0, cfg.vocab_size)
    token = 0
    for s in range(N_PROMPT):  # main.zig:330-334
        token = int(p[s])
        m.forward(s + 1, token, False)
    import ctypes as C

    L = zo.lib()
    seq = N_PROMPT
    for _ in range(warmup):
        token = int(L.zo_gpt_sample_greedy(C.byref(m.gpt), seq + 1, token, C.byref(m.state)))
        seq += 1
    t0 = time.perf_counter()
    for _ in range(steps):
        token = int(L.zo_gpt_sample_greedy(C.byref(m.gpt), seq + 1, token, C.byref(m.state)))
        seq += 1
    dt = time.perf_counter() - t0
    m.close()
    return steps / dt, dt / steps * 1e3, cores, blas


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    steps = min(args.steps, 1024 - N_PROMPT - args.warmup)
    tps, ms, cores, blas = cpu_reference(steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": tps, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD,
                   "note": "reference CPU path = line-for-line C port of src/ops.zig+main.zig (no Zig toolchain in this image) + " + blas},
        "cpu_baseline": {"value": tps, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{steps} greedy tokens after a {N_PROMPT}-token prompt and {args.warmup} warm-up tokens"},
        "e2e": {"value": tps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


def run_ours(args):
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist

        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from zig_gpt2_b200 import gpt as G
    from zig_gpt2_b200 import lib

    L = lib.init(local_rank)
    cfg = SIZES[SIZE]
    K, W = args.steps, max(3, args.warmup)
    K = min(K, cfg.context_size - N_PROMPT - W)
    model = G.gpt_from_numpy(cfg, synth_for_size(SIZE))
    state = G.State(cfg)
    eng = model.engine(state)
    prompt = prompt_for(rank, cfg.vocab_size)
    pp = prompt.ctypes.data_as(lib.c_size_p)
    first = N_PROMPT + W

    def barrier():
        if dist is not None:
            dist.barrier()
        L.zg_sync()

    # untimed: prompt fill (one token at a time, no logits) + W warm-up tokens, then ramp clocks for ~0.3 s
    L.zg_engine_set_prompt(eng, pp, N_PROMPT)
    L.zg_engine_run_steps(eng, 0, first)
    L.zg_sync()
    lib.check()
    t_end = time.perf_counter() + 0.3
    while time.perf_counter() < t_end:
        L.zg_engine_run_steps(eng, first, K)
        L.zg_sync()

    sampler = ClockSampler(local_rank)
    sampler.start()
    trials = []
    launches0 = L.zg_launch_count()
    for _ in range(args.trials):
        barrier()
        L.zg_timer_begin()
        L.zg_engine_run_steps(eng, first, K)  # exactly K decode steps, inputs resident in HBM
        ms = L.zg_timer_end_ms()
        barrier()
        trials.append(ms)
    launches = int(L.zg_launch_count() - launches0) // max(1, args.trials)
    lib.check()
    ms_total = float(np.median(trials))

    # end to end, one call: generate() with host prompt in / host tokens out (prompt fill inside the timed region)
    out = np.zeros(first + K, np.uint64)
    gen_trials = []
    for _ in range(args.trials):
        barrier()
        t0 = time.perf_counter()
        rc = L.zg_engine_generate_greedy(eng, pp, N_PROMPT, first + K, out.ctypes.data_as(lib.c_size_p))
        gen_trials.append(time.perf_counter() - t0)
        if rc:
            lib.check()
            raise RuntimeError(f"generate failed: {rc}")
    gen_s = float(np.median(gen_trials))
    tokens = out.astype(np.int64)

    # end to end, per token: the reference arm's loop (cpu_reference above) through the C-ABI -- prompt fill and W
    # warm-up tokens untimed, then K calls of GPT.sample with a HOST token in and the chosen token read back to the host
    e2e_trials = []
    for _ in range(args.trials):
        token = 0
        for s_ in range(N_PROMPT):  # main.zig:330-334
            token = int(prompt[s_])
            L.zg_engine_forward(eng, s_ + 1, token, 0)
        seq = N_PROMPT
        for _w in range(W):
            token = int(L.zg_engine_sample_greedy(eng, seq + 1, token))
            seq += 1
        barrier()
        t0 = time.perf_counter()
        for _k in range(K):
            token = int(L.zg_engine_sample_greedy(eng, seq + 1, token))
            seq += 1
        e2e_trials.append(time.perf_counter() - t0)
    lib.check()
    e2e_s = float(np.median(e2e_trials))
    clocks = sampler.stop()

    if dist is not None:
        from zig_gpt2_b200.sharding import max_over_ranks

        ms_total, e2e_s, gen_s = max_over_ranks(dist, [ms_total, e2e_s, gen_s], device="cuda")

    bytes_per_launch = sum(cfg.decode_bytes(seq_len=s + 1) for s in range(first, first + K))
    peak, peak_kind = peaks()
    achieved = bytes_per_launch / (float(np.median(trials)) * 1e-3) / 1e9
    value = world * K / (ms_total * 1e-3)
    e2e_value = world * K / e2e_s
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {
            "workload": WORKLOAD, "positions_timed": [first, first + K - 1],
            "sequences": world, "parallelism": f"{world} independent sequence(s), one per GPU, replicated weights, no collective",
            "l2": "inputs larger than L2 (495 MB of weights streamed per token vs 126 MB L2)",
            "trials": args.trials, "timing": "median of trials; CUDA events on the launching stream; max over ranks",
            "weights_init": "N(0,(0.1*sqrt(768/E))^2) linears, N(0,0.05^2) embeddings, N(0,0.02^2) biases (zig_gpt2_b200/weights.py)",
        },
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 8, "d2h_bytes_per_step": 8,
                "call": "K x zg_engine_sample_greedy(seq_len, host token) -> host token (GPT.sample, main.zig:198-207): one launch, "
                        "D2H of the token and a stream synchronise per step; the loop the reference arm times",
                "generate_call": {"value": world * (W + K) / gen_s, "unit": UNIT,
                                  "call": "zg_engine_generate_greedy(host prompt -> host tokens): prompt fill + W + K tokens in "
                                          "one call; generated tokens / wall time of the whole call"}},
        "gpu_launches": launches,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": None, "peak_kind": peak_kind, "kernel": "decode_persistent_kernel",
                     "traffic_note": "not measurable in-run; ncu --set full of this kernel: profiles/r02_decode_persistent_ncu_full.csv",
                     "bytes_per_launch": bytes_per_launch, "frac_of_nominal_8TBs": achieved / 8000.0},
        "clocks": clocks,
        "tokens_tail": [int(t) for t in tokens[-4:]],
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_steps = 48
        tps, ms, cores, blas = cpu_reference(cpu_steps, 2)
        line["cpu_baseline"] = {"value": tps, "unit": UNIT, "cores": cores, "kind": "port", "ms_per_step": ms,
                                "sample": f"{cpu_steps} greedy tokens after a {N_PROMPT}-token prompt + 2 warm-up tokens; "
                                          f"C port of the reference + {blas}"}
    if not args.no_configs:
        # free the batch-1 engine before the large configs take the memory
        model.close()
        del eng, model, state
        line["configs"] = other_configs(args, L, lib, rank, local_rank, world, dist)
    if rank == 0:
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def other_configs(args, L, lib, rank, local_rank, world, dist):
    """BASELINE configs[2..4] as sub-records.  cfg5 shards over the ranks; cfg3 / cfg4 are one-GPU configs and run at N = 1."""
    import bench_configs as BC

    out = {}

    def guarded(name, fn):
        try:
            r = fn()
            out.update(r if name is None else {name: r})
        except Exception as e:  # a failing sub-bench must not take the headline line with it
            lib.load().zg_clear_error()
            out[name or "cfg"] = {"error": f"{type(e).__name__}: {e}"[:300]}

    guarded(None, lambda: BC.measure_cfg5(L, lib, rank, world, dist, trials=3, device_index=local_rank,
                                          both_prompt_modes=(world == 1)))
    if world == 1:
        guarded("cfg3", lambda: BC.measure_cfg3(L, lib, trials=5, device_index=local_rank))
        guarded(None, lambda: BC.measure_cfg4(L, lib, trials=3, steps=4, device_index=local_rank))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=64)  # SURVEY 8d cfg 2: 64 greedy tokens after a 16-token prompt
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--trials", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the cfg3 / cfg4 / cfg5 sub-records")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
