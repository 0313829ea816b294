"""The N>1 path of the bench on CPU: 2 ranks over gloo (127.0.0.1) exercise the sequence sharding, the
max-over-ranks timing reduction and the token gather that bench.py uses around the (GPU-only) hot path."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_partitions_exactly():
    from zig_gpt2_b200.sharding import shard_range

    for n in (0, 1, 7, 1024, 1025):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from zig_gpt2_b200.sharding import gather_token_ids, max_over_ranks, shard_range
    import bench

    n_seq, steps = 5, 7
    lo, hi = shard_range(n_seq, world, rank)
    local = np.array([[1000 * s + t for t in range(steps)] for s in range(lo, hi)], dtype=np.int64).reshape(hi - lo, steps)
    ms = max_over_ranks(dist, [10.0 + rank, 3.0 - rank])
    allt = gather_token_ids(dist, local, n_seq)
    prompts = bench.prompt_for(rank, 50257)
    q.put((rank, ms, allt.tolist(), prompts.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_gloo_sharding_timing_and_gather():
    world, port = 2, 29500 + (os.getpid() % 2000)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    results.sort()
    want = [[1000 * s + t for t in range(7)] for s in range(5)]
    for rank, ms, allt, prompts in results:
        assert ms == [11.0, 3.0]          # slowest rank wins, element-wise
        assert allt == want               # rank-ordered concatenation of the shards
    assert results[0][3] != results[1][3]  # every rank decodes its own synthetic sequence
