"""Multi-GPU mode of this path: independent sequences sharded across GPUs, replicated weights, no collective on
the hot path.  Sequences never interact in the reference (one sequence per process, main.zig),
so a rank owns a contiguous slice of the sequence list and the only cross-rank traffic is bookkeeping after
the timed region: the max-over-ranks time and the gathered token ids."""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np


def shard_range(n_sequences: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous split: rank r of N owns sequences [r*S/N, (r+1)*S/N)."""
    if not 0 <= rank < world_size:
        raise ValueError("rank out of range")
    return (rank * n_sequences) // world_size, ((rank + 1) * n_sequences) // world_size


def max_over_ranks(dist, values: Sequence[float], device=None) -> List[float]:
    """Element-wise max of per-rank timings (the bench reports the slowest rank)."""
    import torch

    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t.cpu()]


def gather_token_ids(dist, local: np.ndarray, n_sequences: int, device=None) -> np.ndarray:
    """Concatenate per-rank [local_sequences, steps] token-id blocks in rank order (off the hot path)."""
    import torch

    local = np.ascontiguousarray(local, np.int64)
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    steps = local.shape[1]
    pad = (n_sequences + world - 1) // world
    buf = torch.full((pad, steps), -1, dtype=torch.int64, device=device)
    buf[: local.shape[0]] = torch.from_numpy(local).to(buf.device)
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf)
    rows = []
    for r, block in enumerate(out):
        lo, hi = shard_range(n_sequences, world, r)
        rows.append(block[: hi - lo].cpu().numpy())
    return np.concatenate(rows, axis=0)
