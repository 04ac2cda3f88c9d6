"""Batch sharding across the GPUs of one box: every variant treats images independently in
eval mode, so rank r of W simply owns images [lo, hi) and no collective sits on the hot path.  The only
collective offered is the optional result gather (NCCL all-gather over NVLink; gloo on CPU in the tests)."""
from __future__ import annotations

import torch


def shard_range(batch, rank, world):
    """Contiguous, balanced split: the first (batch % world) ranks get one extra image."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world {world}")
    base, rem = divmod(batch, world)
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def shard_batch(x, rank, world):
    lo, hi = shard_range(x.shape[0], rank, world)
    return x[lo:hi]


def gather_outputs(y_local, batch, group=None):
    """All-gather per-rank results back into the full batch order (optional; never inside the timed hot path)."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    sizes = [shard_range(batch, r, world) for r in range(world)]
    max_n = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((max_n,) + tuple(y_local.shape[1:]), dtype=y_local.dtype, device=y_local.device)
    pad[: y_local.shape[0]] = y_local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([p[: hi - lo] for p, (lo, hi) in zip(parts, sizes)], dim=0)
