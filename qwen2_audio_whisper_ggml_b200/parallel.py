"""Data-parallel sharding of independent 30 s windows over the GPUs of one box.

One process per GPU (torch.distributed plumbing), full weight replica per rank, window w -> rank floor(w * G / B)
(contiguous blocks), NO collective on the compute path.  Only when the caller asks for all embeddings on one
rank is a gather issued (NCCL on GPUs, gloo in the CPU tests), ordered by window index.
"""
from __future__ import annotations

import numpy as np


def shard_bounds(n_windows: int, rank: int, world: int) -> tuple[int, int]:
    """[start, end) of the windows rank owns: w belongs to rank floor(w * world / n_windows)"""
    if world < 1 or not (0 <= rank < world) or n_windows < 0:
        raise ValueError("bad shard arguments")
    start = -(-rank * n_windows // world)
    end = -(-(rank + 1) * n_windows // world)
    return start, min(end, n_windows)


def shard_sizes(n_windows: int, world: int) -> list[int]:
    return [shard_bounds(n_windows, r, world)[1] - shard_bounds(n_windows, r, world)[0] for r in range(world)]


def encode_sharded(encode_fn, windows, n_samples=None, gather_to: int | None = None, group=None):
    """Encode this rank's shard with encode_fn(windows[start:end], n_samples[start:end]) -> [b, n_out, n_state] (numpy or torch).

    Returns (local_result, (start, end)) or, when gather_to is given, (all windows in order on rank gather_to else None, bounds).
    """
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    B = len(windows)
    s, e = shard_bounds(B, rank, world)
    local = encode_fn(windows[s:e], None if n_samples is None else n_samples[s:e]) if e > s else None
    if gather_to is None or world == 1:
        return local, (s, e)
    sizes = shard_sizes(B, world)
    mx = max(sizes)
    is_np = isinstance(local, np.ndarray) or local is None
    # shape of one window's result, agreed through a tiny all_reduce so empty shards can take part
    shp = torch.zeros(2, dtype=torch.int64)
    if local is not None:
        shp[0], shp[1] = int(local.shape[1]), int(local.shape[2])
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    shp = shp.to(dev)
    dist.all_reduce(shp, op=dist.ReduceOp.MAX, group=group)
    n_out, n_state = int(shp[0]), int(shp[1])
    pad = torch.zeros((mx, n_out, n_state), dtype=torch.float32, device=dev)
    if local is not None:
        t = torch.from_numpy(local) if isinstance(local, np.ndarray) else local
        pad[: e - s] = t.to(dev)
    parts = [torch.empty_like(pad) for _ in range(world)] if rank == gather_to else None
    dist.gather(pad, parts, dst=gather_to, group=group)
    if rank != gather_to:
        return None, (s, e)
    out = torch.cat([parts[r][: sizes[r]] for r in range(world)], dim=0)
    return (out.cpu().numpy() if is_np else out), (s, e)
