"""Multi-GPU sharding of independent frames.

Frames never depend on each other, so each rank decodes a contiguous range of the frame table, balanced
by compressed + decompressed bytes.  There is no collective on the data path: the only exchange is
a gather of the per-frame {status, decoded size} words (and, in bench.py, of the timings).
"""
from __future__ import annotations

import numpy as np


def partition(comp_sizes, out_sizes, world_size: int) -> list[tuple[int, int]]:
    """Contiguous [lo, hi) frame ranges, one per rank, balanced by sum(comp + out) bytes."""
    w = np.asarray(comp_sizes, dtype=np.float64) + np.asarray(out_sizes, dtype=np.float64)
    n = len(w)
    if world_size <= 1 or n == 0:
        return [(0, n)] + [(n, n)] * (max(world_size, 1) - 1)
    cum = np.concatenate([[0.0], np.cumsum(w)])
    total = cum[-1]
    bounds = [0]
    for r in range(1, world_size):
        target = total * r / world_size
        k = int(np.searchsorted(cum, target, side="left"))
        k = min(max(k, bounds[-1]), n)
        bounds.append(k)
    bounds.append(n)
    return [(bounds[r], bounds[r + 1]) for r in range(world_size)]


def gather_frame_results(status, out_size, ranges, rank: int, world_size: int, device=None):
    """Host-side gather of per-frame status / size words onto every rank (torch.distributed; works with
    gloo on CPU and nccl on GPUs).  Returns (status[u32], out_size[u64]) for the whole frame table."""
    import torch
    import torch.distributed as dist
    n_total = ranges[-1][1]
    if world_size == 1:
        return np.asarray(status, dtype=np.uint32), np.asarray(out_size, dtype=np.uint64)
    longest = max(hi - lo for lo, hi in ranges)
    dev = device if device is not None else "cpu"
    mine = torch.zeros(2, longest, dtype=torch.int64, device=dev)
    k = ranges[rank][1] - ranges[rank][0]
    if k:
        mine[0, :k] = torch.as_tensor(np.asarray(status, dtype=np.int64), device=dev)
        mine[1, :k] = torch.as_tensor(np.asarray(out_size, dtype=np.uint64).astype(np.int64), device=dev)
    parts = [torch.zeros_like(mine) for _ in range(world_size)]
    dist.all_gather(parts, mine)
    st = np.zeros(n_total, np.uint32)
    sz = np.zeros(n_total, np.uint64)
    for r, (lo, hi) in enumerate(ranges):
        p = parts[r].cpu().numpy()
        st[lo:hi] = p[0, :hi - lo].astype(np.uint32)
        sz[lo:hi] = p[1, :hi - lo].astype(np.uint64)
    return st, sz
