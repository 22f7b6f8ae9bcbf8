"""Frame sharding across the GPUs of one box.

The path shards by independent units: every (frame, bbox) is independent in 2D (reference loop
pose_pipeline/wrappers/mmpose.py:60-76 carries nothing between iterations) and every 3D output frame
depends only on a 243-frame window of tiny 2D keypoints.  So there is NO data-path collective: rank r
decodes and stages its own contiguous frame range from host, and the only exchange is the gather of the
(N_r, K, 3) float32 result rows (204 B per frame) -- torch.distributed all_gather, NCCL over NVLink on the
GPU box, gloo in the CPU tests.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous range [start, stop) of rank `rank`: ceil(n/world) frames per rank, last ranks may be short/empty."""
    per = (n + world - 1) // world if world > 0 else n
    start = min(n, rank * per)
    return start, min(n, start + per)


def dist_info() -> Tuple[int, int]:
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
    except Exception:
        pass
    return 0, 1


def gather_rows(local: np.ndarray, n_total: int, device=None) -> np.ndarray:
    """All ranks contribute their contiguous row block (shard_range order); every rank gets the (n_total, ...) array."""
    rank, world = dist_info()
    if world == 1:
        return local
    import torch
    import torch.distributed as dist
    per = (n_total + world - 1) // world
    tail = local.shape[1:]
    buf = np.zeros((per,) + tail, np.float64 if local.dtype == np.float64 else np.float32)
    buf[: local.shape[0]] = local
    dev = device if device is not None else ("cuda" if dist.get_backend() == "nccl" else "cpu")
    t = torch.from_numpy(buf).to(dev)
    outs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(outs, t)
    full = np.concatenate([o.cpu().numpy() for o in outs], axis=0)[:n_total]
    return full


def max_over_ranks(value: int) -> int:
    """max of an integer over the ranks (1 rank: the value itself)."""
    rank, world = dist_info()
    if world == 1:
        return int(value)
    import torch
    import torch.distributed as dist
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([int(value)], dtype=torch.int64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return int(t.item())
