"""Multi-GPU plumbing (one process per GPU, torch.distributed).  Round 1 shards CLIPS over ranks (independent replicas, no data-path
collective); the frame-sharded path (K/V exchange + GroupNorm-statistics all-reduce) builds on these helpers."""
from __future__ import annotations

import os
from typing import List, Tuple

import torch


def env_world() -> Tuple[int, int, int]:
    """(rank, world_size, local_rank) from the torchrun environment (1-process defaults)."""
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def init(backend: str, device: torch.device | None = None):
    import torch.distributed as dist
    rank, world, _ = env_world()
    if world > 1 and not dist.is_initialized():
        kw = {}
        if backend == "nccl" and device is not None:
            kw["device_id"] = device
        dist.init_process_group(backend, **kw)
    return rank, world


def barrier():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.barrier()


def max_over_ranks(value: float, device: torch.device) -> float:
    """Timing rule of the bench contract: a multi-GPU duration is the MAX over ranks."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def shard_frames(num_frames: int, world: int, rank: int) -> List[int]:
    """Contiguous frame block of `rank` (frames of one clip over the GPUs of one NVSwitch box)."""
    if num_frames % world != 0:
        raise ValueError(f"{num_frames} frames do not shard evenly over {world} ranks")
    per = num_frames // world
    return list(range(rank * per, (rank + 1) * per))


def shard_clips(num_clips: int, world: int, rank: int) -> List[int]:
    """Round-robin clip assignment for the replica mode."""
    return list(range(rank, num_clips, world))
