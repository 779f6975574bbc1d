"""Multi-GPU plumbing (one process per GPU, torch.distributed): rendezvous, frame slicing / gathering of clip tensors, index helpers of the
frame-sharded forward.  The forward's own exchanges do NOT go through torch.distributed: they are peer-memory kernels (p2p.py, csrc/fz_p2p.cu);
`allreduce_set_sums` below is the reference semantics of the GroupNorm statistics exchange, kept for the CPU (gloo) tests."""
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


def frame_slice(x: torch.Tensor, rank: int, world: int, dim: int = 2) -> torch.Tensor:
    """This rank's contiguous frame block of a [B, C, F, H, W] clip tensor."""
    frames = shard_frames(x.shape[dim], world, rank)
    return x.narrow(dim, frames[0], len(frames)).contiguous()


def gather_frames(x_local: torch.Tensor, world: int, dim: int = 2, group=None) -> torch.Tensor:
    """Inverse of frame_slice on every rank (all-gather over the frame axis)."""
    import torch.distributed as dist
    if world <= 1:
        return x_local
    parts = [torch.empty_like(x_local) for _ in range(world)]
    dist.all_gather(parts, x_local.contiguous(), group=group)
    return torch.cat(parts, dim=dim)


def gathered_source_rows(frame_index: List[int], rank: int, world: int, frames_local: int, batch: int) -> List[int]:
    """Row blocks of the all-gathered K / V^T buffer ([world][batch][frames_local] frames) that the local query frames read:
    frame_index[g] is the GLOBAL source frame of global query frame g (engine.sc_frame_indices over the whole clip)."""
    nb = batch * frames_local
    out = []
    for b in range(batch):
        for f in range(frames_local):
            g = frame_index[rank * frames_local + f]
            out.append((g // frames_local) * nb + b * frames_local + g % frames_local)
    return out


def allreduce_set_sums(image_sums: torch.Tensor, frames_local: int, group=None) -> torch.Tensor:
    """GroupNorm statistics exchange: image_sums [batch * frames_local, G, 2] -> [batch, G, 2] summed over the local frames AND the ranks."""
    import torch.distributed as dist
    nb, g, two = image_sums.shape
    s = image_sums.view(nb // frames_local, frames_local, g, two).sum(1)
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(s, group=group)
    return s
