"""Frame sharding across ranks + gather of the per-box records.

Frames are independent (``/root/reference/src/mapper/pose_processor.py:91-115`` carries no
cross-frame state), so rank ``r`` of ``R`` lifts the contiguous frame range
``[floor(r*F/R), floor((r+1)*F/R))`` with zero input exchange; the only collective is an
all-gather of the 96-byte records (NCCL over NVLink on the GPU box, gloo in the CPU tests).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def shard_range(F: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous frame shard of ``rank``."""
    return (rank * F) // world, ((rank + 1) * F) // world


def shard_boxes(frame_off, rank: int, world: int):
    """``(f0, f1, b0, b1, local_frame_off)`` for this rank's CSR slice."""
    F = len(frame_off) - 1
    f0, f1 = shard_range(F, rank, world)
    b0, b1 = int(frame_off[f0]), int(frame_off[f1])
    local = frame_off[f0 : f1 + 1] - frame_off[f0]
    return f0, f1, b0, b1, local


def all_gather_records(records: torch.Tensor, counts: list[int] | None = None, group=None) -> torch.Tensor:
    """All-gather ``[B_r,24]`` float32 record tensors into ``[sum B_r,24]`` in rank order.

    Equal ``B_r`` on every rank (the synthetic configs) is a single
    ``all_gather_into_tensor``; ragged shards pass ``counts`` (boxes per rank) and are padded
    to the maximum, then trimmed -- record order equals the single-GPU order either way."""
    world = dist.get_world_size(group)
    if world == 1:
        return records
    B, Wd = records.shape
    if counts is None:
        out = torch.empty((world * B, Wd), dtype=records.dtype, device=records.device)
        dist.all_gather_into_tensor(out, records.contiguous(), group=group)
        return out
    mx = max(counts)
    pad = records
    if B < mx:
        pad = torch.zeros((mx, Wd), dtype=records.dtype, device=records.device)
        pad[:B] = records
    out = torch.empty((world * mx, Wd), dtype=records.dtype, device=records.device)
    dist.all_gather_into_tensor(out, pad.contiguous(), group=group)
    return torch.cat([out[r * mx : r * mx + c] for r, c in enumerate(counts)], dim=0)


def gather_counts(n_local: int, device, group=None) -> list[int]:
    world = dist.get_world_size(group)
    t = torch.tensor([n_local], dtype=torch.int64, device=device)
    out = torch.empty((world,), dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(out, t, group=group)
    return [int(v) for v in out.cpu()]
