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


class _DevBuffer:
    """A raw device allocation exposed through ``__cuda_array_interface__`` so torch can view it."""

    def __init__(self, ptr: int, nbytes: int):
        self.ptr, self.nbytes = int(ptr), int(nbytes)
        self.__cuda_array_interface__ = {"shape": (self.nbytes,), "typestr": "|u1", "data": (self.ptr, False), "version": 2}


class PeerGather:
    """Record gather FUSED into the lift (``lm3d_lift_boxes_gather``): no collective call.

    Every rank owns ``slots`` gather buffers of ``world * n_records`` records (``cudaMalloc`` + CUDA IPC handle,
    exchanged once through the process group); each rank maps all its peers' buffers (NVLink peer access) and hands
    the lift a table of ``world`` pointers per slot.  The lift kernels then store every finished record to all of
    them at ``rank * n_records + b`` -- the all-gather happens as a side effect of the kernel epilogues, spread over
    the kernel's duration, without a communication kernel competing for the SMs of a persistent grid (what limited
    the NCCL path: ``PipelinedGather`` below, kept as the baseline and as the parity reference).

    The stores of a call are complete when its stream reaches the end of the call on every rank: ``barrier()``
    (stream sync + process-group barrier) before reading ``buffer(slot)``.  Equal ``n_records`` on every rank."""

    def __init__(self, n_records: int, device, slots: int = 2, group=None):
        import ctypes as C

        from . import _capi

        self.lib = _capi.load()
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.n_records = int(n_records)
        self.device = torch.device(device)
        self.slots = int(slots)
        nbytes = self.world * self.n_records * 96
        self._own, self._opened, handles = [], [], []
        with torch.cuda.device(self.device):
            for _ in range(self.slots):
                ptr = C.c_void_p()
                h = C.create_string_buffer(64)
                _capi.check(self.lib.lm3d_gather_alloc(nbytes, C.byref(ptr), h), "lm3d_gather_alloc")
                self._own.append(ptr.value)
                handles.append(h.raw)
            everyone = [None] * self.world
            dist.all_gather_object(everyone, handles, group=group)
            self.tables = []  # per slot: ctypes array of `world` device pointers (index = destination rank)
            for s in range(self.slots):
                tab = (C.c_void_p * self.world)()
                for r in range(self.world):
                    if r == self.rank:
                        tab[r] = self._own[s]
                    else:
                        ptr = C.c_void_p()
                        _capi.check(self.lib.lm3d_gather_open(everyone[r][s], C.byref(ptr)), "lm3d_gather_open")
                        self._opened.append(ptr.value)
                        tab[r] = ptr.value
                self.tables.append(tab)
        self._views = [torch.as_tensor(_DevBuffer(p, nbytes), device=self.device).view(torch.float32).view(self.world * self.n_records, 24)
                       for p in self._own]

    @property
    def box_offset(self) -> int:
        return self.rank * self.n_records

    def buffer(self, slot: int) -> torch.Tensor:
        """This rank's gathered ``[world * n_records, 24]`` float32 view of ``slot`` (valid after ``barrier()``)."""
        return self._views[slot]

    def barrier(self):
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)

    def close(self):
        with torch.cuda.device(self.device):
            self.barrier()
            for p in self._opened:
                self.lib.lm3d_gather_close(p)
            self.barrier()
            for p in self._own:
                self.lib.lm3d_gather_free(p)
        self._opened, self._own, self._views = [], [], []


class PipelinedGather:
    """Overlap the record all-gather of step ``i`` with the lift of step ``i+1``.

    The gather runs on the process group's own stream (``async_op=True``) into one of two
    output buffers; the caller alternates two ``LiftPlan`` record buffers the same way and calls
    ``ready(slot)`` before reusing a slot.  At 8 GPUs the gather of a C2 shard (19.2 MB out,
    134 MB in per rank) costs ~0.3 ms against ~2 ms of lift, so hiding it is what takes the
    weak-scaling efficiency from 6.9x to the compute-only figure."""

    def __init__(self, n_records: int, device, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.out = [torch.empty((self.world * n_records, 24), dtype=torch.float32, device=device) for _ in range(2)]
        self.pending = [None, None]

    def ready(self, slot: int):
        """Make the current stream wait until the gather that last used ``slot`` has finished."""
        if self.pending[slot] is not None:
            self.pending[slot].wait()
            self.pending[slot] = None

    def launch(self, slot: int, records: torch.Tensor):
        if self.world == 1:
            return
        self.pending[slot] = dist.all_gather_into_tensor(self.out[slot], records, group=self.group, async_op=True)

    def drain(self):
        for s in (0, 1):
            self.ready(s)
