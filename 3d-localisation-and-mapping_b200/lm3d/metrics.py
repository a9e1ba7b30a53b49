"""Harness-side accounting (untimed): algorithmic bytes of a lift call."""
from __future__ import annotations

import torch


def union_pixels(rect4: torch.Tensor, frame_off: torch.Tensor, H: int, W: int, chunk: int | None = None) -> torch.Tensor:
    """``U_f`` = number of distinct pixels covered by >=1 rect of frame f, exactly, via a 2-D
    difference array per frame (int32 ``[chunk,H+1,W+1]`` on the rects' device)."""
    F = frame_off.numel() - 1
    dev = rect4.device
    B = rect4.shape[0]
    if chunk is None:  # ~256 MB of int32 difference array per chunk
        chunk = max(1, min(1024, (1 << 26) // ((H + 1) * (W + 1))))
    box_frame = torch.searchsorted(frame_off[1:].contiguous(), torch.arange(B, device=dev), right=True)
    U = torch.zeros(F, dtype=torch.int64, device=dev)
    r = rect4.long()
    for f0 in range(0, F, chunk):
        f1 = min(F, f0 + chunk)
        sel = (box_frame >= f0) & (box_frame < f1)
        if not bool(sel.any()):
            continue
        rr, ff = r[sel], box_frame[sel] - f0
        diff = torch.zeros((f1 - f0, H + 1, W + 1), dtype=torch.int32, device=dev)
        flat = diff.view(-1)
        S = (H + 1) * (W + 1)

        def add(y, x, v):
            flat.index_add_(0, ff * S + y * (W + 1) + x, torch.full_like(x, v, dtype=torch.int32))

        add(rr[:, 1], rr[:, 0], 1)
        add(rr[:, 1], rr[:, 2] + 1, -1)
        add(rr[:, 3] + 1, rr[:, 0], -1)
        add(rr[:, 3] + 1, rr[:, 2] + 1, 1)
        cover = diff.cumsum(1).cumsum(2)[:, :H, :W]
        U[f0:f1] = (cover > 0).sum(dim=(1, 2))
    return U


def algorithmic_bytes(rect4: torch.Tensor, frame_off: torch.Tensor, H: int, W: int) -> int:
    """``sum_f 4*U_f + 16*B_f + 48 + 96*B_f`` (each needed depth value read once, rects in,
    pose+intrinsics table, records out)."""
    F = frame_off.numel() - 1
    B = rect4.shape[0]
    U = union_pixels(rect4, frame_off, H, W)
    return int(4 * int(U.sum()) + 16 * B + 48 * F + 96 * B)
