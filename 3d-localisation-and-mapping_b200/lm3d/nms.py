"""3-D non-maximum suppression over lifted boxes (``lm3d_nms_boxes``).

Replaces ``BoundingBoxProcessor.suppress_bboxes`` (``/root/reference/task_def.py:145-149``).  The class's source is
not in the reference repository; the rules are NMS-SPEC v0 (``DESIGN.md`` 4.8).  No CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _capi
from .lift import _stream_ptr

DEFAULT_IOU_THR = 0.1
DEFAULT_PAD_M = 0.03  # the reference's bbox_depth_buffer (pose_processor.py:50)


def nms_boxes(corners: torch.Tensor, conf: torch.Tensor, label: torch.Tensor, iou_thr: float = DEFAULT_IOU_THR,
              pad_m: float = DEFAULT_PAD_M, want_parent: bool = True):
    """corners: CUDA float32 ``[B,24]`` records (``lift_boxes`` output) or packed ``[B,12]`` / ``[B,4,3]``; conf
    float32 ``[B]``; label int32 ``[B]``.  Returns ``(keep uint8[B], parent int32[B] | None, rounds)``.
    Synchronises the current stream."""
    lib = _capi.load()
    if corners.dtype != torch.float32 or not corners.is_cuda or not corners.is_contiguous():
        raise ValueError("corners must be a contiguous CUDA float32 tensor")
    B = corners.shape[0]
    stride = corners.numel() // B if B else 12
    if B and stride not in (12, _capi.RECORD_WORDS):
        raise ValueError("corners must be [B,24] records or [B,12] / [B,4,3] packed corners")
    for t, name, dt in ((conf, "conf", torch.float32), (label, "label", torch.int32)):
        if t.dtype != dt or not t.is_cuda or not t.is_contiguous() or t.shape != (B,) or t.device != corners.device:
            raise ValueError(f"{name} must be a contiguous CUDA {dt} tensor of shape [B] on the corners' device")
    dev = corners.device
    keep = torch.empty((B,), dtype=torch.uint8, device=dev)
    parent = torch.empty((B,), dtype=torch.int32, device=dev) if want_parent else None
    ws_bytes = int(lib.lm3d_nms_workspace_bytes(B))
    ws = torch.empty((max(ws_bytes, 16),), dtype=torch.uint8, device=dev)
    rounds = C.c_int32(0)
    with torch.cuda.device(dev):
        st = lib.lm3d_nms_boxes(corners.data_ptr(), stride, conf.data_ptr(), label.data_ptr(), B, float(iou_thr),
                                float(pad_m), keep.data_ptr(), parent.data_ptr() if parent is not None else None,
                                C.byref(rounds), ws.data_ptr(), ws.numel(), _stream_ptr(dev))
    _capi.check(st, "lm3d_nms_boxes")
    return keep, parent, int(rounds.value)
