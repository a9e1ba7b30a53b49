"""Torch-facing wrappers over the C ABI (``include/lm3d.h``).

PyTorch is plumbing here (device memory, streams); the work is done by the hand-written
sm_100a kernels in ``csrc/`` reached through ``ctypes``.  Nothing in this module computes
the lift on the CPU or in torch ops -- a missing ``liblm3d.so`` raises.

Replaces, for whole sequences, the per-frame / per-box Python loops of
``ProcessPose._3d_processing`` (``/root/reference/src/mapper/pose_processor.py:124-240``).
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import _capi

#: numpy view of ``lm3d_box_out`` (96 bytes)
RECORD_DTYPE = np.dtype(
    [
        ("corners", "<f4", (4, 3)),
        ("centroid", "<f4", (3,)),
        ("aabb_min", "<f4", (3,)),
        ("aabb_max", "<f4", (3,)),
        ("z_q", "<f4"),
        ("n_valid", "<i4"),
        ("n_pix", "<i4"),
    ]
)
assert RECORD_DTYPE.itemsize == _capi.RECORD_BYTES


def _chk(t: torch.Tensor, name: str, dtype, ndim=None, cuda=True):
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name}: expected torch.Tensor, got {type(t).__name__}")
    if t.dtype != dtype:
        raise TypeError(f"{name}: expected dtype {dtype}, got {t.dtype}")
    if cuda and not t.is_cuda:
        raise ValueError(f"{name}: must be a CUDA tensor (no CPU fallback exists)")
    if not t.is_contiguous():
        raise ValueError(f"{name}: must be contiguous")
    if ndim is not None and t.dim() != ndim:
        raise ValueError(f"{name}: expected {ndim} dims, got {t.dim()}")
    return t


def _stream_ptr(device) -> int:
    return int(torch.cuda.current_stream(device).cuda_stream)


def workspace_bytes(F: int, B: int, H: int | None = None, W: int | None = None) -> int:
    """Workspace of ``lm3d_lift_boxes``: the minimum for F frames / B boxes, or -- with the frame shape -- the
    size that also holds the tile-pyramid scratch (``lm3d_lift_workspace_bytes``)."""
    lib = _capi.load()
    if H is None or W is None:
        return int(lib.lm3d_workspace_bytes(int(F), int(B)))
    return int(lib.lm3d_lift_workspace_bytes(int(F), int(H), int(W), int(B)))


def scale_boxes(boxes_xyxy, image_wh, frame_off, depth_w: int, depth_h: int, out=None):
    """Detector boxes (RGB px, fp64 ``[B,4]``) -> inclusive clamped int32 rects ``[B,4]`` at
    depth resolution (``lm3d_scale_boxes``; reference ``pose_processor.py:174-181,186-187``)."""
    lib = _capi.load()
    _chk(boxes_xyxy, "boxes_xyxy", torch.float64, 2)
    _chk(image_wh, "image_wh", torch.float64, 2)
    _chk(frame_off, "frame_off", torch.int64, 1)
    B, F = boxes_xyxy.shape[0], frame_off.shape[0] - 1
    if boxes_xyxy.shape[1] != 4 or image_wh.shape != (F, 2):
        raise ValueError("boxes_xyxy must be [B,4] and image_wh [F,2]")
    if out is None:
        out = torch.empty((B, 4), dtype=torch.int32, device=boxes_xyxy.device)
    _chk(out, "out", torch.int32, 2)
    with torch.cuda.device(boxes_xyxy.device):
        st = lib.lm3d_scale_boxes(
            boxes_xyxy.data_ptr(), image_wh.data_ptr(), frame_off.data_ptr(), F, B, int(depth_w), int(depth_h),
            out.data_ptr(), _stream_ptr(boxes_xyxy.device),
        )
    _capi.check(st, "lm3d_scale_boxes")
    return out


class LiftPlan:
    """Pre-allocated output + workspace for repeated ``lift_boxes`` calls of one shape."""

    def __init__(self, F: int, B: int, device, order_stats: bool = False, H: int | None = None, W: int | None = None):
        self.F, self.B = int(F), int(B)
        self.device = torch.device(device)
        self.records = torch.empty((self.B, _capi.RECORD_WORDS), dtype=torch.float32, device=self.device)
        self.order_stats = (
            torch.empty((self.B, 2), dtype=torch.float32, device=self.device) if order_stats else None
        )
        self.ws_bytes = workspace_bytes(self.F, self.B, H, W)
        self.workspace = torch.empty((max(self.ws_bytes, 16),), dtype=torch.uint8, device=self.device)


def lift_boxes(
    depth,
    pose7,
    intr4,
    rect4,
    frame_off,
    scale_depth: float = 1000.0,
    max_depth_mm: float = math.inf,
    q: float = 50.0,
    plan: LiftPlan | None = None,
    order_stats: bool = False,
    gather=None,
    gather_slot: int = 0,
):
    """Lift every box of a sequence (``lm3d_lift_boxes``; with ``gather`` -- an ``lm3d.dist.PeerGather`` -- the
    multi-GPU entry ``lm3d_lift_boxes_gather``, which also stores every record into slot ``gather_slot`` of every
    rank's gather buffer).

    depth ``[F,H,W]`` f32 mm, pose7 ``[F,7]`` f64, intr4 ``[F,4]`` f64 (depth resolution),
    rect4 ``[B,4]`` i32, frame_off ``[F+1]`` i64 -- all CUDA, contiguous.  Returns the
    records as a ``[B,24]`` float32 tensor (``records_to_numpy`` gives the struct view), plus
    the ``[B,2]`` order statistics when ``order_stats`` is set.  Asynchronous on the current
    stream."""
    lib = _capi.load()
    _chk(depth, "depth", torch.float32, 3)
    F, H, W = depth.shape
    _chk(pose7, "pose7", torch.float64, 2)
    _chk(intr4, "intr4", torch.float64, 2)
    _chk(rect4, "rect4", torch.int32, 2)
    _chk(frame_off, "frame_off", torch.int64, 1)
    B = rect4.shape[0]
    if pose7.shape != (F, 7) or intr4.shape != (F, 4) or frame_off.shape[0] != F + 1 or rect4.shape[1] != 4:
        raise ValueError("shape mismatch: pose7 [F,7], intr4 [F,4], rect4 [B,4], frame_off [F+1]")
    for t in (pose7, intr4, rect4, frame_off):
        if t.device != depth.device:
            raise ValueError("all tensors must live on the same device")
    if plan is None:
        plan = LiftPlan(F, B, depth.device, order_stats, H, W)
    elif plan.F < F or plan.B < B or plan.device != depth.device:
        raise ValueError("LiftPlan too small for this call")
    os_ptr = plan.order_stats.data_ptr() if plan.order_stats is not None else None
    with torch.cuda.device(depth.device):
        if gather is None:
            st = lib.lm3d_lift_boxes(
                depth.data_ptr(), F, H, W, pose7.data_ptr(), intr4.data_ptr(), rect4.data_ptr(), frame_off.data_ptr(), B,
                float(scale_depth), float(max_depth_mm), float(q), plan.records.data_ptr(), os_ptr,
                plan.workspace.data_ptr(), plan.workspace.numel(), _stream_ptr(depth.device),
            )
        else:
            if B > gather.n_records:
                raise ValueError("more boxes than the gather buffers hold per rank")
            st = lib.lm3d_lift_boxes_gather(
                depth.data_ptr(), F, H, W, pose7.data_ptr(), intr4.data_ptr(), rect4.data_ptr(), frame_off.data_ptr(), B,
                float(scale_depth), float(max_depth_mm), float(q), plan.records.data_ptr(), os_ptr,
                plan.workspace.data_ptr(), plan.workspace.numel(), gather.tables[gather_slot], gather.world,
                gather.box_offset, _stream_ptr(depth.device),
            )
    _capi.check(st, "lm3d_lift_boxes")
    rec = plan.records[:B]
    if plan.order_stats is not None:
        return rec, plan.order_stats[:B]
    return rec


def records_to_numpy(records: torch.Tensor) -> np.ndarray:
    """``[B,24]`` float32 tensor -> structured ``RECORD_DTYPE[B]`` (synchronising D2H copy)."""
    host = records.detach().contiguous().cpu().numpy()
    return host.view(RECORD_DTYPE).reshape(-1)


def lift_frame_cloud(depth, pose7, intr4, scale_depth: float = 1000.0, max_depth_mm: float = math.inf):
    """Full-frame world point cloud ``[F,H,W,3]`` (NaN where invalid) and per-frame valid counts
    (``lm3d_lift_frame_cloud``; reference ``pose_processor.py:154-156,262-271``)."""
    lib = _capi.load()
    _chk(depth, "depth", torch.float32, 3)
    F, H, W = depth.shape
    _chk(pose7, "pose7", torch.float64, 2)
    _chk(intr4, "intr4", torch.float64, 2)
    if pose7.shape != (F, 7) or intr4.shape != (F, 4):
        raise ValueError("shape mismatch: pose7 [F,7], intr4 [F,4]")
    if pose7.device != depth.device or intr4.device != depth.device:
        raise ValueError("all tensors must live on the same device")
    xyz = torch.empty((F, H, W, 3), dtype=torch.float32, device=depth.device)
    n_valid = torch.empty((F,), dtype=torch.int32, device=depth.device)
    ws = torch.empty((max(int(lib.lm3d_cloud_workspace_bytes(F)), 16),), dtype=torch.uint8, device=depth.device)
    with torch.cuda.device(depth.device):
        st = lib.lm3d_lift_frame_cloud(
            depth.data_ptr(), F, H, W, pose7.data_ptr(), intr4.data_ptr(), float(scale_depth), float(max_depth_mm),
            xyz.data_ptr(), n_valid.data_ptr(), ws.data_ptr(), ws.numel(), _stream_ptr(depth.device),
        )
    _capi.check(st, "lm3d_lift_frame_cloud")
    return xyz, n_valid


def lift_boxes_host(
    depth: np.ndarray,
    pose7: np.ndarray,
    intr4: np.ndarray,
    boxes_xyxy: np.ndarray,
    image_wh: np.ndarray,
    frame_off: np.ndarray,
    scale_depth: float = 1000.0,
    max_depth_mm: float = math.inf,
    q: float = 50.0,
    device: int = 0,
    out: np.ndarray | None = None,
) -> np.ndarray:
    """HOST-buffer entry (``lm3d_lift_boxes_host``): numpy in, structured records out; the
    library does the chunked H2D / lift / D2H itself.  This is the call the drop-in
    ``ProcessPose`` makes and the one ``bench.py`` times as ``e2e``."""
    lib = _capi.load()

    def arr(a, dt, name):
        a = np.ascontiguousarray(a, dtype=dt)
        return a

    depth = arr(depth, np.float32, "depth")
    if depth.ndim != 3:
        raise ValueError("depth must be [F,H,W]")
    F, H, W = depth.shape
    pose7 = arr(pose7, np.float64, "pose7")
    intr4 = arr(intr4, np.float64, "intr4")
    boxes_xyxy = arr(boxes_xyxy, np.float64, "boxes_xyxy").reshape(-1, 4)
    image_wh = arr(image_wh, np.float64, "image_wh")
    frame_off = arr(frame_off, np.int64, "frame_off")
    B = boxes_xyxy.shape[0]
    if pose7.shape != (F, 7) or intr4.shape != (F, 4) or image_wh.shape != (F, 2) or frame_off.shape != (F + 1,):
        raise ValueError("shape mismatch: pose7 [F,7], intr4 [F,4], image_wh [F,2], frame_off [F+1]")
    if out is None:
        out = np.empty(B, dtype=RECORD_DTYPE)
    elif out.dtype != RECORD_DTYPE or out.shape != (B,) or not out.flags.c_contiguous:
        raise ValueError("out must be a contiguous RECORD_DTYPE[B] array")
    st = lib.lm3d_lift_boxes_host(
        depth.ctypes.data, F, H, W, pose7.ctypes.data, intr4.ctypes.data, boxes_xyxy.ctypes.data,
        image_wh.ctypes.data, frame_off.ctypes.data, B, float(scale_depth), float(max_depth_mm), float(q),
        out.ctypes.data, int(device),
    )
    _capi.check(st, "lm3d_lift_boxes_host")
    return out
