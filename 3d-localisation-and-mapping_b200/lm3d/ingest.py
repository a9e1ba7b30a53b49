"""Depth / calibration ingest on the GPU (SURVEY.md 8f "next" #2).

``ImageDataset._load_depth_image`` (``/root/reference/src/detector/dataset.py:68-81``) decodes one depth PNG per
``__getitem__`` -- 8UC4 pixels that are the bytes of fp32 metres -- reinterprets and scales it to millimetres on
the CPU.  Here the decoded bytes of a whole sequence are converted in one kernel (``lm3d_ingest_depth``), in place
if wanted, and the calibration dicts become the ``[F,4]`` table ``lm3d_lift_boxes`` takes.  PNG inflate itself
stays with cv2 / the caller.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _capi


def decode_depth(raw_8uc4: torch.Tensor, scale: float = 1000.0, out: torch.Tensor | None = None) -> torch.Tensor:
    """``[...,H,W,4]`` uint8 CUDA tensor (decoded depth PNGs) -> ``[...,H,W]`` float32 millimetres (``dataset.py:70-77``)."""
    lib = _capi.load()
    if not raw_8uc4.is_cuda:
        raise ValueError("raw_8uc4 must be a CUDA tensor: there is no CPU fallback")
    if raw_8uc4.dtype != torch.uint8 or raw_8uc4.shape[-1] != 4 or not raw_8uc4.is_contiguous():
        raise ValueError("raw_8uc4 must be a contiguous [...,H,W,4] uint8 tensor")
    shape = tuple(raw_8uc4.shape[:-1])
    n = int(np.prod(shape)) if shape else 1
    if out is None:
        out = torch.empty(shape, dtype=torch.float32, device=raw_8uc4.device)
    if out.dtype != torch.float32 or not out.is_contiguous() or out.numel() != n or out.device != raw_8uc4.device:
        raise ValueError("out must be a contiguous float32 tensor with one element per pixel on the same device")
    with torch.cuda.device(raw_8uc4.device):
        st = lib.lm3d_ingest_depth(raw_8uc4.data_ptr(), n, float(scale), out.data_ptr(),
                                   torch.cuda.current_stream().cuda_stream)
    _capi.check(st, "lm3d_ingest_depth")
    return out


def intrinsics_table(calibrations, depth_width: int) -> np.ndarray:
    """Calibration dicts (``dataset.py:102-121``) -> ``[F,4]`` fp64 ``fx fy cx cy`` at depth resolution
    (``pose_processor.py:133-137``: all four divided by the WIDTH ratio ``image_width / depth_width``)."""
    tab = np.empty((len(calibrations), 4), dtype=np.float64)
    for i, c in enumerate(calibrations):
        s = c["image_width"] / depth_width
        tab[i] = (c["fx"] / s, c["fy"] / s, c["cx"] / s, c["cy"] / s)
    return tab
