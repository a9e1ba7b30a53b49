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


# ---------------------------------------------------------------------------------------
# Batched loader: depth PNGs + calibration YAMLs of a scan -> [F,H,W] on the device
# ---------------------------------------------------------------------------------------
def _natural_key(name: str):
    """Order ``2.png`` before ``10.png`` (the reference orders file names with ``natsorted``, ``dataset.py:32-33``)."""
    import re

    return [int(t) if t.isdigit() else t.lower() for t in re.split(r"(\d+)", name)]


def load_calibration(path) -> tuple:
    """One calibration YAML -> ``(fx, fy, cx, cy, image_width, image_height)`` at RGB resolution, the entries
    ``ImageDataset._load_calibration`` reads (``dataset.py:102-121``: ``camera_matrix.data[0,4,2,5]``)."""
    import yaml

    loader = getattr(yaml, "CSafeLoader", yaml.SafeLoader)
    with open(path, "r") as fh:
        c = yaml.load(fh, Loader=loader)
    m = c["camera_matrix"]["data"]
    return (float(m[0]), float(m[4]), float(m[2]), float(m[5]), float(c.get("image_width")), float(c.get("image_height")))


class DepthSequence:
    """The depth + calibration side of ``ImageDataset`` (``/root/reference/src/detector/dataset.py:12-121``) as a
    BATCHED loader for the lift: where the reference decodes, reinterprets and scales one frame per
    ``__getitem__`` on the CPU, ``batch_device(frames)`` decodes the PNGs of a whole chunk on a thread pool
    (``cv2.imread`` releases the GIL) straight into a pinned staging ring, copies each slot to the device while the
    next one is being decoded, and converts the raw 8UC4 bytes to fp32 millimetres IN PLACE on the device
    (``lm3d_ingest_depth``).  The result is the ``[n,H,W]`` tensor + ``[n,6]`` calibration table the drop-in
    ``ProcessPose`` lifts from without a host-side depth array.  No CPU conversion path exists."""

    def __init__(self, depth_paths, calib_paths, depth_width: int = 192, depth_height: int = 256, device=0,
                 workers: int | None = None, slot_frames: int = 256):
        import os

        if len(depth_paths) != len(calib_paths):
            raise ValueError("one calibration file per depth image")
        self.depth_paths = [str(p) for p in depth_paths]
        self.calib_paths = [str(p) for p in calib_paths]
        self.depth_width, self.depth_height = int(depth_width), int(depth_height)
        self.device = torch.device("cuda", device) if isinstance(device, int) else torch.device(device)
        self.workers = int(workers or min(32, max(1, len(os.sched_getaffinity(0)))))
        self.slot_frames = int(slot_frames)
        self._cal = {}
        self._ring = None
        self.decode_seconds = 0.0  # wall time spent waiting for PNG decode (reported by bench.py)

    @classmethod
    def from_dirs(cls, depth_image_dir, calibration_dir, image_dir=None, **kw):
        """Pair files like ``ImageDataset._pair_filenames`` (``dataset.py:39-49``): natural order of the RGB file
        names, a frame exists when ``N.jpg`` has its ``N.png``; without an RGB directory every depth PNG counts."""
        import os

        depth_names = set(os.listdir(depth_image_dir))
        if image_dir is not None:
            stems = [n[: -len(".jpg")] for n in sorted(os.listdir(image_dir), key=_natural_key) if n.endswith(".jpg")]
            stems = [s for s in stems if s + ".png" in depth_names]
        else:
            stems = [n[: -len(".png")] for n in sorted(depth_names, key=_natural_key) if n.endswith(".png")]
        return cls([os.path.join(depth_image_dir, s + ".png") for s in stems],
                   [os.path.join(calibration_dir, s + ".yaml") for s in stems], **kw)

    def __len__(self):
        return len(self.depth_paths)

    def calibration(self, frames) -> np.ndarray:
        """``[n,6]`` fp64: fx, fy, cx, cy, image_width, image_height of ``frames`` (parsed once, cached)."""
        out = np.empty((len(frames), 6), dtype=np.float64)
        for i, f in enumerate(frames):
            f = int(f)
            if f not in self._cal:
                self._cal[f] = load_calibration(self.calib_paths[f])
            out[i] = self._cal[f]
        return out

    def _decode_into(self, slot: np.ndarray, i: int, path: str):
        import cv2

        img = cv2.imread(path, cv2.IMREAD_UNCHANGED)  # [H,W,4] uint8 = the bytes of fp32 metres (dataset.py:70-74)
        if img is None:
            raise FileNotFoundError(path)
        if img.dtype != np.uint8 or img.size != self.depth_height * self.depth_width * 4:
            raise ValueError(f"{path}: expected an 8UC4 PNG of {self.depth_height}x{self.depth_width}, got {img.shape} {img.dtype}")
        slot[i] = img.reshape(self.depth_height, self.depth_width, 4)

    def batch_device(self, frames, out: torch.Tensor | None = None):
        """Depth of ``frames`` as a CUDA ``[n,H,W]`` float32 tensor in millimetres + their ``[n,6]`` calibration rows."""
        import time
        from concurrent.futures import ThreadPoolExecutor

        _capi.load()  # fail before any file is touched when the CUDA library is missing
        n, H, W = len(frames), self.depth_height, self.depth_width
        if out is None:
            out = torch.empty((n, H, W), dtype=torch.float32, device=self.device)
        if out.shape != (n, H, W) or out.dtype != torch.float32 or not out.is_cuda or not out.is_contiguous():
            raise ValueError("out must be a contiguous CUDA float32 [n,H,W] tensor")
        S = max(1, min(self.slot_frames, n))
        if self._ring is None or self._ring[0].shape[0] < S:
            self._ring = [torch.empty((S, H, W, 4), dtype=torch.uint8, pin_memory=True) for _ in range(2)]
            self._done = [torch.cuda.Event(), torch.cuda.Event()]
        raw_dev = out.view(torch.uint8).view(n, H, W, 4)  # the fp32 output IS the raw byte buffer: converted in place
        stream = torch.cuda.current_stream(self.device)
        with ThreadPoolExecutor(self.workers) as pool:
            def submit(k):
                lo, hi = k * S, min(n, (k + 1) * S)
                slot = self._ring[k & 1].numpy()
                return [pool.submit(self._decode_into, slot, i - lo, self.depth_paths[int(frames[i])]) for i in range(lo, hi)]

            n_slots = (n + S - 1) // S
            pending = submit(0) if n_slots else []
            for k in range(n_slots):
                lo, hi = k * S, min(n, (k + 1) * S)
                t0 = time.perf_counter()
                for fut in pending:
                    fut.result()
                self.decode_seconds += time.perf_counter() - t0
                if k + 1 < n_slots:
                    self._done[(k + 1) & 1].synchronize()  # the copy that last read that slot has finished
                    pending = submit(k + 1)
                raw_dev[lo:hi].copy_(self._ring[k & 1][: hi - lo], non_blocking=True)
                self._done[k & 1].record(stream)
        decode_depth(raw_dev, 1000.0, out=out)  # reinterpret + x1000 on the device, in place (dataset.py:73-77)
        return out, self.calibration(frames)
