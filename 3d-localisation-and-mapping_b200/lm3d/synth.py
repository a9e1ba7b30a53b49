"""Seeded synthetic RGB-D scan sequences (SURVEY.md section 8d generator).

The reference's input data is git-ignored (``/root/reference/.gitignore:1-11``), so every
test and benchmark runs on sequences made here.  Layouts follow the reference:

* depth ``[F,H,W]`` fp32 **millimetres**, H rows x W cols (``src/detector/dataset.py:68-81``),
* pose rows ``[tx,ty,tz,qx,qy,qz,qw]`` (``src/mapper/database_query.py:22-24``),
* intrinsics at RGB resolution ``{image_width,image_height,fx,fy,cx,cy}``
  (``src/detector/dataset.py:114-121``),
* boxes ``[x1,y1,x2,y2,damage_cls,conf,label]`` in RGB pixels
  (``src/detector/detector.py:126-127,155``).

``make_sequence`` is the numpy generator (bit-reproducible, used by parity tests);
``make_sequence_torch`` builds the same *law* directly in HBM for the full-size bench
configs (a 10k-frame numpy sequence would take minutes and a 2 GB H2D copy).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

# iPhone-like RGB intrinsics [DEFINED in SURVEY 8d]
RGB_W, RGB_H = 1440, 1920
RGB_FX = RGB_FY = 1450.0
RGB_CX, RGB_CY = 720.0, 960.0

#: BASELINE.json configs -> (F, H, W, boxes/frame).  C4 is sharded across ranks.
CONFIGS = {
    "C1": (100, 256, 192, 10),
    "C2": (10_000, 256, 192, 20),
    "C3": (10_000, 1920, 1440, 50),
    "C4": (1_000_000, 256, 192, 20),
    "C5": (1_000, 1920, 1440, 256),
}


@dataclass
class Sequence:
    depth: np.ndarray  # [F,H,W] f32 mm
    pose7: np.ndarray  # [F,7] f64
    intrinsics: list  # F dicts at RGB resolution
    boxes: np.ndarray  # [F,B,4] f64 RGB px (x1,y1,x2,y2)
    damage_cls: np.ndarray  # [F,B] int
    conf: np.ndarray  # [F,B] f64
    label: np.ndarray  # [F,B] int
    depth_width: int = 192
    depth_height: int = 256
    meta: dict = field(default_factory=dict)

    # ---- views in the reference's own input types -------------------------------------
    def bbox_coordinates(self) -> dict:
        """``{frame_index: [[x1,y1,x2,y2,damage_cls,conf,label], ...]}`` as produced by
        ``ObjectDetector.forward`` (``src/detector/detector.py:126-129``)."""
        out = {}
        F, B = self.boxes.shape[:2]
        for f in range(F):
            out[f] = [
                [
                    float(self.boxes[f, b, 0]),
                    float(self.boxes[f, b, 1]),
                    float(self.boxes[f, b, 2]),
                    float(self.boxes[f, b, 3]),
                    int(self.damage_cls[f, b]),
                    float(self.conf[f, b]),
                    int(self.label[f, b]),
                ]
                for b in range(B)
            ]
        return out

    def pose_dataframe(self):
        """DataFrame with the columns of ``PoseDataExtractor.fetch_data``
        (``src/mapper/database_query.py:20-25``)."""
        import pandas as pd

        F = self.pose7.shape[0]
        df = pd.DataFrame(self.pose7, columns=["tx", "ty", "tz", "qx", "qy", "qz", "qw"])
        df.insert(0, "timestamp", pd.to_datetime(np.arange(F, dtype=np.float64) / 30.0, unit="s"))
        return df

    def dataset(self):
        """Object with ``ds[i] -> (rgb, depth[H,W], intrinsics)`` like ``ImageDataset``
        (``src/detector/dataset.py:53-66``); rgb is a zero-size placeholder."""
        return ArrayDataset(self.depth, self.intrinsics)

    # ---- flat, C-ABI shaped views ------------------------------------------------------
    def intr4_depth_res(self) -> np.ndarray:
        """``[F,4]`` fp64 ``fx,fy,cx,cy`` after the reference's rescale
        (``pose_processor.py:133-137``: all four divided by the width ratio)."""
        out = np.empty((len(self.intrinsics), 4), dtype=np.float64)
        for f, ci in enumerate(self.intrinsics):
            s = ci["image_width"] / self.depth_width
            out[f] = (ci["fx"] / s, ci["fy"] / s, ci["cx"] / s, ci["cy"] / s)
        return out

    def image_wh(self) -> np.ndarray:
        return np.array([[ci["image_width"], ci["image_height"]] for ci in self.intrinsics], dtype=np.float64)

    def frame_off(self) -> np.ndarray:
        F, B = self.boxes.shape[:2]
        return (np.arange(F + 1, dtype=np.int64) * B).astype(np.int64)


class ArrayDataset:
    """In-memory stand-in for ``ImageDataset`` (``src/detector/dataset.py:12-66``)."""

    def __init__(self, depth, intrinsics):
        self.depth = depth
        self.intrinsics = intrinsics

    def __len__(self):
        return self.depth.shape[0]

    def __getitem__(self, idx):
        return None, self.depth[idx], self.intrinsics[idx]

    def batch(self, frames, depth_out):
        """Batched form used by the drop-in ``ProcessPose``: returns ``(cal, depth)`` -- ``cal [n,6]`` = fx, fy, cx, cy,
        image_width, image_height (RGB resolution) and the depth of ``frames`` as a C-contiguous ``[n,H,W]`` float32
        array: a zero-copy view of the store when the frames are consecutive, else ``depth_out`` filled."""
        idx = np.asarray(frames, dtype=np.int64)
        if len(idx) and int(idx[-1]) - int(idx[0]) + 1 == len(idx) and (len(idx) == 1 or bool(np.all(np.diff(idx) == 1))):
            depth = self.depth[int(idx[0]) : int(idx[-1]) + 1]
            if depth.dtype != np.float32 or not depth.flags.c_contiguous:
                depth_out[...] = depth
                depth = depth_out
        else:
            np.take(self.depth, idx, axis=0, out=depth_out)
            depth = depth_out
        cal = np.empty((len(idx), 6), dtype=np.float64)
        last, row = None, None
        for i, f in enumerate(idx):
            ci = self.intrinsics[int(f)]
            if ci is not last:
                last, row = ci, (ci["fx"], ci["fy"], ci["cx"], ci["cy"], ci["image_width"], ci["image_height"])
            cal[i] = row
        return cal, depth


def _quat_mul(a, b):
    ax, ay, az, aw = a
    bx, by, bz, bw = b
    return np.array(
        [
            aw * bx + ax * bw + ay * bz - az * by,
            aw * by - ax * bz + ay * bw + az * bx,
            aw * bz + ax * by - ay * bx + az * bw,
            aw * bw - ax * bx - ay * by - az * bz,
        ]
    )


def make_poses(F: int, rng) -> np.ndarray:
    """Smooth trajectory: ``t`` random walk sigma=2 cm/frame, ``q`` = random unit quaternion
    composed with a small per-frame rotation."""
    t = np.cumsum(rng.normal(0.0, 0.02, size=(F, 3)), axis=0)
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    out = np.empty((F, 7), dtype=np.float64)
    small = rng.normal(0.0, 0.01, size=(F, 3))
    for f in range(F):
        dq = np.array([*(0.5 * small[f]), 1.0])
        dq /= np.linalg.norm(dq)
        q = _quat_mul(q, dq)
        q /= np.linalg.norm(q)
        out[f, :3] = t[f]
        out[f, 3:] = q
    return out


def make_boxes(F: int, B: int, rng) -> np.ndarray:
    """RGB-pixel boxes: ``w/W, h/H ~ U(0.10,0.40)``, top-left uniform with the box inside."""
    wf = rng.uniform(0.10, 0.40, size=(F, B))
    hf = rng.uniform(0.10, 0.40, size=(F, B))
    x1 = rng.uniform(0.0, 1.0, size=(F, B)) * (1.0 - wf)
    y1 = rng.uniform(0.0, 1.0, size=(F, B)) * (1.0 - hf)
    return np.stack([x1 * RGB_W, y1 * RGB_H, (x1 + wf) * RGB_W, (y1 + hf) * RGB_H], axis=-1)


def make_sequence(F: int, H: int, W: int, B: int, seed: int = 1234, patches: bool = True) -> Sequence:
    """Numpy generator (``numpy.random.default_rng(seed)``), SURVEY 8d law:
    tilted plane ``1500 + 0.8u + 0.5v`` mm + N(0,15); per box a flat "sign" patch 80-200 mm
    nearer over the inner 60 % of the box (+N(0,2)); 2 % pixels 0 and 0.1 % NaN."""
    rng = np.random.default_rng(seed)
    pose7 = make_poses(F, rng)
    boxes = make_boxes(F, B, rng)
    u = np.arange(W, dtype=np.float32)[None, :]
    v = np.arange(H, dtype=np.float32)[:, None]
    plane = (1500.0 + 0.8 * u + 0.5 * v).astype(np.float32)
    depth = np.empty((F, H, W), dtype=np.float32)
    sx, sy = W / RGB_W, H / RGB_H
    for f in range(F):
        d = plane + rng.normal(0.0, 15.0, size=(H, W)).astype(np.float32)
        if patches:
            offs = rng.uniform(80.0, 200.0, size=B)
            for b in range(B):
                x1, y1, x2, y2 = boxes[f, b]
                bw, bh = (x2 - x1) * sx, (y2 - y1) * sy
                px0 = int(x1 * sx + 0.2 * bw)
                px1 = int(x1 * sx + 0.8 * bw)
                py0 = int(y1 * sy + 0.2 * bh)
                py1 = int(y1 * sy + 0.8 * bh)
                if px1 <= px0 or py1 <= py0:
                    continue
                centre = plane[min((py0 + py1) // 2, H - 1), min((px0 + px1) // 2, W - 1)]
                d[py0:py1, px0:px1] = (
                    centre - offs[b] + rng.normal(0.0, 2.0, size=(py1 - py0, px1 - px0))
                ).astype(np.float32)
        r = rng.random(size=(H, W))
        d[r < 0.02] = 0.0
        d[r > 0.999] = np.nan
        depth[f] = d
    intr = [
        dict(image_width=RGB_W, image_height=RGB_H, fx=RGB_FX, fy=RGB_FY, cx=RGB_CX, cy=RGB_CY)
        for _ in range(F)
    ]
    return Sequence(
        depth=depth,
        pose7=pose7,
        intrinsics=intr,
        boxes=boxes,
        damage_cls=rng.integers(0, 2, size=(F, B)),
        conf=rng.uniform(0.25, 1.0, size=(F, B)),
        label=rng.integers(0, 8, size=(F, B)),
        depth_width=W,
        depth_height=H,
        meta=dict(seed=seed, F=F, H=H, W=W, B=B),
    )


def make_config(name: str, frames: int | None = None) -> Sequence:
    """Sequence of a BASELINE.json config shape (seed = 1234 + config number), optionally
    with fewer frames (parity tests scale the frame count down, never the frame shape)."""
    F, H, W, B = CONFIGS[name]
    return make_sequence(frames or F, H, W, B, seed=1234 + int(name[1:]))


# ---------------------------------------------------------------------------------------
# Device-side generator for the full-size bench (same law, torch RNG, generated in HBM)
# ---------------------------------------------------------------------------------------
def make_sequence_torch(F: int, H: int, W: int, B: int, seed: int, device, chunk: int = 512):
    """Returns a dict of device tensors shaped for ``lm3d.lift.lift_boxes``:
    ``depth [F,H,W] f32``, ``pose7 [F,7] f64``, ``intr4 [F,4] f64`` (depth resolution),
    ``boxes [F*B,4] f64`` RGB px, ``image_wh [F,2] f64``, ``frame_off [F+1] i64``."""
    import torch

    g = torch.Generator(device=device)
    g.manual_seed(seed)
    rng = np.random.default_rng(seed)
    pose7 = torch.from_numpy(make_poses(F, rng)).to(device)
    boxes = torch.from_numpy(make_boxes(F, B, rng)).to(device)  # [F,B,4]
    s = RGB_W / W
    intr4 = torch.tensor([RGB_FX / s, RGB_FY / s, RGB_CX / s, RGB_CY / s], dtype=torch.float64, device=device)
    intr4 = intr4.expand(F, 4).contiguous()
    image_wh = torch.tensor([RGB_W, RGB_H], dtype=torch.float64, device=device).expand(F, 2).contiguous()
    depth = torch.empty((F, H, W), dtype=torch.float32, device=device)
    u = torch.arange(W, dtype=torch.float32, device=device)[None, None, :]
    v = torch.arange(H, dtype=torch.float32, device=device)[None, :, None]
    plane = 1500.0 + 0.8 * u + 0.5 * v
    sx, sy = W / RGB_W, H / RGB_H
    offs = 80.0 + 120.0 * torch.rand((F, B), generator=g, device=device)
    for c0 in range(0, F, chunk):
        c1 = min(F, c0 + chunk)
        n = c1 - c0
        d = plane + 15.0 * torch.randn((n, H, W), generator=g, device=device)
        noise2 = 2.0 * torch.randn((n, H, W), generator=g, device=device)
        bx = boxes[c0:c1]
        bw = (bx[..., 2] - bx[..., 0]) * sx
        bh = (bx[..., 3] - bx[..., 1]) * sy
        px0 = (bx[..., 0] * sx + 0.2 * bw).floor()
        px1 = (bx[..., 0] * sx + 0.8 * bw).floor()
        py0 = (bx[..., 1] * sy + 0.2 * bh).floor()
        py1 = (bx[..., 1] * sy + 0.8 * bh).floor()
        for b in range(B):
            mx = (u >= px0[:, b, None, None]) & (u < px1[:, b, None, None])
            my = (v >= py0[:, b, None, None]) & (v < py1[:, b, None, None])
            ucen = ((px0[:, b] + px1[:, b]) * 0.5).floor().float()
            vcen = ((py0[:, b] + py1[:, b]) * 0.5).floor().float()
            level = (1500.0 + 0.8 * ucen + 0.5 * vcen - offs[c0:c1, b].float())[:, None, None]
            d = torch.where(mx & my, level + noise2, d)
        r = torch.rand((n, H, W), generator=g, device=device)
        d = torch.where(r < 0.02, torch.zeros((), device=device), d)
        d = torch.where(r > 0.999, torch.full((), float("nan"), device=device), d)
        depth[c0:c1] = d
    frame_off = torch.arange(F + 1, dtype=torch.int64, device=device) * B
    return dict(
        depth=depth,
        pose7=pose7,
        intr4=intr4,
        boxes=boxes.reshape(F * B, 4).contiguous(),
        image_wh=image_wh,
        frame_off=frame_off,
    )
