"""Host-side geometry helpers of the mapper (``Transforms``).

The reference imports ``src.utils.transformations.Transforms`` but does not ship it
(``/root/reference/src/mapper/pose_processor.py:17``); the methods below are restated from
their call sites so the reference's entry points keep importing.  The per-box arithmetic
itself runs in CUDA (``lm3d.lift``): these scalar helpers exist for API compatibility
(``display_3d`` overlays, consumers such as ``Mapping``) and are NOT on the lift's hot path.
"""
from __future__ import annotations

import numpy as np


class Transforms:
    def get_translation(self, pose_data):
        """``[tx,ty,tz]`` of a pose row (call site ``pose_processor.py:228``)."""
        return np.asarray(pose_data[:3], dtype=np.float64)

    def get_rotation(self, pose_data):
        """3x3 rotation of a pose row, quaternion scalar-last (call site ``pose_processor.py:229``)."""
        x, y, z, w = (float(v) for v in pose_data[3:7])
        n = np.sqrt(x * x + y * y + z * z + w * w)
        x, y, z, w = x / n, y / n, z / n, w / n
        return np.array(
            [
                [1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)],
            ],
            dtype=np.float64,
        )

    def get_transformation_matrix(self, pose_data):
        """4x4 camera->world ``[[R,t],[0,1]]`` (call sites ``pose_processor.py:140,254``)."""
        T = np.eye(4, dtype=np.float64)
        T[:3, :3] = self.get_rotation(pose_data)
        T[:3, 3] = self.get_translation(pose_data)
        return T

    def scale_bounding_box(self, bbox, image_size, depth_size):
        """RGB-pixel box -> depth-pixel box, tail passed through (call site ``:174-178``).  The depth size is
        remembered: ``bbox_to_3d`` needs it to clamp the rect the percentile depth is taken over."""
        iw, ih = image_size
        dw, dh = depth_size
        self._depth_size = (int(dw), int(dh))
        out = list(bbox)
        out[0] = float(bbox[0]) * dw / iw
        out[1] = float(bbox[1]) * dh / ih
        out[2] = float(bbox[2]) * dw / iw
        out[3] = float(bbox[3]) * dh / ih
        return out

    def bbox_to_3d(self, scaled_bbox, img_size=None):
        """Four ``(x, y)`` corners TL, BL, BR, TR (call site ``:181``; order of the in-repo precedent
        ``src/detector/detector.py:202``).  The corners stay float -- the reference truncates them with ``int()``
        itself (``:186-187``).  Also fixes the box's inclusive pixel rect (``int()`` truncation, clamped to the
        frame) for the ``_depth_to_3d`` calls that follow: the reference comments those calls "z-values from median
        over bbox (x, y) range" (``:183``) but passes only a corner and the depth image, so the range has to
        travel from here (ORACLE-SPEC v0, R6 / R9)."""
        x1, y1, x2, y2 = (float(v) for v in scaled_bbox[:4])
        size = getattr(self, "_depth_size", None)
        if size is not None:
            dw, dh = size

            def px(v, hi):
                return min(max(int(v), 0), hi)

            xa, xb, ya, yb = px(x1, dw - 1), px(x2, dw - 1), px(y1, dh - 1), px(y2, dh - 1)
            self._rect = (min(xa, xb), min(ya, yb), max(xa, xb), max(ya, yb))
        else:
            self._rect = None
        self._rect_depth = None
        return [(x1, y1), (x1, y2), (x2, y2), (x2, y1)]

    percentile = 50.0  # the reference's "median"; ProcessPose(percentile=...) of the CUDA path mirrors it

    def _depth_to_3d(self, x, y, depth, fx, fy, cx, cy, scale_depth):
        """Pixel + depth -> camera-frame ``[X,Y,Z]`` (call site ``:184-196``).

        ``depth`` is the ``[H,W]`` depth image (what the reference passes) or a scalar depth in depth units.
        With an image, z is the percentile (median) of the VALID depths (finite, > 0) inside the rect the
        preceding ``bbox_to_3d`` fixed -- the same rule the CUDA lift and the oracle implement (R8-R10) -- and
        the pixel is clamped to the frame (a box touching the right edge gives ``int(x) == W``).  Without a
        preceding ``bbox_to_3d`` there is no rect and the pixel's own depth is used."""
        if np.ndim(depth) == 2:
            H, W = depth.shape
            x = min(max(int(x), 0), W - 1)
            y = min(max(int(y), 0), H - 1)
            rect = getattr(self, "_rect", None)
            if rect is None:
                d = float(depth[y, x])
            else:
                if getattr(self, "_rect_depth", None) is None:
                    x0, y0, x1, y1 = rect
                    patch = np.asarray(depth[y0 : y1 + 1, x0 : x1 + 1], dtype=np.float32)
                    with np.errstate(invalid="ignore"):
                        ok = np.isfinite(patch) & (patch > 0)
                    vals = patch[ok].astype(np.float64)
                    self._rect_depth = (
                        float(np.percentile(vals, self.percentile, method="linear")) if vals.size else float("nan")
                    )
                d = self._rect_depth
        else:
            d = float(depth)
        z = d / scale_depth
        return np.array([(x - cx) * z / fx, (y - cy) * z / fy, z], dtype=np.float64)

    def create_3d_bounding_box(self, corners, buffer):
        """Extrude 4 coplanar world corners to 8 along the face normal by ``+-buffer``
        (call sites ``pose_processor.py:204-206``, ``mapping.py:163-165``).  Display only."""
        c = [np.asarray(p, dtype=np.float64) for p in corners[:4]]
        n = np.cross(c[1] - c[0], c[3] - c[0])
        ln = np.linalg.norm(n)
        n = n / ln if ln > 0 else np.array([0.0, 0.0, 1.0])
        return [p - buffer * n for p in c] + [p + buffer * n for p in c]

    def get_camera_direction(self, pose_df):
        """Optical-axis (+z) direction of every pose row (call sites ``mapping.py:187``,
        ``mapper/database_query.py:37``)."""
        cols = ["tx", "ty", "tz", "qx", "qy", "qz", "qw"]
        rows = pose_df[cols].to_numpy(dtype=np.float64) if hasattr(pose_df, "columns") else np.asarray(pose_df)
        return np.stack([self.get_rotation(r)[:, 2] for r in rows], axis=0)
