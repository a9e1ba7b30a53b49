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
        """RGB-pixel box -> depth-pixel box, tail passed through (call site ``:174-178``)."""
        iw, ih = image_size
        dw, dh = depth_size
        out = list(bbox)
        out[0] = float(bbox[0]) * dw / iw
        out[1] = float(bbox[1]) * dh / ih
        out[2] = float(bbox[2]) * dw / iw
        out[3] = float(bbox[3]) * dh / ih
        return out

    def bbox_to_3d(self, scaled_bbox, img_size=None):
        """Four ``(x, y)`` corners TL, BL, BR, TR (call site ``:181``; order of the in-repo
        precedent ``src/detector/detector.py:202``)."""
        x1, y1, x2, y2 = (float(v) for v in scaled_bbox[:4])
        return [(x1, y1), (x1, y2), (x2, y2), (x2, y1)]

    def _depth_to_3d(self, x, y, depth, fx, fy, cx, cy, scale_depth):
        """Pixel + depth -> camera-frame ``[X,Y,Z]`` (call site ``:184-196``).  ``depth`` may be
        a scalar depth in depth units, or an ``[H,W]`` image (then the pixel's own value)."""
        d = float(depth[int(y), int(x)]) if np.ndim(depth) == 2 else float(depth)
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
