"""Minimal ``Visualiser`` so the mapper's entry points import without Open3D.

Only ``parse_images`` matters to the lift (depth layout: ``[H,W]`` fp32 millimetres, call site
``/root/reference/src/mapper/pose_processor.py:94-97``).  The interactive overlays the reference
draws with Open3D / cv2 windows are out of scope and raise if requested.
"""
from __future__ import annotations

import numpy as np


class Visualiser:
    def parse_images(self, rgb_tensor, depth_tensor):
        rgb = None
        if rgb_tensor is not None:
            rgb = np.asarray(rgb_tensor)
            if rgb.ndim == 3 and rgb.shape[0] in (1, 3):  # CHW tensor -> HWC image
                rgb = np.transpose(rgb, (1, 2, 0))
        depth = np.ascontiguousarray(np.asarray(depth_tensor), dtype=np.float32)
        return rgb, depth

    def _no_gui(self, *_a, **_k):
        raise NotImplementedError(
            "interactive Open3D / cv2 display is outside the B200 lift's scope (SURVEY.md section 2)"
        )

    display_imgs = gen_rgbd = gen_point_cloud = overlay_3d_bbox = _no_gui
    _overlay_camera_frustum = overlay_pose = overlay_pose_directions = _no_gui
