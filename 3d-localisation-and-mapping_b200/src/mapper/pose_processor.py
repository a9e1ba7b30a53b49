"""Drop-in ``ProcessPose``: same constructor, same ``get_global_coordinates()`` result as
``/root/reference/src/mapper/pose_processor.py:38-122``, with the per-frame / per-box Python
loops replaced by ONE batched call into the sm_100a CUDA library (``lm3d``).

Output format is unchanged (``pose_processor.py:115,208``):
``{frame_index: [[c0, c1, c2, c3, damage_cls, conf, label], ...]}`` with ``c_i`` a
``np.ndarray`` of shape ``(3,)`` in world coordinates, dict order = ``bbox_coordinates`` order,
frames without boxes map to ``[]``, everything picklable.  The richer per-box record
(centroid, extents, counts) is kept in ``self.records`` after the call.

There is no CPU fallback: without a CUDA device / ``liblm3d.so`` the call raises.
"""
from __future__ import annotations

import logging

import numpy as np

from lm3d import lift
from src.utils.transformations import Transforms
from src.utils.visualisation import Visualiser


class ProcessPose:
    def __init__(
        self,
        pose,
        dataset,
        bbox_coordinates,
        img_size,
        depth_width,
        depth_height,
        display_rgbd=False,
        display_3d=False,
        scale_depth=1000,
        bbox_depth_buffer=0.03,
        verbose=False,
        device=0,
        percentile=50.0,
        max_depth_mm=float("inf"),
    ):
        """Parameters as the reference (``pose_processor.py:39-52``); ``device``,
        ``percentile`` and ``max_depth_mm`` are additions with reference-preserving defaults."""
        self.pose = pose
        self.dataset = dataset
        self.bbox_coordinates = bbox_coordinates
        self.img_size = img_size
        self.depth_width = depth_width
        self.depth_height = depth_height
        self.display_rgbd = display_rgbd
        self.display_3d = display_3d
        self.scale_depth = scale_depth
        self.bbox_depth_buffer = bbox_depth_buffer
        self.verbose = verbose
        self.device = device
        self.percentile = percentile
        self.max_depth_mm = max_depth_mm
        self.records = None
        self.record_frames = None

        self.visualiser = Visualiser()
        self.transforms = Transforms()

        logging.basicConfig(level=logging.INFO)
        self.logger = logging.getLogger(__name__)
        self.logger.info("Processing Pose.")

    # ------------------------------------------------------------------------------------
    def _gather(self):
        """Batch the sequence: one pass over ``bbox_coordinates`` (reference loop header,
        ``pose_processor.py:89-93``) collecting depth, pose row, rescaled intrinsics, boxes."""
        frames = list(self.bbox_coordinates.keys())
        F = len(frames)
        H, W = int(self.depth_height), int(self.depth_width)
        depth = np.empty((F, H, W), dtype=np.float32)
        pose7 = np.empty((F, 7), dtype=np.float64)
        intr4 = np.empty((F, 4), dtype=np.float64)
        image_wh = np.empty((F, 2), dtype=np.float64)
        frame_off = np.zeros(F + 1, dtype=np.int64)
        boxes = []
        # pose row i = frame i, first column is the timestamp (pose_processor.py:109): one conversion for the whole
        # DataFrame instead of one pandas row lookup per frame (50 us each: half a second per 10 k frames)
        pose_np = (self.pose.iloc[:, 1:8].to_numpy(dtype=np.float64) if hasattr(self.pose, "iloc") else None)
        for i, frame_index in enumerate(frames):
            rgb_tensor, depth_tensor, ci = self.dataset[frame_index]
            _, depth_image = self.visualiser.parse_images(None, depth_tensor)
            if depth_image.shape != (H, W):
                raise ValueError(f"frame {frame_index}: depth is {depth_image.shape}, expected {(H, W)}")
            depth[i] = depth_image
            pose7[i] = pose_np[frame_index] if pose_np is not None else np.asarray(self.pose[frame_index], dtype=np.float64)
            # intrinsics rescale: every entry, cy included, by the WIDTH ratio (:133-137)
            s = ci["image_width"] / self.depth_width
            intr4[i] = (ci["fx"] / s, ci["fy"] / s, ci["cx"] / s, ci["cy"] / s)
            image_wh[i] = (ci["image_width"], ci["image_height"])
            bxs = self.bbox_coordinates[frame_index]
            frame_off[i + 1] = frame_off[i] + len(bxs)
            boxes.extend([float(b[0]), float(b[1]), float(b[2]), float(b[3])] for b in bxs)
        boxes = np.asarray(boxes, dtype=np.float64).reshape(-1, 4)
        return frames, depth, pose7, intr4, image_wh, frame_off, boxes

    def get_global_coordinates(self):
        if self.display_rgbd or self.display_3d:
            raise NotImplementedError("display_rgbd / display_3d need the reference's Open3D GUI (out of scope)")
        frames, depth, pose7, intr4, image_wh, frame_off, boxes = self._gather()
        rec = lift.lift_boxes_host(
            depth, pose7, intr4, boxes, image_wh, frame_off,
            scale_depth=float(self.scale_depth), max_depth_mm=float(self.max_depth_mm),
            q=float(self.percentile), device=int(self.device),
        )
        self.records = rec
        self.record_frames = frames
        corners = rec["corners"].astype(np.float64)  # [B,4,3]
        global_bboxes = {}
        for i, frame_index in enumerate(frames):
            rows = []
            b0 = int(frame_off[i])
            for j, bbox in enumerate(self.bbox_coordinates[frame_index]):
                c = corners[b0 + j]
                # row = global_corners + bbox[-3:]  (pose_processor.py:208)
                rows.append([c[0].copy(), c[1].copy(), c[2].copy(), c[3].copy()] + list(bbox[-3:]))
                if self.verbose:
                    self.logger.info(f"\tOriginal 2D Corners: {bbox}")
                    self.logger.info(f"\tGlobal 3D Coordinates: {rows[-1][:4]}\n")
            global_bboxes[frame_index] = rows
        return global_bboxes

    def _transform_to_global(self, local_point, pose_data):
        """Kept for callers of the reference's helper (``pose_processor.py:242-260``)."""
        transformation = self.transforms.get_transformation_matrix(pose_data)
        local_point = np.array([(*local_point, 1)])
        return (transformation @ local_point.T)[:3, 0]
