"""Drop-in ``BoundingBoxProcessor``: the 3-D NMS the reference runs on the result of
``ProcessPose.get_global_coordinates()`` (``/root/reference/task_def.py:145-149``), as one batched call into the
sm_100a CUDA library.

The reference's own ``src/mapper/bbox_optimiser.py`` is not in its repository.  Constructor arguments and the
method name follow its only call site; the result keeps the shape its consumer reads
(``src/mapper/mapping.py:170-176``: ``for frame_index, bbox_list in optimised_bboxes.items()``, ``bbox[:4]`` = the
four world corners): every input frame key in input order, each with the KEPT rows of that frame, rows unchanged.
The suppression rules are NMS-SPEC v0 (``DESIGN.md`` 4.8).  There is no CPU fallback.
"""
from __future__ import annotations

import numpy as np
import torch

from lm3d import nms


class BoundingBoxProcessor:
    def __init__(self, global_bboxes_data, pose=None, iou_thresh=nms.DEFAULT_IOU_THR, bbox_depth_buffer=nms.DEFAULT_PAD_M,
                 device=0):
        """``global_bboxes_data``: ``{frame: [[c0, c1, c2, c3, damage_cls, conf, label], ...]}``; ``pose`` (the pose
        DataFrame the reference passes) is accepted and unused: the rows are already in world coordinates."""
        self.global_bboxes_data = global_bboxes_data
        self.pose = pose
        self.iou_thresh = float(iou_thresh)
        self.bbox_depth_buffer = float(bbox_depth_buffer)
        self.device = torch.device("cuda", device) if isinstance(device, int) else torch.device(device)
        self.keep = None      # uint8 [B] after suppress_bboxes(), rows in dict order
        self.parent = None    # int32 [B]: the kept row each suppressed row was merged into
        self.rounds = 0

    def suppress_bboxes(self):
        rows = [(f, r) for f, lst in self.global_bboxes_data.items() for r in lst]
        out = {f: [] for f in self.global_bboxes_data}
        if not rows:
            self.keep = np.zeros(0, dtype=np.uint8)
            self.parent = np.zeros(0, dtype=np.int32)
            return out
        corners = np.array([[np.asarray(c, dtype=np.float64) for c in r[:4]] for _, r in rows], dtype=np.float32)
        conf = np.array([float(r[5]) for _, r in rows], dtype=np.float32)
        ids: dict = {}
        label = np.array([ids.setdefault(r[6], len(ids)) for _, r in rows], dtype=np.int32)
        keep, parent, self.rounds = nms.nms_boxes(
            torch.from_numpy(corners.reshape(-1, 12)).to(self.device), torch.from_numpy(conf).to(self.device),
            torch.from_numpy(label).to(self.device), self.iou_thresh, self.bbox_depth_buffer)
        self.keep = keep.cpu().numpy()
        self.parent = parent.cpu().numpy()
        for (f, r), k in zip(rows, self.keep):
            if k:
                out[f].append(r)
        return out
