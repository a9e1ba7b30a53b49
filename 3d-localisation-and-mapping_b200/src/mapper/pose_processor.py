"""Drop-in ``ProcessPose``: same constructor, same ``get_global_coordinates()`` result as
``/root/reference/src/mapper/pose_processor.py:38-122``, with the per-frame / per-box Python
loops replaced by batched calls into the sm_100a CUDA library (``lm3d``).

Output format is unchanged (``pose_processor.py:115,208``):
``{frame_index: [[c0, c1, c2, c3, damage_cls, conf, label], ...]}`` with ``c_i`` a
``np.ndarray`` of shape ``(3,)`` in world coordinates, dict order = ``bbox_coordinates`` order,
frames without boxes map to ``[]``, everything picklable.

Beside it, the columnar form of the same result: ``get_global_records()``
returns ``LiftedRecords`` -- the ``lm3d_box_out`` records as one structured array plus the CSR
frame offsets, frame keys and the passthrough tail -- without building a Python object per
box, and ``LiftedRecords.save`` / ``load`` replace the stage pickle of nested lists
(``task_def.py:62-72``, ``pose_processor.py:316-320``) with one ``.npz``.

There is no CPU fallback: without a CUDA device / ``liblm3d.so`` the call raises.
"""
from __future__ import annotations

import logging
from dataclasses import dataclass
from itertools import chain
from operator import itemgetter

import numpy as np

from lm3d import lift
from src.utils.transformations import Transforms
from src.utils.visualisation import Visualiser

#: frames are gathered and lifted in chunks of about this many bytes of depth (the reference holds one frame at a
#: time, ``pose_processor.py:91-93``; the whole sequence at once would be 110 GB of host memory for config C3)
GATHER_CHUNK_BYTES = 256 << 20


@dataclass
class LiftedRecords:
    """Columnar result of the lift: one ``RECORD_DTYPE`` row per box, CSR by frame.

    ``records[frame_off[i]:frame_off[i+1]]`` are the boxes of ``frames[i]`` in input order;
    ``tail`` holds each box's passthrough ``[damage_cls, conf, label]`` (``pose_processor.py:208``)
    as three columns (labels as strings when they are not numeric)."""

    records: np.ndarray      # RECORD_DTYPE[B]
    frame_off: np.ndarray    # int64 [F+1]
    frames: np.ndarray       # [F] frame keys in bbox_coordinates order
    damage_cls: np.ndarray   # [B]
    conf: np.ndarray         # [B] float64
    label: np.ndarray        # [B]

    def save(self, path) -> None:
        """One ``.npz`` (no pickle inside): the wire format between the mapper stage and its consumers."""
        np.savez(path, records=self.records.view(np.uint8).reshape(-1, lift.RECORD_DTYPE.itemsize),
                 frame_off=self.frame_off, frames=self.frames, damage_cls=self.damage_cls, conf=self.conf,
                 label=self.label)

    @classmethod
    def load(cls, path) -> "LiftedRecords":
        with np.load(path, allow_pickle=False) as z:
            rec = np.ascontiguousarray(z["records"]).view(lift.RECORD_DTYPE).reshape(-1)
            return cls(rec, z["frame_off"], z["frames"], z["damage_cls"], z["conf"], z["label"])

    def to_rows(self) -> dict:
        """The reference's nested-list form (``pose_processor.py:115,208``).  The four corners of a box are
        ``(3,)`` float64 views of one ``[B,4,3]`` array (no per-corner copy); they pickle as independent arrays."""
        corners = self.records["corners"].astype(np.float64).reshape(-1, 3)
        cs = list(corners)  # 4*B views, made in one C-level iteration
        dc, cf, lb = self.damage_cls.tolist(), self.conf.tolist(), self.label.tolist()
        off = self.frame_off.tolist()
        out = {}
        for i, key in enumerate(self.frames.tolist()):
            out[key] = [[cs[4 * b], cs[4 * b + 1], cs[4 * b + 2], cs[4 * b + 3], dc[b], cf[b], lb[b]]
                        for b in range(off[i], off[i + 1])]
        return out


def _column(values):
    """Passthrough column as a numeric array when it is one, else as strings (labels may be class names)."""
    try:
        a = np.asarray(values)
        if a.dtype.kind in "biuf":
            return a
    except Exception:
        pass
    return np.asarray([str(v) for v in values])


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
        self.lifted = None

        self.visualiser = Visualiser()
        self.transforms = Transforms()

        logging.basicConfig(level=logging.INFO)
        self.logger = logging.getLogger(__name__)
        self.logger.info("Processing Pose.")

    # ------------------------------------------------------------------------------------
    def _gather_tables(self):
        """Everything but the depth, for the whole sequence, in ``bbox_coordinates`` order (reference loop header,
        ``pose_processor.py:89-93``): frame keys, pose rows, CSR offsets, the boxes' first four entries."""
        frames = list(self.bbox_coordinates.keys())
        F = len(frames)
        counts = np.fromiter((len(self.bbox_coordinates[k]) for k in frames), dtype=np.int64, count=F)
        frame_off = np.zeros(F + 1, dtype=np.int64)
        np.cumsum(counts, out=frame_off[1:])
        B = int(frame_off[-1])
        all_rows = chain.from_iterable(self.bbox_coordinates[k] for k in frames)
        boxes = np.array(list(map(itemgetter(0, 1, 2, 3), all_rows)), dtype=np.float64).reshape(B, 4)
        # pose row i = frame i, first column is the timestamp (pose_processor.py:109): one conversion for the whole
        # DataFrame instead of one pandas row lookup per frame (50 us each: half a second per 10 k frames)
        if hasattr(self.pose, "iloc"):
            pose_np = self.pose.iloc[:, 1:8].to_numpy(dtype=np.float64)
            pose7 = np.ascontiguousarray(pose_np[np.asarray(frames, dtype=np.int64)]) if F else np.empty((0, 7))
        else:
            pose7 = np.array([np.asarray(self.pose[k], dtype=np.float64) for k in frames], dtype=np.float64).reshape(F, 7)
        return frames, frame_off, boxes, pose7

    def _gather_depth(self, frames, depth_out, intr4, image_wh):
        """Depth + intrinsics of ``frames``: fills ``intr4`` / ``image_wh`` and returns the ``[n,H,W]`` float32 HOST
        array to lift from.  A dataset with a ``batch(frames, out)`` method (the in-memory ``ArrayDataset``) answers in
        one call and may hand back a zero-copy view of its own storage; anything else is read frame by frame into
        ``depth_out`` like the reference does (``dataset[i]``, ``pose_processor.py:93``).  (Datasets that deliver
        depth on the DEVICE -- ``lm3d.ingest.DepthSequence`` -- take ``_lift_from_device_loader`` instead.)"""
        H, W = int(self.depth_height), int(self.depth_width)
        batch = getattr(self.dataset, "batch", None)
        if batch is not None:
            cal, depth = batch(frames, depth_out)  # cal [n,6]: fx fy cx cy image_width image_height
            cal = np.asarray(cal, dtype=np.float64)
            if depth.shape != (len(frames), H, W):
                raise ValueError(f"dataset.batch returned depth of shape {depth.shape}, expected {(len(frames), H, W)}")
            s = cal[:, 4] / float(self.depth_width)
            intr4[:] = cal[:, :4] / s[:, None]
            image_wh[:] = cal[:, 4:6]
            return depth
        for i, frame_index in enumerate(frames):
            _rgb, depth_tensor, ci = self.dataset[frame_index]
            _, depth_image = self.visualiser.parse_images(None, depth_tensor)
            if depth_image.shape != (H, W):
                raise ValueError(f"frame {frame_index}: depth is {depth_image.shape}, expected {(H, W)}")
            depth_out[i] = depth_image
            # intrinsics rescale: every entry, cy included, by the WIDTH ratio (:133-137)
            s = ci["image_width"] / self.depth_width
            intr4[i] = (ci["fx"] / s, ci["fy"] / s, ci["cx"] / s, ci["cy"] / s)
            image_wh[i] = (ci["image_width"], ci["image_height"])
        return depth_out

    def _gather(self):
        """Whole-sequence form of the two gathers above (tests, small sequences)."""
        frames, frame_off, boxes, pose7 = self._gather_tables()
        F, H, W = len(frames), int(self.depth_height), int(self.depth_width)
        depth = np.empty((F, H, W), dtype=np.float32)
        intr4 = np.empty((F, 4), dtype=np.float64)
        image_wh = np.empty((F, 2), dtype=np.float64)
        depth = self._gather_depth(frames, depth, intr4, image_wh)
        return frames, depth, pose7, intr4, image_wh, frame_off, boxes

    _staging = {}  # (shape, dtype) -> pinned torch tensor, shared by the instances of a process (pinning 256 MB costs ~0.1 s)

    def _pinned(self, shape, dtype):
        """Staging buffer for one chunk: pinned when torch can pin it (H2D copies of the C call go asynchronous)."""
        key = (tuple(shape), np.dtype(dtype).name)
        t = ProcessPose._staging.get(key)
        if t is None:
            try:
                import torch

                t = torch.empty(shape, dtype=getattr(torch, np.dtype(dtype).name), pin_memory=True)
                ProcessPose._staging.clear()
                ProcessPose._staging[key] = t
            except Exception:
                return np.empty(shape, dtype=dtype)
        return t.numpy()

    # ------------------------------------------------------------------------------------
    def get_global_records(self) -> LiftedRecords:
        """Lift every box of the sequence and return the columnar result (no per-box Python objects)."""
        if self.display_rgbd or self.display_3d:
            raise NotImplementedError("display_rgbd / display_3d need the reference's Open3D GUI (out of scope)")
        frames, frame_off, boxes, pose7 = self._gather_tables()
        F, H, W = len(frames), int(self.depth_height), int(self.depth_width)
        B = int(frame_off[-1])
        rec = np.empty(B, dtype=lift.RECORD_DTYPE)
        chunk = max(1, min(F, GATHER_CHUNK_BYTES // max(1, H * W * 4)))
        if hasattr(self.dataset, "batch_device"):
            self._lift_from_device_loader(frames, frame_off, boxes, pose7, rec, chunk)
            return self._finish(frames, frame_off, rec)
        depth = self._pinned((chunk, H, W), np.float32) if F else None
        intr4 = np.empty((chunk, 4), dtype=np.float64)
        image_wh = np.empty((chunk, 2), dtype=np.float64)
        for f0 in range(0, F, chunk):
            f1 = min(F, f0 + chunk)
            n = f1 - f0
            b0, b1 = int(frame_off[f0]), int(frame_off[f1])
            if b1 == b0:
                continue  # no boxes in this chunk: its frames are not even read
            d = self._gather_depth(frames[f0:f1], depth[:n], intr4[:n], image_wh[:n])
            lift.lift_boxes_host(
                d, pose7[f0:f1], intr4[:n], boxes[b0:b1], image_wh[:n], frame_off[f0 : f1 + 1] - b0,
                scale_depth=float(self.scale_depth), max_depth_mm=float(self.max_depth_mm),
                q=float(self.percentile), device=int(self.device), out=rec[b0:b1],
            )
        return self._finish(frames, frame_off, rec)

    def _lift_from_device_loader(self, frames, frame_off, boxes, pose7, rec, chunk):
        """Datasets that deliver depth on the device (``lm3d.ingest.DepthSequence.batch_device``: PNG decode ->
        pinned ring -> H2D -> in-place conversion): the lift runs on the device tensors directly, only the records
        come back to the host."""
        import torch

        F = len(frames)
        dev = torch.device("cuda", int(self.device))
        plan = None
        for f0 in range(0, F, chunk):
            f1 = min(F, f0 + chunk)
            b0, b1 = int(frame_off[f0]), int(frame_off[f1])
            if b1 == b0:
                continue
            depth_dev, cal = self.dataset.batch_device(frames[f0:f1])
            s = cal[:, 4] / float(self.depth_width)  # intrinsics rescale by the WIDTH ratio (:133-137)
            fo = torch.from_numpy(frame_off[f0 : f1 + 1] - b0).to(dev)
            rect4 = lift.scale_boxes(torch.from_numpy(boxes[b0:b1]).to(dev), torch.from_numpy(np.ascontiguousarray(cal[:, 4:6])).to(dev),
                                     fo, int(self.depth_width), int(self.depth_height))
            if plan is None or plan.F < f1 - f0 or plan.B < b1 - b0:
                plan = lift.LiftPlan(f1 - f0, b1 - b0, dev, False, int(self.depth_height), int(self.depth_width))
            out = lift.lift_boxes(depth_dev, torch.from_numpy(pose7[f0:f1]).to(dev), torch.from_numpy(cal[:, :4] / s[:, None]).to(dev),
                                  rect4, fo, scale_depth=float(self.scale_depth), max_depth_mm=float(self.max_depth_mm),
                                  q=float(self.percentile), plan=plan)
            rec[b0:b1] = lift.records_to_numpy(out)

    def _finish(self, frames, frame_off, rec):
        F = len(frames)
        tails = list(map(itemgetter(-3, -2, -1), chain.from_iterable(self.bbox_coordinates[k] for k in frames)))
        dc, cf, lb = (list(c) for c in zip(*tails)) if tails else ([], [], [])
        self.lifted = LiftedRecords(rec, frame_off, _column(frames) if F else np.empty(0, dtype=np.int64),
                                    _column(dc), np.asarray(cf, dtype=np.float64), _column(lb))
        self.records = rec
        self.record_frames = frames
        return self.lifted

    def get_global_coordinates(self):
        lifted = self.get_global_records()
        corners = lifted.records["corners"].astype(np.float64).reshape(-1, 3)
        cs = list(corners)  # 4*B (3,) views in one C-level iteration instead of 4*B .copy() calls
        off = lifted.frame_off.tolist()
        global_bboxes = {}
        for i, frame_index in enumerate(self.record_frames):
            bxs = self.bbox_coordinates[frame_index]
            b0 = off[i]
            # row = global_corners + bbox[-3:]  (pose_processor.py:208)
            rows = [[cs[4 * (b0 + j)], cs[4 * (b0 + j) + 1], cs[4 * (b0 + j) + 2], cs[4 * (b0 + j) + 3], *bbox[-3:]]
                    for j, bbox in enumerate(bxs)]
            if self.verbose:
                for bbox, row in zip(bxs, rows):
                    self.logger.info(f"\tOriginal 2D Corners: {bbox}")
                    self.logger.info(f"\tGlobal 3D Coordinates: {row[:4]}\n")
            global_bboxes[frame_index] = rows
        return global_bboxes

    def _transform_to_global(self, local_point, pose_data):
        """Kept for callers of the reference's helper (``pose_processor.py:242-260``)."""
        transformation = self.transforms.get_transformation_matrix(pose_data)
        local_point = np.array([(*local_point, 1)])
        return (transformation @ local_point.T)[:3, 0]
