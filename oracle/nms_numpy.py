"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the 3-D NMS that follows the lift.

Only ``tests/`` may import this module; the product path (``lm3d.nms``, ``liblm3d.so``) never does.

PARITY UNPINNED.  The reference calls ``BoundingBoxProcessor(global_bboxes_data, pose_df).suppress_bboxes()``
(``/root/reference/task_def.py:145-149``) but its source, ``src/mapper/bbox_optimiser.py``, is not in the
repository, and no test or fixture pins its output.  What IS pinned by the callers:

* input  = the result of ``ProcessPose.get_global_coordinates()``: ``{frame: [[c0, c1, c2, c3, damage_cls, conf,
  label], ...]}``, ``c_i`` world XYZ (``pose_processor.py:115,208``);
* output = a dict iterated with ``.items()`` whose values are lists of rows read as ``bbox[:4]`` = four world
  corners (``src/mapper/mapping.py:170-176``), i.e. rows of the same shape as the input rows.

NMS-SPEC v0 (every rule DEFINED here, fp32 arithmetic, one rounding per operation, no fused multiply-add):

* N1  extent of a box = axis-aligned bounds of its four world corners, grown by ``pad`` metres on every side
      (a sign seen head-on has no thickness along its normal; ``pad`` = the reference's ``bbox_depth_buffer``,
      ``pose_processor.py:50``, default 0.03): ``lo = min_c(corner) - pad``, ``hi = max_c(corner) + pad``.
* N2  a box takes part iff its 12 corner coordinates AND its confidence are finite (lift records with
      ``n_valid == 0`` are NaN; a NaN confidence has no place in the greedy order).
* N3  volume ``vol = ((hi.x - lo.x) * (hi.y - lo.y)) * (hi.z - lo.z)``.
* N4  for two boxes ``d_k = min(hi_a.k, hi_b.k) - max(lo_a.k, lo_b.k)``; they overlap iff every ``d_k > 0`` and
      ``inter = (d.x * d.y) * d.z``, ``union = (vol_a + vol_b) - inter``, ``inter > thr * union``.
* N5  only boxes with the same ``label`` compete.
* N6  greedy order: confidence descending, ties by ascending box index.  A box is KEPT iff no KEPT box earlier
      in that order overlaps it (N4, N5); otherwise it is suppressed by the FIRST such box in the order.
* N7  result: ``keep[B]`` (uint8) and ``parent[B]`` (int32: own index if kept, index of the suppressing kept box,
      -1 if the box does not take part).
"""
from __future__ import annotations

import numpy as np

F32 = np.float32


def box_extents(corners: np.ndarray, pad: float):
    """N1-N3.  corners ``[B,4,3]`` float32 -> lo ``[B,3]``, hi ``[B,3]``, vol ``[B]``, valid ``[B]``."""
    c = np.asarray(corners, dtype=F32).reshape(-1, 4, 3)
    valid = np.isfinite(c).all(axis=(1, 2))
    with np.errstate(invalid="ignore"):
        lo = (c.min(axis=1) - F32(pad)).astype(F32)
        hi = (c.max(axis=1) + F32(pad)).astype(F32)
        e = (hi - lo).astype(F32)
        vol = ((e[:, 0] * e[:, 1]).astype(F32) * e[:, 2]).astype(F32)
    return lo, hi, vol, valid


def overlaps(lo_a, hi_a, vol_a, lo_b, hi_b, vol_b, thr: float) -> np.ndarray:
    """N4 for one box a against arrays of boxes b (vectorised over b)."""
    d = (np.minimum(hi_a, hi_b) - np.maximum(lo_a, lo_b)).astype(F32)
    pos = (d > 0).all(axis=-1)
    inter = ((d[..., 0] * d[..., 1]).astype(F32) * d[..., 2]).astype(F32)
    union = ((vol_a + vol_b).astype(F32) - inter).astype(F32)
    return pos & (inter > (F32(thr) * union).astype(F32))


def nms_3d(corners, conf, label, thr: float = 0.1, pad: float = 0.03):
    """N6/N7, the plain sequential greedy loop (O(B * kept))."""
    lo, hi, vol, valid = box_extents(corners, pad)
    conf = np.asarray(conf, dtype=F32)
    valid = valid & np.isfinite(conf)  # N2
    label = np.asarray(label, dtype=np.int32)
    B = lo.shape[0]
    keep = np.zeros(B, dtype=np.uint8)
    parent = np.full(B, -1, dtype=np.int32)
    order = np.lexsort((np.arange(B), -np.where(valid, conf, 0).astype(np.float64)))  # conf descending, index ascending
    kept: list[int] = []
    for i in order:
        if not valid[i]:
            continue
        if kept:
            k = np.asarray(kept)
            hit = overlaps(lo[i], hi[i], vol[i], lo[k], hi[k], vol[k], thr) & (label[k] == label[i])
            if hit.any():
                parent[i] = k[np.argmax(hit)]  # kept[] is in greedy order: the first hit is the suppressor
                continue
        keep[i] = 1
        parent[i] = i
        kept.append(int(i))
    return keep, parent


def suppress_rows(global_bboxes_data: dict, thr: float = 0.1, pad: float = 0.03) -> dict:
    """Row-level form (the shape ``BoundingBoxProcessor.suppress_bboxes`` returns): every input frame key, in input
    order, with the kept rows of that frame in input order."""
    rows = [(f, r) for f, lst in global_bboxes_data.items() for r in lst]
    if not rows:
        return {f: [] for f in global_bboxes_data}
    corners = np.array([[np.asarray(c, dtype=np.float64) for c in r[:4]] for _, r in rows], dtype=F32)
    conf = np.array([float(r[5]) for _, r in rows], dtype=F32)
    ids: dict = {}
    label = np.array([ids.setdefault(r[6], len(ids)) for _, r in rows], dtype=np.int32)
    keep, _ = nms_3d(corners, conf, label, thr, pad)
    out = {f: [] for f in global_bboxes_data}
    for (f, r), k in zip(rows, keep):
        if k:
            out[f].append(r)
    return out
