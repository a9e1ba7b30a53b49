"""CPU oracle for the src/mapper 2D-box -> 3D lift.  TEST INFRASTRUCTURE ONLY.

This file is the checker, never the product: only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it.  The product
path (``lm3d`` + ``csrc``) never does and fails loudly when the CUDA library is missing.

PARITY STATUS: **partially pinned**.  The reference ships the loop structure, the intrinsics
rescale, the ``int()`` pixel truncation, the homogeneous pose multiply and the output row
format (``/root/reference/src/mapper/pose_processor.py:88-260``), but the arithmetic helpers it
calls (``src/utils/transformations.py``, ``src/utils/visualisation.py``) are absent from the
reference tree and it has no tests or golden vectors.  What CAN be pinned is pinned:
``tests/golden/make_golden.py`` executes the reference's own ``ProcessPose`` class from
``/root/reference`` (absent modules stubbed, this file's helpers standing in for the absent
``Transforms``) and commits its rows; ``tests/test_oracle_golden.py`` checks this oracle
against them.  The un-shipped helper arithmetic follows SURVEY.md section 8c "ORACLE-SPEC v0"
(rules R1..R12); every function below cites the rule and the reference line it follows.

All maths is fp64 numpy on fp32 depth, like the reference (Python floats / numpy defaults).
Two forms: a *loop form* shaped like the reference (per frame, per box, pose matrix rebuilt
per corner) used as the CPU baseline, and a *vectorised form* used by the parity tests.
"""
from __future__ import annotations

import numpy as np

# ---------------------------------------------------------------------------------------
# Record layout (mirrors include/lm3d.h :: lm3d_box_out, but fp64 where the kernel is fp32)
# ---------------------------------------------------------------------------------------
ORACLE_RECORD = np.dtype(
    [
        ("corners", np.float64, (4, 3)),
        ("centroid", np.float64, (3,)),
        ("aabb_min", np.float64, (3,)),
        ("aabb_max", np.float64, (3,)),
        ("z_q", np.float64),
        ("n_valid", np.int32),
        ("n_pix", np.int32),
        # the two selected order statistics (raw fp32 millimetres) -- bit-exact parity fields
        ("d_lo", np.float32),
        ("d_hi", np.float32),
    ]
)


# ---------------------------------------------------------------------------------------
# R3 / R4 : pose -> 4x4, point transform
# ---------------------------------------------------------------------------------------
def get_transformation_matrix(pose7) -> np.ndarray:
    """R3: ``[tx,ty,tz,qx,qy,qz,qw]`` -> 4x4 camera->world ``[[R(q),t],[0,1]]``.

    Restates the absent ``Transforms.get_transformation_matrix`` as called at
    ``pose_processor.py:140`` and ``:254``; quaternion is scalar-last as in the RTAB-Map pose
    file (``src/mapper/database_query.py:22``).  The quaternion is normalised first.
    """
    p = np.asarray(pose7, dtype=np.float64)
    tx, ty, tz, x, y, z, w = (float(v) for v in p)
    n = np.sqrt(x * x + y * y + z * z + w * w)
    x, y, z, w = x / n, y / n, z / n, w / n
    T = np.eye(4, dtype=np.float64)
    T[0, 0] = 1.0 - 2.0 * (y * y + z * z)
    T[0, 1] = 2.0 * (x * y - z * w)
    T[0, 2] = 2.0 * (x * z + y * w)
    T[1, 0] = 2.0 * (x * y + z * w)
    T[1, 1] = 1.0 - 2.0 * (x * x + z * z)
    T[1, 2] = 2.0 * (y * z - x * w)
    T[2, 0] = 2.0 * (x * z - y * w)
    T[2, 1] = 2.0 * (y * z + x * w)
    T[2, 2] = 1.0 - 2.0 * (x * x + y * y)
    T[0, 3], T[1, 3], T[2, 3] = tx, ty, tz
    return T


def transform_to_global(local_point, pose7) -> np.ndarray:
    """R4: ``(T @ [X,Y,Z,1]^T)[:3]`` -- follows ``pose_processor.py:242-260`` line by line
    (the matrix is rebuilt from the pose on every call, exactly as the reference does)."""
    T = get_transformation_matrix(pose7)
    hp = np.array([(*local_point, 1)], dtype=np.float64)
    return (T @ hp.T)[:3, 0]


# ---------------------------------------------------------------------------------------
# R1 : intrinsics rescale
# ---------------------------------------------------------------------------------------
def rescale_intrinsics(camera_intrinsics: dict, depth_width: int):
    """R1, ``pose_processor.py:133-137``: every intrinsic (``cy`` included) is divided by the
    WIDTH ratio ``image_width / depth_width``."""
    s = camera_intrinsics["image_width"] / depth_width
    return (
        camera_intrinsics["fx"] / s,
        camera_intrinsics["fy"] / s,
        camera_intrinsics["cx"] / s,
        camera_intrinsics["cy"] / s,
    )


# ---------------------------------------------------------------------------------------
# R5 / R6 / R7 : box scaling, pixel rect, corner order
# ---------------------------------------------------------------------------------------
def scale_bounding_box(bbox, image_size, depth_size):
    """R5 (absent ``Transforms.scale_bounding_box``, call site ``pose_processor.py:174-178``):
    ``x*dw/iw``, ``y*dh/ih`` in fp64 on the first four entries; the tail is passed through."""
    iw, ih = image_size
    dw, dh = depth_size
    out = list(bbox)
    out[0] = float(bbox[0]) * dw / iw
    out[1] = float(bbox[1]) * dh / ih
    out[2] = float(bbox[2]) * dw / iw
    out[3] = float(bbox[3]) * dh / ih
    return out


def pixel_rect(scaled_bbox, depth_width: int, depth_height: int):
    """R6: ``int()`` truncation toward zero (``pose_processor.py:186-187``) then clamp to the
    frame; rect is inclusive on both ends, ``x0<=x1``, ``y0<=y1``."""

    def px(v, hi):
        return min(max(int(v), 0), hi)

    xa, xb = px(scaled_bbox[0], depth_width - 1), px(scaled_bbox[2], depth_width - 1)
    ya, yb = px(scaled_bbox[1], depth_height - 1), px(scaled_bbox[3], depth_height - 1)
    return min(xa, xb), min(ya, yb), max(xa, xb), max(ya, yb)


def rect_corners(rect):
    """R7: TL, BL, BR, TR on the integer rect -- the order of the one in-repo precedent,
    ``src/detector/detector.py:202`` (``[x1,y1],[x1,y2],[x2,y2],[x2,y1]``)."""
    x0, y0, x1, y1 = rect
    return [(x0, y0), (x0, y1), (x1, y1), (x1, y0)]


# ---------------------------------------------------------------------------------------
# R8 / R9 : validity, percentile depth
# ---------------------------------------------------------------------------------------
def valid_mask(d: np.ndarray, max_depth_mm: float = np.inf) -> np.ndarray:
    """R8: valid <=> finite and ``0 < d <= max_depth_mm`` (depth is fp32 millimetres,
    ``src/detector/dataset.py:76-77``)."""
    with np.errstate(invalid="ignore"):
        return np.isfinite(d) & (d > 0) & (d <= max_depth_mm)


def percentile_depth(valid_d: np.ndarray, q: float):
    """R9: ``numpy.percentile(valid d, q, method="linear")`` in fp64.  Returns
    ``(d_q, d_lo, d_hi)`` where ``d_lo``/``d_hi`` are the two order statistics (fp32) the
    interpolation used.  Empty input -> NaNs."""
    n = int(valid_d.size)
    if n == 0:
        return float("nan"), np.float32("nan"), np.float32("nan")
    s = np.sort(valid_d.astype(np.float32, copy=False), kind="stable")
    d_q = float(np.percentile(s.astype(np.float64), q, method="linear"))
    h = (n - 1) * (q / 100.0)
    lo = int(np.floor(h))
    hi = min(lo + 1, n - 1)
    if h == lo:  # no interpolation partner is touched
        hi = lo
    return d_q, s[lo], s[hi]


def depth_to_3d(x: int, y: int, d_q: float, fx, fy, cx, cy, scale_depth):
    """R10 (absent ``Transforms._depth_to_3d``, call site ``pose_processor.py:184-196``):
    pixel + the box's percentile depth -> camera-frame ``[X,Y,Z]``; ``u``=col, ``v``=row,
    integer pixel centres, no +0.5."""
    z = d_q / scale_depth
    return np.array([(x - cx) * z / fx, (y - cy) * z / fy, z], dtype=np.float64)


# ---------------------------------------------------------------------------------------
# R8..R11 : one box -> record
# ---------------------------------------------------------------------------------------
def lift_box(depth, rect, T, fx, fy, cx, cy, scale_depth=1000.0, max_depth_mm=np.inf, q=50.0):
    """One box -> ``ORACLE_RECORD`` scalar (rules R6..R11).  ``depth`` is ``[H,W]`` fp32 mm."""
    x0, y0, x1, y1 = rect
    rec = np.zeros((), dtype=ORACLE_RECORD)
    patch = depth[y0 : y1 + 1, x0 : x1 + 1]
    rec["n_pix"] = patch.size
    m = valid_mask(patch, max_depth_mm)
    vd = patch[m]
    n_valid = int(vd.size)
    rec["n_valid"] = n_valid
    if n_valid == 0:
        for k in ("corners", "centroid", "aabb_min", "aabb_max", "z_q"):
            rec[k] = np.nan
        rec["d_lo"] = np.nan
        rec["d_hi"] = np.nan
        return rec
    d_q, d_lo, d_hi = percentile_depth(vd, q)
    rec["d_lo"], rec["d_hi"] = d_lo, d_hi
    rec["z_q"] = d_q / scale_depth
    R, t = T[:3, :3], T[:3, 3]
    for i, (cxp, cyp) in enumerate(rect_corners(rect)):
        rec["corners"][i] = R @ depth_to_3d(cxp, cyp, d_q, fx, fy, cx, cy, scale_depth) + t
    # R11: per-pixel lift of the valid pixels (the reference's full-frame Open3D unprojection,
    # pose_processor.py:154-156/262-271, restricted to the box)
    vv, uu = np.nonzero(m)
    u = (uu + x0).astype(np.float64)
    v = (vv + y0).astype(np.float64)
    z = vd.astype(np.float64) / scale_depth
    pts = np.stack([(u - cx) * z / fx, (v - cy) * z / fy, z], axis=0)  # [3,n]
    w = R @ pts + t[:, None]
    rec["centroid"] = w.mean(axis=1)
    rec["aabb_min"] = w.min(axis=1)
    rec["aabb_max"] = w.max(axis=1)
    return rec


# ---------------------------------------------------------------------------------------
# Loop form: shaped like ProcessPose._3d_processing / get_global_coordinates
# ---------------------------------------------------------------------------------------
def process_frame_loop(
    pose_data,
    depth_image,
    bboxes,
    camera_intrinsics,
    depth_width,
    depth_height,
    scale_depth=1000.0,
    max_depth_mm=np.inf,
    q=50.0,
    with_cloud=False,
    rich=False,
):
    """Per-frame body, same step order as ``pose_processor.py:124-240``:
    rescale intrinsics (:133-137), pose matrix (:140), [optional dead full-frame cloud
    :154-156], per box: scale (:174-178) -> corners (:181) -> ``_depth_to_3d(int(x),int(y))``
    x4 (:184-196) -> ``_transform_to_global`` x4 with the matrix re-derived per corner
    (:199-201,:254) -> row = corners + bbox[-3:] (:208).
    ``rich=True`` additionally returns the per-box ORACLE_RECORDs (R11 extras)."""
    fx, fy, cx, cy = rescale_intrinsics(camera_intrinsics, depth_width)
    T = get_transformation_matrix(pose_data)
    if with_cloud:
        full_frame_cloud(depth_image, T, fx, fy, cx, cy, scale_depth)
    rows, recs = [], []
    for bbox in bboxes:
        scaled = scale_bounding_box(
            bbox,
            (camera_intrinsics["image_width"], camera_intrinsics["image_height"]),
            (depth_width, depth_height),
        )
        rect = pixel_rect(scaled, depth_width, depth_height)
        x0, y0, x1, y1 = rect
        patch = depth_image[y0 : y1 + 1, x0 : x1 + 1]
        vd = patch[valid_mask(patch, max_depth_mm)]
        d_q, _, _ = percentile_depth(vd, q)
        corners_3d = [
            depth_to_3d(int(x), int(y), d_q, fx, fy, cx, cy, scale_depth)
            for x, y in rect_corners(rect)
        ]
        global_corners = [transform_to_global(c, pose_data) for c in corners_3d]
        rows.append(global_corners + list(bbox[-3:]))
        if rich:
            recs.append(lift_box(depth_image, rect, T, fx, fy, cx, cy, scale_depth, max_depth_mm, q))
    return (rows, recs) if rich else rows


def full_frame_cloud(depth_image, T, fx, fy, cx, cy, scale_depth=1000.0, depth_trunc=np.inf):
    """a7 (``pose_processor.py:262-271``, Open3D semantics restated): ``z=d/scale``; pixels
    with ``z<=0``, non-finite or ``z>=depth_trunc`` dropped; ``p=((u-cx)z/fx,(v-cy)z/fy,z)``;
    world = ``T . p``.  Returns ``[n,3]`` fp64 in row-major pixel order."""
    H, W = depth_image.shape
    z = depth_image.astype(np.float64) / scale_depth
    with np.errstate(invalid="ignore"):
        m = np.isfinite(z) & (z > 0) & (z < depth_trunc)
    vv, uu = np.nonzero(m)
    zz = z[m]
    pts = np.stack([(uu - cx) * zz / fx, (vv - cy) * zz / fy, zz], axis=0)
    return (T[:3, :3] @ pts + T[:3, 3:4]).T


def get_global_coordinates_loop(
    pose_rows,
    dataset,
    bbox_coordinates,
    depth_width,
    depth_height,
    scale_depth=1000.0,
    max_depth_mm=np.inf,
    q=50.0,
    with_cloud=False,
):
    """Frame loop of ``ProcessPose.get_global_coordinates`` (``pose_processor.py:88-122``).
    ``pose_rows[i]`` is the 7-vector of frame ``i`` (``:109``); ``dataset[i]`` returns
    ``(rgb, depth[H,W] fp32 mm, intrinsics dict)`` (``src/detector/dataset.py:66``)."""
    out = {}
    for frame_index, bboxes in bbox_coordinates.items():
        _, depth, intr = dataset[frame_index]
        depth = np.asarray(depth, dtype=np.float32)
        out[frame_index] = process_frame_loop(
            pose_rows[frame_index],
            depth,
            bboxes,
            intr,
            depth_width,
            depth_height,
            scale_depth,
            max_depth_mm,
            q,
            with_cloud,
        )
    return out


# ---------------------------------------------------------------------------------------
# Vectorised / batched form (the parity checker for the C-ABI call)
# ---------------------------------------------------------------------------------------
def boxes_to_rects(boxes_xyxy, image_wh, depth_wh):
    """R5+R6 over an array: ``boxes_xyxy [B,4]`` fp64 RGB pixels, ``image_wh [B,2]`` (or
    ``[2]``) -> int32 ``[B,4]`` inclusive clamped rects.  Same op order as the scalar form
    (``x*dw/iw``), so results are bit-identical to it."""
    b = np.asarray(boxes_xyxy, dtype=np.float64).reshape(-1, 4)
    iwh = np.broadcast_to(np.asarray(image_wh, dtype=np.float64), (b.shape[0], 2))
    dw, dh = float(depth_wh[0]), float(depth_wh[1])
    xs0 = b[:, 0] * dw / iwh[:, 0]
    ys0 = b[:, 1] * dh / iwh[:, 1]
    xs1 = b[:, 2] * dw / iwh[:, 0]
    ys1 = b[:, 3] * dh / iwh[:, 1]

    def px(v, hi):
        with np.errstate(invalid="ignore"):
            t = np.trunc(v)
        t = np.where(np.isnan(t), 0.0, t)
        return np.clip(t, 0, hi).astype(np.int32)

    xa, xb = px(xs0, depth_wh[0] - 1), px(xs1, depth_wh[0] - 1)
    ya, yb = px(ys0, depth_wh[1] - 1), px(ys1, depth_wh[1] - 1)
    return np.stack(
        [np.minimum(xa, xb), np.minimum(ya, yb), np.maximum(xa, xb), np.maximum(ya, yb)], axis=1
    ).astype(np.int32)


def lift_boxes(
    depth,
    pose7,
    intr4,
    rect4,
    frame_off,
    scale_depth=1000.0,
    max_depth_mm=np.inf,
    q=50.0,
):
    """Batched checker with the C-ABI's argument meaning (``include/lm3d.h ::
    lm3d_lift_boxes``): ``depth [F,H,W]`` fp32 mm, ``pose7 [F,7]`` fp64, ``intr4 [F,4]``
    fp64 already at depth resolution (R1 applied), ``rect4 [B,4]`` int32 inclusive,
    ``frame_off [F+1]`` CSR.  Returns ``ORACLE_RECORD[B]``."""
    depth = np.asarray(depth)
    B = int(rect4.shape[0])
    out = np.zeros(B, dtype=ORACLE_RECORD)
    F = depth.shape[0]
    for f in range(F):
        b0, b1 = int(frame_off[f]), int(frame_off[f + 1])
        if b0 == b1:
            continue
        T = get_transformation_matrix(pose7[f])
        fx, fy, cx, cy = (float(v) for v in intr4[f])
        for b in range(b0, b1):
            out[b] = lift_box(
                depth[f], tuple(int(v) for v in rect4[b]), T, fx, fy, cx, cy, scale_depth, max_depth_mm, q
            )
    return out


def union_pixels_per_frame(rect4, frame_off, H, W):
    """``U_f`` = number of distinct pixels covered by >=1 rect of frame ``f``
    (2-D difference array).  Used by the bench harness for algorithmic bytes (untimed)."""
    F = len(frame_off) - 1
    U = np.zeros(F, dtype=np.int64)
    for f in range(F):
        diff = np.zeros((H + 1, W + 1), dtype=np.int32)
        for b in range(int(frame_off[f]), int(frame_off[f + 1])):
            x0, y0, x1, y1 = (int(v) for v in rect4[b])
            diff[y0, x0] += 1
            diff[y0, x1 + 1] -= 1
            diff[y1 + 1, x0] -= 1
            diff[y1 + 1, x1 + 1] += 1
        cover = diff.cumsum(0).cumsum(1)[:H, :W]
        U[f] = int((cover > 0).sum())
    return U


# ---------------------------------------------------------------------------------------
# Depth / calibration ingest (SURVEY.md 8f "next" #2).  Unlike the lift helpers this part of the
# reference IS in the tree, and tests/golden/depth_ingest.npz is produced by running it unmodified.
# ---------------------------------------------------------------------------------------
def decode_depth_8uc4(raw_u8: np.ndarray, scale: float = 1000.0) -> np.ndarray:
    """``[...,H,W,4]`` uint8 (a depth PNG as ``cv2.imread(..., IMREAD_UNCHANGED)`` returns it: the four bytes of
    each fp32 METRE value) -> ``[...,H,W]`` float32 MILLIMETRES.  Follows
    ``/root/reference/src/detector/dataset.py:70-77``: reinterpret the bytes as float32, multiply IN float32."""
    raw_u8 = np.ascontiguousarray(raw_u8, dtype=np.uint8)
    if raw_u8.shape[-1] != 4:
        raise ValueError("expected [...,H,W,4] uint8")
    metres = raw_u8.view(np.float32).reshape(raw_u8.shape[:-1])
    return metres * np.float32(scale)


def intrinsics_row(calibration: dict, depth_width: int) -> np.ndarray:
    """Calibration dict of ``dataset.py:102-121`` -> ``[fx,fy,cx,cy]`` at depth resolution (R1,
    ``pose_processor.py:133-137``: all four divided by the WIDTH ratio)."""
    fx, fy, cx, cy = rescale_intrinsics(calibration, depth_width)
    return np.array([fx, fy, cx, cy], dtype=np.float64)
