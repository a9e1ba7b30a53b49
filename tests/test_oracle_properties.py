"""Property tests of the oracle (hypothesis).  CPU only."""
import numpy as np
from hypothesis import given, settings
from hypothesis import strategies as st

from oracle import reference_numpy as ora


def _case(seed, H=24, W=20):
    rng = np.random.default_rng(seed)
    depth = (800 + 900 * rng.random((H, W))).astype(np.float32)
    depth[rng.random((H, W)) < 0.05] = 0
    q = rng.normal(size=4)
    pose = np.concatenate([rng.normal(size=3), q / np.linalg.norm(q)])
    x0, x1 = sorted(rng.integers(0, W, 2).tolist())
    y0, y1 = sorted(rng.integers(0, H, 2).tolist())
    return depth, pose, (x0, y0, x1, y1)


@settings(max_examples=40, deadline=None)
@given(st.integers(0, 10_000))
def test_translation_equivariance(seed):
    depth, pose, rect = _case(seed)
    a = ora.lift_box(depth, rect, ora.get_transformation_matrix(pose), 15.0, 15.0, 10.0, 12.0)
    shift = np.array([3.0, -2.0, 7.0])
    pose2 = pose.copy()
    pose2[:3] += shift
    b = ora.lift_box(depth, rect, ora.get_transformation_matrix(pose2), 15.0, 15.0, 10.0, 12.0)
    if int(a["n_valid"]) == 0:
        return
    for k in ("centroid", "aabb_min", "aabb_max"):
        np.testing.assert_allclose(b[k], a[k] + shift, atol=1e-9)
    np.testing.assert_allclose(b["corners"], a["corners"] + shift, atol=1e-9)
    assert b["z_q"] == a["z_q"] and b["n_valid"] == a["n_valid"]


@settings(max_examples=40, deadline=None)
@given(st.integers(0, 10_000))
def test_box_growth_is_monotone_and_aabb_contains_centroid(seed):
    depth, pose, (x0, y0, x1, y1) = _case(seed)
    T = ora.get_transformation_matrix(pose)
    small = ora.lift_box(depth, (x0, y0, x1, y1), T, 15.0, 15.0, 10.0, 12.0)
    H, W = depth.shape
    big = ora.lift_box(depth, (max(x0 - 2, 0), max(y0 - 2, 0), min(x1 + 2, W - 1), min(y1 + 2, H - 1)), T, 15.0, 15.0, 10.0, 12.0)
    assert big["n_pix"] >= small["n_pix"] and big["n_valid"] >= small["n_valid"]
    if int(small["n_valid"]):
        assert (big["aabb_min"] <= small["aabb_min"] + 1e-12).all() and (big["aabb_max"] >= small["aabb_max"] - 1e-12).all()
        assert (small["aabb_min"] - 1e-9 <= small["centroid"]).all() and (small["centroid"] <= small["aabb_max"] + 1e-9).all()
        assert small["d_lo"] <= small["d_hi"]


@settings(max_examples=15, deadline=None)
@given(st.integers(0, 10_000))
def test_loop_form_equals_batched_form(seed):
    from lm3d import synth

    seq = synth.make_sequence(2, 32, 24, 3, seed=seed)
    rows = ora.get_global_coordinates_loop(seq.pose7, seq.dataset(), seq.bbox_coordinates(), 24, 32)
    rect4 = ora.boxes_to_rects(seq.boxes.reshape(-1, 4), np.repeat(seq.image_wh(), 3, axis=0), (24, 32))
    rec = ora.lift_boxes(seq.depth, seq.pose7, seq.intr4_depth_res(), rect4, seq.frame_off())
    for f in range(2):
        for b in range(3):
            np.testing.assert_allclose(np.stack(rows[f][b][:4]), rec["corners"][f * 3 + b], atol=1e-12, equal_nan=True)
