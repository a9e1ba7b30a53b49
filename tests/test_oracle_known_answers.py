"""Hand-computed known answers for ORACLE-SPEC v0.  CPU only."""
import numpy as np
import pytest

from oracle import reference_numpy as ora

IDENT = np.array([0, 0, 0, 0, 0, 0, 1.0])


def test_identity_pose_and_quaternion_normalisation():
    np.testing.assert_allclose(ora.get_transformation_matrix(IDENT), np.eye(4))
    np.testing.assert_allclose(ora.get_transformation_matrix([1, 2, 3, 0, 0, 0, 5.0])[:3, :3], np.eye(3))
    np.testing.assert_allclose(ora.get_transformation_matrix([1, 2, 3, 0, 0, 0, 5.0])[:3, 3], [1, 2, 3])


def test_rotation_90_about_z_scalar_last():
    s = np.sqrt(0.5)
    T = ora.get_transformation_matrix([0, 0, 0, 0, 0, s, s])  # qz=qw -> +90 deg about z
    np.testing.assert_allclose(T[:3, :3] @ [1, 0, 0], [0, 1, 0], atol=1e-15)
    np.testing.assert_allclose(ora.transform_to_global([1.0, 0.0, 2.0], [5, 6, 7, 0, 0, s, s]), [5, 7, 9], atol=1e-15)


def test_intrinsics_rescale_uses_width_ratio_for_cy_too():
    fx, fy, cx, cy = ora.rescale_intrinsics(dict(image_width=1440, image_height=1920, fx=1450.0, fy=1400.0, cx=720.0, cy=960.0), 192)
    s = 1440 / 192
    assert (fx, fy, cx, cy) == (1450.0 / s, 1400.0 / s, 720.0 / s, 960.0 / s)


def test_pixel_rect_truncation_clamp_and_ordering():
    assert ora.pixel_rect([10.9, 20.9, 30.1, 40.999], 192, 256) == (10, 20, 30, 40)
    assert ora.pixel_rect([0.0, 0.0, 192.0, 256.0], 192, 256) == (0, 0, 191, 255)       # int(192.0) is out of range
    assert ora.pixel_rect([-3.5, -0.2, 5.0, 5.0], 192, 256) == (0, 0, 5, 5)             # int() truncates toward zero
    assert ora.pixel_rect([30.0, 40.0, 10.0, 20.0], 192, 256) == (10, 20, 30, 40)       # inverted box
    assert ora.rect_corners((1, 2, 3, 4)) == [(1, 2), (1, 4), (3, 4), (3, 2)]           # TL, BL, BR, TR
    r = ora.boxes_to_rects(np.array([[10.9, 20.9, 30.1, 40.999], [0, 0, 1440, 1920.0]]) * [[1], [1]],
                           np.array([[192.0, 256.0], [1440.0, 1920.0]]), (192, 256))
    assert r.tolist() == [[10, 20, 30, 40], [0, 0, 191, 255]]


def test_valid_mask():
    d = np.array([0.0, -1.0, 1.0, np.nan, np.inf, 1500.0, 3000.0], dtype=np.float32)
    assert ora.valid_mask(d).tolist() == [False, False, True, False, False, True, True]
    assert ora.valid_mask(d, 2000.0).tolist() == [False, False, True, False, False, True, False]


@pytest.mark.parametrize("vals,q,dq,lo,hi", [
    ([5.0], 50, 5.0, 5.0, 5.0),
    ([1, 2, 3, 4], 50, 2.5, 2.0, 3.0),         # even count -> interpolated
    ([3, 1, 2], 50, 2.0, 2.0, 2.0),            # odd count -> exact element
    ([1, 2, 3, 4, 5], 25, 2.0, 2.0, 2.0),
    ([10, 20], 75, 17.5, 10.0, 20.0),
    ([1, 2, 3, 4], 0, 1.0, 1.0, 1.0),
    ([1, 2, 3, 4], 100, 4.0, 4.0, 4.0),
])
def test_percentile_depth(vals, q, dq, lo, hi):
    got = ora.percentile_depth(np.array(vals, dtype=np.float32), q)
    assert got[0] == pytest.approx(dq, abs=1e-12) and float(got[1]) == lo and float(got[2]) == hi


def test_constant_plane_box_closed_form():
    H, W = 8, 6
    depth = np.full((H, W), 2000.0, dtype=np.float32)
    fx = fy = 4.0
    cx, cy = 2.0, 3.0
    rec = ora.lift_box(depth, (1, 2, 4, 5), np.eye(4), fx, fy, cx, cy)
    assert int(rec["n_pix"]) == 16 and int(rec["n_valid"]) == 16
    assert float(rec["z_q"]) == 2.0
    # corners: ((u-cx)z/fx, (v-cy)z/fy, z) at (1,2),(1,5),(4,5),(4,2)
    np.testing.assert_allclose(rec["corners"], [[-0.5, -0.5, 2], [-0.5, 1.0, 2], [1.0, 1.0, 2], [1.0, -0.5, 2]])
    np.testing.assert_allclose(rec["centroid"], [0.25, 0.25, 2.0])
    np.testing.assert_allclose(rec["aabb_min"], [-0.5, -0.5, 2.0])
    np.testing.assert_allclose(rec["aabb_max"], [1.0, 1.0, 2.0])


def test_single_pixel_and_all_invalid():
    depth = np.array([[0.0, 1500.0], [np.nan, -2.0]], dtype=np.float32)
    rec = ora.lift_box(depth, (1, 0, 1, 0), np.eye(4), 1.0, 1.0, 0.0, 0.0)
    assert int(rec["n_valid"]) == 1 and float(rec["d_lo"]) == 1500.0
    np.testing.assert_allclose(rec["centroid"], [1.5, 0.0, 1.5])
    rec = ora.lift_box(depth, (0, 0, 0, 1), np.eye(4), 1.0, 1.0, 0.0, 0.0)
    assert int(rec["n_valid"]) == 0 and int(rec["n_pix"]) == 2
    for k in ("corners", "centroid", "aabb_min", "aabb_max", "z_q"):
        assert np.isnan(rec[k]).all()


def test_union_pixels():
    rect4 = np.array([[0, 0, 1, 1], [1, 1, 2, 2], [5, 5, 5, 5]], dtype=np.int32)
    U = ora.union_pixels_per_frame(rect4, np.array([0, 2, 3]), 8, 8)
    assert U.tolist() == [7, 1]


def test_full_frame_cloud_matches_per_box_lift():
    rng = np.random.default_rng(1)
    depth = (1000 + 100 * rng.random((6, 5))).astype(np.float32)
    depth[2, 3] = 0
    T = ora.get_transformation_matrix([0.1, 0.2, 0.3, 0.1, -0.2, 0.3, 0.9])
    pts = ora.full_frame_cloud(depth, T, 3.0, 3.5, 2.0, 2.5)
    assert pts.shape == (29, 3)
    rec = ora.lift_box(depth, (0, 0, 4, 5), T, 3.0, 3.5, 2.0, 2.5)
    np.testing.assert_allclose(pts.mean(0), rec["centroid"], rtol=1e-12)
    np.testing.assert_allclose(pts.min(0), rec["aabb_min"], rtol=1e-12)


def test_rotation_rule_r3_against_scipy():
    """R3 (quaternion -> rotation, scalar LAST as in RTAB-Map's pose file, src/mapper/database_query.py:22) checked
    against an independent implementation: scipy.spatial.transform.Rotation.from_quat is scalar-last too."""
    Rotation = pytest.importorskip("scipy.spatial.transform").Rotation
    rng = np.random.default_rng(11)
    for _ in range(200):
        q = rng.normal(size=4) * rng.uniform(0.1, 10.0)        # un-normalised on purpose: R3 normalises first
        t = rng.normal(size=3)
        T = ora.get_transformation_matrix([*t, *q])
        np.testing.assert_allclose(T[:3, :3], Rotation.from_quat(q).as_matrix(), atol=1e-14)
        np.testing.assert_allclose(T[:3, 3], t)
        p = rng.normal(size=3)
        np.testing.assert_allclose(ora.transform_to_global(p, [*t, *q]), Rotation.from_quat(q).apply(p) + t, atol=1e-13)
    # and the shipped Transforms (the class the reference's ProcessPose calls) is the same rule
    from src.utils.transformations import Transforms

    q = np.array([0.1, -0.7, 0.3, 0.64])
    np.testing.assert_allclose(Transforms().get_rotation([0, 0, 0, *q]), Rotation.from_quat(q).as_matrix(), atol=1e-14)
