"""Host-side logic (no GPU): synthetic generator, Transforms mirror, pose parsing, batching."""
import numpy as np
import pytest

from oracle import reference_numpy as ora


def test_synth_is_deterministic_and_shaped():
    from lm3d import synth

    a = synth.make_config("C1", frames=3)
    b = synth.make_config("C1", frames=3)
    assert a.depth.shape == (3, 256, 192) and a.depth.dtype == np.float32
    assert np.array_equal(a.depth, b.depth, equal_nan=True) and np.array_equal(a.pose7, b.pose7)
    assert a.boxes.shape == (3, 10, 4)
    frac_zero = float((a.depth == 0).mean())
    assert 0.01 < frac_zero < 0.03 and np.isnan(a.depth).any()
    assert np.allclose(np.linalg.norm(a.pose7[:, 3:], axis=1), 1.0)
    bc = a.bbox_coordinates()
    assert list(bc.keys()) == [0, 1, 2] and len(bc[0][0]) == 7
    assert a.frame_off().tolist() == [0, 10, 20, 30]
    assert list(a.pose_dataframe().columns) == ["timestamp", "tx", "ty", "tz", "qx", "qy", "qz", "qw"]


def test_transforms_mirror_agrees_with_oracle():
    from src.utils.transformations import Transforms

    t = Transforms()
    rng = np.random.default_rng(0)
    for _ in range(20):
        pose = rng.normal(size=7)
        np.testing.assert_allclose(t.get_transformation_matrix(pose), ora.get_transformation_matrix(pose), atol=1e-15)
    bbox = [100.5, 200.25, 400.0, 900.0, 1, 0.9, 3]
    assert t.scale_bounding_box(bbox, (1440, 1920), (192, 256)) == ora.scale_bounding_box(bbox, (1440, 1920), (192, 256))
    assert t.bbox_to_3d([1, 2, 3, 4], 640) == [(1.0, 2.0), (1.0, 4.0), (3.0, 4.0), (3.0, 2.0)]
    np.testing.assert_allclose(t._depth_to_3d(3, 4, 2000.0, 10.0, 10.0, 1.0, 2.0, 1000), ora.depth_to_3d(3, 4, 2000.0, 10.0, 10.0, 1.0, 2.0, 1000))
    box8 = t.create_3d_bounding_box([np.array(p, float) for p in [(0, 0, 0), (0, 1, 0), (1, 1, 0), (1, 0, 0)]], 0.03)
    assert len(box8) == 8 and abs(abs(box8[0][2]) - 0.03) < 1e-12


def test_pose_extractor_reads_rtabmap_pose_file(tmp_path):
    from src.mapper.database_query import PoseDataExtractor

    p = tmp_path / "poses.txt"
    p.write_text("#timestamp x y z qx qy qz qw id\n1.5 0.1 0.2 0.3 0 0 0 1 7\n2.5 1.1 1.2 1.3 0 0 1 0 8\n")
    df = PoseDataExtractor(str(p)).fetch_data()
    assert list(df.columns) == ["timestamp", "tx", "ty", "tz", "qx", "qy", "qz", "qw"]
    assert df.iloc[1][1:].to_numpy().astype(float).tolist() == [1.1, 1.2, 1.3, 0, 0, 1, 0]


def test_process_pose_batches_like_the_reference_loop():
    """_gather() walks bbox_coordinates in dict order, takes pose row i for frame i, applies
    the width-ratio intrinsics rescale -- no GPU involved."""
    from lm3d import synth
    from src.mapper.pose_processor import ProcessPose

    seq = synth.make_sequence(4, 32, 24, 3, seed=9)
    bc = seq.bbox_coordinates()
    bc = {2: bc[2], 0: [], 3: bc[3][:1]}  # custom order, an empty frame, a ragged frame
    pp = ProcessPose(seq.pose_dataframe(), seq.dataset(), bc, 640, 24, 32)
    frames, depth, pose7, intr4, image_wh, frame_off, boxes = pp._gather()
    assert frames == [2, 0, 3] and frame_off.tolist() == [0, 3, 3, 4] and boxes.shape == (4, 4)
    assert np.array_equal(depth[0], seq.depth[2], equal_nan=True)
    np.testing.assert_array_equal(pose7[2], seq.pose7[3])
    np.testing.assert_array_equal(intr4, seq.intr4_depth_res()[[2, 0, 3]])
    np.testing.assert_array_equal(boxes[3], seq.boxes[3, 0])


def test_process_pose_gather_frame_by_frame_equals_batched():
    """A dataset without ``batch`` is read with ``dataset[i]`` like the reference; the result is the same."""
    from lm3d import synth
    from src.mapper.pose_processor import ProcessPose

    seq = synth.make_sequence(5, 32, 24, 2, seed=10)
    bc = seq.bbox_coordinates()

    class PlainDataset:  # the reference's ImageDataset interface only
        def __getitem__(self, i):
            return None, seq.depth[i], seq.intrinsics[i]

    a = ProcessPose(seq.pose_dataframe(), seq.dataset(), bc, 640, 24, 32)._gather()
    b = ProcessPose(seq.pose_dataframe(), PlainDataset(), bc, 640, 24, 32)._gather()
    assert a[0] == b[0]
    for x, y in zip(a[1:], b[1:]):
        assert np.array_equal(x, y, equal_nan=True)


def test_lifted_records_round_trip_and_rows(tmp_path):
    """The columnar wire format: save -> load is lossless without pickle; to_rows() gives the
    reference's nested-list form (pose_processor.py:208) with (3,) float64 corners."""
    import pickle

    from lm3d import lift
    from src.mapper.pose_processor import LiftedRecords

    rng = np.random.default_rng(1)
    B = 7
    rec = np.zeros(B, dtype=lift.RECORD_DTYPE)
    rec["corners"] = rng.normal(size=(B, 4, 3)).astype(np.float32)
    rec["n_valid"] = np.arange(B)
    rec["centroid"][2] = np.nan
    lr = LiftedRecords(rec, np.array([0, 3, 3, 7], dtype=np.int64), np.array([5, 2, 9]), np.arange(B) % 2,
                       rng.random(B), np.array(["stop", "yield", "stop", "a", "b", "c", "d"]))
    path = tmp_path / "lifted.npz"
    lr.save(path)
    back = LiftedRecords.load(path)
    assert back.records.tobytes() == rec.tobytes() and back.records.dtype == lift.RECORD_DTYPE
    assert back.frame_off.tolist() == [0, 3, 3, 7] and back.frames.tolist() == [5, 2, 9]
    assert back.label.tolist() == lr.label.tolist() and np.array_equal(back.conf, lr.conf)
    rows = back.to_rows()
    assert list(rows.keys()) == [5, 2, 9] and [len(v) for v in rows.values()] == [3, 0, 4]
    r = rows[9][1]   # box 4
    assert len(r) == 7 and all(isinstance(c, np.ndarray) and c.shape == (3,) and c.dtype == np.float64 for c in r[:4])
    np.testing.assert_array_equal(np.stack(r[:4]), rec["corners"][4].astype(np.float64))
    assert r[4:] == [0, float(lr.conf[4]), "b"]
    pickle.loads(pickle.dumps(rows))


def test_shard_ranges_cover_all_frames():
    from lm3d import dist as ldist

    for F in (1, 7, 100, 1001):
        for world in (1, 2, 3, 8):
            spans = [ldist.shard_range(F, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == F
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
    f0, f1, b0, b1, local = ldist.shard_boxes(np.array([0, 2, 2, 5, 9]), 1, 2)
    assert (f0, f1, b0, b1) == (2, 4, 2, 9) and local.tolist() == [0, 3, 7]


def test_record_dtype_matches_header_struct():
    from lm3d import lift

    assert lift.RECORD_DTYPE.itemsize == 96
    assert lift.RECORD_DTYPE.fields["n_valid"][1] == 88 and lift.RECORD_DTYPE.fields["z_q"][1] == 84


def test_torch_wrappers_reject_cpu_tensors():
    import torch

    from lm3d import lift

    with pytest.raises(ValueError, match="CUDA"):
        lift.lift_boxes(torch.zeros(1, 4, 4), torch.zeros(1, 7, dtype=torch.float64), torch.ones(1, 4, dtype=torch.float64),
                        torch.zeros(1, 4, dtype=torch.int32), torch.tensor([0, 1]))


def test_nms_wrapper_rejects_cpu_tensors_and_bad_shapes():
    import torch

    from lm3d import nms

    with pytest.raises(ValueError, match="CUDA"):
        nms.nms_boxes(torch.zeros(3, 12), torch.zeros(3), torch.zeros(3, dtype=torch.int32))


def test_bbox_processor_shape_without_boxes_and_no_cpu_fallback():
    """Frames without rows come back as empty lists without touching the GPU; with rows and no CUDA device the class
    raises (there is no CPU path behind it)."""
    import torch

    from src.mapper.bbox_optimiser import BoundingBoxProcessor

    assert BoundingBoxProcessor({3: [], 7: []}, None).suppress_bboxes() == {3: [], 7: []}
    if not torch.cuda.is_available():
        row = [np.zeros(3), np.array([0.0, 1.0, 0.0]), np.array([1.0, 1.0, 0.0]), np.array([1.0, 0.0, 0.0]), 0, 0.9, "sign"]
        with pytest.raises(Exception):
            BoundingBoxProcessor({0: [row]}, None).suppress_bboxes()
