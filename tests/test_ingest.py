"""Depth / calibration ingest: the oracle against the fixture produced by the REFERENCE's own loaders, the host
table builder against the oracle, and (GPU) the kernel against both, bit for bit."""
import os

import numpy as np
import pytest

from oracle import reference_numpy as ora

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "depth_ingest.npz")


def _gold():
    g = np.load(GOLD)
    calib = {str(k): float(v) for k, v in zip(g["calib_keys"], g["calib_vals"])}
    return g["raw_8uc4"], g["depth_mm"], calib


def test_oracle_decode_matches_reference_loader_bit_for_bit():
    raw, want, _ = _gold()
    got = ora.decode_depth_8uc4(raw)
    assert got.dtype == np.float32 and got.shape == want.shape
    assert got.tobytes() == want.tobytes()  # NaN / inf / denormal / negative pixels included


def test_intrinsics_table_matches_oracle_and_reference_dict():
    from lm3d import ingest

    _, _, calib = _gold()
    assert calib == {"image_width": 1440.0, "image_height": 1920.0, "fx": 1450.25, "fy": 1449.75, "cx": 721.5, "cy": 958.25}
    tab = ingest.intrinsics_table([calib, calib], 192)
    assert np.array_equal(tab[0], ora.intrinsics_row(calib, 192)) and np.array_equal(tab[0], tab[1])
    assert np.array_equal(tab[0], np.array([1450.25, 1449.75, 721.5, 958.25]) / (1440.0 / 192))


def test_decode_rejects_cpu_tensors():
    import torch
    from lm3d import ingest

    with pytest.raises(ValueError):
        ingest.decode_depth(torch.zeros((2, 2, 4), dtype=torch.uint8))


@pytest.mark.gpu
def test_gpu_decode_matches_reference_loader_bit_for_bit(cuda_device):
    import torch
    from lm3d import ingest

    raw, want, _ = _gold()
    got = ingest.decode_depth(torch.from_numpy(raw).to(cuda_device)).cpu().numpy()
    # bit-exact wherever the value is a number (denormals, infinities, negatives included); a NaN stays a NaN
    # (the GPU multiplier returns the canonical quiet NaN, numpy keeps the payload -- both are "invalid depth")
    nan = np.isnan(want)
    assert np.array_equal(np.isnan(got), nan)
    assert np.array_equal(got.view(np.uint32)[~nan], want.view(np.uint32)[~nan])
    # in place, odd pixel counts, other scales
    rng = np.random.default_rng(1)
    for n in (1, 3, 4, 5, 1027, 49152 * 3 + 2):
        m = (rng.random(n, dtype=np.float32) * 5).astype(np.float32)
        r = torch.from_numpy(m.view(np.uint8).reshape(n, 4).copy()).to(cuda_device)
        o = ingest.decode_depth(r, scale=1000.0, out=r.view(torch.float32).reshape(n))
        torch.cuda.synchronize()
        assert o.cpu().numpy().tobytes() == (m * np.float32(1000.0)).tobytes()


@pytest.mark.gpu
def test_gpu_ingest_then_lift_matches_oracle(cuda_device):
    """Decoded PNG bytes -> lm3d_ingest_depth -> lm3d_lift_boxes equals the oracle run on the reference-decoded depth."""
    import torch
    from lm3d import ingest, lift, synth
    from parity import assert_records_match

    raw, want_mm, calib = _gold()
    F, H, W = want_mm.shape
    rng = np.random.default_rng(3)
    pose7 = synth.make_poses(F, rng)
    intr4 = ingest.intrinsics_table([calib] * F, W)
    rect4 = np.array([[2, 3, 40, 50], [0, 0, W - 1, H - 1], [10, 10, 10, 10]] * F, dtype=np.int32)
    frame_off = np.arange(F + 1, dtype=np.int64) * 3
    dev = cuda_device
    depth = ingest.decode_depth(torch.from_numpy(raw).to(dev))
    rec, os_ = lift.lift_boxes(depth, torch.from_numpy(pose7).to(dev), torch.from_numpy(intr4).to(dev),
                               torch.from_numpy(rect4).to(dev), torch.from_numpy(frame_off).to(dev), order_stats=True)
    torch.cuda.synchronize()
    want = ora.lift_boxes(want_mm, pose7, intr4, rect4, frame_off)
    assert_records_match(lift.records_to_numpy(rec), os_.cpu().numpy(), want)


# ---------------------------------------------------------------------------------------
# batched loader (SURVEY 8f row 2 as written): PNG + YAML files -> [F,H,W] on the device
# ---------------------------------------------------------------------------------------
def _write_scan(tmp_path, raw, yaml_text, names=None):
    """Files as the reference's extraction step leaves them (detector/database_query.py:28-42): N.png depth in 8UC4,
    N.yaml calibration, N.jpg colour (empty here: the lift never reads it)."""
    import cv2

    for d in ("rgb", "depth", "calib"):
        (tmp_path / d).mkdir()
    names = names or [str(i + 1) for i in range(len(raw))]
    for n, img in zip(names, raw):
        assert cv2.imwrite(str(tmp_path / "depth" / f"{n}.png"), img)
        (tmp_path / "rgb" / f"{n}.jpg").write_bytes(b"")
        (tmp_path / "calib" / f"{n}.yaml").write_text(yaml_text)
    return str(tmp_path / "depth"), str(tmp_path / "calib"), str(tmp_path / "rgb")


def test_depth_sequence_pairs_files_in_natural_order_and_parses_calibration(tmp_path):
    from lm3d import ingest

    g = np.load(GOLD)
    raw = g["raw_8uc4"]
    names = ["10", "2", "1"]   # natural order: 1, 2, 10 (dataset.py:32-33 uses natsorted)
    depth_dir, calib_dir, rgb_dir = _write_scan(tmp_path, raw, str(g["yaml_text"]), names)
    (tmp_path / "depth" / "99.png").write_bytes(b"x")   # a depth file without its jpg is not a frame (dataset.py:39-49)
    seq = ingest.DepthSequence.from_dirs(depth_dir, calib_dir, image_dir=rgb_dir, depth_width=48, depth_height=64)
    assert len(seq) == 3 and [os.path.basename(p) for p in seq.depth_paths] == ["1.png", "2.png", "10.png"]
    cal = seq.calibration([0, 2])
    _, _, calib = _gold()
    assert cal.shape == (2, 6)
    assert cal[0].tolist() == [calib["fx"], calib["fy"], calib["cx"], calib["cy"], calib["image_width"], calib["image_height"]]
    import torch

    if not torch.cuda.is_available():   # no CPU conversion path behind the loader
        with pytest.raises(Exception):
            seq.batch_device([0, 1])


@pytest.mark.gpu
def test_depth_sequence_matches_reference_loader_and_feeds_process_pose(cuda_device, tmp_path):
    """PNG / YAML files -> DepthSequence.batch_device == the reference's own _load_depth_image output (fixture), for
    any frame order and slot size; ProcessPose fed by the loader == ProcessPose fed by the decoded arrays."""
    import torch
    from lm3d import ingest, synth
    from src.mapper.pose_processor import ProcessPose

    g = np.load(GOLD)
    raw, want = g["raw_8uc4"], g["depth_mm"]
    F, H, W = want.shape
    depth_dir, calib_dir, rgb_dir = _write_scan(tmp_path, raw, str(g["yaml_text"]))
    for slot in (1, 2, 256):
        seq = ingest.DepthSequence.from_dirs(depth_dir, calib_dir, image_dir=rgb_dir, depth_width=W, depth_height=H,
                                             slot_frames=slot, workers=3)
        order = [2, 0, 1, 0]
        got, cal = seq.batch_device(order)
        torch.cuda.synchronize()
        got = got.cpu().numpy()
        nan = np.isnan(want[order])
        assert np.array_equal(np.isnan(got), nan)
        assert np.array_equal(got.view(np.uint32)[~nan], want[order].view(np.uint32)[~nan])
        assert cal.shape == (4, 6) and cal[0, 4] == 1440.0
    # the drop-in on top of it
    rng = np.random.default_rng(5)
    pose = synth.Sequence(depth=want, pose7=synth.make_poses(F, rng), intrinsics=[], boxes=np.zeros((F, 1, 4)),
                          damage_cls=None, conf=None, label=None, depth_width=W, depth_height=H).pose_dataframe()
    bc = {0: [[100.0, 200.0, 900.0, 1500.0, 0, 0.9, "stop"], [0.0, 0.0, 1440.0, 1920.0, 1, 0.8, "yield"]], 2: [], 1: [[30.0, 40.0, 700.0, 800.0, 0, 0.7, "stop"]]}
    _, _, calib = _gold()
    a = ProcessPose(pose, seq, bc, 640, W, H).get_global_records()
    b = ProcessPose(pose, synth.ArrayDataset(want, [calib] * F), bc, 640, W, H).get_global_records()
    assert a.records.tobytes() == b.records.tobytes() and a.label.tolist() == ["stop", "yield", "stop"]
