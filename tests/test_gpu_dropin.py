"""GPU: the drop-in ProcessPose (reference API, CUDA inside) against the reference-generated
fixture and the oracle; the full-frame cloud; a larger C2-shaped run; size-independent
properties at a scale the oracle cannot follow."""
import os
import pickle

import numpy as np
import pytest
import torch

from oracle import reference_numpy as ora
from parity import assert_close_coords, assert_records_match

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "reference_rows.npz")


def test_process_pose_matches_reference_rows(cuda_device):
    from lm3d import synth
    from src.mapper.pose_processor import ProcessPose

    g = np.load(GOLD)
    seq = synth.make_sequence(int(g["F"]), int(g["H"]), int(g["W"]), int(g["B"]), seed=int(g["seed"]))
    seq.boxes[...] = g["boxes"]
    bc = seq.bbox_coordinates()
    pp = ProcessPose(pose=seq.pose_dataframe(), dataset=seq.dataset(), bbox_coordinates=bc, img_size=640,
                     depth_width=192, depth_height=256)
    out = pp.get_global_coordinates()
    assert list(out.keys()) == list(bc.keys())
    for f in out:
        assert len(out[f]) == len(bc[f])
        for b, row in enumerate(out[f]):
            assert len(row) == 7 and all(isinstance(c, np.ndarray) and c.shape == (3,) for c in row[:4])
            assert row[4:] == bc[f][b][-3:]                      # passthrough tail (pose_processor.py:208)
            assert_close_coords(np.stack(row[:4]), g["corners"][f, b], f"frame {f} box {b}")
    pickle.loads(pickle.dumps(out))                              # stage pickle must keep working (task_def.py:62-72)


def test_process_pose_order_empty_and_ragged_frames(cuda_device):
    from lm3d import synth
    from src.mapper.pose_processor import ProcessPose

    seq = synth.make_sequence(5, 256, 192, 4, seed=8)
    full = seq.bbox_coordinates()
    bc = {3: full[3], 1: [], 4: full[4][:2], 0: full[0]}
    out = ProcessPose(seq.pose_dataframe(), seq.dataset(), bc, 640, 192, 256).get_global_coordinates()
    assert list(out.keys()) == [3, 1, 4, 0] and out[1] == [] and len(out[4]) == 2
    want = ora.get_global_coordinates_loop(seq.pose7, seq.dataset(), bc, 192, 256)
    for f in bc:
        for b in range(len(bc[f])):
            assert_close_coords(np.stack(out[f][b][:4]), np.stack(want[f][b][:4]), f"frame {f} box {b}")


def test_columnar_records_equal_rows_and_chunked_gather(cuda_device, tmp_path, monkeypatch):
    """get_global_records() == get_global_coordinates() (same corners, same order), survives save / load, and a
    gather in small frame chunks (bounded host memory) gives byte-identical records."""
    from lm3d import synth
    from src.mapper import pose_processor as ppm

    seq = synth.make_sequence(9, 256, 192, 5, seed=31)
    full = seq.bbox_coordinates()
    bc = {k: full[k] for k in (4, 0, 7, 2, 8, 1)}
    bc[7] = []
    pp = ppm.ProcessPose(seq.pose_dataframe(), seq.dataset(), bc, 640, 192, 256)
    lr = pp.get_global_records()
    rows = pp.get_global_coordinates()
    assert lr.frames.tolist() == [4, 0, 7, 2, 8, 1] and lr.frame_off.tolist() == [0, 5, 10, 10, 15, 20, 25]
    lr.save(tmp_path / "r.npz")
    back = ppm.LiftedRecords.load(tmp_path / "r.npz").to_rows()
    assert list(back.keys()) == list(rows.keys())
    for f in rows:
        assert len(rows[f]) == len(back[f])
        for ra, rb in zip(rows[f], back[f]):
            assert all(np.array_equal(x, y) for x, y in zip(ra[:4], rb[:4])) and ra[4:] == rb[4:]
    monkeypatch.setattr(ppm, "GATHER_CHUNK_BYTES", 2 * 256 * 192 * 4)   # two frames per chunk
    lr2 = ppm.ProcessPose(seq.pose_dataframe(), seq.dataset(), bc, 640, 192, 256).get_global_records()
    assert lr2.records.tobytes() == lr.records.tobytes()
    # frame by frame through the reference's dataset[i] interface: the same bytes again
    class Plain:
        def __getitem__(self, i):
            return None, seq.depth[i], seq.intrinsics[i]

    lr3 = ppm.ProcessPose(seq.pose_dataframe(), Plain(), bc, 640, 192, 256).get_global_records()
    assert lr3.records.tobytes() == lr.records.tobytes()


def test_host_entry_rejects_bad_csr_and_restores_the_device(cuda_device):
    from lm3d import _capi, lift, synth

    seq = synth.make_sequence(3, 256, 192, 2, seed=1)
    args = (seq.depth, seq.pose7, seq.intr4_depth_res(), seq.boxes.reshape(-1, 4), seq.image_wh())
    for bad in ([0, 4, 2, 6], [1, 2, 4, 6], [0, 2, 4, 5]):   # non-monotone, off[0] != 0, off[F] != B
        with pytest.raises(_capi.Lm3dError):
            lift.lift_boxes_host(*args, np.array(bad, dtype=np.int64))
    before = torch.cuda.current_device()
    lift.lift_boxes_host(*args, seq.frame_off(), device=0)
    assert torch.cuda.current_device() == before


def test_frame_cloud_matches_oracle(cuda_device):
    from lm3d import lift, synth

    seq = synth.make_sequence(3, 256, 192, 1, seed=2)
    xyz, n_valid = lift.lift_frame_cloud(torch.from_numpy(seq.depth).to(cuda_device), torch.from_numpy(seq.pose7).to(cuda_device),
                                         torch.from_numpy(seq.intr4_depth_res()).to(cuda_device))
    xyz, n_valid = xyz.cpu().numpy(), n_valid.cpu().numpy()
    for f in range(3):
        fx, fy, cx, cy = seq.intr4_depth_res()[f]
        want = ora.full_frame_cloud(seq.depth[f], ora.get_transformation_matrix(seq.pose7[f]), fx, fy, cx, cy)
        m = ora.valid_mask(seq.depth[f])
        assert int(n_valid[f]) == int(m.sum()) == want.shape[0]
        assert np.array_equal(np.isnan(xyz[f, ..., 0]), ~m)      # validity mask bit-exact
        assert_close_coords(xyz[f][m], want, f"cloud frame {f}")


def test_c2_shape_many_frames(cuda_device):
    from lm3d import synth
    from test_gpu_parity import run_cuda, run_oracle

    seq = synth.make_config("C2", frames=150)                    # 3000 boxes, every warp-box size class
    rec, os_, rect4 = run_cuda(seq, cuda_device)
    assert_records_match(rec, os_, run_oracle(seq, rect4))


def test_full_size_properties(cuda_device):
    """BASELINE config C2 at full size (10k frames, 200k boxes): properties that need no oracle,
    plus an oracle check of a random sample of boxes."""
    from lm3d import lift, synth

    F, H, W, B = synth.CONFIGS["C2"]
    d = synth.make_sequence_torch(F, H, W, B, seed=99, device=cuda_device)
    rect4 = lift.scale_boxes(d["boxes"], d["image_wh"], d["frame_off"], W, H)
    rec_t, os_t = lift.lift_boxes(d["depth"], d["pose7"], d["intr4"], rect4, d["frame_off"], order_stats=True)
    rec = lift.records_to_numpy(rec_t)
    os_ = os_t.cpu().numpy()
    r = rect4.cpu().numpy().astype(np.int64)
    assert np.array_equal(rec["n_pix"], (r[:, 2] - r[:, 0] + 1) * (r[:, 3] - r[:, 1] + 1))
    assert (rec["n_valid"] <= rec["n_pix"]).all() and (rec["n_valid"] > 0).all()
    assert (os_[:, 0] <= os_[:, 1]).all()
    zq = rec["z_q"].astype(np.float64) * 1000.0
    assert ((zq >= os_[:, 0] * (1 - 1e-6)) & (zq <= os_[:, 1] * (1 + 1e-6))).all()     # percentile lies between its order stats
    assert ((rec["aabb_min"] <= rec["centroid"] + 1e-4) & (rec["centroid"] <= rec["aabb_max"] + 1e-4)).all()
    # idempotence / determinism: a second run is byte-identical
    rec2 = lift.records_to_numpy(lift.lift_boxes(d["depth"], d["pose7"], d["intr4"], rect4, d["frame_off"]))
    assert rec2.tobytes() == rec.tobytes()
    # oracle on a random sample of boxes (their frames are pulled back to the host)
    rng = np.random.default_rng(0)
    pick = np.sort(rng.choice(F * B, size=200, replace=False))
    frames = np.unique(pick // B)
    depth_h = d["depth"][torch.from_numpy(frames).to(cuda_device)].cpu().numpy()
    pose_h, intr_h = d["pose7"].cpu().numpy(), d["intr4"].cpu().numpy()
    fmap = {int(f): i for i, f in enumerate(frames)}
    want = np.zeros(len(pick), dtype=ora.ORACLE_RECORD)
    for i, b in enumerate(pick):
        f = int(b // B)
        fx, fy, cx, cy = intr_h[f]
        want[i] = ora.lift_box(depth_h[fmap[f]], tuple(int(v) for v in r[b]), ora.get_transformation_matrix(pose_h[f]), fx, fy, cx, cy)
    assert_records_match(rec[pick], os_[pick], want)
