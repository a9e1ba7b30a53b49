"""GPU: the CUDA 3-D NMS (lm3d_nms_boxes, through the C ABI) against the oracle: keep flags and parents bit-exact."""
import os

import numpy as np
import pytest
import torch

from nms_cases import chain_boxes, clustered_boxes
from oracle import nms_numpy as ora

pytestmark = pytest.mark.gpu


def run_cuda(corners, conf, label, dev, thr=0.1, pad=0.03):
    from lm3d import nms

    keep, parent, rounds = nms.nms_boxes(torch.from_numpy(np.ascontiguousarray(corners.reshape(len(conf), -1))).to(dev),
                                         torch.from_numpy(conf).to(dev), torch.from_numpy(label).to(dev), thr, pad)
    return keep.cpu().numpy(), parent.cpu().numpy(), rounds


@pytest.mark.parametrize("signs,per_sign,seed", [(1, 1, 0), (30, 20, 1), (200, 25, 2), (5, 600, 3)])
def test_clustered_boxes_match_oracle(cuda_device, signs, per_sign, seed):
    corners, conf, label = clustered_boxes(signs, per_sign, seed=seed)
    keep, parent, _ = run_cuda(corners, conf, label, cuda_device)
    want_keep, want_parent = ora.nms_3d(corners, conf, label)
    assert np.array_equal(keep, want_keep)
    assert np.array_equal(parent, want_parent)


@pytest.mark.parametrize("thr,pad", [(0.0, 0.03), (0.5, 0.1), (0.1, 0.0)])
def test_threshold_and_padding(cuda_device, thr, pad):
    corners, conf, label = clustered_boxes(60, 15, seed=11, jitter=0.1)
    keep, parent, _ = run_cuda(corners, conf, label, cuda_device, thr, pad)
    want_keep, want_parent = ora.nms_3d(corners, conf, label, thr, pad)
    assert np.array_equal(keep, want_keep) and np.array_equal(parent, want_parent)


def test_long_dependency_chain(cuda_device):
    """Every box depends on its predecessor: the relaxation needs about one round per box, many batches of rounds
    and the reuse of the header's round slots."""
    n = 601  # (rounds ~ n when nothing propagates inside a round; the margin over 64 allows for a lot of propagation)
    corners, conf, label = chain_boxes(n)
    keep, parent, rounds = run_cuda(corners, conf, label, cuda_device)
    assert rounds > 64
    assert keep.tolist() == [1 - (i & 1) for i in range(n)]
    assert parent.tolist() == [i - (i & 1) for i in range(n)]


def test_ties_invalid_and_duplicates(cuda_device):
    corners, conf, label = clustered_boxes(8, 40, seed=5)
    conf[:] = np.round(conf, 1)                     # heavy confidence ties: index order decides
    corners[::7] = np.nan                           # lift records with n_valid == 0
    corners[3] = corners[10]; label[3] = label[10]  # exact duplicates
    conf[20] = np.nan                               # N2: a non-finite confidence takes a box out, like a NaN corner
    conf[21] = np.inf
    keep, parent, _ = run_cuda(corners, conf, label, cuda_device)
    want_keep, want_parent = ora.nms_3d(corners, conf, label)
    assert np.array_equal(keep, want_keep) and np.array_equal(parent, want_parent)
    assert (parent[::7] == -1).all() and parent[20] == -1 and parent[21] == -1


def test_empty_and_all_invalid(cuda_device):
    from lm3d import nms

    z = torch.zeros((0, 12), dtype=torch.float32, device=cuda_device)
    keep, parent, rounds = nms.nms_boxes(z, torch.zeros(0, device=cuda_device), torch.zeros(0, dtype=torch.int32, device=cuda_device))
    assert keep.numel() == 0 and parent.numel() == 0 and rounds == 0
    corners = np.full((50, 4, 3), np.nan, dtype=np.float32)
    keep, parent, _ = run_cuda(corners, np.ones(50, np.float32), np.zeros(50, np.int32), cuda_device)
    assert not keep.any() and (parent == -1).all()


def test_records_of_a_lifted_sequence_and_the_drop_in_class(cuda_device):
    """The [B,24] records the lift produces go straight in (stride 24), and the reference-shaped class gives the
    rows the oracle's row-level form gives."""
    from lm3d import lift, nms, synth
    from src.mapper.bbox_optimiser import BoundingBoxProcessor
    from src.mapper.pose_processor import ProcessPose

    seq = synth.make_sequence(40, 256, 192, 12, seed=33)
    seq.pose7[:, :3] *= 0.05                        # a camera that barely moves: boxes of nearby frames overlap in 3-D
    dev = cuda_device
    fo = torch.from_numpy(seq.frame_off()).to(dev)
    rect4 = lift.scale_boxes(torch.from_numpy(seq.boxes.reshape(-1, 4)).to(dev), torch.from_numpy(seq.image_wh()).to(dev), fo,
                             seq.depth_width, seq.depth_height)
    rec = lift.lift_boxes(torch.from_numpy(seq.depth).to(dev), torch.from_numpy(seq.pose7).to(dev),
                          torch.from_numpy(seq.intr4_depth_res()).to(dev), rect4, fo)
    B = rec.shape[0]
    rng = np.random.default_rng(3)
    conf = rng.uniform(0.2, 0.99, B).astype(np.float32)
    label = rng.integers(0, 3, B).astype(np.int32)
    keep, parent, _ = nms.nms_boxes(rec, torch.from_numpy(conf).to(dev), torch.from_numpy(label).to(dev), 0.05, 0.05)
    corners = rec[:, :12].cpu().numpy().reshape(B, 4, 3)
    want_keep, want_parent = ora.nms_3d(corners, conf, label, 0.05, 0.05)
    assert np.array_equal(keep.cpu().numpy(), want_keep) and np.array_equal(parent.cpu().numpy(), want_parent)
    assert 0 < want_keep.sum() < B                  # the case really suppresses something

    pp = ProcessPose(pose=seq.pose_dataframe(), dataset=seq.dataset(), bbox_coordinates=seq.bbox_coordinates(), img_size=640,
                     depth_width=192, depth_height=256)
    rows = pp.get_global_coordinates()
    proc = BoundingBoxProcessor(rows, seq.pose_dataframe())
    out = proc.suppress_bboxes()
    want = ora.suppress_rows(rows)
    assert list(out.keys()) == list(rows.keys())
    for f in rows:
        assert [id(r) for r in out[f]] == [id(r) for r in want[f]]


def test_properties_at_scale(cuda_device):
    """200 k boxes (the C2 box count): size-independent properties instead of the O(B * kept) oracle."""
    from lm3d import nms

    corners, conf, label = clustered_boxes(4000, 50, seed=9)
    dev = cuda_device
    c_t, f_t, l_t = torch.from_numpy(corners.reshape(-1, 12)).to(dev), torch.from_numpy(conf).to(dev), torch.from_numpy(label).to(dev)
    keep, parent, rounds = nms.nms_boxes(c_t, f_t, l_t)
    keep, parent = keep.cpu().numpy().astype(bool), parent.cpu().numpy()
    assert 4000 <= keep.sum() < 0.2 * len(keep) and rounds < 64
    assert np.array_equal(parent[keep], np.flatnonzero(keep))
    sup = ~keep
    p = parent[sup]
    assert keep[p].all() and (label[p] == label[sup]).all()
    assert ((conf[p] > conf[sup]) | ((conf[p] == conf[sup]) & (p < np.flatnonzero(sup)))).all()
    lo, hi, vol, _ = ora.box_extents(corners, 0.03)
    assert all(ora.overlaps(lo[i], hi[i], vol[i], lo[j][None], hi[j][None], vol[j][None], 0.1)[0]
               for i, j in zip(np.flatnonzero(sup)[:2000], p[:2000]))
    k_idx = np.flatnonzero(keep)                    # idempotence: the survivors do not suppress each other
    keep2, _, _ = nms.nms_boxes(c_t[torch.from_numpy(k_idx).to(dev)].contiguous(), f_t[torch.from_numpy(k_idx).to(dev)].contiguous(),
                                l_t[torch.from_numpy(k_idx).to(dev)].contiguous())
    assert bool(keep2.all())


def test_committed_fixture(cuda_device):
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "nms_cases.npz"))
    for name in ("clustered", "loose", "chain"):
        keep, parent, _ = run_cuda(g[f"{name}_corners"], g[f"{name}_conf"], g[f"{name}_label"], cuda_device,
                                   float(g[f"{name}_thr"]), float(g[f"{name}_pad"]))
        assert np.array_equal(keep, g[f"{name}_keep"]) and np.array_equal(parent, g[f"{name}_parent"]), name
