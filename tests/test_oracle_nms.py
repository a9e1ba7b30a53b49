"""CPU: the 3-D NMS oracle (oracle/nms_numpy.py, NMS-SPEC v0) against known answers, an independent scalar
restatement, and its own properties.  (Parity unpinned: the reference's bbox_optimiser.py is absent.)"""
import os

import numpy as np

from nms_cases import brute_force_nms, chain_boxes, clustered_boxes
from oracle import nms_numpy as nms


def square(x, y=0.0, z=0.0, s=1.0):
    return np.array([[x, y, z], [x, y + s, z], [x + s, y + s, z], [x + s, y, z]], dtype=np.float32)


def test_known_answers():
    c = np.stack([square(0), square(0.05), square(5.0)])
    keep, parent = nms.nms_3d(c, [0.5, 0.9, 0.2], [0, 0, 0])
    assert keep.tolist() == [0, 1, 1] and parent.tolist() == [1, 1, 2]
    keep, _ = nms.nms_3d(c, [0.5, 0.9, 0.2], [0, 1, 0])          # labels differ: no competition
    assert keep.tolist() == [1, 1, 1]
    keep, parent = nms.nms_3d(c, [0.7, 0.7, 0.7], [0, 0, 0])     # equal confidence: the lower index wins
    assert keep.tolist() == [1, 0, 1] and parent.tolist() == [0, 0, 2]


def test_chain_keeps_every_other_box():
    corners, conf, label = chain_boxes(9)
    keep, parent = nms.nms_3d(corners, conf, label, thr=0.1, pad=0.03)
    assert keep.tolist() == [1, 0, 1, 0, 1, 0, 1, 0, 1]
    assert parent.tolist() == [0, 0, 2, 2, 4, 4, 6, 6, 8]


def test_invalid_boxes_do_not_take_part():
    c = np.stack([square(0), square(0), square(0)])
    c[0, 2, 1] = np.nan
    keep, parent = nms.nms_3d(c, [0.9, 0.5, 0.4], [0, 0, 0])
    assert keep.tolist() == [0, 1, 0] and parent.tolist() == [-1, 1, 1]


def test_non_finite_confidence_does_not_take_part():
    """N2 covers the confidence too (the CUDA kernel, the header and the oracle agree): NaN / inf confidences neither
    suppress nor get suppressed."""
    c = np.stack([square(0), square(0), square(0), square(0)])
    keep, parent = nms.nms_3d(c, [np.nan, 0.5, np.inf, 0.4], [0, 0, 0, 0])
    assert keep.tolist() == [0, 1, 0, 0] and parent.tolist() == [-1, 1, -1, 1]
    want = brute_force_nms(c, np.array([np.nan, 0.5, np.inf, 0.4], np.float32), np.zeros(4, np.int32), 0.1, 0.03)
    assert np.array_equal(keep, want[0]) and np.array_equal(parent, want[1])


def test_iou_threshold_and_padding():
    # two unit squares shifted by 0.5 along x, pad 0.5: extents 2 x 2 x 1, inter 1.5 x 2 x 1 = 3, union 5 -> IoU 0.6
    c = np.stack([square(0), square(0.5)])
    assert nms.nms_3d(c, [0.9, 0.8], [0, 0], thr=0.59, pad=0.5)[0].tolist() == [1, 0]
    assert nms.nms_3d(c, [0.9, 0.8], [0, 0], thr=0.61, pad=0.5)[0].tolist() == [1, 1]
    assert nms.nms_3d(c, [0.9, 0.8], [0, 0], thr=0.0, pad=0.0)[0].tolist() == [1, 1]   # flat boxes have no volume


def test_matches_scalar_restatement():
    for seed in range(4):
        corners, conf, label = clustered_boxes(12, 9, seed=seed)
        corners[seed] = np.nan
        conf[5] = conf[6]
        want = brute_force_nms(corners, conf, label, 0.1, 0.03)
        got = nms.nms_3d(corners, conf, label, 0.1, 0.03)
        assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])


def test_properties():
    corners, conf, label = clustered_boxes(40, 25, seed=7)
    keep, parent = nms.nms_3d(corners, conf, label)
    k = keep.astype(bool)
    assert 40 <= k.sum() < 0.2 * len(keep)                      # about one survivor per sign (+ label noise)
    keep2, _ = nms.nms_3d(corners[k], conf[k], label[k])        # idempotent
    assert keep2.all()
    assert np.array_equal(parent[k], np.flatnonzero(k))
    sup = ~k
    assert k[parent[sup]].all() and (label[parent[sup]] == label[sup]).all() and (conf[parent[sup]] >= conf[sup]).all()
    perm = np.random.default_rng(0).permutation(len(keep))      # input order only matters through ties
    keep_p, _ = nms.nms_3d(corners[perm], conf[perm], label[perm])
    assert np.array_equal(keep_p, keep[perm])


def test_suppress_rows_shape():
    corners, conf, label = clustered_boxes(3, 4, seed=1)
    rows = [[*[c.astype(np.float64) for c in corners[i]], 0, float(conf[i]), f"sign{label[i]}"] for i in range(12)]
    data = {10: rows[:5], 11: [], 12: rows[5:]}
    out = nms.suppress_rows(data)
    assert list(out.keys()) == [10, 11, 12] and out[11] == []
    keep, _ = nms.nms_3d(corners, conf, label)
    assert sum(len(v) for v in out.values()) == int(keep.sum())
    kept_ids = [id(r) for v in out.values() for r in v]
    assert kept_ids == [id(rows[i]) for i in np.flatnonzero(keep)]


GOLD = os.path.join(os.path.dirname(__file__), "golden", "nms_cases.npz")


def test_oracle_reproduces_the_committed_fixture():
    g = np.load(GOLD)
    for name in ("clustered", "loose", "chain"):
        keep, parent = nms.nms_3d(g[f"{name}_corners"], g[f"{name}_conf"], g[f"{name}_label"], float(g[f"{name}_thr"]),
                                  float(g[f"{name}_pad"]))
        assert np.array_equal(keep, g[f"{name}_keep"]) and np.array_equal(parent, g[f"{name}_parent"]), name
