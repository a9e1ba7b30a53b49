"""The oracle against the fixture produced by the REFERENCE's own ProcessPose
(tests/golden/make_golden.py).  CPU only."""
import os

import numpy as np

from oracle import reference_numpy as ora

GOLD = os.path.join(os.path.dirname(__file__), "golden", "reference_rows.npz")


def load_case():
    from lm3d import synth

    g = np.load(GOLD)
    seq = synth.make_sequence(int(g["F"]), int(g["H"]), int(g["W"]), int(g["B"]), seed=int(g["seed"]))
    seq.boxes[...] = g["boxes"]
    return g, seq


def test_loop_form_matches_reference_rows():
    g, seq = load_case()
    out = ora.get_global_coordinates_loop(seq.pose7, seq.dataset(), seq.bbox_coordinates(), 192, 256)
    assert list(out.keys()) == list(range(int(g["F"])))
    for f in range(int(g["F"])):
        assert len(out[f]) == int(g["B"])
        for b, row in enumerate(out[f]):
            assert len(row) == 7  # [c0,c1,c2,c3,damage_cls,conf,label]  (pose_processor.py:208)
            got = np.stack(row[:4])
            np.testing.assert_allclose(got, g["corners"][f, b], rtol=0, atol=1e-12)
            np.testing.assert_allclose(np.array(row[4:], dtype=np.float64), g["tail"][f, b])


def test_batched_form_matches_reference_rows():
    g, seq = load_case()
    B = int(g["B"])
    rect4 = ora.boxes_to_rects(seq.boxes.reshape(-1, 4), np.repeat(seq.image_wh(), B, axis=0), (192, 256))
    rec = ora.lift_boxes(seq.depth, seq.pose7, seq.intr4_depth_res(), rect4, seq.frame_off())
    np.testing.assert_allclose(rec["corners"].reshape(g["corners"].shape), g["corners"], rtol=0, atol=1e-12)
    # the fixture's edge boxes really are clamped rects
    assert tuple(rect4[0]) == (0, 0, 191, 255)
    assert tuple(rect4[1 * B + 1])[2:] == (191, 255)
