"""Writes tests/golden/nms_cases.npz: inputs and oracle outputs (oracle/nms_numpy.py, NMS-SPEC v0) of two seeded
cases.  The reference has no implementation of this stage in its repository (src/mapper/bbox_optimiser.py is
absent), so unlike reference_rows.npz this fixture pins the SPEC as frozen in round 1, not the reference: a change
of the oracle or of the CUDA path that alters a keep flag or a parent shows up against it.

    python tests/golden/make_golden_nms.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from nms_cases import chain_boxes, clustered_boxes  # noqa: E402
from oracle import nms_numpy as nms  # noqa: E402

out = {}
for name, (corners, conf, label), thr, pad in (
    ("clustered", clustered_boxes(60, 20, seed=101), 0.1, 0.03),
    ("loose", clustered_boxes(25, 30, seed=102, jitter=0.12, label_noise=0.2), 0.3, 0.05),
    ("chain", chain_boxes(41), 0.1, 0.03),
):
    corners = corners.copy()
    if name == "clustered":
        corners[::97] = np.nan
        conf = np.round(conf, 2)
    keep, parent = nms.nms_3d(corners, conf, label, thr, pad)
    out.update({f"{name}_corners": corners, f"{name}_conf": conf, f"{name}_label": label, f"{name}_thr": np.float32(thr),
                f"{name}_pad": np.float32(pad), f"{name}_keep": keep, f"{name}_parent": parent})
np.savez_compressed(os.path.join(HERE, "nms_cases.npz"), **out)
print({k: int(v.sum()) for k, v in out.items() if k.endswith("_keep")})
