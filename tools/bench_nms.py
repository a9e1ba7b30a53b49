"""Times lm3d_nms_boxes (device-resident inputs, CUDA events around the C-ABI call) on clustered boxes:
`signs` physical signs seen `per_sign` times each.  Usage: python tools/bench_nms.py [signs per_sign] ..."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "3d-localisation-and-mapping_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
from lm3d import nms
from nms_cases import clustered_boxes

dev = torch.device("cuda:0")
cases = [(4000, 50), (400, 500), (40000, 50)] if len(sys.argv) < 3 else [tuple(map(int, sys.argv[i:i + 2])) for i in range(1, len(sys.argv) - 1, 2)]
for signs, per in cases:
    corners, conf, label = clustered_boxes(signs, per, seed=1, extent=40.0 * (signs / 4000) ** (1 / 3))
    c, f, l = (torch.from_numpy(corners.reshape(-1, 12)).to(dev), torch.from_numpy(conf).to(dev), torch.from_numpy(label).to(dev))
    for _ in range(3):
        keep, parent, rounds = nms.nms_boxes(c, f, l)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    torch.cuda.synchronize(); ev[0].record()
    reps = 10
    for _ in range(reps):
        keep, parent, rounds = nms.nms_boxes(c, f, l)
    ev[1].record(); torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / reps
    B = len(conf)
    print(f"{B:8d} boxes ({signs} signs x {per}): {ms:8.3f} ms per call, {B / ms / 1e3:8.2f} M boxes/s, kept {int(keep.sum())}, rounds {rounds}", flush=True)
