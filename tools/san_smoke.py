"""dev helper: one small lift per kernel family for compute-sanitizer (run: compute-sanitizer --tool memcheck python tools/san_smoke.py)."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "3d-localisation-and-mapping_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
from lm3d import lift, synth
from oracle import reference_numpy as ora
from parity import assert_records_match

dev = torch.device("cuda:0")
for name, frames in (("C1", 6), ("C2", 4), ("C3", 1)):
    seq = synth.make_config(name, frames=frames)
    fo = torch.from_numpy(seq.frame_off()).to(dev)
    rect4 = lift.scale_boxes(torch.from_numpy(seq.boxes.reshape(-1, 4)).to(dev), torch.from_numpy(seq.image_wh()).to(dev), fo,
                             seq.depth_width, seq.depth_height)
    rec, os_ = lift.lift_boxes(torch.from_numpy(seq.depth).to(dev), torch.from_numpy(seq.pose7).to(dev),
                               torch.from_numpy(seq.intr4_depth_res()).to(dev), rect4, fo, order_stats=True)
    torch.cuda.synchronize()
    want = ora.lift_boxes(seq.depth, seq.pose7, seq.intr4_depth_res(), rect4.cpu().numpy(), seq.frame_off())
    assert_records_match(lift.records_to_numpy(rec), os_.cpu().numpy(), want)
    print(name, "ok", rect4.shape[0], "boxes", flush=True)
