"""dev helper: the warp kernel on quantised depth (heavy ties): time per lift, boxes that took the exact select
(deferred to lift_resolve_kernel on the quad path) and histogram passes beyond the first."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "3d-localisation-and-mapping_b200"))
from lm3d import lift, synth
dev = torch.device("cuda:0")
seq = synth.make_config("C2", frames=1500)
for step in (0.0, 1.0, 10.0, 100.0):
    d = seq.depth.copy()
    if step > 0:
        q = np.round(d / step) * step
        d = np.where(np.isfinite(d) & (d > 0), q, d).astype(np.float32)
    depth = torch.from_numpy(d).to(dev)
    fo = torch.from_numpy(seq.frame_off()).to(dev)
    rect4 = lift.scale_boxes(torch.from_numpy(seq.boxes.reshape(-1, 4)).to(dev), torch.from_numpy(seq.image_wh()).to(dev), fo, 192, 256)
    pose7 = torch.from_numpy(seq.pose7).to(dev); intr4 = torch.from_numpy(seq.intr4_depth_res()).to(dev)
    plan = lift.LiftPlan(depth.shape[0], rect4.shape[0], dev)
    for _ in range(3):
        lift.lift_boxes(depth, pose7, intr4, rect4, fo, plan=plan)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    torch.cuda.synchronize(); ev[0].record()
    for _ in range(10):
        lift.lift_boxes(depth, pose7, intr4, rect4, fo, plan=plan)
    ev[1].record(); torch.cuda.synchronize()
    rare = [int(v) for v in plan.workspace[:64].view(torch.int32)[4:7].cpu()]
    print(f"quantisation {step:6.1f} mm: {ev[0].elapsed_time(ev[1]) / 10:.3f} ms per lift of {rect4.shape[0]} boxes; "
          f"exact_selects={rare[0]} extra_histogram_passes={rare[1]}", flush=True)
