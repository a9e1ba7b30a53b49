"""dev experiment: 200k C2-law boxes over F frames (F small => depth is L2-resident); prints lift_small ms."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "3d-localisation-and-mapping_b200"))
import numpy as np, torch
from lm3d import _capi, lift, synth
dev = torch.device("cuda:0")
lib = _capi.load()
for F in (100, 400, 10000):
    per = 200000 // F
    d = synth.make_sequence_torch(F, 256, 192, 20, seed=1236, device=dev)
    rng = np.random.default_rng(0)
    boxes = torch.from_numpy(synth.make_boxes(F, per, rng)).to(dev).reshape(F * per, 4).contiguous()
    frame_off = torch.arange(F + 1, dtype=torch.int64, device=dev) * per
    rect4 = lift.scale_boxes(boxes, d["image_wh"], frame_off, 192, 256)
    plan = lift.LiftPlan(F, F * per, dev)
    for _ in range(3):
        lift.lift_boxes(d["depth"], d["pose7"], d["intr4"], rect4, frame_off, plan=plan)
    lib.lm3d_profile_enable(1)
    ms4 = (ctypes.c_float * 4)(); acc = 0.0
    for _ in range(10):
        lift.lift_boxes(d["depth"], d["pose7"], d["intr4"], rect4, frame_off, plan=plan)
        lib.lm3d_profile_read(ms4); acc += ms4[2]
    lib.lm3d_profile_enable(0)
    print(f"F={F} boxes={F*per} depth_MB={F*256*192*4/1e6:.0f} lift_small_ms={acc/10:.3f}", flush=True)
    del d, plan
