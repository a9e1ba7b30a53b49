"""dev helper: why tile_box_kernel hands boxes over (a -DLM3D_DEBUG_REASONS build: make variant NAME=dr EXTRA=-DLM3D_DEBUG_REASONS;
LM3D_LIB=.../liblm3d_dr.so python tools/tile_reasons.py [C3 frames] [C5 frames])"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "3d-localisation-and-mapping_b200"))
import torch
from lm3d import lift, synth
dev = torch.device("cuda:0")
for name, F in (("C3", int(sys.argv[1]) if len(sys.argv) > 1 else 64), ("C5", int(sys.argv[2]) if len(sys.argv) > 2 else 16)):
    _, H, W, B = synth.CONFIGS[name]
    d = synth.make_sequence_torch(F, H, W, B, seed=1234 + int(name[1:]), device=dev)
    rect4 = lift.scale_boxes(d["boxes"], d["image_wh"], d["frame_off"], W, H)
    plan = lift.LiftPlan(F, F * B, dev, H=H, W=W)
    lift.lift_boxes(d["depth"], d["pose7"], d["intr4"], rect4, d["frame_off"], plan=plan)
    torch.cuda.synchronize()
    c = plan.workspace[:128].view(torch.int32).cpu().numpy()
    n = max(1, int(c[16:21].sum()))
    print(f"{name} F={F} boxes={F*B} handed over {c[14]}  reasons ok/column/miss/ties/strips = {list(c[16:21])}  "
          f"mean collected keys {c[22]/n:.0f}, captured strip keys {c[21]/n:.0f}, listed tiles {c[23]/n:.0f}")
