"""dev helper: cycles per phase of tile_box_kernel (a -DLM3D_TILE_TIMING build: make variant NAME=tt EXTRA=-DLM3D_TILE_TIMING;
LM3D_LIB=.../liblm3d_tt.so python tools/tile_phases.py [C3 frames] [C5 frames])"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "3d-localisation-and-mapping_b200"))
import torch
from lm3d import _capi, lift, synth
dev = torch.device("cuda:0")
lib = _capi.load()
NAMES = ["setup", "bracket", "-", "tiles+strips", "passA", "narrow", "-", "passB", "select+write", "claim", "sum items"]
for name, F in (("C3", int(sys.argv[1]) if len(sys.argv) > 1 else 128), ("C5", int(sys.argv[2]) if len(sys.argv) > 2 else 64)):
    _, H, W, B = synth.CONFIGS[name]
    d = synth.make_sequence_torch(F, H, W, B, seed=1234 + int(name[1:]), device=dev)
    rect4 = lift.scale_boxes(d["boxes"], d["image_wh"], d["frame_off"], W, H)
    plan = lift.LiftPlan(F, F * B, dev, H=H, W=W)
    for _ in range(2):
        lift.lift_boxes(d["depth"], d["pose7"], d["intr4"], rect4, d["frame_off"], plan=plan)
    torch.cuda.synchronize()
    out = (ctypes.c_ulonglong * 16)()
    lib.lm3d_debug_tile_prof(out, 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    lift.lift_boxes(d["depth"], d["pose7"], d["intr4"], rect4, d["frame_off"], plan=plan)
    e1.record()
    torch.cuda.synchronize()
    lib.lm3d_debug_tile_prof(out, 0)
    tot = sum(out[:11])
    print(f"{name} F={F} boxes={F*B} lift_ms={e0.elapsed_time(e1):.3f} cycles/box={tot/(F*B):.0f}")
    for i, n in enumerate(NAMES):
        print(f"  {n:12s} {100.0*out[i]/tot:5.1f} %   {out[i]/(F*B):9.0f} cycles/box")
    del d, plan
