"""Dev helper: run the lift on an F-frame C2-law sequence with the bounds-checked debug
library (LM3D_LIB=.../liblm3d_dbg.so) and print what it recorded."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "3d-localisation-and-mapping_b200"))
import torch

from lm3d import _capi, lift, synth

F = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
dev = torch.device("cuda:0")
d = synth.make_sequence_torch(F, 256, 192, 20, seed=1236, device=dev)
rect4 = lift.scale_boxes(d["boxes"], d["image_wh"], d["frame_off"], 192, 256)
torch.cuda.synchronize()
print("scale ok", flush=True)
try:
    rec = lift.lift_boxes(d["depth"], d["pose7"], d["intr4"], rect4, d["frame_off"])
    torch.cuda.synchronize()
    print("lift ok", flush=True)
except Exception as e:
    print("lift failed:", str(e).split("\n")[0], flush=True)
lib = _capi.load()
if hasattr(lib, "lm3d_debug_read"):
    buf = (ctypes.c_int * 16)()
    st = lib.lm3d_debug_read(buf)
    print("debug_read status", st, list(buf), flush=True)
