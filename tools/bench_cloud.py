"""dev helper: lm3d_lift_frame_cloud streaming bandwidth (4 B in, 12 B out per pixel) on 2 000 frames of 256x192
(393 MB in, 1.18 GB out: larger than L2) and on 40 frames of 1920x1440.  The C ABI is called with a preallocated
output (what a pipeline does), CUDA events around each call; prints the median and the spread."""
import os, sys, json
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "3d-localisation-and-mapping_b200"))
from lm3d import _capi, synth
from lm3d.lift import _stream_ptr
lib = _capi.load()
dev = torch.device("cuda:0")
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
for F, H, W in ((2000, 256, 192), (40, 1920, 1440)):
    data = synth.make_sequence_torch(F, H, W, 2, seed=5, device=dev)
    depth, pose7, intr4 = data["depth"], data["pose7"], data["intr4"]
    xyz = torch.empty((F, H, W, 3), dtype=torch.float32, device=dev)
    nv = torch.empty((F,), dtype=torch.int32, device=dev)
    ws = torch.empty((int(lib.lm3d_cloud_workspace_bytes(F)),), dtype=torch.uint8, device=dev)
    def call():
        _capi.check(lib.lm3d_lift_frame_cloud(depth.data_ptr(), F, H, W, pose7.data_ptr(), intr4.data_ptr(), 1000.0, float("inf"),
                                              xyz.data_ptr(), nv.data_ptr(), ws.data_ptr(), ws.numel(), _stream_ptr(dev)), "lm3d_lift_frame_cloud")
    for _ in range(10):
        call()
    reps = 40
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    torch.cuda.synchronize(); ev[0].record()
    for i in range(reps):
        call(); ev[i + 1].record()
    torch.cuda.synchronize()
    t = np.array([ev[i].elapsed_time(ev[i + 1]) for i in range(reps)])
    ms = float(np.median(t))
    nbytes = 16 * F * H * W
    gbs = nbytes / (ms * 1e-3) / 1e9
    print(json.dumps({"kernel": "frame_cloud_kernel (+ frame table)", "frames": F, "H": H, "W": W, "ms_median": round(ms, 4),
                      "ms_min": round(float(t.min()), 4), "ms_max": round(float(t.max()), 4), "bytes": nbytes, "GBps": round(gbs, 1),
                      "frac_of_peak": round(gbs / peak, 3), "frames_per_s": round(F / (ms * 1e-3))}), flush=True)
