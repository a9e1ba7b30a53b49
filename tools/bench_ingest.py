"""dev helper: lm3d_ingest_depth streaming bandwidth on C2-sized input (10 k frames of 256x192 8UC4 = 1.97 GB in, 1.97 GB out)."""
import os, sys, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "3d-localisation-and-mapping_b200"))
from lm3d import ingest
dev = torch.device("cuda:0")
F, H, W = 10000, 256, 192
raw = (torch.rand((F, H, W), device=dev) * 4 + 0.3).view(torch.uint8).reshape(F, H, W, 4)
out = torch.empty((F, H, W), dtype=torch.float32, device=dev)
for _ in range(3):
    ingest.decode_depth(raw, out=out)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
torch.cuda.synchronize(); ev[0].record()
for _ in range(20):
    ingest.decode_depth(raw, out=out)
ev[1].record(); torch.cuda.synchronize()
ms = ev[0].elapsed_time(ev[1]) / 20
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
gbs = 2 * F * H * W * 4 / (ms * 1e-3) / 1e9
print(json.dumps({"kernel": "ingest_depth_kernel", "ms": ms, "bytes": 2 * F * H * W * 4, "GBps": gbs, "frac_of_peak": gbs / peak, "frames_per_s": F / (ms * 1e-3)}))
