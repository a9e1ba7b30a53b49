"""dev helper (one GPU): what the fused record gather costs the lift kernels apart from the link -- lm3d_lift_boxes_gather with
0 / 2 / 8 "peer" buffers that all live on this GPU (the stores are the same instructions as over NVLink).
usage: [LM3D_LIB=...] python tools/bench_gather_local.py [frames]"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "3d-localisation-and-mapping_b200"))
import torch
from lm3d import _capi, lift, synth
dev = torch.device("cuda:0")
lib = _capi.load()
F = int(sys.argv[1]) if len(sys.argv) > 1 else 40000
_, H, W, B = synth.CONFIGS["C2"]
d = synth.make_sequence_torch(F, H, W, B, seed=7, device=dev)
rect4 = lift.scale_boxes(d["boxes"], d["image_wh"], d["frame_off"], W, H)
nb = F * B
plan = lift.LiftPlan(F, nb, dev, False, H, W)
out = {}
for n_peers in (0, 2, 8, 0, 8):
    peers = [torch.empty((nb, 24), dtype=torch.float32, device=dev) for _ in range(n_peers)]
    tab = (C.c_void_p * max(n_peers, 1))(*[p.data_ptr() for p in peers]) if n_peers else None
    def call():
        st = lib.lm3d_lift_boxes_gather(d["depth"].data_ptr(), F, H, W, d["pose7"].data_ptr(), d["intr4"].data_ptr(), rect4.data_ptr(),
                                        d["frame_off"].data_ptr(), nb, 1000.0, float("inf"), 50.0, plan.records.data_ptr(), None,
                                        plan.workspace.data_ptr(), plan.workspace.numel(), tab, n_peers, 0, torch.cuda.current_stream().cuda_stream)
        _capi.check(st, "lm3d_lift_boxes_gather")
    for _ in range(3): call()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): call()
    e1.record(); torch.cuda.synchronize()
    ok = all(torch.equal(p.view(torch.int32), plan.records[:nb].view(torch.int32)) for p in peers)
    print(f"peers={n_peers}: {e0.elapsed_time(e1)/10:.3f} ms per lift of {F} frames, peers == records: {ok}")
    del peers
