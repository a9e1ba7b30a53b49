import os, sys, numpy as np, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/3d-localisation-and-mapping_b200")
from lm3d import lift, synth
dev = torch.device("cuda:0")
seq = synth.make_config("C2", frames=2000)
depth = torch.from_numpy(seq.depth).to(dev); fo = torch.from_numpy(seq.frame_off()).to(dev)
rect4 = lift.scale_boxes(torch.from_numpy(seq.boxes.reshape(-1, 4)).to(dev), torch.from_numpy(seq.image_wh()).to(dev), fo, 192, 256)
pose7 = torch.from_numpy(seq.pose7).to(dev); intr4 = torch.from_numpy(seq.intr4_depth_res()).to(dev)
plan = lift.LiftPlan(depth.shape[0], rect4.shape[0], dev)
for _ in range(3): lift.lift_boxes(depth, pose7, intr4, rect4, fo, plan=plan)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
torch.cuda.synchronize(); ev[0].record()
for _ in range(10): lift.lift_boxes(depth, pose7, intr4, rect4, fo, plan=plan)
ev[1].record(); torch.cuda.synchronize()
c = plan.workspace[:64].view(torch.int32).cpu().tolist()
print(os.environ.get("LM3D_LIB","default")[-16:], f"{ev[0].elapsed_time(ev[1])/10:.4f} ms / {rect4.shape[0]} boxes; skipped {c[7]}, outside window {c[12]}, overflow {c[13]}, refinements {c[5]}")
