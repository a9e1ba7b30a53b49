#!/usr/bin/env python
"""bench.py -- throughput of the bbox->3D lift (BASELINE.json metric) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl lm3d|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A *step* is one pass of the hot path (lm3d_scale_boxes + lm3d_lift_boxes) over one synthetic sequence PER GPU.

  N = 1   config C2 of BASELINE.json: 10k frames, 256x192 fp32 depth, 20 boxes/frame -- the config the metric is
          quoted on.
  N > 1   config C4: the 1M-frame fleet of C2-shaped scans, one 125k-frame shard (1M / 8) per GPU -- weak scaling,
          frames sharded, no data-path exchange except the per-box records, which every rank ends up holding:
          the gather is FUSED into the lift's epilogue (lm3d_lift_boxes_gather: peer-mapped buffers, NVLink
          stores from the kernels); LM3D_GATHER=nccl runs the NCCL all-gather baseline instead.  The per-GPU step
          of the SAME size at N = 1 is in other_configs["C4_shard"], so efficiency can be read against it.
Rank 0 prints ONE JSON line.

  value          frames/s, whole job, inputs resident in HBM, CUDA-event timed, max over ranks
  e2e            same metric through the HOST-buffer C-ABI call (lm3d_lift_boxes_host): pinned host inputs -> H2D ->
                 lift -> D2H records, all inside the timed region (each rank pinned to its GPU's NUMA node)
  dropin_e2e     the reference-facing Python API itself: ProcessPose.get_global_coordinates() (nested rows) and
                 .get_global_records() (columnar) on the C2 sequence held in host memory (rank 0, N = 1)
  roofline       the dominant stage timed by events bracketing it inside the C ABI; achieved = algorithmic bytes
                  / that time
  other_configs  C3 / C5 (large frames, tile path) and the C4 shard at N = 1: value + roofline each
  cpu_baseline   the numpy oracle (loop form, mirrors pose_processor.py:91-208) on the host cores of this box,
                 bounded sample (rank 0, N = 1 only); variants: one process, vectorised, with the dead cloud

--impl reference times that CPU path alone (all host cores) and prints the same line shape.
"""
from __future__ import annotations

import argparse
import glob
import hashlib
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "3d-localisation-and-mapping_b200")
for _p in (ROOT, PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

METRIC = "frames_per_s_lifted"
UNIT = "frames/s"
C4_SHARD_FRAMES = 125_000  # 1M frames / 8 GPUs


# ------------------------------------------------------------------------------------------
# CPU reference arm (the only place outside tests/ and smoke() that executes oracle/)
# ------------------------------------------------------------------------------------------
_G = {}


def _cpu_worker(span):
    from oracle import reference_numpy as ora

    f0, f1 = span
    seq, boxes = _G["seq"], _G["boxes"]
    n = 0
    for f in range(f0, f1):
        rows = ora.process_frame_loop(
            seq.pose7[f], seq.depth[f], boxes[f], seq.intrinsics[f], seq.depth_width, seq.depth_height,
            with_cloud=_G.get("with_cloud", False),
        )
        n += len(rows)
    return n


class CpuReference:
    """Oracle loop form, frame-sharded over host processes (fork), in-memory arrays."""

    def __init__(self, frames_per_step: int, procs: int, with_cloud: bool = False):
        import multiprocessing as mp

        from lm3d import synth

        F, H, W, B = synth.CONFIGS["C2"]
        self.procs = procs
        self.frames = frames_per_step
        self.boxes_per_frame = B
        seq = synth.make_sequence(frames_per_step, H, W, B, seed=1234 + 2)
        _G["seq"] = seq
        _G["boxes"] = seq.bbox_coordinates()
        _G["with_cloud"] = with_cloud
        self.pool = mp.get_context("fork").Pool(procs) if procs > 1 else None
        edges = [(i * frames_per_step) // procs for i in range(procs + 1)]
        self.spans = [(edges[i], edges[i + 1]) for i in range(procs) if edges[i + 1] > edges[i]]

    def step(self) -> float:
        t0 = time.perf_counter()
        if self.pool is None:
            n = _cpu_worker(self.spans[0])
        else:
            n = sum(self.pool.map(_cpu_worker, self.spans))
        dt = time.perf_counter() - t0
        assert n == self.frames * self.boxes_per_frame
        return dt

    def step_vectorised(self) -> float:
        """The batched oracle form (one numpy call chain per box, no per-corner pose rebuild), one process."""
        import numpy as np

        from oracle import reference_numpy as ora

        seq = _G["seq"]
        B = self.boxes_per_frame
        t0 = time.perf_counter()
        rect4 = ora.boxes_to_rects(seq.boxes.reshape(-1, 4), np.repeat(seq.image_wh(), B, axis=0), (seq.depth_width, seq.depth_height))
        ora.lift_boxes(seq.depth, seq.pose7, seq.intr4_depth_res(), rect4, seq.frame_off())
        return time.perf_counter() - t0

    def close(self):
        if self.pool is not None:
            self.pool.close()
            self.pool.join()


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def workload_text(name, F, H, W, B):
    return f"{name}: {F} frames x {H}x{W} fp32 depth, {B} boxes/frame per GPU (SURVEY 8d law, generated in HBM)"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from lm3d import synth

    F, H, W, B = synth.CONFIGS["C2"]
    cores = host_cores()
    per_step = 24 * cores
    ref = CpuReference(per_step, cores)
    for _ in range(args.warmup):
        ref.step()
    times = [ref.step() for _ in range(args.steps)]
    ref.close()
    total = sum(times)
    fps = per_step * args.steps / total
    name = "C2" if args.gpus == 1 else "C4 (1M-frame fleet, one 1/8 shard per GPU)"
    Fg = F if args.gpus == 1 else C4_SHARD_FRAMES
    line = {
        "impl": "reference",
        "metric": METRIC,
        "value": fps,
        "unit": UNIT,
        "boxes_per_s": fps * B,
        "n_gpus": args.gpus,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f64",
        "data": "synthetic",
        "config": {
            "workload": workload_text(name, Fg, H, W, B),
            "sample": f"{per_step} frames/step of the same law (CPU-bounded sample; the rate does not depend on the frame count)",
        },
        "cpu_baseline": {
            "value": fps,
            "unit": UNIT,
            "cores": cores,
            "kind": "port",
            "sample": f"{per_step} frames x {B} boxes per step, numpy oracle loop form, {cores} processes",
        },
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------
# clocks sampling
# ------------------------------------------------------------------------------------------
class ClockSampler:
    Q = (
        "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
        "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    )

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=open(self.path, "w"), stderr=subprocess.DEVNULL,
            )
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for ln in open(self.path):
                p = [x.strip() for x in ln.split(",")]
                if len(p) < 9:
                    continue
                try:
                    sm.append(float(p[1]))
                    mx.append(float(p[2]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            sm.sort()
            out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max(mx), samples=len(sm))
        out["reasons"] = sorted(reasons)
        return out


def pin_to_gpu_numa_node(local: int):
    """Bind this rank (and the pinned buffers it allocates afterwards: first touch) to the NUMA node its GPU hangs
    off, so that 8 ranks do not funnel their H2D traffic through one socket.  Best effort; returns what was done."""
    try:
        import torch

        prop = torch.cuda.get_device_properties(local)
        bus = f"{getattr(prop, 'pci_domain_id', 0):04x}:{prop.pci_bus_id:02x}:{getattr(prop, 'pci_device_id', 0):02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read().strip())
        if node < 0:
            n_nodes = len(glob.glob("/sys/devices/system/node/node[0-9]*"))
            return {"numa_node": None, "note": f"no NUMA affinity reported for the GPU ({n_nodes} NUMA node(s) visible, "
                                               f"{len(os.sched_getaffinity(0))} CPUs allowed): nothing to pin"}
        cpus = []
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus += list(range(int(a), int(b or a) + 1))
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if allowed:
            os.sched_setaffinity(0, allowed)
        return {"numa_node": node, "cpus": len(allowed)}
    except Exception as exc:  # noqa: BLE001
        return {"numa_node": None, "note": f"not pinned: {type(exc).__name__}"}


def source_sha() -> str:
    h = hashlib.sha256()
    for f in sorted(glob.glob(os.path.join(PKG, "csrc", "*.cu*"))):
        h.update(open(f, "rb").read())
    return h.hexdigest()[:16]


# ------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------
STAGES = ["prep_frames", "prep_boxes", "lift_tma", "lift_warp", "tile_path", "lift_block"]
STAGE_KERNELS = {
    "prep_frames": "prep_frames_kernel", "prep_boxes": "prep_boxes_kernel (+ tile_route_kernel)", "lift_tma": "lift_tma_kernel",
    "lift_warp": "lift_quad_kernel (+ lift_resolve_kernel)", "tile_path": "tile_sum_kernel + tile_box_kernel (all frame chunks)",
    "lift_block": "lift_block_kernel",
}


class Workload:
    """One synthetic sequence resident in HBM + the plans to lift it."""

    def __init__(self, name, F, H, W, B, seed, dev, plans=1):
        import torch

        from lm3d import lift, synth

        self.name, self.F, self.H, self.W, self.B = name, F, H, W, B
        d = synth.make_sequence_torch(F, H, W, B, seed=seed, device=dev, chunk=512 if H * W < 1_000_000 else 8)
        self.depth, self.pose7, self.intr4 = d["depth"], d["pose7"], d["intr4"]
        self.boxes, self.image_wh, self.frame_off = d["boxes"], d["image_wh"], d["frame_off"]
        self.nb = self.boxes.shape[0]
        self.plans = [lift.LiftPlan(F, self.nb, dev, False, H, W) for _ in range(plans)]
        self.rect4 = torch.empty((self.nb, 4), dtype=torch.int32, device=dev)

    def step(self, slot=0, gather=None):
        from lm3d import lift

        lift.scale_boxes(self.boxes, self.image_wh, self.frame_off, self.W, self.H, out=self.rect4)
        return lift.lift_boxes(self.depth, self.pose7, self.intr4, self.rect4, self.frame_off, plan=self.plans[slot],
                               gather=gather, gather_slot=slot)


def measure_roofline(wl, lib, peak, peak_src, reps):
    """Per-stage events inside the C ABI -> the dominant stage's achieved algorithmic GB/s."""
    import ctypes

    import numpy as np
    import torch

    from lm3d import _capi, lift, metrics

    alg_bytes = metrics.algorithmic_bytes(wl.rect4, wl.frame_off, wl.H, wl.W)
    lib.lm3d_profile_enable(1)
    ms = (ctypes.c_float * 6)()
    kern = np.zeros(6)
    for _ in range(reps):
        lift.lift_boxes(wl.depth, wl.pose7, wl.intr4, wl.rect4, wl.frame_off, plan=wl.plans[0])
        _capi.check(lib.lm3d_profile_read(ms), "lm3d_profile_read")
        kern += np.array(list(ms))
    lib.lm3d_profile_enable(0)
    kern /= reps
    c = [int(v) for v in wl.plans[0].workspace[:128].view(torch.int32).cpu()]
    dom = 2 + int(np.argmax(kern[2:6]))
    lift_ms = float(kern[2:6].sum())  # every lift stage (the tile path hands a few boxes to lift_block_kernel)
    achieved = alg_bytes / (kern[dom] * 1e-3) / 1e9
    return {
        "bound": "hbm",
        "kernel": STAGE_KERNELS[STAGES[dom]],
        "achieved": achieved,
        "peak": peak,
        "peak_source": peak_src,
        "unit": "GB/s",
        "frac": achieved / peak,
        "frac_of_nominal_8000": achieved / 8000.0,  # SURVEY 8d asks for both denominators
        "frac_all_lift_stages": alg_bytes / (lift_ms * 1e-3) / 1e9 / peak,
        "traffic": None,
        "algorithmic_bytes_per_launch": alg_bytes,
        "kernel_ms": {STAGES[i]: float(kern[i]) for i in range(6)},
        "warp_path": os.environ.get("LM3D_WARP_PATH", "quad"),
        "rare_paths": {"exact_selects": c[4], "refinement_passes": c[5], "deferred_boxes": c[10],
                       "tile_path_handed_to_block": c[14], "cta_boxes": c[1], "warp_boxes": c[0] + c[8]},
    }


def timed_steps(torch, step, steps, warmup, barrier):
    for _ in range(warmup):
        step()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(steps):
        step()
    ev1.record()
    barrier()
    return ev0.elapsed_time(ev1) / steps


def run_e2e(args, torch, lift, wl, local, barrier, max_over_ranks, world, pin_note):
    """Same metric through the reference-facing HOST-buffer call: pinned host inputs -> H2D -> lift ->
    D2H of the records, every step, all inside the timed region (wall clock around the blocking call)."""
    def pinned(t):
        h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        h.copy_(t)
        return h

    tensors = (wl.depth, wl.pose7, wl.intr4, wl.boxes, wl.image_wh, wl.frame_off)
    host = [pinned(t) for t in tensors]
    h_out_t = torch.empty((wl.nb, 24), dtype=torch.float32, pin_memory=True)
    h_out = h_out_t.numpy().view(lift.RECORD_DTYPE).reshape(-1)
    torch.cuda.synchronize()
    h2d = sum(t.numel() * t.element_size() for t in host)
    d2h = wl.nb * 96
    arrs = [t.numpy() for t in host]

    def e2e_step():
        lift.lift_boxes_host(*arrs, device=local, out=h_out)

    e2e_steps = max(2, min(args.steps, 5))
    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    mine = (time.perf_counter() - t0) / e2e_steps
    barrier()
    e2e_s = max_over_ranks(mine)
    wl.step()
    torch.cuda.synchronize()
    dev_rec = lift.records_to_numpy(wl.plans[0].records[: wl.nb])
    return {
        "value": world * wl.F / e2e_s,
        "unit": UNIT,
        "h2d_bytes_per_step": h2d,
        "d2h_bytes_per_step": d2h,
        "api": "lm3d_lift_boxes_host (pinned host buffers, chunked H2D overlapped with compute)",
        "frames_per_gpu": wl.F,
        "steps": e2e_steps,
        "h2d_gbs_this_rank": h2d / mine / 1e9,
        "numa": pin_note,
        "matches_device_path": bool(dev_rec.tobytes() == h_out.tobytes()),
    }


def run_dropin(wl, local):
    """The drop-in itself: ProcessPose over the C2 sequence held in host memory (what task_def.py:133-142 calls)."""
    import numpy as np
    import pandas as pd

    from lm3d import synth
    from src.mapper.pose_processor import ProcessPose

    depth = wl.depth.cpu().numpy()
    boxes = wl.boxes.cpu().numpy().reshape(wl.F, wl.B, 4)
    rng = np.random.default_rng(0)
    cls, conf, lab = rng.integers(0, 2, (wl.F, wl.B)).tolist(), rng.uniform(0.25, 1, (wl.F, wl.B)).tolist(), rng.integers(0, 8, (wl.F, wl.B)).tolist()
    bl = boxes.tolist()
    bc = {f: [[*bl[f][b], cls[f][b], conf[f][b], lab[f][b]] for b in range(wl.B)] for f in range(wl.F)}
    pose = pd.DataFrame(wl.pose7.cpu().numpy(), columns=["tx", "ty", "tz", "qx", "qy", "qz", "qw"])
    pose.insert(0, "timestamp", np.arange(wl.F) / 30.0)
    intr = dict(image_width=synth.RGB_W, image_height=synth.RGB_H, fx=synth.RGB_FX, fy=synth.RGB_FY, cx=synth.RGB_CX, cy=synth.RGB_CY)
    ds = synth.ArrayDataset(depth, [intr] * wl.F)

    class PerFrame:  # the reference's ImageDataset interface only: dataset[i]
        def __getitem__(self, i):
            return None, depth[i], intr

    out = {"frames": wl.F, "boxes": wl.nb, "unit": UNIT}
    for key, dataset, fn in (("rows_batched_dataset", ds, "get_global_coordinates"), ("records_batched_dataset", ds, "get_global_records"),
                             ("rows_per_frame_dataset", PerFrame(), "get_global_coordinates")):
        pp = ProcessPose(pose, dataset, bc, 640, wl.W, wl.H, device=local)
        getattr(pp, fn)()  # warm-up (staging buffers of the library, page faults)
        t0 = time.perf_counter()
        res = getattr(pp, fn)()
        dt = time.perf_counter() - t0
        out[key] = {"frames_per_s": wl.F / dt, "seconds": dt}
        del res
    return out


def run_ingest(dev, n_frames=2048, H=256, W=192):
    """Batched depth loader (PNG files -> pinned ring -> device, in-place conversion): frames/s from disk."""
    import cv2
    import numpy as np
    import torch

    from lm3d import ingest

    rng = np.random.default_rng(1)
    with tempfile.TemporaryDirectory() as tmp:
        os.makedirs(os.path.join(tmp, "depth"))
        os.makedirs(os.path.join(tmp, "calib"))
        metres = (0.5 + 3.0 * rng.random((8, H, W))).astype(np.float32)
        yaml_text = ("image_width: 1440\nimage_height: 1920\ncamera_matrix:\n  rows: 3\n  cols: 3\n"
                     "  data: [1450.0, 0.0, 720.0, 0.0, 1450.0, 960.0, 0.0, 0.0, 1.0]\n")
        for i in range(n_frames):
            cv2.imwrite(os.path.join(tmp, "depth", f"{i + 1}.png"), metres[i % 8].view(np.uint8).reshape(H, W, 4))
            with open(os.path.join(tmp, "calib", f"{i + 1}.yaml"), "w") as fh:
                fh.write(yaml_text)
        seq = ingest.DepthSequence.from_dirs(os.path.join(tmp, "depth"), os.path.join(tmp, "calib"), depth_width=W, depth_height=H, device=dev)
        frames = list(range(n_frames))
        seq.batch_device(frames[:256])
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        seq.batch_device(frames)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    return {"frames": n_frames, "frames_per_s": n_frames / dt, "workers": seq.workers,
            "what": "DepthSequence.batch_device: cv2 PNG decode on a thread pool -> pinned ring -> H2D -> lm3d_ingest_depth in place (+ YAML)"}


def run_lm3d(args):
    import torch
    import torch.distributed as dist

    from lm3d import _capi, lift, synth
    from lm3d import dist as ldist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl lm3d needs a CUDA device: the lift has no CPU fallback")
    dev = torch.device(f"cuda:{local}")
    torch.cuda.set_device(dev)
    pin_note = pin_to_gpu_numa_node(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    lib = _capi.load()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v: float) -> float:
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"

    # ---- workload: resident in HBM, generated untimed ---------------------------------------------------
    name = args.workload or ("C2" if world == 1 else "C4")
    F, H, W, B = synth.CONFIGS["C2" if name == "C4" else name]
    if name == "C4":
        F = C4_SHARD_FRAMES
    if args.frames:
        F = args.frames
    wl = Workload(name, F, H, W, B, seed=1234 + int(name[1:]) + 1000 * rank, dev=dev, plans=2)
    gather_mode = os.environ.get("LM3D_GATHER", "peer") if world > 1 else "none"
    pg = nccl = None
    gather_note = None
    if gather_mode == "peer":
        try:
            pg = ldist.PeerGather(wl.nb, dev, slots=2)
        except Exception as exc:  # noqa: BLE001  (no peer access between the GPUs of this box: the NCCL baseline)
            gather_mode, gather_note = "nccl", f"peer buffers unavailable ({type(exc).__name__}: {exc}); NCCL all-gather used"
    if gather_mode == "nccl":
        nccl = ldist.PipelinedGather(wl.nb, dev)
    step_no = [0]

    def step():
        slot = step_no[0] & 1
        step_no[0] += 1
        if nccl is not None:
            nccl.ready(slot)  # the gather that last read plans[slot].records must be done
        rec = wl.step(slot, gather=pg)
        if nccl is not None:
            nccl.launch(slot, rec)
        return rec

    def drain():
        if nccl is not None:
            nccl.drain()

    for _ in range(max(args.warmup, 3)):
        step()
    drain()
    barrier()

    # ---- timed region: value ------------------------------------------------------------------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    barrier()
    launches0 = lib.lm3d_kernel_launches()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        step()
    drain()  # NCCL mode: every step's gather has landed before the clock stops (peer mode: the stores ARE the kernels)
    ev1.record()
    barrier()
    launches = lib.lm3d_kernel_launches() - launches0
    my_ms = ev0.elapsed_time(ev1) / args.steps
    ms_per_step = max_over_ranks(my_ms)
    value = world * F / (ms_per_step * 1e-3)
    ms_ranks = [my_ms]
    if world > 1:  # the spread over ranks (every GPU lifts an independent shard of the same size)
        t = torch.zeros(world, dtype=torch.float64, device=dev)
        t[rank] = my_ms
        dist.all_reduce(t)
        ms_ranks = [float(v) for v in t.cpu()]

    # ---- multi-GPU: the gathered bytes must equal the concatenation of the per-rank records (untimed) -----------------
    gather_check = None
    if world > 1:
        slot = (step_no[0] - 1) & 1
        mine = wl.plans[slot].records[: wl.nb]
        ref = ldist.all_gather_records(mine)  # NCCL, the baseline collective
        got = pg.buffer(slot) if pg is not None else nccl.out[slot]
        ok = bool(torch.equal(got.view(torch.int32), ref.view(torch.int32)))
        own = bool(torch.equal(got[rank * wl.nb : (rank + 1) * wl.nb].view(torch.int32), mine.view(torch.int32)))
        t = torch.tensor([int(ok and own)], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        gather_check = bool(t.item())
        del ref

    # ---- roofline leg: per-stage events inside the C ABI ----------------------------------------------------------
    roofline = measure_roofline(wl, lib, peak, peak_src, max(3, min(args.steps, 10)))
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath) and not args.frames:
        try:
            ent = json.load(open(tpath)).get("C2" if name == "C4" else name)
            if isinstance(ent, dict) and ent.get("source_sha") == source_sha():  # only a capture of THIS source counts
                roofline["traffic"] = ent["dram_bytes_per_launch"] * (F / ent["frames"])
                roofline["traffic_source"] = ent.get("source")
        except Exception:
            pass

    # ---- e2e: HOST buffers through lm3d_lift_boxes_host (C2-sized sequence per rank) ---------------------------------
    e2e = None
    if not args.no_e2e:
        wl_e2e = wl
        if F > 20_000:  # a C4 shard is 24.6 GB: the host-buffer leg runs on a C2-sized sequence of the same law
            wl_e2e = Workload("C2", 10_000, H, W, B, seed=99 + rank, dev=dev)
        e2e = run_e2e(args, torch, lift, wl_e2e, local, barrier, max_over_ranks, world, pin_note)
        if wl_e2e is not wl:
            del wl_e2e
    clocks = sampler.stop() if rank == 0 else None

    other, dropin, ingest, cpu = None, None, None, None
    if rank == 0 and world == 1 and not args.frames and name == "C2":
        # ---- the drop-in API itself -------------------------------------------------------------------------------
        if not args.no_dropin:
            dropin = run_dropin(wl, local)
            try:
                ingest = run_ingest(dev)
            except Exception as exc:  # noqa: BLE001
                ingest = {"error": f"{type(exc).__name__}: {exc}"}
        # ---- other configs: C3 / C5 (tile path) and the C4 shard, one GPU ---------------------------------------------
        if not args.no_other:
            other = {}
            del wl.plans[1:]
            for oname, oF in (("C3", args.c3_frames), ("C5", args.c5_frames), ("C4_shard", C4_SHARD_FRAMES)):
                cfg = synth.CONFIGS["C2" if oname == "C4_shard" else oname]
                torch.cuda.empty_cache()
                try:
                    ow = Workload(oname, oF, cfg[1], cfg[2], cfg[3], seed=1234 + (4 if oname == "C4_shard" else int(oname[1:])), dev=dev)
                    oms = timed_steps(torch, ow.step, 5, 3, barrier)
                    oroof = measure_roofline(ow, lib, peak, peak_src, 3)
                    other[oname] = {
                        "workload": workload_text(oname, oF, cfg[1], cfg[2], cfg[3]),
                        "value": oF / (oms * 1e-3), "unit": UNIT, "boxes_per_s": oF * cfg[3] / (oms * 1e-3), "ms_per_step": oms, "steps": 5,
                        "roofline": oroof,
                    }
                    del ow
                except Exception as exc:  # noqa: BLE001
                    other[oname] = {"error": f"{type(exc).__name__}: {exc}"}
        # ---- CPU baselines ------------------------------------------------------------------------------------------
        if not args.no_cpu:
            cores = host_cores()
            per_step = 24 * cores
            ref = CpuReference(per_step, cores)
            ref.step()
            ts = [ref.step() for _ in range(3)]
            t_vec = ref.step_vectorised()
            ref.close()
            one = CpuReference(24, 1)
            t_one = one.step()
            one.close()
            onec = CpuReference(24, 1, with_cloud=True)
            t_onec = onec.step()
            onec.close()
            cpu = {
                "value": per_step * len(ts) / sum(ts),
                "unit": UNIT,
                "cores": cores,
                "kind": "port",
                "sample": f"{per_step} frames x {B} boxes per step x {len(ts)} steps, numpy oracle loop form "
                          f"(pose_processor.py:91-208 shape), {cores} processes",
                "variants": {
                    "cpu_1_loop_form": {"value": 24 / t_one, "cores": 1, "sample": "24 frames, one process (the reference is single-threaded)"},
                    "cpu_1_loop_form_with_dead_cloud": {
                        "value": 24 / t_onec, "cores": 1,
                        "sample": "24 frames, one process, plus the full-frame unprojection the reference computes and drops (pose_processor.py:154-156)"},
                    "cpu_1_vectorised": {"value": per_step / t_vec, "cores": 1, "sample": f"{per_step} frames, batched numpy oracle form, one process"},
                },
            }

    if pg is not None:
        pg.close()
    if rank == 0:
        line = {
            "metric": METRIC,
            "value": value,
            "unit": UNIT,
            "boxes_per_s": value * B,
            "n_gpus": world,
            "steps": args.steps,
            "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step,
            "ms_per_step_by_rank": ms_ranks,
            "higher_is_better": True,
            "scaling": "weak",
            "vs_baseline": None,
            "dtype": "f32",
            "data": "synthetic",
            "config": {
                "workload": workload_text(name if name != "C4" else "C4 (1M-frame fleet, one 1/8 shard per GPU)", F, H, W, B),
                "frames_per_gpu": F,
                "boxes_per_gpu": wl.nb,
                "l2": f"inputs larger than L2 ({wl.depth.numel() * 4 / 1e6:.0f} MB depth per GPU vs 126 MB)",
                "step": "lm3d_scale_boxes + lm3d_lift_boxes"
                        + {"peer": " with the record gather fused into the kernels' epilogue (lm3d_lift_boxes_gather: NVLink peer stores, "
                                   f"{wl.nb * 96 / 1e6:.0f} MB per rank per step to each of {world} ranks)",
                           "nccl": " + NCCL all-gather of records (async, overlapped with the next step's lift)", "none": ""}[gather_mode],
                "gather": gather_mode,
                "gather_note": gather_note,
            },
            "roofline": roofline,
            "cpu_baseline": cpu,
            "e2e": e2e,
            "dropin_e2e": dropin,
            "ingest": ingest,
            "other_configs": other,
            "gather_check": gather_check,
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="lm3d", choices=["lm3d", "reference"])
    ap.add_argument("--frames", type=int, default=0, help="override frames per GPU (diagnostics only)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--workload", default=None, choices=["C1", "C2", "C3", "C4", "C5"],
                    help="diagnostics: another BASELINE config shape (the driver never passes this)")
    ap.add_argument("--no-e2e", action="store_true", help="diagnostics: skip the host-buffer leg")
    ap.add_argument("--no-other", action="store_true", help="skip other_configs (C3 / C5 / C4 shard at N = 1)")
    ap.add_argument("--no-dropin", action="store_true", help="skip the ProcessPose / loader legs")
    ap.add_argument("--c3-frames", type=int, default=2000)
    ap.add_argument("--c5-frames", type=int, default=1000)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_lm3d(args)


if __name__ == "__main__":
    main()
