#!/usr/bin/env python
"""bench.py -- throughput of the bbox->3D lift (BASELINE.json metric) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl lm3d|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A *step* is one pass of the hot path (lm3d_scale_boxes + lm3d_lift_boxes, and for N>1 the
NCCL all-gather of the per-box records) over one synthetic sequence of config C2
(10k frames, 256x192 fp32 depth, 20 boxes/frame) PER GPU -- weak scaling, frames sharded,
no data-path exchange except the record gather.  Rank 0 prints ONE JSON line.

  value         frames/s, whole job, inputs resident in HBM, CUDA-event timed, max over ranks
  e2e           same metric through the HOST-buffer C-ABI call (lm3d_lift_boxes_host):
                pinned host inputs -> H2D -> lift -> D2H records, all inside the timed region
  roofline      the dominant kernel (warp-per-box lift) timed by events bracketing it inside
                the C ABI; achieved = algorithmic bytes / that time
  cpu_baseline  the numpy oracle (loop form, mirrors pose_processor.py:91-208) on the host
                cores of this box, bounded sample (rank 0, N=1 only)

--impl reference times that CPU path alone (all host cores) and prints the same line shape.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "3d-localisation-and-mapping_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

WORKLOAD = "C2"  # the config BASELINE.json's metric is quoted on; --workload picks another (diagnostics only)
METRIC = "frames_per_s_lifted"
UNIT = "frames/s"


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
            seq.pose7[f], seq.depth[f], boxes[f], seq.intrinsics[f], seq.depth_width, seq.depth_height
        )
        n += len(rows)
    return n


class CpuReference:
    """Oracle loop form, frame-sharded over host processes (fork), in-memory arrays."""

    def __init__(self, frames_per_step: int, procs: int):
        import multiprocessing as mp

        from lm3d import synth

        F, H, W, B = synth.CONFIGS[WORKLOAD]
        self.procs = procs
        self.frames = frames_per_step
        self.boxes_per_frame = B
        seq = synth.make_sequence(frames_per_step, H, W, B, seed=1234 + 2)
        _G["seq"] = seq
        _G["boxes"] = seq.bbox_coordinates()
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

    def close(self):
        if self.pool is not None:
            self.pool.close()
            self.pool.join()


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from lm3d import synth

    F, H, W, B = synth.CONFIGS[WORKLOAD]
    cores = host_cores()
    per_step = 24 * cores
    ref = CpuReference(per_step, cores)
    for _ in range(args.warmup):
        ref.step()
    times = [ref.step() for _ in range(args.steps)]
    ref.close()
    total = sum(times)
    fps = per_step * args.steps / total
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
            "workload": f"{WORKLOAD}: {F} frames x {H}x{W} fp32 depth, {B} boxes/frame per GPU (law of SURVEY 8d)",
            "sample": f"{per_step} frames/step of the same law (CPU-bounded sample)",
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


# ------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------
def run_e2e(args, torch, lift, tensors, nb, local, barrier, max_over_ranks, frames_total, plan):
    """Same metric through the reference-facing HOST-buffer call: pinned host inputs -> H2D -> lift ->
    D2H of the records, every step, all inside the timed region (wall clock around the blocking call)."""
    def pinned(t):
        h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        h.copy_(t)
        return h

    host = [pinned(t) for t in tensors]
    h_out_t = torch.empty((nb, 24), dtype=torch.float32, pin_memory=True)
    h_out = h_out_t.numpy().view(lift.RECORD_DTYPE).reshape(-1)
    torch.cuda.synchronize()
    h2d = sum(t.numel() * t.element_size() for t in host)
    d2h = nb * 96
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
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0) / e2e_steps
    # the host entry and the device entry must agree byte for byte (same deterministic kernels)
    dev_rec = lift.records_to_numpy(plan.records[:nb])
    return {
        "value": frames_total / e2e_s,
        "unit": UNIT,
        "h2d_bytes_per_step": h2d,
        "d2h_bytes_per_step": d2h,
        "api": "lm3d_lift_boxes_host (pinned host buffers, chunked H2D overlapped with compute)",
        "steps": e2e_steps,
        "matches_device_path": bool(dev_rec.tobytes() == h_out.tobytes()),
    }


def run_lm3d(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    from lm3d import _capi, lift, metrics, synth
    from lm3d import dist as ldist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl lm3d needs a CUDA device: the lift has no CPU fallback")
    dev = torch.device(f"cuda:{local}")
    torch.cuda.set_device(dev)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    lib = _capi.load()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v: float) -> float:
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(v: float) -> float:
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ---- workload: one C2-shaped shard per GPU, generated in HBM (untimed) -------------------
    workload = args.workload
    F, H, W, B = synth.CONFIGS[workload]
    if args.frames:
        F = args.frames
    data = synth.make_sequence_torch(F, H, W, B, seed=1234 + int(workload[1:]) + 1000 * rank, device=dev,
                                     chunk=512 if H * W < 1_000_000 else 8)
    depth, pose7, intr4 = data["depth"], data["pose7"], data["intr4"]
    boxes, image_wh, frame_off = data["boxes"], data["image_wh"], data["frame_off"]
    nb = boxes.shape[0]
    plans = [lift.LiftPlan(F, nb, dev, False, H, W), lift.LiftPlan(F, nb, dev, False, H, W)]  # double-buffered: gather(i) overlaps lift(i+1)
    plan = plans[0]
    rect4 = torch.empty((nb, 4), dtype=torch.int32, device=dev)
    gather = ldist.PipelinedGather(nb, dev) if world > 1 else None
    step_no = [0]

    def step():
        slot = step_no[0] & 1
        step_no[0] += 1
        if gather is not None:
            gather.ready(slot)  # the gather that last read plans[slot].records must be done
        lift.scale_boxes(boxes, image_wh, frame_off, W, H, out=rect4)
        rec = lift.lift_boxes(depth, pose7, intr4, rect4, frame_off, plan=plans[slot])
        if gather is not None:
            gather.launch(slot, rec)
        return rec

    def drain():
        if gather is not None:
            gather.drain()

    for _ in range(max(args.warmup, 3)):
        step()
    drain()
    barrier()

    # ---- timed region: value --------------------------------------------------------------
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
    drain()  # every step's gather has landed before the clock stops
    ev1.record()
    barrier()
    launches = lib.lm3d_kernel_launches() - launches0
    ms_total = max_over_ranks(ev0.elapsed_time(ev1))
    ms_per_step = ms_total / args.steps
    frames_total = sum_over_ranks(float(F))
    value = frames_total / (ms_per_step * 1e-3)

    # ---- roofline leg: per-kernel events inside the C ABI -----------------------------------
    alg_bytes = metrics.algorithmic_bytes(rect4, frame_off, H, W)
    lib.lm3d_profile_enable(1)
    import ctypes

    ms4 = (ctypes.c_float * 6)()
    kern = np.zeros(6)
    reps = max(3, min(args.steps, 10))
    for _ in range(reps):
        lift.lift_boxes(depth, pose7, intr4, rect4, frame_off, plan=plan)
        _capi.check(lib.lm3d_profile_read(ms4), "lm3d_profile_read")
        kern += np.array(list(ms4))
    lib.lm3d_profile_enable(0)
    kern /= reps
    # rare-path counters of the last call (workspace words 4..6): exact selects (generic fallbacks; on the quad path
    # the boxes deferred to lift_resolve_kernel), histogram passes beyond the first, candidate overflows
    counters16 = [int(v) for v in plan.workspace[:128].view(torch.int32).cpu()]
    rare = counters16[4:8]
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    dom = 2 + int(np.argmax(kern[2:6]))
    achieved = alg_bytes / (kern[dom] * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get(workload if not args.frames else "", None)
        except Exception:
            traffic = None
    roofline = {
        "bound": "hbm",
        "kernel": ["prep_frames_kernel", "prep_boxes_kernel", "lift_tma_kernel",
                   {"hist": "lift_hist_kernel"}.get(os.environ.get("LM3D_WARP_PATH", ""), "lift_quad_kernel"),
                   "tile_map_kernel + tile_build_kernel + tile_box_kernel (all frame chunks)", "lift_block_kernel"][dom],
        "achieved": achieved,
        "peak": peak,
        "peak_source": peak_src,
        "unit": "GB/s",
        "frac": achieved / peak,
        "frac_of_nominal_8000": achieved / 8000.0,  # SURVEY 8d asks for both denominators
        "traffic": traffic,
        "algorithmic_bytes_per_launch": alg_bytes,
        "kernel_ms": {"prep_frames": kern[0], "prep_boxes": kern[1], "lift_tma": kern[2], "lift_warp": kern[3],
                      "tile_path": kern[4], "lift_block": kern[5]},
        "warp_path": os.environ.get("LM3D_WARP_PATH", "quad"),
        "rare_paths": {"global_fallbacks": rare[0], "narrowing_passes": rare[1], "candidate_overflows": rare[2],
                       "pass2_skipped": rare[3], "tile_path_handed_to_block": counters16[14], "tile_level2": counters16[13], "tile_handed_why": counters16[16:20], "cta_boxes": counters16[1]},
    }

    # ---- e2e: HOST buffers through lm3d_lift_boxes_host --------------------------------------
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(args, torch, lift, (depth, pose7, intr4, boxes, image_wh, frame_off), nb, local, barrier,
                      max_over_ranks, frames_total, plan)
    clocks = sampler.stop() if rank == 0 else None

    # ---- CPU baseline (rank 0, N=1 only) ------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cores = host_cores()
        per_step = 24 * cores
        ref = CpuReference(per_step, cores)
        ref.step()
        ts = [ref.step() for _ in range(3)]
        ref.close()
        cpu = {
            "value": per_step * len(ts) / sum(ts),
            "unit": UNIT,
            "cores": cores,
            "kind": "port",
            "sample": f"{per_step} frames x {B} boxes per step x {len(ts)} steps, numpy oracle loop form "
                      f"(pose_processor.py:91-208 shape), {cores} processes",
        }

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
            "higher_is_better": True,
            "scaling": "weak",
            "vs_baseline": None,
            "dtype": "f32",
            "data": "synthetic",
            "config": {
                "workload": f"{workload}: {F} frames x {H}x{W} fp32 depth, {B} boxes/frame per GPU "
                            f"(SURVEY 8d law, generated in HBM)",
                "frames_per_gpu": F,
                "boxes_per_gpu": nb,
                "l2": f"inputs larger than L2 ({depth.numel() * 4 / 1e6:.0f} MB depth per GPU vs 126 MB)",
                "step": "lm3d_scale_boxes + lm3d_lift_boxes"
                        + (" + NCCL all-gather of records (async, overlapped with the next step's lift)" if world > 1 else ""),
            },
            "roofline": roofline,
            "cpu_baseline": cpu,
            "e2e": e2e,
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
    ap.add_argument("--frames", type=int, default=0, help="override frames per GPU (debug only)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--workload", default=WORKLOAD, choices=["C1", "C2", "C3", "C5"],
                    help="diagnostics: another BASELINE config shape (use with --frames; the driver never passes this)")
    ap.add_argument("--no-e2e", action="store_true", help="diagnostics: skip the host-buffer leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_lm3d(args)


if __name__ == "__main__":
    main()
