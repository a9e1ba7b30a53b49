"""GPU: the record gather fused into the lift's epilogue (lm3d_lift_boxes_gather).

Single GPU: the "peers" are plain device buffers -- every kernel path that writes a record (warp kernels and their
deferred-box resolver, the CTA-per-box kernel, the tile path) must push the same bytes to them, at the box offset.
Two or more GPUs (skipped otherwise): one process per GPU over NCCL + CUDA IPC peer buffers; the gathered buffer on
every rank must equal the 1-rank result byte for byte, and so must the NCCL all-gather baseline."""
import ctypes as C
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _call_gather(dev, seq, rect4_np, n_peers, box_offset, cap, q=50.0):
    from lm3d import _capi, lift

    lib = _capi.load()
    F, H, W = seq.depth.shape
    depth = torch.from_numpy(seq.depth).to(dev)
    pose7 = torch.from_numpy(seq.pose7).to(dev)
    intr4 = torch.from_numpy(seq.intr4_depth_res()).to(dev)
    fo = torch.from_numpy(seq.frame_off()).to(dev)
    rect4 = torch.from_numpy(rect4_np).to(dev)
    B = rect4.shape[0]
    plan = lift.LiftPlan(F, B, dev, False, H, W)
    peers = [torch.full((cap, 24), float("nan"), dtype=torch.float32, device=dev) for _ in range(n_peers)]
    tab = (C.c_void_p * n_peers)(*[p.data_ptr() for p in peers])
    st = lib.lm3d_lift_boxes_gather(depth.data_ptr(), F, H, W, pose7.data_ptr(), intr4.data_ptr(), rect4.data_ptr(), fo.data_ptr(), B,
                                    1000.0, float("inf"), q, plan.records.data_ptr(), None, plan.workspace.data_ptr(),
                                    plan.workspace.numel(), tab, n_peers, box_offset, torch.cuda.current_stream().cuda_stream)
    _capi.check(st, "lm3d_lift_boxes_gather")
    torch.cuda.synchronize()
    return plan.records[:B].clone(), peers, plan.workspace[:128].view(torch.int32).cpu().numpy()


@pytest.mark.parametrize("path", ["quad", "hist", "tma"])
def test_every_record_writer_pushes_to_the_peers(cuda_device, monkeypatch, path):
    from lm3d import synth
    from oracle import reference_numpy as ora

    monkeypatch.setenv("LM3D_WARP_PATH", path)
    monkeypatch.setenv("LM3D_TILE_PATH", "on")
    monkeypatch.setenv("LM3D_TILE_COVER", "1.0")
    seq = synth.make_sequence(4, 256, 192, 8, seed=3)
    d = seq.depth
    seq.depth[1] = np.where(np.isfinite(d[1]) & (d[1] > 0), np.round(d[1] / 250.0) * 250.0, d[1])  # ties -> deferred boxes
    rect4 = ora.boxes_to_rects(seq.boxes.reshape(-1, 4), np.repeat(seq.image_wh(), 8, axis=0), (192, 256))
    rect4[0] = (0, 0, 191, 255)      # CTA boxes; frame 0 is covered more than once -> tile path
    rect4[1] = (5, 9, 180, 250)
    rect4[16] = (10, 20, 150, 200)   # frame 2: a CTA box under the cover threshold -> lift_block_kernel
    rec, peers, c = _call_gather(cuda_device, seq, rect4, 3, 5, 32 + 5)
    B = rect4.shape[0]
    assert c[1] >= 1 and c[0] + c[8] >= 20
    for p in peers:
        got = p[5 : 5 + B]
        assert torch.equal(got.view(torch.int32), rec.view(torch.int32))
        assert torch.isnan(p[:5]).all() and torch.isnan(p[5 + B :]).all()   # nothing outside [offset, offset + B)


def _worker(rank, world, port, out_q):
    for p in (ROOT, os.path.join(ROOT, "3d-localisation-and-mapping_b200")):
        sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist

    dev = torch.device(f"cuda:{rank}")
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from lm3d import dist as ldist
    from lm3d import lift, synth

    try:
        F, Bf = 8 * world, 6
        seq = synth.make_sequence(F, 256, 192, Bf, seed=17)
        seq.boxes[0, 0] = [0.0, 0.0, 1440.0, 1920.0]            # a CTA box in rank 0's shard
        seq.boxes[F - 1, 1] = [100.0, 100.0, 1300.0, 1800.0]    # and one in the last rank's
        frame_off = seq.frame_off()

        def lift_range(f0, f1, gather=None, slot=0):
            fo = torch.from_numpy(frame_off[f0 : f1 + 1] - frame_off[f0]).to(dev)
            b0, b1 = int(frame_off[f0]), int(frame_off[f1])
            rect4 = lift.scale_boxes(torch.from_numpy(seq.boxes.reshape(-1, 4)[b0:b1]).to(dev), torch.from_numpy(seq.image_wh()[f0:f1]).to(dev),
                                     fo, 192, 256)
            return lift.lift_boxes(torch.from_numpy(seq.depth[f0:f1]).to(dev), torch.from_numpy(seq.pose7[f0:f1]).to(dev),
                                   torch.from_numpy(seq.intr4_depth_res()[f0:f1]).to(dev), rect4, fo, gather=gather, gather_slot=slot).clone()

        whole = lift_range(0, F)                                  # the 1-rank result, computed on every rank's own GPU
        f0, f1, b0, b1, _ = ldist.shard_boxes(frame_off, rank, world)
        pg = ldist.PeerGather(b1 - b0, dev, slots=2)
        ok = True
        for step in range(3):                                     # slots alternate; a slot is rewritten on step 2
            mine = lift_range(f0, f1, gather=pg, slot=step & 1)
            pg.barrier()
            got = pg.buffer(step & 1)
            ok = ok and bool(torch.equal(got.view(torch.int32), whole.view(torch.int32)))
            ok = ok and bool(torch.equal(mine.view(torch.int32), whole[b0:b1].view(torch.int32)))
            pg.barrier()
        nccl = ldist.all_gather_records(mine)                     # the NCCL baseline gives the same bytes
        ok = ok and bool(torch.equal(nccl.view(torch.int32), whole.view(torch.int32)))
        pg.close()
        out_q.put((rank, ok, ""))
    except Exception as exc:  # noqa: BLE001
        import traceback

        out_q.put((rank, False, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_multi_gpu_gather_equals_single_rank_result(world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res), [m for _, ok, m in res if not ok]
