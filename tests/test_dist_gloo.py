"""N>1 host path on CPU: world_size-2 gloo.  Each rank takes its contiguous frame shard, produces
records for it (the oracle stands in for the GPU lift here -- the subject of the test is the
sharding + gather logic of lm3d.dist), and the all-gathered result must equal the 1-rank result
row for row, for equal and for ragged shards."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _records_as_tensor(seq, f0, f1, frame_off, rect4):
    from lm3d import lift
    from oracle import reference_numpy as ora

    loc = frame_off[f0 : f1 + 1] - frame_off[f0]
    b0, b1 = int(frame_off[f0]), int(frame_off[f1])
    rec = ora.lift_boxes(seq.depth[f0:f1], seq.pose7[f0:f1], seq.intr4_depth_res()[f0:f1], rect4[b0:b1], loc)
    out = np.zeros(b1 - b0, dtype=lift.RECORD_DTYPE)
    for k in ("corners", "centroid", "aabb_min", "aabb_max", "z_q", "n_valid", "n_pix"):
        out[k] = rec[k]
    return torch.from_numpy(out.view(np.float32).reshape(-1, 24).copy())


def _worker(rank, world, port, ragged, q):
    for p in (ROOT, os.path.join(ROOT, "3d-localisation-and-mapping_b200")):
        sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from lm3d import dist as ldist
    from lm3d import synth
    from oracle import reference_numpy as ora

    seq = synth.make_sequence(6, 48, 40, 4, seed=21)
    counts = [4, 0, 3, 4, 1, 4] if ragged else [4] * 6
    keep = np.concatenate([np.arange(f * 4, f * 4 + c) for f, c in enumerate(counts)])
    frame_off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    rect4 = ora.boxes_to_rects(seq.boxes.reshape(-1, 4)[keep], np.repeat(seq.image_wh(), 4, axis=0)[keep], (40, 48))
    f0, f1, b0, b1, local = ldist.shard_boxes(frame_off, rank, world)
    mine = _records_as_tensor(seq, f0, f1, frame_off, rect4)
    assert mine.shape[0] == b1 - b0 == int(local[-1])
    if ragged:
        cnt = ldist.gather_counts(mine.shape[0], torch.device("cpu"))
        gathered = ldist.all_gather_records(mine, counts=cnt)
    else:
        gathered = ldist.all_gather_records(mine)
    whole = _records_as_tensor(seq, 0, 6, frame_off, rect4)
    ok = gathered.shape == whole.shape and bool(torch.equal(gathered.view(torch.int32), whole.view(torch.int32)))
    if not ragged:  # the double-buffered async gather used by bench.py gives the same bytes, slot after slot
        pg = ldist.PipelinedGather(mine.shape[0], torch.device("cpu"))
        for i in range(4):
            pg.ready(i & 1)
            pg.launch(i & 1, mine)
        pg.drain()
        ok = ok and all(bool(torch.equal(o.view(torch.int32), whole.view(torch.int32))) for o in pg.out)
    q.put((rank, ok))
    dist.destroy_process_group()


def _run(ragged, port):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ragged, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(res) == [(0, True), (1, True)]


def test_two_rank_gather_equal_shards():
    _run(False, 29611)


def test_two_rank_gather_ragged_shards():
    _run(True, 29612)
