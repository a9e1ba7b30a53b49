"""GPU parity of the tile path (csrc/lm3d_lift_tiles.cuh): large frames summarised once in 16x16 tiles; a box takes
the sums / extents / counts of the tiles it covers completely, scans only the tiles whose depth range straddles its
percentile bracket, and walks its boundary strips pixel by pixel.  Everything goes through the C ABI and is compared with the numpy oracle; the order
statistics must be bit-exact, and the path must really have been taken (workspace counters)."""
import numpy as np
import pytest
import torch

from oracle import reference_numpy as ora
from parity import assert_records_match

pytestmark = pytest.mark.gpu


def lift_with_plan(seq, dev, rect4=None, q=50.0, max_depth_mm=float("inf"), frame_off=None):
    from lm3d import lift

    depth = torch.from_numpy(seq.depth).to(dev)
    F, H, W = depth.shape
    fo_np = seq.frame_off() if frame_off is None else frame_off
    fo = torch.from_numpy(fo_np).to(dev)
    if rect4 is None:
        rect4_t = lift.scale_boxes(torch.from_numpy(seq.boxes.reshape(-1, 4)).to(dev), torch.from_numpy(seq.image_wh()).to(dev),
                                   fo, seq.depth_width, seq.depth_height)
    else:
        rect4_t = torch.from_numpy(np.asarray(rect4, dtype=np.int32)).to(dev)
    plan = lift.LiftPlan(F, rect4_t.shape[0], dev, True, H, W)
    rec, os_ = lift.lift_boxes(depth, torch.from_numpy(seq.pose7).to(dev), torch.from_numpy(seq.intr4_depth_res()).to(dev),
                               rect4_t, fo, q=q, max_depth_mm=max_depth_mm, plan=plan)
    torch.cuda.synchronize()
    counters = plan.workspace[:64].view(torch.int32).cpu().numpy()
    want = ora.lift_boxes(seq.depth, seq.pose7, seq.intr4_depth_res(), rect4_t.cpu().numpy(), fo_np, 1000.0, max_depth_mm, q)
    return lift.records_to_numpy(rec), os_.cpu().numpy(), want, counters


def make_seq(depth, boxes_per_frame):
    from lm3d import synth

    F, H, W = depth.shape
    rng = np.random.default_rng(0)
    return synth.Sequence(
        depth=depth.astype(np.float32), pose7=synth.make_poses(F, rng),
        intrinsics=[dict(image_width=W * 7.5, image_height=H * 7.5, fx=1450.0, fy=1450.0, cx=720.0, cy=960.0)] * F,
        boxes=np.zeros((F, boxes_per_frame, 4)), damage_cls=None, conf=None, label=None, depth_width=W, depth_height=H)


@pytest.mark.parametrize("name,frames", [("C3", 8), ("C5", 4)])
def test_large_frame_configs_take_the_tile_path(cuda_device, name, frames):
    from lm3d import synth

    seq = synth.make_config(name, frames=frames)
    rec, os_, want, counters = lift_with_plan(seq, cuda_device)
    assert_records_match(rec, os_, want)
    n_large = int((want["n_pix"] > 8160).sum())
    assert n_large > 0.9 * len(want)
    # the CTA-per-box kernel saw only what the tile path handed over (catch-all bins / overfull target bins)
    assert counters[15] == 0, "tile path: collected keys != histogram count"
    assert counters[1] == counters[14] and counters[14] < 0.05 * n_large, (counters[1], counters[14], n_large)


def test_tile_path_equals_block_path_bit_for_bit_on_order_stats(cuda_device, monkeypatch):
    from lm3d import synth

    seq = synth.make_config("C3", frames=2)
    rec_t, os_t, want, c_t = lift_with_plan(seq, cuda_device, q=37.5)
    monkeypatch.setenv("LM3D_TILE_PATH", "off")
    rec_b, os_b, _, c_b = lift_with_plan(seq, cuda_device, q=37.5)
    assert c_b[1] == len(want) and c_t[1] < len(want)
    assert np.array_equal(os_t.view(np.uint32), os_b.view(np.uint32))
    assert np.array_equal(rec_t["n_valid"], rec_b["n_valid"])
    assert_records_match(rec_t, os_t, want)


def _alignment_rects(H, W):
    r = []
    for x0, x1 in ((0, W - 1), (32, 95), (31, 96), (33, 94), (1, W - 2), (64, W - 1), (5, 70), (40, 199 if W > 200 else W - 3)):
        for y0, y1 in ((0, H - 1), (32, 127), (31, 128), (33, 126), (7, H - 9), (96, H - 1)):
            if x1 < W and y1 < H and (x1 - x0 + 1) * (y1 - y0 + 1) > 8160:
                r.append((x0, y0, x1, y1))
    return r


@pytest.mark.parametrize("H,W", [(256, 192), (200, 168), (160, 260)])
@pytest.mark.parametrize("q", [50.0, 12.5])
def test_tile_alignment_sweep_forced_on_small_frames(cuda_device, monkeypatch, H, W, q):
    """LM3D_TILE_PATH=on takes the tile path for any frame; rect edges on / next to tile edges, rects with one row or
    column of interior tiles, rects that reach the right / bottom edge of frames whose size is not a multiple of 32
    (partial edge tiles count as covered), 3 % invalid pixels, two frames."""
    monkeypatch.setenv("LM3D_TILE_PATH", "on")
    monkeypatch.setenv("LM3D_TILE_COVER", "0.01")
    rng = np.random.default_rng(H * 1000 + W)
    depth = (900 + 700 * rng.random((2, H, W)) + 0.5 * np.arange(W)[None, None, :]).astype(np.float32)
    depth[rng.random(depth.shape) < 0.03] = 0.0
    depth[rng.random(depth.shape) < 0.002] = np.nan
    rects = _alignment_rects(H, W)
    assert len(rects) >= 8
    seq = make_seq(depth, len(rects))
    rec, os_, want, counters = lift_with_plan(seq, cuda_device, rect4=rects + rects, q=q)
    assert_records_match(rec, os_, want)
    assert counters[15] == 0 and counters[1] == counters[14]
    assert counters[14] < len(rects), "every box fell back to the CTA-per-box kernel"


def test_tile_path_mixed_with_warp_boxes_and_cover_threshold(cuda_device, monkeypatch):
    """Frames below the cover threshold keep the CTA-per-box kernel; small boxes keep the warp kernel; ragged CSR."""
    monkeypatch.setenv("LM3D_TILE_PATH", "on")
    rng = np.random.default_rng(12)
    H, W = 256, 192
    depth = (1000 + 400 * rng.random((4, H, W))).astype(np.float32)
    big = [(0, 0, 191, 255), (10, 10, 180, 240), (30, 40, 190, 250)]          # frame 0: cover > 1 -> tiles
    small = [(5, 5, 40, 40), (100, 100, 150, 150)]
    rects = big + small + small + [(20, 20, 130, 140)] + small[:1] + big[:2] + small   # frames: 0 | 1 | 2 (cover < 1) | 3
    frame_off = np.array([0, 5, 7, 9, 13], dtype=np.int64)
    seq = make_seq(depth, 1)
    rec, os_, want, counters = lift_with_plan(seq, cuda_device, rect4=rects, frame_off=frame_off)
    assert_records_match(rec, os_, want)
    assert counters[0] == 7                      # warp boxes
    assert counters[1] - counters[14] == 1       # frame 2's large box was routed to the CTA-per-box list directly


@pytest.mark.parametrize("step_mm", [0.25, 5.0, 250.0])
def test_tile_path_quantised_depth_falls_back_exactly(cuda_device, monkeypatch, step_mm):
    """Runs of equal keys overfill the target bins (more than the collect capacity): those boxes are handed to
    lift_block_kernel; the result stays bit-exact either way."""
    from lm3d import synth

    seq = synth.make_config("C3", frames=1)
    d = seq.depth
    seq.depth[...] = np.where(np.isfinite(d) & (d > 0), np.round(d / step_mm) * step_mm, d).astype(np.float32)
    rec, os_, want, counters = lift_with_plan(seq, cuda_device)
    assert_records_match(rec, os_, want)
    assert counters[15] == 0
    if step_mm >= 250.0:
        assert counters[14] > 0, "no box fell back although whole boxes sit on a few values"


@pytest.mark.parametrize("q", [0.0, 100.0, 99.9])
def test_tile_path_extreme_percentiles(cuda_device, q):
    """The minimum / maximum of a box usually sit in the frame map's outer bins or catch-alls."""
    from lm3d import synth

    seq = synth.make_config("C3", frames=1)
    seq.depth[0, 100:110, 100:400] = 9000.0   # far outliers inside many boxes
    seq.depth[0, 300:305, 200:900] = 12.5     # near outliers
    rec, os_, want, counters = lift_with_plan(seq, cuda_device, q=q)
    assert_records_match(rec, os_, want)


def test_tile_path_small_workspace_degrades_to_blocks(cuda_device):
    """A workspace of only lm3d_workspace_bytes(F, B) (the documented minimum) has no room for tiles: same records."""
    from lm3d import lift, synth

    dev = cuda_device
    seq = synth.make_config("C3", frames=1)
    fo = torch.from_numpy(seq.frame_off()).to(dev)
    rect4 = lift.scale_boxes(torch.from_numpy(seq.boxes.reshape(-1, 4)).to(dev), torch.from_numpy(seq.image_wh()).to(dev), fo,
                             seq.depth_width, seq.depth_height)
    plan = lift.LiftPlan(1, rect4.shape[0], dev, True)  # no H, W: minimum workspace
    rec, os_ = lift.lift_boxes(torch.from_numpy(seq.depth).to(dev), torch.from_numpy(seq.pose7).to(dev),
                               torch.from_numpy(seq.intr4_depth_res()).to(dev), rect4, fo, plan=plan)
    torch.cuda.synchronize()
    c = plan.workspace[:64].view(torch.int32).cpu().numpy()
    assert c[14] == 0 and c[1] == rect4.shape[0]
    want = ora.lift_boxes(seq.depth, seq.pose7, seq.intr4_depth_res(), rect4.cpu().numpy(), seq.frame_off())
    assert_records_match(lift.records_to_numpy(rec), os_.cpu().numpy(), want)


def test_tile_path_over_range_and_garbage_pixels(cuda_device):
    """max_depth cuts into the boxes (the listed tiles hold over-range pixels: pass A takes its validity-testing form) and
    10 % of the pixels are 0 / negative / NaN / +-inf: the key-space tests of pass B and of the strips must drop all of
    them without a validity test."""
    from lm3d import synth

    seq = synth.make_config("C3", frames=1)
    rng = np.random.default_rng(5)
    r = rng.random(seq.depth.shape)
    d = seq.depth
    d[r < 0.02] = -d[r < 0.02]
    d[(r >= 0.02) & (r < 0.04)] = np.inf
    d[(r >= 0.04) & (r < 0.06)] = -np.inf
    d[(r >= 0.06) & (r < 0.08)] = -0.0
    d[(r >= 0.08) & (r < 0.10)] = np.nan
    for max_depth in (float("inf"), 2400.0, 1700.0):
        rec, os_, want, counters = lift_with_plan(seq, cuda_device, max_depth_mm=max_depth)
        assert_records_match(rec, os_, want)
        assert counters[14] < 0.2 * len(want), (max_depth, counters[14])


def test_tile_path_constant_plane_hands_over_exactly(cuda_device):
    """Every key of a box ties: the strips alone hold more keys of the (degenerate) bracket than the capture buffer, the
    columns would overflow -- the boxes go to lift_block_kernel, the records stay exact."""
    from lm3d import synth

    seq = synth.make_config("C3", frames=1)
    seq.depth[...] = 1234.5
    seq.depth[0, ::7, ::5] = 0.0
    rec, os_, want, counters = lift_with_plan(seq, cuda_device)
    assert_records_match(rec, os_, want)
    assert counters[14] > 0


@pytest.mark.parametrize("chunk", [1, 2])
def test_tile_path_frame_chunks(cuda_device, monkeypatch, chunk):
    """Five frames in chunks of 1 / 2 (the TileSum scratch is reused chunk after chunk): same records as the oracle."""
    monkeypatch.setenv("LM3D_TILE_PATH", "on")
    monkeypatch.setenv("LM3D_TILE_COVER", "0.01")
    monkeypatch.setenv("LM3D_TILE_CHUNK", str(chunk))
    H, W = 200, 168
    rng = np.random.default_rng(77)
    depth = (900 + 700 * rng.random((5, H, W)) + 0.5 * np.arange(W)[None, None, :]).astype(np.float32)
    depth[rng.random(depth.shape) < 0.03] = 0.0
    rects = _alignment_rects(H, W)[:6]
    seq = make_seq(depth, len(rects))
    rec, os_, want, counters = lift_with_plan(seq, cuda_device, rect4=rects * 5)
    assert_records_match(rec, os_, want)
    assert counters[14] < len(rects) * 5
