"""GPU parity: the CUDA path, called through the C ABI, against the numpy oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import reference_numpy as ora
from parity import assert_records_match

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=["quad", "hist", "tma"])
def warp_kernel_path(request, monkeypatch):
    """Every parity case runs through each warp-box kernel: "quad" (default: float4 loads + histogram
    percentile), "hist" (scalar loads + histogram percentile; also what W % 4 != 0 tensors take), "tma" (TMA tile ring
    + histogram percentile; rects no tile class fits and W % 4 != 0 tensors still take "hist")."""
    monkeypatch.setenv("LM3D_WARP_PATH", request.param)
    return request.param


def run_cuda(seq, dev, rect4=None, q=50.0, max_depth_mm=float("inf")):
    from lm3d import lift

    depth = torch.from_numpy(seq.depth).to(dev)
    pose7 = torch.from_numpy(seq.pose7).to(dev)
    intr4 = torch.from_numpy(seq.intr4_depth_res()).to(dev)
    frame_off = torch.from_numpy(seq.frame_off()).to(dev)
    if rect4 is None:
        boxes = torch.from_numpy(seq.boxes.reshape(-1, 4)).to(dev)
        image_wh = torch.from_numpy(seq.image_wh()).to(dev)
        rect4_t = lift.scale_boxes(boxes, image_wh, frame_off, seq.depth_width, seq.depth_height)
    else:
        rect4_t = torch.from_numpy(rect4).to(dev)
    rec, os_ = lift.lift_boxes(depth, pose7, intr4, rect4_t, frame_off, q=q, max_depth_mm=max_depth_mm, order_stats=True)
    torch.cuda.synchronize()
    return lift.records_to_numpy(rec), os_.cpu().numpy(), rect4_t.cpu().numpy()


def run_oracle(seq, rect4, q=50.0, max_depth_mm=float("inf")):
    return ora.lift_boxes(seq.depth, seq.pose7, seq.intr4_depth_res(), rect4, seq.frame_off(), 1000.0, max_depth_mm, q)


@pytest.mark.parametrize("name,frames", [("C1", 100), ("C2", 24), ("C3", 2), ("C5", 1)])
def test_config_shapes(cuda_device, name, frames):
    from lm3d import synth

    seq = synth.make_config(name, frames=frames)
    rec, os_, rect4 = run_cuda(seq, cuda_device)
    want_rect = ora.boxes_to_rects(seq.boxes.reshape(-1, 4), np.repeat(seq.image_wh(), seq.boxes.shape[1], axis=0),
                                   (seq.depth_width, seq.depth_height))
    assert np.array_equal(rect4, want_rect), "rect4 not bit-exact"
    assert_records_match(rec, os_, run_oracle(seq, rect4))


def test_large_boxes_legacy_cta_kernel(cuda_device, monkeypatch):
    """CTA boxes default to lift_block_kernel (float4 quads + histogram percentile); LM3D_LARGE_PATH=legacy keeps
    the round-1a CTA kernel (also what W % 4 != 0 tensors take) -- both must agree with the oracle."""
    from lm3d import synth

    monkeypatch.setenv("LM3D_LARGE_PATH", "legacy")
    seq = synth.make_config("C3", frames=2)
    rec, os_, rect4 = run_cuda(seq, cuda_device)
    assert_records_match(rec, os_, run_oracle(seq, rect4))


@pytest.mark.parametrize("q", [0.0, 10.0, 37.5, 50.0, 90.0, 100.0])
def test_percentiles(cuda_device, q):
    from lm3d import synth

    seq = synth.make_sequence(6, 256, 192, 12, seed=77)
    rec, os_, rect4 = run_cuda(seq, cuda_device, q=q)
    assert_records_match(rec, os_, run_oracle(seq, rect4, q=q))


def test_max_depth_and_invalid(cuda_device):
    from lm3d import synth

    seq = synth.make_sequence(4, 256, 192, 10, seed=5)
    seq.depth[0, :, :] = np.nan  # a frame with no valid pixel at all
    seq.depth[1, ::2, :] = -5.0
    seq.depth[2, :, ::3] = np.inf
    rec, os_, rect4 = run_cuda(seq, cuda_device, max_depth_mm=1650.0)
    want = run_oracle(seq, rect4, max_depth_mm=1650.0)
    assert (want["n_valid"][:10] == 0).all()
    assert_records_match(rec, os_, want)


def _rect_case(dev, depth, rects, q=50.0):
    from lm3d import synth

    F, H, W = depth.shape
    rng = np.random.default_rng(0)
    seq = synth.Sequence(
        depth=depth.astype(np.float32), pose7=synth.make_poses(F, rng),
        intrinsics=[dict(image_width=W * 7.5, image_height=H * 7.5, fx=1450.0, fy=1450.0, cx=720.0, cy=960.0)] * F,
        boxes=np.zeros((F, len(rects) // F, 4)), damage_cls=None, conf=None, label=None, depth_width=W, depth_height=H)
    rect4 = np.asarray(rects, dtype=np.int32)
    rec, os_, _ = run_cuda(seq, dev, rect4=rect4, q=q)
    assert_records_match(rec, os_, run_oracle(seq, rect4, q=q))
    return rec


def test_edge_rects(cuda_device):
    rng = np.random.default_rng(3)
    depth = (1000 + 500 * rng.random((1, 256, 192))).astype(np.float32)
    rects = [
        (0, 0, 0, 0),            # single pixel
        (191, 255, 191, 255),    # last pixel
        (0, 0, 191, 255),        # whole frame (large path)
        (10, 10, 10, 200),       # one column
        (5, 7, 190, 7),          # one row
        (3, 3, 10, 10),          # 64 px: exact-sample path
        (3, 3, 11, 10),          # 72 px
        (0, 0, 63, 127),         # 8192 px: largest warp box
        (0, 0, 64, 127),         # 8320 px: smallest CTA box
        (100, 50, 131, 81),      # even count
        (100, 50, 130, 80),      # odd count
    ]
    _rect_case(cuda_device, depth, rects)
    _rect_case(cuda_device, depth, rects, q=25.0)


def test_tile_classes_and_alignment(cuda_device):
    """TMA tiles start at x0 & ~3 and come in 16-column classes: sweep start phases, widths around the
    class boundaries, chunk-boundary heights, frame edges (zero-filled tile elements) and several frames."""
    rng = np.random.default_rng(8)
    depth = (900 + 700 * rng.random((3, 256, 192))).astype(np.float32)
    depth[rng.random(depth.shape) < 0.03] = 0.0
    rects = []
    for f in range(3):
        per = []
        for x0 in (0, 1, 2, 3, 61):
            for w in (1, 13, 16, 17, 29, 31, 32, 33, 47, 48, 49, 77, 80):
                if x0 + w <= 192:
                    per.append((x0, 5 + f, x0 + w - 1, 5 + f + [1, 20, 21, 22, 43, 64, 97][(x0 + w) % 7] - 1))
        per += [(0, 250, 191, 255), (1, 0, 190, 40), (129, 200, 191, 255), (188, 0, 191, 255), (0, 0, 3, 255),
                (2, 100, 5, 227), (96, 128, 191, 200)]
        rects += per
    n = len(rects) // 3
    _rect_case(cuda_device, depth, rects[:n] + rects[n:2 * n] + rects[2 * n:3 * n])
    _rect_case(cuda_device, depth, rects[:n] + rects[n:2 * n] + rects[2 * n:3 * n], q=90.0)


def test_widths_not_multiple_of_16_or_4(cuda_device):
    """W = 100 (tile classes wider than the frame are never needed) and W = 50 (global stride not a
    multiple of 16 bytes: the tensor map cannot be built, the direct-load kernel takes every box)."""
    rng = np.random.default_rng(9)
    for W in (100, 50):
        depth = (900 + 700 * rng.random((2, 64, W))).astype(np.float32)
        rects = []
        for f in range(2):
            rects += [(0, 0, W - 1, 63), (1, 2, W - 2, 30), (W - 7, 10, W - 1, 60), (3, 3, 3, 3), (5, 0, 40, 63),
                      (W // 2, 1, W - 1, 9)]
        _rect_case(cuda_device, depth, rects)


def test_ties_and_constant_planes(cuda_device):
    rng = np.random.default_rng(4)
    depth = np.full((4, 256, 192), 1234.5, dtype=np.float32)           # constant plane
    depth[1] = np.where(rng.random((256, 192)) < 0.5, 1000.0, 2000.0)  # two values
    depth[2] = np.round(1000 + 30 * rng.random((256, 192)))            # ~30 distinct values
    depth[3, :, :96] = 1000.0                                          # two values, split down the middle: rects centred
    depth[3, :, 96:] = 2000.0                                          # on x = 96 have the two median ranks on either side
    rects = []
    for f in range(3):
        rects += [(0, 0, 191, 255), (20, 30, 90, 140), (5, 5, 40, 40), (0, 0, 7, 7)]
    rects += [(0, 0, 191, 255), (56, 10, 135, 49), (76, 100, 115, 139), (92, 7, 99, 14)]
    # frame_off is uniform: 4 rects per frame
    for q in (50.0, 33.0, 0.0, 100.0):
        _rect_case(cuda_device, depth, rects, q=q)


@pytest.mark.parametrize("step_mm", [1.0, 10.0, 250.0])
def test_quantised_depth(cuda_device, step_mm):
    """Depth rounded to a grid (sensor quantisation): the percentile bins hold runs of equal keys, up to whole boxes on
    one value -- the boxes the histogram map cannot split (deferred to the exact key-space select of
    lift_resolve_kernel on the quad path) must stay bit-exact."""
    from lm3d import synth

    seq = synth.make_sequence(6, 256, 192, 14, seed=21)
    d = seq.depth
    q = np.round(d / step_mm) * step_mm
    seq.depth[...] = np.where(np.isfinite(d) & (d > 0), q, d).astype(np.float32)
    rec, os_, rect4 = run_cuda(seq, cuda_device)
    assert_records_match(rec, os_, run_oracle(seq, rect4))
    rec, os_, rect4 = run_cuda(seq, cuda_device, q=25.0)
    assert_records_match(rec, os_, run_oracle(seq, rect4, q=25.0))


def test_quantised_depth_takes_the_exact_select(cuda_device):
    """The case above must really exercise the exact select (workspace counter 4), not pass because the fast path
    happened to cope: 250 mm steps put hundreds of equal keys in the percentile bin of most boxes."""
    from lm3d import lift, synth

    seq = synth.make_sequence(6, 256, 192, 14, seed=21)
    d = seq.depth
    seq.depth[...] = np.where(np.isfinite(d) & (d > 0), np.round(d / 250.0) * 250.0, d).astype(np.float32)
    dev = cuda_device
    fo = torch.from_numpy(seq.frame_off()).to(dev)
    rect4 = lift.scale_boxes(torch.from_numpy(seq.boxes.reshape(-1, 4)).to(dev), torch.from_numpy(seq.image_wh()).to(dev), fo,
                             seq.depth_width, seq.depth_height)
    plan = lift.LiftPlan(seq.depth.shape[0], rect4.shape[0], dev, True)
    rec, os_ = lift.lift_boxes(torch.from_numpy(seq.depth).to(dev), torch.from_numpy(seq.pose7).to(dev),
                               torch.from_numpy(seq.intr4_depth_res()).to(dev), rect4, fo, plan=plan)
    torch.cuda.synchronize()
    if os.environ.get("LM3D_WARP_PATH", "quad") == "quad":
        assert int(plan.workspace[:64].view(torch.int32)[4].cpu()) > 0, "no box took the exact select"
    assert_records_match(lift.records_to_numpy(rec), os_.cpu().numpy(), run_oracle(seq, rect4.cpu().numpy()))


def test_out_of_range_rects_are_clamped(cuda_device):
    rng = np.random.default_rng(5)
    depth = (1000 + 500 * rng.random((1, 64, 48))).astype(np.float32)
    from lm3d import lift, synth

    seq = synth.Sequence(depth=depth, pose7=synth.make_poses(1, rng),
                         intrinsics=[dict(image_width=360.0, image_height=480.0, fx=400.0, fy=400.0, cx=180.0, cy=240.0)],
                         boxes=np.zeros((1, 2, 4)), damage_cls=None, conf=None, label=None, depth_width=48, depth_height=64)
    raw = np.array([[-5, -5, 100, 100], [40, 60, 10, 20]], dtype=np.int32)
    clamped = np.array([[0, 0, 47, 63], [10, 20, 40, 60]], dtype=np.int32)
    rec, os_, _ = run_cuda(seq, cuda_device, rect4=raw)
    assert_records_match(rec, os_, run_oracle(seq, clamped))


def test_ragged_and_empty_frames(cuda_device):
    from lm3d import lift, synth

    seq = synth.make_sequence(5, 256, 192, 6, seed=11)
    counts = [0, 6, 1, 0, 3]
    keep = np.concatenate([np.arange(f * 6, f * 6 + c) for f, c in enumerate(counts)]).astype(np.int64)
    frame_off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    dev = cuda_device
    boxes = torch.from_numpy(seq.boxes.reshape(-1, 4)[keep]).to(dev)
    fo = torch.from_numpy(frame_off).to(dev)
    rect4 = lift.scale_boxes(boxes, torch.from_numpy(seq.image_wh()).to(dev), fo, 192, 256)
    rec, os_ = lift.lift_boxes(torch.from_numpy(seq.depth).to(dev), torch.from_numpy(seq.pose7).to(dev),
                               torch.from_numpy(seq.intr4_depth_res()).to(dev), rect4, fo, order_stats=True)
    want = ora.lift_boxes(seq.depth, seq.pose7, seq.intr4_depth_res(), rect4.cpu().numpy(), frame_off)
    assert_records_match(lift.records_to_numpy(rec), os_.cpu().numpy(), want)
    # B == 0 is a no-op, not an error
    empty = lift.lift_boxes(torch.from_numpy(seq.depth).to(dev), torch.from_numpy(seq.pose7).to(dev),
                            torch.from_numpy(seq.intr4_depth_res()).to(dev),
                            torch.empty((0, 4), dtype=torch.int32, device=dev),
                            torch.zeros(6, dtype=torch.int64, device=dev))
    assert empty.shape[0] == 0


def test_host_entry_matches_device_entry(cuda_device):
    from lm3d import lift, synth

    seq = synth.make_config("C1", frames=40)
    rec_dev, _, rect4 = run_cuda(seq, cuda_device)
    rec_host = lift.lift_boxes_host(seq.depth, seq.pose7, seq.intr4_depth_res(), seq.boxes.reshape(-1, 4),
                                    seq.image_wh(), seq.frame_off())
    assert rec_host.tobytes() == rec_dev.tobytes()  # deterministic kernel => byte-identical
