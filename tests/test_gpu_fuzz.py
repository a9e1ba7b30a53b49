"""GPU parity, randomised: hypothesis draws frame shapes (W % 4 != 0 included), rects of both classes (warp boxes up to
8160 px, CTA boxes above), percentiles, invalid-pixel densities and depth quantisation steps; every draw goes through
the C ABI and is compared with the oracle (bit-exact counts / order statistics, 1e-4 coordinates).  Plus constructed
inputs that FORCE the rare paths of the CTA-per-box kernel (bracket refinement, bisection select), with the workspace
counters checked so the paths are known to have run."""
import numpy as np
import pytest
import torch

from oracle import reference_numpy as ora
from parity import assert_records_match

pytestmark = pytest.mark.gpu
hypothesis = pytest.importorskip("hypothesis")
from hypothesis import HealthCheck, given, settings  # noqa: E402
from hypothesis import strategies as st  # noqa: E402


def _lift(dev, depth, rects, frame_off, q, max_depth_mm=float("inf")):
    from lm3d import lift, synth

    F, H, W = depth.shape
    rng = np.random.default_rng(0)
    pose7 = synth.make_poses(F, rng)
    s = 1440.0 / W
    intr4 = np.tile(np.array([1450.0 / s, 1450.0 / s, 720.0 / s, 960.0 / s]), (F, 1))
    rect_t = torch.from_numpy(np.asarray(rects, dtype=np.int32).reshape(-1, 4)).to(dev)
    plan = lift.LiftPlan(F, rect_t.shape[0], dev, True, H, W)
    rec, os_ = lift.lift_boxes(torch.from_numpy(depth).to(dev), torch.from_numpy(pose7).to(dev), torch.from_numpy(intr4).to(dev),
                               rect_t, torch.from_numpy(frame_off).to(dev), q=q, max_depth_mm=max_depth_mm, plan=plan)
    torch.cuda.synchronize()
    want = ora.lift_boxes(depth, pose7, intr4, np.asarray(rects, dtype=np.int32).reshape(-1, 4), frame_off, 1000.0, max_depth_mm, q)
    return lift.records_to_numpy(rec), os_.cpu().numpy(), want, plan.workspace[:128].view(torch.int32).cpu().numpy()


@st.composite
def cases(draw):
    H = draw(st.integers(8, 200))
    W = draw(st.sampled_from([16, 48, 50, 97, 100, 128, 192, 203, 256]))
    F = draw(st.integers(1, 3))
    seed = draw(st.integers(0, 2**31 - 1))
    q = draw(st.sampled_from([0.0, 5.0, 25.0, 50.0, 50.0, 61.8, 99.0, 100.0]))
    p_zero = draw(st.sampled_from([0.0, 0.02, 0.3, 0.9]))
    p_nan = draw(st.sampled_from([0.0, 0.001, 0.05]))
    step = draw(st.sampled_from([0.0, 0.0, 0.5, 7.0, 300.0]))
    n_boxes = draw(st.lists(st.integers(0, 6), min_size=F, max_size=F))
    rects = []
    for f in range(F):
        for _ in range(n_boxes[f]):
            xa, xb = sorted((draw(st.integers(0, W - 1)), draw(st.integers(0, W - 1))))
            ya, yb = sorted((draw(st.integers(0, H - 1)), draw(st.integers(0, H - 1))))
            if draw(st.booleans()):   # bias towards big rects so the CTA class is reached on the larger frames
                xa, xb, ya, yb = xa // 4, W - 1 - (W - 1 - xb) // 4, ya // 4, H - 1 - (H - 1 - yb) // 4
            rects.append((xa, ya, xb, yb))
    max_depth = draw(st.sampled_from([float("inf"), 1900.0]))
    return H, W, F, seed, q, p_zero, p_nan, step, n_boxes, rects, max_depth


@settings(max_examples=40, deadline=None, suppress_health_check=list(HealthCheck), derandomize=True)
@given(cases())
def test_fuzz_against_oracle(cuda_device, case):
    H, W, F, seed, q, p_zero, p_nan, step, n_boxes, rects, max_depth = case
    rng = np.random.default_rng(seed)
    depth = (1200 + 900 * rng.random((F, H, W)) + 2.0 * np.arange(W)[None, None, :]).astype(np.float32)
    if step > 0:
        depth = (np.round(depth / step) * step).astype(np.float32)
    r = rng.random(depth.shape)
    depth[r < p_zero] = 0.0
    depth[r > 1.0 - p_nan] = np.nan
    frame_off = np.concatenate([[0], np.cumsum(n_boxes)]).astype(np.int64)
    if not rects:
        return
    rec, os_, want, _ = _lift(cuda_device, depth, rects, frame_off, q, max_depth)
    assert_records_match(rec, os_, want)


@st.composite
def tile_cases(draw):
    H = draw(st.sampled_from([48, 64, 100, 131, 160, 200]))
    W = draw(st.sampled_from([48, 64, 100, 128, 192, 256]))
    F = draw(st.integers(1, 3))
    seed = draw(st.integers(0, 2**31 - 1))
    q = draw(st.sampled_from([0.0, 5.0, 25.0, 50.0, 50.0, 61.8, 99.0, 100.0]))
    p_zero = draw(st.sampled_from([0.0, 0.02, 0.3, 0.9]))
    p_bad = draw(st.sampled_from([0.0, 0.001, 0.05]))
    step = draw(st.sampled_from([0.0, 0.0, 0.0, 0.5, 7.0, 300.0]))
    patch = draw(st.booleans())
    n_boxes = draw(st.lists(st.integers(0, 5), min_size=F, max_size=F))
    rects = []
    for f in range(F):
        for _ in range(n_boxes[f]):
            xa, xb = sorted((draw(st.integers(0, W - 1)), draw(st.integers(0, W - 1))))
            ya, yb = sorted((draw(st.integers(0, H - 1)), draw(st.integers(0, H - 1))))
            xa, xb, ya, yb = xa // 3, W - 1 - (W - 1 - xb) // 3, ya // 3, H - 1 - (H - 1 - yb) // 3   # big rects: interior tiles exist
            rects.append((xa, ya, xb, yb))
    max_depth = draw(st.sampled_from([float("inf"), 1900.0, 1500.0]))
    return H, W, F, seed, q, p_zero, p_bad, step, patch, n_boxes, rects, max_depth


@settings(max_examples=40, deadline=None, suppress_health_check=list(HealthCheck), derandomize=True)
@given(tile_cases())
def test_fuzz_tile_path_against_oracle(cuda_device, case):
    """The same draw, forced through the tile path (tile_sum_kernel + tile_box_kernel: sampled histogram, key-space counting):
    small frames, so a box has a handful of interior tiles and wide strips; flat patches (many keys on a few values), invalid
    pixels of every kind, quantised depth, max_depth inside the data."""
    import os

    H, W, F, seed, q, p_zero, p_bad, step, patch, n_boxes, rects, max_depth = case
    rng = np.random.default_rng(seed)
    depth = (1200 + 900 * rng.random((F, H, W)) + 2.0 * np.arange(W)[None, None, :]).astype(np.float32)
    if patch:
        depth[:, H // 4 : 3 * H // 4, W // 4 : 3 * W // 4] = (1600.0 + 2.0 * rng.standard_normal((F, 3 * H // 4 - H // 4, 3 * W // 4 - W // 4))).astype(np.float32)
    if step > 0:
        depth = (np.round(depth / step) * step).astype(np.float32)
    r = rng.random(depth.shape)
    depth[r < p_zero] = 0.0
    bad = r > 1.0 - p_bad
    depth[bad] = rng.choice(np.array([np.nan, np.inf, -np.inf, -5.0, 1e9], dtype=np.float32), size=int(bad.sum()))
    frame_off = np.concatenate([[0], np.cumsum(n_boxes)]).astype(np.int64)
    if not rects:
        return
    old = {k: os.environ.get(k) for k in ("LM3D_TILE_PATH", "LM3D_TILE_COVER", "LM3D_TILE_CHUNK")}
    os.environ.update(LM3D_TILE_PATH="on", LM3D_TILE_COVER="0.01", LM3D_TILE_CHUNK="2")
    try:
        rec, os_, want, counters = _lift(cuda_device, depth, rects, frame_off, q, max_depth)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    assert_records_match(rec, os_, want)
    assert counters[15] == 0
    _TILE_FUZZ["large"] += int((want["n_pix"] > 8160).sum())
    _TILE_FUZZ["handed"] += int(counters[14])


_TILE_FUZZ = {"large": 0, "handed": 0}


def test_fuzz_tile_path_was_exercised(cuda_device):
    """(runs after the fuzz above) most CTA-class boxes of the draws were finished by tile_box_kernel itself."""
    assert _TILE_FUZZ["large"] >= 40, _TILE_FUZZ
    assert _TILE_FUZZ["handed"] < 0.6 * _TILE_FUZZ["large"], _TILE_FUZZ


def _lattice_mask(n_pix, samples):
    idx = (np.arange(samples, dtype=np.int64) * n_pix + (n_pix >> 1)) // samples
    m = np.zeros(n_pix, dtype=bool)
    m[idx] = True
    return m


def test_block_kernel_refinement_and_bisection_paths_are_reached(cuda_device, monkeypatch):
    """lift_block_kernel brackets the percentile from a 2048-pixel LATTICE sample of the rect.  Poisoning exactly the
    lattice pixels puts the bracket far from the true median: the histogram pass reports a miss, two refinement
    passes follow (counter 5), and when those miss too the bisection select finishes (counter 4) -- exact either way."""
    monkeypatch.setenv("LM3D_TILE_PATH", "off")
    H, W = 256, 192
    rng = np.random.default_rng(3)
    depth = (1000 + 50 * rng.random((3, H, W))).astype(np.float32)
    rect = (0, 0, W - 1, H - 1)
    lat = _lattice_mask(H * W, 2048).reshape(H, W)   # kBlkBinnedSample (csrc/lm3d_lift_block.cuh)
    n_lat = int(lat.sum())
    depth[0][lat] = (30000.0 + 100.0 * rng.random(n_lat)).astype(np.float32)  # bracket at 30 m: the rank is far below -> two refinements miss -> bisection
    depth[1][lat] = (1.0 + rng.random(n_lat)).astype(np.float32)             # bracket at 1 mm: the rank is far above it
    depth[2][lat] = (1052.0 + rng.random(n_lat)).astype(np.float32)          # a mild miss: one refinement pass catches it
    rects = [rect, (3, 5, W - 2, H - 7)] * 3
    frame_off = np.array([0, 2, 4, 6], dtype=np.int64)
    for q in (50.0, 20.0):
        rec, os_, want, c = _lift(cuda_device, depth, rects, frame_off, q)
        assert_records_match(rec, os_, want)
        assert c[1] == 6 and c[5] >= 3, f"refinement passes: {c[5]}"
        assert c[4] >= 1, f"bisection selects: {c[4]}"


def test_block_kernel_overfull_bins_take_the_exact_paths(cuda_device, monkeypatch):
    """Ties: whole CTA boxes on one or two values overfill the target bins and the private columns."""
    monkeypatch.setenv("LM3D_TILE_PATH", "off")
    H, W = 256, 192
    rng = np.random.default_rng(4)
    depth = np.empty((3, H, W), dtype=np.float32)
    depth[0] = 1234.5
    depth[1] = np.where(rng.random((H, W)) < 0.5, 1000.0, 2000.0)
    depth[2] = np.round(1000 + 3 * rng.random((H, W)))
    rects = [(0, 0, W - 1, H - 1), (10, 20, 180, 240)] * 3
    frame_off = np.array([0, 2, 4, 6], dtype=np.int64)
    for q in (50.0, 0.0, 100.0, 33.0):
        rec, os_, want, c = _lift(cuda_device, depth, rects, frame_off, q)
        assert_records_match(rec, os_, want)
    assert c[4] + c[5] > 0
