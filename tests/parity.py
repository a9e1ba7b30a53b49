"""Shared parity checker: CUDA records (through the C ABI) vs the numpy oracle.

Tolerances are the ones BASELINE.json's north_star states: bit-exact for counts, rects and
the two selected order statistics; |a-b| <= 1e-4 * max(|b|, 1 m) for every coordinate.
"""
import numpy as np

RTOL = 1e-4
SCENE_SCALE = 1.0  # metres; guards near-zero world coordinates


def assert_close_coords(got, want, what):
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape, what
    nan_g, nan_w = np.isnan(got), np.isnan(want)
    assert np.array_equal(nan_g, nan_w), f"{what}: NaN pattern differs"
    tol = RTOL * np.maximum(np.abs(want), SCENE_SCALE)
    err = np.abs(got - want)
    bad = ~nan_w & (err > tol)
    assert not bad.any(), f"{what}: max err {np.nanmax(err / tol):.3g} x tol at {np.argwhere(bad)[:5].tolist()}"


def assert_records_match(rec, order_stats, ora):
    """rec: lm3d RECORD_DTYPE[B]; order_stats: float32 [B,2] or None; ora: ORACLE_RECORD[B]."""
    assert rec.shape == ora.shape
    assert np.array_equal(rec["n_pix"], ora["n_pix"]), "n_pix"
    assert np.array_equal(rec["n_valid"], ora["n_valid"]), "n_valid"
    if order_stats is not None:
        os_ = np.asarray(order_stats, dtype=np.float32)
        want = np.stack([ora["d_lo"], ora["d_hi"]], axis=1)
        # bit-exact (NaN == NaN): compare the raw words of non-empty boxes, NaN-ness of empty ones
        ne = ora["n_valid"] > 0
        assert np.array_equal(os_[ne].view(np.uint32), want[ne].view(np.uint32)), "order statistics not bit-exact"
        assert np.isnan(os_[~ne]).all()
    for k in ("corners", "centroid", "aabb_min", "aabb_max", "z_q"):
        assert_close_coords(rec[k], ora[k], k)
