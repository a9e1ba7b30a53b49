"""The C-ABI shared library loads and exports every symbol include/lm3d.h declares.
No compute calls: runs without a GPU."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from lm3d import _capi

    if not os.path.exists(_capi.LIB_PATH):
        import __graft_entry__

        __graft_entry__.build()
    return _capi.load()


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "lm3d.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(lm3d_[a-z_0-9]+)\s*\(", hdr)))


def test_header_symbols_are_exported(lib):
    from lm3d import _capi

    syms = declared_symbols()
    assert "lm3d_lift_boxes" in syms and len(syms) >= 8
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in lm3d.h but not exported by liblm3d.so"
    assert sorted(_capi.SYMBOLS) == syms, "lm3d/_capi.py SYMBOLS out of sync with include/lm3d.h"


def test_version_status_and_workspace(lib):
    assert lib.lm3d_version() == 100
    assert lib.lm3d_status_string(0) == b"ok"
    assert b"workspace" in lib.lm3d_status_string(-2)
    w1 = lib.lm3d_workspace_bytes(10, 100)
    w2 = lib.lm3d_workspace_bytes(1000, 20000)
    assert 0 < w1 < w2 and w2 % 256 == 0
    assert lib.lm3d_workspace_bytes(-1, 5) == 0


def test_lift_workspace_adds_the_tile_scratch(lib, monkeypatch):
    """lm3d_lift_workspace_bytes = the minimum + per-frame box areas + one cursor per frame chunk + 64-byte summaries of every
    complete 16 x 16 tile for one chunk of frames (256 by default, LM3D_TILE_CHUNK); small frames and W % 4 != 0 add nothing."""
    for k in ("LM3D_TILE_PATH", "LM3D_TILE_CHUNK"):
        monkeypatch.delenv(k, raising=False)
    F, H, W, B = 1000, 1440, 1920, 50000
    base = lib.lm3d_workspace_bytes(F, B)
    full = lib.lm3d_lift_workspace_bytes(F, H, W, B)
    per_frame = (W // 16) * (H // 16) * 64
    assert base < full and full % 256 == 0
    assert 256 * per_frame <= full - base <= 256 * per_frame + 4 * F + 4096
    assert lib.lm3d_lift_workspace_bytes(F, 192, 256, B) == base           # under 2^18 pixels: no tile path
    assert lib.lm3d_lift_workspace_bytes(F, H, W + 2, B) == base           # W % 4 != 0
    assert lib.lm3d_lift_workspace_bytes(10, H, W, B) - lib.lm3d_workspace_bytes(10, B) < 11 * per_frame   # fewer frames than a chunk
    monkeypatch.setenv("LM3D_TILE_CHUNK", "8")
    assert 8 * per_frame <= lib.lm3d_lift_workspace_bytes(F, H, W, B) - base <= 8 * per_frame + 8 * F + 4096
    monkeypatch.setenv("LM3D_TILE_PATH", "off")
    assert lib.lm3d_lift_workspace_bytes(F, H, W, B) == base


def test_argument_validation_happens_before_any_cuda_call(lib):
    # q outside [0,100], bad sizes and null pointers are rejected with LM3D_ERR_BAD_ARG (-1)
    z = ctypes.c_void_p(0)
    assert lib.lm3d_lift_boxes(z, 1, 4, 4, z, z, z, z, 1, 1000.0, float("inf"), 150.0, z, z, z, 0, z) == -1
    assert lib.lm3d_lift_boxes(z, 1, 0, 4, z, z, z, z, 1, 1000.0, float("inf"), 50.0, z, z, z, 0, z) == -1
    assert lib.lm3d_lift_boxes(z, 1, 4, 4, z, z, z, z, 1, 1000.0, float("inf"), 50.0, z, z, z, 0, z) == -1
    assert lib.lm3d_lift_boxes(z, 1, 4, 4, z, z, z, z, 0, 1000.0, float("inf"), 50.0, z, z, z, 0, z) == 0  # B == 0: no-op
    assert lib.lm3d_scale_boxes(z, z, z, 1, 3, 4, 4, z, z) == -1
    assert lib.lm3d_kernel_launches() == 0


def test_nms_argument_validation(lib):
    z = ctypes.c_void_p(0)
    one = ctypes.c_void_p(16)          # any non-null, 16-byte aligned value: rejected before it is dereferenced
    r = ctypes.c_int32(7)
    f = ctypes.c_float
    assert lib.lm3d_nms_workspace_bytes(-1) == 0
    w1, w2 = lib.lm3d_nms_workspace_bytes(100), lib.lm3d_nms_workspace_bytes(200000)
    assert 0 < w1 < w2 and w2 % 256 == 0
    assert lib.lm3d_nms_boxes(z, 24, z, z, 0, f(0.1), f(0.03), z, z, ctypes.byref(r), z, 0, z) == 0 and r.value == 0  # B == 0
    assert lib.lm3d_nms_boxes(one, 24, one, one, -1, f(0.1), f(0.03), one, z, None, one, 1 << 20, z) == -1
    assert lib.lm3d_nms_boxes(one, 11, one, one, 5, f(0.1), f(0.03), one, z, None, one, 1 << 20, z) == -1   # stride < 12
    assert lib.lm3d_nms_boxes(one, 24, one, one, 5, f(-0.1), f(0.03), one, z, None, one, 1 << 20, z) == -1  # thr < 0
    assert lib.lm3d_nms_boxes(one, 24, one, one, 5, f(float("nan")), f(0.03), one, z, None, one, 1 << 20, z) == -1
    assert lib.lm3d_nms_boxes(z, 24, one, one, 5, f(0.1), f(0.03), one, z, None, one, 1 << 20, z) == -1     # null corners
    assert lib.lm3d_nms_boxes(one, 24, one, one, 5, f(0.1), f(0.03), one, z, None, ctypes.c_void_p(24), 1 << 20, z) == -3
    assert lib.lm3d_nms_boxes(one, 24, one, one, 5, f(0.1), f(0.03), one, z, None, one, 64, z) == -2        # workspace too small
    assert b"invariant" in lib.lm3d_status_string(-6)
    assert lib.lm3d_kernel_launches() == 0


def test_no_cpu_fallback_when_library_missing(monkeypatch, tmp_path):
    from lm3d import _capi

    monkeypatch.setattr(_capi, "_lib", None)
    monkeypatch.setattr(_capi, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_capi.Lm3dLibraryError):
        _capi.load()
