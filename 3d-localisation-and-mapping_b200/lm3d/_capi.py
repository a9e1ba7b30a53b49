"""ctypes binding of ``liblm3d.so`` (the C ABI in ``include/lm3d.h``).

There is no fallback: if the shared library is missing or does not load, importing the
lift raises ``Lm3dLibraryError``.  Build it with ``python __graft_entry__.py build`` (or
``make -C 3d-localisation-and-mapping_b200/csrc``).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LM3D_LIB", os.path.join(_HERE, "liblm3d.so"))  # LM3D_LIB: debug build only

RECORD_BYTES = 96
RECORD_WORDS = 24

#: every symbol ``include/lm3d.h`` declares (checked by tests/test_capi_symbols.py)
SYMBOLS = (
    "lm3d_version",
    "lm3d_status_string",
    "lm3d_workspace_bytes",
    "lm3d_lift_workspace_bytes",
    "lm3d_scale_boxes",
    "lm3d_lift_boxes",
    "lm3d_lift_boxes_gather",
    "lm3d_gather_alloc",
    "lm3d_gather_open",
    "lm3d_gather_close",
    "lm3d_gather_free",
    "lm3d_cloud_workspace_bytes",
    "lm3d_lift_frame_cloud",
    "lm3d_ingest_depth",
    "lm3d_lift_boxes_host",
    "lm3d_nms_workspace_bytes",
    "lm3d_nms_boxes",
    "lm3d_kernel_launches",
    "lm3d_profile_enable",
    "lm3d_profile_read",
)


class Lm3dLibraryError(RuntimeError):
    pass


class Lm3dError(RuntimeError):
    def __init__(self, status: int, what: str):
        super().__init__(f"{what}: lm3d status {status} ({status_string(status)})")
        self.status = status


_lib = None


def load():
    """Load ``liblm3d.so`` once and declare the prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise Lm3dLibraryError(
            f"{LIB_PATH} not found: the CUDA extension is not built. "
            "Run `python __graft_entry__.py build` (nvcc, sm_100a). There is no CPU fallback."
        )
    try:
        lib = C.CDLL(LIB_PATH)
    except OSError as exc:  # pragma: no cover - depends on the box
        raise Lm3dLibraryError(f"cannot load {LIB_PATH}: {exc}") from exc
    vp, i64, i32, dbl, sz = C.c_void_p, C.c_int64, C.c_int32, C.c_double, C.c_size_t
    lib.lm3d_version.restype = C.c_int
    lib.lm3d_version.argtypes = []
    lib.lm3d_status_string.restype = C.c_char_p
    lib.lm3d_status_string.argtypes = [C.c_int]
    lib.lm3d_workspace_bytes.restype = sz
    lib.lm3d_workspace_bytes.argtypes = [i64, i64]
    lib.lm3d_lift_workspace_bytes.restype = sz
    lib.lm3d_lift_workspace_bytes.argtypes = [i64, i32, i32, i64]
    lib.lm3d_scale_boxes.restype = C.c_int
    lib.lm3d_scale_boxes.argtypes = [vp, vp, vp, i64, i64, i32, i32, vp, vp]
    lib.lm3d_lift_boxes.restype = C.c_int
    lib.lm3d_lift_boxes.argtypes = [vp, i64, i32, i32, vp, vp, vp, vp, i64, dbl, dbl, dbl, vp, vp, vp, sz, vp]
    lib.lm3d_lift_boxes_gather.restype = C.c_int
    lib.lm3d_lift_boxes_gather.argtypes = [vp, i64, i32, i32, vp, vp, vp, vp, i64, dbl, dbl, dbl, vp, vp, vp, sz, vp, i32, i64, vp]
    lib.lm3d_gather_alloc.restype = C.c_int
    lib.lm3d_gather_alloc.argtypes = [sz, C.POINTER(vp), vp]
    lib.lm3d_gather_open.restype = C.c_int
    lib.lm3d_gather_open.argtypes = [vp, C.POINTER(vp)]
    lib.lm3d_gather_close.restype = C.c_int
    lib.lm3d_gather_close.argtypes = [vp]
    lib.lm3d_gather_free.restype = C.c_int
    lib.lm3d_gather_free.argtypes = [vp]
    lib.lm3d_lift_frame_cloud.restype = C.c_int
    lib.lm3d_lift_frame_cloud.argtypes = [vp, i64, i32, i32, vp, vp, dbl, dbl, vp, vp, vp, sz, vp]
    lib.lm3d_cloud_workspace_bytes.restype = sz
    lib.lm3d_cloud_workspace_bytes.argtypes = [i64]
    lib.lm3d_ingest_depth.restype = C.c_int
    lib.lm3d_ingest_depth.argtypes = [vp, i64, C.c_float, vp, vp]
    lib.lm3d_lift_boxes_host.restype = C.c_int
    lib.lm3d_lift_boxes_host.argtypes = [vp, i64, i32, i32, vp, vp, vp, vp, vp, i64, dbl, dbl, dbl, vp, C.c_int]
    lib.lm3d_nms_workspace_bytes.restype = sz
    lib.lm3d_nms_workspace_bytes.argtypes = [i64]
    lib.lm3d_nms_boxes.restype = C.c_int
    lib.lm3d_nms_boxes.argtypes = [vp, i64, vp, vp, i64, C.c_float, C.c_float, vp, vp, C.POINTER(C.c_int32), vp, sz, vp]
    lib.lm3d_kernel_launches.restype = i64
    lib.lm3d_kernel_launches.argtypes = []
    lib.lm3d_profile_enable.restype = C.c_int
    lib.lm3d_profile_enable.argtypes = [C.c_int]
    lib.lm3d_profile_read.restype = C.c_int
    lib.lm3d_profile_read.argtypes = [vp]
    _lib = lib
    return lib


def status_string(status: int) -> str:
    return load().lm3d_status_string(int(status)).decode()


def check(status: int, what: str):
    if status != 0:
        raise Lm3dError(status, what)
