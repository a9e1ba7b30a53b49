"""Generate tests/golden/reference_rows.npz by running the REFERENCE's own ProcessPose.

    python tests/golden/make_golden.py          (needs /root/reference; run in the build container)

What is real and what is stubbed
--------------------------------
The class executed is ``/root/reference/src/mapper/pose_processor.py::ProcessPose`` -- its frame
loop (:88-122), intrinsics rescale (:133-137), corner loop with ``int()`` truncation (:184-196),
``_transform_to_global`` (:242-260) and row assembly (:208) all run unmodified.  Three imports
of that file cannot be satisfied in this container and are stubbed *before* import:

* ``open3d``  (not installed): only ``o3d.camera.PinholeCameraIntrinsic`` is touched on this path
  (:144-151) and its result only feeds the dead full-frame cloud -> inert stand-in;
* ``natsort`` (not installed): imported by ``src/detector/dataset.py`` only -> inert stand-in;
* ``src.utils.{config,transformations,visualisation}``: ABSENT from the reference tree.  ``Transforms`` is the
  repo's SHIPPED restatement (``3d-localisation-and-mapping_b200/src/utils/transformations.py``, ORACLE-SPEC v0 of
  SURVEY.md 8c, independent of ``oracle/``); ``Visualiser.parse_images`` passes depth through, the Open3D wrappers
  return None.

So the fixture pins the reference's own control flow / truncation / pose multiply / row format
around the restated helper arithmetic -- the most that can be pinned.
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"

sys.path.insert(0, os.path.join(ROOT, "3d-localisation-and-mapping_b200", "lm3d"))
import synth  # noqa: E402  (imported as a plain module: the package dir also holds a `src` mirror)


def install_stubs():
    o3d = types.ModuleType("open3d")
    o3d.camera = types.SimpleNamespace(PinholeCameraIntrinsic=lambda *a, **k: None)
    o3d.geometry = types.SimpleNamespace(PointCloud=lambda *a, **k: None)
    sys.modules["open3d"] = o3d
    nat = types.ModuleType("natsort")
    nat.natsorted = sorted
    sys.modules["natsort"] = nat

    # The absent src/utils/transformations.py: the repo's own SHIPPED restatement (ORACLE-SPEC v0), loaded from its
    # file so that the fixture is "the reference's ProcessPose + the helper class a user of this repo gets" -- one
    # restatement, not a private stub.  It does not import the oracle.
    import importlib.util

    spec = importlib.util.spec_from_file_location(
        "_shipped_transformations", os.path.join(ROOT, "3d-localisation-and-mapping_b200", "src", "utils", "transformations.py"))
    shipped = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(shipped)
    Transforms = shipped.Transforms

    class Visualiser:
        def parse_images(self, rgb, depth):
            return None, np.asarray(depth, dtype=np.float32)

        def gen_rgbd(self, *a, **k):
            return None

        def gen_point_cloud(self, *a, **k):
            return None

    for name, attrs in (
        ("src.utils", {}),
        ("src.utils.config", {"ConfigLoader": object}),
        ("src.utils.transformations", {"Transforms": Transforms}),
        ("src.utils.visualisation", {"Visualiser": Visualiser}),
    ):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m


def main():
    install_stubs()
    sys.path.insert(0, REF)
    from src.mapper.pose_processor import ProcessPose  # the reference's class

    assert ProcessPose.__module__ == "src.mapper.pose_processor"
    import inspect

    assert inspect.getsourcefile(ProcessPose).startswith(REF), "must execute the reference's own file"

    seq = synth.make_sequence(6, 256, 192, 5, seed=4242)
    # boxes touching / crossing the frame edges exercise the truncation + clamp path
    seq.boxes[0, 0] = [0.0, 0.0, 1440.0, 1920.0]
    seq.boxes[1, 1] = [1300.0, 1700.0, 1440.0, 1920.0]
    seq.boxes[2, 2] = [10.2, 10.7, 17.4, 18.1]
    pp = ProcessPose(
        pose=seq.pose_dataframe(), dataset=seq.dataset(), bbox_coordinates=seq.bbox_coordinates(),
        img_size=640, depth_width=192, depth_height=256,
    )
    out = pp.get_global_coordinates()
    corners = np.array([[np.stack(row[:4]) for row in out[f]] for f in range(6)])  # [F,B,4,3]
    tail = np.array([[row[4:] for row in out[f]] for f in range(6)], dtype=np.float64)
    np.savez_compressed(
        os.path.join(HERE, "reference_rows.npz"),
        seed=4242, F=6, H=256, W=192, B=5, boxes=seq.boxes, corners=corners, tail=tail,
    )
    print("wrote reference_rows.npz", corners.shape, float(np.abs(corners).max()))


if __name__ == "__main__":
    main()
