"""Generate tests/golden/depth_ingest.npz by running the REFERENCE's own depth / calibration loaders.

    python tests/golden/make_golden_ingest.py     (needs /root/reference, cv2, torchvision; build container only)

``/root/reference/src/detector/dataset.py::ImageDataset._load_depth_image`` (:68-81) and
``_load_calibration`` (:102-121) run UNMODIFIED on files written here: depth PNGs in the 8UC4
encoding RTAB-Map exports (the four bytes of each fp32 metre value as one BGRA pixel, the way
``src/detector/database_query.py:28-42`` writes them) and an OpenCV-style calibration YAML.  Only
``natsort`` (not installed; used by the constructor to order file names) is replaced by ``sorted``.
The fixture stores the decoded PNG bytes exactly as ``cv2.imread`` returns them and the tensors /
dict the reference produced, so it pins the ingest row of SURVEY.md 8(f) on the reference itself.
"""
import os
import sys
import tempfile
import types

import cv2
import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    nat = types.ModuleType("natsort")
    nat.natsorted = sorted
    sys.modules["natsort"] = nat
    sys.path.insert(0, REF)
    from src.detector.dataset import ImageDataset

    H, W, F = 64, 48, 3
    rng = np.random.default_rng(2024)
    metres = (0.3 + 4.0 * rng.random((F, H, W))).astype(np.float32)
    metres[0, :4, :] = 0.0
    metres[1, 5, 5] = np.nan
    metres[1, 6, 6] = np.inf
    metres[2, 7, :] = -1.25
    metres[2, 8, 8] = np.float32(1e-42)  # denormal
    with tempfile.TemporaryDirectory() as tmp:
        for d in ("rgb", "depth", "calib"):
            os.makedirs(os.path.join(tmp, d))
        for f in range(F):
            png = metres[f].view(np.uint8).reshape(H, W, 4)
            assert cv2.imwrite(os.path.join(tmp, "depth", f"{f + 1}.png"), png)
            open(os.path.join(tmp, "rgb", f"{f + 1}.jpg"), "wb").close()
        yaml_text = (
            "image_width: 1440\nimage_height: 1920\ncamera_matrix:\n  rows: 3\n  cols: 3\n"
            "  data: [1450.25, 0.0, 721.5, 0.0, 1449.75, 958.25, 0.0, 0.0, 1.0]\n"
        )
        with open(os.path.join(tmp, "calib", "1.yaml"), "w") as fh:
            fh.write(yaml_text)
        ds = ImageDataset(os.path.join(tmp, "rgb"), os.path.join(tmp, "depth"), os.path.join(tmp, "calib"), 640,
                          depth_width=W, depth_height=H, processing=False)
        raw = np.stack([cv2.imread(os.path.join(tmp, "depth", f"{f + 1}.png"), cv2.IMREAD_UNCHANGED) for f in range(F)])
        want = np.stack([ds._load_depth_image(os.path.join(tmp, "depth", f"{f + 1}.png")).numpy() for f in range(F)])
        calib = ds._load_calibration(os.path.join(tmp, "calib", "1.yaml"))
    assert raw.dtype == np.uint8 and raw.shape == (F, H, W, 4) and want.dtype == np.float32
    out = os.path.join(HERE, "depth_ingest.npz")
    np.savez_compressed(out, raw_8uc4=raw, depth_mm=want,
                        calib_keys=np.array(sorted(calib)), calib_vals=np.array([float(calib[k]) for k in sorted(calib)]),
                        yaml_text=np.array(yaml_text))
    print("wrote", out, raw.shape, want.shape, calib)


if __name__ == "__main__":
    main()
