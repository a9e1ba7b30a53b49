"""Synthetic inputs for the 3-D NMS tests (shared by the CPU and GPU suites)."""
import numpy as np


def clustered_boxes(n_signs: int, per_sign: int, seed: int, jitter: float = 0.03, label_noise: float = 0.05,
                    extent: float = 40.0, n_labels: int = 6):
    """``n_signs`` planar rectangles ("signs") scattered in a cube, each observed ``per_sign`` times with corner
    jitter -- the shape of the lift's output on a real scan (many per-frame boxes per physical sign)."""
    rng = np.random.default_rng(seed)
    B = n_signs * per_sign
    centre = rng.uniform(-extent / 2, extent / 2, size=(n_signs, 3))
    a = rng.normal(size=(n_signs, 3)); a /= np.linalg.norm(a, axis=1, keepdims=True)
    b = np.cross(a, rng.normal(size=(n_signs, 3))); b /= np.linalg.norm(b, axis=1, keepdims=True)
    w = rng.uniform(0.3, 0.9, size=(n_signs, 1)); h = rng.uniform(0.3, 0.9, size=(n_signs, 1))
    quad = np.stack([centre - a * w / 2 - b * h / 2, centre - a * w / 2 + b * h / 2,
                     centre + a * w / 2 + b * h / 2, centre + a * w / 2 - b * h / 2], axis=1)  # [S,4,3]
    corners = np.repeat(quad, per_sign, axis=0) + rng.normal(scale=jitter, size=(B, 4, 3))
    sign_label = rng.integers(0, n_labels, size=n_signs)
    label = np.repeat(sign_label, per_sign)
    flip = rng.random(B) < label_noise
    label = np.where(flip, rng.integers(0, n_labels, size=B), label).astype(np.int32)
    conf = rng.uniform(0.3, 0.99, size=B).astype(np.float32)
    perm = rng.permutation(B)  # observations of one sign are not adjacent in the input
    return corners[perm].astype(np.float32), conf[perm], label[perm]


def chain_boxes(n: int):
    """``n`` unit squares along x, each overlapping only its neighbours, confidence falling along the chain: greedy
    NMS keeps 0, 2, 4, ... and the parallel relaxation needs ~n rounds (the worst case for the round loop)."""
    x = np.arange(n, dtype=np.float64) * 0.5
    corners = np.zeros((n, 4, 3))
    corners[:, :, 0] = x[:, None] + np.array([0.0, 0.0, 1.0, 1.0])
    corners[:, :, 1] = np.array([0.0, 1.0, 1.0, 0.0])
    conf = np.linspace(0.99, 0.01, n).astype(np.float32)
    return corners.astype(np.float32), conf, np.zeros(n, dtype=np.int32)


def brute_force_nms(corners, conf, label, thr, pad):
    """Independent scalar restatement of NMS-SPEC v0 (pure Python loops, float32 per operation)."""
    f = np.float32
    B = len(conf)
    lo = np.zeros((B, 3), f); hi = np.zeros((B, 3), f); vol = np.zeros(B, f); ok = np.zeros(B, bool)
    for i in range(B):
        c = corners[i].reshape(4, 3)
        ok[i] = bool(np.isfinite(c).all()) and bool(np.isfinite(conf[i]))  # N2
        for k in range(3):
            lo[i, k] = f(min(c[:, k])) - f(pad)
            hi[i, k] = f(max(c[:, k])) + f(pad)
        vol[i] = f(f((hi[i, 0] - lo[i, 0]) * (hi[i, 1] - lo[i, 1])) * (hi[i, 2] - lo[i, 2]))
    order = sorted((i for i in range(B) if ok[i]), key=lambda i: (-float(conf[i]), i))
    keep = np.zeros(B, np.uint8); parent = np.full(B, -1, np.int32); kept = []
    for i in order:
        if not ok[i]:
            continue
        sup = -1
        for j in kept:
            if label[j] != label[i]:
                continue
            d = [f(min(hi[i, k], hi[j, k]) - max(lo[i, k], lo[j, k])) for k in range(3)]
            if not all(x > 0 for x in d):
                continue
            inter = f(f(d[0] * d[1]) * d[2])
            union = f(f(vol[i] + vol[j]) - inter)
            if inter > f(f(thr) * union):
                sup = j
                break
        if sup >= 0:
            parent[i] = sup
        else:
            keep[i] = 1; parent[i] = i; kept.append(i)
    return keep, parent
