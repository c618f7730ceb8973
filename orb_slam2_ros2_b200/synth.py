"""Seeded synthetic inputs shaped like the reference's datasets.

Pure numpy (PCG64) so the same frames are produced in the build container and on the GPU box:
  * KITTI-shaped stereo pairs 1241x376 (config/kitti_config_00.yaml), constant disparity + right-image noise,
  * TUM-shaped gray 640x480 + uint16 depth scaled by 5208 (config/tum_config_f2.yaml),
  * 1920x1080 pairs for the high-resolution stress configuration.
"""
from __future__ import annotations

import numpy as np

KITTI = dict(width=1241, height=376, n_features=2000, n_levels=8, scale_factor=1.2, ini_th=20, min_th=7,
             fx=718.856, fy=718.856, cx=607.1928, cy=185.2157, bl=0.537166, dist=(0.0, 0.0, 0.0, 0.0, 0.0), depth_scale=1.0)
TUM = dict(width=640, height=480, n_features=1000, n_levels=8, scale_factor=1.2, ini_th=20, min_th=7,
           fx=520.908620, fy=521.007327, cx=325.141442, cy=249.701764, bl=0.0767889,
           dist=(0.231222, -0.784899, -0.003257, -0.000105, 0.917205), depth_scale=5208.0)
HD = dict(width=1920, height=1080, n_features=5000, n_levels=12, scale_factor=1.2, ini_th=20, min_th=7,
          fx=1400.0, fy=1400.0, cx=960.0, cy=540.0, bl=0.12, dist=(0.0, 0.0, 0.0, 0.0, 0.0), depth_scale=1.0)


def _gauss_sep(img: np.ndarray, sigma: float) -> np.ndarray:
    r = max(1, int(np.ceil(3 * sigma)))
    x = np.arange(-r, r + 1, dtype=np.float64)
    k = np.exp(-0.5 * (x / sigma) ** 2)
    k = (k / k.sum()).astype(np.float32)
    p = np.pad(img, ((0, 0), (r, r)), mode="reflect")
    out = np.zeros_like(img)
    for i in range(2 * r + 1):
        out += k[i] * p[:, i : i + img.shape[1]]
    p = np.pad(out, ((r, r), (0, 0)), mode="reflect")
    out2 = np.zeros_like(img)
    for i in range(2 * r + 1):
        out2 += k[i] * p[i : i + img.shape[0], :]
    return out2


def synth_image(h: int, w: int, seed: int) -> np.ndarray:
    """Multi-scale blocky texture + noise, lightly blurred: a few thousand FAST corners per pyramid level."""
    rng = np.random.default_rng(seed)
    img = np.zeros((h, w), np.float32)
    for s, a in [(64, 60.0), (24, 50.0), (8, 40.0), (3, 25.0)]:
        gh, gw = (h + s - 1) // s + 1, (w + s - 1) // s + 1
        g = rng.random((gh, gw)).astype(np.float32)
        img += a * np.kron(g, np.ones((s, s), np.float32))[:h, :w]
    img += rng.normal(0.0, 3.0, (h, w)).astype(np.float32)
    img = _gauss_sep(img, 0.8)
    return np.clip(img, 0, 255).astype(np.uint8)


def synth_stereo_pair(h: int, w: int, seed: int, disparity: int = 17):
    """left/right views of one wide base image: x_right = x_left - disparity; +-2 grey-level noise on the right."""
    base = synth_image(h, w + 128, seed)
    left = np.ascontiguousarray(base[:, 64 : 64 + w])
    right = base[:, 64 + disparity : 64 + disparity + w].astype(np.int16)
    right = right + np.random.default_rng(seed + 100).integers(-2, 3, right.shape, dtype=np.int16)
    return left, np.clip(right, 0, 255).astype(np.uint8)


def synth_depth_u16(h: int, w: int, seed: int, depth_scale: float = 5208.0) -> np.ndarray:
    """Smooth random depth in [0.5, 6] m with ~10 % invalid (0) pixels, as uint16 = round(depth_scale * z)."""
    rng = np.random.default_rng(seed + 7)
    g = rng.random(((h + 31) // 32 + 1, (w + 31) // 32 + 1)).astype(np.float32)
    z = np.kron(g, np.ones((32, 32), np.float32))[:h, :w]
    z = 0.5 + 5.5 * _gauss_sep(z, 6.0)
    raw = np.clip(np.round(depth_scale * z), 0, 65535).astype(np.uint16)
    raw[rng.random((h, w)) < 0.10] = 0
    return raw


def synth_stereo_pool(h: int, w: int, n: int, seed0: int = 0, disparity: int = 17):
    """-> (left[n,h,w], right[n,h,w]) uint8"""
    ls, rs = [], []
    for i in range(n):
        l, r = synth_stereo_pair(h, w, seed0 + i, disparity)
        ls.append(l)
        rs.append(r)
    return np.stack(ls), np.stack(rs)


AREA_QUERY_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("radius", "<f4"), ("octave", "<i4"), ("min_level", "<i4"), ("max_level", "<i4")])


def synth_area_queries(kps: np.ndarray, desc: np.ndarray, n: int, seed: int, width: int, height: int, n_levels: int = 8, th: float = 15.0,
                       jitter: float = 3.0, flip_bits: int = 6, mode: str = "window"):
    """Queries shaped like the tracker's searchByProjection calls (src/ORBMatcher.cc:296-314, :575-583) against a frame
    with keypoints `kps` / descriptors `desc`: n keypoints drawn from the frame ("the same point seen in the last frame"),
    moved by N(0, jitter) px, `flip_bits` random descriptor bits flipped, octave window per `mode`:
      "window"  [octave-1, octave+1] clamped      "up"  [octave, n_levels-1]      "down"  [0, octave]
    Every 16th query is sent outside the image or given a huge radius (window clipping), and ~30 % of the frame's
    keypoints are marked excluded ("already has a map point").  -> (queries, query_desc, exclude, source_idx)"""
    rng = np.random.default_rng(seed)
    m = len(kps)
    src = rng.integers(0, max(m, 1), n)
    q = np.zeros(n, AREA_QUERY_DTYPE)
    if m == 0:
        return q, np.zeros((n, 32), np.uint8), np.zeros(0, np.uint8), src
    q["x"] = (kps["x"][src] + rng.normal(0, jitter, n)).astype(np.float32)
    q["y"] = (kps["y"][src] + rng.normal(0, jitter, n)).astype(np.float32)
    q["x"] = np.clip(q["x"], 0, width - 1)
    q["y"] = np.clip(q["y"], 0, height - 1)
    q["octave"] = kps["octave"][src]
    q["radius"] = np.float32(th)
    if mode == "up":
        q["min_level"], q["max_level"] = q["octave"], n_levels - 1
    elif mode == "down":
        q["min_level"], q["max_level"] = 0, q["octave"]
    else:
        q["min_level"], q["max_level"] = np.maximum(0, q["octave"] - 1), np.minimum(n_levels - 1, q["octave"] + 1)
    far = np.arange(n) % 16 == 5
    q["x"][far] += np.float32(width)  # right of the image: empty window
    big = np.arange(n) % 16 == 11
    q["radius"][big] = np.float32(400.0)  # covers the whole image after the sf^2 scaling
    q["min_level"][big], q["max_level"][big] = 0, n_levels - 1
    qd = desc[src].copy()
    bits = rng.integers(0, 256, (n, flip_bits))
    for k in range(flip_bits):
        qd[np.arange(n), bits[:, k] // 8] ^= (1 << (bits[:, k] % 8)).astype(np.uint8)
    exclude = (rng.random(m) < 0.3).astype(np.uint8)
    return q, qd, exclude, src


def synth_vocabulary(k: int, L: int, seed: int, early_leaf: float = 0.03, stop_fraction: float = 0.02, flip_bits: int = 24):
    """A DBoW3-shaped vocabulary tree in the node order of ORB-SLAM2's text format (one record per non-root node, parents
    before children, the k children of a node consecutive): -> dict(k, L, parent[int32 n], is_leaf[uint8 n], desc[uint8 n,32],
    weight[float64 n]) where record i describes node i + 1.  Children are their parent's descriptor with `flip_bits`
    random bits flipped (so that descriptors near a leaf descend to it), a few nodes stop early (k-means ran out of
    points), leaf weights are idf-like positive doubles and a fraction is exactly 0 (stopped words)."""
    rng = np.random.default_rng(seed)
    parents, leaves, descs = [], [], []
    level_ids = np.array([0], np.int64)
    level_desc = rng.integers(0, 256, (1, 32), dtype=np.uint8)
    next_id = 1
    for depth in range(1, L + 1):
        n_par = len(level_ids)
        par = np.repeat(level_ids, k)
        d = np.repeat(level_desc, k, axis=0)
        bits = rng.integers(0, 256, (n_par * k, flip_bits))
        for b in range(flip_bits):
            d[np.arange(n_par * k), bits[:, b] // 8] ^= (1 << (bits[:, b] % 8)).astype(np.uint8)
        leaf = np.ones(n_par * k, np.uint8) if depth == L else (rng.random(n_par * k) < early_leaf).astype(np.uint8)
        ids = next_id + np.arange(n_par * k, dtype=np.int64)
        next_id += n_par * k
        parents.append(par)
        leaves.append(leaf)
        descs.append(d)
        keep = leaf == 0
        level_ids, level_desc = ids[keep], d[keep]
    parent = np.concatenate(parents).astype(np.int32)
    is_leaf = np.concatenate(leaves)
    desc = np.concatenate(descs)
    n = len(parent)
    weight = np.where(is_leaf > 0, np.log(1.0 + 50.0 * rng.random(n)) + 0.01, 0.0)
    weight[(is_leaf > 0) & (rng.random(n) < stop_fraction)] = 0.0
    return dict(k=k, L=L, parent=parent, is_leaf=is_leaf, desc=desc, weight=weight.astype(np.float64))


def write_vocabulary_text(path: str, voc: dict, scoring: int = 0, weighting: int = 0) -> str:
    """ORB-SLAM2 / DBoW3 text vocabulary: 'k L scoring weighting', then 'parent isLeaf d0 .. d31 weight' per node
    (scoring 0 = L1_NORM, weighting 0 = TF_IDF)."""
    with open(path, "w") as f:
        f.write(f"{voc['k']} {voc['L']} {scoring} {weighting}\n")
        for p, l, d, w in zip(voc["parent"], voc["is_leaf"], voc["desc"], voc["weight"]):
            f.write(f"{int(p)} {int(l)} " + " ".join(str(int(x)) for x in d) + f" {float(w)!r}\n")
    return path
