"""ctypes bindings for the CPU oracle (oracle/liborb_oracle.so) and the compiled reference (oracle/_ref/libref.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs.  The product package (orb_slam2_ros2_b200) never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "liborb_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libref.so")

KP_DTYPE = np.dtype(
    [("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"), ("octave", "<i4"), ("class_id", "<i4")]
)
assert KP_DTYPE.itemsize == 28
MAX_LEVELS = 32
AREA_QUERY_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("radius", "<f4"), ("octave", "<i4"), ("min_level", "<i4"), ("max_level", "<i4")])
assert AREA_QUERY_DTYPE.itemsize == 24

u8p = C.POINTER(C.c_uint8)
i32p = C.POINTER(C.c_int)
f32p = C.POINTER(C.c_float)
f64p = C.POINTER(C.c_double)


def build(force: bool = False) -> None:
    """Compile the oracle (and oracle/_ref when /root/reference is present); `make` is incremental."""
    if force or not os.path.exists(ORACLE_SO) or os.path.isdir("/root/reference"):
        subprocess.run(["make", "-C", HERE, "all"], check=True, capture_output=True)


def _ptr(a, t):
    return a.ctypes.data_as(t) if a is not None else None


class _Pyramid(C.Structure):
    _fields_ = [
        ("n_levels", C.c_int),
        ("w", C.c_int * MAX_LEVELS),
        ("h", C.c_int * MAX_LEVELS),
        ("sf", C.c_float * MAX_LEVELS),
        ("quota", C.c_int * MAX_LEVELS),
        ("img", u8p * MAX_LEVELS),
        ("blur", u8p * MAX_LEVELS),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(ORACLE_SO):
            build()
        _lib = C.CDLL(ORACLE_SO)
        L = _lib
        L.oracle_fast9_nms.restype = C.c_int
        L.oracle_fast9_nms.argtypes = [u8p, C.c_int, C.c_int, C.c_size_t, C.c_int, i32p, i32p, i32p, C.c_int]
        L.oracle_fast_cells.restype = C.c_int
        L.oracle_fast_cells.argtypes = [u8p, C.c_int, C.c_int, C.c_size_t, C.c_int, C.c_int, i32p, i32p, i32p, C.c_int, i32p]
        L.oracle_resize_linear_u8.argtypes = [u8p, C.c_int, C.c_int, C.c_size_t, u8p, C.c_int, C.c_int, C.c_size_t]
        L.oracle_gaussian_blur7_u8.argtypes = [u8p, C.c_int, C.c_int, C.c_size_t, u8p, C.c_size_t]
        L.oracle_undistort_points.argtypes = [f32p, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, f32p, C.c_int]
        L.oracle_quadtree_select.restype = C.c_int
        L.oracle_quadtree_select.argtypes = [C.c_int, C.c_int, C.c_int, f32p, f32p, f32p, C.c_int, i32p, C.POINTER(C.c_long)]
        L.oracle_ic_angle.restype = C.c_double
        L.oracle_ic_angle.argtypes = [u8p, C.c_size_t, C.c_int, C.c_int]
        L.oracle_brief.argtypes = [u8p, C.c_size_t, C.c_float, C.c_float, C.c_double, f32p, u8p]
        L.oracle_pyramid_build.restype = C.c_int
        L.oracle_pyramid_build.argtypes = [C.POINTER(_Pyramid), u8p, C.c_int, C.c_int, C.c_size_t, C.c_int, C.c_int, C.c_float]
        L.oracle_pyramid_free.argtypes = [C.POINTER(_Pyramid)]
        L.oracle_extract.restype = C.c_int
        L.oracle_extract.argtypes = [C.POINTER(_Pyramid), C.c_int, C.c_int, f32p, C.c_void_p, u8p, f64p, i32p]
        L.oracle_init_grid.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, i32p, i32p, i32p, C.c_int, i32p]
        L.oracle_find_features_in_area.restype = C.c_int
        L.oracle_find_features_in_area.argtypes = [C.c_void_p, i32p, i32p, C.c_int, C.c_int, f32p, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float,
                                                   C.c_int, C.c_int, C.c_int, i32p]
        L.oracle_search_in_area.argtypes = [C.c_void_p, u8p, C.c_int, i32p, i32p, C.c_int, C.c_int, f32p, C.c_float, C.c_float, C.c_void_p, u8p, C.c_int,
                                            u8p, i32p, i32p, f32p, i32p]
        L.oracle_verify_angle.restype = C.c_int
        L.oracle_verify_angle.argtypes = [C.c_int, i32p, i32p, f32p, C.c_void_p, C.c_void_p]
        L.oracle_bow_transform.restype = C.c_int
        L.oracle_bow_transform.argtypes = [C.c_void_p, u8p, C.c_int, C.c_int, i32p, f64p, i32p, i32p, i32p, i32p]
        L.oracle_search_by_bow.restype = C.c_int
        L.oracle_search_by_bow.argtypes = [i32p, i32p, i32p, C.c_int, u8p, u8p, i32p, i32p, i32p, C.c_int, u8p, u8p, i32p, i32p, i32p, f32p, i32p]
        L.oracle_serialize_keyframe.restype = C.c_size_t
        L.oracle_serialize_keyframe.argtypes = [C.c_void_p, u8p, f64p, f64p, C.c_int, C.c_uint64, C.c_float, C.c_float, C.c_float, C.c_float, f32p, C.c_int,
                                                u8p, C.c_size_t]
        L.oracle_search_by_stereo.restype = C.c_int
        L.oracle_search_by_stereo.argtypes = [C.POINTER(_Pyramid), C.POINTER(_Pyramid), C.c_void_p, u8p, C.c_int, C.c_void_p, u8p, C.c_int,
                                              C.c_float, C.c_float, f64p, f64p, i32p]
        L.oracle_rgbd_lookup.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_size_t, C.c_float, C.c_void_p, C.c_void_p, C.c_int, C.c_float,
                                         f64p, f64p]
        L.oracle_scale_factors.argtypes = [C.c_float, C.c_int, f32p]
        L.oracle_level_quotas.argtypes = [C.c_int, C.c_float, C.c_int, i32p]
        L.oracle_level_sizes.restype = C.c_int
        L.oracle_level_sizes.argtypes = [C.c_int, C.c_int, f32p, C.c_int, i32p, i32p]
        L.oracle_umax.argtypes = [i32p]
        L.oracle_fast9_arc_value.restype = C.c_int
        L.oracle_fast9_arc_value.argtypes = [u8p, C.c_size_t]
    return _lib


def default_pattern() -> np.ndarray:
    """The 256x4 BRIEF pattern as float32, parsed from include/orbx_pattern.h (the repo's built-in table)."""
    import re

    txt = open(os.path.join(HERE, "..", "include", "orbx_pattern.h")).read()
    body = txt[txt.index("{") + 1 : txt.index("};")]
    vals = [int(v) for v in re.findall(r"-?\d+", body)]
    assert len(vals) == 1024
    return np.asarray(vals, dtype=np.float32).reshape(256, 4)


def write_template_file(path: str, pattern: np.ndarray | None = None) -> str:
    """Write a template file in the reference's format (header line + 256 tab-separated rows)."""
    p = default_pattern() if pattern is None else pattern
    with open(path, "w") as f:
        f.write("x1  y1  x2  y2\n")
        for r in p:
            f.write("\t".join(str(int(v)) for v in r) + "\n")
    return path


# ----------------------------------------------------------------------------------------------------------------
def scale_factors(scale: float, n_levels: int) -> np.ndarray:
    sf = np.zeros(n_levels, np.float32)
    lib().oracle_scale_factors(scale, n_levels, _ptr(sf, f32p))
    return sf


def level_quotas(n_features: int, scale: float, n_levels: int) -> np.ndarray:
    q = np.zeros(n_levels, np.int32)
    lib().oracle_level_quotas(n_features, scale, n_levels, _ptr(q, i32p))
    return q


def level_sizes(w: int, h: int, scale: float, n_levels: int):
    sf = scale_factors(scale, n_levels)
    lw = np.zeros(n_levels, np.int32)
    lh = np.zeros(n_levels, np.int32)
    rc = lib().oracle_level_sizes(w, h, _ptr(sf, f32p), n_levels, _ptr(lw, i32p), _ptr(lh, i32p))
    return rc, lw, lh


def umax() -> np.ndarray:
    u = np.zeros(16, np.int32)
    lib().oracle_umax(_ptr(u, i32p))
    return u


def resize_linear(img: np.ndarray, dw: int, dh: int) -> np.ndarray:
    img = np.ascontiguousarray(img, np.uint8)
    out = np.empty((dh, dw), np.uint8)
    lib().oracle_resize_linear_u8(_ptr(img, u8p), img.shape[1], img.shape[0], img.strides[0], _ptr(out, u8p), dw, dh, dw)
    return out


def gaussian_blur7(img: np.ndarray) -> np.ndarray:
    img = np.ascontiguousarray(img, np.uint8)
    out = np.empty_like(img)
    lib().oracle_gaussian_blur7_u8(_ptr(img, u8p), img.shape[1], img.shape[0], img.strides[0], _ptr(out, u8p), out.strides[0])
    return out


def fast9_nms(img: np.ndarray, threshold: int):
    """-> (n,3) int32 array of (x, y, score), row-major order"""
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    cap = (h // 2 + 1) * (w // 2 + 1)
    xs, ys, sc = (np.zeros(cap, np.int32) for _ in range(3))
    n = lib().oracle_fast9_nms(_ptr(img, u8p), w, h, img.strides[0], threshold, _ptr(xs, i32p), _ptr(ys, i32p), _ptr(sc, i32p), cap)
    return np.stack([xs[:n], ys[:n], sc[:n]], 1)


def fast_cells(img: np.ndarray, ini_th: int = 20, min_th: int = 7):
    """-> ((n,3) int32 (x_roi, y_roi, score) in detection order, n_fallback_cells)"""
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    cap = h * w // 4 + 16
    xs, ys, sc = (np.zeros(cap, np.int32) for _ in range(3))
    nfb = C.c_int(0)
    n = lib().oracle_fast_cells(_ptr(img, u8p), w, h, img.strides[0], ini_th, min_th, _ptr(xs, i32p), _ptr(ys, i32p), _ptr(sc, i32p), cap, C.byref(nfb))
    if n < 0:
        raise ValueError("level too small for the 30-px cell grid")
    return np.stack([xs[:n], ys[:n], sc[:n]], 1), nfb.value


def quadtree_select(roi_w: int, roi_h: int, xs, ys, resp, need: int):
    xs = np.ascontiguousarray(xs, np.float32)
    ys = np.ascontiguousarray(ys, np.float32)
    resp = np.ascontiguousarray(resp, np.float32)
    out = np.zeros(max(need, 1) + 4, np.int32)
    pops = C.c_long(0)
    n = lib().oracle_quadtree_select(roi_w, roi_h, len(xs), _ptr(xs, f32p), _ptr(ys, f32p), _ptr(resp, f32p), need, _ptr(out, i32p), C.byref(pops))
    return out[:n].copy(), pops.value


def undistort_points(xy: np.ndarray, fx, fy, cx, cy, dist) -> np.ndarray:
    out = np.ascontiguousarray(xy, np.float32).copy()
    d = np.ascontiguousarray(dist, np.float32)
    lib().oracle_undistort_points(_ptr(out, f32p), out.shape[0], fx, fy, cx, cy, _ptr(d, f32p), len(d))
    return out


def ic_angle(img: np.ndarray, x: int, y: int) -> float:
    img = np.ascontiguousarray(img, np.uint8)
    return lib().oracle_ic_angle(_ptr(img, u8p), img.strides[0], x, y)


def brief(blurred: np.ndarray, px: float, py: float, theta: float, pattern: np.ndarray) -> np.ndarray:
    blurred = np.ascontiguousarray(blurred, np.uint8)
    pat = np.ascontiguousarray(pattern, np.float32)
    d = np.zeros(32, np.uint8)
    lib().oracle_brief(_ptr(blurred, u8p), blurred.strides[0], px, py, theta, _ptr(pat, f32p), _ptr(d, u8p))
    return d


class Pyramid:
    """ORBExtractor constructor state (pyramid + blurred pyramid + per-level tables)."""

    def __init__(self, img: np.ndarray, n_features: int, n_levels: int, scale: float):
        img = np.ascontiguousarray(img, np.uint8)
        self._p = _Pyramid()
        rc = lib().oracle_pyramid_build(C.byref(self._p), _ptr(img, u8p), img.shape[1], img.shape[0], img.strides[0], n_features, n_levels, scale)
        if rc != 0:
            raise ValueError("ImageSizeError" if rc == -1 else "bad level count")
        self.n_levels = n_levels
        self.w = list(self._p.w[:n_levels])
        self.h = list(self._p.h[:n_levels])
        self.sf = np.array(self._p.sf[:n_levels], np.float32)
        self.quota = list(self._p.quota[:n_levels])

    def level(self, l: int) -> np.ndarray:
        return np.ctypeslib.as_array(self._p.img[l], shape=(self.h[l], self.w[l])).copy()

    def blurred(self, l: int) -> np.ndarray:
        return np.ctypeslib.as_array(self._p.blur[l], shape=(self.h[l], self.w[l])).copy()

    def __del__(self):
        try:
            lib().oracle_pyramid_free(C.byref(self._p))
        except Exception:
            pass


@dataclass
class Extracted:
    kps: np.ndarray  # KP_DTYPE
    desc: np.ndarray  # (n,32) u8
    theta: np.ndarray  # (n,) f64 radians
    level_counts: np.ndarray
    pyr: Pyramid


def extract(img: np.ndarray, n_features=2000, n_levels=8, scale=1.2, ini_th=20, min_th=7, pattern=None) -> Extracted:
    pyr = Pyramid(img, n_features, n_levels, scale)
    pat = np.ascontiguousarray(default_pattern() if pattern is None else pattern, np.float32)
    kps = np.zeros(n_features + 8, KP_DTYPE)
    desc = np.zeros((n_features + 8, 32), np.uint8)
    th = np.zeros(n_features + 8, np.float64)
    lc = np.zeros(n_levels, np.int32)
    n = lib().oracle_extract(C.byref(pyr._p), ini_th, min_th, _ptr(pat, f32p), kps.ctypes.data, _ptr(desc, u8p), _ptr(th, f64p), _ptr(lc, i32p))
    if n < 0:
        raise ValueError("level too small for the cell grid")
    return Extracted(kps[:n].copy(), desc[:n].copy(), th[:n].copy(), lc, pyr)


def search_by_stereo(left: Extracted, right: Extracted, fx: float, bf: float, kps_left_undist=None):
    kl = np.ascontiguousarray(left.kps if kps_left_undist is None else kps_left_undist)
    kr = np.ascontiguousarray(right.kps)
    dl = np.ascontiguousarray(left.desc)
    dr = np.ascontiguousarray(right.desc)
    nl, nr = len(kl), len(kr)
    ur = np.zeros(max(nl, 1), np.float64)
    dp = np.zeros(max(nl, 1), np.float64)
    mi = np.zeros(max(nl, 1), np.int32)
    nm = lib().oracle_search_by_stereo(C.byref(left.pyr._p), C.byref(right.pyr._p), kl.ctypes.data, _ptr(dl, u8p), nl, kr.ctypes.data, _ptr(dr, u8p), nr,
                                       fx, bf, _ptr(ur, f64p), _ptr(dp, f64p), _ptr(mi, i32p))
    return nm, ur[:nl], dp[:nl], mi[:nl]


def rgbd_lookup(depth_img: np.ndarray, depth_scale: float, kps_raw: np.ndarray, kps_undist: np.ndarray, bf: float):
    is_float = depth_img.dtype == np.float32
    d = np.ascontiguousarray(depth_img, np.float32 if is_float else np.uint16)
    n = len(kps_raw)
    ur = np.zeros(max(n, 1), np.float64)
    dp = np.zeros(max(n, 1), np.float64)
    kr = np.ascontiguousarray(kps_raw)
    ku = np.ascontiguousarray(kps_undist)
    lib().oracle_rgbd_lookup(d.ctypes.data, int(is_float), d.shape[1], d.shape[0], d.strides[0] // d.itemsize, depth_scale, kr.ctypes.data, ku.ctypes.data,
                             n, bf, _ptr(ur, f64p), _ptr(dp, f64p))
    return ur[:n], dp[:n]


def init_grid(kps: np.ndarray, min_u, min_v, max_u, max_v):
    """VirtualFrame::initGrid (src/Frame.cc:53-69), numpy restatement: rows = cvCeil((float)(maxV - minV) / 48), cols =
    cvCeil((float)(maxU - minU) / 64); keypoint i goes to cell (cvFloor(pt.y / 48), cvFloor(pt.x / 64)), float division,
    in keypoint order.  (Frame.cc cannot be compiled in isolation -- DBoW3 / KeyFrame dependencies -- so this piece is a
    restatement only.)  -> list[rows][cols] of index arrays"""
    f32 = np.float32
    rows = int(np.ceil(f32(f32(max_v) - f32(min_v)) / f32(48)))
    cols = int(np.ceil(f32(f32(max_u) - f32(min_u)) / f32(64)))
    grid = [[[] for _ in range(cols)] for _ in range(rows)]
    r = np.floor(kps["y"].astype(f32) / f32(48)).astype(np.int64)
    c = np.floor(kps["x"].astype(f32) / f32(64)).astype(np.int64)
    for i in range(len(kps)):
        if 0 <= r[i] < rows and 0 <= c[i] < cols:  # anything else indexes mGrids out of range in the reference
            grid[r[i]][c[i]].append(i)
    return [[np.asarray(cell, np.int32) for cell in row] for row in grid]


def grid_csr(kps, min_u, min_v, max_u, max_v):
    """oracle_init_grid (VirtualFrame::initGrid, src/Frame.cc:53-69) -> (rows, cols, start[rows*cols+1], entries)"""
    kps = np.ascontiguousarray(kps, KP_DTYPE)
    rows, cols = C.c_int(0), C.c_int(0)
    cap = 1 << 16
    start = np.zeros(cap + 1, np.int32)
    entries = np.zeros(max(len(kps), 1), np.int32)
    lib().oracle_init_grid(kps.ctypes.data, len(kps), min_u, min_v, max_u, max_v, C.byref(rows), C.byref(cols), _ptr(start, i32p), cap, _ptr(entries, i32p))
    nc = rows.value * cols.value
    return rows.value, cols.value, start[: nc + 1].copy(), entries[: start[nc]].copy()


def search_in_area(kps, desc, bounds, sf, queries, q_desc, exclude=None):
    """oracle_search_in_area.  bounds = (min_u, min_v, max_u, max_v) -> dict(best_idx, best_dist, ratio, n_cand)"""
    kps = np.ascontiguousarray(kps, KP_DTYPE)
    desc = np.ascontiguousarray(desc, np.uint8)
    queries = np.ascontiguousarray(queries, AREA_QUERY_DTYPE)
    q_desc = np.ascontiguousarray(q_desc, np.uint8)
    sf = np.ascontiguousarray(sf, np.float32)
    ex = None if exclude is None else np.ascontiguousarray(exclude, np.uint8)
    rows, cols, start, entries = grid_csr(kps, *bounds)
    n = len(queries)
    bi, bd, nc = np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros(n, np.int32)
    ra = np.zeros(n, np.float32)
    lib().oracle_search_in_area(kps.ctypes.data, _ptr(desc, u8p), len(kps), _ptr(start, i32p), _ptr(entries, i32p), rows, cols, _ptr(sf, f32p), bounds[2],
                                bounds[3], queries.ctypes.data, _ptr(q_desc, u8p), n, _ptr(ex, u8p), _ptr(bi, i32p), _ptr(bd, i32p), _ptr(ra, f32p),
                                _ptr(nc, i32p))
    return dict(best_idx=bi, best_dist=bd, ratio=ra, n_cand=nc)


def verify_angle(query_idx, train_idx, distance, kps1, kps2):
    """oracle_verify_angle -> (query_idx, train_idx, distance) of the surviving matches"""
    qi = np.ascontiguousarray(query_idx, np.int32).copy()
    ti = np.ascontiguousarray(train_idx, np.int32).copy()
    di = np.ascontiguousarray(distance, np.float32).copy()
    k1, k2 = np.ascontiguousarray(kps1, KP_DTYPE), np.ascontiguousarray(kps2, KP_DTYPE)
    m = lib().oracle_verify_angle(len(qi), _ptr(qi, i32p), _ptr(ti, i32p), _ptr(di, f32p), k1.ctypes.data, k2.ctypes.data)
    return qi[:m], ti[:m], di[:m]


def serialize_keyframe(kps, desc, u_right, depth, kf_id, bounds, pose_rt=None, with_map_points=True) -> bytes:
    """oracle_serialize_keyframe: orbslam2.KeyFrameData bytes (proto/Keyframe.proto:45-64, writer src/KeyFrame.cc:553-647).
    bounds = (min_u, min_v, max_u, max_v) as returned by grid_info"""
    kps = np.ascontiguousarray(kps, KP_DTYPE)
    desc = np.ascontiguousarray(desc, np.uint8)
    ur, dp = np.ascontiguousarray(u_right, np.float64), np.ascontiguousarray(depth, np.float64)
    n = len(kps)
    pose = None if pose_rt is None else np.ascontiguousarray(pose_rt, np.float32)
    out = np.zeros(256 + 96 * n, np.uint8)
    m = lib().oracle_serialize_keyframe(kps.ctypes.data, _ptr(desc, u8p), _ptr(ur, f64p), _ptr(dp, f64p), n, kf_id, bounds[2], bounds[3], bounds[0], bounds[1],
                                        _ptr(pose, f32p), int(with_map_points), _ptr(out, u8p), out.size)
    assert m > 0
    return out[:m].tobytes()


class _Vocab(C.Structure):
    _fields_ = [("k", C.c_int32), ("L", C.c_int32), ("n_nodes", C.c_int32), ("child_start", i32p), ("child_ids", i32p), ("desc", u8p), ("weight", f64p),
                ("word_id", i32p)]


class Vocabulary:
    """A DBoW3 tree built from ORB-SLAM2 text-format records (record i = node i + 1; ids and word ids in record order, as
    DBoW3's text loader assigns them) for oracle_bow_transform."""

    def __init__(self, k, L, parent, is_leaf, desc, weight):
        n = len(parent) + 1
        self.k, self.L, self.n_nodes = int(k), int(L), n
        par = np.concatenate([[-1], np.asarray(parent, np.int64)])
        order = np.argsort(par[1:], kind="stable") + 1  # children grouped by parent, ascending id inside a group
        self.child_ids = np.ascontiguousarray(order, np.int32)
        cnt = np.bincount(par[1:], minlength=n)
        self.child_start = np.zeros(n + 1, np.int32)
        self.child_start[1:] = np.cumsum(cnt)
        self.desc = np.zeros((n, 32), np.uint8)
        self.desc[1:] = desc
        self.weight = np.zeros(n, np.float64)
        self.weight[1:] = weight
        self.word_id = np.full(n, -1, np.int32)
        leaf = np.concatenate([[0], np.asarray(is_leaf)]) > 0
        self.word_id[leaf] = np.arange(int(leaf.sum()), dtype=np.int32)
        self._c = _Vocab(self.k, self.L, n, _ptr(self.child_start, i32p), _ptr(self.child_ids, i32p), _ptr(self.desc, u8p), _ptr(self.weight, f64p),
                         _ptr(self.word_id, i32p))


def bow_transform(voc: Vocabulary, desc, levelsup: int = 4):
    """oracle_bow_transform -> dict(bow_ids, bow_vals, fv_nodes, fv_start, fv_feats)"""
    desc = np.ascontiguousarray(desc, np.uint8)
    n = len(desc)
    ids, vals = np.zeros(n + 1, np.int32), np.zeros(n + 1, np.float64)
    fn_, fs, ff = np.zeros(n + 1, np.int32), np.zeros(n + 2, np.int32), np.zeros(n + 1, np.int32)
    cnt = C.c_int(0)
    m = lib().oracle_bow_transform(C.byref(voc._c), _ptr(desc, u8p), n, levelsup, _ptr(ids, i32p), _ptr(vals, f64p), _ptr(fn_, i32p), _ptr(fs, i32p),
                                   _ptr(ff, i32p), C.byref(cnt))
    c = cnt.value
    return dict(bow_ids=ids[:m].copy(), bow_vals=vals[:m].copy(), fv_nodes=fn_[:c].copy(), fv_start=fs[: c + 1].copy(), fv_feats=ff[: fs[c]].copy())


def search_by_bow(f_bow: dict, f_desc, k_bow: dict, k_desc, frame_cand_ok=None, kf_query_ok=None):
    """oracle_search_by_bow on two bow_transform() results -> dict(kf_idx, best_idx, best_dist, ratio, n_cand), one row per
    visited keyframe feature with candidates"""
    fd, kd = np.ascontiguousarray(f_desc, np.uint8), np.ascontiguousarray(k_desc, np.uint8)
    fn, fs, ff = (np.ascontiguousarray(f_bow[k], np.int32) for k in ("fv_nodes", "fv_start", "fv_feats"))
    kn, ks, kf = (np.ascontiguousarray(k_bow[k], np.int32) for k in ("fv_nodes", "fv_start", "fv_feats"))
    fm = None if frame_cand_ok is None else np.ascontiguousarray(frame_cand_ok, np.uint8)
    km = None if kf_query_ok is None else np.ascontiguousarray(kf_query_ok, np.uint8)
    cap = max(len(kf), 1)
    ki, bi, bd, nc = (np.zeros(cap, np.int32) for _ in range(4))
    ra = np.zeros(cap, np.float32)
    n = lib().oracle_search_by_bow(_ptr(fn, i32p), _ptr(fs, i32p), _ptr(ff, i32p), len(fn), _ptr(fd, u8p), _ptr(fm, u8p), _ptr(kn, i32p), _ptr(ks, i32p),
                                   _ptr(kf, i32p), len(kn), _ptr(kd, u8p), _ptr(km, u8p), _ptr(ki, i32p), _ptr(bi, i32p), _ptr(bd, i32p), _ptr(ra, f32p),
                                   _ptr(nc, i32p))
    return dict(kf_idx=ki[:n].copy(), best_idx=bi[:n].copy(), best_dist=bd[:n].copy(), ratio=ra[:n].copy(), n_cand=nc[:n].copy())


# ----------------------------------------------------------------------------------------------------------------
# the compiled reference (oracle/_ref/libref.so)
_ref = None


def have_ref() -> bool:
    return os.path.exists(REF_SO)


def ref():
    global _ref
    if _ref is None:
        if not have_ref():
            build()
        R = C.CDLL(REF_SO)
        R.ref_extract.restype = C.c_int
        R.ref_extract.argtypes = [u8p, C.c_int, C.c_int, C.c_size_t, C.c_int, C.c_int, C.c_float, C.c_char_p, C.c_int, C.c_int, C.c_void_p, u8p, C.c_int,
                                  u8p, i32p, i32p, f32p]
        R.ref_blurred_pyramid.restype = C.c_int
        R.ref_blurred_pyramid.argtypes = [u8p, C.c_int, C.c_int, C.c_size_t, C.c_int, C.c_int, C.c_float, C.c_char_p, u8p]
        R.ref_quadtree.restype = C.c_int
        R.ref_quadtree.argtypes = [C.c_int, C.c_int, C.c_int, f32p, f32p, f32p, C.c_int, i32p]
        R.ref_stereo.restype = C.c_int
        R.ref_stereo.argtypes = [u8p, u8p, C.c_int, C.c_int, C.c_size_t, C.c_int, C.c_int, C.c_float, C.c_char_p, C.c_int, C.c_int, C.c_void_p, u8p, i32p,
                                 C.c_void_p, u8p, i32p, f64p, f64p, C.c_int]
        R.ref_set_camera.argtypes = [C.c_float] * 5 + [f32p]
        R.ref_get_bf.restype = C.c_float
        R.ref_undistort.argtypes = [f32p, C.c_int]
        R.ref_init_grid.restype = C.c_int
        R.ref_init_grid.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, i32p, i32p, i32p, C.c_int, i32p]
        R.ref_search_in_area.argtypes = [C.c_void_p, u8p, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, f32p, C.c_int, C.c_void_p, u8p, C.c_int, u8p,
                                         i32p, i32p, f32p, i32p]
        R.ref_verify_angle.restype = C.c_int
        R.ref_verify_angle.argtypes = [C.c_int, i32p, i32p, f32p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        R.ref_rgbd.restype = C.c_int
        R.ref_rgbd.argtypes = [u8p, C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_int, C.c_size_t, C.c_float, C.c_int, C.c_int, C.c_float, C.c_char_p, C.c_int,
                               C.c_int, C.c_void_p, u8p, f64p, f64p, C.c_int]
        R.ref_bench_stereo.restype = C.c_double
        R.ref_bench_stereo.argtypes = [u8p, u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                       C.POINTER(C.c_long)]
        _ref = R
    return _ref


def ref_reset():
    ref().ref_reset()


def ref_set_camera(fx, fy, cx, cy, bl, dist5=None):
    d = np.zeros(5, np.float32) if dist5 is None else np.ascontiguousarray(dist5, np.float32)
    ref().ref_set_camera(fx, fy, cx, cy, bl, _ptr(d, f32p))
    return ref().ref_get_bf()


def ref_extract(img, template_path, n_features=2000, n_levels=8, scale=1.2, ini_th=20, min_th=7, want_pyramid=False):
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    cap = n_features + 8
    kps = np.zeros(cap, KP_DTYPE)
    desc = np.zeros((cap, 32), np.uint8)
    lw = np.zeros(n_levels, np.int32)
    lh = np.zeros(n_levels, np.int32)
    sf = np.zeros(n_levels, np.float32)
    pyr = np.zeros(4 * w * h, np.uint8) if want_pyramid else None
    n = ref().ref_extract(_ptr(img, u8p), w, h, img.strides[0], n_features, n_levels, scale, template_path.encode(), ini_th, min_th, kps.ctypes.data,
                          _ptr(desc, u8p), cap, _ptr(pyr, u8p), _ptr(lw, i32p), _ptr(lh, i32p), _ptr(sf, f32p))
    if n < 0:
        return n, None, None, None
    levels = None
    if want_pyramid:
        levels, off = [], 0
        for l in range(n_levels):
            levels.append(pyr[off : off + lw[l] * lh[l]].reshape(lh[l], lw[l]).copy())
            off += lw[l] * lh[l]
    return n, kps[:n].copy(), desc[:n].copy(), dict(levels=levels, lw=lw, lh=lh, sf=sf)


def ref_blurred(img, template_path, n_features, n_levels, scale, lw, lh):
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    out = np.zeros(int(np.sum(lw.astype(np.int64) * lh)), np.uint8)
    rc = ref().ref_blurred_pyramid(_ptr(img, u8p), w, h, img.strides[0], n_features, n_levels, scale, template_path.encode(), _ptr(out, u8p))
    assert rc == 0
    res, off = [], 0
    for l in range(n_levels):
        res.append(out[off : off + lw[l] * lh[l]].reshape(lh[l], lw[l]))
        off += lw[l] * lh[l]
    return res


def ref_quadtree(roi_w, roi_h, xs, ys, resp, need):
    xs = np.ascontiguousarray(xs, np.float32)
    ys = np.ascontiguousarray(ys, np.float32)
    resp = np.ascontiguousarray(resp, np.float32)
    out = np.zeros(max(need, 1) + 4, np.int32)
    n = ref().ref_quadtree(roi_w, roi_h, len(xs), _ptr(xs, f32p), _ptr(ys, f32p), _ptr(resp, f32p), need, _ptr(out, i32p))
    return out[:n].copy()


def ref_stereo(left, right, template_path, n_features=2000, n_levels=8, scale=1.2, ini_th=20, min_th=7):
    left = np.ascontiguousarray(left, np.uint8)
    right = np.ascontiguousarray(right, np.uint8)
    assert left.shape == right.shape and left.strides == right.strides
    h, w = left.shape
    cap = n_features + 8
    kl, kr = np.zeros(cap, KP_DTYPE), np.zeros(cap, KP_DTYPE)
    dl, dr = np.zeros((cap, 32), np.uint8), np.zeros((cap, 32), np.uint8)
    ur, dp = np.zeros(cap, np.float64), np.zeros(cap, np.float64)
    nl, nr = C.c_int(0), C.c_int(0)
    nm = ref().ref_stereo(_ptr(left, u8p), _ptr(right, u8p), w, h, left.strides[0], n_features, n_levels, scale, template_path.encode(), ini_th, min_th,
                          kl.ctypes.data, _ptr(dl, u8p), C.byref(nl), kr.ctypes.data, _ptr(dr, u8p), C.byref(nr), _ptr(ur, f64p), _ptr(dp, f64p), cap)
    if nm < 0:
        return dict(status=nm)
    a, b = nl.value, nr.value
    return dict(status=0, n_matches=nm, kl=kl[:a].copy(), dl=dl[:a].copy(), kr=kr[:b].copy(), dr=dr[:b].copy(), u_right=ur[:a].copy(), depth=dp[:a].copy())


def ref_rgbd(gray, depth_img, depth_scale, template_path, n_features=1000, n_levels=8, scale=1.2, ini_th=20, min_th=7):
    """The reference's OWN RGB-D Frame ctor body (src/Frame.cc:130-158) -> (kps undistorted, desc, u_right, depth)"""
    g = np.ascontiguousarray(gray, np.uint8)
    is_float = depth_img.dtype == np.float32
    d = np.ascontiguousarray(depth_img, np.float32 if is_float else np.uint16)
    h, w = g.shape
    cap = n_features + 8
    kps, desc = np.zeros(cap, KP_DTYPE), np.zeros((cap, 32), np.uint8)
    ur, dp = np.zeros(cap, np.float64), np.zeros(cap, np.float64)
    n = ref().ref_rgbd(_ptr(g, u8p), w, h, g.strides[0], d.ctypes.data, int(is_float), d.strides[0], depth_scale, n_features, n_levels, scale,
                       template_path.encode(), ini_th, min_th, kps.ctypes.data, _ptr(desc, u8p), _ptr(ur, f64p), _ptr(dp, f64p), cap)
    if n < 0:
        raise RuntimeError(f"ref_rgbd failed ({n})")
    return kps[:n], desc[:n], ur[:n], dp[:n]


def ref_undistort(xy):
    out = np.ascontiguousarray(xy, np.float32).copy()
    ref().ref_undistort(_ptr(out, f32p), out.shape[0])
    return out


def ref_grid_csr(kps, min_u, min_v, max_u, max_v):
    kps = np.ascontiguousarray(kps, KP_DTYPE)
    rows, cols = C.c_int(0), C.c_int(0)
    cap = 1 << 16
    start = np.zeros(cap + 1, np.int32)
    entries = np.zeros(max(len(kps), 1), np.int32)
    nc = ref().ref_init_grid(kps.ctypes.data, len(kps), min_u, min_v, max_u, max_v, C.byref(rows), C.byref(cols), _ptr(start, i32p), cap, _ptr(entries, i32p))
    assert nc >= 0
    return rows.value, cols.value, start[: nc + 1].copy(), entries[: start[nc]].copy()


def ref_search_in_area(kps, desc, bounds, sf, queries, q_desc, exclude=None):
    kps = np.ascontiguousarray(kps, KP_DTYPE)
    desc = np.ascontiguousarray(desc, np.uint8)
    queries = np.ascontiguousarray(queries, AREA_QUERY_DTYPE)
    q_desc = np.ascontiguousarray(q_desc, np.uint8)
    sf = np.ascontiguousarray(sf, np.float32)
    ex = None if exclude is None else np.ascontiguousarray(exclude, np.uint8)
    n = len(queries)
    bi, bd, nc = np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros(n, np.int32)
    ra = np.zeros(n, np.float32)
    ref().ref_search_in_area(kps.ctypes.data, _ptr(desc, u8p), len(kps), bounds[0], bounds[1], bounds[2], bounds[3], _ptr(sf, f32p), len(sf),
                             queries.ctypes.data, _ptr(q_desc, u8p), n, _ptr(ex, u8p), _ptr(bi, i32p), _ptr(bd, i32p), _ptr(ra, f32p), _ptr(nc, i32p))
    return dict(best_idx=bi, best_dist=bd, ratio=ra, n_cand=nc)


def ref_verify_angle(query_idx, train_idx, distance, kps1, kps2):
    qi = np.ascontiguousarray(query_idx, np.int32).copy()
    ti = np.ascontiguousarray(train_idx, np.int32).copy()
    di = np.ascontiguousarray(distance, np.float32).copy()
    k1, k2 = np.ascontiguousarray(kps1, KP_DTYPE), np.ascontiguousarray(kps2, KP_DTYPE)
    m = ref().ref_verify_angle(len(qi), _ptr(qi, i32p), _ptr(ti, i32p), _ptr(di, f32p), k1.ctypes.data, len(k1), k2.ctypes.data, len(k2))
    return qi[:m], ti[:m], di[:m]


def ref_bench_stereo(left_pool, right_pool, template_path, n_frames, workers, n_features=2000, n_levels=8, scale=1.2, ini_th=20, min_th=7):
    """-> (seconds, total matches).  left_pool/right_pool: (pool, H, W) uint8, dense."""
    lp = np.ascontiguousarray(left_pool, np.uint8)
    rp = np.ascontiguousarray(right_pool, np.uint8)
    pool, h, w = lp.shape
    m = C.c_long(0)
    secs = ref().ref_bench_stereo(_ptr(lp, u8p), _ptr(rp, u8p), pool, w, h, n_features, n_levels, scale, template_path.encode(), ini_th, min_th, n_frames,
                                  workers, C.byref(m))
    return secs, m.value
