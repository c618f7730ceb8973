"""The C++17 host mirror (include/orbx/orb_slam2_shim.hpp): same class names / signatures as the reference, driven by
tests/cpp/shim_main.cpp the way the reference's call sites do, compiled against the minimal cv:: types of oracle/stub
(this image has no C++ OpenCV) and linked with liborbx.so."""
import os
import struct
import subprocess

import numpy as np
import pytest

from orb_slam2_ros2_b200 import api, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def shim_binary(tmp_path_factory):
    api.load_library()
    out = str(tmp_path_factory.mktemp("shim") / "shim_main")
    libdir = os.path.join(ROOT, "orb_slam2_ros2_b200")
    cmd = ["/usr/bin/g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "oracle", "stub"),
           os.path.join(ROOT, "tests", "cpp", "shim_main.cpp"), "-o", out, "-L", libdir, "-lorbx", f"-Wl,-rpath,{libdir}", "-pthread"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return out


def test_shim_compiles_and_fails_loudly_without_gpu(shim_binary, template_path):
    import torch

    r = subprocess.run([shim_binary, "errors", template_path], capture_output=True, text=True)
    if torch.cuda.is_available():
        assert r.returncode == 0 and "errors ok=2" in r.stdout
    else:
        assert r.returncode == 2 and "no CUDA device" in r.stderr  # no CPU fallback


class _Reader:
    def __init__(self, path):
        self.b = open(path, "rb").read()
        self.o = 0

    def i32(self):
        v = struct.unpack_from("<i", self.b, self.o)[0]
        self.o += 4
        return v

    def f32(self):
        v = struct.unpack_from("<f", self.b, self.o)[0]
        self.o += 4
        return v

    def arr(self, dtype, n):
        a = np.frombuffer(self.b, dtype, n, self.o).copy()
        self.o += a.nbytes
        return a

    def kps_desc(self):
        n = self.i32()
        return self.arr(api.KP_DTYPE, n), self.arr(np.uint8, 32 * n).reshape(n, 32)


def _same(a, b, da, db):
    assert len(a) == len(b)
    for f in ("x", "y", "size", "response", "octave", "class_id"):
        assert np.array_equal(a[f], b[f]), f
    assert np.abs(a["angle"] - b["angle"]).max(initial=0) <= np.degrees(1e-4)
    assert int(np.unpackbits(da ^ db).sum()) <= 2


@pytest.mark.gpu
def test_shim_stereo_matches_oracle(shim_binary, template_path, oracle, tmp_path):
    c = synth.KITTI
    left, right = synth.synth_stereo_pair(c["height"], c["width"], 3, 17)
    left.tofile(tmp_path / "l.raw")
    right.tofile(tmp_path / "r.raw")
    out = tmp_path / "out.bin"
    voc = synth.synth_vocabulary(8, 3, 5)
    voc_path = synth.write_vocabulary_text(str(tmp_path / "voc.txt"), voc)
    r = subprocess.run([shim_binary, "stereo", str(c["width"]), str(c["height"]), "2000", "8", "1.2", template_path, str(tmp_path / "l.raw"),
                        str(tmp_path / "r.raw"), str(out), voc_path], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    rd = _Reader(out)
    el, er = oracle.extract(left), oracle.extract(right)
    k, d = rd.kps_desc()  # ORBExtractor::extract on the left image
    _same(k, el.kps, d, el.desc)
    assert rd.i32() == 8 and rd.i32() == el.pyr.w[7] and rd.f32() == el.pyr.sf[7]
    kl, dl = rd.kps_desc()
    kr, dr = rd.kps_desc()
    _same(kl, el.kps, dl, el.desc)
    _same(kr, er.kps, dr, er.desc)
    n_matches = rd.i32()
    ur, dp = rd.arr(np.float64, len(kl)), rd.arr(np.float64, len(kl))
    cam = api.Camera()
    nm, our, odp, _ = oracle.search_by_stereo(el, er, np.float32(cam.fx), cam.bf)
    assert n_matches == nm and np.abs(ur - our).max() <= 1e-3 and np.abs(dp - odp).max() <= 1e-3 * np.abs(odp).max()
    assert (rd.i32(), rd.i32(), rd.i32(), rd.i32()) == (8, 20, len(kl), 1)  # mGrids: 8 x 20 cells, all keypoints, ascending
    # ORBMatcher::searchByProjection (constant-velocity form) + verifyAngle against the oracle's composition of the same steps
    nm = rd.i32()
    got = rd.arr(np.dtype([("q", "<i4"), ("t", "<i4"), ("d", "<f4")]), nm)
    q = np.zeros(len(kl), oracle.AREA_QUERY_DTYPE)
    q["x"], q["y"], q["octave"], q["radius"] = kl["x"] + np.float32(2), kl["y"] - np.float32(1), kl["octave"], 15
    q["min_level"], q["max_level"] = np.maximum(0, kl["octave"] - 1), np.minimum(kl["octave"] + 1, 7)
    has1 = np.zeros(len(kl), np.uint8)
    has1[::3] = 1
    e = oracle.search_in_area(kl, dl, (0.0, 0.0, float(c["width"]), float(c["height"])), el.pyr.sf, q, dl, has1)
    ok = (e["n_cand"] > 0) & (e["ratio"] < np.float32(0.6)) & (e["best_dist"] < 50)
    assert nm == int(ok.sum()) and nm > 500
    assert np.array_equal(got["q"], e["best_idx"][ok]) and np.array_equal(got["t"], np.nonzero(ok)[0]) and np.array_equal(got["d"], e["best_dist"][ok])
    k2 = kl.copy()
    vq, vt, vd = oracle.verify_angle(got["q"], got["t"], got["d"], kl, k2)
    nv = rd.i32()
    gv = rd.arr(np.dtype([("q", "<i4"), ("t", "<i4"), ("d", "<f4")]), nv)
    assert nv == len(vq) and np.array_equal(gv["q"], vq) and np.array_equal(gv["t"], vt) and np.array_equal(gv["d"], vd)
    # Frame::serializeKeyFrameData == the oracle's KeyFrameData bytes for the same frame
    rec = bytes(rd.arr(np.uint8, rd.i32()))
    assert rec == oracle.serialize_keyframe(kl, dl, ur, dp, 42, (0.0, 0.0, float(c["width"]), float(c["height"])), None, True)
    # Frame::computeBow == the oracle's DBoW3 restatement
    OV = oracle.Vocabulary(**voc)
    e = oracle.bow_transform(OV, dl, 4)
    assert rd.i32() == int((OV.word_id >= 0).sum())
    nb = rd.i32()
    bv = rd.arr(np.dtype([("id", "<i4"), ("v", "<f8")]), nb)
    assert np.array_equal(bv["id"], e["bow_ids"]) and np.array_equal(bv["v"], e["bow_vals"])
    nf = rd.i32()
    assert nf == len(e["fv_nodes"])
    for j in range(nf):
        assert rd.i32() == e["fv_nodes"][j]
        m = rd.i32()
        assert np.array_equal(rd.arr(np.int32, m), e["fv_feats"][e["fv_start"][j] : e["fv_start"][j + 1]])
    # ORBMatcher::searchByBow (frame against itself, loop mode) == the oracle's composition + verifyAngle
    sb = oracle.search_by_bow(e, dl, e, dl)
    keep = ~((sb["best_dist"] > 50) | (sb["ratio"] > np.float32(0.6)))
    vq, vt, vd = oracle.verify_angle(sb["best_idx"][keep], sb["kf_idx"][keep], sb["best_dist"][keep].astype(np.float32), kl, kl)
    nbm = rd.i32()
    gm = rd.arr(np.dtype([("q", "<i4"), ("t", "<i4"), ("d", "<f4")]), nbm)
    assert nbm == len(vq) and nbm > 100 and np.array_equal(gm["q"], vq) and np.array_equal(gm["t"], vt) and np.array_equal(gm["d"], vd)


@pytest.mark.gpu
def test_shim_rgbd_matches_oracle(shim_binary, template_path, oracle, tmp_path):
    c = synth.TUM
    gray = synth.synth_image(c["height"], c["width"], 5)
    depth = synth.synth_depth_u16(c["height"], c["width"], 5, c["depth_scale"])
    gray.tofile(tmp_path / "g.raw")
    depth.tofile(tmp_path / "d.raw")
    out = tmp_path / "out.bin"
    r = subprocess.run([shim_binary, "rgbd", str(c["width"]), str(c["height"]), "1000", "8", "1.2", template_path, str(tmp_path / "g.raw"),
                        str(tmp_path / "d.raw"), str(out), str(c["depth_scale"])], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    rd = _Reader(out)
    k, d = rd.kps_desc()
    n_depth = rd.i32()
    ur, dp = rd.arr(np.float64, len(k)), rd.arr(np.float64, len(k))
    e = oracle.extract(gray, 1000)
    xy = oracle.undistort_points(np.stack([e.kps["x"], e.kps["y"]], 1), c["fx"], c["fy"], c["cx"], c["cy"], np.array(c["dist"], np.float32))
    ku = e.kps.copy()
    ku["x"], ku["y"] = xy[:, 0], xy[:, 1]
    assert np.abs(k["x"] - ku["x"]).max() <= 1e-3 and np.abs(k["y"] - ku["y"]).max() <= 1e-3
    assert np.array_equal(d, e.desc)
    our, odp = oracle.rgbd_lookup(depth, c["depth_scale"], e.kps, ku, np.float32(c["fx"]) * np.float32(c["bl"]))
    assert np.array_equal(dp, odp) and np.abs(ur - our).max() <= 1e-3 and n_depth == int((odp > 0).sum())


@pytest.mark.gpu
def test_shim_two_threads_same_configuration(shim_binary, template_path, oracle, tmp_path):
    """src/Frame.cc:100-105: two extractors of one configuration extract on two std::threads; each thread gets its own
    context from the shim's cache, results equal the oracle, getPyramid() works before / long after extract(), and a Frame
    whose device state has been replaced refuses device-side queries."""
    c = synth.KITTI
    a, b = synth.synth_image(c["height"], c["width"], 21), synth.synth_image(c["height"], c["width"], 22)
    a.tofile(tmp_path / "a.raw")
    b.tofile(tmp_path / "b.raw")
    out = tmp_path / "out.bin"
    r = subprocess.run([shim_binary, "threads", str(c["width"]), str(c["height"]), "1000", "8", "1.2", template_path, str(tmp_path / "a.raw"),
                        str(tmp_path / "b.raw"), str(out)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    rd = _Reader(out)
    ea, eb = oracle.extract(a, 1000), oracle.extract(b, 1000)
    ka, da = rd.kps_desc()
    kb, db = rd.kps_desc()
    _same(ka, ea.kps, da, ea.desc)
    _same(kb, eb.kps, db, eb.desc)
    for e in (ea, eb, eb):
        assert rd.i32() == 8
        w, h = rd.i32(), rd.i32()
        assert (w, h) == (e.pyr.w[7], e.pyr.h[7])
        assert np.array_equal(rd.arr(np.uint8, w * h).reshape(h, w), e.pyr.level(7))
    assert (rd.i32(), rd.i32(), rd.i32()) == (8, 1, 8)  # f1's grid, f1 refused after f2 was made, f2's grid
