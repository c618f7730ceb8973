"""GPU parity of the tracking-side matchers through the C ABI: findFeaturesInArea +
exclusion + getBestMatch (orbx_search_in_area*) and verifyAngle (orbx_verify_angle) against the CPU oracle, which is
itself pinned to the reference's compiled lines (tests/test_oracle_vs_ref.py).

The oracle is run on the keypoints / descriptors the GPU frame produced, so every compared value is integer or an exactly
rounded float division: the bar is bit-exact."""
import numpy as np
import pytest

from orb_slam2_ros2_b200 import api, synth

pytestmark = pytest.mark.gpu


def _camera(c, dist=None):
    return api.Camera(c["fx"], c["fy"], c["cx"], c["cy"], c["bl"], tuple(c["dist"]) if dist is None else dist, c["depth_scale"])


def _same(got, exp, what):
    for k in ("n_cand", "best_idx", "best_dist", "ratio"):
        assert np.array_equal(got[k], exp[k], equal_nan=True), f"{what}: {k} differs at {np.nonzero(got[k] != exp[k])[0][:5]}"


@pytest.mark.parametrize("mode", ["window", "up", "down"])
def test_search_in_area_kitti_frame(oracle, mode):
    c = synth.KITTI
    left, right = synth.synth_stereo_pair(c["height"], c["width"], 3, 17)
    ctx = api.Context(c["width"], c["height"], 2000, 8, 1.2, camera=_camera(c))
    r = ctx.stereo_frame(left, right)
    rows, cols, mnu, mnv, mxu, mxv = ctx.grid_info()
    sf = ctx.scaled_factors()
    for th, seed in [(15.0, 1), (30.0, 2), (3.0, 3)]:
        q, qd, ex, src = synth.synth_area_queries(r.kps_left, r.desc_left, 1800, seed, c["width"], c["height"], 8, th, mode=mode)
        for e in (None, ex):
            got = ctx.search_in_area(q, qd, e)
            exp = oracle.search_in_area(r.kps_left, r.desc_left, (mnu, mnv, mxu, mxv), sf, q, qd, e)
            _same(got, exp, f"{mode} th={th} exclude={e is not None}")
        hit = got["best_idx"] == src
        assert (got["n_cand"] > 0).mean() > 0.6 and hit.mean() > 0.3
    ctx.close()


def test_search_in_area_ties_zero_distance_and_empty(oracle):
    """exact descriptor copies (distance 0, ratio 0/0 = NaN when the runner-up is 0 too), windows that contain nothing, n = 0"""
    c = synth.KITTI
    left, right = synth.synth_stereo_pair(c["height"], c["width"], 4, 17)
    ctx = api.Context(c["width"], c["height"], 2000, 8, 1.2, camera=_camera(c))
    r = ctx.stereo_frame(left, right)
    b = ctx.grid_info()[2:]
    sf = ctx.scaled_factors()
    q, qd, ex, src = synth.synth_area_queries(r.kps_left, r.desc_left, 900, 9, c["width"], c["height"], 8, 40.0, flip_bits=0)
    got, exp = ctx.search_in_area(q, qd), oracle.search_in_area(r.kps_left, r.desc_left, b, sf, q, qd, None)
    _same(got, exp, "zero distance")
    assert (got["best_dist"] == 0).sum() > 500
    # nothing in the octave window / everything excluded
    q2 = q.copy()
    q2["min_level"], q2["max_level"] = 7, 3
    got = ctx.search_in_area(q2, qd)
    assert (got["best_idx"] == -1).all() and (got["n_cand"] == 0).all()
    got = ctx.search_in_area(q, qd, np.ones(2000, np.uint8))
    assert (got["best_idx"] == -1).all()
    assert len(ctx.search_in_area(q[:0], qd[:0])["best_idx"]) == 0
    ctx.close()


def test_search_in_area_tum_rgbd_with_distortion(oracle):
    """640x480: the grid bounds come from the undistorted image corners and floor(mfMaxU / 64) is one past the grid"""
    c = synth.TUM
    dist = (0.231222, -0.28, -0.003257, -0.000105, 0.0)
    for dd in ((0.0,) * 5, dist):
        ctx = api.Context(c["width"], c["height"], 1000, 8, 1.2, camera=_camera(c, dd))
        gray = synth.synth_image(c["height"], c["width"], 12)
        r = ctx.rgbd_frame(gray, synth.synth_depth_u16(c["height"], c["width"], 12, c["depth_scale"]))
        rows, cols, mnu, mnv, mxu, mxv = ctx.grid_info()
        q, qd, ex, src = synth.synth_area_queries(r.kps, r.desc, 1200, 2, c["width"], c["height"], 8, 15.0)
        for e in (None, ex):
            got = ctx.search_in_area(q, qd, e)
            exp = oracle.search_in_area(r.kps, r.desc, (mnu, mnv, mxu, mxv), ctx.scaled_factors(), q, qd, e)
            _same(got, exp, f"tum dist={dd[0] != 0}")
        assert (got["n_cand"] > 0).mean() > 0.5
        ctx.close()


def test_search_in_area_batch_device(oracle):
    import torch

    c = synth.KITTI
    n, nq = 4, 1500
    lefts, rights = synth.synth_stereo_pool(c["height"], c["width"], n, seed0=70)
    ctx = api.Context(c["width"], c["height"], 2000, 8, 1.2, camera=_camera(c), max_batch=n)
    dl, dr = torch.from_numpy(lefts).cuda(), torch.from_numpy(rights).cuda()
    stream = torch.cuda.Stream()
    ctx.set_stream(stream.cuda_stream)
    res = ctx.stereo_batch_device(n, dl.data_ptr(), dr.data_ptr(), c["width"], c["width"] * c["height"])
    kps = ctx.read_device(res.kps_und, (2 * n, 2000), api.KP_DTYPE)[0::2]
    desc = ctx.read_device(res.desc, (2 * n, 2000, 32), np.uint8)[0::2]
    nk = ctx.read_device(res.n_kps, (2 * n,), np.int32)[0::2]
    qs = np.zeros((n, nq), api.AREA_QUERY_DTYPE)
    qds = np.zeros((n, nq, 32), np.uint8)
    exs = np.zeros((n, 2000), np.uint8)
    nqs = np.array([nq, nq - 1, 700, 0], np.int32)
    for f in range(n):
        q, qd, ex, _ = synth.synth_area_queries(kps[f][: nk[f]], desc[f][: nk[f]], nq, 30 + f, c["width"], c["height"], 8, 15.0)
        qs[f], qds[f], exs[f, : len(ex)] = q, qd, ex
    before = ctx.launch_count
    with torch.cuda.stream(stream):
        t_q = torch.from_numpy(qs.view(np.uint8).reshape(n, -1)).cuda()
        t_d = torch.from_numpy(qds).cuda()
        t_x = torch.from_numpy(exs).cuda()
        t_n = torch.from_numpy(nqs).cuda()
        o_idx = torch.full((n, nq), -7, dtype=torch.int32, device="cuda")
        o_dist = torch.zeros((n, nq), dtype=torch.int32, device="cuda")
        o_nc = torch.zeros((n, nq), dtype=torch.int32, device="cuda")
        o_ratio = torch.zeros((n, nq), dtype=torch.float32, device="cuda")
        ctx.search_in_area_batch_device(n, nq, t_q.data_ptr(), t_d.data_ptr(), t_n.data_ptr(), t_x.data_ptr(), o_idx.data_ptr(), o_dist.data_ptr(),
                                        o_ratio.data_ptr(), o_nc.data_ptr())
    assert ctx.launch_count - before == 1
    stream.synchronize()
    b = ctx.grid_info()[2:]
    for f in range(n):
        m = int(nqs[f])
        exp = oracle.search_in_area(kps[f][: nk[f]], desc[f][: nk[f]], b, ctx.scaled_factors(), qs[f][:m], qds[f][:m], exs[f][: nk[f]])
        got = dict(best_idx=o_idx[f, :m].cpu().numpy(), best_dist=o_dist[f, :m].cpu().numpy(), ratio=o_ratio[f, :m].cpu().numpy(),
                   n_cand=o_nc[f, :m].cpu().numpy())
        _same(got, exp, f"frame {f}")
        assert (o_idx[f, m:] == -7).all()  # queries beyond n_queries[f] are not touched
    ctx.set_stream(None)
    ctx.close()


def test_verify_angle(oracle):
    ctx = api.Context(320, 240, 300, 4, 1.2)
    rng = np.random.default_rng(11)
    for n, n1, spread in [(1500, 2000, 25.0), (5000, 900, 90.0), (257, 300, 3.0), (3, 10, 1.0), (1, 1, 0.0), (0, 5, 0.0)]:
        k1, k2 = np.zeros(n1, api.KP_DTYPE), np.zeros(max(n, 1), api.KP_DTYPE)
        k1["angle"] = rng.uniform(-180, 180, n1).astype(np.float32)
        qi = rng.integers(0, n1, n).astype(np.int32)
        ti = rng.permutation(max(n, 1))[:n].astype(np.int32)
        ang = k1["angle"][qi].astype(np.float64) + rng.normal(0, spread, n)
        k2["angle"][ti] = (180.0 - np.mod(180.0 - ang, 360.0)).astype(np.float32)
        k2["angle"][ti[: n // 9]] = k1["angle"][qi[: n // 9]]
        di = rng.integers(0, 50, n).astype(np.float32)
        got, exp = ctx.verify_angle(qi, ti, di, k1, k2), oracle.verify_angle(qi, ti, di, k1, k2)
        assert all(np.array_equal(a, b) for a, b in zip(got, exp)), n
    with pytest.raises(ValueError):
        ctx.verify_angle([5], [0], [1.0], np.zeros(3, api.KP_DTYPE), np.zeros(3, api.KP_DTYPE))
    ctx.close()


def test_state_errors_and_matcher_mirror(oracle, template_path):
    c = synth.KITTI
    ctx = api.Context(c["width"], c["height"], 2000, 8, 1.2, camera=_camera(c))
    q = np.zeros(4, api.AREA_QUERY_DTYPE)
    with pytest.raises(ValueError):
        ctx.search_in_area(q, np.zeros((4, 32), np.uint8))  # no frame yet
    left, right = synth.synth_stereo_pair(c["height"], c["width"], 0, 17)
    ctx.stereo_frame(left, right)
    with pytest.raises(ValueError):
        ctx.search_in_area(q, np.zeros((4, 32), np.uint8), frame=1)
    ctx.close()

    # ORBMatcher::searchByProjection (constant-velocity form, src/ORBMatcher.cc:265-347): "last frame" = the same scene
    # shifted by 4 px; its keypoints with map points are searched in the current frame
    cur = api.Frame.createStereo(left, right, 2000, template_path, 20, 7, None, 8, 1.2, camera=_camera(c))
    l2, r2 = synth.synth_stereo_pair(c["height"], c["width"] + 4, 0, 17)
    last = api.Frame.createStereo(np.ascontiguousarray(l2[:, 4:]), np.ascontiguousarray(r2[:, 4:]), 2000, template_path, 20, 7, None, 8, 1.2,
                                  camera=_camera(c))
    k_last, d_last = last.mvFeatsLeft.copy(), last.mvLeftDescriptor.copy()
    cur = api.Frame.createStereo(left, right, 2000, template_path, 20, 7, None, 8, 1.2, camera=_camera(c))  # the context now holds `cur`
    valid2 = last.mvDepths > 0
    has1 = np.zeros(len(cur.mvFeatsLeft), bool)
    has1[::5] = True
    m = api.ORBMatcher(0.6)
    for z, bl in [(0.0, 0.5), (1.0, 0.5), (-1.0, 0.5)]:
        got = m.searchByProjection(cur, k_last, d_last, valid2, has1, 15, tlc_z=z, baseline=bl)
        # the same decision list composed from the oracle
        idx = np.nonzero(valid2)[0]
        qq = np.zeros(len(idx), oracle.AREA_QUERY_DTYPE)
        qq["x"], qq["y"], qq["octave"], qq["radius"] = k_last["x"][idx], k_last["y"][idx], k_last["octave"][idx], 15
        if z > bl:
            qq["min_level"], qq["max_level"] = qq["octave"], 7
        elif -z > bl:
            qq["min_level"], qq["max_level"] = 0, qq["octave"]
        else:
            qq["min_level"], qq["max_level"] = np.maximum(0, qq["octave"] - 1), np.minimum(qq["octave"] + 1, 7)
        ctx2 = cur.extra["ctx"]
        e = oracle.search_in_area(cur.mvFeatsLeft, cur.mvLeftDescriptor, ctx2.grid_info()[2:], ctx2.scaled_factors(), qq, d_last[idx], has1.astype(np.uint8))
        ok = (e["n_cand"] > 0) & (e["ratio"] < np.float32(0.6)) & (e["best_dist"] < 50)
        exp = np.stack([e["best_idx"][ok], idx[ok], e["best_dist"][ok]], 1).astype(np.int32)
        assert np.array_equal(got, exp) and len(got) > 50
        assert not has1[got[:, 0]].any()
    # local-map form (:561-621): first map point to claim a free keypoint wins
    uv = np.stack([k_last["x"], k_last["y"]], 1)[valid2]
    nm, assigned = m.searchByProjectionMapPoints(cur, uv, k_last["octave"][valid2], np.full(len(uv), 0.999, np.float32), d_last[valid2], 3.0, has1)
    assert nm == int(has1.sum()) + len(assigned) and len(np.unique(assigned[:, 0])) == len(assigned) and not has1[assigned[:, 0]].any()
