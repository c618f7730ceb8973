"""The quadtree kernel alone, on arbitrary corner lists, against the reference's Quadtree (compiled reference when
available, else the C restatement): corners on split lines, starved levels, need in {0,1,2}, deep single-pixel clusters
(beyond the 9 key levels), wide ROIs with more than 31 root strips (generic root path), levels denser than the
shared-memory list (global-scratch path)."""
import numpy as np
import pytest

from orb_slam2_ros2_b200 import api

pytestmark = pytest.mark.gpu


def _expected(oracle, w, h, xs, ys, resp, need):
    if oracle.have_ref() and not (len(xs) == 0 and need == 1):
        idx = oracle.ref_quadtree(w, h, xs, ys, resp, need)
    else:
        idx, _ = oracle.quadtree_select(w, h, xs, ys, resp, need)
    return np.stack([xs[idx], ys[idx], resp[idx]], 1).astype(np.int32).reshape(-1, 3)


def _ctx(w, h, need):
    # one level whose ROI is w x h and whose quota is `need` (a single level gets the whole budget, ORBExtractor.cc:301)
    return api.Context(w + 32, h + 32, max(need, 1), 1, 1.2)


def _check(oracle, ctx, w, h, xs, ys, resp, need, level=0):
    got = ctx.run_quadtree(level, xs, ys, resp)
    exp = _expected(oracle, w, h, xs.astype(np.float32), ys.astype(np.float32), resp.astype(np.float32), need)
    assert np.array_equal(got, exp), (w, h, len(xs), need, len(got), len(exp))


def _unique_points(rng, w, h, n):
    n = min(n, (w - 6) * (h - 6))
    flat = rng.choice((w - 6) * (h - 6), n, replace=False)
    flat.sort()  # detection order is row-major inside a cell; any fixed order works for the quadtree itself
    return (flat % (w - 6) + 3).astype(np.int32), (flat // (w - 6) + 3).astype(np.int32)


@pytest.mark.parametrize("fast", ["1", "0"], ids=["loopfree", "sequential"])
@pytest.mark.parametrize("shape", [(1209, 344), (608, 448), (314, 73), (100, 400), (64, 64), (1888, 1048)])
def test_random_lists(oracle, shape, fast, monkeypatch):
    """both formulations of Quadtree::split(): the loop-free one (default) and the sequential loop (ORBX_QT_FAST=0)"""
    monkeypatch.setenv("ORBX_QT_FAST", fast)
    w, h = shape
    rng = np.random.default_rng(w * 7 + h)
    for need in (2, 7, 60, 434):
        ctx = _ctx(w, h, need)
        for n in (0, 1, 5, need - 1, need, need + 3, 3 * need, 2500):
            if n < 0:
                continue
            xs, ys = _unique_points(rng, w, h, min(n, w * h // 5))  # a level's cell slots hold about w*h/4 corners
            resp = rng.integers(6, 120, len(xs)).astype(np.int32)
            _check(oracle, ctx, w, h, xs, ys, resp, need)
        n_fast, n_seq = ctx.quadtree_stats()
        if fast == "0":
            assert n_fast == 0
        elif shape in ((1209, 344), (608, 448), (1888, 1048)):
            assert n_fast >= 6, (n_fast, n_seq)  # non-dyadic strips: random corners never force the fallback
        ctx.close()


def test_loop_free_path_is_the_one_that_runs_on_images(oracle):
    """every level of ordinary images is solved without the sequential loop -- and still equals the oracle"""
    from orb_slam2_ros2_b200 import synth

    for (h, w, nf, nl, seed) in ((376, 1241, 2000, 8, 0), (480, 640, 1000, 8, 3), (1080, 1920, 5000, 12, 4), (376, 1241, 4000, 8, 2)):
        img = synth.synth_image(h, w, seed)
        ctx = api.Context(w, h, nf, nl, 1.2)
        kps, desc = ctx.extract(img)
        e = oracle.extract(img, nf, nl, 1.2)
        assert len(kps) == len(e.kps) and np.array_equal(kps["x"], e.kps["x"]) and np.array_equal(kps["y"], e.kps["y"])
        assert np.array_equal(kps["octave"], e.kps["octave"]) and np.array_equal(kps["response"], e.kps["response"])
        assert ctx.quadtree_stats() == (nl, 0)
        ctx.close()


def test_corners_on_split_lines(oracle):
    w, h = 1208, 344  # strips of exactly 302 px, midlines on integer coordinates at several depths
    rng = np.random.default_rng(3)
    ctx = _ctx(w, h, 300)
    for trial in range(6):
        xs, ys = _unique_points(rng, w, h, 1500)
        k = len(xs) // 3
        xs[:k] = rng.choice([302, 604, 906, 151, 453, 755, 1057, 75, 226], k)  # strip bounds and depth-1/2 midlines (x)
        ys[k : 2 * k] = rng.choice([172, 86, 258, 43, 129], k)                 # depth-1/2/3 midlines (y)
        key = xs.astype(np.int64) * 4096 + ys
        _, first = np.unique(key, return_index=True)
        first.sort()
        xs, ys = xs[first], ys[first]
        resp = rng.integers(6, 120, len(xs)).astype(np.int32)
        _check(oracle, ctx, w, h, xs, ys, resp, 300)
    ctx.close()


def test_starved_and_tiny_quotas(oracle):
    w, h = 600, 200
    rng = np.random.default_rng(5)
    xs, ys = _unique_points(rng, w, h, 150)
    resp = rng.integers(6, 120, len(xs)).astype(np.int32)
    for need in (1, 2, 3, 149, 150, 151, 400):
        ctx = _ctx(w, h, need)
        _check(oracle, ctx, w, h, xs, ys, resp, need)
        ctx.close()
    # need == 0: the second level of a 2-level context with nFeatures == 1 has quota 0
    ctx = api.Context(300, 200, 1, 2, 1.2)
    lw, lh, _, q = ctx.level_info(1)
    assert q == 0
    got = ctx.run_quadtree(1, xs[:50] % (lw - 38) + 3, ys[:50] % (lh - 38) + 3, resp[:50])
    assert len(got) == 0
    ctx.close()


def test_deep_clusters_beyond_key_levels(oracle):
    """tight clusters of adjacent pixels force splits well below depth 9 (the geometry fallback in fixed point)"""
    w, h = 1209, 344
    rng = np.random.default_rng(9)
    ctx = _ctx(w, h, 400)
    for trial in range(5):
        cx, cy = rng.integers(20, w - 20, 60), rng.integers(20, h - 20, 60)
        pts = set()
        for a, b in zip(cx, cy):
            for dx in range(3):
                for dy in range(3):
                    pts.add((int(a + dx), int(b + dy)))
        pts = sorted(pts, key=lambda p: (p[1], p[0]))
        xs = np.array([p[0] for p in pts], np.int32)
        ys = np.array([p[1] for p in pts], np.int32)
        resp = rng.integers(6, 120, len(xs)).astype(np.int32)
        _check(oracle, ctx, w, h, xs, ys, resp, 400)
    ctx.close()


def test_many_root_strips_generic_path(oracle):
    w, h = 3400, 90  # round(w / h) = 38 strips > 31: keys are not used
    rng = np.random.default_rng(11)
    ctx = _ctx(w, h, 200)
    for n in (50, 600, 3000):
        xs, ys = _unique_points(rng, w, h, n)
        resp = rng.integers(6, 120, len(xs)).astype(np.int32)
        _check(oracle, ctx, w, h, xs, ys, resp, 200)
    ctx.close()


def test_dense_level_global_scratch_path(oracle):
    w, h = 1209, 344
    rng = np.random.default_rng(13)
    ctx = _ctx(w, h, 434)
    for n in (3584, 3585, 9000, 40000):
        xs, ys = _unique_points(rng, w, h, n)
        resp = rng.integers(6, 120, len(xs)).astype(np.int32)
        _check(oracle, ctx, w, h, xs, ys, resp, 434)
    ctx.close()


def test_equal_responses_pick_lowest_index(oracle):
    w, h = 400, 300
    rng = np.random.default_rng(17)
    xs, ys = _unique_points(rng, w, h, 900)
    resp = np.full(len(xs), 30, np.int32)
    ctx = _ctx(w, h, 120)
    _check(oracle, ctx, w, h, xs, ys, resp, 120)
    ctx.close()
