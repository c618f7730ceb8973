"""Pin the oracle's OpenCV primitives against the real OpenCV (cv2 4.13.0): resize, GaussianBlur, FAST, undistort.

The reference calls these at src/ORBExtractor.cc:316,319,365,367 and src/Camera.cc:36; OpenCV is not vendored in the
reference, so cv2 is the only executable ground truth for them.
"""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")

from orb_slam2_ros2_b200 import synth

SHAPES = [(376, 1241), (480, 640), (1080, 1920), (97, 131)]


def _images(h, w):
    rng = np.random.default_rng(h * 7 + w)
    yield synth.synth_image(h, w, 1)
    yield rng.integers(0, 256, (h, w), dtype=np.uint8)
    yield (rng.random((h, w)) > 0.5).astype(np.uint8) * 255


@pytest.mark.parametrize("shape", SHAPES)
def test_resize_matches_cv2(oracle, shape):
    h, w = shape
    n_levels = 12 if h >= 1080 else 8
    rc, lw, lh = oracle.level_sizes(w, h, 1.2, n_levels)
    img = next(_images(h, w))
    for l in range(1, n_levels):
        if lw[l] < 8 or lh[l] < 8:
            continue
        a = oracle.resize_linear(img, int(lw[l]), int(lh[l]))
        b = cv2.resize(img, (int(lw[l]), int(lh[l])), interpolation=cv2.INTER_LINEAR)
        assert np.array_equal(a, b), f"level {l}"


@pytest.mark.parametrize("dsize", [(620, 188), (2480, 752), (1300, 400), (100, 50), (1240, 188), (1239, 375), (413, 125)])
def test_resize_odd_ratios_matches_cv2(oracle, dsize):
    img = synth.synth_image(376, 1240, 3)
    a = oracle.resize_linear(img, *dsize)
    b = cv2.resize(img, dsize, interpolation=cv2.INTER_LINEAR)
    assert np.array_equal(a, b)


@pytest.mark.parametrize("shape", SHAPES + [(38, 38), (62, 75)])
def test_blur_matches_cv2(oracle, shape):
    for img in _images(*shape):
        a = oracle.gaussian_blur7(img)
        b = cv2.GaussianBlur(img, (7, 7), 2, sigmaY=2, borderType=cv2.BORDER_REFLECT_101)
        assert np.array_equal(a, b)


def _cv_fast(img, t):
    det = cv2.FastFeatureDetector_create(threshold=t, nonmaxSuppression=True, type=cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    k = det.detect(np.ascontiguousarray(img))
    return np.array([[p.pt[0], p.pt[1], p.response] for p in k], dtype=np.float64).astype(np.int32).reshape(-1, 3)


@pytest.mark.parametrize("t", [20, 7, 1, 60])
def test_fast_matches_cv2(oracle, t):
    for h, w in [(376, 1241), (45, 45), (36, 37), (7, 7), (8, 30), (6, 40)]:
        for img in _images(h, w):
            a = oracle.fast9_nms(img, t)
            b = _cv_fast(img, t)
            assert np.array_equal(a, b), (h, w, t, len(a), len(b))


def test_fast_patch_views_match_cv2(oracle):
    """cell patches are strided views of the level image, exactly like image.rowRange().colRange() (:363)"""
    img = synth.synth_image(376, 1241, 5)
    rng = np.random.default_rng(0)
    for _ in range(40):
        y0, x0 = int(rng.integers(0, 330)), int(rng.integers(0, 1190))
        ph, pw = int(rng.integers(7, 46)), int(rng.integers(7, 46))
        patch = img[y0 : y0 + ph, x0 : x0 + pw]
        for t in (20, 7):
            assert np.array_equal(oracle.fast9_nms(patch, t), _cv_fast(patch, t))


def test_fast_score_is_arc_value_minus_one(oracle):
    img = synth.synth_image(120, 160, 9)
    res = oracle.fast9_nms(img, 7)
    import ctypes as C

    for x, y, s in res[:200]:
        sub = np.ascontiguousarray(img[y - 3 : y + 4, x - 3 : x + 4])
        m = oracle.lib().oracle_fast9_arc_value(sub[3:, 3:].ctypes.data_as(C.POINTER(C.c_uint8)), sub.strides[0])
        assert m - 1 == s and m > 7


def test_undistort_matches_cv2(oracle):
    c = synth.TUM
    rng = np.random.default_rng(4)
    pts = np.stack([rng.uniform(0, 640, 2000), rng.uniform(0, 480, 2000)], 1).astype(np.float32)
    K = np.array([[c["fx"], 0, c["cx"]], [0, c["fy"], c["cy"]], [0, 0, 1]], np.float32)
    # 5-coefficient TUM model, and a 4-coefficient model (System.cc:63-73 drops k3 when it is 0); the 4-coefficient
    # case uses a milder k2 because the TUM k1/k2 without k3 make the fixed-point iteration diverge at the corners
    for dist in (np.array(c["dist"], np.float32), np.array([0.231222, -0.12, -0.003257, -0.000105], np.float32)):
        ref = cv2.undistortPoints(pts.reshape(-1, 1, 2), K, dist, None, K).reshape(-1, 2)
        got = oracle.undistort_points(pts, c["fx"], c["fy"], c["cx"], c["cy"], dist)
        assert np.abs(got - ref).max() < 1e-3  # tolerance of the uRight/depth requirement
        assert np.abs(got - pts).max() > 1.0  # the distortion really moves points


def test_umax_table(oracle):
    assert list(oracle.umax()) == [15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3]


def test_level_tables_kitti(oracle):
    # SURVEY.md Appendix B (float32-exact values derived from src/ORBExtractor.cc:283-317)
    assert list(oracle.level_quotas(2000, 1.2, 8)) == [434, 362, 302, 252, 210, 175, 146, 119]
    assert list(oracle.level_quotas(500, 1.2, 8)) == [109, 91, 76, 63, 52, 43, 36, 30]
    assert list(oracle.level_quotas(1000, 1.2, 8)) == [217, 181, 151, 126, 105, 88, 73, 59]
    assert list(oracle.level_quotas(4000, 1.2, 8)) == [869, 724, 603, 502, 418, 348, 290, 246]
    assert list(oracle.level_quotas(5000, 1.2, 12)) == [939, 782, 652, 543, 452, 377, 314, 262, 218, 182, 152, 127]
    rc, lw, lh = oracle.level_sizes(1241, 376, 1.2, 8)
    assert rc == 0 and list(lw) == [1241, 1034, 862, 718, 598, 499, 416, 346] and list(lh) == [376, 313, 261, 218, 181, 151, 126, 105]
    rc, lw, lh = oracle.level_sizes(640, 480, 1.2, 8)
    assert list(lw) == [640, 533, 444, 370, 309, 257, 214, 179] and list(lh) == [480, 400, 333, 278, 231, 193, 161, 134]
    rc, lw, lh = oracle.level_sizes(1920, 1080, 1.2, 12)
    assert list(lw) == [1920, 1600, 1333, 1111, 926, 772, 643, 536, 447, 372, 310, 258]
    rc, _, _ = oracle.level_sizes(100, 60, 1.2, 8)
    assert rc == -1  # ImageSizeError (:310-314)
