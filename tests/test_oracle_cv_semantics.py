"""Pins oracle/cv_semantics.py (numpy restatement) against cv2 4.13.0 itself: the library the
reference delegates its arithmetic to."""
import numpy as np
import pytest

cv2 = pytest.importorskip('cv2')

from oracle import cv_semantics as cs
from synth_euroc import SlidingTextureStream

LK = dict(winSize=(15, 15), maxLevel=3,
          criteria=(cv2.TERM_CRITERIA_EPS | cv2.TERM_CRITERIA_COUNT, 30, 0.01),
          flags=cv2.OPTFLOW_USE_INITIAL_FLOW)


@pytest.fixture(scope='module')
def frames():
    s = SlidingTextureStream(n_frames=3, seed=1, sigma=2.5)
    return s.frame(0), s.frame(1)


def test_pyr_down_bit_exact(frames):
    img = frames[0].cam0_image
    for im in (img, img[:479, :751], img[:61, :95], np.ascontiguousarray(img[:, :188])):
        assert np.array_equal(cs.pyr_down(im), cv2.pyrDown(im))


def test_pyr_down_ramp_known_answer():
    ramp = np.tile(np.arange(0, 64, dtype=np.uint8) * 4, (16, 1))
    out = cs.pyr_down(ramp)
    assert out.shape == (8, 32)
    assert np.array_equal(out, cv2.pyrDown(ramp))
    assert np.array_equal(out[3, 2:30], (np.arange(2, 30) * 8).astype(np.uint8))


def test_scharr_bit_exact(frames):
    img = frames[0].cam0_image
    dx, dy = cs.scharr(img)
    assert np.array_equal(dx, cv2.Scharr(img, cv2.CV_16S, 1, 0))
    assert np.array_equal(dy, cv2.Scharr(img, cv2.CV_16S, 0, 1))


def _cv_fast(img, thr, mask=None):
    det = cv2.FastFeatureDetector_create(thr)
    kps = det.detect(img, mask=mask) if mask is not None else det.detect(img)
    return np.array([(k.pt[0], k.pt[1], k.response) for k in kps]).reshape(-1, 3)


@pytest.mark.parametrize('thr', [15, 25])
def test_fast_exact_incl_order_and_response(frames, thr):
    img = frames[0].cam0_image
    xs, ys, rs = cs.fast_detect(img, thr)
    ref = _cv_fast(img, thr)
    assert len(ref) > 20
    assert np.array_equal(ref, np.stack([xs, ys, rs], 1).astype(float))


def test_fast_mask_is_post_filter(frames):
    img = frames[0].cam0_image
    mask = np.ones_like(img)
    mask[100:200, 100:300] = 0
    mask[::7, ::5] = 0
    xs, ys, rs = cs.fast_detect(img, 15, mask)
    assert np.array_equal(_cv_fast(img, 15, mask), np.stack([xs, ys, rs], 1).astype(float))


def test_fast_handmade_corner_known_answer():
    img = np.full((9, 9), 100, np.uint8)
    img[4, 4] = 200                       # isolated bright pixel: all 16 ring pixels darker by 100
    xs, ys, rs = cs.fast_detect(img, 15)
    assert (list(xs), list(ys), list(rs)) == ([4], [4], [99])
    assert np.array_equal(_cv_fast(img, 15), [[4.0, 4.0, 99.0]])


def test_fast_empty_and_tiny():
    assert len(cs.fast_detect(np.zeros((6, 6), np.uint8), 15)[0]) == 0
    assert len(cs.fast_detect(np.full((32, 32), 7, np.uint8), 15)[0]) == 0


def _points(img, n, seed):
    xs, ys, _ = cs.fast_detect(img, 15)
    idx = np.random.default_rng(seed).choice(len(xs), n, replace=False)
    pts = np.stack([xs[idx], ys[idx]], 1).astype(np.float32)
    pts += np.random.default_rng(seed + 1).uniform(0, 1, pts.shape).astype(np.float32)
    edge = np.array([[2.5, 3.5], [750.2, 478.9], [5, 240], [400, 1.2], [748, 100], [-4, 50],
                     [760, 100], [0, 0], [751, 479], [-9, -9], [770, 500]], np.float32)
    return np.vstack([pts, edge])


def test_lk_temporal_matches_cv2(frames):
    f0, f1 = frames
    pts = _points(f0.cam0_image, 120, 0)
    guess = pts + np.float32([1.0, 0.4])
    q, st, _ = cv2.calcOpticalFlowPyrLK(f0.cam0_image, f1.cam0_image, pts, guess.copy(), **LK)
    q2, st2 = cs.lk_track(cs.build_pyramid(f0.cam0_image, 3), cs.build_pyramid(f1.cam0_image, 3),
                          pts, guess)
    st = st.reshape(-1)
    assert (st == st2).mean() >= 0.995
    both = (st == 1) & (st2 == 1)
    assert both.sum() > 100
    assert np.abs(q - q2)[both].max() < 1e-3          # 0.01 px budget; observed 0
    assert np.abs(q - q2).max() < 1e-3                # stored points agree even when lost


def test_lk_stereo_and_large_guess_error(frames):
    f0, _ = frames
    pts = _points(f0.cam0_image, 80, 5)
    for off in ([-10.0, 0.0], [-30.0, 6.0], [0.0, 0.0]):
        guess = pts + np.float32(off)
        q, st, _ = cv2.calcOpticalFlowPyrLK(f0.cam0_image, f0.cam1_image, pts, guess.copy(), **LK)
        q2, st2 = cs.lk_track(cs.build_pyramid(f0.cam0_image, 3), cs.build_pyramid(f0.cam1_image, 3),
                              pts, guess)
        st = st.reshape(-1)
        assert (st == st2).mean() >= 0.995
        both = (st == 1) & (st2 == 1)
        assert np.abs(q - q2)[both].max() < 1e-3


def test_lk_five_levels_textureless_and_offframe():
    s = SlidingTextureStream(width=640, height=512, n_frames=2, seed=4, sigma=2.0)
    a, b = s.frame(0).cam0_image.copy(), s.frame(1).cam0_image.copy()
    a[200:300, 200:400] = 128             # flat patch -> minEig rejection at level 0
    b[200:300, 200:400] = 128
    lk = dict(LK, maxLevel=4)
    g = np.random.default_rng(0)
    pts = np.vstack([g.uniform([0, 0], [640, 512], (150, 2)), g.uniform([210, 210], [390, 290], (30, 2)),
                     g.uniform([-20, -20], [660, 530], (40, 2))]).astype(np.float32)
    guess = pts + g.normal(0, 2.0, pts.shape).astype(np.float32)
    q, st, _ = cv2.calcOpticalFlowPyrLK(a, b, pts, guess.copy(), **lk)
    q2, st2 = cs.lk_track(cs.build_pyramid(a, 4), cs.build_pyramid(b, 4), pts, guess)
    st = st.reshape(-1)
    assert 0 < st.sum() < len(st)
    assert (st == st2).mean() >= 0.995
    both = (st == 1) & (st2 == 1)
    d = np.abs(q - q2)[both].max(axis=1)
    assert np.mean(d < 0.01) >= 0.995     # random points include ill-conditioned windows


K0 = [458.654, 457.296, 367.215, 248.375]
D0 = np.array([-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05])


@pytest.mark.parametrize('dtype', [np.float32, np.float64])
def test_undistort_distort_match_cv2(dtype):
    g = np.random.default_rng(3)
    pts = g.uniform([0, 0], [752, 480], (500, 2)).astype(dtype)
    Km = np.array([[K0[0], 0, K0[2]], [0, K0[1], K0[3]], [0, 0, 1]])
    R = cv2.Rodrigues(np.array([0.01, -0.02, 0.005]))[0]
    for rot in (None, R):
        ref = cv2.undistortPoints(pts.reshape(-1, 1, 2), Km, D0, None,
                                  np.eye(3) if rot is None else rot, np.eye(3)).reshape(-1, 2)
        got = cs.undistort_radtan(pts, K0, D0, rot)
        assert got.dtype == ref.dtype == dtype
        tol = 0 if dtype == np.float32 else 4e-15
        assert np.abs(ref - got).max() <= tol
        h = cv2.convertPointsToHomogeneous(ref)
        pr, _ = cv2.projectPoints(h, np.zeros(3), np.zeros(3), Km, D0)
        got2 = cs.distort_radtan(ref, K0, D0)
        assert got2.dtype == pr.dtype
        assert np.abs(pr.reshape(-1, 2) - got2).max() <= (0 if dtype == np.float32 else 1e-12)


def test_undistort_principal_point_known_answer():
    out = cs.undistort_radtan(np.array([[K0[2], K0[3]]]), K0, D0)
    assert np.allclose(out, 0.0, atol=1e-7)


def test_rodrigues_matches_cv2():
    from oracle.pipeline_port import rodrigues
    g = np.random.default_rng(0)
    for v in list(g.normal(0, 0.05, (20, 3))) + [np.zeros(3), np.array([1e-20, 0, 0])]:
        assert np.abs(rodrigues(v) - cv2.Rodrigues(v)[0]).max() < 1e-15


@pytest.mark.parametrize('dtype', [np.float32, np.float64])
def test_equidistant_model_matches_cv2_fisheye(dtype):
    """camera_model.py:41-43, 69-70 (cv2.fisheye.undistortPoints / distortPoints).  Groundwork only: the reference's
    own distort_points raises under cv2 4.13 for this model (fisheye.distortPoints wants a 2-channel array and gets
    (N, 2)), so its stereo matcher cannot run on an equidistant camera and libavb does not build the model."""
    Ke = [461.6, 460.3, 362.7, 248.1]
    De = np.array([-0.0091, 0.0666, -0.1028, 0.0612])
    Km = np.array([[Ke[0], 0, Ke[2]], [0, Ke[1], Ke[3]], [0, 0, 1.0]])
    g = np.random.default_rng(5)
    pts = np.stack([g.uniform(-50, 800, 800), g.uniform(-50, 530, 800)], 1).astype(dtype)
    pts[0] = (Ke[2], Ke[3])
    R = cv2.Rodrigues(np.array([0.01, -0.02, 0.03]))[0]
    tol = 0 if dtype == np.float32 else 4e-15
    for rot in (None, R):
        ref = cv2.fisheye.undistortPoints(pts.reshape(-1, 1, 2), Km, De, R=np.eye(3) if rot is None else rot,
                                          P=np.eye(3)).reshape(-1, 2)
        got = cs.undistort_equidistant(pts, Ke, De, rot)
        assert got.dtype == ref.dtype == dtype and np.abs(ref - got).max() <= tol
    und = cs.undistort_equidistant(pts, Ke, De)
    ref = cv2.fisheye.distortPoints(und.reshape(-1, 1, 2), Km, De).reshape(-1, 2)
    got = cs.distort_equidistant(und, Ke, De)
    assert np.abs(ref - got).max() <= (0 if dtype == np.float32 else 1e-12)
    assert np.abs(got.astype(np.float64) - pts).max() < (1e-3 if dtype == np.float32 else 1e-9)      # round trip
    # points the Newton iteration cannot invert are flagged (-1e6, -1e6), exactly where cv2 flags them
    hard = np.array([0.5, -2.0, 3.0, -4.0])
    far = np.stack([g.uniform(-500, 1500, 300), g.uniform(-500, 1200, 300)], 1).astype(dtype)
    ref = cv2.fisheye.undistortPoints(far.reshape(-1, 1, 2), Km, hard, R=np.eye(3), P=np.eye(3)).reshape(-1, 2)
    got = cs.undistort_equidistant(far, Ke, hard)
    assert np.array_equal(ref[:, 0] == -1e6, got[:, 0] == -1e6) and (got[:, 0] == -1e6).sum() > 100
    assert np.abs(ref - got).max() <= tol
    with pytest.raises(cv2.error):                            # the reference's call shape (camera_model.py:70)
        cv2.fisheye.distortPoints(und, Km, De)


def test_integer_stages_against_cv2_on_random_shapes():
    """Property check over random sizes and contents (incl. odd sizes, constant and saturated images): pyrDown, Scharr
    and FAST (keypoints in scan order with responses, with and without a mask) equal cv2's, bit for bit."""
    from hypothesis import given, settings, strategies as st, HealthCheck

    det = cv2.FastFeatureDetector_create(15)

    @settings(max_examples=30, deadline=None, suppress_health_check=list(HealthCheck))
    @given(st.integers(8, 70), st.integers(8, 90), st.integers(0, 2 ** 31 - 1), st.sampled_from(['noise', 'smooth', 'flat', 'sat']))
    def check(h, w, seed, kind):
        g = np.random.default_rng(seed)
        if kind == 'noise':
            img = g.integers(0, 256, (h, w)).astype(np.uint8)
        elif kind == 'smooth':
            img = cv2.GaussianBlur(g.integers(0, 256, (h, w)).astype(np.uint8), (0, 0), 1.5)
        elif kind == 'flat':
            img = np.full((h, w), int(g.integers(0, 256)), np.uint8)
        else:
            img = (g.integers(0, 2, (h, w)) * 255).astype(np.uint8)
        assert np.array_equal(cs.pyr_down(img), cv2.pyrDown(img))
        dx, dy = cs.scharr(img)
        assert np.array_equal(dx, cv2.Scharr(img, cv2.CV_16S, 1, 0)) and np.array_equal(dy, cv2.Scharr(img, cv2.CV_16S, 0, 1))
        mask = (g.integers(0, 4, (h, w)) > 0).astype(np.uint8)
        for m in (None, mask):
            kps = det.detect(img, mask=m) if m is not None else det.detect(img)
            xs, ys, rs = cs.fast_detect(img, 15, m)
            assert [(int(k.pt[0]), int(k.pt[1]), int(k.response)) for k in kps] == list(zip(xs.tolist(), ys.tolist(), rs.tolist()))

    check()
