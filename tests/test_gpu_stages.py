"""GPU parity, stage by stage, through the C-ABI (ctypes): CUDA kernels vs the oracle on identical inputs.
Bit-exact for the integer stages (pyramid, FAST); <= 0.01 px and >= 99.5 % status agreement for LK."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import cv_semantics as cs
from frontend_config import FrontEndConfig, config_c2, config_c3, config_default
from oracle.pipeline_port import FrontEndPort
from synth_euroc import SlidingTextureStream

try:
    import cv2
except Exception:                                   # pragma: no cover
    cv2 = None


@pytest.fixture(scope='module')
def ctx752():
    from image_processing import _native
    c = _native.Context(config_c2(), 752, 480, use_graph=False)
    yield c
    c.close()


@pytest.fixture(scope='module')
def frames752():
    s = SlidingTextureStream(n_frames=3, seed=1, sigma=2.5)
    return s.frame(0), s.frame(1)


def test_pyramid_bit_exact(ctx752, frames752):
    f0, _ = frames752
    ctx752.upload(f0.cam0_image, f0.cam1_image)
    ctx752.build_pyramids()
    for slot, img in ((0, f0.cam0_image), (1, f0.cam1_image)):
        ref = cs.build_pyramid(img, 3)
        for lvl in range(4):
            got = ctx752.download_level(slot, lvl)
            assert got.shape == ref[lvl].shape
            assert np.array_equal(got, ref[lvl]), f'slot {slot} level {lvl}'


@pytest.mark.parametrize('w,h,levels', [(1280, 1024, 4), (640, 512, 4), (80 * 16, 400, 3), (96, 96, 1), (752, 481, 3),
                                        (1040, 1030, 5)])      # last: level 5 (33 px wide) is built from an odd-width level (65 px)
def test_pyramid_other_sizes(w, h, levels):
    from image_processing import _native
    cfg = FrontEndConfig(pyramid_levels=levels, width=w, height=h)
    g = np.random.default_rng(w + h)
    img0 = g.integers(0, 256, (h, w)).astype(np.uint8)
    img1 = g.integers(0, 256, (h, w)).astype(np.uint8)
    c = _native.Context(cfg, w, h, use_graph=False)
    try:
        c.upload(img0, img1)
        c.build_pyramids()
        for slot, img in ((0, img0), (1, img1)):
            ref = cs.build_pyramid(img, levels)
            for lvl in range(levels + 1):
                assert np.array_equal(c.download_level(slot, lvl), ref[lvl]), (slot, lvl)
    finally:
        c.close()


def test_fast_exact_incl_order_response_and_mask(ctx752, frames752):
    f0, _ = frames752
    ctx752.upload(f0.cam0_image, f0.cam1_image)
    xs, ys, rs = ctx752.fast_detect()
    exs, eys, ers = cs.fast_detect(f0.cam0_image, 15)
    assert len(exs) > 1000
    assert np.array_equal(xs, exs) and np.array_equal(ys, eys) and np.array_equal(rs, ers)
    mask = np.ones_like(f0.cam0_image)
    mask[100:200, 100:300] = 0
    mask[::7, ::5] = 0
    xs, ys, rs = ctx752.fast_detect(mask)
    exs, eys, ers = cs.fast_detect(f0.cam0_image, 15, mask)
    assert np.array_equal(xs, exs) and np.array_equal(ys, eys) and np.array_equal(rs, ers)
    if cv2 is not None:
        kps = cv2.FastFeatureDetector_create(15).detect(f0.cam0_image, mask=mask)
        assert np.array_equal(np.array([(k.pt[0], k.pt[1], k.response) for k in kps]),
                              np.stack([xs, ys, rs], 1).astype(float))


def test_fast_dense_noise_and_flat_images():
    from image_processing import _native
    c = _native.Context(config_default(), 752, 480, use_graph=False)
    try:
        g = np.random.default_rng(5)
        noise = g.integers(0, 256, (480, 752)).astype(np.uint8)       # worst case corner density
        c.upload(noise, noise)
        xs, ys, rs = c.fast_detect()
        exs, eys, ers = cs.fast_detect(noise, 15)
        assert len(exs) > 10000
        assert np.array_equal(xs, exs) and np.array_equal(ys, eys) and np.array_equal(rs, ers)
        flat = np.full((480, 752), 77, np.uint8)
        c.upload(flat, flat)
        assert len(c.fast_detect()[0]) == 0
    finally:
        c.close()


def test_maximum_image_size():
    """The largest image the library accepts (4096 wide, W * H < 2^24: the scan index of a FAST key has 24 bits): pyramid
    bit-exact on all 5 levels, FAST exact incl. order at the far corner of the index range, LK against the numpy oracle."""
    from image_processing import _native
    from synth_euroc import make_texture
    w, h, levels = 4096, 4095, 5
    # 16 x 16 cells of 256 x 255 px: a cell's FAST bucket must fit the selection kernel's shared memory (cell area <= ~200 k px)
    cfg = FrontEndConfig(grid_row=16, grid_col=16, pyramid_levels=levels, width=w, height=h)
    a = np.clip(np.rint(make_texture(h + 8, w + 8, 21, 2.0)), 0, 255).astype(np.uint8)
    img0, img1 = np.ascontiguousarray(a[2:h + 2, 3:w + 3]), np.ascontiguousarray(a[:h, :w])
    c = _native.Context(cfg, w, h, use_graph=False)
    try:
        c.upload(img0, img0)
        c.build_pyramids()
        ref0 = cs.build_pyramid(img0, levels)
        for lvl in range(levels + 1):
            assert np.array_equal(c.download_level(0, lvl), ref0[lvl]), lvl
        xs, ys, rs = c.fast_detect()
        if cv2 is not None:
            kps = cv2.FastFeatureDetector_create(cfg.fast_threshold).detect(img0)
            rx = np.array([int(k.pt[0]) for k in kps]); ry = np.array([int(k.pt[1]) for k in kps])
            rr = np.array([int(k.response) for k in kps])
        else:
            rx, ry, rr = cs.fast_detect(img0, cfg.fast_threshold)
        assert len(xs) == len(rx) > 100000
        assert np.array_equal(xs, rx) and np.array_equal(ys, ry) and np.array_equal(rs, rr)
        assert ys.max() >= h - 8 and xs.max() >= w - 8            # keypoints next to the last rows / columns
        c.advance()
        c.upload(img1, img1)
        c.build_pyramids()
        g = np.random.default_rng(2)
        far = np.stack([xs[-60:], ys[-60:]], 1).astype(np.float32)          # the bottom rows of the image
        pts = np.vstack([far, g.uniform([10, 10], [w - 10, h - 10], (60, 2)).astype(np.float32)])
        guess = pts + np.float32([2.0, 1.5])
        q, st = c.klt_track(2, 0, pts, guess)
    finally:
        c.close()
    q_ref, st_ref = cs.lk_track(ref0, cs.build_pyramid(img1, levels), pts, guess)
    assert np.array_equal(st, st_ref) and np.array_equal(q[st_ref == 1], q_ref[st_ref == 1])
    assert st.sum() > 60
    print(f'4096x4095: {len(xs)} keypoints exact, 6 pyramid levels exact, {int(st.sum())} of {len(st)} tracks identical to the oracle')


def _points(img, n, seed):
    xs, ys, _ = cs.fast_detect(img, 15)
    idx = np.random.default_rng(seed).choice(len(xs), n, replace=False)
    pts = np.stack([xs[idx], ys[idx]], 1).astype(np.float32)
    pts += np.random.default_rng(seed + 1).uniform(0, 1, pts.shape).astype(np.float32)
    h, w = img.shape
    edge = np.array([[2.5, 3.5], [w - 1.8, h - 1.1], [5, h / 2], [w / 2, 1.2], [w - 4, 100], [-4, 50],
                     [w + 8, 100], [0, 0], [w - 1, h - 1], [-9, -9], [w + 18, h + 20]], np.float32)
    return np.vstack([pts, edge])


def _assert_lk(q, st, q_ref, st_ref, what):
    st_ref = st_ref.reshape(-1)
    agree = (st == st_ref).mean()
    both = (st == 1) & (st_ref == 1)
    d = np.abs(q - q_ref)[both].max(axis=1) if both.any() else np.zeros(1)
    print(f'{what}: n={len(st)} status agreement {agree:.4f}, max |d| {d.max():.3g} px, '
          f'bit-exact {np.mean(d == 0):.4f}, tracked {both.sum()}')
    assert agree >= 0.995, what
    assert d.max() <= 0.01, what                       # tolerance of BASELINE.json: 0.01 px


def test_klt_temporal_vs_numpy_oracle(ctx752, frames752):
    f0, f1 = frames752
    ctx752.upload(f0.cam0_image, f0.cam1_image)
    ctx752.build_pyramids()
    ctx752.advance()
    ctx752.upload(f1.cam0_image, f1.cam1_image)
    ctx752.build_pyramids()
    pts = _points(f0.cam0_image, 150, 0)
    guess = pts + np.float32([1.0, 0.4])
    q, st = ctx752.klt_track(2, 0, pts, guess)
    q_ref, st_ref = cs.lk_track(cs.build_pyramid(f0.cam0_image, 3), cs.build_pyramid(f1.cam0_image, 3), pts, guess)
    _assert_lk(q, st, q_ref, st_ref, 'temporal vs numpy oracle')
    assert np.array_equal(q, q_ref) and np.array_equal(st, st_ref)     # same exact-sum arithmetic -> identical


@pytest.mark.skipif(cv2 is None, reason='cv2 not importable')
def test_klt_vs_cv2_many_points(ctx752, frames752):
    f0, f1 = frames752
    lk = dict(winSize=(15, 15), maxLevel=3, criteria=(cv2.TERM_CRITERIA_EPS | cv2.TERM_CRITERIA_COUNT, 30, 0.01),
              flags=cv2.OPTFLOW_USE_INITIAL_FLOW)
    ctx752.upload(f0.cam0_image, f0.cam1_image)
    ctx752.build_pyramids()
    ctx752.advance()
    ctx752.upload(f1.cam0_image, f1.cam1_image)
    ctx752.build_pyramids()
    pts = _points(f0.cam0_image, 3000, 3)
    g = np.random.default_rng(9)
    guess = pts + g.normal(0, 1.5, pts.shape).astype(np.float32)
    q, st = ctx752.klt_track(2, 0, pts, guess)
    q_ref, st_ref, _ = cv2.calcOpticalFlowPyrLK(f0.cam0_image, f1.cam0_image, pts, guess.copy(), **lk)
    _assert_lk(q, st, q_ref, st_ref, 'temporal vs cv2')
    # stereo direction, current cam0 -> current cam1, poor initial guess
    guess = pts - np.float32([8.0, 0.0])
    q, st = ctx752.klt_track(0, 1, pts, guess)
    q_ref, st_ref, _ = cv2.calcOpticalFlowPyrLK(f1.cam0_image, f1.cam1_image, pts, guess.copy(), **lk)
    _assert_lk(q, st, q_ref, st_ref, 'stereo vs cv2')


@pytest.mark.skipif(cv2 is None, reason='cv2 not importable')
def test_klt_c3_five_levels_textureless_offframe():
    from image_processing import _native
    cfg = config_c3()
    s = SlidingTextureStream(width=1280, height=1024, n_frames=2, seed=4, sigma=2.0)
    a, b = s.frame(0).cam0_image.copy(), s.frame(1).cam0_image.copy()
    a[200:300, 200:400] = 128
    b[200:300, 200:400] = 128
    lk = dict(winSize=(15, 15), maxLevel=4, criteria=(cv2.TERM_CRITERIA_EPS | cv2.TERM_CRITERIA_COUNT, 30, 0.01),
              flags=cv2.OPTFLOW_USE_INITIAL_FLOW)
    g = np.random.default_rng(0)
    pts = np.vstack([g.uniform([0, 0], [1280, 1024], (1500, 2)), g.uniform([210, 210], [390, 290], (100, 2)),
                     g.uniform([-20, -20], [1300, 1044], (400, 2))]).astype(np.float32)
    guess = pts + g.normal(0, 2.0, pts.shape).astype(np.float32)
    c = _native.Context(cfg, 1280, 1024, use_graph=False)
    try:
        c.upload(a, a)
        c.build_pyramids()
        c.advance()
        c.upload(b, b)
        c.build_pyramids()
        q, st = c.klt_track(2, 0, pts, guess)
    finally:
        c.close()
    q_ref, st_ref, _ = cv2.calcOpticalFlowPyrLK(a, b, pts, guess.copy(), **lk)
    st_ref = st_ref.reshape(-1)
    assert 0 < st_ref.sum() < len(st_ref)
    assert (st == st_ref).mean() >= 0.995
    both = (st == 1) & (st_ref == 1)
    d = np.abs(q - q_ref)[both].max(axis=1)
    print(f'C3 random points: status agreement {(st == st_ref).mean():.4f}, within 0.01 px {np.mean(d <= 0.01):.4f}, '
          f'bit-exact {np.mean(d == 0):.4f}')
    assert np.mean(d <= 0.01) >= 0.995               # random points include ill-conditioned windows


def _box_sum(a, r=7):
    """Sum over the (2r+1)^2 window around every pixel (zero outside)."""
    c = np.pad(a, ((r + 1, r), (r + 1, r))).cumsum(0).cumsum(1)
    n = 2 * r + 1
    return c[n:, n:] - c[:-n, n:] - c[n:, :-n] + c[:-n, :-n]


def _contrast_frames(kind):
    """Two 752x480 frames outside the range of the bench texture.  'binary': black/white blobs with sharp edges (structure
    tensor sums above LK_NARROW_MAX = 3e8 -> the wide, two-half exact reduction of avb_lk.cuh); 'faint': mid-grey with
    one or two grey levels of texture, in bands of rising amplitude (minimum eigenvalue around cv2's 1e-4 threshold ->
    the exact sqrt/division branch of lk_tensor_f instead of its band test)."""
    st = SlidingTextureStream(n_frames=2, seed=11, sigma=1.5, drift=(2.0, 1.0))
    if kind == 'binary':
        st._tex = np.where(st._tex > 127.5, 255.0, 0.0)
    else:
        amp = np.linspace(0.2, 12.0, st._tex.shape[1])[None, :]
        st._tex = 128.0 + (st._tex - 127.5) / 127.5 * amp
    return st.frame(0), st.frame(1)


@pytest.mark.parametrize('kind', ['binary', 'faint'])
def test_klt_outside_the_bench_texture_range(ctx752, kind):
    """The LK shortcuts (narrow sums, eigenvalue band test) have an exact path behind them for the inputs they do not
    cover; these images take it.  Against the numpy oracle: identical; against cv2: the BASELINE tolerance."""
    f0, f1 = _contrast_frames(kind)
    ix, iy = cs.scharr(f0.cam0_image)
    q11 = _box_sum(ix.astype(np.int64) ** 2)
    if kind == 'binary':
        assert q11.max() > 3.0e8, q11.max()                # integer template sums are on the same scale as Scharr^2
    else:
        lam = q11[20:-20, 20:-20] / 2.0 ** 20 / 225.0 / 2
        assert lam.min() < 1e-4 < lam.max()
    ctx752.upload(f0.cam0_image, f0.cam1_image)
    ctx752.build_pyramids()
    ctx752.advance()
    ctx752.upload(f1.cam0_image, f1.cam1_image)
    ctx752.build_pyramids()
    g = np.random.default_rng(3)
    if kind == 'binary':                                    # windows on the strongest structure first
        ys, xs = np.unravel_index(np.argsort(q11[10:-10, 10:-10], axis=None)[::-1][:4000:40], q11[10:-10, 10:-10].shape)
        strong = np.stack([xs + 10, ys + 10], 1).astype(np.float32)
    else:
        strong = np.zeros((0, 2), np.float32)
    pts = np.vstack([strong, g.uniform([8, 8], [744, 472], (2000, 2)).astype(np.float32)])
    pts += g.uniform(0, 1, pts.shape).astype(np.float32)
    guess = pts + np.float32([-1.5, -0.7]) + g.normal(0, 0.5, pts.shape).astype(np.float32)
    q, st = ctx752.klt_track(2, 0, pts, guess)
    n = 160
    q_ref, st_ref = cs.lk_track(cs.build_pyramid(f0.cam0_image, 3), cs.build_pyramid(f1.cam0_image, 3), pts[:n], guess[:n])
    assert np.array_equal(st[:n], st_ref) and np.array_equal(q[:n][st_ref == 1], q_ref[st_ref == 1]), kind
    assert 0 < st.sum()
    if kind == 'faint':
        assert st.sum() < len(st)                           # the faint end of the image is below the threshold
    if cv2 is not None:
        lk = dict(winSize=(15, 15), maxLevel=3, criteria=(cv2.TERM_CRITERIA_EPS | cv2.TERM_CRITERIA_COUNT, 30, 0.01),
                  flags=cv2.OPTFLOW_USE_INITIAL_FLOW)
        q_cv, st_cv, _ = cv2.calcOpticalFlowPyrLK(f0.cam0_image, f1.cam0_image, pts, guess.copy(), **lk)
        st_cv = st_cv.reshape(-1)
        both = (st == 1) & (st_cv == 1)
        d = np.abs(q - q_cv)[both].max(axis=1)
        print(f'{kind}: status agreement {(st == st_cv).mean():.4f}, within 0.01 px {np.mean(d <= 0.01):.4f}, '
              f'max {d.max():.3g}, tracked {both.sum()}')
        assert (st == st_cv).mean() >= 0.995 and np.mean(d <= 0.01) >= 0.995


@pytest.mark.skipif(cv2 is None, reason='cv2 not importable')
def test_stereo_match_vs_port(ctx752, frames752):
    f0, _ = frames752
    ctx752.upload(f0.cam0_image, f0.cam1_image)
    ctx752.build_pyramids()
    pts = _points(f0.cam0_image, 2000, 11)
    p1, ok = ctx752.stereo_match(pts)
    port = FrontEndPort(config_c2(), backend='cv2')
    p1_ref, ok_ref = port.stereo_match(f0.cam0_image, f0.cam1_image, pts)
    agree = (ok == ok_ref).mean()
    both = ok & ok_ref
    d = np.abs(p1 - p1_ref)[both].max(axis=1)
    print(f'stereo_match: inlier agreement {agree:.4f} ({ok.sum()} vs {ok_ref.sum()}), max |d| {d.max():.3g}')
    assert ok_ref.sum() > 500
    assert agree >= 0.995
    assert d.max() <= 0.01
    assert ctx752.stereo_match(np.zeros((0, 2), np.float32))[0].shape == (0, 2)      # empty input


def test_undistort_distort_vs_oracle(ctx752):
    cfg = config_c2()
    g = np.random.default_rng(3)
    pts = g.uniform([0, 0], [752, 480], (700, 2))
    R = FrontEndPort(cfg, backend='numpy').R0to1
    for f32 in (False, True):
        x = pts.astype(np.float32) if f32 else pts
        for K, D in ((cfg.cam0_intrinsics, cfg.cam0_distortion_coeffs), (cfg.cam1_intrinsics, cfg.cam1_distortion_coeffs)):
            for rot in (None, R):
                ref = cs.undistort_radtan(x, K, D, rot)
                got = ctx752.undistort_ex(K, D, x, rot, f32_io=f32)
                assert np.array_equal(got.astype(ref.dtype), ref)            # bit-exact, f32 and f64
                ref2 = cs.distort_radtan(ref, K, D)
                got2 = ctx752.distort_ex(K, D, ref, f32_io=f32)
                assert np.array_equal(got2.astype(ref2.dtype), ref2)


def test_camera_model_class_matches_reference_signature(ctx752):
    from image_processing import CameraModel
    cfg = config_c2()
    cm = CameraModel(cfg.cam0_intrinsics, 'radtan', cfg.cam0_distortion_coeffs, context=ctx752)
    pts = np.random.default_rng(1).uniform([0, 0], [752, 480], (50, 2)).astype(np.float32)
    u = cm.undistort_points(pts, cm.intrinsics, cm.distortion_model, cm.distortion_coeffs)
    assert u.dtype == np.float32 and u.shape == (50, 2)
    assert np.array_equal(u, cs.undistort_radtan(pts, cfg.cam0_intrinsics, cfg.cam0_distortion_coeffs))
    back = cm.distort_points(u, cm.intrinsics, cm.distortion_model, cm.distortion_coeffs)
    assert np.abs(back - pts).max() < 0.5          # 5 fixed-point iterations only: loose at the image corners
    assert cm.undistort_points([], cm.intrinsics, 'radtan', cm.distortion_coeffs) == []
    with pytest.raises(RuntimeError):
        cm.undistort_points(pts, cm.intrinsics, 'equidistant', cm.distortion_coeffs)


def test_gather_from_frame_store_places_every_image():
    """avb_store_* + avb_process_frame_gather with more images than one gather launch carries (2 x 130 > 256 pointers):
    level 0 of every stream and camera equals the frame the table pointed at, and the levels built from it equal the
    oracle pyramid."""
    from image_processing import _native
    w, h, S, n = 160, 128, 130, 7
    cfg = FrontEndConfig(grid_row=2, grid_col=2, pyramid_levels=2, width=w, height=h)
    g = np.random.default_rng(9)
    frames = g.integers(0, 256, (n, 2, h, w)).astype(np.uint8)
    store = _native.FrameStore(w, h, n)
    ctx = _native.Context(cfg, w, h, num_streams=S, use_graph=False)
    try:
        assert store.nbytes >= n * 2 * w * h and store.addr.shape == (n, 2)
        for k in range(n):
            store.upload(k, frames[k, 0], frames[k, 1][:, :], timestamp=float(k))
        with pytest.raises(RuntimeError):                       # frame index out of range
            store.upload(n, frames[0, 0], frames[0, 1])
        with pytest.raises(ValueError):                         # wrong shape
            store.upload(0, frames[0, 0][:, :-16], frames[0, 1])
        pick = g.integers(0, n, S)
        ctx.process_gather(store.addr[pick])
        for s in (0, 1, 63, 127, 128, 129):
            for cam in (0, 1):
                assert np.array_equal(ctx.download_level(cam, 0, s=s), frames[pick[s], cam]), (s, cam)
            ref = cs.build_pyramid(frames[pick[s], 1], 2)
            assert np.array_equal(ctx.download_level(1, 2, s=s), ref[2])
        with pytest.raises(ValueError):
            ctx.process_gather(store.addr[pick[:3]])
        bad = store.addr[pick].copy()
        bad[5, 1] += 1                                          # not 16-byte aligned
        with pytest.raises(RuntimeError, match='aligned'):
            ctx.process_gather(bad)
    finally:
        ctx.close()
        store.close()
