"""GPU parity of k_ransac (two-point RANSAC, BASELINE config C3) against oracle/ransac.py, through the C-ABI
(avb_two_point_ransac for flat lists, avb_process_frame for the fused frame).  The reference itself has no such stage
(all-ones stub, feature_tracker.py:135-136): with it switched off the pipeline tests of test_gpu_pipeline.py apply;
here the bar is mask-for-mask equality with the oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import ransac as rs
from frontend_config import FrontEndConfig, config_c3
from replay import run_stream
from oracle.pipeline_port import FrontEndPort
from ransac_cases import D_EUROC, D_EUROC1, K_EUROC, K_EUROC1, make_pixels, undistorted
from synth_euroc import SlidingTextureStream


@pytest.fixture(scope='module')
def ctx():
    from image_processing import _native
    c = _native.Context(FrontEndConfig(), 752, 480)
    yield c
    c.close()


CASES = [
    # n, outliers, kwargs
    (300, 40, {}),
    (2000, 500, {}),
    (5000, 2500, {}),                                          # more points than one pass of 256 threads x 8
    (257, 30, dict(translation=(0.0, 0.12, 0.0))),
    (64, 5, dict(translation=(0.0, 0.0, 0.4), gyro=(0.02, 0.01, -0.03))),
    (200, 4, dict(translation=(0.0, 0.0, 0.0), jitter_px=0.2, outlier_px=(8, 20))),      # pure rotation shortcut
    (100, 95, {}),                                             # hardly any consensus: hypotheses fall below 0.2 N
    (3, 0, {}), (2, 0, {}), (1, 0, {}),
]


@pytest.mark.parametrize('n,outliers,kw', CASES)
def test_flat_list_masks_equal_oracle(ctx, n, outliers, kw):
    total = agree = 0
    for cam, (K, D) in enumerate(((K_EUROC, D_EUROC), (K_EUROC1, D_EUROC1))):
        for frame_index in (1, 2, 77):
            a, b, R, truth = make_pixels(n, outliers, seed=100 * cam + frame_index, K=K, D=D, **kw)
            u1, u2 = undistorted(a, b, R, K, D)
            want = rs.two_point_ransac(u1, u2, K, 3.0, seed=5, frame_index=frame_index, cam=cam)
            got = ctx.two_point_ransac(K, D, a, b, R, threshold=3.0, seed=5, frame_index=frame_index, cam=cam)
            assert got.shape == want.shape
            total += n
            agree += int((got == want).sum())
            assert np.array_equal(got, want), f'cam {cam} frame {frame_index}: {(got != want).sum()} of {n} masks differ'
    print(f'n={n}: {agree}/{total} masks equal')


def test_flat_list_edge_cases(ctx):
    assert ctx.two_point_ransac(K_EUROC, D_EUROC, np.zeros((0, 2)), np.zeros((0, 2))).shape == (0,)
    # everything farther than 50 px: pre-rejection leaves < 3 points -> all outliers
    a, b, R, _ = make_pixels(50, 0, seed=1)
    got = ctx.two_point_ransac(K_EUROC, D_EUROC, a, b + 80.0, R)
    assert not got.any()
    # identical point sets: zero displacement -> degenerate branch keeps everything
    got = ctx.two_point_ransac(K_EUROC, D_EUROC, a, a, None)
    want = rs.two_point_ransac(*undistorted(a, a, None), K_EUROC, 3.0)
    assert np.array_equal(got, want) and got.all()
    with pytest.raises(ValueError):
        ctx.two_point_ransac(K_EUROC, D_EUROC, a, b[:-1])


def _frames(fe, stream, read):
    out = []
    run_stream(fe, stream, on_frame=lambda k, msg, fm: out.append(read(fm)))
    return out


def _run_pair(cfg, kw):
    """Same stream through the oracle port (RANSAC = oracle/ransac.py) and the CUDA front end (k_ransac in the graph)."""
    from image_processing import ImageProcessor
    fe = FrontEndPort(cfg, backend='cv2')
    ref = _frames(fe, SlidingTextureStream(**kw),
                  lambda fm: dict(ids=fe.ids.copy(), cell=fe.cell.copy(), life=fe.life.copy(), p0=fe.p0.astype(np.float64),
                                  p1=fe.p1.astype(np.float64), counters=dict(fe.num_features)))
    ip = ImageProcessor(cfg)

    def read(fm):
        cell, life, p0, p1 = ip.context.features(0)
        hdr, ids, meas = ip.context.result(0)
        return dict(ids=ids.copy(), cell=cell, life=life, p0=p0, p1=p1, counters=dict(ip.num_features))

    got = _frames(ip, SlidingTextureStream(**kw), read)
    ip.context.close()
    return ref, got


def _check(ref, got):
    dropped = 0
    for k, (f, r) in enumerate(zip(got, ref)):
        assert np.array_equal(f['ids'], r['ids']), f'frame {k}: feature ids differ'
        assert np.array_equal(f['cell'], r['cell']) and np.array_equal(f['life'], r['life']), f'frame {k}'
        if len(f['ids']):
            assert np.abs(f['p0'] - r['p0']).max() <= 0.01 and np.abs(f['p1'] - r['p1']).max() <= 0.01
        if k > 0:
            for key in ('before_tracking', 'after_tracking', 'after_matching', 'after_ransac'):
                assert f['counters'].get(key, -1) == r['counters'].get(key, -1), f'frame {k}: {key}'
            dropped += r['counters']['after_matching'] - r['counters']['after_ransac']
    return dropped


def test_pipeline_with_ransac_c2_gyro():
    """C2 geometry with the stage switched on; the stream's moving foreground patch and noise feed it real outliers."""
    cfg = FrontEndConfig(grid_row=6, grid_col=10, two_point_ransac=True, ransac_seed=3)
    kw = dict(n_frames=20, seed=7, sigma=2.2, drift=(1.6, 0.7), gyro=(0.01, -0.02, 0.03), noise=1.0,
              movers=[(200, 150, 40, -3.0, 4.0), (520, 300, 50, 5.0, -2.5)])
    ref, got = _run_pair(cfg, kw)
    dropped = _check(ref, got)
    print(f'C2 + RANSAC: 20 frames identical, {dropped} features rejected by RANSAC in total')
    assert dropped > 50, 'the moving patches must produce outliers for the stage to reject'
    assert len(got[-1]['ids']) > 150


def test_pipeline_with_ransac_c3():
    cfg = config_c3()
    assert cfg.two_point_ransac
    kw = dict(width=1280, height=1024, n_frames=6, seed=11, sigma=1.8, drift=(1.2, 0.9),
              movers=[(300, 300, 80, 4.0, -3.0), (900, 700, 100, -5.0, 2.0), (640, 200, 60, 2.0, 6.0)])
    ref, got = _run_pair(cfg, kw)
    dropped = _check(ref, got)
    print(f'C3 + RANSAC: features/frame {[len(f["ids"]) for f in got]}, {dropped} rejected by RANSAC')
    assert dropped > 50
    assert len(got[-1]['ids']) > 800


def test_ransac_off_is_the_reference():
    """ransac = 0 never launches the kernel: after_ransac == after_matching on every frame (reference stub)."""
    from image_processing import ImageProcessor
    ip = ImageProcessor(FrontEndConfig())
    msgs = run_stream(ip, SlidingTextureStream(n_frames=4, seed=2))
    assert ip.num_features['after_ransac'] == ip.num_features['after_matching']
    assert not ip.context.ransac
    ip.context.close()
