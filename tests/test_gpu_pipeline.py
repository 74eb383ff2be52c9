"""GPU parity, end to end: ImageProcessor.stereo_callback (one CUDA-graph launch per frame) against
(a) golden dumps of the UNMODIFIED reference and (b) the oracle port run live on the same stream."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from frontend_config import FrontEndConfig, config_c3
from replay import run_stream
from oracle.pipeline_port import FrontEndPort
from synth_euroc import SlidingTextureStream
from tools.make_golden import CASES


def _run_gpu(cfg, stream, use_graph=True):
    from image_processing import ImageProcessor
    ip = ImageProcessor(cfg, use_graph=use_graph)
    frames = []

    def on_frame(k, msg, fm):
        cell, life, p0, p1 = ip.context.features(0)
        hdr, ids, meas = ip.context.result(0)
        frames.append(dict(ids=ids.copy(), cell=cell, life=life, p0=p0, p1=p1, pub=meas.copy(),
                           has_new=int(hdr['has_new']), counters=dict(ip.num_features),
                           n_fast=int(hdr['n_fast'])))

    msgs = run_stream(ip, stream, on_frame=on_frame)
    nid = ip.next_feature_id
    ip.context.close()
    return msgs, frames, nid


def _compare(frames, ref_frames, tol=0.01):
    """ids / cells / lifetimes identical, positions within tol px.  Returns the worst deviation."""
    worst = 0.0
    for k, (f, r) in enumerate(zip(frames, ref_frames)):
        assert np.array_equal(f['ids'], r['ids']), f'frame {k}: feature ids differ'
        assert np.array_equal(f['cell'], r['cell']), f'frame {k}: grid cells differ'
        assert np.array_equal(f['life'], r['life']), f'frame {k}: lifetimes differ'
        if len(f['ids']):
            d = max(np.abs(f['p0'].astype(np.float64) - r['p0']).max(), np.abs(f['p1'].astype(np.float64) - r['p1']).max())
            worst = max(worst, d)
            assert d <= tol, f'frame {k}: tracked position off by {d} px'
            assert np.abs(f['pub'] - r['pub']).max() <= 1e-4, f'frame {k}: published normalized coords'
    return worst


@pytest.mark.parametrize('name', list(CASES))
def test_pipeline_matches_reference_golden(name, golden_dir):
    g = np.load(os.path.join(golden_dir, name + '.npz'))
    gr, gc, gmin, gmax, skw = CASES[name]
    cfg = FrontEndConfig(grid_row=gr, grid_col=gc, grid_min=gmin, grid_max=gmax)
    msgs, frames, nid = _run_gpu(cfg, SlidingTextureStream(**skw))
    n = int(g['n_frames'][0])
    assert len(msgs) == n
    ref = [dict(ids=g[f'f{k}_ids'], cell=g[f'f{k}_cell'], life=g[f'f{k}_life'], p0=g[f'f{k}_p0'],
                p1=g[f'f{k}_p1'], pub=g[f'f{k}_pub']) for k in range(n)]
    worst = _compare(frames, ref)
    exact = all(np.array_equal(f['p0'].astype(np.float64), r['p0']) and np.array_equal(f['p1'].astype(np.float64), r['p1'])
                for f, r in zip(frames, ref))
    print(f'{name}: worst position deviation {worst:.3g} px over {n} frames; bit-exact positions: {exact}')
    for k, (fm, f) in enumerate(zip(msgs, frames)):
        assert fm.timestamp == g[f'f{k}_ts'][0]
        assert [x.id for x in fm.features] == list(g[f'f{k}_pub_ids'])
        assert f['has_new'] == int(g[f'f{k}_u0_is_f64'][0]) or len(fm.features) == 0
        if k > 0:
            c = f['counters']
            assert [c.get('before_tracking', -1), c.get('after_tracking', -1), c.get('after_matching', -1),
                    c.get('after_ransac', -1)] == list(g[f'f{k}_counters'])
    assert nid == int(g['next_feature_id'][0])


def _port_frames(cfg, stream):
    fe = FrontEndPort(cfg, backend='cv2')
    out = []

    def on_frame(k, msg, fm):
        pub = np.array([[f.u0, f.v0, f.u1, f.v1] for f in fm.features], np.float64).reshape(-1, 4)
        out.append(dict(ids=fe.ids.copy(), cell=fe.cell.copy(), life=fe.life.copy(), p0=fe.p0.astype(np.float64),
                        p1=fe.p1.astype(np.float64), pub=pub))

    run_stream(fe, stream, on_frame=on_frame)
    return out, fe.next_feature_id


def test_pipeline_vs_port_c2_25_frames_with_and_without_graph():
    cfg = FrontEndConfig(grid_row=6, grid_col=10)
    kw = dict(n_frames=25, seed=7, sigma=2.2, drift=(1.6, 0.7), gyro=(0.01, -0.02, 0.03), noise=1.0)
    ref, ref_nid = _port_frames(cfg, SlidingTextureStream(**kw))
    for use_graph in (True, False):
        msgs, frames, nid = _run_gpu(cfg, SlidingTextureStream(**kw), use_graph=use_graph)
        worst = _compare(frames, ref)
        print(f'C2 25 frames graph={use_graph}: worst deviation {worst:.3g} px; features/frame {len(frames[-1]["ids"])}')
        assert nid == ref_nid
        assert len(frames[-1]['ids']) > 250


def test_pipeline_vs_port_c3_stress():
    cfg = config_c3()
    kw = dict(width=1280, height=1024, n_frames=5, seed=11, sigma=1.8, drift=(1.2, 0.9))
    ref, ref_nid = _port_frames(cfg, SlidingTextureStream(**kw))
    msgs, frames, nid = _run_gpu(cfg, SlidingTextureStream(**kw))
    worst = _compare(frames, ref)
    print(f'C3: worst deviation {worst:.3g} px; features/frame {[len(f["ids"]) for f in frames]}')
    assert nid == ref_nid
    assert len(frames[-1]['ids']) > 800


def test_empty_and_textureless_stream():
    """Flat images: no corners, nothing tracked, empty feature lists every frame (reference behaviour:
    empty inputs short-circuit, stereo_matcher.py:44-45, feature_tracker.py:97-98)."""
    from image_processing import ImageProcessor
    from synth_euroc import img_msg, stereo_msg
    cfg = FrontEndConfig()
    ip = ImageProcessor(cfg)
    flat = np.full((480, 752), 90, np.uint8)
    for k in range(3):
        ts = 5.0 + 0.05 * k
        fm = ip.stereo_callback(stereo_msg(ts, flat, flat, img_msg(ts, flat), img_msg(ts, flat)))
        assert fm.timestamp == ts and fm.features == []
    assert ip.next_feature_id == 0
    assert all(len(c) == 0 for c in ip.prev_features)
    ip.context.close()


def test_multi_stream_context_equals_single_streams():
    """S streams processed by the same launches give exactly what S separate contexts give."""
    from image_processing import _native
    cfg = FrontEndConfig(grid_row=6, grid_col=10)
    streams = [SlidingTextureStream(n_frames=6, seed=20 + s, sigma=2.0 + 0.3 * s, drift=(1.0 + 0.2 * s, 0.5)) for s in range(3)]
    singles = []
    for st in streams:
        c = _native.Context(cfg, 752, 480, num_streams=1)
        rows = []
        for k in range(st.n):
            f = st.frame(k)
            c.process([f.cam0_image], [f.cam1_image])
            hdr, ids, meas = c.result(0)
            rows.append((ids.copy(), meas.copy(), c.features(0)))
        singles.append(rows)
        c.close()
    c = _native.Context(cfg, 752, 480, num_streams=3)
    for k in range(6):
        fr = [st.frame(k) for st in streams]
        c.process([f.cam0_image for f in fr], [f.cam1_image for f in fr])
        for s in range(3):
            hdr, ids, meas = c.result(s)
            assert np.array_equal(ids, singles[s][k][0])
            assert np.array_equal(meas, singles[s][k][1])
            for a, b in zip(c.features(s), singles[s][k][2]):
                assert np.array_equal(a, b)
    c.close()


@pytest.mark.parametrize('name', ['ref_default_s0', 'ref_c1_gyro_s2', 'ref_c2_dark_s4'])
def test_staged_pipeline_matches_reference_golden(name, golden_dir):
    """mode='staged' = the reference's own orchestration over this package's stage classes (PyramidBuilder,
    StereoMatcher, FeatureInitializer, FeatureTracker, FeatureAdder, FeaturePruner, FeaturePublisher), every stage a
    separate libavb call.  Must give what the unmodified reference gave, incl. the u0/v0 dtype quirk (B11)."""
    from image_processing import ImageProcessor
    g = np.load(os.path.join(golden_dir, name + '.npz'))
    gr, gc, gmin, gmax, skw = CASES[name]
    cfg = FrontEndConfig(grid_row=gr, grid_col=gc, grid_min=gmin, grid_max=gmax)
    ip = ImageProcessor(cfg, mode='staged')
    frames = []

    def on_frame(k, msg, fm):
        feats = [f for cell in ip.prev_features for f in cell]
        cells = [c for c, cell in enumerate(ip.prev_features) for _ in cell]
        frames.append(dict(ids=np.array([f.id for f in feats], np.int64), cell=np.array(cells, np.int64),
                           life=np.array([f.lifetime for f in feats], np.int64),
                           p0=np.array([f.cam0_point for f in feats], np.float64).reshape(-1, 2),
                           p1=np.array([f.cam1_point for f in feats], np.float64).reshape(-1, 2),
                           pub=np.array([[f.u0, f.v0, f.u1, f.v1] for f in fm.features], np.float64).reshape(-1, 4),
                           u0_f64=len(fm.features) > 0 and np.asarray(fm.features[0].u0).dtype == np.float64,
                           counters=dict(ip.num_features)))

    msgs = run_stream(ip, SlidingTextureStream(**skw), on_frame=on_frame)
    n = int(g['n_frames'][0])
    ref = [dict(ids=g[f'f{k}_ids'], cell=g[f'f{k}_cell'], life=g[f'f{k}_life'], p0=g[f'f{k}_p0'],
                p1=g[f'f{k}_p1'], pub=g[f'f{k}_pub']) for k in range(n)]
    worst = _compare(frames, ref)
    print(f'staged {name}: worst position deviation {worst:.3g} px')
    for k, (fm, f) in enumerate(zip(msgs, frames)):
        assert [x.id for x in fm.features] == list(g[f'f{k}_pub_ids'])
        if k > 0:
            c = f['counters']
            assert [c.get('before_tracking', -1), c.get('after_tracking', -1), c.get('after_matching', -1),
                    c.get('after_ransac', -1)] == list(g[f'f{k}_counters'])
    assert ip.next_feature_id == int(g['next_feature_id'][0])
    ip.context.close()


def test_stage_classes_individually():
    """StereoMatcher / FastDetector / CameraModel / FeaturePublisher with the reference's call signatures."""
    from image_processing import (CameraModel, FastDetector, FeatureMetaData, FeaturePublisher, IMUProcessor, PyramidBuilder,
                                  StereoMatcher, create_context)
    from oracle import cv_semantics as cs
    cfg = FrontEndConfig()
    st = SlidingTextureStream(n_frames=2, seed=4, sigma=2.5)
    f0 = st.frame(0)
    ctx = create_context(cfg, 752, 480, use_graph=False)
    imu = IMUProcessor(cfg.T_imu_cam0, cfg.T_imu_cam1)
    cam = CameraModel(cfg.cam0_intrinsics, cfg.cam0_distortion_model, cfg.cam0_distortion_coeffs)
    pb = PyramidBuilder(cfg.win_size, cfg.pyramid_levels, f0.cam0_msg, f0.cam1_msg)
    pyr0, pyr1 = pb.create_image_pyramids()
    assert pyr0 is f0.cam0_image and pb.curr_cam1_pyramid is f0.cam1_image
    det = FastDetector(cfg.fast_threshold, lambda img: ctx)
    kps = det.detect(f0.cam0_image)
    exs, eys, ers = cs.fast_detect(f0.cam0_image, cfg.fast_threshold)
    assert [k.pt for k in kps] == [(float(x), float(y)) for x, y in zip(exs, eys)]
    assert [k.response for k in kps] == [float(r) for r in ers]
    sm = StereoMatcher(cfg.lk_params, imu, pb, cam, cfg.stereo_threshold)
    assert len(sm.stereo_match([])[0]) == 0
    pts0 = [k.pt for k in kps[:400]]
    p1, ok = sm.stereo_match(pts0)
    port = FrontEndPort(cfg, backend='cv2')
    ep1, eok = port.stereo_match(f0.cam0_image, f0.cam1_image, pts0)
    assert (ok == eok).mean() >= 0.995
    both = ok & eok
    assert both.sum() > 200 and np.abs(p1[both] - ep1[both]).max() <= 0.01
    pub = FeaturePublisher(cfg.cam0_intrinsics, cfg.cam0_distortion_model, cfg.cam0_distortion_coeffs,
                           cfg.cam1_intrinsics, cfg.cam1_distortion_model, cfg.cam1_distortion_coeffs)
    feats = []
    for i in np.nonzero(both)[0][:20]:
        fm = FeatureMetaData()
        fm.id, fm.lifetime, fm.cam0_point, fm.cam1_point = int(i), 1, pts0[i], p1[i]
        feats.append(fm)
    pub.cam0_curr_img_msg, pub.cam1_curr_img_msg, pub.curr_features = f0.cam0_msg, f0.cam1_msg, [feats]
    msg = pub.publish()
    assert msg.timestamp == f0.cam0_msg.timestamp and [m.id for m in msg.features] == [f.id for f in feats]
    eu0 = cs.undistort_radtan(np.array([f.cam0_point for f in feats]), cfg.cam0_intrinsics, cfg.cam0_distortion_coeffs)
    eu1 = cs.undistort_radtan(np.array([f.cam1_point for f in feats]), cfg.cam1_intrinsics, cfg.cam1_distortion_coeffs)
    got = np.array([[m.u0, m.v0, m.u1, m.v1] for m in msg.features], np.float64)
    assert np.abs(got[:, :2] - eu0).max() < 1e-12 and np.abs(got[:, 2:] - eu1).max() < 1e-6
    ctx.close()


def test_multi_stream_front_end_replay_equals_single_pipelines():
    """MultiStreamFrontEnd + replay (deterministic lock-step driver, one launch chain for all streams) publishes
    exactly what one ImageProcessor per stream publishes, IMU handling included."""
    from image_processing import ImageProcessor
    from multi_stream import MultiStreamFrontEnd, replay, stream_stats
    cfg = FrontEndConfig(grid_row=5, grid_col=6)
    kws = [dict(n_frames=6, seed=30 + s, sigma=2.0 + 0.4 * s, drift=(1.0 + 0.3 * s, -0.5), gyro=(0.01 * s, -0.02, 0.03),
                noise=0.5 * s) for s in range(3)]
    singles = []
    for kw in kws:
        ip = ImageProcessor(cfg)
        singles.append(run_stream(ip, SlidingTextureStream(**kw)))
        ip.context.close()
    fe = MultiStreamFrontEnd(cfg, 752, 480, 3)
    multi = replay(fe, [SlidingTextureStream(**kw) for kw in kws])
    for s in range(3):
        assert len(multi[s]) == len(singles[s]) == 6
        for a, b in zip(multi[s], singles[s]):
            assert a.timestamp == b.timestamp
            assert [(f.id, f.u0, f.v0, f.u1, f.v1) for f in a.features] == [(f.id, f.u0, f.v0, f.u1, f.v1) for f in b.features]
        assert stream_stats(s, multi[s])['features'] == sum(len(m.features) for m in singles[s])
    fe.close()


@pytest.mark.parametrize('name', ['ref_c2_s1', 'ref_c2_dark_s4'])
@pytest.mark.parametrize('wpf', ['1', '4'])
def test_both_lk_lane_mappings_match_reference_golden(wpf, name, golden_dir, monkeypatch):
    """The throughput mapping (one warp per feature, packed loads + dp2a) and the latency mapping (four warps per
    feature, parallel per-level templates) are the same arithmetic: both must reproduce the reference's dumps."""
    monkeypatch.setenv('AVB_WPF', wpf)
    g = np.load(os.path.join(golden_dir, name + '.npz'))
    gr, gc, gmin, gmax, skw = CASES[name]
    cfg = FrontEndConfig(grid_row=gr, grid_col=gc, grid_min=gmin, grid_max=gmax)
    msgs, frames, nid = _run_gpu(cfg, SlidingTextureStream(**skw))
    n = int(g['n_frames'][0])
    ref = [dict(ids=g[f'f{k}_ids'], cell=g[f'f{k}_cell'], life=g[f'f{k}_life'], p0=g[f'f{k}_p0'],
                p1=g[f'f{k}_p1'], pub=g[f'f{k}_pub']) for k in range(n)]
    worst = _compare(frames, ref)
    print(f'AVB_WPF={wpf} {name}: worst position deviation {worst:.3g} px')
    assert nid == int(g['next_feature_id'][0])


def test_pipeline_vs_port_lossy_stream_30_frames():
    """A stream that loses features on every frame (independently moving patches, heavy noise, fast drift + gyro): cells
    empty and refill, pruning by lifetime kicks in, new ids are handed out every frame.  30 frames equal the port."""
    cfg = FrontEndConfig(grid_row=6, grid_col=10)
    kw = dict(n_frames=30, seed=5, sigma=2.0, drift=(2.2, -1.4), gyro=(0.02, 0.01, -0.03), noise=2.5,
              movers=[(200, 150, 40, -3.0, 4.0), (520, 300, 50, 5.0, -2.5), (380, 240, 30, -6.0, -5.0)])
    ref, ref_nid = _port_frames(cfg, SlidingTextureStream(**kw))
    msgs, frames, nid = _run_gpu(cfg, SlidingTextureStream(**kw))
    worst = _compare(frames, ref)
    lost = sum(f['counters'].get('after_tracking', 0) - f['counters'].get('after_matching', 0) for f in frames[1:])
    print(f'lossy stream: 30 frames equal the port (worst {worst:.3g} px); {lost} features lost in stereo matching, '
          f'{nid} ids handed out')
    assert nid == ref_nid and lost > 0


def test_maximum_capacity_context_matches_port():
    """The largest table the library accepts: 16 x 16 cells x 32 slots = 8192 features per stream (AVB_MAX_CAP slots per
    cell, NMAX limit), 1280x1024, RANSAC on: exercises the eight-CTA k_finish, the one-warp LK mapping on a single
    stream and k_ransac's strided passes.  Three frames against the port."""
    cfg = FrontEndConfig(grid_row=16, grid_col=16, grid_min=16, grid_max=32, pyramid_levels=4, width=1280, height=1024,
                         two_point_ransac=True)
    kw = dict(width=1280, height=1024, n_frames=3, seed=3, sigma=1.6, drift=(1.1, 0.8), movers=[(600, 500, 90, 4.0, 3.0)])
    ref, ref_nid = _port_frames(cfg, SlidingTextureStream(**kw))
    msgs, frames, nid = _run_gpu(cfg, SlidingTextureStream(**kw))
    worst = _compare(frames, ref)
    print(f'max capacity: features/frame {[len(f["ids"]) for f in frames]}, worst {worst:.3g} px')
    assert nid == ref_nid
    assert len(frames[-1]['ids']) > 3000


def test_multi_stream_with_ransac_equals_single_pipelines():
    """RANSAC draws are keyed by each stream's own frame index and both camera rotations travel per stream: S streams
    in one context (process_frames, (2, S, 3, 3) rotations) publish what S separate pipelines publish."""
    from image_processing import ImageProcessor
    from multi_stream import MultiStreamFrontEnd, replay
    cfg = FrontEndConfig(grid_row=5, grid_col=6, two_point_ransac=True, ransac_seed=11)
    kws = [dict(n_frames=7, seed=40 + s, sigma=2.0 + 0.3 * s, drift=(1.2 + 0.3 * s, -0.6), gyro=(0.01 * s, -0.02, 0.02),
                noise=0.5, movers=[(250 + 60 * s, 200, 40, 4.0, -3.0 + s)]) for s in range(3)]
    singles, dropped = [], 0
    for kw in kws:
        ip = ImageProcessor(cfg)
        out = []

        def on_frame(k, msg, fm, ip=ip, out=out):
            nonlocal dropped
            out.append(fm)
            if k:
                dropped += ip.num_features['after_matching'] - ip.num_features['after_ransac']

        run_stream(ip, SlidingTextureStream(**kw), on_frame=on_frame)
        singles.append(out)
        ip.context.close()
    fe = MultiStreamFrontEnd(cfg, 752, 480, 3)
    multi = replay(fe, [SlidingTextureStream(**kw) for kw in kws])
    for s in range(3):
        for a, b in zip(multi[s], singles[s]):
            assert [(f.id, f.u0, f.v0, f.u1, f.v1) for f in a.features] == [(f.id, f.u0, f.v0, f.u1, f.v1) for f in b.features]
    fe.close()
    print(f'multi-stream + RANSAC: 3 streams x 7 frames identical to single pipelines; {dropped} features rejected by RANSAC')
    assert dropped > 0


def test_sweep_from_hbm_store_equals_single_pipelines():
    """Sequence x offset sweep (run.bat shape): two sequences cached once in HBM, three time-offset runs each, all six in
    lock-step through one context fed by the gather kernel.  Every run publishes exactly what a separate
    ImageProcessor publishes when it replays the host frames of that run (frames and IMU samples older than the run's
    start time dropped, dataset.py:206-214).  With estimator workers the sweep also returns one trajectory per run."""
    from image_processing import ImageProcessor
    from frontend_config import with_filter_fields
    from sweep import CachedSequence, run_sweep
    cfg = with_filter_fields(FrontEndConfig(grid_row=5, grid_col=6))
    kws = [dict(n_frames=12, seed=60 + q, sigma=2.0 + 0.5 * q, drift=(1.3, -0.4 - 0.3 * q), gyro=(0.02, -0.01 * q, 0.03),
                noise=0.5) for q in range(2)]
    offsets = (0.0, 0.12, 0.27)                                    # counted from the first IMU stamp (50 ms before frame 0,
                                                                   # dataset.py:203); frame period 0.05 s: runs start at frames 0, 2, 5
    seqs = [CachedSequence(SlidingTextureStream(**kw)) for kw in kws]
    assert seqs[0].store.nbytes >= 12 * 2 * 752 * 480
    steps = 6
    res = run_sweep(cfg, seqs, offsets, n_steps=steps, estimator_workers=2, warmup_steps=1)
    assert res['streams'] == 6 and res['steps'] == steps and len(res['trajectories']) == 6
    assert [r.first_frame for r in res['runs']] == [0, 2, 5, 0, 2, 5]
    # the same runs, one pipeline each, host frames
    want = []
    for q, kw in enumerate(kws):
        for off in offsets:
            st = SlidingTextureStream(**kw)
            frames = list(st.frames())
            start = next(iter(st.imu())).timestamp + off
            ip = ImageProcessor(cfg)
            out = []
            for kind, m in st.events():
                if m.timestamp < start:
                    continue
                if kind == 'imu':
                    ip.imu_callback(m)
                else:
                    out.append(ip.stereo_callback(m))
                    if len(out) == steps:
                        break
            ip.context.close()
            want.append(out)
    # replay the sweep once more without estimators, collecting what each run published
    from multi_stream import MultiStreamFrontEnd
    fe = MultiStreamFrontEnd(cfg, 752, 480, 6)
    runs = res['runs']
    pos = [r.first_imu for r in runs]
    for k in range(steps):
        refs, addrs = [], np.empty((6, 2), np.uint64)
        for s, r in enumerate(runs):
            seq = seqs[r.sequence]
            idx = r.first_frame + k
            j1 = int(np.searchsorted(seq.imu_rows[:, 0], seq.timestamps[idx], side='right'))
            for m in seq.imu_msgs[pos[s]:j1]:
                fe.imu_callback(s, m)
            pos[s] = j1
            refs.append(seq.frame_refs[idx])
            addrs[s] = seq.store.addr[idx]
        out = fe.step_from_store(addrs, refs)
        for s, (ts, ids, meas) in enumerate(out):
            fm = want[s][k]
            assert ts == fm.timestamp
            assert ids.tolist() == [f.id for f in fm.features]
            assert meas.tolist() == [[float(f.u0), float(f.v0), float(f.u1), float(f.v1)] for f in fm.features]
            assert res['features'][s, k] == len(ids)
    fe.close()
    for q in seqs:
        q.close()
    print(f'sweep: 6 offset runs x {steps} steps from the HBM store equal 6 single pipelines; '
          f'{int(res["features"].sum())} features published')


def test_live_front_end_plus_host_msckf_reproduces_the_reference_trajectory(golden_dir):
    """The whole VIO on this box: rendered room sequence -> CUDA front end (stereo_callback) -> host MSCKF (msckf.py).
    The front end must publish what the committed B200 dump holds (ids identical, coordinates to 1e-9: same kernels,
    same arithmetic, any box), and the filter fed live must land on the trajectory the UNMODIFIED reference filter
    produced from that dump (tests/golden/ref_msckf_traj.npz, tools/make_msckf_golden.py)."""
    from frontend_config import with_filter_fields
    from image_processing import ImageProcessor
    from msckf import MSCKF
    from tools.ate_parity import GRID, make_stream
    n = 64
    cfg = with_filter_fields(FrontEndConfig(**GRID))
    z = np.load(os.path.join(golden_dir, 'ate_gpu_features.npz'))
    want = np.load(os.path.join(golden_dir, 'ref_msckf_traj.npz'))['traj']
    ip, est = ImageProcessor(cfg), MSCKF(cfg, outfile=False)
    rows, k = [], 0
    for kind, m in make_stream(n).events():
        if kind == 'imu':
            ip.imu_callback(m)
            est.imu_callback(m)
            continue
        fm = ip.stereo_callback(m)
        ids = np.array([f.id for f in fm.features], np.int64)
        meas = np.array([[f.u0, f.v0, f.u1, f.v1] for f in fm.features], np.float64).reshape(-1, 4)
        assert np.array_equal(ids, z[f'f{k}_ids']), f'frame {k}: ids differ from the committed B200 dump'
        assert np.abs(meas - z[f'f{k}_meas']).max() < 1e-9
        r = est.feature_callback(fm)
        if r is not None:
            rows.append([k, r.timestamp, *r.pose.t, *est.imu_state.orientation])
        k += 1
    ip.context.close()
    got = np.array(rows)
    ref = want[want[:, 0] < n]
    assert len(got) == len(ref) > 30 and np.array_equal(got[:, 0], ref[:, 0])
    dp, dq = np.abs(got[:, 2:5] - ref[:, 2:5]).max(), np.abs(got[:, 5:9] - ref[:, 5:9]).max()
    print(f'live VIO on the GPU box: {len(got)} poses, max |dp| {dp:.3g} m, |dq| {dq:.3g} against the reference filter')
    assert dp < 1e-6 and dq < 1e-6


def test_run_sweep_cli_on_euroc_directories(tmp_path):
    """run_sweep.py (the reference's run.bat as one job): two rendered sequences written in the EuRoC layout, two time
    offsets each; every run leaves the reference's trajectory file (msckf.py:152-160) and a row of the metrics summary."""
    import csv
    from euroc import write_euroc
    from run_sweep import COLUMNS, main
    from synth_euroc import RoomSceneStream
    paths = []
    for q in range(2):
        st = RoomSceneStream(FrontEndConfig(), n_frames=34, seed=20 + q, tex_size=512, amp=0.8 + 0.2 * q)
        p = tmp_path / f'SEQ_{q:02d}'
        write_euroc(str(p), st)
        paths.append(str(p))
    out = tmp_path / 'results'
    assert main(['--path', *paths, '--offsets', '0', '0.3', '--workers', '2', '--out', str(out)]) == 0
    rows = list(csv.DictReader(open(out / 'metrics_summary.csv')))
    assert [r['dataset'] for r in rows] == ['SEQ_00', 'SEQ_00', 'SEQ_01', 'SEQ_01'] and list(rows[0]) == COLUMNS
    assert [r['offset'] for r in rows] == ['0', '0.3', '0', '0.3']
    for r in rows:
        lines = open(out / 'txts' / f"output_{r['dataset']}_offset{r['offset']}.txt").read().strip().splitlines()
        assert len(lines) == int(r['poses'])
        if r['offset'] == '0':                                   # 28 frames per run: the filter publishes from frame 20 on
            assert int(r['frames']) == 28 and len(lines) >= 5 and float(r['ate_rmse_m']) < 0.05
            cols = lines[0].split()
            assert len(cols) == 8 and abs(sum(float(c) ** 2 for c in cols[4:]) - 1.0) < 1e-6


def test_page_locked_strided_and_pageable_host_images_give_the_same_messages():
    """The three intake paths of avb_process_frame (page-locked dense images: DMA straight from the caller's memory;
    pageable images: staged through the library's pinned block by two threads; strided views: row-wise staging), each
    through the split-graph steady state (cam0-only work under the cam1 copy): identical feature messages, 14 lossy frames."""
    import torch
    from image_processing import ImageProcessor
    from synth_euroc import img_msg, stereo_msg
    cfg = FrontEndConfig(grid_row=6, grid_col=10)
    kw = dict(n_frames=14, seed=8, sigma=2.0, drift=(1.9, -1.1), gyro=(0.02, 0.0, -0.02), noise=2.0)
    base = SlidingTextureStream(**kw)
    frames = [base.frame(k) for k in range(base.n)]

    def variant(kind):
        out = []
        for f in frames:
            if kind == 'pinned':
                a = torch.empty((2, 480, 752), dtype=torch.uint8).pin_memory().numpy()
                a[0], a[1] = f.cam0_image, f.cam1_image
                i0, i1 = a[0], a[1]
            elif kind == 'strided':
                a = np.zeros((2, 480, 800), np.uint8)
                a[0, :, 16:768], a[1, :, 16:768] = f.cam0_image, f.cam1_image
                i0, i1 = a[0, :, 16:768], a[1, :, 16:768]
                assert i0.strides == (800, 1)
            else:
                i0, i1 = f.cam0_image.copy(), f.cam1_image.copy()
            out.append(stereo_msg(f.timestamp, i0, i1, img_msg(f.timestamp, i0), img_msg(f.timestamp, i1)))
        return out

    got = {}
    for kind in ('pageable', 'pinned', 'strided'):
        st = SlidingTextureStream(**kw)
        msgs_in = variant(kind)
        st.frames = lambda m=msgs_in: iter(m)
        ip = ImageProcessor(cfg)
        msgs = run_stream(ip, st)
        got[kind] = [(m.timestamp, [(f.id, f.u0, f.v0, f.u1, f.v1) for f in m.features]) for m in msgs]
        ip.context.close()
    assert len(got['pageable']) == 14 and len(got['pageable'][-1][1]) > 250
    assert got['pinned'] == got['pageable'] and got['strided'] == got['pageable']


def test_submit_images_then_process_submitted_equals_process_frame():
    """avb_submit_images + avb_process_submitted (the frame in two halves, so that the caller integrates the gyro window
    while the images are on the bus) against the one-call path, 3 streams x 8 frames incl. frame 0 (the general intake) and
    1 stream x 8 frames (the split-graph intake); and the call-order errors."""
    from image_processing import _native
    from test_gpu_many_streams import _rotations, _snapshot
    cfg = FrontEndConfig(grid_row=6, grid_col=10)
    kw = dict(n_frames=12, seed=9, sigma=2.0, drift=(1.6, 0.9), gyro=(0.02, -0.01, 0.02), noise=1.5)
    st = SlidingTextureStream(**kw)
    frames = [st.frame(k) for k in range(st.n)]
    st.frames = lambda: iter(frames)
    Rs = _rotations(cfg, st)
    for S in (3, 1):
        a = _native.Context(cfg, 752, 480, num_streams=S)
        b = _native.Context(cfg, 752, 480, num_streams=S)
        try:
            for k in range(8):
                idx = [k + s for s in range(S)]
                i0, i1 = [frames[i].cam0_image for i in idx], [frames[i].cam1_image for i in idx]
                R0 = None if k == 0 else np.stack([Rs[i][0] for i in idx])
                R1 = None if k == 0 else np.stack([Rs[i][1] for i in idx])
                a.submit_images(i0, i1)
                if k == 3:                                  # call order: nothing else may start a frame in between
                    with pytest.raises(RuntimeError, match='submitted'):
                        a.submit_images(i0, i1)
                    with pytest.raises(RuntimeError, match='submitted'):
                        a.process_staged(R0, R1)
                a.process_submitted(R0, R1)
                b.process(i0, i1, R0, R1)
                for s in range(S):
                    x, y = _snapshot(a, s), _snapshot(b, s)
                    assert x['hdr'] == y['hdr'], (S, k, s)
                    for key in ('ids', 'cell', 'life', 'p0', 'p1', 'meas'):
                        assert np.array_equal(x[key], y[key]), (S, k, s, key)
            with pytest.raises(RuntimeError, match='no images submitted'):
                a.process_submitted(None, None)
            assert _snapshot(a, 0)['hdr'][0] > 250
        finally:
            a.close()
            b.close()
