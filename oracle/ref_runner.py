"""ORACLE (test infrastructure): runs the UNMODIFIED reference (oracle/_ref, copied verbatim from /root/reference/src by
tools/make_oracle_ref.sh) in a process of its own.  Only tests/, bench.py's reference / cpu_baseline legs and the tools
may execute this; nothing in the product package imports it.

    python oracle/ref_runner.py bench  --workload c2 --frames 60 --warmup 5 [--threads T] [--procs P]
        times ImageProcessor.stereo_callback of the reference (image_processing/pipeline.py:46-150) on the synthetic
        workload stream, frames pre-decoded in RAM, IMU delivered by the synchronous driver; one JSON line.
        --threads T  cv2.setNumThreads(T) (0 = leave cv2's default: all cores)
        --procs P    P processes side by side (each its own stream, started together), aggregate frames/s
    python oracle/ref_runner.py vio --frames 64
        the reference's thread harness modules/vio.py (VIO(config, img_q, imu_q)) with ITS MSCKF, consuming the
        image_processing package named by --front-end (`b200`: this repo's CUDA drop-in, `ref`: the reference's own);
        queues fed in the deterministic order of the synchronous driver; trajectory rows as JSON.

A separate process because the reference's module names (`image_processing`, `msckf`, `config`, `utils`) collide with the
drop-in package's on purpose: sys.path order decides which one a process sees.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, 'oracle', '_ref')
PKG = os.path.join(ROOT, 'uav-airvision_b200')
sys.dont_write_bytecode = True


def _paths(front_end='ref'):
    """`ref`: oracle/_ref FIRST (image_processing, msckf, config, utils, feature all resolve to the reference); the package
    directory stays behind it for synth_euroc / frontend_config (plain data + the synthetic stream).  `b200`: the drop-in
    `image_processing` package first, everything else of the reference behind it."""
    if not os.path.isdir(REF):
        raise SystemExit('oracle/_ref missing: run tools/make_oracle_ref.sh where /root/reference exists')
    for p in (PKG, ROOT, REF):
        while p in sys.path:
            sys.path.remove(p)
    sys.path[:0] = [REF, PKG, ROOT] if front_end == 'ref' else [PKG, REF, ROOT]


def ref_config(workload):
    """The reference's own ConfigEuRoC with the front-end fields of the BASELINE workload (frontend_config.py)."""
    from config import ConfigEuRoC
    from frontend_config import config_c2, config_c3
    src = {'c2': config_c2, 'c3': config_c3}[workload]()
    cfg = ConfigEuRoC()
    for k in ('grid_row', 'grid_col', 'grid_num', 'grid_min_feature_num', 'grid_max_feature_num', 'fast_threshold',
              'pyramid_levels', 'patch_size', 'win_size', 'cam0_intrinsics', 'cam1_intrinsics', 'cam0_resolution',
              'cam1_resolution'):
        setattr(cfg, k, getattr(src, k))
    cfg.lk_params = dict(cfg.lk_params, winSize=src.win_size, maxLevel=src.pyramid_levels)
    return cfg


def stream_kwargs(workload, seed_offset=0):
    kw = {'c2': dict(width=752, height=480, seed=7, sigma=2.2, drift=(1.6, 0.7), gyro=(0.01, -0.02, 0.03), noise=1.0),
          'c3': dict(width=1280, height=1024, seed=11, sigma=1.8, drift=(1.2, 0.9))}[workload]
    return dict(kw, seed=kw['seed'] + seed_offset)


def _bench_one(workload, n_frames, warmup, threads, seed_offset, barrier=None, budget_s=120.0):
    import cv2
    if threads > 0:
        cv2.setNumThreads(threads)
    from image_processing import ImageProcessor
    import image_processing
    assert os.path.realpath(image_processing.__file__).startswith(os.path.realpath(REF)), image_processing.__file__
    from synth_euroc import SlidingTextureStream
    st = SlidingTextureStream(n_frames=n_frames, **stream_kwargs(workload, seed_offset))
    frames = [st.frame(k) for k in range(n_frames)]
    st.frames = lambda: iter(frames)
    events = list(st.events())
    ip = ImageProcessor(ref_config(workload))
    times, feats = [], []
    if barrier is not None:
        barrier.wait()
    t_start = time.perf_counter()
    for kind, msg in events:
        if kind == 'imu':
            ip.imu_callback(msg)
            continue
        t0 = time.perf_counter()
        fm = ip.stereo_callback(msg)
        times.append(time.perf_counter() - t0)
        feats.append(len(fm.features))
        if time.perf_counter() - t_start > budget_s:
            break
    w = min(warmup + 1, max(len(times) - 1, 1))               # frame 0 + warm-up frames are not timed
    return dict(times=times, feats=feats, timed_from=w, cv2_threads=cv2.getNumThreads(), cv2=cv2.__version__)


def _bench_worker(args):
    workload, n_frames, warmup, threads, seed_offset, barrier = args
    _paths('ref')
    return _bench_one(workload, n_frames, warmup, threads, seed_offset, barrier)


def bench(a):
    _paths('ref')
    if a.procs <= 1:
        r = _bench_one(a.workload, a.frames, a.warmup, a.threads, 0)
        t = r['times'][r['timed_from']:]
        out = dict(kind='_ref', frames_timed=len(t), seconds=sum(t), fps=len(t) / sum(t) if t else 0.0,
                   frame0_ms=1e3 * r['times'][0], ms_median=1e3 * sorted(t)[len(t) // 2] if t else None,
                   features_per_s=sum(r['feats'][r['timed_from']:]) / sum(t) if t else 0.0,
                   features_last=r['feats'][-1], cv2_threads=r['cv2_threads'], cv2=r['cv2'], procs=1)
    else:
        import multiprocessing as mp
        ctx = mp.get_context('spawn')
        with ctx.Manager() as man:
            bar = man.Barrier(a.procs)
            with ctx.Pool(a.procs) as pool:
                rs = pool.map(_bench_worker, [(a.workload, a.frames, a.warmup, a.threads, i, bar) for i in range(a.procs)])
        # every process runs the same number of frames from a common start: aggregate = total timed frames / slowest span
        spans = [sum(r['times'][r['timed_from']:]) for r in rs]
        n = [len(r['times']) - r['timed_from'] for r in rs]
        out = dict(kind='_ref', frames_timed=sum(n), seconds=max(spans), fps=sum(n) / max(spans),
                   fps_sum_of_rates=sum(k / s for k, s in zip(n, spans)), procs=a.procs, cv2_threads=rs[0]['cv2_threads'],
                   cv2=rs[0]['cv2'], features_per_s=sum(sum(r['feats'][r['timed_from']:]) for r in rs) / max(spans))
    print(json.dumps(out), flush=True)


class _SyncQueue:
    """queue.Queue whose producer can wait until the consumer has finished everything put so far (the consumer is back
    in get() on an empty queue).  The reference's threads (modules/vio.py:26-53) loop `msg = q.get(); handle(msg)`."""

    def __init__(self):
        import queue
        import threading
        self.q = queue.Queue()
        self.cv = threading.Condition()
        self.puts = 0
        self.gets = 0                    # get() calls started

    def put(self, x):
        with self.cv:
            self.puts += 1
        self.q.put(x)

    def get(self):
        with self.cv:
            self.gets += 1
            self.cv.notify_all()
        return self.q.get()

    def wait_idle(self, timeout=120.0):
        with self.cv:
            if not self.cv.wait_for(lambda: self.gets == self.puts + 1, timeout):
                raise RuntimeError('consumer thread did not drain its queue')


def _render_range(job):
    """Pool worker: frames lo .. hi-1 of the rendered room sequence (tools/ate_parity.make_stream) as uint8 arrays."""
    n, lo, hi = job
    for p_ in (PKG, ROOT, os.path.join(ROOT, 'tools')):
        if p_ not in sys.path:
            sys.path.insert(0, p_)
    from ate_parity import make_stream
    st = make_stream(n)
    return lo, [(f.cam0_image, f.cam1_image) for f in (st.frame(k) for k in range(lo, hi))]


def _prerender(stream, n, procs):
    """Renders the n frames of `stream` on `procs` processes (a frame takes ~0.17 s on one core) and makes
    stream.frames() replay them."""
    import multiprocessing as mp
    from synth_euroc import img_msg, stereo_msg
    procs = max(1, min(procs, n))
    step = -(-n // procs)
    with mp.get_context('spawn').Pool(procs) as pool:
        parts = pool.map(_render_range, [(n, lo, min(lo + step, n)) for lo in range(0, n, step)])
    imgs = [im for _, chunk in sorted(parts, key=lambda t: t[0]) for im in chunk]

    def frames():
        for k, (i0, i1) in enumerate(imgs):
            ts = stream.t0 + k / stream.rate
            yield stereo_msg(ts, i0, i1, img_msg(ts, i0), img_msg(ts, i1))
    stream.frames = frames


def vio(a):
    """modules/vio.VIO unchanged, its MSCKF unchanged; `image_processing` = the package under test."""
    import contextlib
    import io
    import tempfile
    import numpy as np
    _paths(a.front_end)
    if a.front_end == 'b200':
        # `msckf` must be the REFERENCE's (the package directory also holds the repo's own msckf.py)
        import importlib.util
        spec = importlib.util.spec_from_file_location('msckf', os.path.join(REF, 'msckf.py'))
        mod = importlib.util.module_from_spec(spec)
        sys.modules['msckf'] = mod
        spec.loader.exec_module(mod)
    import image_processing
    import msckf
    from modules.vio import VIO
    want_ip = PKG if a.front_end == 'b200' else REF
    assert os.path.realpath(image_processing.__file__).startswith(os.path.realpath(want_ip)), image_processing.__file__
    assert os.path.realpath(msckf.__file__).startswith(os.path.realpath(REF)), msckf.__file__
    from config import ConfigEuRoC
    sys.path.insert(0, os.path.join(ROOT, 'tools'))
    from ate_parity import GRID, make_stream
    cfg = ConfigEuRoC()
    cfg.grid_row, cfg.grid_col = GRID['grid_row'], GRID['grid_col']
    cfg.grid_num = cfg.grid_row * cfg.grid_col
    cfg.grid_min_feature_num, cfg.grid_max_feature_num = GRID['grid_min'], GRID['grid_max']
    stream = make_stream(a.frames)
    if a.render_procs > 1:
        _prerender(stream, a.frames, a.render_procs)
    os.environ['DATASET_NAME'], os.environ['TIME_OFFSET'] = 'live_vio_' + a.front_end, '0'
    work = tempfile.mkdtemp(prefix='vio_')
    cwd = os.getcwd()
    os.chdir(work)                                            # the reference MSCKF appends to results/txts/... in the cwd
    rows, feats = [], []
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            img_q, imu_q = _SyncQueue(), _SyncQueue()
            v = VIO(cfg, img_q, imu_q)
            v.feature_queue = feat_q = _SyncQueue()           # harness only: lets the driver wait for the filter
            orig = v.msckf.feature_callback

            def tap(fm, orig=orig):
                r = orig(fm)
                feats.append((fm.timestamp, [f.id for f in fm.features],
                              [[float(f.u0), float(f.v0), float(f.u1), float(f.v1)] for f in fm.features]))
                if r is not None:
                    st = v.msckf.state_server.imu_state
                    rows.append([len(feats) - 1, float(r.timestamp), *np.asarray(r.pose.t, float), *np.asarray(st.orientation, float)])
                return r
            v.msckf.feature_callback = tap
            v.start()
            t0 = time.perf_counter()
            for kind, msg in stream.events():
                if kind == 'imu':
                    imu_q.put(msg)
                    continue
                imu_q.wait_idle()                             # every IMU sample up to this frame has reached both consumers
                img_q.put(msg)
                img_q.wait_idle()
                feat_q.wait_idle()
            wall = time.perf_counter() - t0
            img_q.put(None)
            imu_q.put(None)
            for th in (v.img_thread, v.imu_thread, v.vio_thread):
                th.join(timeout=30)
    finally:
        os.chdir(cwd)
    out = dict(front_end=a.front_end, frames=len(feats), poses=len(rows), wall_s=wall, traj=[] if a.no_traj else rows,
               image_processing=os.path.relpath(image_processing.__file__, ROOT), msckf=os.path.relpath(msckf.__file__, ROOT))
    if a.dump:
        np.savez_compressed(a.dump, traj=np.array(rows), n_frames=np.array([len(feats)]),
                            **{f'f{k}_ids': np.array(f[1], np.int64) for k, f in enumerate(feats)},
                            **{f'f{k}_meas': np.array(f[2], np.float64).reshape(-1, 4) for k, f in enumerate(feats)},
                            **{f'f{k}_ts': np.array([f[0]]) for k, f in enumerate(feats)})
    print(json.dumps(out), flush=True)


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    sub = ap.add_subparsers(dest='cmd', required=True)
    b = sub.add_parser('bench')
    b.add_argument('--workload', default='c2', choices=['c2', 'c3'])
    b.add_argument('--frames', type=int, default=60)
    b.add_argument('--warmup', type=int, default=5)
    b.add_argument('--threads', type=int, default=0)
    b.add_argument('--procs', type=int, default=1)
    v_ = sub.add_parser('vio')
    v_.add_argument('--frames', type=int, default=64)
    v_.add_argument('--front-end', default='b200', choices=['b200', 'ref'])
    v_.add_argument('--dump', default=None)
    v_.add_argument('--render-procs', type=int, default=1, help='render the sequence ahead on this many processes')
    v_.add_argument('--no-traj', action='store_true', help='leave the trajectory rows out of the JSON line (use --dump)')
    a = ap.parse_args()
    (bench if a.cmd == 'bench' else vio)(a)
