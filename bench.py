#!/usr/bin/env python
"""bench.py -- stereo frames/s and tracked features/s of the B200 image front end.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload c2|c3|c5]

A "step" is one stereo frame of one stream through the whole front end (pyramid build, temporal KLT, stereo
KLT + filters, FAST + grid ranking, new-feature stereo match, prune, publish).  At N GPUs every rank owns one
independent synthetic EuRoC-format stream (no collective on the data path: "scaling": "weak"); rank 0 prints ONE
JSON line.

  value     DEVICE-RESIDENT frames/s: the whole frame sequence already lies in HBM, K dependent frames are enqueued on
            the context's stream, CUDA events on that stream, max over ranks.  No host result is awaited per frame, so
            this is NOT the SURVEY 8(d) metric ("frames whose feature_msg is fully materialised on the host"): it
            explains `e2e`.
  e2e       the 8(d) metric: the same frames through the public API, ImageProcessor.stereo_callback(stereo_msg) with
            HOST numpy images: host->pinned copy, H2D, the CUDA-graph frame, D2H of the result block and construction
            of the FeatureMeasurement list are all inside the timed region
  repeats   a timed region of K steps is repeated R times (R chosen so that the regions add up to >= 0.25 s) and the
            MEDIAN region is reported; `steps` stays K
  roofline  the frame's kernel chain against the HBM copy peak of MEASURED_PEAKS.json
  multi_stream / c4_weak   BASELINE config C4 (64 time-offset runs): 64 runs in total sharded over the GPUs (strong) and
            64 runs per GPU (weak); the last frame of the first and last run is checked against single-stream contexts
  c3, c5    compact legs of the other BASELINE configurations (N = 1 only)
  cpu_baseline / --impl reference
            the UNMODIFIED reference front end (oracle/_ref, copied from /root/reference/src by
            tools/make_oracle_ref.sh) on the box's host cores, same frames: one stream with cv2's default threads, one
            stream with cv2.setNumThreads(1), and os.cpu_count() single-thread processes side by side (aggregate)

Inputs are synthetic (seeded sliding-texture stereo + IMU, synth_euroc.py): >= 200 distinct frames (> the 126 MB L2)
rendered once; longer timed sequences walk them forth and back (continuous motion), so between two reads of a frame
more than an L2 of other frames has passed.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, 'uav-airvision_b200')):
    if p not in sys.path:
        sys.path.insert(0, p)
sys.dont_write_bytecode = True

METRIC = 'stereo_frames_per_s'
UNIT = 'frames/s'
DTYPE = 'u8/int32 fixed-point + f32 (LK), f64 (undistort)'
MIN_REGION_S = 0.25
REF_RUNNER = os.path.join(ROOT, 'oracle', 'ref_runner.py')


def workload(name):
    from frontend_config import config_c2, config_c3
    if name == 'c2':
        return config_c2(), dict(width=752, height=480, seed=7, sigma=2.2, drift=(1.6, 0.7),
                                 gyro=(0.01, -0.02, 0.03), noise=1.0), \
            'C2: single synthetic EuRoC-format stereo stream 752x480, grid 6x10 x max 5 = 300 features, ' \
            '4-level pyramid, KLT 15x15'
    if name == 'c3':
        return config_c3(), dict(width=1280, height=1024, seed=11, sigma=1.8, drift=(1.2, 0.9)), \
            'C3: 1280x1024 stereo, grid 10x10 x max 20 = 2000 features, 5-level pyramid (intrinsics scaled with ' \
            'the image), batched two-point RANSAC on (not a reference stage: the reference stubs it out; oracle/ransac.py defines it)'
    raise SystemExit(f'unknown workload {name}')


def line_config(wname, world):
    """`config` of the JSON line: identical for the repo arm and the reference arm (static description only; measured
    quantities live in other keys of the line)."""
    return {'workload': wname, 'streams_per_gpu': 1, 'streams_total': world,
            'frames': 'synthetic sliding-texture stereo + IMU, >= 200 distinct frames (> L2) walked forth and back',
            'timing': f'median of `repeats` regions of `steps` frames each (regions add up to >= {MIN_REGION_S} s)'}


def stage_bytes(w, h, max_level, S):
    """Algorithmic HBM bytes of the image-scan stages for S streams: the input copy reads and writes
    the block, every pyramid level l reads level l-1 and writes level l (both cameras), FAST reads cam0 once."""
    px = [w * h]
    lw, lh = w, h
    for _ in range(max_level):
        lw, lh = (lw + 1) // 2, (lh + 1) // 2
        px.append(lw * lh)
    return {'input_copy': 2 * S * 2 * px[0], 'pyramid': 2 * S * sum(px[l - 1] + px[l] for l in range(1, max_level + 1)),
            'fast': S * px[0]}


def algorithmic_bytes_per_frame(w, h, max_level, n_feat):
    """read both u8 inputs once + write pyramid levels 1..L once + 64 B per feature."""
    px, tot = w * h, 0
    lw, lh = w, h
    for _ in range(max_level):
        lw, lh = (lw + 1) // 2, (lh + 1) // 2
        tot += lw * lh
    return 2 * (px + tot) + 64 * n_feat


class ClockSampler:
    """nvidia-smi clocks + throttle reasons while the timed regions run (B200_PROFILING.md, clocks line)."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(index), f'--query-gpu={self.Q}',
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(',')]))

    def summary(self, windows):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                pass
        rows = [r for t, r in self.rows if any(a - 0.05 <= t <= b + 0.05 for a, b in windows)] or \
               [r for _, r in self.rows]
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        if not sm:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0}
        return {'sm_mhz': float(np.median(sm)), 'sm_max_mhz': float(max(mx)), 'reasons': sorted(reasons),
                'samples': len(sm)}


class ForthAndBack:
    """`total` frames from `n` distinct rendered ones: 0, 1, .., n-1, n-2, .., 1, 0, 1, ..  The sliding texture then
    moves forth and back -- continuous motion, every frame tracks from its predecessor -- while time stamps and IMU
    samples keep increasing as in the base stream."""

    def __init__(self, base, frames, total):
        self.base, self.src, self.n = base, frames, int(total)
        self.w, self.h, self.rate = base.w, base.h, base.rate
        nb = len(frames)
        period = max(2 * (nb - 1), 1)
        m = np.arange(self.n) % period
        self.index = np.where(m < nb, m, period - m).astype(np.int64)

    def frame(self, k):
        from synth_euroc import img_msg, stereo_msg
        f = self.src[int(self.index[k])]
        ts = self.base.t0 + k / self.base.rate
        return stereo_msg(ts, f.cam0_image, f.cam1_image, img_msg(ts, f.cam0_image), img_msg(ts, f.cam1_image))

    def frames(self):
        return (self.frame(k) for k in range(self.n))

    def imu(self):
        from synth_euroc import imu_msg
        b = self.base
        n_imu = int(np.floor((self.n - 1) / b.rate * b.imu_rate)) + 1
        acc = np.array([0.0, 0.0, 9.81])
        lead = int(0.05 * b.imu_rate)
        for j in range(-lead, n_imu + 1):
            yield imu_msg(b.t0 + j / b.imu_rate, b.gyro.copy(), acc.copy())

    def events(self):
        imu_it = iter(self.imu())
        pending = next(imu_it, None)
        for f in self.frames():
            while pending is not None and pending.timestamp <= f.timestamp:
                yield 'imu', pending
                pending = next(imu_it, None)
            yield 'stereo', f


def rotations_for(cfg, stream):
    """(cam0_R_p_c, cam1_R_p_c) per frame exactly as the pipeline's IMUProcessor produces them with the synchronous
    driver; the second one is read by the RANSAC kernel only (workload c3)."""
    from image_processing import IMUProcessor
    imu = IMUProcessor(cfg.T_imu_cam0, cfg.T_imu_cam1)
    out, prev = [], None
    for kind, msg in stream.events():
        if kind == 'imu':
            imu.imu_callback(msg)
            continue
        imu.cam0_prev_img_msg, imu.cam0_curr_img_msg = prev, msg.cam0_msg
        out.append((np.eye(3), np.eye(3)) if prev is None else tuple(imu.integrate_imu_data()))
        prev = msg.cam0_msg
    return out


def repeats_for(K, est_ms_per_step, cap_frames=3000):
    """Number of timed regions of K steps so that they add up to >= MIN_REGION_S (at most 50, at most cap_frames frames)."""
    region = K * est_ms_per_step * 1e-3
    if region >= MIN_REGION_S:
        return 1
    return int(max(1, min(50, math.ceil(MIN_REGION_S / region), cap_frames // max(K, 1))))


# ---- the reference on the host cores (oracle/_ref through oracle/ref_runner.py, its own process) -----------------------
def ref_available():
    return os.path.isdir(os.path.join(ROOT, 'oracle', '_ref'))


def run_ref(wl, frames, warmup, threads=0, procs=1, timeout=900):
    """One oracle/ref_runner.py bench run -> its JSON (None when it failed)."""
    cmd = [sys.executable, REF_RUNNER, 'bench', '--workload', wl, '--frames', str(frames), '--warmup', str(warmup),
           '--threads', str(threads), '--procs', str(procs)]
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout)
        if r.returncode != 0:
            sys.stderr.write(r.stderr[-1500:])
            return None
        return json.loads(r.stdout.strip().splitlines()[-1])
    except Exception as e:                                  # pragma: no cover
        sys.stderr.write(f'reference runner failed: {e}\n')
        return None


def port_front_end(cfg, stream, n_frames, budget_s=30.0):
    """The oracle port (oracle/pipeline_port.py, cv2 backend) on the first frames of the stream: kept BESIDE the
    reference figure (its array-style bookkeeping is faster than the reference's per-keypoint Python loops)."""
    import cv2
    from oracle.pipeline_port import FrontEndPort
    fe = FrontEndPort(cfg, backend='cv2')
    times, feats = [], []
    t_start = time.perf_counter()
    k = 0
    for kind, msg in stream.events():
        if kind == 'imu':
            fe.imu_callback(msg)
            continue
        t0 = time.perf_counter()
        fm = fe.stereo_callback(msg)
        times.append(time.perf_counter() - t0)
        feats.append(len(fm.features))
        k += 1
        if k >= n_frames or (time.perf_counter() - t_start) > budget_s:
            break
    return times, feats, cv2.getNumThreads(), cv2.__version__


def cpu_variants(wl, frames, warmup, cores, with_one_thread=True, with_aggregate=True):
    """(i) one stream, cv2 default threads; (ii) one stream, one thread; (iii) `cores` single-thread processes, aggregate."""
    out = {}
    a = run_ref(wl, frames, warmup, threads=0)
    if a is None:
        return None
    out['one_stream_default_threads'] = {'value': a['fps'], 'cv2_threads': a['cv2_threads'], 'frames_timed': a['frames_timed'],
                                         'ms_median': a['ms_median'], 'frame0_ms': a['frame0_ms'],
                                         'tracked_features_per_s': a['features_per_s']}
    out['cv2'] = a['cv2']
    if with_one_thread:
        b = run_ref(wl, frames, warmup, threads=1)
        if b is not None:
            out['one_stream_one_thread'] = {'value': b['fps'], 'frames_timed': b['frames_timed'], 'ms_median': b['ms_median']}
    if with_aggregate and cores > 1:
        c = run_ref(wl, frames, warmup, threads=1, procs=cores)
        if c is not None:
            out['all_cores_one_stream_each'] = {'value': c['fps'], 'processes': c['procs'], 'frames_timed': c['frames_timed'],
                                                'tracked_features_per_s': c['features_per_s'],
                                                'note': f'{c["procs"]} single-thread processes side by side, each its own stream; '
                                                        'aggregate = timed frames of all / slowest span'}
    return out


def run_reference(args, rank, world):
    """--impl reference: the UNMODIFIED reference front end (oracle/_ref) on the host cores, rank 0 only.  N = 1: one
    stream with all the threads cv2 takes (what a user of the reference gets for this workload); N > 1: the repo arm
    runs N streams on N GPUs, so the like-for-like figure is the whole host: os.cpu_count() single-thread processes."""
    if rank != 0:
        return
    cfg, skw, wname = workload(args.workload)
    W, K = args.warmup, args.steps
    cores = os.cpu_count() or 1
    est_fps = 60.0 if args.workload == 'c2' else 4.0
    frames = int(min(W + 1 + K, W + 1 + max(8, est_fps * 60)))    # a leg stays within about a minute
    line = {'impl': 'reference', 'metric': METRIC, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': K, 'warmup': W,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'u8/int32 fixed-point + f32 (cv2)',
            'data': 'synthetic', 'config': line_config(wname, world), 'gpu_launches': 0}
    if not ref_available():
        # the oracle port is the stand-in (kind "port"): same cv2 calls at the same call sites
        from synth_euroc import SlidingTextureStream
        st = SlidingTextureStream(n_frames=frames, **skw)
        fr = [st.frame(k) for k in range(frames)]
        st.frames = lambda: iter(fr)
        times, feats, threads, cvv = port_front_end(cfg, st, frames, budget_s=120.0)
        w = min(W + 1, max(len(times) - 1, 1))
        tot = float(np.sum(times[w:]))
        val = (len(times) - w) / tot
        line.update(value=val, ms_per_step=1e3 / val, frames_timed=len(times) - w,
                    cpu_baseline={'value': val, 'unit': UNIT, 'cores': threads, 'kind': 'port',
                                  'sample': f'{len(times) - w} frames; oracle/_ref missing (tools/make_oracle_ref.sh was not run)'},
                    e2e={'value': val, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0})
        print(json.dumps(line), flush=True)
        return
    v = cpu_variants(args.workload, frames, W, cores)
    if v is None:
        print(json.dumps({'impl': 'reference', 'unavailable': 'oracle/ref_runner.py failed (see stderr)'}), flush=True)
        return
    one = v['one_stream_default_threads']
    agg = v.get('all_cores_one_stream_each')
    if world > 1 and agg is not None:
        val, used = agg['value'], agg['processes']
        what = f'{agg["processes"]} single-thread reference processes side by side (whole host)'
        feats = agg['tracked_features_per_s']
    else:
        val, used, what = one['value'], one['cv2_threads'], f'one stream, cv2 threads {one["cv2_threads"]}'
        feats = one['tracked_features_per_s']
    line.update(value=val, ms_per_step=1e3 / val if val > 0 else None, frames_timed=one['frames_timed'],
                tracked_features_per_s=feats, frame0_ms=one['frame0_ms'],
                cpu_baseline={'value': val, 'unit': UNIT, 'cores': used, 'kind': 'reference',
                              'sample': f'{what}; {one["frames_timed"]} consecutive frames of the workload stream after frame 0 + '
                                        f'{W} warm-up frames; unmodified reference (oracle/_ref: image_processing/pipeline.py:46-150), '
                                        f'cv2 {v["cv2"]}, host cores {cores}',
                              'variants': v},
                e2e={'value': val, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0})
    print(json.dumps(line), flush=True)


# ---- C5 ------------------------------------------------------------------------------------------------------------------
def _render_sequence(job):
    """Worker of the C5 leg: one rendered EuRoC-geometry sequence (images, IMU rows, ground truth) as arrays."""
    q, n_frames = job
    for p in (ROOT, os.path.join(ROOT, 'uav-airvision_b200')):
        if p not in sys.path:
            sys.path.insert(0, p)
    from frontend_config import config_c2
    from synth_euroc import RoomSceneStream
    st = RoomSceneStream(config_c2(), n_frames=n_frames, seed=100 + q, amp=0.7 + 0.08 * q, tex_size=1536)
    fr = list(st.frames())
    imu = np.array([[m.timestamp, *m.angular_velocity, *m.linear_acceleration] for m in st.imu()])
    gt = list(st.groundtruth())
    return {'q': q, 'ts': np.array([f.timestamp for f in fr]), 'img0': np.stack([f.cam0_image for f in fr]),
            'img1': np.stack([f.cam1_image for f in fr]), 'imu': imu,
            'gt_t': np.array([g.timestamp for g in gt]), 'gt_p': np.array([g.p for g in gt])}


class _ArraySequence:
    """Adapter: arrays of _render_sequence -> the frames()/imu()/groundtruth() shape CachedSequence reads."""

    def __init__(self, d):
        self.d, self.n = d, len(d['ts'])

    def frames(self):
        from synth_euroc import img_msg, stereo_msg
        d = self.d
        for k in range(self.n):
            t = float(d['ts'][k])
            yield stereo_msg(t, d['img0'][k], d['img1'][k], img_msg(t, d['img0'][k]), img_msg(t, d['img1'][k]))

    def imu(self):
        from synth_euroc import imu_msg
        return (imu_msg(float(r[0]), r[1:4], r[4:7]) for r in self.d['imu'])

    def groundtruth(self):
        from synth_euroc import gt_msg
        z = np.zeros(3)
        return (gt_msg(float(t), p, None, z, z, z) for t, p in zip(self.d['gt_t'], self.d['gt_p']))


def run_c5(B, n_seq, n_off, n_steps, n_workers, bound):
    """BASELINE config C5: sequences x time offsets, full front end + host MSCKF, sharded over the GPUs by sequence.
    Every rank renders its sequences (synthetic, EuRoC geometry), uploads each ONCE into an HBM frame store, and runs
    all offset runs of its sequences in lock-step through one context; estimators run in worker processes.  Returns the
    JSON line (rank 0) or None."""
    import multiprocessing as mp
    from frontend_config import config_c2, with_filter_fields
    from metrics import trajectory_metrics
    from multi_stream import shard_streams
    from sweep import CachedSequence, run_sweep
    rank, world, local, torch = B.rank, B.world, B.local, B.torch
    cfg = with_filter_fields(config_c2())
    mine = shard_streams(n_seq, world, rank)
    step_frames = 2                                           # offsets 0, 0.1 s, 0.2 s, ... (20 Hz frames)
    offsets = [o * step_frames / 20.0 for o in range(n_off)]
    n_frames = n_steps + step_frames * (n_off - 1) + 1
    cores = sorted(os.sched_getaffinity(0))
    t0 = time.perf_counter()
    with mp.get_context('spawn').Pool(min(len(mine), max(1, len(cores)))) as pool:
        rendered = pool.map(_render_sequence, [(q, n_frames) for q in mine])
    render_s = time.perf_counter() - t0
    t0 = time.perf_counter()
    seqs = [CachedSequence(_ArraySequence(d), device=local, name=f'room{d["q"]}') for d in rendered]
    torch.cuda.synchronize()
    upload_s = time.perf_counter() - t0
    store_bytes = sum(q.store.nbytes for q in seqs)
    # the ranks of a box share its cores unless each is bound to its own slice
    share = len(cores) if bound else len(cores) // max(1, int(os.environ.get('LOCAL_WORLD_SIZE', world)))
    # one estimator per core of the rank: the driver thread mostly waits on the GPU
    workers = n_workers if n_workers > 0 else max(1, share)
    warm = 2
    sampler = ClockSampler(local) if rank == 0 else None
    # front end alone (what the GPU side of the sweep sustains from the HBM store, results as arrays on the host)
    B.barrier()
    fe_only = run_sweep(cfg, seqs, offsets, device=local, n_steps=n_steps, warmup_steps=warm)
    B.barrier()
    t_a = time.perf_counter()
    full = run_sweep(cfg, seqs, offsets, device=local, n_steps=n_steps, estimator_workers=workers, warmup_steps=warm)
    B.barrier()
    t_b = time.perf_counter()
    assert np.array_equal(fe_only['features'], full['features'])
    S, timed = full['streams'], full['timed_steps']
    wall = B.max_over_ranks(full['wall_s'])
    fe_wall = B.max_over_ranks(fe_only['wall_s'])
    frames_total = B.sum_over_ranks(S * timed)
    feats_total = B.sum_over_ranks(float(full['features'][:, warm:].sum()))
    published = B.sum_over_ranks(float(sum(len(t) for t in full['trajectories'])))
    upload_max = B.max_over_ranks(upload_s)
    # accuracy of the offset-0 run of this rank's first sequence against its ground truth (sanity, not a parity gate)
    ate = None
    tr = full['trajectories'][0]
    if len(tr) > 10 and seqs[0].groundtruth is not None:
        m = trajectory_metrics(tr[:, 0], tr[:, 1:4], *seqs[0].groundtruth)
        ate = {'ate_rmse_m': m['ate_rmse_m'], 'path_m': m['path_m'], 'poses': int(len(tr))}
    img_bytes = 2 * seqs[0].width * seqs[0].height
    for q in seqs:
        q.close()
    clocks = sampler.summary([(t_a, t_b)]) if sampler is not None else None
    if rank != 0:
        return None
    busy = full['estimator']['worker_busy_s']
    return {
        'metric': METRIC, 'value': frames_total / wall, 'unit': UNIT, 'n_gpus': world, 'steps': timed, 'warmup': warm,
        'ms_per_step': 1e3 * wall / timed, 'higher_is_better': True, 'scaling': 'weak' if world <= n_seq else 'strong',
        'vs_baseline': None, 'dtype': 'u8/int32 fixed-point + f32 (LK), f64 (undistort, MSCKF)', 'data': 'synthetic',
        'config': {'workload': f'C5: {n_seq} EuRoC-geometry synthetic sequences x {n_off} time offsets '
                               f'= {n_seq * n_off} runs, full front end (C2 grid, 300 features) + host MSCKF, '
                               f'sharded by sequence over {world} GPU(s)',
                   'runs_per_gpu': S, 'offset_spacing_s': step_frames / 20.0,
                   'estimator_workers_per_gpu': workers, 'host_cores_per_rank': len(cores),
                   'host_cpus_bound': sorted(bound) if bound else None,
                   'frame_source': f'HBM frame store: each sequence uploaded once ({store_bytes / 1e6:.0f} MB on rank 0), '
                                   f'every offset run gathers its frames on the device'},
        'tracked_features_per_s': feats_total / wall,
        'poses_published': int(published),
        'front_end_only': {'value': frames_total / fe_wall, 'unit': UNIT, 'ms_per_step': 1e3 * fe_wall / timed,
                           'note': 'same sweep without estimators: gather + frame chain + result arrays on the host'},
        'estimator': {'worker_busy_s': [round(b, 3) for b in busy],
                      'ms_per_frame': 1e3 * float(np.sum(busy)) / max(full['estimator']['frames'], 1),
                      'errors': full['estimator'].get('errors') or None,
                      'note': 'host MSCKF (uav-airvision_b200/msckf.py) in worker processes; it bounds this leg'},
        'e2e': {'value': frames_total / (wall + upload_max), 'unit': UNIT,
                'h2d_bytes_per_step': int(len(mine) * n_frames * img_bytes / n_steps),
                'd2h_bytes_per_step': int(S * (48 + 300 * 40)),
                'note': 'includes the one-time upload of every sequence frame into the store (amortised over its offset runs); '
                        'rendering the synthetic frames is excluded'},
        'setup_s': {'render': round(render_s, 2), 'upload': round(upload_s, 3)},
        'accuracy_run0': ate, 'gpu_launches': int(timed * full['kernels_per_step']), 'clocks': clocks,
    }


# ---- one workload through the device-resident and the end-to-end leg --------------------------------------------------------
class Bench:
    """Shared state of a repo-arm run: torch handles, rank info, clock sampling windows."""

    def __init__(self, torch, dist, rank, world, local):
        self.torch, self.dist, self.rank, self.world, self.local = torch, dist, rank, world, local
        self.windows = []

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def _reduce(self, x, op):
        a = np.asarray(x, dtype=np.float64)
        if self.world == 1:
            return float(a) if a.ndim == 0 else a
        t = self.torch.tensor(a.reshape(-1), device='cuda', dtype=self.torch.float64)
        self.dist.all_reduce(t, op=getattr(self.dist.ReduceOp, op))
        r = t.cpu().numpy().reshape(a.shape)
        return float(r) if a.ndim == 0 else r

    def max_over_ranks(self, x):
        return self._reduce(x, 'MAX')

    def sum_over_ranks(self, x):
        return self._reduce(x, 'SUM')


def measure(B, wl, W, K, est_ms, n_base, n_prof=12, pageable=True):
    """Device-resident leg + end-to-end leg(s) of one workload.  Returns a dict of raw results."""
    torch = B.torch
    from image_processing import ImageProcessor, _native
    from synth_euroc import SlidingTextureStream, img_msg, stereo_msg
    cfg, skw, wname = workload(wl)
    skw = dict(skw, seed=skw['seed'] + B.rank)            # every rank owns its own stream
    R = repeats_for(K, est_ms)
    base = SlidingTextureStream(n_frames=n_base, **skw)
    frames = [base.frame(k) for k in range(n_base)]       # rendered once, decoded in RAM
    base.frames = lambda: iter(frames)
    T = W + 1 + R * K + n_prof
    seq = ForthAndBack(base, frames, T)
    Rs = rotations_for(cfg, seq)
    width, height = base.w, base.h
    img_bytes = width * height

    # ---- device-resident leg ("value") --------------------------------------------------------------------
    ctx = _native.Context(cfg, width, height, num_streams=1, device=B.local, use_graph=True)
    bb = ctx.block_bytes
    host_blocks = torch.empty((n_base, bb), dtype=torch.uint8).pin_memory()
    hb = host_blocks.numpy()
    Rs_base = rotations_for(cfg, base)                    # rotation sections of the base blocks (the C4 legs read them)
    for k, f in enumerate(frames):
        hb[k, :img_bytes] = f.cam0_image.reshape(-1)
        hb[k, img_bytes:2 * img_bytes] = f.cam1_image.reshape(-1)
        ctx.fill_rotations(hb[k], Rs_base[k][0], Rs_base[k][1])
    dev_base = host_blocks.cuda(non_blocking=False)
    # the timed sequence: T blocks, images gathered from the base frames, every step its own rotation section
    rot_off, rot_len = ctx.rot_offset, ctx.rot_stride
    scratch = np.zeros(bb, np.uint8)
    rot_host = np.empty((T, rot_len), np.uint8)
    for k in range(T):
        ctx.fill_rotations(scratch, Rs[k][0], Rs[k][1])
        rot_host[k] = scratch[rot_off:rot_off + rot_len]
    dev_seq = dev_base[torch.from_numpy(seq.index).cuda()]
    dev_seq[:, rot_off:rot_off + rot_len] = torch.from_numpy(rot_host).cuda()
    torch.cuda.synchronize()
    ptr = dev_seq.data_ptr()
    ext = torch.cuda.ExternalStream(ctx.cuda_stream(), device=B.local)
    ctx.process_device(ptr)                               # frame 0 (first-frame chain, launched without a graph)
    frame0_dev_ms = ctx.last_frame_ms()
    for k in range(1, W + 1):
        ctx.process_device(ptr + k * bb)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(R + 1)]
    B.barrier()
    t_a = time.perf_counter()
    ev[0].record(ext)
    for r in range(R):
        for k in range(W + 1 + r * K, W + 1 + (r + 1) * K):
            ctx.enqueue_device(ptr + k * bb)
        ev[r + 1].record(ext)
    ctx.sync()
    B.barrier()
    B.windows.append((t_a, time.perf_counter()))
    region_ms = np.atleast_1d(B.max_over_ranks(np.array([ev[r].elapsed_time(ev[r + 1]) for r in range(R)])))
    hdr, ids, _ = ctx.result(0)
    last_n_dev = int(hdr['n_features'])
    kernels_per_frame = ctx.kernels_per_frame()
    # per-stage device times of the steady-state chain, serialised (explains `value`; not a bench number)
    stage_ms = {}
    for k in range(W + 1 + R * K, T):
        st = ctx.profile_frame_device(ptr + k * bb)
        for name, v in st.items():
            stage_ms.setdefault(name, []).append(v)
    stage_ms = {k_: float(np.median(v)) for k_, v in stage_ms.items()}
    ctx.close()

    # ---- end-to-end leg through the public API ----------------------------------------------------------------
    # With the frames' numpy arrays living in page-locked memory (the contract's "inputs in pinned host memory": libavb
    # then DMAs straight from them) and, optionally, with ordinary pageable arrays (staged through the library's own
    # pinned block, copy pipelined with the H2D).  The headline e2e is the pinned-input one.
    n_e2e = W + 1 + R * K

    def e2e_leg(pinned_inputs):
        ip = ImageProcessor(cfg, device=B.local, use_graph=True)
        counts, evs, nst = [], [], 0
        for kind, msg in seq.events():
            if kind == 'stereo':
                if nst >= n_e2e:
                    break
                if pinned_inputs:
                    b = int(seq.index[nst])
                    i0 = hb[b, :img_bytes].reshape(height, width)
                    i1 = hb[b, img_bytes:2 * img_bytes].reshape(height, width)
                    msg = stereo_msg(msg.timestamp, i0, i1, img_msg(msg.timestamp, i0), img_msg(msg.timestamp, i1))
                nst += 1
            evs.append((kind, msg))
        state = {'idx': 0, 'frames': 0}

        def pump(until_frames):
            while state['idx'] < len(evs) and state['frames'] < until_frames:
                kind, msg = evs[state['idx']]
                state['idx'] += 1
                if kind == 'imu':
                    ip.imu_callback(msg)
                else:
                    fm = ip.stereo_callback(msg)
                    counts.append(len(fm.features))
                    state['frames'] += 1

        t0 = time.perf_counter()
        pump(1)
        frame0_ms = 1e3 * (time.perf_counter() - t0)
        pump(W + 1)
        B.barrier()
        t_a = time.perf_counter()
        spans = []
        for r in range(R):
            t0 = time.perf_counter()
            pump(W + 1 + (r + 1) * K)
            torch.cuda.synchronize()
            spans.append(time.perf_counter() - t0)
        B.barrier()
        B.windows.append((t_a, time.perf_counter()))
        cap = ip.context.capacity
        ip.context.close()
        return np.atleast_1d(B.max_over_ranks(np.array(spans))), counts, cap, frame0_ms

    e2e_spans, feats_per_frame, cap, frame0_e2e_ms = e2e_leg(True)
    assert feats_per_frame[W + R * K] == last_n_dev, 'device-resident and end-to-end legs disagree on the last frame'
    res = dict(cfg=cfg, wname=wname, width=width, height=height, bb=bb, R=R, W=W, K=K, region_ms=region_ms,
               e2e_spans=e2e_spans, feats_per_frame=feats_per_frame, cap=cap, kernels_per_frame=kernels_per_frame,
               stage_ms=stage_ms, frame0_dev_ms=frame0_dev_ms, frame0_e2e_ms=frame0_e2e_ms, base=base, frames=frames,
               dev_base=dev_base, hb=hb, n_base=n_base,
               d2h_bytes=(int(_native.C.sizeof(_native.AvbFrameHeader)) + cap * 64 + 255) & ~255)
    if pageable:
        pg_spans, feats_pg, _, _ = e2e_leg(False)
        assert feats_pg == feats_per_frame
        res['pageable_spans'] = pg_spans
    del dev_seq
    return res


def c4_leg(B, m, S, KM, with_sweep):
    """BASELINE config C4 on this GPU: S time-offset runs of the bench sequence (run s starts 2*s frames in) lock-stepped in
    one context, inputs resident in HBM.  The last timed frame of the first and the last run is compared with a
    single-stream context fed the same frames (ids and published coordinates identical)."""
    torch = B.torch
    from image_processing import _native
    cfg, width, height = m['cfg'], m['width'], m['height']
    img_bytes = width * height
    WM = 4
    dev_base, bb1 = m['dev_base'], m['bb']
    mctx = _native.Context(cfg, width, height, num_streams=S, device=B.local, use_graph=True)
    mbb = mctx.block_bytes
    rot_off = mctx.rot_offset
    nblk = WM + 1 + KM + 2                            # + 2 frames for the serialised stage timing
    assert 2 * (S - 1) + nblk <= m['n_base']
    mblocks = torch.zeros((nblk, mbb), dtype=torch.uint8, device='cuda')
    imgs = dev_base[:, :2 * img_bytes]
    one = _native.Context(cfg, width, height, num_streams=1, device=B.local, use_graph=True)
    rs_ = one.rot_stride                              # H | cam0_R_p_c | cam1_R_p_c per stream
    Hs = dev_base[:, one.rot_offset:one.rot_offset + rs_]
    for k in range(nblk):
        src = torch.arange(S, device='cuda') * 2 + k      # run s starts 2*s frames into the sequence
        mblocks[k, :S * 2 * img_bytes] = imgs[src].reshape(-1)
        mblocks[k, rot_off:rot_off + S * rs_] = Hs[src].reshape(-1)
    torch.cuda.synchronize()
    mptr = mblocks.data_ptr()
    mext = torch.cuda.ExternalStream(mctx.cuda_stream(), device=B.local)
    for k in range(WM + 1):
        mctx.process_device(mptr + k * mbb)
    m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    B.barrier()
    t_a = time.perf_counter()
    m0.record(mext)
    for k in range(WM + 1, WM + 1 + KM):
        mctx.enqueue_device(mptr + k * mbb)
    m1.record(mext)
    mctx.sync()
    B.barrier()
    B.windows.append((t_a, time.perf_counter()))
    m_ms = B.max_over_ranks(m0.elapsed_time(m1))
    nf = sum(int(mctx.result(s)[0]['n_features']) for s in range(S))
    # parity of the timed shape: the first and the last run against a single-stream context on the same frames
    checked = []
    for s in sorted({0, S - 1}):
        _, ids_m, meas_m = mctx.result(s)
        ids_m, meas_m = ids_m.copy(), meas_m.copy()
        one.reset()
        for k in range(WM + 1 + KM):
            one.process_device(dev_base.data_ptr() + (2 * s + k) * bb1)
        _, ids_1, meas_1 = one.result(0)
        assert np.array_equal(ids_m, ids_1) and np.array_equal(meas_m, meas_1), \
            f'C4 leg: run {s} of the {S}-run context differs from a single-stream context on its last timed frame'
        checked.append(s)
    one.close()
    mctx.profile_frame_device(mptr + (nblk - 2) * mbb)
    st = mctx.profile_frame_device(mptr + (nblk - 1) * mbb)
    kern = mctx.kernels_per_frame()
    mctx.close()
    bpf = algorithmic_bytes_per_frame(width, height, cfg.pyramid_levels, nf / S)
    world = B.world
    out = {'streams_total': S * world, 'streams_per_gpu': S, 'steps': KM, 'value': world * S * KM / (m_ms * 1e-3),
           'unit': UNIT, 'ms_per_step': m_ms / KM, 'features_per_frame': nf / S, 'kernels_per_step': kern,
           'hbm_gbs': bpf * S * KM / (m_ms * 1e-3) / 1e9,
           'stage_ms': {k_: round(v, 4) for k_, v in st.items()},
           'stage_hbm_gbs': {k_: round(b / (st[k_] * 1e-3) / 1e9, 1)
                             for k_, b in stage_bytes(width, height, cfg.pyramid_levels, S).items() if st.get(k_, 0) > 0},
           'parity_checked': f'last timed frame of runs {checked}: ids and published coordinates identical to a single-stream '
                             f'context on the same frames (tests/test_gpu_many_streams.py pins every run and frame to the port)',
           'note': 'S time-offset runs of the sequence (run s starts 2*s frames in), lock-stepped in one context: every '
                   'kernel launch covers all S runs; inputs resident in HBM (device-resident figure, like `value`)'}
    del mblocks
    if with_sweep:
        # the same sweep through the public driver: the sequence cached once in an HBM frame store, every run gathers its
        # frames on the device, results come back as host arrays (sweep.run_sweep; wall clock, IMU windows included)
        from sweep import CachedSequence, run_sweep

        class _Seq:
            def __init__(self, fr, st):
                self.fr, self.st, self.n = fr, st, len(fr)

            def frames(self):
                return iter(self.fr)

            def imu(self):
                return self.st.imu()
        base = m['base']
        cached = CachedSequence(_Seq(m['frames'], base), device=B.local, name='bench sequence')
        sw = run_sweep(cfg, [cached], [(2 * s_ + 0.5) / base.rate for s_ in range(S)], device=B.local,
                       n_steps=min(200, len(m['frames']) - 2 * (S - 1)), warmup_steps=WM + 1)
        B.barrier()
        sw_wall = B.max_over_ranks(sw['wall_s'])
        cached.close()
        out['e2e_from_store'] = {'value': world * S * sw['timed_steps'] / sw_wall, 'unit': UNIT,
                                 'ms_per_step': 1e3 * sw_wall / sw['timed_steps'], 'steps': sw['timed_steps'],
                                 'features_per_frame': float(sw['features'][:, WM + 1:].mean()),
                                 'note': 'sweep.run_sweep: frames gathered from the HBM frame store (sequence uploaded once), '
                                         'per-run IMU windows on the host, ids + measurements of every run copied to host arrays'}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--warmup', type=int, default=10)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--workload', default='c2', choices=['c2', 'c3', 'c5'])
    ap.add_argument('--c5-sequences', type=int, default=8)
    ap.add_argument('--c5-offsets', type=int, default=16)
    ap.add_argument('--c5-steps', type=int, default=60)
    ap.add_argument('--c5-workers', type=int, default=0, help='estimator processes per GPU (0 = host cores of the rank)')
    ap.add_argument('--streams', type=int, default=64,
                    help='total runs of the multi-stream leg (config C4), sharded over the GPUs (0 = skip)')
    ap.add_argument('--ms-steps', type=int, default=20)
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg')
    ap.add_argument('--no-sublegs', action='store_true', help='skip the compact c3 / c5 legs of the default line')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if args.impl == 'reference':
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device: libavb has no CPU fallback')
    torch.cuda.set_device(local)
    # one process per GPU: stay on the host cores next to this GPU (pinned staging / mapped result block are first-touch)
    from multi_stream import bind_to_gpu_numa
    full_affinity = os.sched_getaffinity(0)
    pr = torch.cuda.get_device_properties(local)
    bound = bind_to_gpu_numa(f'{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0', local,
                             int(os.environ.get('LOCAL_WORLD_SIZE', world)))
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    B = Bench(torch, dist, rank, world, local)

    if args.workload == 'c5':
        line = run_c5(B, args.c5_sequences, args.c5_offsets, args.c5_steps, args.c5_workers, bound)
        if rank == 0:
            print(json.dumps(line), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return

    W, K = args.warmup, args.steps
    sampler = ClockSampler(local) if rank == 0 else None
    # multi-stream legs = BASELINE config C4: `--streams` (64) time-offset runs in total, sharded over the GPUs (strong), and
    # `--streams` runs on every GPU (weak)
    c4 = args.streams > 0 and args.workload == 'c2'
    S_strong = max(1, args.streams // world) if c4 else 0
    S_weak = args.streams if c4 else 0
    n_extra = (2 * (max(S_strong, S_weak) - 1) + max(4 + 1 + args.ms_steps + 2, 105)) if c4 else 0
    est = {'c2': 0.12, 'c3': 0.35}[args.workload]
    m = measure(B, args.workload, W, K, est, n_base=max(200 if args.workload == 'c2' else 48, n_extra))
    cfg, wname, width, height = m['cfg'], m['wname'], m['width'], m['height']
    R = m['R']
    timed_feats = float(np.sum(m['feats_per_frame'][W + 1:W + 1 + R * K])) / R      # per region of K frames
    total_feats = B.sum_over_ranks(timed_feats)

    multi = weak = None
    if c4:
        multi = c4_leg(B, m, S_strong, args.ms_steps, with_sweep=True)
        multi['config'] = f'C4 strong: {S_strong * world} independent time-offset runs sharded over {world} GPU(s)'
        if world > 1:
            weak = c4_leg(B, m, S_weak, args.ms_steps, with_sweep=False)
            weak['config'] = f'C4 weak: {S_weak} time-offset runs on each of {world} GPU(s)'
        else:
            weak = {'same_as': 'multi_stream (one GPU: 64 runs in total = 64 runs per GPU)', 'value': multi['value'],
                    'unit': UNIT, 'streams_per_gpu': S_weak, 'streams_total': S_weak}

    # ---- compact legs of the other BASELINE configurations (one GPU only: the driver's BENCH record carries them) ----------
    c3 = c5 = None
    if world == 1 and args.workload == 'c2' and not args.no_sublegs:
        m3 = measure(B, 'c3', 3, 30, 0.35, n_base=40, n_prof=4, pageable=False)
        r3 = float(np.median(m3['region_ms']))
        e3 = float(np.median(m3['e2e_spans']))
        c3 = {'workload': m3['wname'], 'steps': 30, 'warmup': 3, 'repeats': m3['R'],
              'value': 30 / (r3 * 1e-3), 'ms_per_step': r3 / 30, 'unit': UNIT,
              'e2e': {'value': 30 / e3, 'ms_per_step': 1e3 * e3 / 30, 'h2d_bytes_per_step': int(m3['bb']),
                      'd2h_bytes_per_step': int(m3['d2h_bytes'])},
              'features_per_frame': float(np.mean(m3['feats_per_frame'][4:])), 'kernels_per_frame': m3['kernels_per_frame'],
              'frame0_ms': {'device': m3['frame0_dev_ms'], 'e2e': m3['frame0_e2e_ms']},
              'stage_ms_serialised': {k_: round(v, 4) for k_, v in m3['stage_ms'].items()}}
        del m3
        torch.cuda.empty_cache()
        c5line = run_c5(B, 2, 8, 40, 0, bound)
        c5 = {k_: c5line[k_] for k_ in ('value', 'unit', 'steps', 'ms_per_step', 'tracked_features_per_s', 'poses_published',
                                        'front_end_only', 'accuracy_run0', 'setup_s')}
        c5['workload'] = c5line['config']['workload'] + ' (compact: the full C5 is `--workload c5`)'
        c5['estimator'] = {'ms_per_frame': c5line['estimator']['ms_per_frame'],
                           'workers': c5line['config']['estimator_workers_per_gpu'], 'errors': c5line['estimator']['errors']}

    # ---- CPU baseline (rank 0, N=1 only): the unmodified reference on a bounded sample -----------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        os.sched_setaffinity(0, full_affinity)            # the CPU path may use every host core
        cores = os.cpu_count() or 1
        n_cpu = int(min(W + 1 + K, 300))
        port_t, port_f, port_threads, cvv = port_front_end(cfg, ForthAndBack(m['base'], m['frames'], n_cpu), n_cpu, budget_s=12.0)
        pw = min(W + 1, max(len(port_t) - 1, 1))
        port = {'value': (len(port_t) - pw) / float(np.sum(port_t[pw:])), 'kind': 'port', 'cores': port_threads,
                'note': 'oracle/pipeline_port.py (vectorised restatement, cv2 at the same call sites): kept beside the reference figure'}
        v = cpu_variants(args.workload, n_cpu, W, cores) if ref_available() else None
        if v is not None and c3 is not None:
            v3 = run_ref('c3', 3 + 1 + 8, 3, threads=0)
            if v3 is not None:
                c3['cpu_reference'] = {'value': v3['fps'], 'kind': 'reference', 'cores': v3['cv2_threads'],
                                       'frames_timed': v3['frames_timed'],
                                       'note': 'unmodified reference, one stream, cv2 default threads; no RANSAC stage exists there'}
                c3['e2e_vs_cpu_reference'] = c3['e2e']['value'] / v3['fps']
        if v is not None:
            one = v['one_stream_default_threads']
            cpu = {'value': one['value'], 'unit': UNIT, 'cores': one['cv2_threads'], 'kind': 'reference',
                   'sample': f'{one["frames_timed"]} consecutive frames of the same stream (after frame 0 + {W} warm-up frames) through the '
                             f'UNMODIFIED reference front end (oracle/_ref: image_processing/pipeline.py:46-150), one stream, cv2 {v["cv2"]} '
                             f'with its default {one["cv2_threads"]} threads; host cores {cores}',
                   'ms_per_frame_median': one['ms_median'], 'frame0_ms': one['frame0_ms'],
                   'tracked_features_per_s': one['tracked_features_per_s'], 'variants': v, 'port': port}
        else:
            cpu = dict(port, unit=UNIT, sample=f'{len(port_t) - pw} frames; oracle/_ref missing: the PORT stands in (tools/make_oracle_ref.sh)')

    clocks = sampler.summary(B.windows) if sampler is not None else None

    if rank == 0:
        peaks_path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
        else:
            peak, peak_src = 6650.0, 'fallback (B200_PROFILING.md)'
        dev_ms = float(np.median(m['region_ms']))
        e2e_s = float(np.median(m['e2e_spans']))
        pg_s = float(np.median(m['pageable_spans']))
        fps = world * K / (dev_ms * 1e-3)
        nfeat = timed_feats / K
        bpf = algorithmic_bytes_per_frame(width, height, cfg.pyramid_levels, nfeat)
        chain_ms = dev_ms / K
        achieved = bpf / (chain_ms * 1e-3) / 1e9
        bb = m['bb']
        traffic, traffic_src = None, None                 # measured DRAM bytes of one frame chain (ncu --set full capture)
        tpath = os.path.join(ROOT, 'profiles', f'roofline_traffic_{args.workload}.json')
        if os.path.exists(tpath):
            tj = json.load(open(tpath))
            if 'warm' in tj:            # the chain as it runs (caches left alone); the cold-cache replay figure beside it
                traffic = float(tj['warm']['bytes_per_frame']) + float(bb)
                traffic_src = (f"profiles/{os.path.basename(tpath)}: {float(tj['warm']['bytes_per_frame']):.0f} B by the kernels "
                               f"(ncu --cache-control none, DRAM counters of the running chain: the kernels find the just-copied "
                               f"input block and both pyramids in L2) + {int(bb)} B read by the copy engine that places the "
                               f"frame's input block; cold-cache replay of the same kernels (--set full, caches flushed before "
                               f"every kernel): {float(tj['bytes_per_frame']):.0f} B")
            else:
                traffic, traffic_src = float(tj['bytes_per_frame']), f"profiles/{os.path.basename(tpath)} ({tj.get('note', '')})"
        stage_ms = m['stage_ms']
        kernel_stages = {k_: v for k_, v in stage_ms.items() if k_ not in ('input_copy', 'result_copy')}
        dominant = max(kernel_stages, key=kernel_stages.get)
        kpf = m['kernels_per_frame']
        line = {
            'metric': METRIC, 'value': fps, 'unit': UNIT, 'n_gpus': world, 'steps': K, 'warmup': W, 'repeats': R,
            'ms_per_step': chain_ms, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': DTYPE, 'data': 'synthetic', 'config': line_config(wname, world),
            'value_is': 'device-resident (inputs in HBM, no host result awaited per frame): explains e2e; the SURVEY 8(d) metric '
                        '(feature_msg materialised on the host) is `e2e`',
            'region_ms': {'median': dev_ms, 'min': float(np.min(m['region_ms'])), 'max': float(np.max(m['region_ms']))},
            'features_per_frame': nfeat, 'host_cpus_bound': sorted(bound) if bound else None,
            'frame_graph': 'one CUDA-graph launch per frame',
            'tracked_features_per_s': total_feats / (dev_ms * 1e-3),
            'frame0_ms': {'device': m['frame0_dev_ms'], 'e2e': m['frame0_e2e_ms'],
                          'note': 'first frame of a stream: FAST + stereo match of EVERY keypoint (feature_initializer.py:45-85), '
                                  'launched without a graph'},
            'e2e': {'value': world * K / e2e_s, 'unit': UNIT, 'h2d_bytes_per_step': int(bb),
                    'd2h_bytes_per_step': int(m['d2h_bytes']), 'ms_per_step': 1e3 * e2e_s / K, 'repeats': R,
                    'region_s': {'median': e2e_s, 'min': float(np.min(m['e2e_spans'])), 'max': float(np.max(m['e2e_spans']))},
                    'tracked_features_per_s': total_feats / e2e_s,
                    'api': 'ImageProcessor.stereo_callback(stereo_msg) -> feature_msg, host numpy images (page-locked) in, '
                           'FeatureMeasurement list out',
                    'pageable_inputs': {'value': world * K / pg_s, 'ms_per_step': 1e3 * pg_s / K,
                                        'note': 'same call with ordinary pageable numpy arrays: staged through the '
                                                'library pinned block, host copy pipelined with the H2D'}},
            'gpu_launches': int(K * kpf),
            'roofline': {'bound': 'hbm', 'kernel': f'frame chain ({kpf} kernels, one CUDA graph); '
                                                   f'dominant stage by time: {dominant}',
                         'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                         'traffic': traffic, 'traffic_source': traffic_src, 'peak_source': peak_src,
                         'algorithmic_bytes_per_launch': bpf,
                         'stage_ms_serialised': {k_: round(v, 4) for k_, v in stage_ms.items()},
                         'note': 'single stream = latency-bound dependent chain; see multi_stream for the '
                                 'batched figure and DESIGN.md section 5'},
            'multi_stream': multi, 'c4_weak': weak, 'c3': c3, 'c5': c5,
            'cpu_baseline': cpu,
            'clocks': clocks,
        }
        if multi is not None:
            multi['roofline_frac'] = multi['hbm_gbs'] / peak
            # the image-scan kernels are the HBM-shaped part of the path: their own algorithmic bytes / stage time at the
            # many-stream launch shape, against the same measured peak (the LK kernels work out of L1/L2: DESIGN.md section 5)
            line['roofline']['hbm_shaped_stages'] = {
                k_: {'achieved': v, 'unit': 'GB/s', 'frac': round(v / peak, 4), 'streams_per_launch': multi['streams_per_gpu']}
                for k_, v in multi['stage_hbm_gbs'].items()}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
