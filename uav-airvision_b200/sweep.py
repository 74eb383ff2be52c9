"""Sequence x time-offset sweep on one GPU (SURVEY.md section 8 rows e, f1, f2; BASELINE configs C4/C5).

The reference's sweep is a batch file: for every sequence and every offset, `python main.py --path <seq> --offset <s>`
(run.bat:4-12), each run decoding its own frames from its own start time (main.py:12-13, streaming/dataset.py:206-214)
and pacing them by wall clock (streaming/publisher.py:32-53).  Here

  * a sequence is decoded ONCE into an HBM-resident `FrameStore` (`CachedSequence`); all its offset runs read it there;
  * the runs of a GPU advance in lock-step through one multi-stream context: per step one gather kernel + one frame
    chain for all of them, no host image copies, no wall-clock pacing (deterministic);
  * the estimators (host MSCKF, msckf.py) run in worker processes fed with arrays (`EstimatorPool`).

Across GPUs the runs are sharded by sequence (one process per GPU, `multi_stream.shard_streams`), no collective.
"""
from __future__ import annotations

import time
from collections import namedtuple

import numpy as np

from image_processing import _native
from multi_stream import MultiStreamFrontEnd

imu_msg = namedtuple('imu_msg', ['timestamp', 'angular_velocity', 'linear_acceleration'])
frame_ref = namedtuple('frame_ref', ['timestamp'])         # what IMUProcessor reads of an image message
SweepRun = namedtuple('SweepRun', ['sequence', 'offset', 'first_frame', 'first_imu'])


def offset_start(frame_times, imu_times, offset_s):
    """(first frame index, first IMU index) of the run `--offset offset_s`: starttime = first IMU stamp + offset (the
    reference's max(imu.start_time(), stereo.start_time()) sees cam0.starttime = -inf, streaming/dataset.py:184-185,
    203), and every reader drops what is older than it (dataset.py:206-214, 72-75, 119-123)."""
    start = (imu_times[0] if len(imu_times) else frame_times[0]) + float(offset_s)
    return (int(np.searchsorted(frame_times, start, side='left')), int(np.searchsorted(imu_times, start, side='left')))


class CachedSequence:
    """One sequence, decoded once: stereo frames in HBM, IMU samples as rows (t, gyro, acc) + message objects.
    `source` is anything with .frames() (stereo_msg) and .imu() (imu_msg): an EuRoCDataset (euroc.py) or a synthetic
    stream (synth_euroc.py).  `max_frames` bounds the upload."""

    def __init__(self, source, device=0, max_frames=None, name=None):
        self.name = name or getattr(source, 'path', type(source).__name__)
        if hasattr(source, 'frames'):
            frames = source.frames()
        else:                                           # EuRoCDataset: decode ahead on several threads
            frames = source.stereo.prefetch() if hasattr(source.stereo, 'prefetch') else iter(source.stereo)
        first = next(frames)
        h, w = first.cam0_image.shape
        n = max_frames if max_frames is not None else getattr(source, 'n', None)
        if n is None:
            n = len(source.stereo)
        self.store = _native.FrameStore(w, h, n, device=device)
        self.width, self.height = w, h
        ts = []
        k = 0
        f = first
        while f is not None and k < n:
            self.store.upload(k, f.cam0_image, f.cam1_image, f.timestamp)
            ts.append(float(f.timestamp))
            k += 1
            f = next(frames, None)
        self.n_frames = k
        self.timestamps = np.array(ts)
        self.frame_refs = [frame_ref(t) for t in ts]
        imu = source.imu() if callable(getattr(source, 'imu', None)) else iter(source.imu)
        t_end = ts[-1] if ts else -np.inf
        rows = []
        for m in imu:
            if m.timestamp > t_end:
                break
            rows.append([m.timestamp, *m.angular_velocity, *m.linear_acceleration])
        self.imu_rows = np.array(rows, dtype=np.float64).reshape(-1, 7)
        self.imu_msgs = [imu_msg(r[0], r[1:4].copy(), r[4:7].copy()) for r in self.imu_rows]
        gt = getattr(source, 'groundtruth', None)
        self.groundtruth = None
        if gt is not None:
            g = list(gt() if callable(gt) else gt)
            g = [x for x in g if x.timestamp <= t_end + 1.0]
            if g:
                self.groundtruth = (np.array([x.timestamp for x in g]), np.array([x.p for x in g]))

    def run(self, seq_index, offset_s):
        return SweepRun(seq_index, float(offset_s), *offset_start(self.timestamps, self.imu_rows[:, 0], offset_s))

    def close(self):
        self.store.close()


def run_sweep(config, sequences, offsets_s, device=0, n_steps=None, estimator_workers=0, estimator_config=None,
              warmup_steps=0):
    """All (sequence, offset) runs of this GPU in lock-step.  Returns a dict: per-run feature counts and (with
    estimators) trajectories, plus wall-clock timing of the steady-state steps (`warmup_steps` first steps, which hold
    the first-frame chain, are excluded from `frames_per_s`)."""
    runs = [seq.run(q, off) for q, seq in enumerate(sequences) for off in offsets_s]
    S = len(runs)
    if S == 0:
        raise ValueError('empty sweep')
    w, h = sequences[0].width, sequences[0].height
    if any((q.width, q.height) != (w, h) for q in sequences):
        raise ValueError('the sequences of one context share one resolution')
    avail = min(sequences[r.sequence].n_frames - r.first_frame for r in runs)
    steps = avail if n_steps is None else min(int(n_steps), avail)
    if steps <= 0:
        raise ValueError('an offset lies beyond the cached part of its sequence')
    pool = None
    if estimator_workers > 0:
        from estimator_pool import EstimatorPool
        pool = EstimatorPool(estimator_config or config, S, estimator_workers)
    fe = MultiStreamFrontEnd(config, w, h, S, device=device)
    launches = fe.ctx.kernels_per_frame() + (2 * S + 255) // 256       # frame chain + gather launches
    imu_pos = [r.first_imu for r in runs]
    n_feat = np.zeros((S, steps), dtype=np.int32)
    t_mark = None
    fe_s = 0.0
    def host_side(k):
        """IMU messages up to frame k of every run -> the front end's buffers; the step's frame refs and addresses."""
        refs, spans = [], []
        a = np.empty((S, 2), dtype=np.uint64)
        for s, r in enumerate(runs):
            seq = sequences[r.sequence]
            idx = r.first_frame + k
            j0 = imu_pos[s]
            j1 = int(np.searchsorted(seq.imu_rows[:, 0], seq.timestamps[idx], side='right'))
            for m in seq.imu_msgs[j0:j1]:
                fe.imu_callback(s, m)
            imu_pos[s] = j1
            spans.append((j0, j1))
            refs.append(seq.frame_refs[idx])
            a[s] = seq.store.addr[idx]
        fe.prepare_step(refs)
        return a, spans

    try:
        # software pipeline: while the GPU runs step k, the host prepares step k+1 (IMU windows) and hands the results
        # of step k-1 to the estimators
        addrs, spans = host_side(0)
        fe.begin_step_from_store(addrs)
        for k in range(steps):
            if k == warmup_steps:
                t_mark = time.perf_counter()
            nxt = host_side(k + 1) if k + 1 < steps else None
            t0 = time.perf_counter()
            fe.wait_step()
            fe_s += time.perf_counter() - t0 if k >= warmup_steps else 0.0
            # step k+1 goes to the GPU BEFORE step k's results are taken apart: the result blocks of two consecutive
            # frames are both kept (avb_get_result_prev), so the GPU does not idle through the host's copy and loops
            if nxt is not None:
                fe.begin_step_from_store(nxt[0])
            out = fe.take_results(prev=nxt is not None)
            for s, (ts, ids, meas) in enumerate(out):
                n_feat[s, k] = len(ids)
            if pool is not None:
                pool.push_step([(sequences[r.sequence].imu_rows[j0:j1], ts, ids, meas)
                                for r, (j0, j1), (ts, ids, meas) in zip(runs, spans, out)])
            if nxt is not None:
                spans = nxt[1]
        t_fe_done = time.perf_counter()
        traj, pstats = (None, None)
        if pool is not None:
            traj, pstats = pool.finish()
        t_end = time.perf_counter()
    finally:
        fe.close()
        if pool is not None:
            pool.close()
    timed = steps - warmup_steps
    wall = t_end - (t_mark if t_mark is not None else t_end)
    return {'runs': runs, 'steps': steps, 'timed_steps': timed, 'streams': S, 'features': n_feat,
            'trajectories': traj, 'estimator': pstats, 'kernels_per_step': launches,
            'wall_s': wall, 'front_end_wait_s': fe_s, 'front_end_done_s': t_fe_done - (t_mark or t_fe_done),
            'frames_per_s': S * timed / wall if wall > 0 and timed > 0 else 0.0,
            'features_per_s': float(n_feat[:, warmup_steps:].sum()) / wall if wall > 0 and timed > 0 else 0.0}
