"""Host MSCKF consumers in worker processes (BASELINE config C5: "full front end +
host MSCKF").

The reference runs the estimator on a thread of the same Python process as the front end (modules/vio.py:17-19,
46-53): one stream per process, the GIL shared.  With the front end at tens of thousands of frames per second per GPU
the estimator (milliseconds per frame, pure host work) is what bounds a sweep, so every GPU process keeps P estimator
processes, stream s of the context served by worker s mod P.  Messages are plain arrays (IMU rows, ids, normalized
stereo measurements); a worker rebuilds the `feature_msg` the filter expects.  Nothing here touches CUDA, and the module
imports neither torch nor libavb: workers start in well under a second with the `spawn` method.
"""
from __future__ import annotations

import multiprocessing as mp
import time
from collections import namedtuple

import numpy as np

imu_msg = namedtuple('imu_msg', ['timestamp', 'angular_velocity', 'linear_acceleration'])
feature_msg = namedtuple('feature_msg', ['timestamp', 'features'])
Meas = namedtuple('FeatureMeasurement', ['id', 'u0', 'v0', 'u1', 'v1'])


def feed(est, imu_rows, ts, ids, meas):
    """One frame into an estimator: IMU rows (t, gyro xyz, acc xyz) in order, then the feature message."""
    for row in imu_rows:
        est.imu_callback(imu_msg(float(row[0]), row[1:4].copy(), row[4:7].copy()))
    feats = [Meas(int(i), *r) for i, r in zip(ids.tolist(), meas.tolist())]
    return est.feature_callback(feature_msg(float(ts), feats))


def _worker(inbox, conn, config, streams):
    from msckf import MSCKF
    ests = {s: MSCKF(config, outfile=False) for s in streams}
    traj = {s: [] for s in streams}
    busy = 0.0
    frames = 0
    while True:
        batch = inbox.get()
        if batch is None:
            break
        t0 = time.perf_counter()
        for s, imu_rows, ts, ids, meas in batch:
            r = feed(ests[s], imu_rows, ts, ids, meas)
            frames += 1
            if r is not None:
                st = ests[s].imu_state
                traj[s].append([r.timestamp, *r.pose.t, *st.orientation])
        busy += time.perf_counter() - t0
    conn.send({'traj': {s: np.array(v, dtype=np.float64).reshape(-1, 8) for s, v in traj.items()},
               'busy_s': busy, 'frames': frames})
    conn.close()


class EstimatorPool:
    """`n_streams` MSCKF instances spread over `n_workers` processes.  `push_step` queues one frame of every stream
    and returns at once: every worker has its own inbox `depth` steps deep (filter time per frame varies 3x around its
    median, so workers must be allowed to drift apart; a full inbox blocks the producer, which is the back-pressure).
    `finish` collects per stream the published trajectory rows (t, x, y, z, qx, qy, qz, qw: the reference's output-file
    columns, msckf.py:152-160)."""

    def __init__(self, config, n_streams, n_workers, method='spawn', depth=64):
        self.S, self.P = int(n_streams), max(1, min(int(n_workers), int(n_streams)))
        ctx = mp.get_context(method)
        self.conns, self.procs, self.inboxes = [], [], []
        for w in range(self.P):
            parent, child = ctx.Pipe(duplex=False)
            inbox = ctx.Queue(maxsize=int(depth))
            p = ctx.Process(target=_worker, args=(inbox, child, config, list(range(w, self.S, self.P))), daemon=True)
            p.start()
            child.close()
            self.conns.append(parent)
            self.inboxes.append(inbox)
            self.procs.append(p)

    def push_step(self, items):
        """items[s] = (imu_rows float64[m, 7], timestamp, ids int64[n], meas float64[n, 4])."""
        if len(items) != self.S:
            raise ValueError(f'expected {self.S} items')
        for w, inbox in enumerate(self.inboxes):
            inbox.put([(s, *items[s]) for s in range(w, self.S, self.P)])

    def finish(self):
        for inbox in self.inboxes:
            inbox.put(None)
        traj, busy, frames = {}, [], 0
        for conn in self.conns:
            r = conn.recv()
            traj.update(r['traj'])
            busy.append(r['busy_s'])
            frames += r['frames']
        for p in self.procs:
            p.join(timeout=10)
        return [traj[s] for s in range(self.S)], {'worker_busy_s': busy, 'frames': frames}

    def close(self):
        for p in self.procs:
            if p.is_alive():
                p.terminate()
