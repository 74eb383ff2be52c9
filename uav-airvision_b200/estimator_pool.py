"""Host MSCKF consumers in worker processes (BASELINE config C5: "full front end +
host MSCKF").

The reference runs the estimator on a thread of the same Python process as the front end (modules/vio.py:17-19,
46-53): one stream per process, the GIL shared.  With the front end at tens of thousands of frames per second per GPU
the estimator (milliseconds per frame, pure host work) is what bounds a sweep, so every GPU process keeps P estimator
processes, stream s of the context served by worker s mod P.

Transport: one shared-memory ring per worker (`depth` slots; a slot = one step of all the worker's streams: IMU rows,
timestamp, ids, normalized stereo measurements as plain arrays) guarded by two semaphores.  The producer copies arrays
into the slot and posts; nothing is pickled and no feeder thread competes for the producer's GIL (with pickling queues
the producer of a 128-run sweep spent 37 ms per step handing 1.6 MB to 23 feeder threads).  A worker rebuilds the
`feature_msg` the filter expects.  Nothing here touches CUDA (the host driver extension is imported for its list builder only, no context is created)
and torch is never imported: workers start in about a second with the `spawn` method.
"""
from __future__ import annotations

import multiprocessing as mp
import time
from collections import namedtuple
from multiprocessing import shared_memory

import numpy as np

imu_msg = namedtuple('imu_msg', ['timestamp', 'angular_velocity', 'linear_acceleration'])
feature_msg = namedtuple('feature_msg', ['timestamp', 'features'])
Meas = namedtuple('FeatureMeasurement', ['id', 'u0', 'v0', 'u1', 'v1'])

MAX_IMU = 64                # IMU rows per stream and slot; longer runs of IMU samples travel in extra IMU-only slots
STOP, IMU_ONLY, FRAME = -1, 0, 1


try:                                    # the C list builder of the front end's host driver (3 us for 300 measurements)
    from image_processing import FeatureMeasurement as _FM, _avbhost as _host
except Exception:                       # pragma: no cover - e.g. the filter used without the front end's extension built
    _host = None


def feed(est, imu_rows, ts, ids, meas):
    """One frame into an estimator: IMU rows (t, gyro xyz, acc xyz) in order, then the feature message."""
    for row in imu_rows:
        est.imu_callback(imu_msg(float(row[0]), row[1:4].copy(), row[4:7].copy()))
    if hasattr(est, 'feature_callback_arrays'):            # msckf.MSCKF of this package: the arrays as they are
        return est.feature_callback_arrays(float(ts), ids, meas)
    if _host is not None and ids.flags.c_contiguous and meas.flags.c_contiguous and ids.dtype == np.int64 and meas.dtype == np.float64:
        feats = _host.features_from_arrays(ids, meas, _FM)
    else:
        feats = [Meas(int(i), *r) for i, r in zip(ids.tolist(), meas.tolist())]
    return est.feature_callback(feature_msg(float(ts), feats))    # any estimator with the reference's interface


def state_row(est):
    """The columns the reference appends to its output file per published state (msckf.py:152-160):
    IMU-state timestamp, position, JPL orientation quaternion (x, y, z, w)."""
    st = est.imu_state
    return [st.timestamp, *st.position, *st.orientation]


class _Ring:
    """Views into one worker's shared-memory ring: [slot][stream] arrays."""

    def __init__(self, buf, depth, n, cap):
        self.depth, self.n, self.cap = depth, n, cap
        off = 0

        def take(shape, dtype):
            nonlocal off
            a = np.ndarray(shape, dtype=dtype, buffer=buf, offset=off)
            off += a.nbytes
            return a
        self.kind = take((depth,), np.int64)                    # STOP / IMU_ONLY / FRAME
        self.count = take((depth, n, 2), np.int64)              # IMU rows, features
        self.ts = take((depth, n), np.float64)
        self.imu = take((depth, n, MAX_IMU, 7), np.float64)
        self.ids = take((depth, n, cap), np.int64)
        self.meas = take((depth, n, cap, 4), np.float64)
        self.nbytes = off

    @staticmethod
    def size(depth, n, cap):
        return 8 * depth * (1 + 2 * n + n + n * MAX_IMU * 7 + n * cap + n * cap * 4)


def _worker(shm_name, depth, cap, filled, free, conn, config, streams):
    from msckf import MSCKF
    shm = shared_memory.SharedMemory(name=shm_name)
    try:
        ring = _Ring(shm.buf, depth, len(streams), cap)
        ests = [MSCKF(config, outfile=False, blas_threads=1) for _ in streams]     # this process runs nothing but filters
        traj = [[] for _ in streams]
        dead = {}                       # stream -> error text: a filter that raised is retired, the other runs go on
        busy, frames, slot = 0.0, 0, 0
        conn.send('ready')
        while True:
            filled.acquire()
            kind = int(ring.kind[slot])
            if kind == STOP:
                break
            t0 = time.perf_counter()
            for i, est in enumerate(ests):
                if streams[i] in dead:
                    continue
                n_imu, n_feat = (int(v) for v in ring.count[slot, i])
                try:
                    if kind == IMU_ONLY:
                        for row in ring.imu[slot, i, :n_imu]:
                            est.imu_callback(imu_msg(float(row[0]), row[1:4].copy(), row[4:7].copy()))
                        continue
                    r = feed(est, ring.imu[slot, i, :n_imu], ring.ts[slot, i], ring.ids[slot, i, :n_feat], ring.meas[slot, i, :n_feat])
                except Exception as e:  # e.g. a singular innovation covariance on a diverged run (msckf.py:605-612 raises too)
                    dead[streams[i]] = f'{type(e).__name__}: {e}'
                    continue
                frames += 1
                if r is not None:
                    traj[i].append(state_row(est))
            busy += time.perf_counter() - t0
            free.release()
            slot = (slot + 1) % depth
        conn.send({'traj': {s: np.array(v, dtype=np.float64).reshape(-1, 8) for s, v in zip(streams, traj)},
                   'busy_s': busy, 'frames': frames, 'errors': dead})
        conn.close()
        del ring
    finally:
        shm.close()


class EstimatorPool:
    """`n_streams` MSCKF instances spread over `n_workers` processes.  `push_step` hands one frame of every stream to
    the workers and returns at once: every worker has its own ring `depth` steps deep (filter time per frame varies 3x
    around its median, so workers must be allowed to drift apart; a full ring blocks the producer, which is the
    back-pressure).  `capacity` = the most features a frame can carry (grid_num * grid_max_feature_num).  `finish`
    collects per stream the published trajectory rows (t, x, y, z, qx, qy, qz, qw: the reference's output-file
    columns, msckf.py:152-160)."""

    def __init__(self, config, n_streams, n_workers, capacity=None, method='spawn', depth=32):
        self.S, self.P = int(n_streams), max(1, min(int(n_workers), int(n_streams)))
        self.cap = int(capacity if capacity is not None else config.grid_num * config.grid_max_feature_num)
        self.depth = int(depth)
        ctx = mp.get_context(method)
        self.workers = []
        for w in range(self.P):
            streams = list(range(w, self.S, self.P))
            shm = shared_memory.SharedMemory(create=True, size=_Ring.size(self.depth, len(streams), self.cap))
            ring = _Ring(shm.buf, self.depth, len(streams), self.cap)
            filled, free = ctx.Semaphore(0), ctx.Semaphore(self.depth)
            parent, child = ctx.Pipe(duplex=False)
            p = ctx.Process(target=_worker, args=(shm.name, self.depth, self.cap, filled, free, child, config, streams),
                            daemon=True)
            p.start()
            child.close()
            self.workers.append(dict(streams=streams, shm=shm, ring=ring, filled=filled, free=free, conn=parent, proc=p,
                                     slot=0))
        for w in self.workers:          # interpreter start + imports + filter construction: ~1 s, all workers in parallel
            try:
                ok = w['conn'].poll(120) and w['conn'].recv() == 'ready'
            except (EOFError, OSError):
                ok = False
            if not ok:
                self.close()
                raise RuntimeError('estimator worker failed to start (see its traceback above)')

    def _post(self, w, kind, fill):
        while not w['free'].acquire(timeout=5.0):                   # ring full: wait, but not for a dead worker
            if not w['proc'].is_alive():
                raise RuntimeError('estimator worker died (see its traceback above)')
        slot = w['slot']
        if fill is not None:
            fill(w['ring'], slot)
        w['ring'].kind[slot] = kind
        w['slot'] = (slot + 1) % self.depth
        w['filled'].release()

    def push_step(self, items):
        """items[s] = (imu_rows float64[m, 7], timestamp, ids int64[n], meas float64[n, 4])."""
        if len(items) != self.S:
            raise ValueError(f'expected {self.S} items')
        for w in self.workers:
            mine = [items[s] for s in w['streams']]
            done = [0] * len(mine)
            while any(len(it[0]) - d > MAX_IMU for it, d in zip(mine, done)):       # IMU backlog: IMU-only slots first
                def fill_imu(ring, slot):
                    for i, it in enumerate(mine):
                        k = max(0, min(MAX_IMU, len(it[0]) - done[i] - MAX_IMU))
                        ring.imu[slot, i, :k] = it[0][done[i]:done[i] + k]
                        ring.count[slot, i] = (k, 0)
                        done[i] += k
                self._post(w, IMU_ONLY, fill_imu)

            def fill(ring, slot):
                for i, (imu_rows, ts, ids, meas) in enumerate(mine):
                    rows = imu_rows[done[i]:]
                    n = len(ids)
                    if n > self.cap:
                        raise ValueError(f'{n} features in a frame exceed the pool capacity {self.cap}')
                    ring.imu[slot, i, :len(rows)] = rows
                    ring.ts[slot, i] = ts
                    ring.ids[slot, i, :n] = ids
                    ring.meas[slot, i, :n] = meas
                    ring.count[slot, i] = (len(rows), n)
            self._post(w, FRAME, fill)

    def finish(self):
        for w in self.workers:
            self._post(w, STOP, None)
        self._stopped = True
        traj, busy, frames, errors = {}, [], 0, {}
        for w in self.workers:
            try:
                r = w['conn'].recv()
            except (EOFError, OSError):
                raise RuntimeError('estimator worker died before delivering its trajectories') from None
            traj.update(r['traj'])
            busy.append(r['busy_s'])
            frames += r['frames']
            errors.update(r.get('errors', {}))
        # `errors`: stream -> text for filters that raised (their trajectories end where they stopped)
        return [traj[s] for s in range(self.S)], {'worker_busy_s': busy, 'frames': frames, 'errors': errors}

    def _release(self):
        for w in self.workers:
            if w.get('shm') is not None:
                w['ring'] = None
                try:
                    w['shm'].close()
                    w['shm'].unlink()
                except (FileNotFoundError, BufferError):
                    pass
                w['shm'] = None

    def close(self):
        """Joins the workers (they exit on their own after `finish`; anything still running after 5 s is terminated) and
        frees the rings."""
        for w in self.workers:
            if getattr(self, '_stopped', False):
                w['proc'].join(timeout=5)
            if w['proc'].is_alive():
                w['proc'].terminate()
                w['proc'].join(timeout=5)
        self._release()
