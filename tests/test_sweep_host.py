"""Host side of the sequence x offset sweep (SURVEY.md section 8 rows e/f2, BASELINE configs C4/C5) on CPU: the start
rule of an offset run against the EuRoC reader, and the estimator worker pool against the filter run in-process."""
import os

import numpy as np
import pytest

from synth_euroc import SlidingTextureStream


def test_offset_start_equals_the_dataset_reader(tmp_path):
    """`--offset s` drops everything older than first IMU stamp + s (dataset.py:203, 206-214): the cached
    sequence's start indices select exactly the messages the reader yields."""
    from euroc import EuRoCDataset, write_euroc
    from sweep import offset_start
    st = SlidingTextureStream(width=96, height=80, n_frames=14, seed=5, gyro=(0.01, 0.0, 0.02))
    write_euroc(str(tmp_path / 'seq'), st)
    ds = EuRoCDataset(str(tmp_path / 'seq'))
    ds.set_starttime(0)
    frame_t = np.array([f.timestamp for f in ds.stereo])
    imu_t = np.array([m.timestamp for m in ds.imu])
    for off in (0.0, 0.07, 0.22, 0.41):
        ds.set_starttime(off)
        want_f = [f.timestamp for f in ds.stereo]
        want_i = [m.timestamp for m in ds.imu]
        k0, j0 = offset_start(frame_t, imu_t, off)
        assert list(frame_t[k0:]) == want_f and list(imu_t[j0:]) == want_i, off
    k0, j0 = offset_start(frame_t, imu_t, 100.0)             # beyond the end: nothing left
    assert k0 == len(frame_t) and j0 == len(imu_t)


def test_estimator_pool_equals_in_process_filter(golden_dir):
    """3 streams over 2 worker processes: every stream's trajectory equals the filter fed directly."""
    from estimator_pool import EstimatorPool, feed, state_row
    from frontend_config import FrontEndConfig, with_filter_fields
    from msckf import MSCKF
    cfg = with_filter_fields(FrontEndConfig())
    z = np.load(os.path.join(golden_dir, 'ate_gpu_features.npz'))
    imu = np.load(os.path.join(golden_dir, 'ref_msckf_traj.npz'))['imu']
    n = 60
    frames, j = [], 0
    for k in range(n):
        if 22 <= k < 38:                                      # 16 dropped frames: frame 38 arrives with 170 IMU rows, more
            continue                                          # than a ring slot holds -> IMU-only slots travel ahead of it
        ts = float(z[f'f{k}_ts'][0])
        j1 = j
        while j1 < len(imu) and imu[j1, 0] <= ts:
            j1 += 1
        frames.append((imu[j:j1], ts, z[f'f{k}_ids'], z[f'f{k}_meas']))
        j = j1
    est = MSCKF(cfg, outfile=False)
    want = []
    for fr in frames:
        r = feed(est, *fr)
        if r is not None:
            want.append(state_row(est))
    want = np.array(want)
    assert len(want) >= 20 and max(len(f[0]) for f in frames) > 128
    pool = EstimatorPool(cfg, 3, 2, capacity=300, depth=4)
    try:
        for k, fr in enumerate(frames):
            # stream 2 lags one frame behind the others: streams are independent
            pool.push_step([fr, fr, frames[max(k - 1, 0)]])
        traj, stats = pool.finish()
    finally:
        pool.close()
    assert stats['frames'] == 3 * len(frames) and len(stats['worker_busy_s']) == 2
    # the workers cap BLAS at one thread, this process does not: same arithmetic up to the summation order inside BLAS
    assert traj[0].shape == want.shape and np.array_equal(traj[0], traj[1])
    assert np.abs(traj[0] - want).max() < 1e-9 and not stats['errors']
    assert traj[2].shape[1] == 8 and len(traj[2]) >= len(want) - 2
    with pytest.raises(ValueError):
        EstimatorPool.push_step(pool, [frames[0]])


def test_estimator_pool_reports_a_worker_that_cannot_start():
    """A configuration the filter cannot be built from: the pool raises instead of hanging on the ready handshake."""
    from estimator_pool import EstimatorPool
    from frontend_config import FrontEndConfig
    with pytest.raises(RuntimeError, match='failed to start'):
        EstimatorPool(FrontEndConfig(), 2, 1)                 # no filter fields (with_filter_fields not applied)
