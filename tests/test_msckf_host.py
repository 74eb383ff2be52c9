"""Host MSCKF (uav-airvision_b200/msckf.py) against the UNMODIFIED reference filter: the same
400-frame feature dump + IMU stream through both; the reference's trajectory is the committed fixture
tests/golden/ref_msckf_traj.npz (tools/make_msckf_golden.py imports /root/reference/src/msckf.py to write it)."""
import os
import time
from collections import namedtuple

import numpy as np
import pytest

imu_msg = namedtuple('imu_msg', ['timestamp', 'angular_velocity', 'linear_acceleration'])
feature_msg = namedtuple('feature_msg', ['timestamp', 'features'])
Meas = namedtuple('FeatureMeasurement', ['id', 'u0', 'v0', 'u1', 'v1'])


def _msckf_config():
    """The filter fields of the reference's ConfigEuRoC (/root/reference/src/config.py:7-17, 47-72, 93-122)."""
    from frontend_config import FrontEndConfig

    class Opt:
        translation_threshold = -1.0
        huber_epsilon = 0.01
        estimation_precision = 5e-7
        initial_damping = 1e-3
        outer_loop_max_iteration = 5
        inner_loop_max_iteration = 5

    cfg = FrontEndConfig()
    cfg.optimization_config = Opt()
    cfg.gravity = np.array([0.0, 0.0, -9.81])
    cfg.max_cam_state_size = 20
    cfg.position_std_threshold = 2.0
    cfg.gyro_noise, cfg.acc_noise = 0.005 ** 2, 0.05 ** 2
    cfg.gyro_bias_noise, cfg.acc_bias_noise = 0.001 ** 2, 0.01 ** 2
    cfg.observation_noise = 0.035 ** 2
    cfg.velocity = np.zeros(3)
    cfg.velocity_cov, cfg.gyro_bias_cov, cfg.acc_bias_cov = 0.25, 0.01, 0.01
    cfg.extrinsic_rotation_cov, cfg.extrinsic_translation_cov = 3.0462e-4, 2.5e-5
    cfg.T_cn_cnm1 = np.array([
        [0.999997256477881, 0.002312067192424, 0.000376008102415, -0.110073808127187],
        [-0.002317135723281, 0.999898048506644, 0.014089835846648, 0.000399121547014],
        [-0.000343393120525, -0.014090668452714, 0.999900662637729, -0.000853702503357],
        [0, 0, 0, 1.0]])
    cfg.T_imu_body = np.identity(4)
    return cfg


def _replay(golden_dir, n_frames=None, **kw):
    from msckf import MSCKF
    z = np.load(os.path.join(golden_dir, 'ate_gpu_features.npz'))
    g = np.load(os.path.join(golden_dir, 'ref_msckf_traj.npz'))
    imu = g['imu']
    n = int(z['n_frames'][0]) if n_frames is None else n_frames
    est = MSCKF(_msckf_config(), outfile=False, **kw)
    rows, secs, j = [], [], 0
    for k in range(n):
        ts = float(z[f'f{k}_ts'][0])
        while j < len(imu) and imu[j, 0] <= ts:
            est.imu_callback(imu_msg(imu[j, 0], imu[j, 1:4].copy(), imu[j, 4:7].copy()))
            j += 1
        feats = [Meas(int(i), *row) for i, row in zip(z[f'f{k}_ids'], z[f'f{k}_meas'].tolist())]
        t0 = time.perf_counter()
        r = est.feature_callback(feature_msg(ts, feats))
        secs.append(time.perf_counter() - t0)
        if r is not None:
            st = est.imu_state
            rows.append([k, r.timestamp, *r.pose.t, *st.orientation, *st.velocity, len(est.cams), len(est.map_server)])
    return np.array(rows), g['traj'], np.array(secs), float(g['ref_ms_per_frame_median'][0]), est


def test_trajectory_equals_reference_filter(golden_dir):
    got, want, secs, ref_ms, est = _replay(golden_dir)
    assert got.shape == want.shape == (380, 14)
    assert np.array_equal(got[:, 0], want[:, 0]) and np.array_equal(got[:, 1], want[:, 1])
    # window and map bookkeeping identical on every frame: same features used, gated and pruned
    assert np.array_equal(got[:, 12:], want[:, 12:])
    dp = np.abs(got[:, 2:5] - want[:, 2:5]).max()
    dq = np.abs(got[:, 5:9] - want[:, 5:9]).max()
    dv = np.abs(got[:, 9:12] - want[:, 9:12]).max()
    path = np.linalg.norm(np.diff(want[:, 2:5], axis=0), axis=1).sum()
    print(f'host MSCKF vs reference over {len(got)} frames ({path:.2f} m): max |dp| {dp:.3g} m, |dq| {dq:.3g}, |dv| {dv:.3g} m/s; '
          f'{1e3 * np.median(secs[20:]):.2f} ms/frame here vs {ref_ms:.2f} ms/frame for the reference when the fixture was made')
    assert dp < 1e-6 and dq < 1e-6 and dv < 1e-5
    assert est.large_update_count == 0


def test_c_inner_loops_equal_their_numpy_statement(golden_dir):
    """_msckfhost (propagation, triangulation, Jacobian blocks, null-space projection) against the numpy statements kept
    in msckf.py: the same 120 frames through both."""
    a = _replay(golden_dir, 120, use_c=True)[0]
    b = _replay(golden_dir, 120, use_c=False)[0]
    assert a.shape == b.shape and len(a) == 100 and np.array_equal(a[:, 12:], b[:, 12:])
    assert np.abs(a[:, 2:12] - b[:, 2:12]).max() < 1e-9


def test_array_entry_point_equals_the_message_entry_point(golden_dir):
    """feature_callback_arrays(ts, ids, meas) (what a sweep's estimator processes call) is feature_callback(feature_msg)
    without the per-feature objects: identical states on every frame."""
    from msckf import MSCKF
    z = np.load(os.path.join(golden_dir, 'ate_gpu_features.npz'))
    imu = np.load(os.path.join(golden_dir, 'ref_msckf_traj.npz'))['imu']
    a, b = MSCKF(_msckf_config(), outfile=False), MSCKF(_msckf_config(), outfile=False)
    j = 0
    for k in range(70):
        ts = float(z[f'f{k}_ts'][0])
        while j < len(imu) and imu[j, 0] <= ts:
            for est in (a, b):
                est.imu_callback(imu_msg(imu[j, 0], imu[j, 1:4].copy(), imu[j, 4:7].copy()))
            j += 1
        ids, meas = z[f'f{k}_ids'], z[f'f{k}_meas']
        ra = a.feature_callback(feature_msg(ts, [Meas(int(i), *row) for i, row in zip(ids, meas.tolist())]))
        rb = b.feature_callback_arrays(ts, ids, meas)
        assert (ra is None) == (rb is None)
        if ra is not None:
            assert np.array_equal(ra.pose.t, rb.pose.t) and np.array_equal(a.imu_state.orientation, b.imu_state.orientation)
            assert np.array_equal(a.state_cov, b.state_cov) and list(a.map_server) == list(b.map_server)
    assert len(a.cams) == len(b.cams) > 10


def test_returns_none_until_gravity_is_initialised_and_reset(golden_dir):
    from msckf import MSCKF
    est = MSCKF(_msckf_config(), outfile=False)
    assert est.feature_callback(feature_msg(1.0, [])) is None               # msckf.py:182-183
    for k in range(199):
        est.imu_callback(imu_msg(0.005 * k, np.array([0.01, 0.0, 0.0]), np.array([0.0, 0.0, 9.81])))
    assert not est.is_gravity_set
    est.imu_callback(imu_msg(1.0, np.array([0.01, 0.0, 0.0]), np.array([0.0, 0.0, 9.81])))
    assert est.is_gravity_set and abs(est.imu_state.gyro_bias[0] - 0.01) < 1e-12
    assert np.allclose(est.gravity, [0, 0, -9.81])
    r = est.feature_callback(feature_msg(1.0, [Meas(0, 0.1, 0.0, 0.05, 0.0)]))
    assert r is not None and r.pose.R.shape == (3, 3) and len(est.cams) == 1 and len(est.map_server) == 1
    est.reset()
    assert not est.is_gravity_set and len(est.cams) == 0 and est.state_cov.shape == (21, 21)


def test_output_file_format(tmp_path, golden_dir):
    """One line per published state: 't x y z qx qy qz qw' (msckf.py:152-160)."""
    from msckf import MSCKF
    out = tmp_path / 'traj.txt'
    est = MSCKF(_msckf_config(), outfile=str(out))
    for k in range(200):
        est.imu_callback(imu_msg(0.005 * k, np.zeros(3), np.array([0.0, 0.0, 9.81])))
    est.feature_callback(feature_msg(1.0, []))
    cols = out.read_text().strip().split()
    assert len(cols) == 8 and abs(float(cols[0]) - 1.0) < 1e-9


def test_gate_with_an_indefinite_covariance_follows_the_reference_lu_path(golden_dir):
    """The reference's gating_test (msckf.py:605-612) solves S gamma-wise with an LU and only raises when S is exactly
    singular; its (I - KH) P update does not keep P positive definite on a diverging run.  The C gate (Cholesky) must
    not raise then: it marks the feature and the numpy LU statement supplies the statistic."""
    import _msckfhost as C
    g = np.random.default_rng(4)
    F, m, n = 5, 3, 21 + 6 * 4
    Hx, Hf, r = g.normal(size=(F, m, 4, 6)), g.normal(size=(F, 4 * m, 3)), g.normal(size=(F, 4 * m))
    slots = np.tile(np.array([0, 2, 3], np.int64), (F, 1))
    A = g.normal(size=(n, n))
    P_pd = A @ A.T + np.eye(n)
    P_bad = P_pd - 40.0 * np.eye(n)                                         # symmetric, indefinite
    gam = np.empty(F)
    C.gate(Hx, Hf, r, P_pd, slots, 0.035 ** 2, gam)
    assert np.isfinite(gam).all()
    C.gate(Hx, Hf, r, P_bad, slots, 0.035 ** 2, gam)                        # used to raise ArithmeticError
    assert np.isnan(gam).all()
    # through the filter: a covariance made indefinite mid-run; the frame is processed, decisions come from the LU path
    from msckf import MSCKF
    est_c, est_np = _replay(golden_dir, 45)[4], _replay(golden_dir, 45, use_c=False)[4]
    z = np.load(os.path.join(golden_dir, 'ate_gpu_features.npz'))
    imu = np.load(os.path.join(golden_dir, 'ref_msckf_traj.npz'))['imu']
    ts = float(z['f45_ts'][0])
    feats = [Meas(int(i), *row) for i, row in zip(z['f45_ids'], z['f45_meas'].tolist())]
    out = []
    for est in (est_c, est_np):
        est.state_cov[21:, 21:] -= 1e-3 * np.eye(len(est.state_cov) - 21)  # camera blocks: indefinite
        for row in imu[(imu[:, 0] > float(z['f44_ts'][0])) & (imu[:, 0] <= ts)]:
            est.imu_callback(imu_msg(row[0], row[1:4].copy(), row[4:7].copy()))
        res = est.feature_callback(feature_msg(ts, feats))
        assert res is not None
        out.append((res.pose.t.copy(), len(est.map_server), len(est.cams)))
    assert out[0][1:] == out[1][1:] and np.abs(out[0][0] - out[1][0]).max() < 1e-6
