"""Host-side MSCKF (stereo multi-state constraint Kalman filter): the consumer of the front end's feature messages.

Drop-in for the reference's `msckf.MSCKF` (/root/reference/src/msckf.py:96-867 with src/feature/*.py and src/utils.py):
same constructor argument, `imu_callback(imu_msg)`, `feature_callback(feature_msg) -> vio_result | None`, same filter
(Sun et al., "Robust Stereo Visual Inertial Odometry for Fast Autonomous Flight", 2018) with every behaviour of the
reference that shapes the trajectory kept: 21-dimensional IMU error state, third-order transition matrix with the
observability-constrained null-space fix (msckf.py:274-338), RK4 state prediction (:340-388), 6-dof camera-state
augmentation (:390-423), Levenberg-Marquardt triangulation (feature/feature_position_initializer.py:6-76), left
null-space projection, chi-square gating at chi2.ppf(0.05, dof) (:111-113, :605-612), measurement compression,
the `(I - K H) P` covariance update (:600-603), the two-out-of-the-window pruning rule (:678-786), online reset (:822-843).

SURVEY.md section 8(f3): this runs on the host, not on the GPU -- at ~0.13 ms per frame the front end leaves the filter as
the bottleneck of "full front end + host MSCKF" (BASELINE config C5).  What is different from the reference is how the
work is laid out, not what is computed:
  * the camera-state window lives in arrays (quaternions, positions, cached rotation matrices), not in per-state objects;
  * the transition matrix is assembled from its 3x3 blocks in closed form instead of three dense 21x21 products, the
    cross-covariance is propagated once per image with the accumulated transition matrix of the IMU batch;
  * the measurement Jacobians of ALL observations of all features with the same number of camera states are formed in one
    pass (per-observation blocks in C), the null-space basis comes from a complete QR of the 4m x 3 feature Jacobian (the
    update and the gate are invariant to the choice of orthonormal basis), accepted features stay batched up to the
    stacked matrix, which is kept compact in its non-zero columns; the update is computed from H^T H and H^T r on those
    columns (what the reference's thin QR compresses to, without the QR: same K r and (I - K H) P);
  * no printing (the reference prints ~15 lines per frame), no per-frame file open unless an output file is asked for.
Results agree with the reference to rounding (tests/test_msckf_host.py replays a 400-frame feature dump against the
reference filter's committed trajectory).
"""
from __future__ import annotations

import os
from collections import namedtuple
from itertools import chain

import numpy as np

try:                                    # scalar inner loops in C (csrc/msckf_host.c); the numpy statements below stay as
    import _msckfhost as _C             # the readable definition and are compared with it in tests/test_msckf_host.py
except ImportError:                     # pragma: no cover - build.py compiles it
    _C = None

vio_result = namedtuple('vio_result', ['timestamp', 'pose', 'velocity', 'cam0_pose'])

_I3 = np.identity(3)


# ---- small rotation helpers (JPL quaternions [x, y, z, w]; reference src/utils.py) ------------------------------------

def skew(v):
    return np.array([[0.0, -v[2], v[1]], [v[2], 0.0, -v[0]], [-v[1], v[0], 0.0]])


def to_rotation(q):
    """utils.py:12-23 (Trawny & Roumeliotis eq. 78)."""
    q = q / np.sqrt(q @ q)
    v, w = q[:3], q[3]
    return (2.0 * w * w - 1.0) * _I3 - 2.0 * w * skew(v) + 2.0 * np.outer(v, v)


def to_quaternion(R):
    """utils.py:25-47."""
    if R[2, 2] < 0:
        if R[0, 0] > R[1, 1]:
            t = 1 + R[0, 0] - R[1, 1] - R[2, 2]
            q = [t, R[0, 1] + R[1, 0], R[2, 0] + R[0, 2], R[1, 2] - R[2, 1]]
        else:
            t = 1 - R[0, 0] + R[1, 1] - R[2, 2]
            q = [R[0, 1] + R[1, 0], t, R[2, 1] + R[1, 2], R[2, 0] - R[0, 2]]
    else:
        if R[0, 0] < -R[1, 1]:
            t = 1 - R[0, 0] - R[1, 1] + R[2, 2]
            q = [R[0, 2] + R[2, 0], R[2, 1] + R[1, 2], t, R[0, 1] - R[1, 0]]
        else:
            t = 1 + R[0, 0] + R[1, 1] + R[2, 2]
            q = [R[1, 2] - R[2, 1], R[2, 0] - R[0, 2], R[0, 1] - R[1, 0], t]
    q = np.array(q)
    return q / np.sqrt(q @ q)


def quaternion_multiplication(q1, q2):
    """utils.py:61-76."""
    q1 = q1 / np.sqrt(q1 @ q1)
    q2 = q2 / np.sqrt(q2 @ q2)
    x, y, z, w = q1
    q = np.array([w * q2[0] + z * q2[1] - y * q2[2] + x * q2[3],
                  -z * q2[0] + w * q2[1] + x * q2[2] + y * q2[3],
                  y * q2[0] - x * q2[1] + w * q2[2] + z * q2[3],
                  -x * q2[0] - y * q2[1] - z * q2[2] + w * q2[3]])
    return q / np.sqrt(q @ q)


def small_angle_quaternion(dtheta):
    """utils.py:79-93."""
    dq = dtheta / 2.0
    n2 = dq @ dq
    if n2 <= 1:
        return np.array([dq[0], dq[1], dq[2], np.sqrt(1 - n2)])
    return np.array([dq[0], dq[1], dq[2], 1.0]) / np.sqrt(1 + n2)


def from_two_vectors(v0, v1):
    """utils.py:96-120: rotation taking v0 to v1, returned in the JPL convention."""
    v0 = v0 / np.linalg.norm(v0)
    v1 = v1 / np.linalg.norm(v1)
    d = v0 @ v1
    if d < -0.999999:
        axis = np.cross([1, 0, 0], v0)
        if np.linalg.norm(axis) < 0.000001:
            axis = np.cross([0, 1, 0], v0)
        q = np.array([*axis, 0.0])
    elif d > 0.999999:
        q = np.array([0.0, 0.0, 0.0, 1.0])
    else:
        s = np.sqrt((1 + d) * 2)
        q = np.array([*(np.cross(v0, v1) / s), 0.5 * s])
    q = q / np.linalg.norm(q)
    return np.array([-q[0], -q[1], -q[2], q[3]])


def _rotations(q):
    """to_rotation for a stack of unit-or-not quaternions (n, 4) -> (n, 3, 3)."""
    q = q / np.sqrt((q * q).sum(axis=1))[:, None]
    v, w = q[:, :3], q[:, 3]
    R = 2.0 * v[:, :, None] * v[:, None, :]
    d = 2.0 * w * w - 1.0
    R[:, 0, 0] += d
    R[:, 1, 1] += d
    R[:, 2, 2] += d
    tw = 2.0 * w
    R[:, 0, 1] += tw * v[:, 2]
    R[:, 0, 2] -= tw * v[:, 1]
    R[:, 1, 0] -= tw * v[:, 2]
    R[:, 1, 2] += tw * v[:, 0]
    R[:, 2, 0] += tw * v[:, 1]
    R[:, 2, 1] -= tw * v[:, 0]
    return R


class Isometry3d:
    """Rigid transform (utils.py:124-140): what vio_result.pose / cam0_pose carry."""
    __slots__ = ('R', 't')

    def __init__(self, R, t):
        self.R, self.t = R, t

    def matrix(self):
        m = np.identity(4)
        m[:3, :3], m[:3, 3] = self.R, self.t
        return m

    def inverse(self):
        return Isometry3d(self.R.T, -self.R.T @ self.t)

    def __mul__(self, o):
        return Isometry3d(self.R @ o.R, self.R @ o.t + self.t)


# ---- state ------------------------------------------------------------------------------------------------------------

class IMUState:
    """Current IMU state (msckf.py:18-58).  `gravity` and `T_imu_body` are per filter instance here (class attributes
    in the reference, shared by every filter of a process)."""

    def __init__(self):
        self.id = None
        self.timestamp = None
        self.orientation = np.array([0.0, 0.0, 0.0, 1.0])        # world -> IMU
        self.position = np.zeros(3)
        self.velocity = np.zeros(3)
        self.gyro_bias = np.zeros(3)
        self.acc_bias = np.zeros(3)
        self.orientation_null = np.array([0.0, 0.0, 0.0, 1.0])
        self.position_null = np.zeros(3)
        self.velocity_null = np.zeros(3)
        self.R_imu_cam0 = np.identity(3)
        self.t_cam0_imu = np.zeros(3)


class CamWindow:
    """The sliding window of camera states in insertion order (the reference's dict <id, CAMState>, msckf.py:61-77,
    as arrays).  Slot i owns covariance rows 21 + 6 i .. 21 + 6 i + 5."""

    def __init__(self, capacity=64):
        self.ids = []
        self.index = {}                                           # id -> slot
        self.q = np.zeros((capacity, 4))
        self.p = np.zeros((capacity, 3))
        self.q_null = np.zeros((capacity, 4))
        self.p_null = np.zeros((capacity, 3))
        self.R = np.zeros((capacity, 3, 3))                       # to_rotation(q), refreshed whenever q changes
        self.R_null = np.zeros((capacity, 3, 3))

    def __len__(self):
        return len(self.ids)

    def __contains__(self, cam_id):
        return cam_id in self.index

    def clear(self):
        self.ids, self.index = [], {}

    def append(self, cam_id, q, p):
        n = len(self.ids)
        if n == len(self.q):
            for name in ('q', 'p', 'q_null', 'p_null', 'R', 'R_null'):
                a = getattr(self, name)
                setattr(self, name, np.concatenate([a, np.zeros_like(a)]))
        self.ids.append(cam_id)
        self.index[cam_id] = n
        self.q[n], self.p[n] = q, p
        self.q_null[n], self.p_null[n] = q, p
        self.R[n] = self.R_null[n] = to_rotation(q)

    def remove(self, cam_id):
        i = self.index[cam_id]
        n = len(self.ids)
        for name in ('q', 'p', 'q_null', 'p_null', 'R', 'R_null'):
            a = getattr(self, name)
            a[i:n - 1] = a[i + 1:n]
        del self.ids[i]
        self.index = {c: k for k, c in enumerate(self.ids)}
        return i


class Feature:
    """One map feature: its observations by camera state, in the order they arrived (feature/base_feature.py:3-13)."""
    __slots__ = ('id', 'observations', 'position', 'is_initialized')

    def __init__(self, fid):
        self.id = fid
        self.observations = {}                                    # cam state id -> (u0, v0, u1, v1)
        self.position = np.zeros(3)
        self.is_initialized = False


class MSCKF:
    def __init__(self, config, outfile=None, use_c=True, blas_threads=None):
        """`outfile`: None = the reference's rule (results/txts/output_<DATASET_NAME>_offset<TIME_OFFSET>.txt, one line
        appended per published state, msckf.py:10-16,152-160); a path = write there; False = do not write.
        `use_c`: run the IMU propagation and the triangulation through _msckfhost (False: the numpy statements).
        `blas_threads`: opt-in, PROCESS-WIDE BLAS thread limit (None, the default, leaves the hosting process alone).  The
        filter's matrices are at most 1500 x 141 and a threaded OpenBLAS is 2x SLOWER on them than one thread, so the
        estimator worker processes (estimator_pool._worker), which run nothing else, pass 1."""
        self.use_c = bool(use_c) and _C is not None
        self._gamma = None                  # gate statistics of the last _jacobians call (C path)
        if blas_threads is not None:
            try:
                from threadpoolctl import threadpool_limits
                MSCKF._blas_limit = threadpool_limits(limits=int(blas_threads), user_api='blas')
            except Exception:           # pragma: no cover - threadpoolctl is optional
                pass
        self.config = config
        self.optimization_config = config.optimization_config
        self.imu_msg_buffer = []
        self.imu_state = IMUState()
        self.cams = CamWindow()
        self.state_cov = np.zeros((21, 21))
        self.map_server = {}                                      # feature id -> Feature, in first-seen order
        from scipy.stats import chi2
        self.chi_squared_test_table = {i: float(chi2.ppf(0.05, i)) for i in range(1, 100)}    # msckf.py:111-113

        self.imu_state.velocity = np.array(config.velocity, dtype=np.float64)
        self.reset_state_cov()
        # G Qc G^T (msckf.py:122-127, 290-293, 330): block diagonal; R^T (n_a I) R = n_a I
        self._noise = np.zeros(21)
        self._noise[0:3] = config.gyro_noise
        self._noise[3:6] = config.gyro_bias_noise
        self._noise[6:9] = config.acc_noise
        self._noise[9:12] = config.acc_bias_noise
        self.gravity = np.array(config.gravity, dtype=np.float64)

        T_cam0_imu = np.linalg.inv(config.T_imu_cam0)
        self.imu_state.R_imu_cam0 = T_cam0_imu[:3, :3].T.copy()
        self.imu_state.t_cam0_imu = T_cam0_imu[:3, 3].copy()
        T01 = np.asarray(config.T_cn_cnm1, dtype=np.float64)
        self.R_cam0_cam1 = T01[:3, :3].copy()
        self.t_cam0_cam1 = T01[:3, 3].copy()
        Tb = np.asarray(config.T_imu_body, dtype=np.float64)
        self.T_imu_body = Isometry3d(Tb[:3, :3].copy(), Tb[:3, 3].copy())

        self.next_state_id = 0
        self.tracking_rate = None
        self.is_gravity_set = False
        self.is_first_img = True
        self.large_update_count = 0                               # the reference prints a warning (msckf.py:571-573)
        if outfile is None:
            base = 'results/txts'
            os.makedirs(base, exist_ok=True)
            outfile = os.path.join(base, 'output_%s_offset%s.txt' % (os.getenv('DATASET_NAME', 'unknown'),
                                                                     os.getenv('TIME_OFFSET', '0')))
        self._outfile = outfile or None

    # -- callbacks ------------------------------------------------------------------------------------------------------
    def imu_callback(self, imu_msg):
        """msckf.py:162-175: buffer; after 200 samples initialise gravity direction and gyro bias."""
        self.imu_msg_buffer.append(imu_msg)
        if not self.is_gravity_set and len(self.imu_msg_buffer) >= 200:
            self.initialize_gravity_and_bias()
            self.is_gravity_set = True

    def feature_callback(self, feature_msg):
        """msckf.py:177-228."""
        if not self.is_gravity_set:
            return None
        return self._feature_step(feature_msg.timestamp, lambda: self.add_feature_observations(feature_msg))

    def feature_callback_arrays(self, timestamp, ids, meas):
        """feature_callback for a frame that arrives as arrays (ids int64[n], meas float64[n, 4] = u0 v0 u1 v1), the form
        a sweep's estimator processes receive it in: the same step without building and re-reading 300 objects."""
        if not self.is_gravity_set:
            return None
        return self._feature_step(float(timestamp), lambda: self._add_observation_rows(ids.tolist(), meas.tolist()))

    def _feature_step(self, timestamp, add_observations):
        if self.is_first_img:
            self.is_first_img = False
            self.imu_state.timestamp = timestamp
        self.batch_imu_processing(timestamp)
        self.state_augmentation(timestamp)
        add_observations()
        self.remove_lost_features()
        self.prune_cam_state_buffer()
        try:
            return self.publish(timestamp)
        finally:
            self.online_reset()

    def initialize_gravity_and_bias(self):
        """msckf.py:230-249."""
        n = len(self.imu_msg_buffer)
        sum_w, sum_a = np.zeros(3), np.zeros(3)
        for m in self.imu_msg_buffer:
            sum_w += m.angular_velocity
            sum_a += m.linear_acceleration
        self.imu_state.gyro_bias = sum_w / n
        gravity_imu = sum_a / n
        self.gravity = np.array([0.0, 0.0, -np.linalg.norm(gravity_imu)])
        self.imu_state.orientation = from_two_vectors(-self.gravity, gravity_imu)

    # -- propagation ----------------------------------------------------------------------------------------------------
    def batch_imu_processing(self, time_bound):
        """msckf.py:251-272.  The IMU block of the covariance is propagated sample by sample; the cross terms with the
        camera states only see the product of the transition matrices, so they are updated once per image."""
        st = self.imu_state
        if self.use_c:
            return self._batch_imu_processing_c(time_bound)
        used = 0
        phi_total = None
        for msg in self.imu_msg_buffer:
            t = msg.timestamp
            if t < st.timestamp:
                used += 1
                continue
            if t > time_bound:
                break
            phi = self.process_model(t, msg.angular_velocity, msg.linear_acceleration)
            phi_total = phi if phi_total is None else phi @ phi_total
            used += 1
            st.timestamp = t
        if phi_total is not None and self.state_cov.shape[0] > 21:
            P = self.state_cov
            P[:21, 21:] = phi_total @ P[:21, 21:]
            P[21:, :21] = P[:21, 21:].T
        st.id = self.next_state_id
        self.next_state_id += 1
        del self.imu_msg_buffer[:used]

    def _batch_imu_processing_c(self, time_bound):
        st = self.imu_state
        buf = self.imu_msg_buffer
        k = 0
        while k < len(buf) and buf[k].timestamp <= time_bound:
            k += 1
        imu = np.empty((k, 7))
        for i in range(k):
            m = buf[i]
            imu[i, 0] = m.timestamp
            imu[i, 1:4] = m.angular_velocity
            imu[i, 4:7] = m.linear_acceleration
        q, p, v = st.orientation.copy(), st.position.copy(), st.velocity.copy()
        qn, pn, vn = st.orientation_null.copy(), st.position_null.copy(), st.velocity_null.copy()
        P = self.state_cov
        phi_total = np.empty((21, 21))
        used, processed, t = _C.propagate(q, p, v, np.ascontiguousarray(st.gyro_bias), np.ascontiguousarray(st.acc_bias), qn, pn, vn,
                                          P, P.shape[0], self._noise, self.gravity, imu, float(st.timestamp), float(time_bound),
                                          phi_total)
        if processed:
            st.orientation, st.position, st.velocity = q, p, v
            st.orientation_null, st.position_null, st.velocity_null = q, p, v      # aliases, as in the reference
            st.timestamp = t
            if P.shape[0] > 21:
                P[:21, 21:] = phi_total @ P[:21, 21:]
                P[21:, :21] = P[:21, 21:].T
        st.id = self.next_state_id
        self.next_state_id += 1
        del buf[:used]

    def process_model(self, time, m_gyro, m_acc):
        """msckf.py:274-338.  Returns the 21 x 21 transition matrix of this sample."""
        st = self.imu_state
        dt = time - st.timestamp
        gyro = m_gyro - st.gyro_bias
        acc = m_acc - st.acc_bias
        Rt = to_rotation(st.orientation).T                        # IMU -> world

        # Phi = I + F dt + (F dt)^2 / 2 + (F dt)^3 / 6 with F's five non-zero blocks (msckf.py:283-302), in closed form
        A = -skew(gyro) * dt                                      # theta <- theta
        C = -(Rt @ skew(acc)) * dt                                # v <- theta
        D = -Rt * dt                                              # v <- b_a
        A2 = A @ A
        CA = C @ A
        phi = np.identity(21)
        phi[0:3, 0:3] = _I3 + A + A2 / 2.0 + (A2 @ A) / 6.0
        phi[0:3, 3:6] = -dt * (_I3 + A / 2.0 + A2 / 6.0)          # B + A B / 2 + A^2 B / 6,  B = -I dt
        v_th = C + CA / 2.0 + (CA @ A) / 6.0
        phi[6:9, 3:6] = -dt * (C / 2.0 + CA / 6.0)                # C B / 2 + C A B / 6
        phi[6:9, 9:12] = D
        p_th = dt * (C / 2.0 + CA / 6.0)                          # E C / 2 + E C A / 6,  E = I dt
        phi[12:15, 3:6] = -dt * dt * C / 6.0                      # E C B / 6
        phi[12:15, 6:9] = dt * _I3
        phi[12:15, 9:12] = dt * D / 2.0                           # E D / 2

        self.predict_new_state(dt, gyro, acc)

        # observability constraint: keep the null space of the linearised system (msckf.py:311-328)
        R_kk_1 = to_rotation(st.orientation_null)
        phi[0:3, 0:3] = to_rotation(st.orientation) @ R_kk_1.T
        u = R_kk_1 @ self.gravity
        s = u / (u @ u)
        w1 = skew(st.velocity_null - st.velocity) @ self.gravity
        phi[6:9, 0:3] = v_th - np.outer(v_th @ u - w1, s)
        w2 = skew(dt * st.velocity_null + st.position_null - st.position) @ self.gravity
        phi[12:15, 0:3] = p_th - np.outer(p_th @ u - w2, s)

        # P_II <- Phi (P_II + G Qc G^T dt) Phi^T  ( = Phi P Phi^T + Phi G Qc G^T Phi^T dt, msckf.py:330-332)
        P = self.state_cov
        Pii = P[:21, :21].copy()
        Pii[np.diag_indices(21)] += self._noise * dt
        Pii = phi @ Pii @ phi.T
        P[:21, :21] = (Pii + Pii.T) / 2.0

        st.orientation_null = st.orientation
        st.position_null = st.position
        st.velocity_null = st.velocity
        return phi

    def predict_new_state(self, dt, gyro, acc):
        """4th-order Runge-Kutta on position / velocity, closed-form quaternion step (msckf.py:340-388)."""
        st = self.imu_state
        gn = np.sqrt(gyro @ gyro)
        Omega = np.zeros((4, 4))
        Omega[:3, :3] = -skew(gyro)
        Omega[:3, 3] = gyro
        Omega[3, :3] = -gyro
        q, v, p = st.orientation, st.velocity, st.position
        if gn > 1e-5:
            dq_dt = (np.cos(gn * dt * 0.5) * np.identity(4) + np.sin(gn * dt * 0.5) / gn * Omega) @ q
            dq_dt2 = (np.cos(gn * dt * 0.25) * np.identity(4) + np.sin(gn * dt * 0.25) / gn * Omega) @ q
        else:
            dq_dt = np.cos(gn * dt * 0.5) * (np.identity(4) + Omega * dt * 0.5) @ q
            dq_dt2 = np.cos(gn * dt * 0.25) * (np.identity(4) + Omega * dt * 0.25) @ q
        dR_t = to_rotation(dq_dt).T
        dR_t2 = to_rotation(dq_dt2).T
        g = self.gravity
        k1_v_dot = to_rotation(q).T @ acc + g
        k1_v = v + k1_v_dot * dt / 2.0
        k2_v_dot = dR_t2 @ acc + g
        k2_v = v + k2_v_dot * dt / 2
        k3_v_dot = dR_t2 @ acc + g
        k3_v = v + k3_v_dot * dt
        k4_v_dot = dR_t @ acc + g
        st.orientation = dq_dt / np.sqrt(dq_dt @ dq_dt)
        st.velocity = v + (k1_v_dot + 2 * k2_v_dot + 2 * k3_v_dot + k4_v_dot) * dt / 6.0
        st.position = p + (v + 2 * k1_v + 2 * k2_v + k3_v) * dt / 6.0

    # -- augmentation ---------------------------------------------------------------------------------------------------
    def state_augmentation(self, time):
        """msckf.py:390-423: append the current camera pose and its covariance blocks."""
        st = self.imu_state
        R_i_c, t_c_i = st.R_imu_cam0, st.t_cam0_imu
        R_w_i = to_rotation(st.orientation)
        self.cams.append(st.id, to_quaternion(R_i_c @ R_w_i), st.position + R_w_i.T @ t_c_i)

        J = np.zeros((6, 21))
        J[:3, :3] = R_i_c
        J[:3, 15:18] = _I3
        J[3:6, :3] = skew(R_w_i.T @ t_c_i)
        J[3:6, 12:15] = _I3
        J[3:6, 18:21] = _I3
        old = self.state_cov
        n = old.shape[0]
        P = np.empty((n + 6, n + 6))
        P[:n, :n] = old
        cross = J @ old[:21, :]
        P[n:, :n] = cross
        P[:n, n:] = cross.T
        corner = cross[:, :21] @ J.T
        P[n:, n:] = (corner + corner.T) / 2.0
        self.state_cov = P

    def add_feature_observations(self, feature_msg):
        """msckf.py:425-441."""
        state_id = self.imu_state.id
        before = len(self.map_server)
        tracked = 0
        ms = self.map_server
        for f in feature_msg.features:
            feat = ms.get(f.id)
            if feat is None:
                feat = ms[f.id] = Feature(f.id)
            else:
                tracked += 1
            feat.observations[state_id] = (float(f.u0), float(f.v0), float(f.u1), float(f.v1))
        self.tracking_rate = tracked / (before + 1e-5)

    def _add_observation_rows(self, ids, rows):
        """add_feature_observations from plain lists: ids[i] observed at rows[i] = [u0, v0, u1, v1]."""
        state_id = self.imu_state.id
        ms = self.map_server
        before = len(ms)
        tracked = 0
        for fid, row in zip(ids, rows):
            feat = ms.get(fid)
            if feat is None:
                feat = ms[fid] = Feature(fid)
            else:
                tracked += 1
            feat.observations[state_id] = tuple(row)
        self.tracking_rate = tracked / (before + 1e-5)

    # -- triangulation (feature/*.py) -----------------------------------------------------------------------------------
    def check_motion(self, feat):
        """feature/feature_motion_checker.py:6-39."""
        thr = self.optimization_config.translation_threshold
        if thr < 0:
            return True
        ids = list(feat.observations)
        i0, i1 = self.cams.index[ids[0]], self.cams.index[ids[-1]]
        R0 = self.cams.R[i0].T
        z = feat.observations[ids[0]]
        d = np.array([z[0], z[1], 1.0])
        d = R0 @ (d / np.linalg.norm(d))
        tr = self.cams.p[i1] - self.cams.p[i0]
        return np.linalg.norm(tr - (tr @ d) * d) > thr

    def initialize_position(self, feat):
        """Levenberg-Marquardt on inverse depth over every stereo observation
        (feature/feature_position_initializer.py:6-76, feature_observation.py:4-39, feature_depth_estimator.py:4-14)."""
        oc = self.optimization_config
        cams = self.cams
        slots = [cams.index[c] for c in feat.observations if c in cams.index]
        Z = np.array([feat.observations[cams.ids[s]] for s in slots])             # (m, 4)
        m = len(slots)
        # poses of cam0 / cam1 of every observation, expressed relative to the first cam0: x_ci = R x_c0 + t
        Rw = cams.R[slots]                                                       # world -> cam0_i
        pw = cams.p[slots]
        if self.use_c:                  # the statements below, in one call
            x, y, z, valid = _C.triangulate_world(Rw, pw, np.ascontiguousarray(Z), self.R_cam0_cam1, self.t_cam0_cam1,
                                                  oc.huber_epsilon, oc.estimation_precision, oc.initial_damping,
                                                  oc.outer_loop_max_iteration, oc.inner_loop_max_iteration)
            feat.position = np.array([x, y, z])
            feat.is_initialized = bool(valid)
            return feat.is_initialized
        R0w, p0 = Rw[0], pw[0]
        R_c0 = Rw @ R0w.T                                                        # (m, 3, 3)
        t_c0 = np.einsum('mij,mj->mi', Rw, p0 - pw)
        R_c1 = self.R_cam0_cam1 @ R_c0
        t_c1 = t_c0 @ self.R_cam0_cam1.T + self.t_cam0_cam1
        Rs = np.empty((2 * m, 3, 3))
        ts = np.empty((2 * m, 3))
        Rs[0::2], Rs[1::2] = R_c0, R_c1
        ts[0::2], ts[1::2] = t_c0, t_c1
        zs = Z.reshape(2 * m, 2)

        # two-view initial guess from the first stereo pair
        z1, z2 = zs[0], zs[1]
        mm = Rs[1] @ np.array([z1[0], z1[1], 1.0])
        a = mm[:2] - z2 * mm[2]
        b = z2 * ts[1][2] - ts[1][:2]
        depth = (a @ b) / (a @ a)
        sol = np.array([z1[0], z1[1], 1.0 / depth])                              # (alpha, beta, rho)

        if self.use_c:
            sol = np.array(_C.triangulate(Rs, ts, np.ascontiguousarray(zs), sol, oc.huber_epsilon, oc.estimation_precision,
                                          oc.initial_damping, oc.outer_loop_max_iteration, oc.inner_loop_max_iteration))
        else:
            sol = self._levenberg_marquardt(Rs, ts, zs, sol)
        final = np.array([sol[0], sol[1], 1.0]) / sol[2]
        depths = Rs[:, 2, :] @ final + ts[:, 2]
        valid = bool((depths > 0).all())
        feat.position = R0w.T @ final + p0
        feat.is_initialized = valid
        return valid

    def _levenberg_marquardt(self, Rs, ts, zs, sol):
        """numpy statement of the triangulation loop (feature_position_initializer.py:31-70); `inner` counts over the whole
        optimisation, as in the reference."""
        oc = self.optimization_config
        m2 = len(Rs)
        def project(x):
            h = Rs[:, :, 0] * x[0] + Rs[:, :, 1] * x[1] + Rs[:, :, 2] + x[2] * ts
            return h

        def cost(x):
            h = project(x)
            e = h[:, :2] / h[:, 2:3] - zs
            return float((e * e).sum())

        lambd = oc.initial_damping
        outer = inner = 0
        delta_norm = float('inf')
        total = cost(sol)
        W = np.empty((m2, 3, 3))
        W[:, :, :2] = Rs[:, :, :2]
        W[:, :, 2] = ts
        while outer < oc.outer_loop_max_iteration and delta_norm > oc.estimation_precision:
            h = project(sol)
            h3 = h[:, 2]
            J = np.empty((m2, 2, 3))
            J[:, 0] = W[:, 0] / h3[:, None] - W[:, 2] * (h[:, 0] / (h3 * h3))[:, None]
            J[:, 1] = W[:, 1] / h3[:, None] - W[:, 2] * (h[:, 1] / (h3 * h3))[:, None]
            r = h[:, :2] / h3[:, None] - zs
            e = np.sqrt((r * r).sum(axis=1))
            w = np.where(e <= oc.huber_epsilon, 1.0, oc.huber_epsilon / (2 * np.maximum(e, 1e-300)))
            w2 = np.where(w == 1.0, 1.0, w * w)
            A = np.einsum('k,kij,kil->jl', w2, J, J)
            bb = np.einsum('k,kij,ki->j', w2, J, r)
            reduced = False
            while inner < oc.inner_loop_max_iteration and not reduced:
                delta = np.linalg.solve(A + lambd * _I3, bb)
                new_sol = sol - delta
                delta_norm = np.sqrt(delta @ delta)
                new_cost = cost(new_sol)
                if new_cost < total:
                    reduced = True
                    sol, total = new_sol, new_cost
                    lambd = max(lambd / 10.0, 1e-10)
                else:
                    lambd = min(lambd * 10.0, 1e12)
                inner += 1
            outer += 1
        return sol

    # -- measurement model ------------------------------------------------------------------------------------------------
    def _jacobian_blocks(self, R0, p_cam, R_null, p_null, p_w, Z):
        """numpy statement of _msckfhost.jacobians (msckf.py:443-502): H_x (F, m, 4, 6) with the observability
        constraint applied, H_f (F, 4m, 3), r (F, 4m)."""
        F, m = Z.shape[:2]
        R01, t01, g = self.R_cam0_cam1, self.t_cam0_cam1, self.gravity
        d = p_w[:, None, :] - p_cam
        pc0 = np.einsum('fmij,fmj->fmi', R0, d)
        # t_c1_w = t_c0_w - R_w_c1^T t_cam0_cam1  ->  p_c1 = R_cam0_cam1 p_c0 + t_cam0_cam1
        pc1 = pc0 @ R01.T + t01

        def dproj(pc):                                                              # (F, m, 2, 3): d(x/z, y/z)/dp
            iz = 1.0 / pc[..., 2]
            J = np.zeros((F, m, 2, 3))
            J[..., 0, 0] = iz
            J[..., 1, 1] = iz
            J[..., 0, 2] = -pc[..., 0] * iz * iz
            J[..., 1, 2] = -pc[..., 1] * iz * iz
            return J

        J0, J1 = dproj(pc0), dproj(pc1)
        sk = np.zeros((F, m, 3, 3))                                                 # skew(p_c0)
        sk[..., 0, 1], sk[..., 0, 2] = -pc0[..., 2], pc0[..., 1]
        sk[..., 1, 0], sk[..., 1, 2] = pc0[..., 2], -pc0[..., 0]
        sk[..., 2, 0], sk[..., 2, 1] = -pc0[..., 1], pc0[..., 0]
        Hx = np.empty((F, m, 4, 6))
        Hx[..., :2, :3] = J0 @ sk
        Hx[..., :2, 3:] = -(J0 @ R0)
        Hx[..., 2:, :3] = J1 @ (R01 @ sk)
        Hx[..., 2:, 3:] = -(J1 @ (R01 @ R0))
        # observability constraint (msckf.py:496-502)
        u = np.empty((F, m, 6))
        u[..., :3] = R_null @ g
        dn = p_w[:, None, :] - p_null
        u[..., 3] = dn[..., 1] * g[2] - dn[..., 2] * g[1]
        u[..., 4] = dn[..., 2] * g[0] - dn[..., 0] * g[2]
        u[..., 5] = dn[..., 0] * g[1] - dn[..., 1] * g[0]
        Au = np.einsum('fmij,fmj->fmi', Hx, u)
        Hx = Hx - Au[..., None] * (u / (u * u).sum(axis=-1, keepdims=True))[..., None, :]
        Hf = -Hx[..., 3:6].reshape(F, 4 * m, 3)
        r = np.empty((F, m, 4))
        r[..., :2] = Z[..., :2] - pc0[..., :2] / pc0[..., 2:3]
        r[..., 2:] = Z[..., 2:] - pc1[..., :2] / pc1[..., 2:3]
        r = r.reshape(F, 4 * m)
        return Hx, Hf, r

    def _jacobians(self, feats, cam_ids):
        """Measurement Jacobians of F features that are each observed in m camera states (cam_ids[f] lists them), projected
        onto the left null space of the feature Jacobian (msckf.py:443-540), all features at once.
        Returns H (F, 4m - 3, 6m), r (F, 4m - 3), slots (F, m): column block k of H[f] belongs to window slot slots[f, k]."""
        cams = self.cams
        F, m = len(feats), len(cam_ids[0])
        index = cams.index
        slots = np.fromiter((index[c] for ids in cam_ids for c in ids), dtype=np.int64, count=F * m).reshape(F, m)
        Z = np.fromiter(chain.from_iterable(f.observations[c] for f, ids in zip(feats, cam_ids) for c in ids),
                        dtype=np.float64, count=F * m * 4).reshape(F, m, 4)
        p_w = np.array([f.position for f in feats])                                 # (F, 3)
        R01, t01, g = self.R_cam0_cam1, self.t_cam0_cam1, self.gravity
        R0 = cams.R[slots]                                                          # (F, m, 3, 3) world -> cam0
        if self.use_c:
            Hx, Hf, r = np.empty((F, m, 4, 6)), np.empty((F, 4 * m, 3)), np.empty((F, 4 * m))
            _C.jacobians(R0, cams.p[slots], cams.R_null[slots], cams.p_null[slots], np.ascontiguousarray(p_w),
                         np.ascontiguousarray(Z), np.ascontiguousarray(R01), np.ascontiguousarray(t01),
                         np.ascontiguousarray(g), Hx, Hf, r)
        else:
            Hx, Hf, r = self._jacobian_blocks(R0, cams.p[slots], cams.R_null[slots], cams.p_null[slots], p_w, Z)
        # left null space of H_f: the last 4m - 3 columns of its complete QR
        if self.use_c:
            # gate statistic straight from H_x, H_f, r (one 4m x 4m Cholesky per feature, see _msckfhost.gate), then the
            # projection for the stacked Jacobian
            self._gamma = np.empty(F)
            _C.gate(Hx, Hf, r, self.state_cov, np.ascontiguousarray(slots, dtype=np.int64), float(self.config.observation_noise),
                    self._gamma)
            H, rp = np.empty((F, 4 * m - 3, 6 * m)), np.empty((F, 4 * m - 3))
            _C.null_project(Hx, Hf, r, H, rp)
            bad = np.flatnonzero(np.isnan(self._gamma))
            if len(bad):                # innovation covariance not positive definite: the reference's LU path decides
                self._gamma[bad] = self._gamma_lu(H[bad], rp[bad], slots[bad])
            return H, rp, slots
        self._gamma = None
        Q, _ = np.linalg.qr(Hf, mode='complete')
        At = Q[:, :, 3:].transpose(0, 2, 1)                                         # (F, 4m - 3, 4m)
        H = np.einsum('fakr,fkrc->fakc', At.reshape(F, 4 * m - 3, m, 4), Hx).reshape(F, 4 * m - 3, 6 * m)
        return H, np.einsum('fab,fb->fa', At, r), slots

    def _gates(self, H, r, slots, dof):
        """msckf.py:605-612 for a batch.  Returns the boolean decisions."""
        return self._gamma_lu(H, r, slots) < self.chi_squared_test_table[dof]

    def _gamma_lu(self, H, r, slots):
        """The reference's gate statistic r^T (H P H^T + sigma I)^-1 r through an LU solve (msckf.py:605-612; no
        definiteness assumed, a singular S raises as it does there), H by its non-zero column blocks."""
        if (slots == slots[0]).all():                                               # the usual case: one set of states
            c0 = ((21 + 6 * slots[0])[:, None] + np.arange(6)).reshape(-1)
            P = self.state_cov[np.ix_(c0, c0)]                                      # (6m, 6m), broadcast over the features
        else:
            cols = (21 + 6 * slots)[:, :, None] + np.arange(6)                      # (F, m, 6)
            cols = cols.reshape(len(H), -1)
            P = self.state_cov[cols[:, :, None], cols[:, None, :]]                  # (F, 6m, 6m)
        S = H @ P @ H.transpose(0, 2, 1)
        S[:, np.arange(S.shape[1]), np.arange(S.shape[1])] += self.config.observation_noise
        return np.einsum('fa,fa->f', r, np.linalg.solve(S, r[:, :, None])[:, :, 0])

    def _evaluate(self, feats, cam_ids, dof_offset, max_rows=None):
        """Jacobians + gate for a list of features, grouped by their number of camera states so that every group is one
        vectorised pass.  Returns the accepted features as blocks (H (n, a, 6m), r (n, a), slots (n, m)), one per group.
        `max_rows`: the reference stops stacking once more than this many rows are in (msckf.py:665-667): features are
        taken in the order given up to and including the one that crosses the limit."""
        groups = {}
        for i, ids in enumerate(cam_ids):
            groups.setdefault(len(ids), []).append(i)
        res = []
        ok_all = np.zeros(len(feats), dtype=bool)
        rows_all = np.zeros(len(feats), dtype=np.int64)
        for m, idx in groups.items():
            H, r, slots = self._jacobians([feats[i] for i in idx], [cam_ids[i] for i in idx])
            if self._gamma is not None:                                 # msckf.py:605-612, statistic computed in C
                ok = self._gamma < self.chi_squared_test_table[m + dof_offset]
            else:
                ok = self._gates(H, r, slots, m + dof_offset)
            idx = np.asarray(idx)
            ok_all[idx] = ok
            rows_all[idx] = H.shape[1]
            res.append((idx, H, r, slots, ok))
        last = len(feats)
        if max_rows is not None:
            over = np.flatnonzero(np.cumsum(rows_all * ok_all) > max_rows)
            if len(over):
                last = int(over[0]) + 1
        blocks = []
        for idx, H, r, slots, ok in res:
            keep = ok & (idx < last)
            if keep.any():
                blocks.append((H[keep], r[keep], slots[keep]))
        return blocks

    def measurement_update(self, Hc, r, cols):
        """msckf.py:542-603.  The stacked Jacobian arrives compact: `Hc` holds only the columns `cols` (sorted state
        indices) of H -- H is zero everywhere else (the 21 IMU columns always; in the pruning update everything but the
        two camera states that leave: 12 of up to 120 columns)."""
        if len(Hc) == 0 or len(r) == 0:
            return
        # K r and (I - K H) P depend on H only through G = H^T H and b = H^T r (push-through identity:
        # H^T (H P H^T + s I)^-1 = (G P + s I)^-1 H^T), so neither the reference's thin QR of the stacked Jacobian
        # (msckf.py:548-554; 4 ms for 1500 x 120) nor the row-sized S is formed: with E selecting the non-zero columns,
        #   K r = P E (G P_cc + s I)^-1 b,      K H P = P E (G P_cc + s I)^-1 G E^T P
        # -- one nc x nc solve with 1 + n right-hand sides, whatever the number of rows.
        P = self.state_cov
        G = Hc.T @ Hc
        Pc = P[cols]                                                  # (nc, n) = E^T P
        M = G @ Pc[:, cols]
        M[np.diag_indices(len(M))] += self.config.observation_noise
        X = np.linalg.solve(M, np.column_stack([Hc.T @ r, G @ Pc]))   # (nc, 1 + n)
        delta = Pc.T @ X[:, 0]
        KHP = Pc.T @ X[:, 1:]

        st = self.imu_state
        d_imu = delta[:21]
        if np.linalg.norm(d_imu[6:9]) > 0.5 or np.linalg.norm(d_imu[12:15]) > 1.0:
            self.large_update_count += 1
        # Reference quirk (kept: it shapes the transition matrices of the next IMU batch).  The reference corrects velocity
        # and position IN PLACE (msckf.py:579-581) and its *_null attributes are aliases of those arrays since the last
        # propagation step (msckf.py:336-338), so velocity_null / position_null move with the correction, whereas
        # orientation is re-assigned and orientation_null keeps the propagated value.  Same for the camera states below:
        # position_null is an alias of position from the augmentation on (msckf.py:404-405, 590-591).
        st.orientation = quaternion_multiplication(small_angle_quaternion(d_imu[:3]), st.orientation)
        st.gyro_bias = st.gyro_bias + d_imu[3:6]
        st.acc_bias = st.acc_bias + d_imu[9:12]
        aliased_v, aliased_p = st.velocity_null is st.velocity, st.position_null is st.position
        st.velocity = st.velocity + d_imu[6:9]
        st.position = st.position + d_imu[12:15]
        if aliased_v:
            st.velocity_null = st.velocity
        if aliased_p:
            st.position_null = st.position
        st.R_imu_cam0 = to_rotation(small_angle_quaternion(d_imu[15:18])) @ st.R_imu_cam0
        st.t_cam0_imu = st.t_cam0_imu + d_imu[18:21]
        cams = self.cams
        n = len(cams)
        dc = delta[21:].reshape(-1, 6)
        if self.use_c:                  # the statements below, one loop in C
            _C.update_cams(cams.q, cams.p, cams.R, cams.p_null, np.ascontiguousarray(dc))
            Pn = P - KHP                                              # (I - K H) P
            self.state_cov = (Pn + Pn.T) / 2.0
            return
        # every camera state at once: q <- dq(dtheta) * q  (small_angle_quaternion + quaternion_multiplication, vectorised)
        h = dc[:, :3] / 2.0
        n2 = (h * h).sum(axis=1)
        small = n2 <= 1
        dq = np.empty((n, 4))
        dq[:, :3] = h
        dq[:, 3] = np.where(small, np.sqrt(np.where(small, 1 - n2, 0.0)), 1.0)
        dq[~small] /= np.sqrt(1 + n2[~small])[:, None]
        dq /= np.sqrt((dq * dq).sum(axis=1))[:, None]
        q2 = cams.q[:n] / np.sqrt((cams.q[:n] * cams.q[:n]).sum(axis=1))[:, None]
        x, y, z, w = dq[:, 0], dq[:, 1], dq[:, 2], dq[:, 3]
        qn = np.stack([w * q2[:, 0] + z * q2[:, 1] - y * q2[:, 2] + x * q2[:, 3],
                       -z * q2[:, 0] + w * q2[:, 1] + x * q2[:, 2] + y * q2[:, 3],
                       y * q2[:, 0] - x * q2[:, 1] + w * q2[:, 2] + z * q2[:, 3],
                       -x * q2[:, 0] - y * q2[:, 1] - z * q2[:, 2] + w * q2[:, 3]], axis=1)
        qn /= np.sqrt((qn * qn).sum(axis=1))[:, None]
        cams.q[:n] = qn
        cams.R[:n] = _rotations(qn)
        cams.p[:n] += dc[:, 3:]
        cams.p_null[:n] = cams.p[:n]
        Pn = P - KHP                                                  # (I - K H) P
        self.state_cov = (Pn + Pn.T) / 2.0

    def _stack(self, blocks):
        """Stacked Jacobian / residual from blocks (H (n, a, 6m), r (n, a), slots (n, m)), compact in the columns: only the
        camera states that occur in some block get columns (one scatter per block).  Returns Hc, r and the state indices
        `cols` of Hc's columns."""
        rows = sum(b[1].size for b in blocks)
        used = np.unique(np.concatenate([b[2].reshape(-1) for b in blocks])) if blocks else np.zeros(0, np.int64)
        pos = np.full(int(used.max()) + 1 if len(used) else 0, -1, dtype=np.int64)
        pos[used] = np.arange(len(used))
        H = np.zeros((rows, 6 * len(used)))
        r = np.empty(rows)
        at = 0
        for Hs, rs, sl in blocks:
            n, a, c = Hs.shape
            if len(used) * 6 == c and n and (sl == sl[0]).all() and (np.diff(sl[0]) > 0).all():
                H[at:at + n * a] = Hs.reshape(n * a, c)                             # every feature on the same states,
            else:                                                                   # in column order: already compact
                ri = (at + a * np.arange(n))[:, None] + np.arange(a)                # (n, a)
                ci = ((6 * pos[sl])[:, :, None] + np.arange(6)).reshape(n, c)       # (n, 6m)
                H[ri[:, :, None], ci[:, None, :]] = Hs
            r[at:at + n * a] = rs.reshape(-1)
            at += n * a
        cols = (21 + 6 * used[:, None] + np.arange(6)).reshape(-1)
        return H, r, cols

    def remove_lost_features(self):
        """msckf.py:614-676: features that lost tracking are used for an update and leave the map."""
        cur = self.imu_state.id
        invalid, processed = [], []
        for feat in self.map_server.values():
            if cur in feat.observations:
                continue
            if len(feat.observations) < 3:
                invalid.append(feat.id)
                continue
            if not feat.is_initialized:
                if not self.check_motion(feat) or not self.initialize_position(feat):
                    invalid.append(feat.id)
                    continue
            processed.append(feat)
        for fid in invalid:
            del self.map_server[fid]
        if not processed:
            return
        # gate dof = observations - 1 (msckf.py:661); stacking stops after the feature that takes it past 1500 rows
        blocks = self._evaluate(processed, [list(f.observations) for f in processed], -1, max_rows=1500)
        H, r, cols = self._stack(blocks)
        self.measurement_update(H, r, cols)
        for feat in processed:
            del self.map_server[feat.id]

    # -- window management ------------------------------------------------------------------------------------------------
    def find_redundant_cam_states(self):
        """msckf.py:678-709: two states leave: recent ones that barely moved relative to the 4th newest, else the oldest."""
        cams = self.cams
        n = len(cams)
        key = n - 4
        idx, first = key + 1, 0
        key_p, key_R = cams.p[key], cams.R[key]
        rm = []
        for _ in range(2):
            distance = np.linalg.norm(cams.p[idx] - key_p)
            angle = 2 * np.arccos(to_quaternion(cams.R[idx] @ key_R.T)[-1])
            if angle < 0.2618 and distance < 0.4 and self.tracking_rate > 0.5:
                rm.append(cams.ids[idx])
            else:
                rm.append(cams.ids[first])
                first += 1
            idx += 1
        return sorted(rm)

    def prune_cam_state_buffer(self):
        """msckf.py:712-786."""
        if len(self.cams) < self.config.max_cam_state_size:
            return
        rm = self.find_redundant_cam_states()
        todo = []
        for feat in self.map_server.values():
            obs = feat.observations
            involved = [c for c in rm if c in obs]
            if not involved:
                continue
            if len(involved) == 1:
                del obs[involved[0]]
                continue
            if not feat.is_initialized:
                if not self.check_motion(feat) or not self.initialize_position(feat):
                    for c in involved:
                        del obs[c]
                    continue
            todo.append((feat, involved))
        # gate dof = involved states (msckf.py:763)
        blocks = self._evaluate([t[0] for t in todo], [t[1] for t in todo], 0) if todo else []
        for feat, involved in todo:
            for c in involved:
                del feat.observations[c]
        H, r, cols = self._stack(blocks)
        self.measurement_update(H, r, cols)
        for cam_id in rm:
            i = self.cams.remove(cam_id)
            keep = np.r_[0:21 + 6 * i, 27 + 6 * i:self.state_cov.shape[0]]
            self.state_cov = self.state_cov[np.ix_(keep, keep)]

    def reset_state_cov(self):
        """msckf.py:788-798."""
        c = self.config
        P = np.zeros((21, 21))
        P[3:6, 3:6] = c.gyro_bias_cov * _I3
        P[6:9, 6:9] = c.velocity_cov * _I3
        P[9:12, 9:12] = c.acc_bias_cov * _I3
        P[15:18, 15:18] = c.extrinsic_rotation_cov * _I3
        P[18:21, 18:21] = c.extrinsic_translation_cov * _I3
        self.state_cov = P

    def reset(self):
        """msckf.py:800-820."""
        old = self.imu_state
        self.imu_state = IMUState()
        self.imu_state.id = old.id
        self.imu_state.R_imu_cam0, self.imu_state.t_cam0_imu = old.R_imu_cam0, old.t_cam0_imu
        self.cams.clear()
        self.reset_state_cov()
        self.map_server.clear()
        self.imu_msg_buffer.clear()
        self.is_gravity_set = False
        self.is_first_img = True

    def online_reset(self):
        """msckf.py:822-843: drop the window and the map when the position uncertainty explodes."""
        thr = self.config.position_std_threshold
        if thr <= 0:
            return
        P = self.state_cov
        if max(np.sqrt(P[12, 12]), np.sqrt(P[13, 13]), np.sqrt(P[14, 14])) < thr:
            return
        self.cams.clear()
        self.map_server.clear()
        self.reset_state_cov()

    def publish(self, time):
        """msckf.py:845-867."""
        st = self.imu_state
        T_i_w = Isometry3d(to_rotation(st.orientation).T, st.position)
        T_b_w = self.T_imu_body * T_i_w * self.T_imu_body.inverse()
        body_velocity = self.T_imu_body.R @ st.velocity
        R_w_c = st.R_imu_cam0 @ T_i_w.R.T
        t_c_w = st.position + T_i_w.R @ st.t_cam0_imu
        if self._outfile:
            q, p = st.orientation, st.position
            with open(self._outfile, 'a') as f:
                f.write(f'{st.timestamp:.6f} {p[0]:.9f} {p[1]:.9f} {p[2]:.9f} {q[0]:.9f} {q[1]:.9f} {q[2]:.9f} {q[3]:.9f}\n')
        return vio_result(time, T_b_w, body_velocity, Isometry3d(R_w_c.T, t_c_w))
