"""Golden trajectory of the UNMODIFIED reference MSCKF (/root/reference/src/msckf.py) on the committed 400-frame feature
dump of the B200 front end (tests/golden/ate_gpu_features.npz) + the IMU stream of the same rendered sequence.
Runs in the build container only (the reference is importable there); writes tests/golden/ref_msckf_traj.npz:
per published frame the timestamp, body position, IMU orientation quaternion, velocity, plus the IMU samples so that the
CPU test does not have to render the sequence.

    python tools/make_msckf_golden.py
"""
import contextlib
import io
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.dont_write_bytecode = True
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'uav-airvision_b200'))
from tools.ate_parity import make_stream          # noqa: E402


def main():
    z = np.load(os.path.join(ROOT, 'tests', 'golden', 'ate_gpu_features.npz'))
    n = int(z['n_frames'][0])
    stream = make_stream(n)
    imu = [(m.timestamp, *m.angular_velocity, *m.linear_acceleration) for m in stream.imu()]
    imu = np.array(imu, np.float64)
    sys.path.insert(0, '/root/reference/src')
    from config import ConfigEuRoC
    from msckf import MSCKF
    assert MSCKF.__module__ == 'msckf' and '/root/reference' in sys.modules['msckf'].__file__
    from collections import namedtuple
    imu_msg = namedtuple('imu_msg', ['timestamp', 'angular_velocity', 'linear_acceleration'])
    fmsg = namedtuple('feature_msg', ['timestamp', 'features'])
    F = namedtuple('FeatureMeasurement', ['id', 'u0', 'v0', 'u1', 'v1'])
    os.environ['DATASET_NAME'], os.environ['TIME_OFFSET'] = 'golden', '0'
    cwd = os.getcwd()
    os.chdir(tempfile.mkdtemp(prefix='msckf_golden_'))
    rows, per_frame = [], []
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            est = MSCKF(ConfigEuRoC())
            j = 0
            for k in range(n):
                ts = float(z[f'f{k}_ts'][0])
                while j < len(imu) and imu[j, 0] <= ts:
                    est.imu_callback(imu_msg(imu[j, 0], imu[j, 1:4].copy(), imu[j, 4:7].copy()))
                    j += 1
                feats = [F(int(i), *row) for i, row in zip(z[f'f{k}_ids'], z[f'f{k}_meas'].tolist())]
                t0 = time.perf_counter()
                r = est.feature_callback(fmsg(ts, feats))
                per_frame.append(time.perf_counter() - t0)
                if r is not None:
                    st = est.state_server.imu_state
                    rows.append([k, r.timestamp, *r.pose.t, *st.orientation, *st.velocity, len(est.state_server.cam_states),
                                 len(est.map_server)])
    finally:
        os.chdir(cwd)
    rows = np.array(rows, np.float64)
    out = os.path.join(ROOT, 'tests', 'golden', 'ref_msckf_traj.npz')
    np.savez_compressed(out, traj=rows, imu=imu, ref_ms_per_frame_median=np.array([1e3 * np.median(per_frame[20:])]))
    print('reference MSCKF:', len(rows), 'published frames, median', 1e3 * np.median(per_frame[20:]), 'ms/frame ->', out,
          os.path.getsize(out), 'B')


if __name__ == '__main__':
    main()
