"""Downstream MSCKF ATE parity (BASELINE config C2): same rendered EuRoC-format sequence through
  (1) the UNMODIFIED reference front end + the UNMODIFIED reference MSCKF, and
  (2) the B200 front end (feature_msg stream dumped on the GPU box) + the same reference MSCKF.
The MSCKF is the reference's own code (out of build scope); it lives under /root/reference and
cannot travel to the GPU box, so this runs in two steps:

    # on the GPU box: dump what the CUDA front end publishes
    python tools/ate_parity.py dump --frames 400 --out gpurun_out/ate_gpu_features.npz
    # in the build container (reference importable): run both estimators, write profiles/ate_parity.json
    python tools/ate_parity.py compare --features tests/golden/ate_gpu_features.npz

`compare` also reports the feature-level agreement of the two front ends on every frame."""
from __future__ import annotations

import argparse
import contextlib
import io
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.dont_write_bytecode = True
for p in (ROOT, os.path.join(ROOT, 'uav-airvision_b200')):
    if p not in sys.path:
        sys.path.insert(0, p)

SEQ = dict(seed=3, rate=20.0, imu_rate=200.0)
GRID = dict(grid_row=6, grid_col=10, grid_min=3, grid_max=5)        # C2: 300 features


def make_stream(n_frames):
    from frontend_config import FrontEndConfig
    from synth_euroc import RoomSceneStream
    return RoomSceneStream(FrontEndConfig(**GRID), n_frames=n_frames, **SEQ)


def dump(args):
    from image_processing import ImageProcessor
    from frontend_config import FrontEndConfig
    from replay import run_stream
    cfg = FrontEndConfig(**GRID)
    ip = ImageProcessor(cfg)
    rec = {}

    def on_frame(k, msg, fm):
        rec[f'f{k}_ids'] = np.array([f.id for f in fm.features], np.int64)
        rec[f'f{k}_meas'] = np.array([[f.u0, f.v0, f.u1, f.v1] for f in fm.features], np.float64).reshape(-1, 4)
        rec[f'f{k}_ts'] = np.array([fm.timestamp])

    run_stream(ip, make_stream(args.frames), on_frame=on_frame)
    rec['n_frames'] = np.array([args.frames])
    np.savez_compressed(args.out, **rec)
    print('dumped', args.frames, 'frames ->', args.out, os.path.getsize(args.out), 'B; features in the last frame:',
          len(rec[f'f{args.frames - 1}_ids']))


class _Replayed:
    """Front end that replays dumped feature messages (ids + normalized stereo measurements)."""

    def __init__(self, npz):
        from collections import namedtuple
        self.z, self.k = npz, 0
        self.msg = namedtuple('feature_msg', ['timestamp', 'features'])
        self.F = namedtuple('FeatureMeasurement', ['id', 'u0', 'v0', 'u1', 'v1'])

    def imu_callback(self, m):
        pass

    def stereo_callback(self, m):
        k = self.k
        self.k += 1
        ids, meas = self.z[f'f{k}_ids'], self.z[f'f{k}_meas']
        assert abs(float(self.z[f'f{k}_ts'][0]) - m.cam0_msg.timestamp) < 1e-9
        return self.msg(m.cam0_msg.timestamp, [self.F(int(i), *row) for i, row in zip(ids, meas.tolist())])


def run_vio(front_end, stream, tag, workdir):
    """Deterministic driver: reference MSCKF consuming `front_end`'s messages.  Returns (t, positions) of the filter."""
    from msckf import MSCKF                                   # reference, unmodified
    from config import ConfigEuRoC
    from replay import run_stream
    os.environ['DATASET_NAME'], os.environ['TIME_OFFSET'] = tag, '0'
    cwd = os.getcwd()
    os.chdir(workdir)
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            est = MSCKF(ConfigEuRoC())
            poses, msgs = [], []

            class Tap:
                def imu_callback(self, m):
                    est.imu_callback(m)

                def feature_callback(self, fm):
                    r = est.feature_callback(fm)
                    if r is not None:
                        poses.append((r.timestamp, np.array(r.pose.t).copy()))

            msgs = run_stream(front_end, stream, msckf=Tap())
    finally:
        os.chdir(cwd)
    t = np.array([p[0] for p in poses])
    P = np.array([p[1] for p in poses]).reshape(-1, 3)
    return t, P, msgs


def compare(args):
    sys.path.insert(0, '/root/reference/src')
    from config import ConfigEuRoC
    from image_processing import ImageProcessor as RefImageProcessor      # resolves to the REFERENCE package here
    from metrics import trajectory_metrics
    z = np.load(args.features)
    n = int(z['n_frames'][0])
    gt = list(make_stream(n).groundtruth())
    t_gt, p_gt = np.array([g.timestamp for g in gt]), np.array([g.p for g in gt])
    cfg = ConfigEuRoC()
    cfg.grid_row, cfg.grid_col = GRID['grid_row'], GRID['grid_col']
    cfg.grid_num = cfg.grid_row * cfg.grid_col
    cfg.grid_min_feature_num, cfg.grid_max_feature_num = GRID['grid_min'], GRID['grid_max']
    work = tempfile.mkdtemp(prefix='ate_')
    t_r, P_r, ref_msgs = run_vio(RefImageProcessor(cfg), make_stream(n), 'ref_front_end', work)
    t_g, P_g, _ = run_vio(_Replayed(z), make_stream(n), 'b200_front_end', work)
    m_ref = trajectory_metrics(t_r, P_r, t_gt, p_gt)
    m_gpu = trajectory_metrics(t_g, P_g, t_gt, p_gt)
    # feature-level agreement of the two front ends
    same_ids, worst = 0, 0.0
    for k, fm in enumerate(ref_msgs):
        ids = np.array([f.id for f in fm.features], np.int64)
        if np.array_equal(ids, z[f'f{k}_ids']):
            same_ids += 1
            if len(ids):
                meas = np.array([[f.u0, f.v0, f.u1, f.v1] for f in fm.features], np.float64)
                worst = max(worst, float(np.abs(meas - z[f'f{k}_meas']).max()))
    rel = abs(m_gpu['ate_rmse_m'] - m_ref['ate_rmse_m']) / m_ref['ate_rmse_m']
    out = {'sequence': dict(SEQ, frames=n, kind='RoomSceneStream 752x480, C2 grid 6x10 x max 5'),
           'reference_front_end': m_ref, 'b200_front_end': m_gpu, 'ate_rmse_relative_difference': rel,
           'max_position_difference_between_trajectories_m': float(np.abs(P_r - P_g).max()) if P_r.shape == P_g.shape else None,
           'frames_with_identical_feature_ids': same_ids, 'frames': n,
           'worst_normalized_coordinate_difference_on_those_frames': worst,
           'gate': 'ATE within 1 % of the reference (BASELINE.json north_star)', 'pass': bool(rel <= 0.01)}
    print(json.dumps(out, indent=1))
    if args.out:
        with open(args.out, 'w') as f:
            json.dump(out, f, indent=1)
    return out


def live(args):
    """Both front ends LIVE on this box (needs oracle/_ref and a GPU): the reference's thread harness modules/vio.py with
    the reference's MSCKF, once over the reference's own image_processing and once over this repo's CUDA package, on the
    same rendered sequence (>= 60 s for the BASELINE gate); ATE of both against the rendered ground truth."""
    import subprocess
    from metrics import trajectory_metrics
    n = args.frames
    gt = list(make_stream(n).groundtruth())
    t_gt, p_gt = np.array([g.timestamp for g in gt]), np.array([g.p for g in gt])
    runner = os.path.join(ROOT, 'oracle', 'ref_runner.py')
    res = {}
    for fe in ('ref', 'b200'):
        dump = os.path.join(tempfile.mkdtemp(prefix='ate_live_'), fe + '.npz')
        r = subprocess.run([sys.executable, runner, 'vio', '--frames', str(n), '--front-end', fe, '--dump', dump,
                            '--render-procs', str(args.render_procs), '--no-traj'], capture_output=True, text=True)
        if r.returncode != 0:
            raise SystemExit(f'{fe}: {r.stderr[-2000:]}')
        info = json.loads(r.stdout.strip().splitlines()[-1])
        z = np.load(dump)
        traj = z['traj']
        res[fe] = dict(info=info, z=z, metrics=trajectory_metrics(traj[:, 1], traj[:, 2:5], t_gt, p_gt), traj=traj)
    a, b = res['ref'], res['b200']
    same_ids, worst = 0, 0.0
    for k in range(n):
        if np.array_equal(a['z'][f'f{k}_ids'], b['z'][f'f{k}_ids']):
            same_ids += 1
            if len(a['z'][f'f{k}_ids']):
                worst = max(worst, float(np.abs(a['z'][f'f{k}_meas'] - b['z'][f'f{k}_meas']).max()))
    rel = abs(b['metrics']['ate_rmse_m'] - a['metrics']['ate_rmse_m']) / a['metrics']['ate_rmse_m']
    same_shape = a['traj'].shape == b['traj'].shape
    out = {'sequence': dict(SEQ, frames=n, seconds=n / SEQ['rate'], kind='RoomSceneStream 752x480, C2 grid 6x10 x max 5'),
           'harness': 'oracle/_ref modules/vio.py (three threads, two queues) + oracle/_ref msckf.py for both front ends, live on one box',
           'reference_front_end': a['metrics'], 'b200_front_end': b['metrics'], 'ate_rmse_relative_difference': rel,
           'max_position_difference_between_trajectories_m': float(np.abs(a['traj'][:, 2:5] - b['traj'][:, 2:5]).max()) if same_shape else None,
           'frames_with_identical_feature_ids': same_ids, 'frames': n,
           'worst_normalized_coordinate_difference_on_those_frames': worst,
           'wall_s': {'reference_front_end': a['info']['wall_s'], 'b200_front_end': b['info']['wall_s']},
           'gate': 'ATE within 1 % of the reference (BASELINE.json north_star)', 'pass': bool(rel <= 0.01)}
    print(json.dumps(out, indent=1))
    if args.out:
        with open(args.out, 'w') as f:
            json.dump(out, f, indent=1)
    return out


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    sub = ap.add_subparsers(dest='cmd', required=True)
    lv = sub.add_parser('live')
    lv.add_argument('--frames', type=int, default=1200)
    lv.add_argument('--render-procs', type=int, default=os.cpu_count() or 1)
    lv.add_argument('--out', default=os.path.join(ROOT, 'gpurun_out', 'ate_parity_live.json'))
    d = sub.add_parser('dump')
    d.add_argument('--frames', type=int, default=400)
    d.add_argument('--out', default=os.path.join(ROOT, 'gpurun_out', 'ate_gpu_features.npz'))
    c = sub.add_parser('compare')
    c.add_argument('--features', default=os.path.join(ROOT, 'tests', 'golden', 'ate_gpu_features.npz'))
    c.add_argument('--out', default=os.path.join(ROOT, 'profiles', 'ate_parity.json'))
    a = ap.parse_args()
    {'dump': dump, 'compare': compare, 'live': live}[a.cmd](a)
