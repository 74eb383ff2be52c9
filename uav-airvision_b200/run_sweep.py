#!/usr/bin/env python
"""Sequence x time-offset sweep over EuRoC directories: what the reference's batch file does with one
`python main.py --path <seq> --offset <s>` per run (run.bat:4-12, src/main.py:10-34), as ONE job per GPU.

    python uav-airvision_b200/run_sweep.py --path datasets/MH_01_easy datasets/V1_01_easy --offsets 1 5 10 15 20 30 40

Every sequence is decoded once into an HBM frame store; its offset runs advance in lock-step through one libavb context
(sweep.run_sweep); every run has its own host MSCKF in a worker process.  Per run it writes the trajectory file the
reference writes (results/txts/output_<sequence>_offset<o>.txt, one 't x y z qx qy qz qw' line per published state,
msckf.py:10-16, 152-160) and one row of results/metrics_summary.csv with the columns of the reference's summary
(dataset, ate_rmse_m, ate_mean_m, ate_std_m, rte_rmse_m, rte_mean_m, rte_std_m, ate_perc) plus the offset.

Several GPUs: launch one process per GPU (torchrun or by hand with RANK / WORLD_SIZE / LOCAL_RANK set); the sequences are
dealt to the ranks (`shard_streams`), nothing is exchanged; rank r writes metrics_summary.rank<r>.csv when WORLD_SIZE > 1.
"""
from __future__ import annotations

import argparse
import csv
import os
import sys
import time


HERE = os.path.dirname(os.path.abspath(__file__))
if HERE not in sys.path:
    sys.path.insert(0, HERE)

COLUMNS = ['dataset', 'offset', 'ate_rmse_m', 'ate_mean_m', 'ate_std_m', 'rte_rmse_m', 'rte_mean_m', 'rte_std_m', 'ate_perc',
           'poses', 'frames']


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split('\n\n')[0])
    ap.add_argument('--path', nargs='+', required=True, help='EuRoC sequence directories (each holds mav0/)')
    ap.add_argument('--offsets', nargs='+', type=float, default=[1, 5, 10, 15, 20, 30, 40], help='start offsets in seconds (run.bat:5-9)')
    ap.add_argument('--max-frames', type=int, default=None, help='decode at most this many stereo frames per sequence')
    ap.add_argument('--steps', type=int, default=None, help='frames per run (default: to the end of the shortest run)')
    ap.add_argument('--workers', type=int, default=0, help='estimator processes (0 = allowed host cores)')
    ap.add_argument('--out', default='results', help='output directory (txts/ and metrics_summary.csv below it)')
    ap.add_argument('--device', type=int, default=None, help='CUDA device (default: LOCAL_RANK, else 0)')
    a = ap.parse_args(argv)

    from euroc import EuRoCDataset
    from frontend_config import FrontEndConfig, with_filter_fields
    from metrics import trajectory_metrics
    from multi_stream import shard_streams
    from sweep import CachedSequence, run_sweep

    rank, world = int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1'))
    device = a.device if a.device is not None else int(os.environ.get('LOCAL_RANK', '0'))
    mine = [a.path[i] for i in shard_streams(len(a.path), world, rank)]
    if not mine:
        return 0
    cfg = with_filter_fields(FrontEndConfig())                 # the reference's ConfigEuRoC values
    t0 = time.perf_counter()
    seqs, names = [], []
    for path in mine:
        ds = EuRoCDataset(path)
        ds.set_starttime(0)
        name = os.path.basename(os.path.normpath(path))
        seqs.append(CachedSequence(ds, device=device, max_frames=a.max_frames, name=name))
        names.append(name)
    decode_s = time.perf_counter() - t0
    workers = a.workers if a.workers > 0 else max(1, len(os.sched_getaffinity(0)))
    res = run_sweep(cfg, seqs, a.offsets, device=device, n_steps=a.steps, estimator_workers=workers)
    txts = os.path.join(a.out, 'txts')
    os.makedirs(txts, exist_ok=True)
    rows = []
    for run, traj in zip(res['runs'], res['trajectories']):
        name, seq = names[run.sequence], seqs[run.sequence]
        off = int(run.offset) if float(run.offset).is_integer() else run.offset
        with open(os.path.join(txts, f'output_{name}_offset{off}.txt'), 'w') as f:
            for t, x, y, z, qx, qy, qz, qw in traj:
                f.write(f'{t:.6f} {x:.9f} {y:.9f} {z:.9f} {qx:.9f} {qy:.9f} {qz:.9f} {qw:.9f}\n')
        row = {'dataset': name, 'offset': off, 'poses': len(traj), 'frames': res['steps']}
        if seq.groundtruth is not None and len(traj) >= 3:
            try:
                m = trajectory_metrics(traj[:, 0], traj[:, 1:4], *seq.groundtruth)
                row.update({k: m[k] for k in COLUMNS if k in m})
            except ValueError:
                pass
        rows.append(row)
    summary = os.path.join(a.out, 'metrics_summary.csv' if world == 1 else f'metrics_summary.rank{rank}.csv')
    with open(summary, 'w', newline='') as f:
        w = csv.DictWriter(f, fieldnames=COLUMNS)
        w.writeheader()
        w.writerows(rows)
    for q in seqs:
        q.close()
    print(f'rank {rank}: {len(mine)} sequence(s) x {len(a.offsets)} offsets = {res["streams"]} runs, {res["steps"]} frames each; '
          f'decode + upload {decode_s:.1f} s, sweep {res["wall_s"]:.1f} s ({res["frames_per_s"]:.0f} frames/s with {workers} '
          f'estimator processes); trajectories in {txts}, summary {summary}')
    return 0


if __name__ == '__main__':
    sys.exit(main())
