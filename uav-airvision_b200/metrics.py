"""Trajectory metrics for the sequence-level parity gate.

The reference ships results (results/metrics_summary.csv: ate_rmse_m, ate_mean_m, ate_std_m, rte_rmse_m, rte_mean_m,
rte_std_m, ate_perc) but not the scripts that made them (.gitignore:17-21).  These are the standard definitions:
ATE = translation error after a least-squares rigid (SE(3), no scale) alignment of the estimate to ground truth;
RTE = error of the relative translation over consecutive poses; ate_perc = ATE RMSE as % of the path length."""
from __future__ import annotations

import numpy as np


def load_trajectory_txt(path):
    """Lines `t x y z qx qy qz qw` as MSCKF._write_state emits them (reference msckf.py:152-160)."""
    a = np.loadtxt(path, ndmin=2)
    return a[:, 0], a[:, 1:4], a[:, 4:8]


def associate(t_est, t_gt, max_dt=0.01):
    """Index pairs (i_est, i_gt) of nearest timestamps within max_dt."""
    t_gt = np.asarray(t_gt)
    j = np.clip(np.searchsorted(t_gt, t_est), 1, len(t_gt) - 1)
    j = np.where(np.abs(t_gt[j - 1] - t_est) <= np.abs(t_gt[j] - t_est), j - 1, j)
    ok = np.abs(t_gt[j] - t_est) <= max_dt
    return np.nonzero(ok)[0], j[ok]


def align_rigid(est, gt):
    """R, t minimising sum |R est_i + t - gt_i|^2 (Kabsch / Umeyama without scale)."""
    mu_e, mu_g = est.mean(0), gt.mean(0)
    H = (est - mu_e).T @ (gt - mu_g)
    U, _, Vt = np.linalg.svd(H)
    D = np.diag([1.0, 1.0, np.sign(np.linalg.det(Vt.T @ U.T))])
    R = Vt.T @ D @ U.T
    return R, mu_g - R @ mu_e


def trajectory_metrics(t_est, p_est, t_gt, p_gt, max_dt=0.01):
    ie, ig = associate(np.asarray(t_est), np.asarray(t_gt), max_dt)
    if len(ie) < 3:
        raise ValueError('fewer than 3 associated poses')
    e, g = np.asarray(p_est)[ie], np.asarray(p_gt)[ig]
    R, t = align_rigid(e, g)
    ea = e @ R.T + t
    err = np.linalg.norm(ea - g, axis=1)
    de, dg = np.diff(ea, axis=0), np.diff(g, axis=0)
    rerr = np.linalg.norm(de - dg, axis=1)
    path = float(np.linalg.norm(dg, axis=1).sum())
    rms = lambda x: float(np.sqrt(np.mean(x * x)))
    return {'n': int(len(ie)), 'ate_rmse_m': rms(err), 'ate_mean_m': float(err.mean()), 'ate_std_m': float(err.std()),
            'rte_rmse_m': rms(rerr), 'rte_mean_m': float(rerr.mean()), 'rte_std_m': float(rerr.std()),
            'path_m': path, 'ate_perc': 100.0 * rms(err) / path if path > 0 else float('nan')}
