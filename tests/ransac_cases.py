"""Synthetic two-frame point sets for the two-point RANSAC tests: a rigid 3-D scene seen before and after a small
camera motion (rotation known from the "gyro", translation unknown), plus gross outliers."""
import numpy as np

from oracle import cv_semantics as cs

K_EUROC = np.array([458.654, 457.296, 367.215, 248.375])
D_EUROC = np.array([-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05])
K_EUROC1 = np.array([457.587, 456.134, 379.999, 255.238])
D_EUROC1 = np.array([-0.28368365, 0.07451284, -0.00010473, -3.55590700e-05])


def rotation(w):
    w = np.asarray(w, dtype=np.float64)
    th = np.linalg.norm(w)
    if th == 0:
        return np.eye(3)
    k = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]]) / th
    return np.eye(3) + np.sin(th) * k + (1 - np.cos(th)) * k @ k


def make_pixels(n, outliers, seed, translation=(0.08, -0.03, 0.02), gyro=(0.004, -0.006, 0.01), jitter_px=0.15,
                outlier_px=(8.0, 30.0), K=K_EUROC, D=D_EUROC, size=(752, 480)):
    """-> prev_px, cur_px (N, 2) float32, R_p_c (3, 3), truth (N,) bool (True = consistent with the motion)."""
    rng = np.random.default_rng(seed)
    R = rotation(gyro)                                      # previous camera frame -> current camera frame
    t = np.asarray(translation, dtype=np.float64)
    uv = np.stack([rng.uniform(20, size[0] - 20, n), rng.uniform(20, size[1] - 20, n)], axis=1)
    xn = cs.undistort_radtan(uv, K, D)                      # normalized coordinates in the previous frame
    depth = rng.uniform(2.0, 12.0, n)
    P = np.concatenate([xn, np.ones((n, 1))], axis=1) * depth[:, None]
    Q = P @ R.T + t
    cur = cs.distort_radtan(Q[:, :2] / Q[:, 2:3], K, D) + rng.normal(0, jitter_px, (n, 2))
    truth = np.ones(n, bool)
    if outliers:
        bad = rng.choice(n, outliers, replace=False)
        ang = rng.uniform(0, 2 * np.pi, outliers)
        mag = rng.uniform(outlier_px[0], outlier_px[1], outliers)
        cur[bad] += np.stack([np.cos(ang), np.sin(ang)], axis=1) * mag[:, None]
        truth[bad] = False
    return uv.astype(np.float32), cur.astype(np.float32), R, truth


def undistorted(prev_px, cur_px, R, K=K_EUROC, D=D_EUROC):
    """The inputs oracle.ransac.two_point_ransac expects (what the port computes with its backend)."""
    u1 = cs.undistort_radtan(np.asarray(prev_px, np.float32), K, D, R)
    u2 = cs.undistort_radtan(np.asarray(cur_px, np.float32), K, D)
    return u1, u2


def make_case(n, outliers, seed, **kw):
    a, b, R, truth = make_pixels(n, outliers, seed, **kw)
    u1, u2 = undistorted(a, b, R, kw.get('K', K_EUROC), kw.get('D', D_EUROC))
    return u1, u2, truth
