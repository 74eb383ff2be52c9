"""ORACLE (test infrastructure, never shipped, never on the product path).

Two-point RANSAC outlier rejection between the previous and the current frame of one camera.

PARITY UNPINNED BY THE REFERENCE: uav-airvision removed this stage -- the tracker's inlier masks are hard-coded
all-ones (/root/reference/src/image_processing/feature_tracker.py:135-136), `ransac_threshold` is configured
(/root/reference/src/config.py:29) and plumbed (feature_tracker.py:63) but never read, and only the `rescale_points`
helper survives (/root/reference/src/image_processing/camera_model.py:95-108).  BASELINE.json's north star still
names the stage, so it is restated here from the published algorithm the reference descends from: `twoPointRansac`
of the stereo MSCKF-VIO image processor (Sun et al., "Robust Stereo Visual Inertial Odometry for Fast Autonomous
Flight", RA-L 2018).  With `two_point_ransac` off (the default) the front end reproduces the
reference bit for bit; this file defines what "on" means, and libavb's k_ransac is tested against it mask for mask.

Choices where the published algorithm leaves room (all shared by the CUDA kernel):
  * randomness: a counter-based generator (splitmix64) keyed by (seed, frame index, camera, hypothesis) instead of
    a time-seeded engine, so the CPU and GPU draw the same pairs and a replay is deterministic;
  * previous-frame points are undistorted and rotated into the current frame with the full projective mapping
    (`undistort_points(pts, R_p_c)`, i.e. divide by the third coordinate), the operation the reference's own
    `CameraModel.undistort_points` offers (/root/reference/src/image_processing/camera_model.py:24-47);
  * float64 arithmetic on the float32 undistorted coordinates, no fused multiply-add, and the three sums
    (scaling factor, mean distance, its count) are taken in the fixed order `tree_sum` spells out;
  * the hypothesis that wins is the FIRST one with the largest inlier set (the published code compares set sizes
    only); the least-squares refit and its mean error are computed for reporting (`info`) -- they cannot change
    the mask, which is why the CUDA kernel does not compute them.
"""
from __future__ import annotations

import math

import numpy as np

M64 = (1 << 64) - 1
LANES = 256                       # threads per camera in k_ransac; fixes the summation order


def splitmix64(z: int) -> int:
    z = (z + 0x9E3779B97F4A7C15) & M64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M64
    return z ^ (z >> 31)


def draw_pair(seed: int, frame_index: int, cam: int, hyp: int, n: int):
    """Two distinct positions in [0, n): idx1 uniform, idx2 = idx1 + (1..n-1) mod n (Appendix C step 5)."""
    key = ((seed & 0xFFFFFF) << 40) | ((frame_index & 0xFFFFFFF) << 12) | ((cam & 0xF) << 8) | (hyp & 0xFF)
    r = splitmix64(key)
    i1 = ((r >> 32) * n) >> 32
    diff = 1 + (((r & 0xFFFFFFFF) * (n - 1)) >> 32)
    i2 = i1 + diff if i1 + diff < n else i1 + diff - n
    return int(i1), int(i2)


def tree_sum(v: np.ndarray) -> float:
    """Sum of a float64 vector in k_ransac's order: lane t adds v[t], v[t+256], ... in sequence, the 32 lanes of
    a warp meet in an xor butterfly (16, 8, 4, 2, 1), and the 8 warp totals are added in sequence."""
    v = np.asarray(v, dtype=np.float64)
    k = (len(v) + LANES - 1) // LANES
    pad = np.zeros(k * LANES, dtype=np.float64)
    pad[:len(v)] = v
    rows = pad.reshape(k, LANES)
    part = np.zeros(LANES, dtype=np.float64)
    for r in rows:
        part = part + r
    idx = np.arange(LANES)
    for off in (16, 8, 4, 2, 1):
        part = part + part[idx ^ off]
    total = part[0]
    for w in range(1, LANES // 32):
        total = total + part[32 * w]
    return float(total)


def num_iterations(success_probability=0.99):
    return int(math.ceil(math.log(1 - success_probability) / math.log(1 - 0.7 * 0.7)))


def two_point_ransac(u1, u2, intrinsics, inlier_error, seed=0, frame_index=0, cam=0, success_probability=0.99,
                     info=None):
    """u1: previous-frame points, undistorted to normalized coordinates AND rotated by R_p_c; u2: current-frame
    points, undistorted; both (N, 2) float32 (what cv2.undistortPoints returns for float32 input).
    Returns the inlier mask (N,) bool."""
    u1 = np.asarray(u1, dtype=np.float32).reshape(-1, 2).astype(np.float64)
    u2 = np.asarray(u2, dtype=np.float32).reshape(-1, 2).astype(np.float64)
    n = len(u1)
    mask = np.zeros(n, dtype=bool)
    if n == 0:
        return mask
    fx, fy = float(intrinsics[0]), float(intrinsics[1])
    unit = 2.0 / (fx + fy)
    iters = num_iterations(success_probability)

    with np.errstate(all='ignore'):
        # rescale so that the mean point norm is sqrt(2)  (camera_model.py:95-108)
        norms = np.sqrt(u1[:, 0] * u1[:, 0] + u1[:, 1] * u1[:, 1]) + np.sqrt(u2[:, 0] * u2[:, 0] + u2[:, 1] * u2[:, 1])
        sf = (2.0 * n) / tree_sum(norms) * math.sqrt(2.0)
        unit = unit * sf
        x1, y1, x2, y2 = u1[:, 0] * sf, u1[:, 1] * sf, u2[:, 0] * sf, u2[:, 1] * sf
        dx, dy = x1 - x2, y1 - y2
        dist = np.sqrt(dx * dx + dy * dy)
        raw = ~(dist > 50.0 * unit)
        n_raw = int(raw.sum())
        if n_raw < 3:
            return mask
        mean_dist = tree_sum(np.where(raw, dist, 0.0)) / n_raw
        thr = float(inlier_error) * unit
        if mean_dist < unit:                                   # degenerate: (almost) no translation
            return raw & ~(dist > thr)

        c0, c1, c2 = dy, -dx, x1 * y2 - y1 * x2                # epipolar constraint, linear in t = (tx, ty, tz)
        raw_idx = np.nonzero(raw)[0]
        best_count, best_mask, best_err = 0, mask, 1e10
        for h in range(iters):
            s1, s2 = draw_pair(seed, frame_index, cam, h, n_raw)
            i, j = int(raw_idx[s1]), int(raw_idx[s2])
            a = (c0[i], c0[j])
            b = (c1[i], c1[j])
            c = (c2[i], c2[j])
            l1 = (abs(a[0]) + abs(a[1]), abs(b[0]) + abs(b[1]), abs(c[0]) + abs(c[1]))
            base = 0 if (l1[0] <= l1[1] and l1[0] <= l1[2]) else (1 if l1[1] <= l1[2] else 2)   # first minimum

            def solve2(p, q, r):                               # [p q] s = -r, closed-form 2x2 inverse
                det = p[0] * q[1] - q[0] * p[1]
                r0, r1 = -r[0], -r[1]
                return (q[1] * r0 - q[0] * r1) / det, (p[0] * r1 - p[1] * r0) / det

            if base == 0:
                s = solve2(b, c, a)
                model = (1.0, s[0], s[1])
            elif base == 1:
                s = solve2(a, c, b)
                model = (s[0], 1.0, s[1])
            else:
                s = solve2(a, b, c)
                model = (s[0], s[1], 1.0)
            err = (c0 * model[0] + c1 * model[1]) + c2 * model[2]
            inl = raw & (np.abs(err) < thr)
            cnt = int(inl.sum())
            if cnt < 0.2 * n:
                continue
            # least-squares refit on the inlier set (reported only)
            cols = [c0[inl], c1[inl], c2[inl]]
            keep = [k for k in range(3) if k != base]
            A = np.stack([cols[keep[0]], cols[keep[1]]], axis=1)
            sol = np.linalg.lstsq(A, -cols[base], rcond=None)[0]
            better = [0.0, 0.0, 0.0]
            better[base] = 1.0
            better[keep[0]], better[keep[1]] = float(sol[0]), float(sol[1])
            this_err = float(np.abs(c0[inl] * better[0] + c1[inl] * better[1] + c2[inl] * better[2]).mean())
            if cnt > best_count:
                best_count, best_mask, best_err = cnt, inl, this_err
        if info is not None:
            info.update(best_count=best_count, best_error=best_err, unit=unit, scale=sf, n_raw=n_raw)
        return best_mask.copy()


def conjugate_rotation(R01, R0):
    """cam1_R_p_c from cam0_R_p_c when only the latter is supplied: both are the same gyro rotation seen from
    the two camera frames (imu_processor.py:55-64), so R1 = R01 R0 R01^T."""
    R01 = np.asarray(R01, dtype=np.float64).reshape(3, 3)
    return (R01 @ np.asarray(R0, dtype=np.float64).reshape(3, 3)) @ R01.T
