"""Seeded synthetic EuRoC-format stereo + IMU streams (numpy only, deterministic).

The reference ships no data and there is no network, so every test and benchmark in this
repo runs on streams made here.  Message types are the namedtuples the reference's dataset
readers emit (src/streaming/dataset.py:56-57 imu_msg, :101 img_msg, :168-169 stereo_msg),
so the same stream drives the reference pipeline, the oracle port and the CUDA front end.

Kernel-level streams ("sliding texture"): a band-limited noise texture (uniform u8 noise ->
separable Gaussian blur -> min-max normalise) viewed through a crop that drifts by a
sub-pixel amount per frame; cam1 sees the same texture shifted by a fixed disparity.  A
constant gyro rate consistent with zero rotation (i.e. zeros) accompanies it unless
`gyro` is given.
"""
from __future__ import annotations

from collections import namedtuple

import numpy as np

imu_msg = namedtuple('imu_msg', ['timestamp', 'angular_velocity', 'linear_acceleration'])
img_msg = namedtuple('img_msg', ['timestamp', 'image'])
stereo_msg = namedtuple('stereo_msg',
                        ['timestamp', 'cam0_image', 'cam1_image', 'cam0_msg', 'cam1_msg'])


def _gauss_kernel(sigma: float) -> np.ndarray:
    r = max(1, int(np.ceil(3.0 * sigma)))
    x = np.arange(-r, r + 1, dtype=np.float64)
    k = np.exp(-0.5 * (x / sigma) ** 2)
    return k / k.sum()


def _blur(img: np.ndarray, sigma: float) -> np.ndarray:
    k = _gauss_kernel(sigma)
    r = len(k) // 2
    p = np.pad(img, ((0, 0), (r, r)), mode='reflect')
    out = np.zeros_like(img)
    for i, kv in enumerate(k):
        out += kv * p[:, i:i + img.shape[1]]
    p = np.pad(out, ((r, r), (0, 0)), mode='reflect')
    out2 = np.zeros_like(img)
    for i, kv in enumerate(k):
        out2 += kv * p[i:i + img.shape[0], :]
    return out2


def make_texture(h: int, w: int, seed: int, sigma: float = 2.5) -> np.ndarray:
    """Band-limited noise texture, float64 in [0, 255]."""
    rng = np.random.default_rng(seed)
    t = rng.integers(0, 256, size=(h, w)).astype(np.float64)
    t = _blur(t, sigma)
    t -= t.min()
    t *= 255.0 / max(t.max(), 1e-9)
    return t


def _sample(tex: np.ndarray, x0: float, y0: float, h: int, w: int) -> np.ndarray:
    """Bilinear crop of size (h, w) whose top-left corner sits at (x0, y0) in the texture."""
    ix, iy = int(np.floor(x0)), int(np.floor(y0))
    fx, fy = x0 - ix, y0 - iy
    a = tex[iy:iy + h + 1, ix:ix + w + 1]
    v = ((1 - fx) * (1 - fy) * a[:-1, :-1] + fx * (1 - fy) * a[:-1, 1:]
         + (1 - fx) * fy * a[1:, :-1] + fx * fy * a[1:, 1:])
    return np.clip(np.rint(v), 0, 255).astype(np.uint8)


class SlidingTextureStream:
    """Deterministic stereo stream: n_frames frames at `rate` Hz plus `imu_rate` Hz IMU.

    Parameters mirror the survey probe: disparity 12 px, drift 1-2
    px/frame.  `frames()` yields stereo_msg; `imu()` yields imu_msg; `events()` interleaves
    them the way the deterministic driver wants (all IMU with t <= frame t first)."""

    def __init__(self, width=752, height=480, n_frames=20, seed=0, sigma=2.5,
                 disparity=12.0, drift=(1.3, 0.6), rate=20.0, imu_rate=200.0,
                 gyro=(0.0, 0.0, 0.0), t0=1000.0, noise=0.0, movers=(), blackout=None):
        self.w, self.h, self.n = int(width), int(height), int(n_frames)
        self.seed, self.sigma = int(seed), float(sigma)
        self.disparity = float(disparity)
        self.drift = (float(drift[0]), float(drift[1]))
        self.rate, self.imu_rate = float(rate), float(imu_rate)
        self.gyro = np.asarray(gyro, dtype=np.float64)
        self.t0 = float(t0)
        self.noise = float(noise)
        # {frame index: fraction of the image width}: the leftmost columns of both cameras are a flat grey in that frame
        # (1.0 = the whole frame: every feature is lost, the stream restarts from an empty grid without the first-frame
        # initialiser)
        self.blackout = {int(k): float(v) for k, v in (blackout or {}).items()}
        margin_x = int(abs(self.drift[0]) * self.n + abs(self.disparity)) + 8
        margin_y = int(abs(self.drift[1]) * self.n) + 8
        self._mx, self._my = margin_x, margin_y
        self._tex = make_texture(self.h + 2 * margin_y + 2, self.w + 2 * margin_x + 2,
                                 self.seed, self.sigma)
        self._rng_seed = self.seed * 7919 + 13
        # independently moving square patches (cx, cy, half, vx, vy): their own texture, same disparity in both cameras,
        # own image velocity -> stereo-consistent features whose temporal motion contradicts the camera motion
        # (what the two-point RANSAC stage is there to reject)
        self.movers = [tuple(float(v) for v in m) for m in movers]
        self._mover_tex = [make_texture(int(2 * m[2]) + 8, int(2 * m[2]) + 8, self.seed + 1000 + i, self.sigma)
                           for i, m in enumerate(self.movers)]

    def _paste_movers(self, img, k, shift_x):
        for (cx, cy, half, vx, vy), tex in zip(self.movers, self._mover_tex):
            px, py = cx + vx * k - shift_x - half, cy + vy * k - half      # top-left corner of the patch (fractional)
            x_lo, y_lo = int(np.ceil(px)), int(np.ceil(py))
            size = int(2 * half)
            x_a, x_b = max(x_lo, 0), min(x_lo + size, self.w)
            y_a, y_b = max(y_lo, 0), min(y_lo + size, self.h)
            if x_a >= x_b or y_a >= y_b:
                continue
            img[y_a:y_b, x_a:x_b] = _sample(tex, x_a - px + 2.0, y_a - py + 2.0, y_b - y_a, x_b - x_a)
        return img

    def frame(self, k: int):
        ts = self.t0 + k / self.rate
        x0 = self._mx + (self.drift[0] * k if self.drift[0] >= 0 else -self.drift[0] * (self.n - k))
        y0 = self._my + (self.drift[1] * k if self.drift[1] >= 0 else -self.drift[1] * (self.n - k))
        img0 = _sample(self._tex, x0, y0, self.h, self.w)
        # cam1 sees scene content shifted LEFT by the disparity: a cam0 point (x, y)
        # appears at (x - d, y) in cam1
        img1 = _sample(self._tex, x0 + self.disparity, y0, self.h, self.w)
        if self.movers:
            img0 = self._paste_movers(img0, k, 0.0)
            img1 = self._paste_movers(img1, k, self.disparity)
        if self.noise > 0:
            rng = np.random.default_rng(self._rng_seed + k)
            n0 = rng.normal(0.0, self.noise, size=img0.shape)
            n1 = rng.normal(0.0, self.noise, size=img1.shape)
            img0 = np.clip(np.rint(img0 + n0), 0, 255).astype(np.uint8)
            img1 = np.clip(np.rint(img1 + n1), 0, 255).astype(np.uint8)
        if k in self.blackout:
            cols = int(round(self.blackout[k] * self.w))
            img0, img1 = img0.copy(), img1.copy()
            img0[:, :cols] = 90
            img1[:, :cols] = 90
        m0, m1 = img_msg(ts, img0), img_msg(ts, img1)
        return stereo_msg(ts, img0, img1, m0, m1)

    def frames(self):
        for k in range(self.n):
            yield self.frame(k)

    def imu(self):
        n_imu = int(np.floor((self.n - 1) / self.rate * self.imu_rate)) + 1
        acc = np.array([0.0, 0.0, 9.81])
        lead = int(0.05 * self.imu_rate)
        for j in range(-lead, n_imu + 1):
            yield imu_msg(self.t0 + j / self.imu_rate, self.gyro.copy(), acc.copy())

    def events(self):
        """('imu', msg) / ('stereo', msg) in the deterministic-driver order."""
        imu_it = iter(self.imu())
        pending = next(imu_it, None)
        for f in self.frames():
            while pending is not None and pending.timestamp <= f.timestamp:
                yield 'imu', pending
                pending = next(imu_it, None)
            yield 'stereo', f


# ------------------------------------------------------------------------------------------------
# Rendered 3-D sequence (sequence-level tests: downstream MSCKF ATE parity, EuRoC on-disk round trip)
# ------------------------------------------------------------------------------------------------

gt_msg = namedtuple('gt_msg', ['timestamp', 'p', 'q', 'v', 'bw', 'ba'])


def _rodrigues(v):
    th = float(np.sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]))
    if th < 1e-12:
        return np.eye(3)
    r = v / th
    K = np.array([[0, -r[2], r[1]], [r[2], 0, -r[0]], [-r[1], r[0], 0]])
    return np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * (K @ K)


def _quat_wxyz(R):
    """Hamilton quaternion (w, x, y, z) of a rotation matrix, w >= 0 (EuRoC ground-truth column order)."""
    t = np.trace(R)
    if t > 0:
        s = np.sqrt(t + 1.0) * 2
        q = np.array([0.25 * s, (R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s])
    else:
        i = int(np.argmax(np.diag(R)))
        j, k = (i + 1) % 3, (i + 2) % 3
        s = np.sqrt(R[i, i] - R[j, j] - R[k, k] + 1.0) * 2
        q = np.zeros(4)
        q[0] = (R[k, j] - R[j, k]) / s
        q[1 + i] = 0.25 * s
        q[1 + j] = (R[j, i] + R[i, j]) / s
        q[1 + k] = (R[k, i] + R[i, k]) / s
    return q if q[0] >= 0 else -q


class RoomSceneStream:
    """Stereo + IMU stream of a rig moving inside a textured box room, rendered through the radtan / extrinsic model
    of a ConfigEuRoC-shaped calibration (T_imu_cam0/1 = Kalibr T_cam_imu: IMU frame -> camera frame).

    World: z up, gravity (0, 0, -9.81).  The rig rests for `static_s` seconds (the MSCKF initialises gravity from its
    first 200 IMU samples, reference msckf.py:172-175), then follows a smooth Lissajous-like path with small
    rotations.  IMU samples come from central differences of the analytic pose (specific force R^T (a + g), body
    rate), plus optional white noise.  Deterministic for a given seed."""

    G = 9.81

    def __init__(self, calib, n_frames=300, seed=0, rate=20.0, imu_rate=200.0, t0=1000.0, static_s=1.5,
                 half=(4.0, 3.5), floor=-1.5, ceil=2.0, tex_px_per_m=140.0, tex_sigma=3.0, tex_size=2048,
                 gyro_noise=2e-4, acc_noise=2e-3, pixel_noise=0.0, amp=1.0):
        self.calib = calib
        self.w, self.h = int(calib.cam0_resolution[0]), int(calib.cam0_resolution[1])
        self.n, self.seed = int(n_frames), int(seed)
        self.rate, self.imu_rate, self.t0, self.static_s = float(rate), float(imu_rate), float(t0), float(static_s)
        self.lo = np.array([-half[0], -half[1], floor])
        self.hi = np.array([half[0], half[1], ceil])
        self.scale, self.tex_size = float(tex_px_per_m), int(tex_size)
        self.gyro_noise, self.acc_noise, self.pixel_noise, self.amp = gyro_noise, acc_noise, pixel_noise, float(amp)
        self._tex = [make_texture(tex_size, tex_size, seed * 101 + 17 * f + 3, tex_sigma) for f in range(6)]
        self._rays = [self._pixel_rays(calib.cam0_intrinsics, calib.cam0_distortion_coeffs),
                      self._pixel_rays(calib.cam1_intrinsics, calib.cam1_distortion_coeffs)]
        self._T_ic = [np.linalg.inv(calib.T_imu_cam0), np.linalg.inv(calib.T_imu_cam1)]    # camera -> IMU
        # IMU x up, z (the optical axis, roughly) along world +x
        self._R0 = np.array([[0.0, 0.0, 1.0], [0.0, -1.0, 0.0], [1.0, 0.0, 0.0]])
        self._imu_cache = None

    # -- geometry ---------------------------------------------------------------------------------
    def _pixel_rays(self, intr, dist):
        """Normalized ray (x, y, 1) of every pixel centre: inverse of the radtan model by fixed-point iteration."""
        fx, fy, cx, cy = (float(v) for v in intr)
        k1, k2, p1, p2 = (float(v) for v in dist[:4])
        u, v = np.meshgrid(np.arange(self.w, dtype=np.float64), np.arange(self.h, dtype=np.float64))
        x0, y0 = (u - cx) / fx, (v - cy) / fy
        x, y = x0.copy(), y0.copy()
        for _ in range(30):
            r2 = x * x + y * y
            ic = 1.0 / (1.0 + (k2 * r2 + k1) * r2)
            dx = 2 * p1 * x * y + p2 * (r2 + 2 * x * x)
            dy = p1 * (r2 + 2 * y * y) + 2 * p2 * x * y
            x, y = (x0 - dx) * ic, (y0 - dy) * ic
        return np.stack([x.ravel(), y.ravel(), np.ones(x.size)], axis=1)

    def _ramp(self, t):
        s = np.clip((t - self.static_s) / 3.0, 0.0, 1.0)
        return s * s * s * (s * (6 * s - 15) + 10)

    def pose(self, t):
        """(R_wi, p_wi) of the IMU at time t seconds after t0."""
        s = self._ramp(t) * self.amp
        tt = t - self.static_s
        p = s * np.array([0.7 * np.sin(0.55 * tt), 0.9 * np.sin(0.40 * tt + 1.0) - 0.9 * np.sin(1.0),
                          0.35 * np.sin(0.75 * tt)])
        a = s * np.array([0.10 * np.sin(0.45 * tt), 0.12 * np.sin(0.33 * tt + 0.5) - 0.12 * np.sin(0.5),
                          0.22 * np.sin(0.27 * tt)])
        return self._R0 @ _rodrigues(a), p

    def imu_sample(self, t, h=1e-3):
        Rm, pm = self.pose(t - h)
        R, p = self.pose(t)
        Rp, pp = self.pose(t + h)
        acc_w = (pp - 2 * p + pm) / (h * h)
        W = (Rm.T @ Rp - Rp.T @ Rm) / (4 * h)                 # skew(omega_body) + O(h^2)
        gyro = np.array([W[2, 1], W[0, 2], W[1, 0]])
        acc = R.T @ (acc_w + np.array([0.0, 0.0, self.G]))
        return gyro, acc

    # -- rendering --------------------------------------------------------------------------------
    def _render(self, R_wc, c):
        d = self._rays_cam @ R_wc.T
        with np.errstate(divide='ignore', invalid='ignore'):
            tpos = np.where(d > 0, (self.hi - c) / d, np.where(d < 0, (self.lo - c) / d, np.inf))
        axis = np.argmin(tpos, axis=1)
        tmin = tpos[np.arange(len(axis)), axis]
        hit = c + d * tmin[:, None]
        face = axis * 2 + (d[np.arange(len(axis)), axis] > 0)
        out = np.zeros(len(axis))
        for f in range(6):
            m = face == f
            if not m.any():
                continue
            ax = f // 2
            ua, va = [(1, 2), (0, 2), (0, 1)][ax]
            tu = hit[m, ua] * self.scale + 0.37 * self.tex_size
            tv = hit[m, va] * self.scale + 0.61 * self.tex_size
            iu, iv = np.floor(tu).astype(np.int64), np.floor(tv).astype(np.int64)
            fu, fv = tu - iu, tv - iv
            T, n = self._tex[f], self.tex_size
            i0, i1, j0, j1 = iu % n, (iu + 1) % n, iv % n, (iv + 1) % n
            out[m] = ((1 - fu) * (1 - fv) * T[j0, i0] + fu * (1 - fv) * T[j0, i1]
                      + (1 - fu) * fv * T[j1, i0] + fu * fv * T[j1, i1])
        return out.reshape(self.h, self.w)

    def frame(self, k):
        t = k / self.rate
        ts = self.t0 + t
        R_wi, p = self.pose(t)
        imgs = []
        for cam in range(2):
            T = self._T_ic[cam]
            self._rays_cam = self._rays[cam]
            img = self._render(R_wi @ T[:3, :3], p + R_wi @ T[:3, 3])
            if self.pixel_noise > 0:
                img = img + np.random.default_rng(self.seed * 7919 + 2 * k + cam).normal(0.0, self.pixel_noise, img.shape)
            imgs.append(np.clip(np.rint(img), 0, 255).astype(np.uint8))
        m0, m1 = img_msg(ts, imgs[0]), img_msg(ts, imgs[1])
        return stereo_msg(ts, imgs[0], imgs[1], m0, m1)

    def frames(self):
        for k in range(self.n):
            yield self.frame(k)

    def imu(self):
        if self._imu_cache is None:
            rng = np.random.default_rng(self.seed + 12345)
            n_imu = int(np.floor((self.n - 1) / self.rate * self.imu_rate)) + 1
            out = []
            for j in range(n_imu + 1):
                t = j / self.imu_rate
                gyro, acc = self.imu_sample(t)
                gyro = gyro + rng.normal(0.0, self.gyro_noise, 3)
                acc = acc + rng.normal(0.0, self.acc_noise, 3)
                out.append(imu_msg(self.t0 + t, gyro, acc))
            self._imu_cache = out
        return iter(self._imu_cache)

    def groundtruth(self):
        """gt_msg per IMU sample: position, Hamilton quaternion (w, x, y, z) of R_wi, velocity, zero biases."""
        n_imu = int(np.floor((self.n - 1) / self.rate * self.imu_rate)) + 1
        h = 1e-3
        for j in range(n_imu + 1):
            t = j / self.imu_rate
            R, p = self.pose(t)
            v = (self.pose(t + h)[1] - self.pose(t - h)[1]) / (2 * h)
            yield gt_msg(self.t0 + t, p, _quat_wxyz(R), v, np.zeros(3), np.zeros(3))

    def events(self):
        imu_it = iter(self.imu())
        pending = next(imu_it, None)
        for f in self.frames():
            while pending is not None and pending.timestamp <= f.timestamp:
                yield 'imu', pending
                pending = next(imu_it, None)
            yield 'stereo', f
