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
                 gyro=(0.0, 0.0, 0.0), t0=1000.0, noise=0.0):
        self.w, self.h, self.n = int(width), int(height), int(n_frames)
        self.seed, self.sigma = int(seed), float(sigma)
        self.disparity = float(disparity)
        self.drift = (float(drift[0]), float(drift[1]))
        self.rate, self.imu_rate = float(rate), float(imu_rate)
        self.gyro = np.asarray(gyro, dtype=np.float64)
        self.t0 = float(t0)
        self.noise = float(noise)
        margin_x = int(abs(self.drift[0]) * self.n + abs(self.disparity)) + 8
        margin_y = int(abs(self.drift[1]) * self.n) + 8
        self._mx, self._my = margin_x, margin_y
        self._tex = make_texture(self.h + 2 * margin_y + 2, self.w + 2 * margin_x + 2,
                                 self.seed, self.sigma)
        self._rng_seed = self.seed * 7919 + 13

    def frame(self, k: int):
        ts = self.t0 + k / self.rate
        x0 = self._mx + (self.drift[0] * k if self.drift[0] >= 0 else -self.drift[0] * (self.n - k))
        y0 = self._my + (self.drift[1] * k if self.drift[1] >= 0 else -self.drift[1] * (self.n - k))
        img0 = _sample(self._tex, x0, y0, self.h, self.w)
        # cam1 sees scene content shifted LEFT by the disparity: a cam0 point (x, y)
        # appears at (x - d, y) in cam1
        img1 = _sample(self._tex, x0 + self.disparity, y0, self.h, self.w)
        if self.noise > 0:
            rng = np.random.default_rng(self._rng_seed + k)
            n0 = rng.normal(0.0, self.noise, size=img0.shape)
            n1 = rng.normal(0.0, self.noise, size=img1.shape)
            img0 = np.clip(np.rint(img0 + n0), 0, 255).astype(np.uint8)
            img1 = np.clip(np.rint(img1 + n1), 0, 255).astype(np.uint8)
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
