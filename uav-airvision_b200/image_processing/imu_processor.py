"""Gyro pre-integration between two image stamps.  Stays on the host (a few dozen flops per frame);
its 3x3 output is the only IMU-derived input of the CUDA frame chain.

Mirrors the reference's IMUProcessor (image_processing/imu_processor.py:5-67): same constructor, same
attributes, same window rule (B13): first message with t >= t_prev - 0.01 up to (excluding) the first with
t >= t_curr - 0.004; identity and no trimming when either bound is missing; no bias removal."""
from __future__ import annotations

import threading

import numpy as np

from . import _avbhost


def rodrigues(v):
    """Rotation matrix of an axis-angle 3-vector (what cv2.Rodrigues(v)[0] returns)."""
    v = np.asarray(v, dtype=np.float64).reshape(3)
    theta = float(np.sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]))
    if theta < np.finfo(np.float64).eps:
        return np.eye(3)
    r = v / theta
    c, s = np.cos(theta), np.sin(theta)
    rx = np.array([[0, -r[2], r[1]], [r[2], 0, -r[0]], [-r[1], r[0], 0]])
    return c * np.eye(3) + (1.0 - c) * np.outer(r, r) + s * rx


class IMUProcessor:
    def __init__(self, T_imu_cam0, T_imu_cam1):
        self.T_cam0_imu = np.linalg.inv(T_imu_cam0)
        self.R_cam0_imu = self.T_cam0_imu[:3, :3]
        self.t_cam0_imu = self.T_cam0_imu[:3, 3]
        self.T_cam1_imu = np.linalg.inv(T_imu_cam1)
        self.R_cam1_imu = self.T_cam1_imu[:3, :3]
        self.t_cam1_imu = self.T_cam1_imu[:3, 3]
        self._R0 = np.ascontiguousarray(self.R_cam0_imu, dtype=np.float64)
        self._R1 = np.ascontiguousarray(self.R_cam1_imu, dtype=np.float64)
        self.imu_buffer = []
        self.cam0_prev_img_msg = None
        self.cam0_curr_img_msg = None
        # the reference appends from the IMU thread while the image thread re-slices, unsynchronised
        # (vio.py:38-44); here the two sides meet under a lock
        self._lock = threading.Lock()

    def imu_callback(self, msg):
        with self._lock:
            self.imu_buffer.append(msg)

    def integrate_imu_data(self):
        t_prev = self.cam0_prev_img_msg.timestamp
        t_curr = self.cam0_curr_img_msg.timestamp
        cam0_R_p_c, cam1_R_p_c = np.empty((3, 3)), np.empty((3, 3))
        with self._lock:
            buf = self.imu_buffer
            # window scan, mean rate, Rodrigues and the transposes run in C (csrc/avb_host.c: integrate_imu)
            end = _avbhost.integrate_imu(buf, float(t_prev), float(t_curr), self._R0, self._R1,
                                         cam0_R_p_c, cam1_R_p_c)
            if end >= 0:
                self.imu_buffer = buf[end:]
        return cam0_R_p_c, cam1_R_p_c
