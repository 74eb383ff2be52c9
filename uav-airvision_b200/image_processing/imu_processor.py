"""Gyro pre-integration between two image stamps.  Stays on the host (a few dozen flops per frame);
its 3x3 output is the only IMU-derived input of the CUDA frame chain.

Mirrors the reference's IMUProcessor (image_processing/imu_processor.py:5-67): same constructor, same
attributes, same window rule (B13): first message with t >= t_prev - 0.01 up to (excluding) the first with
t >= t_curr - 0.004; identity and no trimming when either bound is missing; no bias removal."""
from __future__ import annotations

import threading

import numpy as np


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
        with self._lock:
            buf = self.imu_buffer
            begin = next((i for i, m in enumerate(buf) if m.timestamp >= t_prev - 0.01), None)
            end = next((i for i, m in enumerate(buf) if m.timestamp >= t_curr - 0.004), None)
            if begin is None or end is None:
                return np.identity(3), np.identity(3)
            window = buf[begin:end]
            self.imu_buffer = buf[end:]
        mean_w = np.zeros(3)
        for m in window:
            mean_w += m.angular_velocity
        if end - begin > 0:
            mean_w /= (end - begin)
        dt = t_curr - t_prev
        cam0_R_p_c = rodrigues((self.R_cam0_imu.T @ mean_w) * dt).T
        cam1_R_p_c = rodrigues((self.R_cam1_imu.T @ mean_w) * dt).T
        return cam0_R_p_c, cam1_R_p_c
