"""FeaturePublisher: flattens the grid and undistorts cam0 points with the cam0 model and cam1 points with the cam1
model to normalized coordinates (image_processing/feature_publisher.py:10-121); the undistortion runs in libavb
(k_undistort).  `cam0_curr_img_msg`, `cam1_curr_img_msg`, `curr_features` are attached after construction
(pipeline.py:131-143)."""
from __future__ import annotations

from collections import namedtuple
from itertools import chain

import numpy as np

from .camera_model import CameraModel
from .feature_measurment import FeatureMeasurement

feature_msg = namedtuple('feature_msg', ['timestamp', 'features'])


class FeaturePublisher:
    def __init__(self, cam0_intrinsics, cam0_dist_model, cam0_dist_coeffs, cam1_intrinsics, cam1_dist_model,
                 cam1_dist_coeffs, context=None):
        self.cam0_intrinsics, self.cam0_dist_model, self.cam0_dist_coeffs = cam0_intrinsics, cam0_dist_model, cam0_dist_coeffs
        self.cam1_intrinsics, self.cam1_dist_model, self.cam1_dist_coeffs = cam1_intrinsics, cam1_dist_model, cam1_dist_coeffs
        self._cm = CameraModel(cam0_intrinsics, cam0_dist_model, cam0_dist_coeffs, context=context)

    def undistort_points(self, pts_in, intrinsics, distortion_model, distortion_coeffs,
                         rectification_matrix=np.identity(3), new_intrinsics=np.array([1, 1, 0, 0])):
        return self._cm.undistort_points(pts_in, intrinsics, distortion_model, distortion_coeffs,
                                         rectification_matrix, new_intrinsics)

    def publish(self):
        feats = list(chain.from_iterable(self.curr_features))
        # np.array over a list that mixes float32 arrays (tracked) and tuples (new) is float64 -> float64 result,
        # an all-float32 list stays float32: the reference's dtype behaviour (Appendix B11) falls out of numpy here too
        u0 = self.undistort_points(np.array([f.cam0_point for f in feats]), self.cam0_intrinsics, self.cam0_dist_model,
                                   self.cam0_dist_coeffs)
        u1 = self.undistort_points(np.array([f.cam1_point for f in feats]), self.cam1_intrinsics, self.cam1_dist_model,
                                   self.cam1_dist_coeffs)
        out = []
        for f, a, b in zip(feats, u0, u1):
            m = FeatureMeasurement()
            m.id, m.u0, m.v0, m.u1, m.v1 = f.id, a[0], a[1], b[0], b[1]
            out.append(m)
        return feature_msg(self.cam0_curr_img_msg.timestamp, out)
