"""FeatureTracker: temporal tracking of the previous frame's features (image_processing/feature_tracker.py:9-177):
gyro-compensated prediction, pyramidal LK previous cam0 -> current cam0 (k_klt_points), image-bounds cull, stereo match
of the survivors, the reference's all-ones RANSAC stub, re-binning into the grid with lifetime + 1."""
from __future__ import annotations

from itertools import chain

import numpy as np

from .feature_meta_data import FeatureMetaData
from .utils import select


class FeatureTracker:
    def __init__(self, lk_params, imu_processor, stereo_matcher, cam0_intrinsics, cam0_distortion_model,
                 cam0_distortion_coeffs, cam1_intrinsics, cam1_distortion_model, cam1_distortion_coeffs,
                 prev_cam0_pyramid, curr_cam0_pyramid, prev_features, curr_features, num_features, grid_row, grid_col,
                 ransac_threshold, context=None):
        self.lk_params = lk_params
        self.integrate_imu_data = imu_processor.integrate_imu_data
        self.R_cam0_imu, self.R_cam1_imu = imu_processor.R_cam0_imu, imu_processor.R_cam1_imu
        self.stereo_match = stereo_matcher.stereo_match
        self.cam0_intrinsics, self.cam0_dist_model, self.cam0_dist_coeffs = \
            cam0_intrinsics, cam0_distortion_model, cam0_distortion_coeffs
        self.cam1_intrinsics, self.cam1_dist_model, self.cam1_dist_coeffs = \
            cam1_intrinsics, cam1_distortion_model, cam1_distortion_coeffs
        self.prev_cam0_pyramid, self.curr_cam0_pyramid = prev_cam0_pyramid, curr_cam0_pyramid
        self.prev_features, self.curr_features = prev_features, curr_features
        self.num_features = num_features
        self.grid_row, self.grid_col = grid_row, grid_col
        self.ransac_threshold = ransac_threshold
        self._ctx = context if context is not None else getattr(stereo_matcher, '_ctx', None)

    def _context(self):
        if self._ctx is None:
            from .pipeline import current_context
            self._ctx = current_context()
        return self._ctx

    def get_grid_size(self, img):
        h, w = img.shape[:2]
        return int(np.ceil(h / self.grid_row)), int(np.ceil(w / self.grid_col))

    def predict_feature_tracking(self, input_pts, R_p_c, intrinsics):
        """p' = K R_p_c K^-1 p in float64, rounded to float32 (feature_tracker.py:159-177)."""
        if len(input_pts) == 0:
            return np.array([], dtype=np.float32)
        fx, fy, cx, cy = (float(v) for v in intrinsics[:4])
        K = np.array([[fx, 0.0, cx], [0.0, fy, cy], [0.0, 0.0, 1.0]])
        H = K @ R_p_c @ np.linalg.inv(K)
        p = np.asarray(input_pts, dtype=np.float64).reshape(-1, 2)
        hom = np.concatenate([p, np.ones((len(p), 1))], axis=1) @ H.T
        return (hom[:, :2] / hom[:, 2:3]).astype(np.float32)

    def track_features(self):
        img = self.curr_cam0_pyramid
        grid_h, grid_w = self.get_grid_size(img)
        cam0_R_p_c, _ = self.integrate_imu_data()
        prev = list(chain.from_iterable(self.prev_features))
        prev_pts = np.array([f.cam0_point for f in prev], dtype=np.float32).reshape(-1, 2)
        self.num_features['before_tracking'] = len(prev_pts)
        if len(prev_pts) == 0:
            return
        pred = self.predict_feature_tracking(prev_pts, cam0_R_p_c, self.cam0_intrinsics)
        # slots: 2 = previous cam0, 0 = current cam0 (include/avb.h)
        curr_pts, status = self._context().klt_track(2, 0, prev_pts, pred)
        h, w = img.shape[:2]
        keep = status.astype(bool) & ~((curr_pts[:, 0] < 0) | (curr_pts[:, 0] > w - 1) |
                                       (curr_pts[:, 1] < 0) | (curr_pts[:, 1] > h - 1))
        tracked = select(prev, keep)
        curr_tracked = curr_pts[keep]
        self.num_features['after_tracking'] = len(curr_tracked)
        cam1_pts, match = self.stereo_match(curr_tracked)
        matched = select(tracked, match)
        cm0, cm1 = select(curr_tracked, match), select(cam1_pts, match)
        self.num_features['after_matching'] = len(cm0)
        # two-point RANSAC is an all-ones stub in the reference (feature_tracker.py:135-136): nothing is dropped
        n = 0
        for f, p0, p1 in zip(matched, cm0, cm1):
            cell = int(p0[1] / grid_h) * self.grid_col + int(p0[0] / grid_w)
            fm = FeatureMetaData()
            fm.id, fm.lifetime, fm.cam0_point, fm.cam1_point = f.id, f.lifetime + 1, p0, p1
            self.curr_features[cell].append(fm)
            n += 1
        self.num_features['after_ransac'] = n
