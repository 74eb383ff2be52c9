"""StereoMatcher.stereo_match on the GPU: infinite-depth prediction (undistort with R_cam0_to_cam1, distort), forward
and backward pyramidal LK, forward-backward / vertical-disparity / bounds / epipolar filters -- one fused kernel
(k_stereo_points -> feature_chain).  Same constructor and return convention as the reference
(image_processing/stereo_matcher.py:7-115)."""
from __future__ import annotations

import numpy as np


class StereoMatcher:
    def __init__(self, lk_params, imu_processor, pyramid_builder, camera_model, stereo_threshold, context=None):
        self.lk_params = lk_params
        self.integrate_imu = imu_processor.integrate_imu_data
        self.R_cam0_imu, self.R_cam1_imu = imu_processor.R_cam0_imu, imu_processor.R_cam1_imu
        self.t_cam0_imu, self.t_cam1_imu = imu_processor.t_cam0_imu, imu_processor.t_cam1_imu
        self.pyr0, self.pyr1 = pyramid_builder.curr_cam0_pyramid, pyramid_builder.curr_cam1_pyramid
        self.camera_model = camera_model
        self.stereo_threshold = stereo_threshold
        self._ctx = context if context is not None else getattr(pyramid_builder, '_ctx', None)

    def _context(self):
        if self._ctx is None:
            from .pipeline import current_context
            self._ctx = current_context()
        return self._ctx

    def stereo_match(self, cam0_points):
        if len(cam0_points) == 0:
            return np.array([]), np.array([], dtype=bool)
        ctx = self._context()
        if float(self.stereo_threshold) != ctx.stereo_threshold:
            raise RuntimeError('stereo_threshold differs from the one the context was created with')
        return ctx.stereo_match(np.asarray(cam0_points, dtype=np.float32))
