"""CameraModel: point undistortion / distortion on the GPU (k_undistort in libavb), same call signatures
as the reference's CameraModel (image_processing/camera_model.py:5-75).  radtan only (EuRoC)."""
from __future__ import annotations

import numpy as np


def _io_dtype(pts_in):
    a = np.asarray(pts_in)
    return a.dtype if a.dtype in (np.float32, np.float64) else np.dtype(np.float64)


class CameraModel:
    def __init__(self, intrinsics, distortion_model, distortion_coeffs, context=None):
        self.intrinsics = intrinsics
        self.distortion_model = distortion_model
        self.distortion_coeffs = distortion_coeffs
        fx, fy, cx, cy = intrinsics
        self.K = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1]], dtype=float)
        self._ctx = context

    def _context(self):
        if self._ctx is not None:
            return self._ctx
        from .pipeline import current_context
        return current_context()

    @staticmethod
    def _check_model(model):
        if model != 'radtan':
            raise RuntimeError(f"distortion model '{model}' is not built into libavb (radtan only)")

    def undistort_points(self, pts_in, intrinsics, distortion_model, distortion_coeffs,
                         rectification_matrix=np.identity(3), new_intrinsics=np.array([1, 1, 0, 0])):
        if len(pts_in) == 0:
            return []
        self._check_model(distortion_model)
        dt = _io_dtype(pts_in)
        K_new = np.array([[new_intrinsics[0], 0.0, new_intrinsics[2]],
                          [0.0, new_intrinsics[1], new_intrinsics[3]], [0.0, 0.0, 1.0]])
        RR = K_new @ np.asarray(rectification_matrix, dtype=np.float64)      # cv2 folds P into R the same way
        out = self._context().undistort_ex(intrinsics, distortion_coeffs, np.reshape(pts_in, (-1, 2)), RR,
                                           f32_io=(dt == np.float32))
        return out.astype(dt)

    def distort_points(self, pts_in, intrinsics, distortion_model, distortion_coeffs):
        if len(pts_in) == 0:
            return []
        self._check_model(distortion_model)
        dt = _io_dtype(pts_in)
        out = self._context().distort_ex(intrinsics, distortion_coeffs, np.reshape(pts_in, (-1, 2)),
                                         f32_io=(dt == np.float32))
        return out.astype(dt)
