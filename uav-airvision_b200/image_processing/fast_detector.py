"""FAST detector object with the cv2 surface the stage classes use: `detector.detect(img, mask=None)` returning
key points with `.pt` and `.response` in row-major scan order (what cv2.FastFeatureDetector_create(thr) gives the
reference at image_processing/pipeline.py:23-25).  The scoring + NMS run in libavb (k_fast)."""
from __future__ import annotations

import numpy as np


class KeyPoint:
    """The two cv2.KeyPoint fields the front end reads (feature_initializer.py:55-65, feature_adder.py:66-90)."""
    __slots__ = ('pt', 'response', 'size', 'angle', 'octave')

    def __init__(self, x, y, response):
        self.pt = (float(x), float(y))
        self.response = float(response)
        self.size, self.angle, self.octave = 7.0, -1.0, 0


class FastDetector:
    def __init__(self, threshold, context_getter):
        self.threshold = int(threshold)
        self._get = context_getter

    def getThreshold(self):
        return self.threshold

    def detect_arrays(self, img, mask=None):
        """(xs, ys, responses) int32 arrays in scan order."""
        ctx = self._get(img)
        ctx.ensure_current_cam0(img)
        return ctx.fast_detect(mask)

    def detect(self, img, mask=None):
        xs, ys, rs = self.detect_arrays(img, mask)
        return [KeyPoint(x, y, r) for x, y, r in zip(xs.tolist(), ys.tolist(), rs.tolist())]
