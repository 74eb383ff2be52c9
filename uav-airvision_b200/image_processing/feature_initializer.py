"""FeatureInitializer: first-frame feature set (image_processing/feature_initializer.py:6-85): FAST on cam0, stereo
match of every key point, per grid cell the `grid_min_feature_num` strongest (stable by scan order) get ids."""
from __future__ import annotations

import numpy as np

from .feature_meta_data import FeatureMetaData


def _grid_size(img, grid_row, grid_col):
    h, w = img.shape[:2]
    return int(np.ceil(h / grid_row)), int(np.ceil(w / grid_col))


def _top_by_response(feats, k):
    # Python's sort is stable: ties keep detection (scan) order, like sorted(..., reverse=True) in the reference
    return sorted(feats, key=lambda f: f.response, reverse=True)[:k]


class FeatureInitializer:
    def __init__(self, detector, stereo_matcher, config, cam0_curr_img_msg, curr_features, next_feature_id, grid_row,
                 grid_col, grid_min_feature_num):
        self.detector = detector
        self.stereo_match = stereo_matcher.stereo_match
        self.config = config
        self.cam0_curr_img_msg = cam0_curr_img_msg
        self.curr_features = curr_features
        self.next_feature_id = next_feature_id
        self.grid_row, self.grid_col = grid_row, grid_col
        self.grid_min_feature_num = grid_min_feature_num

    def get_grid_size(self, img):
        return _grid_size(img, self.grid_row, self.grid_col)

    def initialize_first_frame(self):
        img = self.cam0_curr_img_msg.image
        gh, gw = self.get_grid_size(img)
        kps = self.detector.detect(img)
        pts0 = [kp.pt for kp in kps]
        pts1, inlier = self.stereo_match(pts0)
        cells = [[] for _ in range(self.config.grid_num)]
        for kp, p1, ok in zip(kps, pts1, inlier):
            if not ok:
                continue
            fm = FeatureMetaData()
            fm.response, fm.cam0_point, fm.cam1_point = kp.response, kp.pt, p1
            cells[int(kp.pt[1] / gh) * self.grid_col + int(kp.pt[0] / gw)].append(fm)
        for idx, feats in enumerate(cells):
            for fm in _top_by_response(feats, self.grid_min_feature_num):
                fm.id, fm.lifetime = self.next_feature_id, 1
                self.next_feature_id += 1
                self.curr_features[idx].append(fm)
