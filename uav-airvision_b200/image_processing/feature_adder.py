"""FeatureAdder: new features for sparsely covered cells (image_processing/feature_adder.py:8-108): 7x7 mask around
the live features (with the reference's negative-slice quirk, Appendix B7), masked FAST, per-cell cap at
`grid_max_feature_num` by response, stereo match, per cell the `grid_min_feature_num` strongest inliers get ids."""
from __future__ import annotations

from itertools import chain

import numpy as np

from .feature_initializer import _grid_size, _top_by_response
from .feature_meta_data import FeatureMetaData


class FeatureAdder:
    def __init__(self, detector, stereo_matcher, config, cam0_curr_img_msg, curr_features, next_feature_id, grid_row,
                 grid_col, grid_max_feature_num, grid_min_feature_num):
        self.detector = detector
        self.stereo_matcher = stereo_matcher
        self.stereo_match = stereo_matcher.stereo_match
        self.config = config
        self.cam0_curr_img_msg = cam0_curr_img_msg
        self.curr_features = curr_features
        self.next_feature_id = next_feature_id
        self.grid_row, self.grid_col = grid_row, grid_col
        self.grid_max_feature_num, self.grid_min_feature_num = grid_max_feature_num, grid_min_feature_num

    def get_grid_size(self, img):
        return _grid_size(img, self.grid_row, self.grid_col)

    def add_new_features(self):
        img = self.cam0_curr_img_msg.image
        gh, gw = self.get_grid_size(img)
        mask = np.ones(img.shape[:2], dtype=np.uint8)
        for f in chain.from_iterable(self.curr_features):
            x, y = int(f.cam0_point[0]), int(f.cam0_point[1])
            mask[y - 3:y + 4, x - 3:x + 4] = 0        # a negative start wraps: features within 3 px of the top/left mask nothing
        kps = self.detector.detect(img, mask=mask)
        sieve = [[] for _ in range(self.config.grid_num)]
        for kp in kps:
            sieve[int(kp.pt[1] / gh) * self.grid_col + int(kp.pt[0] / gw)].append(kp)
        cand = []
        for cell in sieve:
            cand.extend(_top_by_response(cell, self.grid_max_feature_num) if len(cell) > self.grid_max_feature_num else cell)
        pts0 = [kp.pt for kp in cand]
        pts1, inlier = self.stereo_match(pts0)
        cells = [[] for _ in range(self.config.grid_num)]
        for kp, p1, ok in zip(cand, pts1, inlier):
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
