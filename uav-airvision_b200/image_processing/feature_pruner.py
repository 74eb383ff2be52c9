"""FeaturePruner: caps every grid cell at `grid_max_feature_num`, keeping the longest-lived features, ties in list
order (image_processing/feature_pruner.py:1-19).  `curr_features` and `config` are attached after construction, as the
reference's pipeline does (pipeline.py:126-129)."""
from __future__ import annotations


class FeaturePruner:
    def __init__(self, grid_max_feature_num):
        self.grid_max_feature_num = grid_max_feature_num

    def prune_features(self):
        cap = self.config.grid_max_feature_num
        for i, feats in enumerate(self.curr_features):
            if len(feats) > cap:
                self.curr_features[i] = sorted(feats, key=lambda f: f.lifetime, reverse=True)[:cap]
