"""ImageProcessingPipeline: the reference's per-frame orchestrator (image_processing/pipeline.py:14-150)
with the whole stage sequence -- pyramid build, temporal KLT, stereo KLT + filters, FAST + grid ranking,
new-feature stereo match, prune, publish -- executed as ONE CUDA-graph launch inside libavb
(avb_process_frame).  Stereo message in, feature_msg(timestamp, [FeatureMeasurement]) out."""
from __future__ import annotations

from collections import defaultdict, namedtuple

import numpy as np

from . import _avbhost, _native
from .camera_model import CameraModel
from .fast_detector import FastDetector
from .feature_adder import FeatureAdder
from .feature_initializer import FeatureInitializer
from .feature_measurment import FeatureMeasurement
from .feature_meta_data import FeatureMetaData
from .feature_pruner import FeaturePruner
from .feature_publisher import FeaturePublisher
from .feature_tracker import FeatureTracker
from .imu_processor import IMUProcessor
from .pyramid_builder import PyramidBuilder
from .stereo_matcher import StereoMatcher

feature_msg = namedtuple('feature_msg', ['timestamp', 'features'])


class ImageProcessingPipeline:
    """mode='fused' (default): the whole frame is one CUDA-graph launch (avb_process_frame).
    mode='staged': the reference's own orchestration (pipeline.py:46-150) over the stage classes of this package, each
    stage a separate libavb call -- slower, there to keep every stage individually callable and to cross-check the
    fused path."""

    def __init__(self, config, device=0, use_graph=True, mode='fused'):
        if mode not in ('fused', 'staged'):
            raise ValueError("mode must be 'fused' or 'staged'")
        self.config = config
        self.mode = mode
        self.prev_cam0_msg = None
        self.imu_processor = IMUProcessor(config.T_imu_cam0, config.T_imu_cam1)
        self.detector = FastDetector(config.fast_threshold, self._ensure_context)   # FAST lives in libavb (k_fast)
        self.camera_model = CameraModel(config.cam0_intrinsics, config.cam0_distortion_model,
                                        config.cam0_distortion_coeffs)
        self._staged_prev = [[] for _ in range(config.grid_num)]
        self._staged_curr = [[] for _ in range(config.grid_num)]
        self.next_feature_id = 0
        self.num_features = defaultdict(int)
        self.first_frame = True
        self.prev_pyr0 = None
        self._device, self._use_graph = device, use_graph
        self._ctx = None
        self._grid_cache = None
        self._new = FeatureMeasurement.__new__

    # -- plumbing ------------------------------------------------------------------------------------------
    @property
    def context(self):
        return self._ctx

    def _ensure_context(self, img):
        if self._ctx is None:
            h, w = img.shape[:2]
            self._ctx = _native.Context(self.config, w, h, num_streams=1, device=self._device,
                                        use_graph=self._use_graph)
            _native_register(self._ctx)
        return self._ctx

    def imu_callback(self, imu_msg):
        self.imu_processor.imu_callback(imu_msg)

    # -- the reference's orchestration over the stage classes -------------------------------------------------
    def _stereo_callback_staged(self, stereo_msg):
        cfg = self.config
        cam0_msg, cam1_msg = stereo_msg.cam0_msg, stereo_msg.cam1_msg
        ctx = self._ctx or self._ensure_context(cam0_msg.image)
        self.camera_model._ctx = ctx
        self.imu_processor.cam0_prev_img_msg = self.prev_cam0_msg
        self.imu_processor.cam0_curr_img_msg = cam0_msg
        builder = PyramidBuilder(cfg.win_size, cfg.pyramid_levels, cam0_msg, cam1_msg, context=ctx)
        pyr0, _ = builder.create_image_pyramids()
        matcher = StereoMatcher(cfg.lk_params, self.imu_processor, builder, self.camera_model, cfg.stereo_threshold)
        curr = self._staged_curr
        if self.first_frame:
            init = FeatureInitializer(detector=self.detector, stereo_matcher=matcher, config=cfg, cam0_curr_img_msg=cam0_msg,
                                      curr_features=curr, next_feature_id=self.next_feature_id, grid_row=cfg.grid_row,
                                      grid_col=cfg.grid_col, grid_min_feature_num=cfg.grid_min_feature_num)
            init.initialize_first_frame()
            self.next_feature_id = init.next_feature_id
            self.first_frame = False
        else:
            tracker = FeatureTracker(lk_params=cfg.lk_params, imu_processor=self.imu_processor, stereo_matcher=matcher,
                                     cam0_intrinsics=cfg.cam0_intrinsics, cam0_distortion_model=cfg.cam0_distortion_model,
                                     cam0_distortion_coeffs=cfg.cam0_distortion_coeffs, cam1_intrinsics=cfg.cam1_intrinsics,
                                     cam1_distortion_model=cfg.cam1_distortion_model,
                                     cam1_distortion_coeffs=cfg.cam1_distortion_coeffs, prev_cam0_pyramid=self.prev_pyr0,
                                     curr_cam0_pyramid=pyr0, prev_features=self._staged_prev, curr_features=curr,
                                     num_features=self.num_features, grid_row=cfg.grid_row, grid_col=cfg.grid_col,
                                     ransac_threshold=cfg.ransac_threshold)
            tracker.track_features()
            adder = FeatureAdder(detector=self.detector, stereo_matcher=matcher, config=cfg, cam0_curr_img_msg=cam0_msg,
                                 curr_features=curr, next_feature_id=self.next_feature_id, grid_row=cfg.grid_row,
                                 grid_col=cfg.grid_col, grid_max_feature_num=cfg.grid_max_feature_num,
                                 grid_min_feature_num=cfg.grid_min_feature_num)
            adder.add_new_features()
            self.next_feature_id = adder.next_feature_id
            pruner = FeaturePruner(cfg.grid_max_feature_num)
            pruner.curr_features, pruner.config = curr, cfg
            pruner.prune_features()
        publisher = FeaturePublisher(cfg.cam0_intrinsics, cfg.cam0_distortion_model, cfg.cam0_distortion_coeffs,
                                     cfg.cam1_intrinsics, cfg.cam1_distortion_model, cfg.cam1_distortion_coeffs, context=ctx)
        publisher.cam0_curr_img_msg, publisher.cam1_curr_img_msg, publisher.curr_features = cam0_msg, cam1_msg, curr
        msg = publisher.publish()
        self.prev_cam0_msg = cam0_msg
        self._staged_prev = curr
        self._staged_curr = [[] for _ in range(cfg.grid_num)]
        self.prev_pyr0 = pyr0
        return msg

    # -- the hot path -----------------------------------------------------------------------------------------
    def stereo_callback(self, stereo_msg):
        if self.mode == 'staged':
            return self._stereo_callback_staged(stereo_msg)
        cam0_msg, cam1_msg = stereo_msg.cam0_msg, stereo_msg.cam1_msg
        ctx = self._ctx or self._ensure_context(cam0_msg.image)
        imu = self.imu_processor
        imu.cam0_prev_img_msg = self.prev_cam0_msg
        imu.cam0_curr_img_msg = cam0_msg
        # host images -> pinned staging -> H2D -> CUDA-graph frame -> D2H -> FeatureMeasurement list, all in C.  The
        # image copies (and FAST behind cam0's) go out first; the gyro window is integrated while they are on the bus
        # (the reference integrates at the top of stereo_callback, pipeline.py:52-55: same inputs, same result).
        h = ctx._h.value
        if self.first_frame:
            feats, hdr = _avbhost.process_frame(h, cam0_msg.image, cam1_msg.image, None, FeatureMeasurement)
        else:
            _avbhost.submit_images(h, cam0_msg.image, cam1_msg.image)
            R = None
            try:
                R, R1 = imu.integrate_imu_data()
                if ctx.ransac:          # the RANSAC kernel compensates each camera with its own gyro rotation
                    R = np.ascontiguousarray(np.stack([R, R1]), dtype=np.float64)
            finally:                    # a submitted frame is always finished (identity rotation if the window failed)
                feats, hdr = _avbhost.finish_frame(h, R, FeatureMeasurement)
        self.next_feature_id = hdr[1]
        if not self.first_frame:
            nf = self.num_features
            nf['before_tracking'] = hdr[2]
            if hdr[2]:
                nf['after_tracking'], nf['after_matching'], nf['after_ransac'] = hdr[3], hdr[4], hdr[5]
        self.first_frame = False
        self.prev_cam0_msg = cam0_msg
        self.prev_pyr0 = cam0_msg.image
        self._grid_cache = None
        return feature_msg(cam0_msg.timestamp, feats)

    # -- state read-back (pipeline.prev_features / curr_features of the reference) ---------------------------
    @property
    def prev_features(self):
        """Grid of FeatureMetaData lists as the reference holds it after the roll (pipeline.py:145-148)."""
        if self.mode == 'staged':
            return self._staged_prev
        if self._grid_cache is None:
            grid = [[] for _ in range(self.config.grid_num)]
            if self._ctx is not None and not self.first_frame:
                _, ids, _ = self._ctx.result(0)
                cell, life, p0, p1 = self._ctx.features(0)
                for i in range(len(ids)):
                    f = FeatureMetaData()
                    f.id, f.lifetime = int(ids[i]), int(life[i])
                    f.cam0_point, f.cam1_point = p0[i].copy(), p1[i].copy()
                    grid[int(cell[i])].append(f)
            self._grid_cache = grid
        return self._grid_cache

    @property
    def curr_features(self):
        return self._staged_curr if self.mode == 'staged' else [[] for _ in range(self.config.grid_num)]


_current_ctx = None


def _native_register(ctx):
    global _current_ctx
    _current_ctx = ctx


def current_context():
    """Most recently created libavb context (the stage classes use it when none is passed)."""
    if _current_ctx is None:
        raise RuntimeError('no libavb context yet: create an ImageProcessingPipeline / call create_context first')
    return _current_ctx


def create_context(config, width, height, num_streams=1, device=0, use_graph=True):
    ctx = _native.Context(config, width, height, num_streams=num_streams, device=device, use_graph=use_graph)
    _native_register(ctx)
    return ctx
