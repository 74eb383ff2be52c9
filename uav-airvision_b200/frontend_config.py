"""Config objects shaped like the reference's ConfigEuRoC (/root/reference/src/config.py:19-123) for the BASELINE
configurations C1/C2/C3.  Only the front-end fields are carried; values are the EuRoC
calibration.  Plain data: used by bench.py, the tools and the tests."""
from __future__ import annotations

import numpy as np

TERM_COUNT, TERM_EPS, OPTFLOW_USE_INITIAL_FLOW = 1, 2, 4      # cv2 constants


class FrontEndConfig:
    def __init__(self, grid_row=4, grid_col=5, grid_min=3, grid_max=5, pyramid_levels=3,
                 patch_size=15, fast_threshold=15, width=752, height=480, two_point_ransac=False, ransac_seed=0):
        self.grid_row, self.grid_col = grid_row, grid_col
        self.grid_num = grid_row * grid_col
        self.grid_min_feature_num, self.grid_max_feature_num = grid_min, grid_max
        self.fast_threshold = fast_threshold
        self.ransac_threshold = 3
        # not reference fields: the reference's RANSAC is an all-ones stub (feature_tracker.py:135-136); True switches
        # on the two-point RANSAC of oracle/ransac.py / libavb k_ransac (BASELINE config C3)
        self.two_point_ransac = bool(two_point_ransac)
        self.ransac_seed = int(ransac_seed)
        self.stereo_threshold = 5
        self.max_iteration = 30
        self.track_precision = 0.01
        self.pyramid_levels = pyramid_levels
        self.patch_size = patch_size
        self.win_size = (patch_size, patch_size)
        self.lk_params = dict(winSize=self.win_size, maxLevel=pyramid_levels,
                              criteria=(TERM_EPS | TERM_COUNT, self.max_iteration,
                                        self.track_precision),
                              flags=OPTFLOW_USE_INITIAL_FLOW)
        self.T_imu_cam0 = np.array([
            [0.014865542981794, 0.999557249008346, -0.025774436697440, 0.065222909535531],
            [-0.999880929698575, 0.014967213324719, 0.003756188357967, -0.020706385492719],
            [0.004140296794224, 0.025715529947966, 0.999660727177902, -0.008054602460030],
            [0, 0, 0, 1.0]])
        self.cam0_camera_model = 'pinhole'
        self.cam0_distortion_model = 'radtan'
        self.cam0_distortion_coeffs = np.array([-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05])
        self.cam0_intrinsics = np.array([458.654, 457.296, 367.215, 248.375])
        self.cam0_resolution = np.array([width, height])
        self.T_imu_cam1 = np.array([
            [0.012555267089103, 0.999598781151433, -0.025389800891747, -0.044901980682509],
            [-0.999755099723116, 0.013011905181504, 0.017900583825251, -0.020569771258915],
            [0.018223771455443, 0.025158836311552, 0.999517347077547, -0.008638135126028],
            [0, 0, 0, 1.0]])
        self.cam1_camera_model = 'pinhole'
        self.cam1_distortion_model = 'radtan'
        self.cam1_distortion_coeffs = np.array([-0.28368365, 0.07451284, -0.00010473, -3.55590700e-05])
        self.cam1_intrinsics = np.array([457.587, 456.134, 379.999, 255.238])
        self.cam1_resolution = np.array([width, height])
        if (width, height) != (752, 480):
            # 1280x1024 stress config: scale the pinhole parameters with the image (stated in bench output)
            sx, sy = width / 752.0, height / 480.0
            for k in ('cam0_intrinsics', 'cam1_intrinsics'):
                v = getattr(self, k)
                setattr(self, k, np.array([v[0] * sx, v[1] * sy, v[2] * sx, v[3] * sy]))


class OptimizationConfig:
    """Triangulation settings (reference config.py:7-17, OptimizationConfigEuRoC)."""
    translation_threshold = -1.0
    huber_epsilon = 0.01
    estimation_precision = 5e-7
    initial_damping = 1e-3
    outer_loop_max_iteration = 5
    inner_loop_max_iteration = 5


def with_filter_fields(cfg):
    """Adds the MSCKF fields of the reference's ConfigEuRoC (config.py:47-72, 93-122) to a front-end config, so one
    object configures the front end and its consumer (msckf.MSCKF) the way ConfigEuRoC does."""
    cfg.optimization_config = OptimizationConfig()
    cfg.gravity = np.array([0.0, 0.0, -9.81])
    cfg.max_cam_state_size = 20
    cfg.position_std_threshold = 2.0
    cfg.gyro_noise, cfg.acc_noise = 0.005 ** 2, 0.05 ** 2
    cfg.gyro_bias_noise, cfg.acc_bias_noise = 0.001 ** 2, 0.01 ** 2
    cfg.observation_noise = 0.035 ** 2
    cfg.velocity = np.zeros(3)
    cfg.velocity_cov, cfg.gyro_bias_cov, cfg.acc_bias_cov = 0.25, 0.01, 0.01
    cfg.extrinsic_rotation_cov, cfg.extrinsic_translation_cov = 3.0462e-4, 2.5e-5
    cfg.T_cn_cnm1 = np.array([
        [0.999997256477881, 0.002312067192424, 0.000376008102415, -0.110073808127187],
        [-0.002317135723281, 0.999898048506644, 0.014089835846648, 0.000399121547014],
        [-0.000343393120525, -0.014090668452714, 0.999900662637729, -0.000853702503357],
        [0, 0, 0, 1.0]])
    cfg.T_imu_body = np.identity(4)
    return cfg


def config_default():            # the reference's shipped config: 4x5 cells, cap 100
    return FrontEndConfig()


def config_c1():                 # 752x480, 150 features
    return FrontEndConfig(grid_row=5, grid_col=6)


def config_c2():                 # 752x480, 300 features, 4 levels, win 15
    return FrontEndConfig(grid_row=6, grid_col=10)


def config_c3(two_point_ransac=True):   # 1280x1024, 2000 features, 5 levels, batched two-point RANSAC
    return FrontEndConfig(grid_row=10, grid_col=10, grid_min=10, grid_max=20, pyramid_levels=4,
                          width=1280, height=1024, two_point_ransac=two_point_ransac)
