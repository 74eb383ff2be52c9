"""CPU-side checks of the drop-in boundary: libavb.so loads, exports every symbol include/avb.h declares,
struct layouts agree, and there is NO CPU fallback (context creation fails loudly without a GPU)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    src = open(os.path.join(ROOT, 'include', 'avb.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(avb_[a-z0-9_]+)\s*\(', src)))


def test_header_cites_reference_interfaces():
    src = open(os.path.join(ROOT, 'include', 'avb.h')).read()
    for needle in ('pipeline.py:46-150', 'stereo_matcher.py:33-115', 'feature_tracker.py:102-108',
                   'camera_model.py:24-47', 'pyramid_builder.py:22-48', 'feature_adder.py:64'):
        assert needle in src


def test_library_exports_every_declared_symbol():
    from image_processing import _native
    lib = _native.load()
    names = _declared_functions()
    assert len(names) >= 25
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert set(names) == set(_native.EXPORTS)
    assert lib.avb_abi_version() == 5


def test_struct_layouts_match_header(tmp_path):
    """sizeof / offsetof of every avb_config and avb_frame_header field, as gcc lays out include/avb.h, equal the
    ctypes mirrors."""
    import subprocess
    from image_processing import _native
    prog = ['#include <stdio.h>', '#include <stddef.h>', '#include "avb.h"', 'int main(void) {']
    for cname, cls in (('avb_config', _native.AvbConfig), ('avb_frame_header', _native.AvbFrameHeader)):
        prog.append(f'  printf("{cname} %zu\\n", sizeof({cname}));')
        for f in cls._fields_:
            prog.append(f'  printf("{cname}.{f[0]} %zu\\n", offsetof({cname}, {f[0]}));')
    prog += ['  return 0;', '}']
    src = tmp_path / 'layout.c'
    src.write_text('\n'.join(prog))
    exe = tmp_path / 'layout'
    subprocess.run(['gcc', '-I', os.path.join(ROOT, 'include'), str(src), '-o', str(exe)], check=True)
    got = dict(line.split() for line in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    for cname, cls in (('avb_config', _native.AvbConfig), ('avb_frame_header', _native.AvbFrameHeader)):
        assert int(got[cname]) == C.sizeof(cls), cname
        for f in cls._fields_:
            assert int(got[f'{cname}.{f[0]}']) == getattr(cls, f[0]).offset, f'{cname}.{f[0]}'
    assert C.sizeof(_native.AvbFrameHeader) == 48


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip('a GPU is present; the no-device path cannot be exercised here')
    from image_processing import _native
    from frontend_config import config_default
    with pytest.raises(RuntimeError, match='no CUDA device|AVB|avb_create'):
        _native.Context(config_default(), 752, 480)
    from image_processing import ImageProcessor
    from synth_euroc import SlidingTextureStream
    ip = ImageProcessor(config_default())
    with pytest.raises(RuntimeError):
        ip.stereo_callback(SlidingTextureStream(n_frames=1).frame(0))
    with pytest.raises(RuntimeError, match='avb_store_create'):      # the HBM frame store has no host stand-in either
        _native.FrameStore(752, 480, 4)
    lib = _native.load()
    assert lib.avb_store_num_frames(None) == 0 and lib.avb_store_image(None, 0, 0) is None
    assert lib.avb_process_frame_gather(None, None, None, None) == -1   # AVB_E_INVALID


def test_product_package_does_not_import_oracle():
    pkg = os.path.join(ROOT, 'uav-airvision_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                text = open(os.path.join(dirpath, f)).read()
                assert 'import oracle' not in text and 'from oracle' not in text, f
                assert 'import cv2' not in text, f'{f}: the product path must not lean on cv2'


def test_imu_processor_matches_port():
    """Host-side gyro integration (C helper behind the Python class) against the oracle port, incl. the window
    quirks (B13).  Equal to a few ulp: the C code sums R^T w in a fixed order, numpy's matmul in its own."""
    from image_processing import IMUProcessor
    from frontend_config import config_default
    from oracle.pipeline_port import FrontEndPort
    from synth_euroc import img_msg, imu_msg
    cfg = config_default()
    a = IMUProcessor(cfg.T_imu_cam0, cfg.T_imu_cam1)
    b = FrontEndPort.__new__(FrontEndPort)
    T0, T1 = np.linalg.inv(cfg.T_imu_cam0), np.linalg.inv(cfg.T_imu_cam1)
    b.R_cam0_imu, b.R_cam1_imu, b.imu_buffer = T0[:3, :3], T1[:3, :3], []
    g = np.random.default_rng(0)
    msgs = [imu_msg(10.0 + i * 0.005, g.normal(0, 0.3, 3), np.zeros(3)) for i in range(60)]
    for m in msgs[:45]:
        a.imu_callback(m)
        b.imu_buffer.append(m)
    for t_prev, t_curr in ((10.0, 10.05), (10.05, 10.1), (10.1, 10.15), (10.15, 10.6)):
        a.cam0_prev_img_msg, a.cam0_curr_img_msg = img_msg(t_prev, None), img_msg(t_curr, None)
        Ra0, Ra1 = a.integrate_imu_data()
        Rb0, Rb1 = b._integrate_imu(t_prev, t_curr)
        assert np.abs(Ra0 - Rb0).max() < 1e-15 and np.abs(Ra1 - Rb1).max() < 1e-15
        assert len(a.imu_buffer) == len(b.imu_buffer)
    assert np.array_equal(Ra0, np.eye(3))          # last window has no end message: identity, no trim


def test_host_extension_imu_integration_matches_python_rule():
    """_avbhost.integrate_imu (C) == the window rule of the reference's IMUProcessor.integrate_imu_data
    (imu_processor.py:28-67): first t >= t_prev-0.01 .. first t >= t_curr-0.004, identity + no trim when open."""
    from image_processing.imu_processor import IMUProcessor, rodrigues
    from frontend_config import config_default
    from synth_euroc import img_msg, imu_msg
    cfg = config_default()
    imu = IMUProcessor(cfg.T_imu_cam0, cfg.T_imu_cam1)
    rng = np.random.default_rng(3)
    t0 = 50.0
    # no IMU yet -> identity, buffer untouched
    imu.cam0_prev_img_msg, imu.cam0_curr_img_msg = img_msg(t0 - 0.05, None), img_msg(t0, None)
    R0, R1 = imu.integrate_imu_data()
    assert np.array_equal(R0, np.eye(3)) and np.array_equal(R1, np.eye(3)) and imu.imu_buffer == []
    for k in range(40):
        for j in range(10):
            imu.imu_callback(imu_msg(t0 + k * 0.05 + j * 0.005, rng.normal(size=3) * 0.4, np.zeros(3)))
        tp, tc = t0 + k * 0.05 - 0.05, t0 + k * 0.05
        imu.cam0_prev_img_msg, imu.cam0_curr_img_msg = img_msg(tp, None), img_msg(tc, None)
        buf = list(imu.imu_buffer)
        R0, R1 = imu.integrate_imu_data()
        b = next((i for i, m in enumerate(buf) if m.timestamp >= tp - 0.01), None)
        e = next((i for i, m in enumerate(buf) if m.timestamp >= tc - 0.004), None)
        if b is None or e is None:
            assert np.array_equal(R0, np.eye(3)) and len(imu.imu_buffer) == len(buf)
            continue
        w = np.zeros(3)
        for m in buf[b:e]:
            w += m.angular_velocity
        if e - b > 0:
            w /= (e - b)
        assert np.abs(R0 - rodrigues((imu.R_cam0_imu.T @ w) * (tc - tp)).T).max() < 1e-15
        assert np.abs(R1 - rodrigues((imu.R_cam1_imu.T @ w) * (tc - tp)).T).max() < 1e-15
        assert imu.imu_buffer == buf[e:]


def test_package_exports_match_the_reference_surface():
    """Same export list as the reference's image_processing/__init__.py:1-12 (+ ImageProcessor with the legacy alias,
    :14-27); constructor parameter names of the stage classes as the reference's pipeline passes them."""
    import inspect
    import image_processing as ip
    for name in ('ImageProcessingPipeline', 'CameraModel', 'IMUProcessor', 'PyramidBuilder', 'FeatureMetaData',
                 'FeatureMeasurement', 'FeatureInitializer', 'FeatureAdder', 'FeatureTracker', 'FeaturePruner',
                 'StereoMatcher', 'FeaturePublisher', 'ImageProcessor'):
        assert hasattr(ip, name), name
    assert ip.ImageProcessor.stareo_callback is ip.ImageProcessingPipeline.stereo_callback
    want = {
        ip.PyramidBuilder: ['win_size', 'pyramid_levels', 'cam0_curr_img_msg', 'cam1_curr_img_msg'],
        ip.StereoMatcher: ['lk_params', 'imu_processor', 'pyramid_builder', 'camera_model', 'stereo_threshold'],
        ip.FeatureInitializer: ['detector', 'stereo_matcher', 'config', 'cam0_curr_img_msg', 'curr_features',
                                'next_feature_id', 'grid_row', 'grid_col', 'grid_min_feature_num'],
        ip.FeatureAdder: ['detector', 'stereo_matcher', 'config', 'cam0_curr_img_msg', 'curr_features', 'next_feature_id',
                          'grid_row', 'grid_col', 'grid_max_feature_num', 'grid_min_feature_num'],
        ip.FeatureTracker: ['lk_params', 'imu_processor', 'stereo_matcher', 'cam0_intrinsics', 'cam0_distortion_model',
                            'cam0_distortion_coeffs', 'cam1_intrinsics', 'cam1_distortion_model', 'cam1_distortion_coeffs',
                            'prev_cam0_pyramid', 'curr_cam0_pyramid', 'prev_features', 'curr_features', 'num_features',
                            'grid_row', 'grid_col', 'ransac_threshold'],
        ip.FeaturePruner: ['grid_max_feature_num'],
        ip.FeaturePublisher: ['cam0_intrinsics', 'cam0_dist_model', 'cam0_dist_coeffs', 'cam1_intrinsics',
                              'cam1_dist_model', 'cam1_dist_coeffs'],
        ip.CameraModel: ['intrinsics', 'distortion_model', 'distortion_coeffs'],
        ip.IMUProcessor: ['T_imu_cam0', 'T_imu_cam1'],
    }
    for cls, names in want.items():
        params = [p for p in inspect.signature(cls.__init__).parameters if p != 'self']
        assert params[:len(names)] == names, (cls.__name__, params)


def test_host_side_stage_logic_without_gpu():
    """FeaturePruner and FeatureTracker.predict_feature_tracking are pure host code: check them against the rules
    (stable lifetime ranking, B9; H = K R K^-1 in float64 -> float32)."""
    from image_processing import FeatureMetaData, FeaturePruner, FeatureTracker, IMUProcessor
    from frontend_config import config_default
    cfg = config_default()
    feats = []
    for i, life in enumerate([3, 7, 7, 1, 7, 2, 9]):
        f = FeatureMetaData()
        f.id, f.lifetime = i, life
        feats.append(f)
    pr = FeaturePruner(cfg.grid_max_feature_num)
    pr.curr_features, pr.config = [list(feats), feats[:2]], cfg
    pr.prune_features()
    assert [f.id for f in pr.curr_features[0]] == [6, 1, 2, 4, 0] and len(pr.curr_features[1]) == 2

    class _SM:
        stereo_match = None
    imu = IMUProcessor(cfg.T_imu_cam0, cfg.T_imu_cam1)
    tr = FeatureTracker(cfg.lk_params, imu, _SM(), cfg.cam0_intrinsics, 'radtan', cfg.cam0_distortion_coeffs,
                        cfg.cam1_intrinsics, 'radtan', cfg.cam1_distortion_coeffs, None, None, [], [], {}, 4, 5, 3)
    from image_processing.imu_processor import rodrigues
    R = rodrigues(np.array([0.01, -0.02, 0.015]))
    pts = np.array([[10.5, 20.25], [700.0, 400.0]], np.float32)
    got = tr.predict_feature_tracking(pts, R, cfg.cam0_intrinsics)
    fx, fy, cx, cy = cfg.cam0_intrinsics
    K = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1.0]])
    H = K @ R @ np.linalg.inv(K)
    exp = []
    for p in pts:
        h = H @ np.array([p[0], p[1], 1.0])
        exp.append([h[0] / h[2], h[1] / h[2]])
    assert got.dtype == np.float32 and np.array_equal(got, np.array(exp, np.float32))
    assert tr.get_grid_size(np.zeros((480, 752), np.uint8)) == (120, 151)


def test_feature_measurement_objects_and_list_builder():
    """FeatureMeasurement (C type, feature_measurment.py:1-9 attribute set): constructor, attributes, subclassing; the
    list builder used by the hot path and the estimator workers; freed objects are recycled without leaking state."""
    from image_processing import FeatureMeasurement, _avbhost
    f = FeatureMeasurement(7, 0.1, 0.2, 0.3, 0.4)
    assert (f.id, f.u0, f.v0, f.u1, f.v1) == (7, 0.1, 0.2, 0.3, 0.4) and 'id=7' in repr(f)
    g = FeatureMeasurement()
    g.id, g.u0 = 3, -1.5
    assert g.id == 3 and g.u0 == -1.5 and g.v1 == 0.0

    class Sub(FeatureMeasurement):
        pass
    s = Sub(1, 2.0, 3.0, 4.0, 5.0)
    s.extra = 'x'
    assert s.v1 == 5.0 and s.extra == 'x'
    del s
    ids = np.arange(100, 400, dtype=np.int64)
    meas = np.random.default_rng(0).normal(size=(300, 4))
    a = _avbhost.features_from_arrays(ids, meas, FeatureMeasurement)
    assert len(a) == 300 and [x.id for x in a] == ids.tolist()
    assert np.array_equal(np.array([[x.u0, x.v0, x.u1, x.v1] for x in a]), meas)
    keep = a[5]
    del a                                                     # 299 objects go to the free list, one stays alive
    b = _avbhost.features_from_arrays(ids[:50] + 1000, meas[:50] * 2, FeatureMeasurement)
    assert keep.id == 105 and keep.u0 == meas[5, 0]           # not recycled while referenced
    assert [x.id for x in b] == (ids[:50] + 1000).tolist() and b[7].v0 == 2 * meas[7, 1]
    assert all(x is not keep for x in b)
    # any class with the five attributes can be requested instead (generic path)
    class Plain:
        pass
    c = _avbhost.features_from_arrays(ids[:3], meas[:3], Plain)
    assert [type(x) for x in c] == [Plain] * 3 and c[2].id == 102 and c[2].u1 == meas[2, 2]
    assert _avbhost.features_from_arrays(ids[:0], meas[:0], FeatureMeasurement) == []
    with pytest.raises(ValueError):
        _avbhost.features_from_arrays(ids, meas[:10], FeatureMeasurement)
