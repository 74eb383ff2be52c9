"""ctypes binding of libavb.so (include/avb.h).  No CPU fallback: if the library or a CUDA device is
missing, loading / context creation raises."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), 'lib', 'libavb.so')

EXPORTS = (
    'avb_abi_version', 'avb_last_error', 'avb_create', 'avb_destroy', 'avb_capacity', 'avb_num_cells',
    'avb_reset', 'avb_input_staging', 'avb_input_block_bytes', 'avb_input_rotation_offset',
    'avb_input_rotation_stride',
    'avb_fill_rotations', 'avb_process_frame', 'avb_submit_images', 'avb_process_submitted', 'avb_process_frame_device', 'avb_enqueue_frame_device',
    'avb_sync', 'avb_get_result', 'avb_get_result_prev', 'avb_get_features', 'avb_upload_stereo', 'avb_advance',
    'avb_build_pyramids', 'avb_download_level', 'avb_fast_detect', 'avb_klt_track', 'avb_stereo_match',
    'avb_undistort_points', 'avb_distort_points', 'avb_two_point_ransac', 'avb_last_frame_ms', 'avb_kernels_per_frame',
    'avb_cuda_stream', 'avb_time_pyramid', 'avb_profile_frame_device', 'avb_get_geometry',
    'avb_store_create', 'avb_store_destroy', 'avb_store_num_frames', 'avb_store_bytes', 'avb_store_upload',
    'avb_store_image', 'avb_process_frame_gather', 'avb_enqueue_frame_gather',
)


class AvbConfig(C.Structure):
    _fields_ = [
        ('width', C.c_int32), ('height', C.c_int32), ('max_level', C.c_int32), ('win_size', C.c_int32),
        ('max_iteration', C.c_int32), ('fast_threshold', C.c_int32), ('grid_row', C.c_int32),
        ('grid_col', C.c_int32), ('grid_min_feature_num', C.c_int32), ('grid_max_feature_num', C.c_int32),
        ('num_streams', C.c_int32), ('device', C.c_int32), ('use_graph', C.c_int32), ('ransac', C.c_int32),
        ('ransac_seed', C.c_int32), ('reserved0', C.c_int32),
        ('track_precision', C.c_double), ('min_eig_threshold', C.c_double), ('stereo_threshold', C.c_double),
        ('ransac_threshold', C.c_double),
        ('cam0_intrinsics', C.c_double * 4), ('cam0_distortion', C.c_double * 4),
        ('cam1_intrinsics', C.c_double * 4), ('cam1_distortion', C.c_double * 4),
        ('R_cam0_to_cam1', C.c_double * 9), ('essential', C.c_double * 9),
    ]


class AvbFrameHeader(C.Structure):
    _fields_ = [
        ('n_features', C.c_int64), ('next_feature_id', C.c_int64),
        ('before_tracking', C.c_int32), ('after_tracking', C.c_int32), ('after_matching', C.c_int32),
        ('after_ransac', C.c_int32), ('has_new', C.c_int32), ('n_fast', C.c_int32),
        ('n_candidates', C.c_int32), ('frame_index', C.c_int32),
    ]


HEADER_DTYPE = np.dtype([
    ('n_features', '<i8'), ('next_feature_id', '<i8'), ('before_tracking', '<i4'), ('after_tracking', '<i4'),
    ('after_matching', '<i4'), ('after_ransac', '<i4'), ('has_new', '<i4'), ('n_fast', '<i4'),
    ('n_candidates', '<i4'), ('frame_index', '<i4')])
assert HEADER_DTYPE.itemsize == C.sizeof(AvbFrameHeader) == 48

_lib = None


def load():
    """Loads libavb.so (once).  Raises OSError with a build hint when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise OSError(f'{LIB_PATH} not found: build it with `python uav-airvision_b200/build.py` '
                      '(nvcc, sm_100a). There is no CPU fallback.')
    lib = C.CDLL(LIB_PATH)
    vp, ip, u8p, f32p, f64p, i32p = C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p
    lib.avb_abi_version.restype = C.c_int
    lib.avb_last_error.restype = C.c_char_p
    lib.avb_last_error.argtypes = [vp]
    lib.avb_create.argtypes = [C.POINTER(AvbConfig), C.POINTER(vp)]
    lib.avb_destroy.argtypes = [vp]
    lib.avb_destroy.restype = None
    for name in ('avb_capacity', 'avb_num_cells', 'avb_reset', 'avb_sync', 'avb_advance', 'avb_build_pyramids',
                 'avb_kernels_per_frame'):
        getattr(lib, name).argtypes = [vp]
    lib.avb_input_staging.argtypes = [vp]
    lib.avb_input_staging.restype = C.c_void_p
    lib.avb_input_block_bytes.argtypes = [vp]
    lib.avb_input_block_bytes.restype = C.c_size_t
    lib.avb_input_rotation_offset.argtypes = [vp]
    lib.avb_input_rotation_offset.restype = C.c_size_t
    lib.avb_input_rotation_stride.argtypes = [vp]
    lib.avb_input_rotation_stride.restype = C.c_size_t
    lib.avb_fill_rotations.argtypes = [vp, u8p, f64p, f64p]
    lib.avb_process_frame.argtypes = [vp, vp, vp, ip, f64p, f64p]
    lib.avb_submit_images.argtypes = [vp, vp, vp, ip]
    lib.avb_process_submitted.argtypes = [vp, f64p, f64p]
    lib.avb_process_frame_device.argtypes = [vp, vp]
    lib.avb_enqueue_frame_device.argtypes = [vp, vp]
    lib.avb_get_result.argtypes = [vp, ip, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]
    lib.avb_get_result_prev.argtypes = [vp, ip, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]
    lib.avb_get_features.argtypes = [vp, ip, i32p, i32p, f32p, f32p]
    lib.avb_upload_stereo.argtypes = [vp, ip, u8p, u8p, ip]
    lib.avb_download_level.argtypes = [vp, ip, ip, ip, u8p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.avb_fast_detect.argtypes = [vp, ip, u8p, i32p, i32p, i32p, C.POINTER(C.c_int)]
    lib.avb_klt_track.argtypes = [vp, ip, ip, ip, f32p, f32p, ip, f32p, u8p]
    lib.avb_stereo_match.argtypes = [vp, ip, f32p, ip, f32p, u8p]
    lib.avb_undistort_points.argtypes = [vp, f64p, f64p, f64p, ip, f64p, ip, f64p]
    lib.avb_distort_points.argtypes = [vp, f64p, f64p, f64p, ip, ip, f64p]
    lib.avb_two_point_ransac.argtypes = [vp, f64p, f64p, f32p, f32p, ip, f64p, C.c_double, ip, ip, ip, u8p]
    lib.avb_last_frame_ms.argtypes = [vp, C.POINTER(C.c_float)]
    lib.avb_cuda_stream.argtypes = [vp]
    lib.avb_cuda_stream.restype = C.c_void_p
    lib.avb_time_pyramid.argtypes = [vp, ip, C.POINTER(C.c_float)]
    lib.avb_profile_frame_device.argtypes = [vp, vp, C.POINTER(C.c_float)]
    lib.avb_store_create.argtypes = [ip, ip, ip, ip, C.POINTER(vp)]
    lib.avb_store_destroy.argtypes = [vp]
    lib.avb_store_destroy.restype = None
    lib.avb_store_num_frames.argtypes = [vp]
    lib.avb_store_bytes.argtypes = [vp]
    lib.avb_store_bytes.restype = C.c_size_t
    lib.avb_store_upload.argtypes = [vp, ip, u8p, u8p, ip]
    lib.avb_store_image.argtypes = [vp, ip, ip]
    lib.avb_store_image.restype = C.c_void_p
    lib.avb_process_frame_gather.argtypes = [vp, vp, f64p, f64p]
    lib.avb_enqueue_frame_gather.argtypes = [vp, vp, f64p, f64p]
    if lib.avb_abi_version() != 5:
        raise OSError('libavb.so ABI version mismatch: rebuild')
    _lib = lib
    return lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def stereo_geometry(cfg):
    """R_cam0_to_cam1 and the essential matrix exactly as StereoMatcher forms them
    (reference stereo_matcher.py:49, 90-91; imu_processor.py:11-17)."""
    T0 = np.linalg.inv(cfg.T_imu_cam0)
    T1 = np.linalg.inv(cfg.T_imu_cam1)
    R0, t0, R1, t1 = T0[:3, :3], T0[:3, 3], T1[:3, :3], T1[:3, 3]
    R01 = R1.T @ R0
    t01 = R1.T @ (t0 - t1)
    x, y, z = t01
    E = np.array([[0, -z, y], [z, 0, -x], [-y, x, 0]]) @ R01
    return R01, E


class FrameStore:
    """A decoded stereo sequence resident in HBM (avb_store_*): uploaded once, read by every time-offset run of the
    sequence (reference: each `main.py --offset` run re-reads its PNGs, streaming/dataset.py:101-117, 206-214)."""

    def __init__(self, width, height, n_frames, device=0):
        self._lib = load()
        self._h = C.c_void_p()
        rc = self._lib.avb_store_create(int(device), int(width), int(height), int(n_frames), C.byref(self._h))
        if rc != 0:
            raise RuntimeError(f'avb_store_create failed ({rc}): no CUDA device, bad geometry (width % 16) or out of memory')
        self.width, self.height, self.n_frames, self.device = int(width), int(height), int(n_frames), int(device)
        self.timestamps = [None] * self.n_frames
        # device addresses [frame][cam], looked up once: the per-step pointer table is a numpy fancy index
        self.addr = np.array([[self._lib.avb_store_image(self._h, k, cam) for cam in (0, 1)]
                              for k in range(self.n_frames)], dtype=np.uint64)

    @property
    def nbytes(self):
        return int(self._lib.avb_store_bytes(self._h))

    def upload(self, k, img0, img1, timestamp=None):
        for im in (img0, img1):
            if im.dtype != np.uint8 or im.shape != (self.height, self.width) or im.strides[1] != 1:
                raise ValueError(f'frames must be ({self.height}, {self.width}) uint8 with unit column stride')
        if img0.strides[0] != img1.strides[0]:
            img0, img1 = np.ascontiguousarray(img0), np.ascontiguousarray(img1)
        rc = self._lib.avb_store_upload(self._h, int(k), _ptr(img0), _ptr(img1), int(img0.strides[0]))
        if rc != 0:
            raise RuntimeError(f'avb_store_upload failed ({rc})')
        self.timestamps[k] = timestamp

    def close(self):
        if getattr(self, '_h', None) is not None and self._h.value:
            self._lib.avb_store_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Context:
    """One libavb context = S lock-stepped stereo streams on one CUDA device."""

    def __init__(self, cfg, width, height, num_streams=1, device=0, use_graph=True):
        lib = load()
        for k in ('cam0_distortion_model', 'cam1_distortion_model'):
            if getattr(cfg, k, 'radtan') != 'radtan':
                raise RuntimeError('libavb implements the radtan model only (EuRoC, config.py:99,116)')
        ac = AvbConfig()
        ac.width, ac.height = int(width), int(height)
        ac.max_level = int(cfg.lk_params['maxLevel'])
        ws = cfg.lk_params['winSize']
        if ws[0] != ws[1]:
            raise RuntimeError('square LK window required')
        ac.win_size = int(ws[0])
        crit = cfg.lk_params['criteria']
        ac.max_iteration, ac.track_precision = int(crit[1]), float(crit[2])
        ac.min_eig_threshold = float(cfg.lk_params.get('minEigThreshold', 1e-4))
        ac.fast_threshold = int(cfg.fast_threshold)
        ac.grid_row, ac.grid_col = int(cfg.grid_row), int(cfg.grid_col)
        ac.grid_min_feature_num, ac.grid_max_feature_num = int(cfg.grid_min_feature_num), int(cfg.grid_max_feature_num)
        ac.num_streams, ac.device, ac.use_graph = int(num_streams), int(device), int(bool(use_graph))
        # not reference fields (the reference's RANSAC is an all-ones stub): see frontend_config.FrontEndConfig
        ac.ransac = int(bool(getattr(cfg, 'two_point_ransac', False)))
        ac.ransac_seed = int(getattr(cfg, 'ransac_seed', 0))
        ac.stereo_threshold = float(cfg.stereo_threshold)
        ac.ransac_threshold = float(getattr(cfg, 'ransac_threshold', 3))
        ac.cam0_intrinsics[:] = [float(v) for v in cfg.cam0_intrinsics]
        ac.cam0_distortion[:] = [float(v) for v in cfg.cam0_distortion_coeffs[:4]]
        ac.cam1_intrinsics[:] = [float(v) for v in cfg.cam1_intrinsics]
        ac.cam1_distortion[:] = [float(v) for v in cfg.cam1_distortion_coeffs[:4]]
        R01, E = stereo_geometry(cfg)
        ac.R_cam0_to_cam1[:] = [float(v) for v in R01.reshape(-1)]
        ac.essential[:] = [float(v) for v in E.reshape(-1)]
        self._lib = lib
        self._h = C.c_void_p()
        rc = lib.avb_create(C.byref(ac), C.byref(self._h))
        if rc != 0:
            raise RuntimeError(f'avb_create failed ({rc}): {lib.avb_last_error(None).decode()}')
        self.width, self.height, self.S = ac.width, ac.height, ac.num_streams
        self.capacity = lib.avb_capacity(self._h)
        self.num_cells = lib.avb_num_cells(self._h)
        self.max_level = ac.max_level
        self.block_bytes = lib.avb_input_block_bytes(self._h)
        self.rot_offset = lib.avb_input_rotation_offset(self._h)
        self.rot_stride = lib.avb_input_rotation_stride(self._h)
        self.ransac = bool(ac.ransac)
        base = lib.avb_input_staging(self._h)
        img_bytes = self.S * 2 * self.width * self.height
        self._staging_block = np.ctypeslib.as_array(C.cast(base, C.POINTER(C.c_uint8)), shape=(self.block_bytes,))
        self.staging = self._staging_block[:img_bytes].reshape(self.S, 2, self.height, self.width)
        self._ident = np.tile(np.eye(3).reshape(-1), self.S)
        self._views = {}
        self._blocks = {}
        self.stereo_threshold = float(ac.stereo_threshold)
        self._staged_frames = 0          # frames made current through begin_frame (stage-class flow)
        self._cur0 = None

    # -- lifecycle -------------------------------------------------------------------------------------
    def close(self):
        if getattr(self, '_h', None) is not None and self._h.value:
            self._lib.avb_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise RuntimeError(f'libavb error {rc}: {self._lib.avb_last_error(self._h).decode()}')

    def reset(self):
        self._ck(self._lib.avb_reset(self._h))

    # -- hot path ----------------------------------------------------------------------------------------
    def process_staged(self, R_p_c0=None, R_p_c1=None):
        """One frame from the pinned staging block (caller filled self.staging[s, cam])."""
        R = self._ident if R_p_c0 is None else np.ascontiguousarray(R_p_c0, dtype=np.float64).reshape(-1)
        R1 = None if (R_p_c0 is None or R_p_c1 is None) else np.ascontiguousarray(R_p_c1, dtype=np.float64).reshape(-1)
        self._ck(self._lib.avb_process_frame(self._h, None, None, self.width, _ptr(R), _ptr(R1)))

    def process(self, imgs0, imgs1, R_p_c0=None, R_p_c1=None):
        """One frame: imgs0[s], imgs1[s] are (H, W) uint8 arrays (copied into pinned staging)."""
        for s in range(self.S):
            self.staging[s, 0] = imgs0[s]
            self.staging[s, 1] = imgs1[s]
        self.process_staged(R_p_c0, R_p_c1)

    def submit_images(self, imgs0, imgs1):
        """First half of a frame straight from the callers' arrays (avb_submit_images): imgs0[s], imgs1[s] are (H, W) uint8
        arrays with unit column stride and one common row stride; they must stay alive until process_submitted()."""
        a0 = [np.asarray(a) for a in imgs0]
        a1 = [np.asarray(a) for a in imgs1]
        stride = a0[0].strides[0]
        for a in a0 + a1:
            if a.dtype != np.uint8 or a.shape != (self.height, self.width) or a.strides != (stride, 1):
                raise ValueError('images must be (H, W) uint8 arrays with unit column stride and equal row stride')
        p0 = (C.c_void_p * self.S)(*[a.ctypes.data for a in a0])
        p1 = (C.c_void_p * self.S)(*[a.ctypes.data for a in a1])
        self._submitted = (a0, a1)                      # keeps the arrays alive
        self._ck(self._lib.avb_submit_images(self._h, p0, p1, int(stride)))

    def process_submitted(self, R_p_c0=None, R_p_c1=None):
        """Second half (avb_process_submitted): rotations in, results in host memory when it returns."""
        R = None if R_p_c0 is None else np.ascontiguousarray(R_p_c0, dtype=np.float64).reshape(-1)
        R1 = None if (R is None or R_p_c1 is None) else np.ascontiguousarray(R_p_c1, dtype=np.float64).reshape(-1)
        try:
            self._ck(self._lib.avb_process_submitted(self._h, _ptr(R), _ptr(R1)))
        finally:
            self._submitted = None

    def process_gather(self, image_addrs, R_p_c0=None, R_p_c1=None, wait=True):
        """One frame from device-resident images: image_addrs = uint64[S, 2] device addresses (FrameStore.addr rows).
        wait=False only enqueues (sync() before reading results or enqueuing again)."""
        tab = np.ascontiguousarray(image_addrs, dtype=np.uint64).reshape(-1)
        if tab.size != 2 * self.S:
            raise ValueError(f'expected {2 * self.S} image addresses')
        R = None if R_p_c0 is None else np.ascontiguousarray(R_p_c0, dtype=np.float64).reshape(-1)
        R1 = None if (R is None or R_p_c1 is None) else np.ascontiguousarray(R_p_c1, dtype=np.float64).reshape(-1)
        fn = self._lib.avb_process_frame_gather if wait else self._lib.avb_enqueue_frame_gather
        self._ck(fn(self._h, _ptr(tab), _ptr(R), _ptr(R1)))

    def fill_rotations(self, block, R_p_c0=None, R_p_c1=None):
        """Rotation section of an input block: cam0_R_p_c / cam1_R_p_c per stream (None: identity / conjugated)."""
        R = None if R_p_c0 is None else np.ascontiguousarray(R_p_c0, dtype=np.float64).reshape(-1)
        R1 = None if (R is None or R_p_c1 is None) else np.ascontiguousarray(R_p_c1, dtype=np.float64).reshape(-1)
        self._ck(self._lib.avb_fill_rotations(self._h, _ptr(block), _ptr(R), _ptr(R1)))

    def process_device(self, d_block_ptr: int):
        self._ck(self._lib.avb_process_frame_device(self._h, C.c_void_p(d_block_ptr)))

    def enqueue_device(self, d_block_ptr: int):
        self._ck(self._lib.avb_enqueue_frame_device(self._h, C.c_void_p(d_block_ptr)))

    def sync(self):
        self._ck(self._lib.avb_sync(self._h))

    def result(self, s=0):
        """(header record, ids int64[n], meas float64[n,4]) of the last frame -- views into pinned memory, valid until the
        frame after next (the result blocks alternate by frame parity)."""
        hp, ip_, mp = C.c_void_p(), C.c_void_p(), C.c_void_p()
        self._ck(self._lib.avb_get_result(self._h, s, C.byref(hp), C.byref(ip_), C.byref(mp)))
        v = self._views.get(hp.value)
        if v is None:
            hdr = np.ctypeslib.as_array(C.cast(hp, C.POINTER(C.c_uint8)), shape=(48,)).view(HEADER_DTYPE)
            ids = np.ctypeslib.as_array(C.cast(ip_, C.POINTER(C.c_int64)), shape=(self.capacity,))
            meas = np.ctypeslib.as_array(C.cast(mp, C.POINTER(C.c_double)), shape=(self.capacity, 4))
            v = (hdr, ids, meas)
            self._views[hp.value] = v
        hdr, ids, meas = v
        n = int(hdr['n_features'][0])
        return hdr[0], ids[:n], meas[:n]

    def result_block(self, prev=False):
        """The result blocks of all S streams as one array: (block uint8[S, stride] -- a view of the pinned memory --,
        ids offset, meas offset).  Stream s: header at block[s, :48] (HEADER_DTYPE), ids int64[capacity] at the ids
        offset, meas float64[capacity, 4] at the meas offset.  prev=True: the frame before the last one enqueued (what a
        driver reads while the next frame runs, avb_get_result_prev)."""
        fn = self._lib.avb_get_result_prev if prev else self._lib.avb_get_result

        def ptrs(s):
            hp, ip_, mp = C.c_void_p(), C.c_void_p(), C.c_void_p()
            self._ck(fn(self._h, s, C.byref(hp), C.byref(ip_), C.byref(mp)))
            return hp.value, ip_.value, mp.value
        h0, i0, m0 = ptrs(0)
        blk = self._blocks.get(h0)
        if blk is None:
            stride = (ptrs(1)[0] - h0) if self.S > 1 else (m0 - h0) + self.capacity * 32
            a = np.ctypeslib.as_array(C.cast(C.c_void_p(h0), C.POINTER(C.c_uint8)), shape=(self.S * stride,))
            blk = (a.reshape(self.S, stride), i0 - h0, m0 - h0)
            self._blocks[h0] = blk
        return blk

    def features(self, s=0):
        """Grid-ordered state of stream s: (cell, lifetime, cam0_xy, cam1_xy)."""
        n = int(self.result(s)[0]['n_features'])
        cell = np.empty(n, np.int32)
        life = np.empty(n, np.int32)
        p0 = np.empty((n, 2), np.float32)
        p1 = np.empty((n, 2), np.float32)
        self._ck(self._lib.avb_get_features(self._h, s, _ptr(cell), _ptr(life), _ptr(p0), _ptr(p1)))
        return cell, life, p0, p1

    def last_frame_ms(self):
        ms = C.c_float()
        self._ck(self._lib.avb_last_frame_ms(self._h, C.byref(ms)))
        return ms.value

    def kernels_per_frame(self):
        return self._lib.avb_kernels_per_frame(self._h)

    def cuda_stream(self):
        return self._lib.avb_cuda_stream(self._h)

    STAGES = ('input_copy', 'fast', 'pyramid', 'track', 'select', 'stereo_new', 'finish', 'spec_match',
              'result_copy')

    def profile_frame_device(self, d_block_ptr: int):
        ms = (C.c_float * 9)()
        self._ck(self._lib.avb_profile_frame_device(self._h, C.c_void_p(d_block_ptr), ms))
        return dict(zip(self.STAGES, [float(v) for v in ms]))

    def time_pyramid(self, iters=20):
        ms = C.c_float()
        self._ck(self._lib.avb_time_pyramid(self._h, iters, C.byref(ms)))
        return ms.value

    # -- per-stage entry points -----------------------------------------------------------------------------
    def upload(self, img0, img1, s=0):
        img0 = np.ascontiguousarray(img0, dtype=np.uint8)
        img1 = np.ascontiguousarray(img1, dtype=np.uint8)
        if img0.shape != (self.height, self.width) or img1.shape != (self.height, self.width):
            raise RuntimeError(f'image shape {img0.shape} does not match the context ({self.height}, {self.width})')
        self._ck(self._lib.avb_upload_stereo(self._h, s, _ptr(img0), _ptr(img1), self.width))

    def advance(self):
        self._ck(self._lib.avb_advance(self._h))

    def begin_frame(self, img0, img1, s=0):
        """Stage-class flow (PyramidBuilder.create_image_pyramids): the frame that was current becomes the previous
        one (its device pyramids stay), the new images are uploaded and their pyramids built."""
        if self._staged_frames:
            self.advance()
        self.upload(img0, img1, s)
        self.build_pyramids()
        self._staged_frames += 1
        self._cur0 = img0

    def ensure_current_cam0(self, img):
        """FAST runs on the current cam0 image of the context; a detector handed a different image uploads it first."""
        if img is self._cur0:
            return
        if self._cur0 is not None and img.shape == self._cur0.shape and np.array_equal(img, self._cur0):
            return
        self.upload(img, img)
        self._cur0 = img

    def build_pyramids(self):
        self._ck(self._lib.avb_build_pyramids(self._h))

    def download_level(self, slot, level, s=0):
        w, h = C.c_int(), C.c_int()
        self._ck(self._lib.avb_download_level(self._h, s, slot, level, None, C.byref(w), C.byref(h)))
        out = np.empty((h.value, w.value), np.uint8)
        self._ck(self._lib.avb_download_level(self._h, s, slot, level, _ptr(out), C.byref(w), C.byref(h)))
        return out

    def fast_detect(self, mask=None, s=0):
        cap = (self.width * self.height) // 4 + 16
        xs, ys, rs = np.empty(cap, np.int32), np.empty(cap, np.int32), np.empty(cap, np.int32)
        n = C.c_int(cap)
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        self._ck(self._lib.avb_fast_detect(self._h, s, _ptr(m), _ptr(xs), _ptr(ys), _ptr(rs), C.byref(n)))
        return xs[:n.value].copy(), ys[:n.value].copy(), rs[:n.value].copy()

    def klt_track(self, slot_from, slot_to, prev_xy, guess_xy, s=0):
        prev = np.ascontiguousarray(prev_xy, dtype=np.float32).reshape(-1, 2)
        guess = np.ascontiguousarray(guess_xy, dtype=np.float32).reshape(-1, 2)
        n = len(prev)
        out = np.zeros((n, 2), np.float32)
        st = np.zeros(n, np.uint8)
        self._ck(self._lib.avb_klt_track(self._h, s, slot_from, slot_to, _ptr(prev), _ptr(guess), n, _ptr(out), _ptr(st)))
        return out, st

    def stereo_match(self, cam0_xy, s=0):
        p0 = np.ascontiguousarray(cam0_xy, dtype=np.float32).reshape(-1, 2)
        n = len(p0)
        p1 = np.zeros((n, 2), np.float32)
        ok = np.zeros(n, np.uint8)
        self._ck(self._lib.avb_stereo_match(self._h, s, _ptr(p0), n, _ptr(p1), _ptr(ok)))
        return p1, ok.astype(bool)

    def undistort_ex(self, intrinsics, distortion, xy, R=None, f32_io=False):
        p = np.ascontiguousarray(xy, dtype=np.float64).reshape(-1, 2)
        out = np.zeros_like(p)
        K = np.ascontiguousarray(intrinsics, dtype=np.float64)
        D = np.ascontiguousarray(np.asarray(distortion, dtype=np.float64)[:4])
        Rm = None if R is None else np.ascontiguousarray(R, dtype=np.float64).reshape(-1)
        self._ck(self._lib.avb_undistort_points(self._h, _ptr(K), _ptr(D), _ptr(p), len(p), _ptr(Rm), int(f32_io), _ptr(out)))
        return out

    def distort_ex(self, intrinsics, distortion, xy, f32_io=False):
        p = np.ascontiguousarray(xy, dtype=np.float64).reshape(-1, 2)
        out = np.zeros_like(p)
        K = np.ascontiguousarray(intrinsics, dtype=np.float64)
        D = np.ascontiguousarray(np.asarray(distortion, dtype=np.float64)[:4])
        self._ck(self._lib.avb_distort_points(self._h, _ptr(K), _ptr(D), _ptr(p), len(p), int(f32_io), _ptr(out)))
        return out

    def two_point_ransac(self, intrinsics, distortion, prev_xy, cur_xy, R_p_c=None, threshold=3.0, seed=0,
                         frame_index=0, cam=0):
        """Inlier mask of the two-point RANSAC between two frames of one camera (k_ransac; not a reference stage:
        feature_tracker.py:135-136 is an all-ones stub).  prev_xy / cur_xy: (N, 2) pixel positions."""
        a = np.ascontiguousarray(prev_xy, dtype=np.float32).reshape(-1, 2)
        b = np.ascontiguousarray(cur_xy, dtype=np.float32).reshape(-1, 2)
        if len(a) != len(b):
            raise ValueError('prev_xy and cur_xy differ in length')
        out = np.zeros(len(a), dtype=np.uint8)
        K = np.ascontiguousarray(intrinsics, dtype=np.float64)
        D = np.ascontiguousarray(np.asarray(distortion, dtype=np.float64)[:4])
        R = None if R_p_c is None else np.ascontiguousarray(R_p_c, dtype=np.float64).reshape(-1)
        self._ck(self._lib.avb_two_point_ransac(self._h, _ptr(K), _ptr(D), _ptr(a), _ptr(b), len(a), _ptr(R),
                                                float(threshold), int(seed), int(frame_index), int(cam), _ptr(out)))
        return out.astype(bool)
