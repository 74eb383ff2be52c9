"""ORACLE (test infrastructure, never shipped, never on the product path).

CPU restatement of the reference's stereo front end, `ImageProcessingPipeline.stereo_callback`
(/root/reference/src/image_processing/pipeline.py:46-150) and the stage classes it drives.
It is written array-style (one flat feature table in grid order instead of lists of
objects) but reproduces every order-sensitive rule and quirk listed in SURVEY.md
Appendix B; tests/test_oracle_pipeline_port.py pins it frame by frame against golden
dumps of the real reference (tests/golden/, made by tools/make_golden.py which imports
/root/reference/src in the build container).

Backends for the OpenCV arithmetic:
  backend='cv2'    call cv2 exactly where the reference does (this is also the timed
                   "port" CPU baseline of bench.py: same library, same call pattern)
  backend='numpy'  oracle/cv_semantics.py (slow; proves the restated semantics close the
                   loop without cv2)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this.
"""
from __future__ import annotations

from collections import namedtuple

import numpy as np

from . import cv_semantics as cs
from . import ransac as rs

feature_msg = namedtuple('feature_msg', ['timestamp', 'features'])
PortFeature = namedtuple('PortFeature', ['id', 'u0', 'v0', 'u1', 'v1'])


def rodrigues(v):
    """cv2.Rodrigues(v)[0] for a 3-vector (imu_processor.py:63-64)."""
    v = np.asarray(v, dtype=np.float64).reshape(3)
    theta = float(np.sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]))
    if theta < np.finfo(np.float64).eps:
        return np.eye(3)
    r = v / theta
    c, s = np.cos(theta), np.sin(theta)
    c1 = 1.0 - c
    rrt = np.outer(r, r)
    rx = np.array([[0, -r[2], r[1]], [r[2], 0, -r[0]], [-r[1], r[0], 0]])
    return c * np.eye(3) + c1 * rrt + s * rx


class _CvBackend:
    name = 'cv2'

    def __init__(self, cfg):
        import cv2
        self.cv2 = cv2
        self.det = cv2.FastFeatureDetector_create(cfg.fast_threshold)
        self.lk = dict(cfg.lk_params)

    def fast(self, img, mask=None):
        kps = self.det.detect(img, mask=mask) if mask is not None else self.det.detect(img)
        n = len(kps)
        xy = np.empty((n, 2), dtype=np.float64)
        rs = np.empty(n, dtype=np.float64)
        for i, k in enumerate(kps):
            xy[i] = k.pt
            rs[i] = k.response
        return xy, rs

    def lk_track(self, img_a, img_b, pts, guess):
        q, st, _ = self.cv2.calcOpticalFlowPyrLK(img_a, img_b, pts, guess, **self.lk)
        return q, st.reshape(-1)

    def undistort(self, pts, intr, dist, R=None):
        K = np.array([[intr[0], 0, intr[2]], [0, intr[1], intr[3]], [0, 0, 1.0]])
        R = np.eye(3) if R is None else R
        return self.cv2.undistortPoints(np.reshape(pts, (-1, 1, 2)), K, np.asarray(dist),
                                        None, R, np.eye(3)).reshape(-1, 2)

    def distort(self, pts, intr, dist):
        K = np.array([[intr[0], 0, intr[2]], [0, intr[1], intr[3]], [0, 0, 1.0]])
        h = self.cv2.convertPointsToHomogeneous(pts)
        out, _ = self.cv2.projectPoints(h, np.zeros(3), np.zeros(3), K, np.asarray(dist))
        return out.reshape(-1, 2)


class _NumpyBackend:
    name = 'numpy'

    def __init__(self, cfg):
        self.thr = cfg.fast_threshold
        self.levels = cfg.lk_params['maxLevel']
        self.win = cfg.lk_params['winSize'][0]
        crit = cfg.lk_params['criteria']
        self.max_iter, self.eps = crit[1], crit[2]
        self._pyr_cache = {}

    def _pyr(self, img):
        key = id(img)
        hit = self._pyr_cache.get(key)
        if hit is None or hit[0] is not img:
            pyr = cs.build_pyramid(img, self.levels)
            hit = (img, pyr, [cs.scharr(l) for l in pyr])
            if len(self._pyr_cache) > 8:
                self._pyr_cache.clear()
            self._pyr_cache[key] = hit
        return hit[1], hit[2]

    def fast(self, img, mask=None):
        xs, ys, rs = cs.fast_detect(img, self.thr, mask)
        return np.stack([xs, ys], 1).astype(np.float64), rs.astype(np.float64)

    def lk_track(self, img_a, img_b, pts, guess):
        pa, da = self._pyr(img_a)
        pb, _ = self._pyr(img_b)
        return cs.lk_track(pa, pb, pts, guess, win=self.win, max_iter=self.max_iter,
                           eps=self.eps, derivs=da)

    def undistort(self, pts, intr, dist, R=None):
        return cs.undistort_radtan(np.asarray(pts), intr, dist, R)

    def distort(self, pts, intr, dist):
        return cs.distort_radtan(np.asarray(pts), intr, dist)


def _cells(p0, gh, gw, cols):
    # int(y / grid_h) * grid_col + int(x / grid_w)   (feature_tracker.py:144-146)
    return (p0[:, 1] / gh).astype(np.int64) * cols + (p0[:, 0] / gw).astype(np.int64)


class FrontEndPort:
    """Drop-in shaped like ImageProcessor: imu_callback / stereo_callback -> feature_msg.

    Feature table (grid order = cell-major, list order inside a cell) kept as arrays:
    ids, life, cell, p0 (cam0 xy), p1 (cam1 xy), fresh (1 = created this frame)."""

    def __init__(self, cfg, backend='cv2', tap=None):
        self.cfg = cfg
        self.be = _CvBackend(cfg) if backend == 'cv2' else _NumpyBackend(cfg)
        self.tap = tap
        T0 = np.linalg.inv(cfg.T_imu_cam0)
        T1 = np.linalg.inv(cfg.T_imu_cam1)
        self.R_cam0_imu, self.t_cam0_imu = T0[:3, :3], T0[:3, 3]
        self.R_cam1_imu, self.t_cam1_imu = T1[:3, :3], T1[:3, 3]
        self.R0to1 = self.R_cam1_imu.T @ self.R_cam0_imu
        t01 = self.R_cam1_imu.T @ (self.t_cam0_imu - self.t_cam1_imu)
        tx = np.array([[0, -t01[2], t01[1]], [t01[2], 0, -t01[0]], [-t01[1], t01[0], 0]])
        self.E = tx @ self.R0to1
        self.imu_buffer = []
        self.prev_img0 = None
        self.prev_ts = None
        self.next_feature_id = 0
        self.first_frame = True
        self.frame_index = 0
        self.num_features = {}
        self._set_table(*self._empty())

    # -- feature table -----------------------------------------------------------------
    @staticmethod
    def _empty():
        return (np.zeros(0, np.int64), np.zeros(0, np.int64), np.zeros(0, np.int64),
                np.zeros((0, 2), np.float32), np.zeros((0, 2), np.float32), np.zeros(0, bool))

    def _set_table(self, ids, life, cell, p0, p1, fresh):
        self.ids, self.life, self.cell, self.p0, self.p1, self.fresh = ids, life, cell, p0, p1, fresh

    # -- IMU (imu_processor.py:22-67) --------------------------------------------------------
    def imu_callback(self, msg):
        self.imu_buffer.append(msg)

    def _integrate_imu(self, t_prev, t_curr):
        buf = self.imu_buffer
        ib = next((i for i, m in enumerate(buf) if m.timestamp >= t_prev - 0.01), None)
        ie = next((i for i, m in enumerate(buf) if m.timestamp >= t_curr - 0.004), None)
        if ib is None or ie is None:
            return np.eye(3), np.eye(3)
        w = np.zeros(3)
        for m in buf[ib:ie]:
            w += m.angular_velocity
        if ie - ib > 0:
            w /= (ie - ib)
        dt = t_curr - t_prev
        R0 = rodrigues((self.R_cam0_imu.T @ w) * dt).T
        R1 = rodrigues((self.R_cam1_imu.T @ w) * dt).T
        self.imu_buffer = buf[ie:]
        return R0, R1

    # -- stereo matching (stereo_matcher.py:33-115) ------------------------------------------
    def stereo_match(self, img0, img1, pts0):
        cfg, be = self.cfg, self.be
        if len(pts0) == 0:
            return np.zeros((0, 2), np.float32), np.zeros(0, bool)
        pts0 = np.asarray(pts0, dtype=np.float32).reshape(-1, 2)
        K, D = cfg.cam0_intrinsics, cfg.cam0_distortion_coeffs     # cam0 model both ways (B3)
        proj1 = be.distort(be.undistort(pts0, K, D, self.R0to1), K, D)
        p1, st_f = be.lk_track(img0, img1, pts0, np.array(proj1, dtype=np.float32))
        p0r, _ = be.lk_track(img1, img0, p1, pts0.copy())           # backward status ignored (B5)
        err = np.linalg.norm(pts0 - p0r, axis=1)
        disp = np.abs(proj1[:, 1] - p1[:, 1])
        ok = st_f.astype(bool) & (err < 3) & (disp < 20)
        h, w = img1.shape[:2]
        ok &= ~((p1[:, 0] < 0) | (p1[:, 0] >= w) | (p1[:, 1] < 0) | (p1[:, 1] >= h))
        u0 = be.undistort(pts0, K, D)
        u1 = be.undistort(p1, K, D)
        unit = 4.0 / (2 * K[0] + 2 * K[1])
        h0 = np.concatenate([u0.astype(np.float64), np.ones((len(u0), 1))], axis=1)
        line = h0 @ self.E.T
        # only the x-term of the element-wise product survives in the reference (B4)
        epi = np.abs(u1[:, 0].astype(np.float64) * line[:, 0]) / np.sqrt(
            line[:, 0] ** 2 + line[:, 1] ** 2)
        ok &= ~(epi > cfg.stereo_threshold * unit)
        if self.tap is not None:
            self.tap('stereo_match', dict(pts0=pts0.copy(), proj1=np.array(proj1), p1=p1.copy(),
                                          st_f=st_f.copy(), p0r=p0r.copy(), ok=ok.copy()))
        return p1, ok

    # -- frame 0 (feature_initializer.py:45-85) -----------------------------------------------
    def _initialize(self, img0, img1, gh, gw):
        cfg = self.cfg
        xy, resp = self.be.fast(img0)
        if self.tap is not None:
            self.tap('fast', dict(xy=xy.copy(), resp=resp.copy()))
        p1, ok = self.stereo_match(img0, img1, xy)
        xy, resp, p1 = xy[ok], resp[ok], p1[ok]
        cell = _cells(xy, gh, gw, cfg.grid_col) if len(xy) else np.zeros(0, np.int64)
        rows = []
        for c in range(cfg.grid_num):
            m = np.nonzero(cell == c)[0]
            m = m[np.argsort(-resp[m], kind='stable')][:cfg.grid_min_feature_num]
            rows.append(m)
        sel = np.concatenate(rows) if rows else np.zeros(0, np.int64)
        n = len(sel)
        ids = self.next_feature_id + np.arange(n, dtype=np.int64)
        self.next_feature_id += n
        self._set_table(ids, np.ones(n, np.int64), cell[sel], xy[sel].astype(np.float32),
                        p1[sel].astype(np.float32), np.ones(n, bool))

    # -- frame k>0: tracker (feature_tracker.py:74-157) ------------------------------------------
    def _track(self, img0, img1, R_p_c, gh, gw, R_p_c1=None):
        cfg = self.cfg
        nf = self.num_features
        nf['before_tracking'] = len(self.ids)
        if len(self.ids) == 0:
            self._set_table(*self._empty())
            return
        K = cfg.cam0_intrinsics
        Km = np.array([[K[0], 0, K[2]], [0, K[1], K[3]], [0, 0, 1.0]])
        H = Km @ R_p_c @ np.linalg.inv(Km)
        ph = np.concatenate([self.p0.astype(np.float64), np.ones((len(self.p0), 1))], axis=1) @ H.T
        pred = (ph[:, :2] / ph[:, 2:3]).astype(np.float32)
        cur, st = self.be.lk_track(self.prev_img0, img0, self.p0, pred)
        h, w = img0.shape[:2]
        keep = st.astype(bool) & ~((cur[:, 0] < 0) | (cur[:, 0] > w - 1) |
                                   (cur[:, 1] < 0) | (cur[:, 1] > h - 1))
        if self.tap is not None:
            self.tap('temporal', dict(prev=self.p0.copy(), pred=pred.copy(), cur=cur.copy(),
                                      st=st.copy(), keep=keep.copy()))
        ids, life, cur = self.ids[keep], self.life[keep], cur[keep]
        prev0, prev1 = self.p0[keep], self.p1[keep]
        nf['after_tracking'] = len(cur)
        p1, ok = self.stereo_match(img0, img1, cur)
        ids, life, cur, p1, prev0, prev1 = ids[ok], life[ok], cur[ok], p1[ok], prev0[ok], prev1[ok]
        nf['after_matching'] = len(cur)
        if getattr(cfg, 'two_point_ransac', False) and len(cur):
            # NOT in the reference (all-ones stub, B2): oracle/ransac.py, applied per camera, survivors need both
            be, seed = self.be, int(getattr(cfg, 'ransac_seed', 0))
            inl = np.ones(len(cur), bool)
            for cam, (a, b, K, D, R) in enumerate((
                    (prev0, cur, cfg.cam0_intrinsics, cfg.cam0_distortion_coeffs, R_p_c),
                    (prev1, p1, cfg.cam1_intrinsics, cfg.cam1_distortion_coeffs, R_p_c1))):
                u_prev = be.undistort(np.asarray(a, np.float32), K, D, R)
                u_cur = be.undistort(np.asarray(b, np.float32), K, D)
                m = rs.two_point_ransac(u_prev, u_cur, K, cfg.ransac_threshold, seed=seed,
                                        frame_index=self.frame_index, cam=cam)
                if self.tap is not None:
                    self.tap('ransac', dict(cam=cam, prev=np.array(a), cur=np.array(b), R=np.array(R), mask=m.copy(),
                                            frame_index=self.frame_index))
                inl &= m
            ids, life, cur, p1 = ids[inl], life[inl], cur[inl], p1[inl]
        nf['after_ransac'] = len(cur)                      # reference: all-ones stub (B2) -> same as after_matching
        cell = _cells(cur, gh, gw, cfg.grid_col) if len(cur) else np.zeros(0, np.int64)
        order = np.argsort(cell, kind='stable')            # per-cell append in list order
        self._set_table(ids[order], life[order] + 1, cell[order], cur[order], p1[order],
                        np.zeros(len(order), bool))

    # -- frame k>0: adder (feature_adder.py:52-108) -----------------------------------------------
    def _add(self, img0, img1, gh, gw):
        cfg = self.cfg
        h, w = img0.shape[:2]
        mask = np.ones((h, w), dtype=np.uint8)
        for x, y in self.p0.astype(np.int64):              # int() truncation (B7)
            if x < 3 or y < 3:
                continue                                   # negative slice start -> empty slice
            mask[y - 3:y + 4, x - 3:x + 4] = 0
        xy, resp = self.be.fast(img0, mask)
        if self.tap is not None:
            self.tap('fast', dict(xy=xy.copy(), resp=resp.copy()))
        cell = _cells(xy, gh, gw, cfg.grid_col) if len(xy) else np.zeros(0, np.int64)
        cand = []
        for c in range(cfg.grid_num):
            m = np.nonzero(cell == c)[0]
            if len(m) > cfg.grid_max_feature_num:
                m = m[np.argsort(-resp[m], kind='stable')][:cfg.grid_max_feature_num]
            cand.append(m)
        cand = np.concatenate(cand) if cand else np.zeros(0, np.int64)
        xy, resp = xy[cand], resp[cand]
        p1, ok = self.stereo_match(img0, img1, xy)
        xy, resp, p1 = xy[ok], resp[ok], p1[ok]
        cell = _cells(xy, gh, gw, cfg.grid_col) if len(xy) else np.zeros(0, np.int64)
        parts = []
        for c in range(cfg.grid_num):
            old = np.nonzero(self.cell == c)[0]
            m = np.nonzero(cell == c)[0]
            m = m[np.argsort(-resp[m], kind='stable')][:cfg.grid_min_feature_num]   # B8
            k = len(m)
            new_ids = self.next_feature_id + np.arange(k, dtype=np.int64)
            self.next_feature_id += k
            parts.append((np.concatenate([self.ids[old], new_ids]),
                          np.concatenate([self.life[old], np.ones(k, np.int64)]),
                          np.full(len(old) + k, c, np.int64),
                          np.concatenate([self.p0[old], xy[m].astype(np.float32)]),
                          np.concatenate([self.p1[old], p1[m].astype(np.float32)]),
                          np.concatenate([np.zeros(len(old), bool), np.ones(k, bool)])))
        self._set_table(*[np.concatenate([p[i] for p in parts]) for i in range(6)])

    # -- frame k>0: pruner (feature_pruner.py:8-19) ------------------------------------------------
    def _prune(self):
        cfg = self.cfg
        keep = []
        for c in range(cfg.grid_num):
            m = np.nonzero(self.cell == c)[0]
            if len(m) > cfg.grid_max_feature_num:
                m = m[np.argsort(-self.life[m], kind='stable')][:cfg.grid_max_feature_num]
            keep.append(m)
        keep = np.concatenate(keep) if keep else np.zeros(0, np.int64)
        self._set_table(self.ids[keep], self.life[keep], self.cell[keep], self.p0[keep],
                        self.p1[keep], self.fresh[keep])

    # -- publish (feature_publisher.py:90-121) --------------------------------------------------------
    def _publish(self, ts):
        cfg = self.cfg
        if len(self.ids) == 0:
            return feature_msg(ts, [])
        # u0,v0 come out f64 when the frame holds any new (tuple-typed) point, else f32 (B11)
        p0 = self.p0.astype(np.float64) if self.fresh.any() else self.p0
        u0 = self.be.undistort(p0, cfg.cam0_intrinsics, cfg.cam0_distortion_coeffs)
        u1 = self.be.undistort(self.p1, cfg.cam1_intrinsics, cfg.cam1_distortion_coeffs)
        feats = [PortFeature(int(i), a[0], a[1], b[0], b[1]) for i, a, b in zip(self.ids, u0, u1)]
        return feature_msg(ts, feats)

    def stereo_callback(self, msg):
        cfg = self.cfg
        img0, img1 = msg.cam0_msg.image, msg.cam1_msg.image
        ts = msg.cam0_msg.timestamp
        h, w = img0.shape[:2]
        gh = int(np.ceil(h / cfg.grid_row))                 # B12
        gw = int(np.ceil(w / cfg.grid_col))
        if self.first_frame:
            self._initialize(img0, img1, gh, gw)
            self.first_frame = False
        else:
            R0, R1 = self._integrate_imu(self.prev_ts, ts)
            self._track(img0, img1, R0, gh, gw, R1)
            self._add(img0, img1, gh, gw)
            self._prune()
        out = self._publish(ts)
        if self.tap is not None:
            self.tap('grid', dict(ids=self.ids.copy(), life=self.life.copy(), cell=self.cell.copy(),
                                  p0=self.p0.copy(), p1=self.p1.copy(), fresh=self.fresh.copy()))
        self.prev_img0, self.prev_ts = img0, ts
        self.frame_index += 1
        return out

    stareo_callback = stereo_callback
