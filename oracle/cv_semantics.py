"""ORACLE (test infrastructure, never shipped, never on the product path).

CPU restatement, in numpy, of the third-party arithmetic the reference front end
delegates to OpenCV (`opencv-python`, unpinned in /root/reference/requirements.txt:2-3;
pinned here to cv2 4.13.0, the version in this image).  The reference call sites are

  * cv2.FastFeatureDetector_create(15).detect   src/image_processing/pipeline.py:23-25,
                                                feature_initializer.py:52, feature_adder.py:64
  * cv2.calcOpticalFlowPyrLK                    feature_tracker.py:102-108,
                                                stereo_matcher.py:64-68, 70-74
  * cv2.undistortPoints / cv2.projectPoints     camera_model.py:45, 72-74,
                                                feature_publisher.py:57

OpenCV's sources are not under /root/reference, so every routine below restates the
published algorithm (OpenCV `fast.cpp`, `pyramids.cpp`, `lkpyramid.cpp`, `undistort`)
and is pinned against cv2 4.13.0 itself by tests/test_oracle_cv_semantics.py.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this.

One deliberate, documented deviation: cv2 accumulates the LK structure tensor and the
residual vector in float32 in a SIMD-lane dependent order; this restatement (and the
CUDA kernels) form the exact integer sums and round once to float32.  The observed
difference against cv2 is < 1e-3 px (tests assert it), inside the 0.01 px budget.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32

# --------------------------------------------------------------------------------------
# border helper
# --------------------------------------------------------------------------------------

def reflect101(i, n):
    """BORDER_REFLECT_101 index map (…2 1 | 0 1 2 … n-1 | n-2 n-3 …), single reflection."""
    i = np.asarray(i)
    i = np.where(i < 0, -i, i)
    i = np.where(i >= n, 2 * n - 2 - i, i)
    return i


# --------------------------------------------------------------------------------------
# A.2  pyrDown  (levels 1..maxLevel of the LK pyramid)
# --------------------------------------------------------------------------------------

_K5 = np.array([1, 4, 6, 4, 1], dtype=np.int32)


def pyr_down(img: np.ndarray) -> np.ndarray:
    """cv2.pyrDown for uint8: 5x5 binomial, REFLECT_101, (sum + 128) >> 8, size (n+1)//2."""
    assert img.dtype == np.uint8 and img.ndim == 2
    h, w = img.shape
    dh, dw = (h + 1) // 2, (w + 1) // 2
    src = img.astype(np.int32)
    # horizontal pass at the even columns
    cols = 2 * np.arange(dw)
    hsum = np.zeros((h, dw), dtype=np.int32)
    for t in range(5):
        hsum += _K5[t] * src[:, reflect101(cols + t - 2, w)]
    rows = 2 * np.arange(dh)
    out = np.zeros((dh, dw), dtype=np.int32)
    for t in range(5):
        out += _K5[t] * hsum[reflect101(rows + t - 2, h), :]
    return ((out + 128) >> 8).astype(np.uint8)


def build_pyramid(img: np.ndarray, max_level: int):
    """Levels 0..max_level as cv2.buildOpticalFlowPyramid makes them (no early stop needed
    at the BASELINE sizes; asserted)."""
    pyr = [np.ascontiguousarray(img)]
    for _ in range(max_level):
        nxt = pyr_down(pyr[-1])
        pyr.append(nxt)
    return pyr


# --------------------------------------------------------------------------------------
# A.3  Scharr derivative (int16, unnormalised), REFLECT_101 at the level's edge
# --------------------------------------------------------------------------------------

def scharr(img: np.ndarray):
    """Returns (dx, dy) int16 = cv2.Scharr(img, CV_16S, 1, 0) / (0, 1), border REFLECT_101."""
    h, w = img.shape
    s = img.astype(np.int32)
    ys = np.arange(h)
    xs = np.arange(w)
    up = s[reflect101(ys - 1, h), :]
    dn = s[reflect101(ys + 1, h), :]
    sm = 3 * up + 10 * s + 3 * dn        # vertical smoothing  -> for dx
    df = dn - up                         # vertical difference -> for dy
    xl = reflect101(xs - 1, w)
    xr = reflect101(xs + 1, w)
    dx = sm[:, xr] - sm[:, xl]
    dy = 3 * df[:, xl] + 10 * df + 3 * df[:, xr]
    return dx.astype(np.int16), dy.astype(np.int16)


# --------------------------------------------------------------------------------------
# A.1  FAST-9/16 with 3x3 non-maximum suppression
# --------------------------------------------------------------------------------------

FAST_RING = ((0, 3), (1, 3), (2, 2), (3, 1), (3, 0), (3, -1), (2, -2), (1, -3),
             (0, -3), (-1, -3), (-2, -2), (-3, -1), (-3, 0), (-3, 1), (-2, 2), (-1, 3))


def fast_score_map(img: np.ndarray, threshold: int) -> np.ndarray:
    """Per-pixel response as cv2 reports it (0 where the pixel is not a FAST-9 corner).

    best = max over the 16 arcs of 9 contiguous ring pixels of min(ring - c) (bright) or
    min(c - ring) (dark); corner <=> best > threshold; response = best - 1.
    """
    h, w = img.shape
    s = img.astype(np.int16)
    score = np.zeros((h, w), dtype=np.int32)
    if h < 7 or w < 7:
        return score
    c = s[3:h - 3, 3:w - 3]
    d = np.stack([s[3 + dy:h - 3 + dy, 3 + dx:w - 3 + dx] - c for dx, dy in FAST_RING])
    best = np.zeros(c.shape, dtype=np.int16)
    for sign in (1, -1):
        dd = d * sign
        for k in range(16):
            m = dd[k]
            for j in range(1, 9):
                m = np.minimum(m, dd[(k + j) % 16])
            best = np.maximum(best, m)
    inner = np.where(best > threshold, best.astype(np.int32) - 1, 0)
    score[3:h - 3, 3:w - 3] = inner
    return score


def fast_detect(img: np.ndarray, threshold: int, mask: np.ndarray | None = None):
    """cv2.FastFeatureDetector_create(threshold).detect(img, mask): TYPE_9_16, NMS on.

    Returns (xs, ys, responses) int arrays in row-major scan order.  NMS keeps a corner iff
    its response is strictly greater than all 8 neighbours' responses (non-corners = 0);
    the mask is a post-filter (NMS runs first)."""
    sc = fast_score_map(img, threshold)
    h, w = sc.shape
    p = np.zeros((h + 2, w + 2), dtype=np.int32)
    p[1:-1, 1:-1] = sc
    keep = sc > 0
    for dy in (-1, 0, 1):
        for dx in (-1, 0, 1):
            if dx == 0 and dy == 0:
                continue
            keep &= sc > p[1 + dy:1 + dy + h, 1 + dx:1 + dx + w]
    if mask is not None:
        keep &= mask != 0
    ys, xs = np.nonzero(keep)           # row-major
    return xs.astype(np.int32), ys.astype(np.int32), sc[ys, xs].astype(np.int32)


# --------------------------------------------------------------------------------------
# A.4  pyramidal Lucas-Kanade  (cv2.calcOpticalFlowPyrLK, OPTFLOW_USE_INITIAL_FLOW)
# --------------------------------------------------------------------------------------

W_BITS = 14
FLT_SCALE = F32(1.0 / (1 << 20))
FLT_EPSILON = F32(1.1920929e-07)


def _rint(x):
    return int(np.rint(x))              # round-half-even, like cvRound


def _weights(a: np.float32, b: np.float32):
    one = F32(1.0)
    s = F32(1 << W_BITS)
    iw00 = _rint(F32(F32(one - a) * F32(one - b)) * s)
    iw01 = _rint(F32(a * F32(one - b)) * s)
    iw10 = _rint(F32(F32(one - a) * b) * s)
    iw11 = (1 << W_BITS) - iw00 - iw01 - iw10
    return iw00, iw01, iw10, iw11


def _patch_u8(img: np.ndarray, ix: int, iy: int, n: int):
    """(n+1)x(n+1) intensities starting at (ix, iy), REFLECT_101 outside the level."""
    h, w = img.shape
    ys = reflect101(iy + np.arange(n + 1), h)
    xs = reflect101(ix + np.arange(n + 1), w)
    return img[np.ix_(ys, xs)].astype(np.int64)


def _patch_deriv(d: np.ndarray, ix: int, iy: int, n: int):
    """(n+1)x(n+1) derivative samples; ZERO outside the level (constant border)."""
    h, w = d.shape
    ys = iy + np.arange(n + 1)
    xs = ix + np.arange(n + 1)
    oky = (ys >= 0) & (ys < h)
    okx = (xs >= 0) & (xs < w)
    out = d[np.ix_(np.clip(ys, 0, h - 1), np.clip(xs, 0, w - 1))].astype(np.int64)
    out[~oky, :] = 0
    out[:, ~okx] = 0
    return out


def _bilin(p, iw, shift):
    iw00, iw01, iw10, iw11 = iw
    v = p[:-1, :-1] * iw00 + p[:-1, 1:] * iw01 + p[1:, :-1] * iw10 + p[1:, 1:] * iw11
    return (v + (1 << (shift - 1))) >> shift


def lk_track(prev_pyr, next_pyr, prev_pts, guess_pts, win: int = 15, max_iter: int = 30,
             eps: float = 0.01, min_eig_thr: float = 1e-4, derivs=None):
    """calcOpticalFlowPyrLK(prev, next, prev_pts, guess, winSize=(win,win),
    maxLevel=len(pyr)-1, criteria=(EPS|COUNT, max_iter, eps), flags=USE_INITIAL_FLOW).

    prev_pyr / next_pyr: lists of uint8 levels (build_pyramid).  Returns (pts f32 (N,2),
    status u8 (N,)).  Float32 arithmetic follows OpenCV's Point2f code statement by
    statement; sums are exact integers rounded once (see module docstring)."""
    prev_pts = np.asarray(prev_pts, dtype=F32).reshape(-1, 2)
    guess_pts = np.asarray(guess_pts, dtype=F32).reshape(-1, 2)
    n = len(prev_pts)
    max_level = len(prev_pyr) - 1
    if derivs is None:
        derivs = [scharr(l) for l in prev_pyr]
    max_iter = min(max(max_iter, 0), 100)
    eps = min(max(float(eps), 0.0), 10.0)
    eps2 = eps * eps                                   # double, like criteria.epsilon
    half = F32((win - 1) * 0.5)
    out = np.zeros((n, 2), dtype=F32)
    status = np.ones(n, dtype=np.uint8)
    for i in range(n):
        nx = ny = F32(0)
        for level in range(max_level, -1, -1):
            sc = F32(1.0 / (1 << level))
            I = prev_pyr[level]
            J = next_pyr[level]
            dxl, dyl = derivs[level]
            rows, cols = I.shape
            px = F32(prev_pts[i, 0] * sc)
            py = F32(prev_pts[i, 1] * sc)
            if level == max_level:
                nx = F32(guess_pts[i, 0] * sc)
                ny = F32(guess_pts[i, 1] * sc)
            else:
                nx = F32(nx * F32(2.0))
                ny = F32(ny * F32(2.0))
            # nextPts[i] = nextPt  (stored, un-shifted)
            px = F32(px - half)
            py = F32(py - half)
            ipx = int(np.floor(px))
            ipy = int(np.floor(py))
            if ipx < -win or ipx >= cols or ipy < -win or ipy >= rows:
                if level == 0:
                    status[i] = 0
                continue
            a = F32(px - F32(ipx))
            b = F32(py - F32(ipy))
            iw = _weights(a, b)
            Ip = _bilin(_patch_u8(I, ipx, ipy, win), iw, W_BITS - 5)
            Ix = _bilin(_patch_deriv(dxl, ipx, ipy, win), iw, W_BITS)
            Iy = _bilin(_patch_deriv(dyl, ipx, ipy, win), iw, W_BITS)
            A11 = F32(F32(int((Ix * Ix).sum())) * FLT_SCALE)
            A12 = F32(F32(int((Ix * Iy).sum())) * FLT_SCALE)
            A22 = F32(F32(int((Iy * Iy).sum())) * FLT_SCALE)
            D = F32(F32(A11 * A22) - F32(A12 * A12))
            dif = F32(A11 - A22)
            rad = F32(F32(dif * dif) + F32(F32(F32(4.0) * A12) * A12))
            min_eig = F32(F32(F32(A22 + A11) - F32(np.sqrt(rad))) / F32(2 * win * win))
            if float(min_eig) < float(min_eig_thr) or D < FLT_EPSILON:   # float vs double threshold
                if level == 0:
                    status[i] = 0
                continue
            D = F32(F32(1.0) / D)
            stored_x, stored_y = nx, ny                 # nextPts[i] as stored before the loop
            nx = F32(nx - half)
            ny = F32(ny - half)
            pdx = pdy = F32(0)
            for j in range(max_iter):
                inx = int(np.floor(nx))
                iny = int(np.floor(ny))
                if inx < -win or inx >= cols or iny < -win or iny >= rows:
                    if level == 0:
                        status[i] = 0
                    break
                a = F32(nx - F32(inx))
                b = F32(ny - F32(iny))
                iw = _weights(a, b)
                Jp = _bilin(_patch_u8(J, inx, iny, win), iw, W_BITS - 5)
                diff = Jp - Ip
                b1 = F32(F32(int((diff * Ix).sum())) * FLT_SCALE)
                b2 = F32(F32(int((diff * Iy).sum())) * FLT_SCALE)
                dx = F32(F32(F32(A12 * b2) - F32(A22 * b1)) * D)
                dy = F32(F32(F32(A12 * b1) - F32(A11 * b2)) * D)
                nx = F32(nx + dx)
                ny = F32(ny + dy)
                stored_x = F32(nx + half)
                stored_y = F32(ny + half)
                if float(dx) * float(dx) + float(dy) * float(dy) <= eps2:
                    break
                if j > 0 and abs(float(F32(dx + pdx))) < 0.01 and abs(float(F32(dy + pdy))) < 0.01:
                    stored_x = F32(stored_x - F32(dx * F32(0.5)))
                    stored_y = F32(stored_y - F32(dy * F32(0.5)))
                    break
                pdx, pdy = dx, dy
            nx, ny = stored_x, stored_y
        out[i, 0] = nx
        out[i, 1] = ny
        if status[i]:
            # the err block (the Python binding always asks for err) re-tests the window
            fx = int(np.floor(F32(nx - half)))
            fy = int(np.floor(F32(ny - half)))
            rows, cols = next_pyr[0].shape
            if fx < -win or fx >= cols or fy < -win or fy >= rows:
                status[i] = 0
    return out, status


# --------------------------------------------------------------------------------------
# A.5  radtan undistort / distort
# --------------------------------------------------------------------------------------

def undistort_radtan(pts, intr, dist, R=None):
    """cv2.undistortPoints(pts, K, D, None, R, I): 5 fixed-point iterations in f64, then R,
    then perspective divide.  Output dtype = input dtype (f32 in -> one rounding to f32)."""
    pts = np.asarray(pts)
    out_dtype = pts.dtype if pts.dtype in (np.float32, np.float64) else np.float64
    p = pts.astype(np.float64).reshape(-1, 2)
    fx, fy, cx, cy = (float(v) for v in intr)
    k1, k2, p1, p2 = (float(v) for v in dist[:4])
    ifx, ify = 1.0 / fx, 1.0 / fy
    x0 = (p[:, 0] - cx) * ifx
    y0 = (p[:, 1] - cy) * ify
    x, y = x0.copy(), y0.copy()
    for _ in range(5):
        r2 = x * x + y * y
        icdist = 1.0 / (1.0 + (k2 * r2 + k1) * r2)
        dx = 2 * p1 * x * y + p2 * (r2 + 2 * x * x)
        dy = p1 * (r2 + 2 * y * y) + 2 * p2 * x * y
        x = (x0 - dx) * icdist
        y = (y0 - dy) * icdist
    if R is not None:
        R = np.asarray(R, dtype=np.float64)
        X = R[0, 0] * x + R[0, 1] * y + R[0, 2]
        Y = R[1, 0] * x + R[1, 1] * y + R[1, 2]
        Wv = 1.0 / (R[2, 0] * x + R[2, 1] * y + R[2, 2])
        x, y = X * Wv, Y * Wv
    return np.stack([x, y], axis=1).astype(out_dtype)


def distort_radtan(pts, intr, dist):
    """cv2.projectPoints([x,y,1], 0, 0, K, D) (camera_model.py:72-74); f32 out for f32 in."""
    pts = np.asarray(pts)
    out_dtype = pts.dtype if pts.dtype in (np.float32, np.float64) else np.float64
    p = pts.astype(np.float64).reshape(-1, 2)
    fx, fy, cx, cy = (float(v) for v in intr)
    k1, k2, p1, p2 = (float(v) for v in dist[:4])
    x, y = p[:, 0], p[:, 1]
    r2 = x * x + y * y
    r4 = r2 * r2
    a1 = 2 * x * y
    a2 = r2 + 2 * x * x
    a3 = r2 + 2 * y * y
    cdist = 1 + k1 * r2 + k2 * r4
    xd = x * cdist + p1 * a1 + p2 * a2
    yd = y * cdist + p1 * a3 + p2 * a1
    return np.stack([xd * fx + cx, yd * fy + cy], axis=1).astype(out_dtype)


# --------------------------------------------------------------------------------------
# A.5b  equidistant (fisheye) undistort / distort -- camera_model.py:41-43, 69-70
# --------------------------------------------------------------------------------------

def undistort_equidistant(pts, intr, dist, R=None):
    """cv2.fisheye.undistortPoints(pts, K, D, R, P=I) (OpenCV 4.13 calib3d/fisheye.cpp): theta_d = |(p - c) / f| clipped
    to pi/2, Newton on theta (1 + k1 th^2 + k2 th^4 + k3 th^6 + k4 th^8) = theta_d, at most 10 steps, stop when a step is
    smaller than 1e-8; scale = tan(theta) / theta_d; then R and the perspective divide.  A point that does not converge
    or whose theta flips sign becomes (-1e6, -1e6).  f64 arithmetic, output dtype = input dtype."""
    pts = np.asarray(pts)
    out_dtype = pts.dtype if pts.dtype in (np.float32, np.float64) else np.float64
    p = pts.astype(np.float64).reshape(-1, 2)
    fx, fy, cx, cy = (float(v) for v in intr)
    k = [float(v) for v in dist[:4]]
    Rm = np.eye(3) if R is None else np.asarray(R, dtype=np.float64)
    out = np.empty_like(p)
    for i in range(len(p)):
        wx, wy = (p[i, 0] - cx) / fx, (p[i, 1] - cy) / fy
        theta_d = min(max(-np.pi / 2.0, float(np.sqrt(wx * wx + wy * wy))), np.pi / 2.0)
        converged, theta, scale = False, theta_d, 0.0
        if abs(theta_d) > 1e-8:
            for _ in range(10):
                t2 = theta * theta
                t4 = t2 * t2
                t6 = t4 * t2
                t8 = t6 * t2
                k0t2, k1t4, k2t6, k3t8 = k[0] * t2, k[1] * t4, k[2] * t6, k[3] * t8
                fix = (theta * (1 + k0t2 + k1t4 + k2t6 + k3t8) - theta_d) / (1 + 3 * k0t2 + 5 * k1t4 + 7 * k2t6 + 9 * k3t8)
                theta = theta - fix
                if abs(fix) < 1e-8:
                    converged = True
                    break
            scale = float(np.tan(theta)) / theta_d
        else:
            converged = True
        flipped = (theta_d < 0 and theta > 0) or (theta_d > 0 and theta < 0)
        if converged and not flipped:
            ux, uy = wx * scale, wy * scale
            X = Rm[0, 0] * ux + Rm[0, 1] * uy + Rm[0, 2]
            Y = Rm[1, 0] * ux + Rm[1, 1] * uy + Rm[1, 2]
            Wv = Rm[2, 0] * ux + Rm[2, 1] * uy + Rm[2, 2]
            out[i] = (X / Wv, Y / Wv)
        else:
            out[i] = (-1000000.0, -1000000.0)
    return out.astype(out_dtype)


def distort_equidistant(pts, intr, dist):
    """cv2.fisheye.distortPoints(pts, K, D): theta = atan(r), theta_d = theta (1 + k1 th^2 + ... + k4 th^8) evaluated
    as theta + k1 th^3 + k2 th^5 + k3 th^7 + k4 th^9, scaled by 1/r (1 when r <= 1e-8)."""
    pts = np.asarray(pts)
    out_dtype = pts.dtype if pts.dtype in (np.float32, np.float64) else np.float64
    p = pts.astype(np.float64).reshape(-1, 2)
    fx, fy, cx, cy = (float(v) for v in intr)
    k = [float(v) for v in dist[:4]]
    out = np.empty_like(p)
    for i in range(len(p)):
        x, y = p[i]
        r = float(np.sqrt(x * x + y * y))
        theta = float(np.arctan(r))
        t2 = theta * theta
        t3 = t2 * theta
        t4 = t2 * t2
        t5 = t4 * theta
        t6 = t3 * t3
        t7 = t6 * theta
        t8 = t4 * t4
        t9 = t8 * theta
        theta_d = theta + k[0] * t3 + k[1] * t5 + k[2] * t7 + k[3] * t9
        cdist = theta_d / r if r > 1e-8 else 1.0
        out[i] = (x * cdist * fx + cx, y * cdist * fy + cy)
    return out.astype(out_dtype)
