// K3/K4 device code: one warp tracks one point through all pyramid levels.
//
// Pyramidal Lucas-Kanade exactly as cv2.calcOpticalFlowPyrLK computes it with
// OPTFLOW_USE_INITIAL_FLOW, winSize 15x15 (reference call sites: feature_tracker.py:102-108,
// stereo_matcher.py:64-68, 70-74; rules: SURVEY.md Appendix A.4):
//   * 14-bit fixed-point bilinear weights (round-half-even), template I = (bilin + 2^8) >> 9 (int16),
//     Ix/Iy = bilinear of the int16 Scharr derivative, (v + 2^13) >> 14; derivative = 0 outside the
//     level, intensities REFLECT_101 outside the level (pyramid border = winSize)
//   * A11/A12/A22 and b1/b2: integer products summed EXACTLY (split 16/16 warp REDUX), rounded once to
//     float32 and scaled by 2^-20 (cv2 accumulates in float32; identical to < 1e-3 px, see oracle/)
//   * all Point2f arithmetic in float32 without FMA contraction (library is built with -fmad=false)
//   * three loop exits (|delta|^2 <= eps^2 in double, oscillation, window out of bounds), status
//     decided at level 0 only, final window re-test (the "err" block).
//
// Lane mapping: lane = 2*row + half.  Row r in 0..15 of the 16x16 bilinear footprint, half 0 owns window
// columns 0..7, half 1 columns 8..14.  Every lane fetches its 9 bytes of row r from L1/L2; row r+1 arrives
// by shuffle from lane+2, so an iteration costs 9 byte loads + 3 shuffles + 4 REDUX per lane.
#pragma once

#include "avb_common.cuh"

#define LK_W_BITS 14

struct LKStatic {                   // lane-constant decomposition
    int lane, r, c0, npx;
};

__device__ __forceinline__ LKStatic lk_static() {
    LKStatic s;
    s.lane = threadIdx.x & 31;
    s.r = s.lane >> 1;
    s.c0 = (s.lane & 1) * 8;
    s.npx = (s.r < AVB_WIN) ? ((s.lane & 1) ? 7 : 8) : 0;
    return s;
}

// exact warp sum of per-lane int32 partials (|v| < 2^31) as int64
__device__ __forceinline__ long long warp_sum_exact(int v) {
    const unsigned lo = __reduce_add_sync(0xffffffffu, (unsigned)(v & 0xffff));
    const int hi = __reduce_add_sync(0xffffffffu, v >> 16);
    return (long long)hi * 65536ll + (long long)lo;
}

__device__ __forceinline__ void lk_weights(float a, float b, int& w00, int& w01, int& w10, int& w11) {
    const float s = (float)(1 << LK_W_BITS);
    const float na = __fsub_rn(1.f, a), nb = __fsub_rn(1.f, b);
    w00 = __float2int_rn(__fmul_rn(__fmul_rn(na, nb), s));
    w01 = __float2int_rn(__fmul_rn(__fmul_rn(a, nb), s));
    w10 = __float2int_rn(__fmul_rn(__fmul_rn(na, b), s));
    w11 = (1 << LK_W_BITS) - w00 - w01 - w10;
}

// n bytes of image row (iy) starting at column ix, REFLECT_101 on both axes
template <int N>
__device__ __forceinline__ void load_row(const uint8_t* __restrict__ img, int w, int h, int pitch, int ix, int iy,
                                         int (&out)[N]) {
    const uint8_t* row = img + (size_t)refl101(iy, h) * pitch;
    if (ix >= 0 && ix + N <= w) {
#pragma unroll
        for (int k = 0; k < N; ++k) out[k] = __ldg(row + ix + k);
    } else {
#pragma unroll
        for (int k = 0; k < N; ++k) out[k] = __ldg(row + refl101(ix + k, w));
    }
}

struct LKParams {
    int nlev;
    int max_iter;
    double min_eig;                 // compared in double, as cv2 does (float minEig vs double threshold)
    double eps2;
};

// Tracks (px0, py0) from pyramid A to pyramid B starting at guess (gx, gy).  All arguments and results are
// warp-uniform.  Returns status (cv2's status byte); (ox, oy) = nextPts[i].
__device__ __forceinline__ bool lk_track_warp(const PyrView& A, const PyrView& B, const Geom& g, float px0, float py0,
                                              float gx, float gy, const LKParams& prm, float& ox, float& oy) {
    const LKStatic L = lk_static();
    bool status = true;
    float nx = 0.f, ny = 0.f;       // nextPts[i]

    for (int level = prm.nlev - 1; level >= 0; --level) {
        const float sc = __int_as_float((127 - level) << 23);       // 2^-level, exact
        const int cols = g.lv[level].w, rows = g.lv[level].h, pitch = g.lv[level].pitch;
        float px = __fmul_rn(px0, sc), py = __fmul_rn(py0, sc);
        if (level == prm.nlev - 1) {
            nx = __fmul_rn(gx, sc);
            ny = __fmul_rn(gy, sc);
        } else {
            nx = __fmul_rn(nx, 2.f);
            ny = __fmul_rn(ny, 2.f);
        }
        px = __fsub_rn(px, (float)AVB_HALF);
        py = __fsub_rn(py, (float)AVB_HALF);
        const int ipx = __float2int_rd(px), ipy = __float2int_rd(py);
        if (ipx < -AVB_WIN || ipx >= cols || ipy < -AVB_WIN || ipy >= rows) {
            if (level == 0) status = false;
            continue;
        }
        int w00, w01, w10, w11;
        lk_weights(__fsub_rn(px, (float)ipx), __fsub_rn(py, (float)ipy), w00, w01, w10, w11);

        // ---- template: I, Ix, Iy for this lane's (up to) 8 window pixels -------------------------
        short tI[8], tIx[8], tIy[8];
        int sA11 = 0, sA12 = 0, sA22 = 0;
        {
            const uint8_t* img = A.lv[level];
            const int y = ipy + L.r, x = ipx + L.c0;            // absolute position of this lane's first sample
            int up[11], mid[11], dn[11];                         // rows y-1, y, y+1; columns x-1 .. x+9
            load_row<11>(img, cols, rows, pitch, x - 1, y - 1, up);
            load_row<11>(img, cols, rows, pitch, x - 1, y, mid);
            load_row<11>(img, cols, rows, pitch, x - 1, y + 1, dn);
            // Scharr derivative at (y, x+k), k = 0..8; zero outside the level
            unsigned dcur[9];
            const bool yin = (y >= 0) && (y < rows);
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                int dx = 3 * (up[k + 2] - up[k]) + 10 * (mid[k + 2] - mid[k]) + 3 * (dn[k + 2] - dn[k]);
                int dy = 3 * (dn[k] - up[k]) + 10 * (dn[k + 1] - up[k + 1]) + 3 * (dn[k + 2] - up[k + 2]);
                const bool in = yin && (x + k >= 0) && (x + k < cols);
                dx = in ? dx : 0;
                dy = in ? dy : 0;
                dcur[k] = ((unsigned)dx & 0xffffu) | ((unsigned)dy << 16);
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const unsigned dnxt0 = __shfl_down_sync(0xffffffffu, dcur[k], 2);
                const unsigned dnxt1 = __shfl_down_sync(0xffffffffu, dcur[k + 1], 2);
                const int iv = (mid[k + 1] * w00 + mid[k + 2] * w01 + dn[k + 1] * w10 + dn[k + 2] * w11 + (1 << (LK_W_BITS - 6))) >>
                               (LK_W_BITS - 5);
                const int dx00 = (short)(dcur[k] & 0xffff), dy00 = (int)dcur[k] >> 16;
                const int dx01 = (short)(dcur[k + 1] & 0xffff), dy01 = (int)dcur[k + 1] >> 16;
                const int dx10 = (short)(dnxt0 & 0xffff), dy10 = (int)dnxt0 >> 16;
                const int dx11 = (short)(dnxt1 & 0xffff), dy11 = (int)dnxt1 >> 16;
                int ixv = (dx00 * w00 + dx01 * w01 + dx10 * w10 + dx11 * w11 + (1 << (LK_W_BITS - 1))) >> LK_W_BITS;
                int iyv = (dy00 * w00 + dy01 * w01 + dy10 * w10 + dy11 * w11 + (1 << (LK_W_BITS - 1))) >> LK_W_BITS;
                const bool act = k < L.npx;
                ixv = act ? ixv : 0;
                iyv = act ? iyv : 0;
                tI[k] = (short)iv;
                tIx[k] = (short)ixv;
                tIy[k] = (short)iyv;
                sA11 += ixv * ixv;
                sA12 += ixv * iyv;
                sA22 += iyv * iyv;
            }
        }
        const float A11 = __fmul_rn(__ll2float_rn(warp_sum_exact(sA11)), 9.5367431640625e-07f);
        const float A12 = __fmul_rn(__ll2float_rn(warp_sum_exact(sA12)), 9.5367431640625e-07f);
        const float A22 = __fmul_rn(__ll2float_rn(warp_sum_exact(sA22)), 9.5367431640625e-07f);
        float D = __fsub_rn(__fmul_rn(A11, A22), __fmul_rn(A12, A12));
        const float dif = __fsub_rn(A11, A22);
        const float rad = __fadd_rn(__fmul_rn(dif, dif), __fmul_rn(__fmul_rn(4.f, A12), A12));
        const float min_eig = __fdiv_rn(__fsub_rn(__fadd_rn(A22, A11), __fsqrt_rn(rad)), (float)(2 * AVB_WIN * AVB_WIN));
        if ((double)min_eig < prm.min_eig || D < 1.1920929e-07f) {
            if (level == 0) status = false;
            continue;
        }
        D = __fdiv_rn(1.f, D);

        // ---- iterations on image B ------------------------------------------------------------------
        float cx = __fsub_rn(nx, (float)AVB_HALF), cy = __fsub_rn(ny, (float)AVB_HALF);
        float pdx = 0.f, pdy = 0.f;
        const uint8_t* imgB = B.lv[level];
        for (int j = 0; j < prm.max_iter; ++j) {
            const int inx = __float2int_rd(cx), iny = __float2int_rd(cy);
            if (inx < -AVB_WIN || inx >= cols || iny < -AVB_WIN || iny >= rows) {
                if (level == 0) status = false;
                break;
            }
            lk_weights(__fsub_rn(cx, (float)inx), __fsub_rn(cy, (float)iny), w00, w01, w10, w11);
            int cur[9];
            load_row<9>(imgB, cols, rows, pitch, inx + L.c0, iny + L.r, cur);
            const unsigned p0 = (unsigned)cur[0] | ((unsigned)cur[1] << 8) | ((unsigned)cur[2] << 16) | ((unsigned)cur[3] << 24);
            const unsigned p1 = (unsigned)cur[4] | ((unsigned)cur[5] << 8) | ((unsigned)cur[6] << 16) | ((unsigned)cur[7] << 24);
            const unsigned q0 = __shfl_down_sync(0xffffffffu, p0, 2);
            const unsigned q1 = __shfl_down_sync(0xffffffffu, p1, 2);
            const int q2 = __shfl_down_sync(0xffffffffu, cur[8], 2);
            int nxt[9];
            nxt[0] = q0 & 0xff; nxt[1] = (q0 >> 8) & 0xff; nxt[2] = (q0 >> 16) & 0xff; nxt[3] = q0 >> 24;
            nxt[4] = q1 & 0xff; nxt[5] = (q1 >> 8) & 0xff; nxt[6] = (q1 >> 16) & 0xff; nxt[7] = q1 >> 24;
            nxt[8] = q2;
            int sb1 = 0, sb2 = 0;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int jv = (cur[k] * w00 + cur[k + 1] * w01 + nxt[k] * w10 + nxt[k + 1] * w11 + (1 << (LK_W_BITS - 6))) >>
                               (LK_W_BITS - 5);
                const int diff = jv - (int)tI[k];
                sb1 += diff * (int)tIx[k];              // tIx/tIy are zero for inactive slots
                sb2 += diff * (int)tIy[k];
            }
            const float b1 = __fmul_rn(__ll2float_rn(warp_sum_exact(sb1)), 9.5367431640625e-07f);
            const float b2 = __fmul_rn(__ll2float_rn(warp_sum_exact(sb2)), 9.5367431640625e-07f);
            const float dx = __fmul_rn(__fsub_rn(__fmul_rn(A12, b2), __fmul_rn(A22, b1)), D);
            const float dy = __fmul_rn(__fsub_rn(__fmul_rn(A12, b1), __fmul_rn(A11, b2)), D);
            cx = __fadd_rn(cx, dx);
            cy = __fadd_rn(cy, dy);
            nx = __fadd_rn(cx, (float)AVB_HALF);
            ny = __fadd_rn(cy, (float)AVB_HALF);
            if (__dadd_rn(__dmul_rn((double)dx, (double)dx), __dmul_rn((double)dy, (double)dy)) <= prm.eps2) break;
            if (j > 0 && (double)fabsf(__fadd_rn(dx, pdx)) < 0.01 && (double)fabsf(__fadd_rn(dy, pdy)) < 0.01) {
                nx = __fsub_rn(nx, __fmul_rn(dx, 0.5f));
                ny = __fsub_rn(ny, __fmul_rn(dy, 0.5f));
                break;
            }
            pdx = dx;
            pdy = dy;
        }
    }
    if (status) {                   // the err block re-tests the final window position
        const int fx = __float2int_rd(__fsub_rn(nx, (float)AVB_HALF)), fy = __float2int_rd(__fsub_rn(ny, (float)AVB_HALF));
        if (fx < -AVB_WIN || fx >= g.lv[0].w || fy < -AVB_WIN || fy >= g.lv[0].h) status = false;
    }
    ox = nx;
    oy = ny;
    return status;
}

// ---- radtan undistort / distort in double (cv2.undistortPoints / projectPoints, Appendix A.5) ----------
__device__ __forceinline__ void undistort_pt(const CamModel& c, double u, double v, const double* R, double& ox, double& oy) {
    const double ifx = 1.0 / c.fx, ify = 1.0 / c.fy;
    const double x0 = (u - c.cx) * ifx, y0 = (v - c.cy) * ify;
    double x = x0, y = y0;
#pragma unroll 1
    for (int j = 0; j < 5; ++j) {
        const double r2 = x * x + y * y;
        const double icdist = 1.0 / (1.0 + ((0.0 * r2 + c.k2) * r2 + c.k1) * r2);
        const double dX = 2.0 * c.p1 * x * y + c.p2 * (r2 + 2.0 * x * x);
        const double dY = c.p1 * (r2 + 2.0 * y * y) + 2.0 * c.p2 * x * y;
        x = (x0 - dX) * icdist;
        y = (y0 - dY) * icdist;
    }
    if (R) {
        const double xx = R[0] * x + R[1] * y + R[2];
        const double yy = R[3] * x + R[4] * y + R[5];
        const double ww = 1.0 / (R[6] * x + R[7] * y + R[8]);
        x = xx * ww;
        y = yy * ww;
    }
    ox = x;
    oy = y;
}

__device__ __forceinline__ void distort_pt(const CamModel& c, double x, double y, double& ou, double& ov) {
    const double r2 = x * x + y * y, r4 = r2 * r2;
    const double a1 = 2.0 * x * y, a2 = r2 + 2.0 * x * x, a3 = r2 + 2.0 * y * y;
    const double cdist = 1.0 + c.k1 * r2 + c.k2 * r4;
    const double xd = x * cdist + c.p1 * a1 + c.p2 * a2;
    const double yd = y * cdist + c.p1 * a3 + c.p2 * a1;
    ou = xd * c.fx + c.cx;
    ov = yd * c.fy + c.cy;
}

// StereoMatcher.stereo_match for one cam0 point (stereo_matcher.py:33-115), warp-uniform.
// The cam0 model is used for both cameras (Appendix B3); backward-LK status is ignored (B5);
// the epipolar error keeps only the x-term of the element-wise product (B4).
__device__ __forceinline__ bool stereo_match_warp(const PyrView& P0, const PyrView& P1, const Geom& g, const LKParams& prm,
                                                  float x0, float y0, float& x1, float& y1) {
    double ux, uy, pu, pv;
    undistort_pt(g.cam0, (double)x0, (double)y0, g.R01, ux, uy);
    distort_pt(g.cam0, (double)(float)ux, (double)(float)uy, pu, pv);
    const float gx = (float)pu, gy = (float)pv;                 // proj1 (float32)
    float fx, fy, bx, by;
    const bool st_f = lk_track_warp(P0, P1, g, x0, y0, gx, gy, prm, fx, fy);
    x1 = fx;
    y1 = fy;
    if (!st_f) return false;
    lk_track_warp(P1, P0, g, fx, fy, x0, y0, prm, bx, by);
    const float ex = __fsub_rn(x0, bx), ey = __fsub_rn(y0, by);
    const float err = __fsqrt_rn(__fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey)));
    const float disp = fabsf(__fsub_rn(gy, fy));
    bool ok = (err < 3.f) && (disp < 20.f);
    ok = ok && !(fx < 0.f || fx >= (float)g.W || fy < 0.f || fy >= (float)g.H);
    if (!ok) return false;
    double a0, b0, a1, b1;
    undistort_pt(g.cam0, (double)x0, (double)y0, nullptr, a0, b0);
    undistort_pt(g.cam0, (double)fx, (double)fy, nullptr, a1, b1);
    const double u0x = (double)(float)a0, u0y = (double)(float)b0, u1x = (double)(float)a1;
    const double l0 = g.E[0] * u0x + g.E[1] * u0y + g.E[2];
    const double l1 = g.E[3] * u0x + g.E[4] * u0y + g.E[5];
    const double epi = fabs(u1x * l0) / sqrt(l0 * l0 + l1 * l1);
    return !(epi > g.epi_thr);
}
