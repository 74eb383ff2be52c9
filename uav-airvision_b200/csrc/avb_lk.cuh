// K3/K4 device code: a team of WPF warps tracks one point through all pyramid levels.
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
// Two lane mappings of the 16x16 bilinear footprint (15x15 window + the row/column below/right):
//   WPF = 1  (throughput: many streams per launch)   lane = 2*row + half; half 0 owns window columns 0..7,
//            half 1 columns 8..14; one warp per feature, no shared memory, no block barrier.
//   WPF = 4  (latency: one stream)   4 warps per feature; a warp owns 4 window rows + the supply row below them,
//            lane = 6*rr + cg, cg < 5 owns columns 3cg..3cg+2; a thread touches 3 pixels per iteration instead
//            of 8, the four warps meet in one block barrier per reduction.
// In both, row r+1 arrives by shuffle from lane + LPR, so every pixel of image B is fetched once per iteration.
#pragma once

#include <limits.h>

#include "avb_common.cuh"

#define LK_W_BITS 14

template <int WPF>
struct LKMap;
template <>
struct LKMap<1> {
    static constexpr int LPR = 2, NPX = 8;
};
template <>
struct LKMap<4> {
    static constexpr int LPR = 6, NPX = 3;
};

struct LKStatic {                   // lane-constant decomposition
    int row, c0, npx;               // window row (0..15; 15 only supplies), first column, active pixels (0 = none)
};

template <int WPF>
__device__ __forceinline__ LKStatic lk_static() {
    LKStatic s;
    const int lane = threadIdx.x & 31;
    if (WPF == 1) {
        s.row = lane >> 1;
        s.c0 = (lane & 1) * 8;
        s.npx = (s.row < AVB_WIN) ? ((lane & 1) ? 7 : 8) : 0;
    } else {
        const int w = (threadIdx.x >> 5) & 3, rr = lane / 6, cg = lane - rr * 6;
        s.row = 4 * w + rr;
        s.c0 = cg * 3;
        s.npx = (lane < 30 && cg < 5 && rr < 4 && s.row < AVB_WIN) ? 3 : 0;
    }
    return s;
}

struct LKShared;

// Exact warp sums of N per-lane int32 partials (|v| < 2^31) as int64: split 16/16 warp REDUX.  This is the general path;
// almost every level takes the one-REDUX path below (see lk_track_warp).
template <int WPF, int N>
__device__ __forceinline__ void team_sum_exact(const int (&v)[N], long long (&out)[N], LKShared*, int&) {
    static_assert(WPF == 1, "the 4-warp mapping reduces inline");
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const unsigned lo = __reduce_add_sync(0xffffffffu, (unsigned)(v[i] & 0xffff));
        const int hi = __reduce_add_sync(0xffffffffu, v[i] >> 16);
        out[i] = (long long)hi * 65536ll + (long long)lo;
    }
}

// Largest structure-tensor diagonal sum (integer, before the 2^-20 scale) for which the residual sums b1, b2 of a level
// are guaranteed to fit in int32, so that ONE 32-bit REDUX per sum is exact (two's-complement wrap-around cancels as long
// as the true total fits):  |b| = |sum d_k g_k| <= sqrt(sum d_k^2) sqrt(sum g_k^2) <= sqrt(225) * 8160 * sqrt(q)  with
// |d_k| <= 8160 (template and sample are both (bilinear + 2^8) >> 9 of bytes), q = sum g_k^2 = q11 or q22;
// 122400 * sqrt(3.0e8) = 2.12e9 < 2^31.  Measured on the bench texture: max q = 6.6e7 (every level narrow).
#define LK_NARROW_MAX 300000000u

// 2-way dot product of two SIGNED 16-bit values (operand a) with two UNSIGNED bytes (lower / upper half of b) plus c.
// The fourth bilinear weight is 2^14 minus the other three rounded weights and can come out as -1, hence signed.
__device__ __forceinline__ int dp2a_lo_su(unsigned a, unsigned b, int c) {
    int d;
    asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ int dp2a_hi_su(unsigned a, unsigned b, int c) {
    int d;
    asm("dp2a.hi.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
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
    float eps2_lo, eps2_hi;         // float band around eps2 outside which the float estimate decides (lk_converged)
    float eig_accept;               // D > eig_accept * (A11 + A22) proves minEig >= threshold without the sqrt / division
};


// Template of one window: I, Ix, Iy of this lane's (up to) NPX pixels and the lane's partial structure-tensor sums.
// (ipx, ipy) = floor(prevPt - halfWin) at this level; weights from its fractional part.
template <int NPX, int LPR>
__device__ __forceinline__ void lk_template(const uint8_t* __restrict__ img, int cols, int rows, int pitch, int ipx, int ipy,
                                            int w00, int w01, int w10, int w11, const LKStatic& L, short (&tI)[NPX],
                                            short (&tIx)[NPX], short (&tIy)[NPX], int& sA11, int& sA12, int& sA22) {
    const int y = ipy + L.row, x = ipx + L.c0;          // absolute position of this lane's first sample
    int up[NPX + 3], mid[NPX + 3], dn[NPX + 3];          // rows y-1, y, y+1; columns x-1 .. x+NPX+1
    const bool inside = LPR == 2 && NPX == 8 && ipx >= 1 && ipx + 18 <= cols && ipy >= 1 && ipy + 17 <= rows;
    if (inside) {
        // Warp-uniform: the 18 x 18 footprint lies inside the level.  A lane fetches only ITS row, as aligned 32-bit
        // words (11 bytes from column x-1: at most 4 words), and takes the rows above and below from lanes -2 / +2;
        // the first and the last row of lanes fetch the one row nobody owns.  A warp-wide byte load of 16 rows costs 16
        // L1 wavefronts; this is 4 + 2 loads per template instead of 33 (the L1 data pipe was the busiest unit of the
        // 64-stream LK kernels, and 3/4 of its wavefronts came from here).
        const uint8_t* a = img + (size_t)y * pitch + (x - 1);
        const unsigned o = (unsigned)(size_t)a & 3u, sh = o * 8u;
        const unsigned* aw = reinterpret_cast<const unsigned*>((size_t)a & ~(size_t)3);
        unsigned M[3], U[3], D[3];
        {
            const unsigned W0 = __ldg(aw), W1 = __ldg(aw + 1), W2 = __ldg(aw + 2), W3 = o >= 2 ? __ldg(aw + 3) : 0u;
            M[0] = __funnelshift_r(W0, W1, sh);
            M[1] = __funnelshift_r(W1, W2, sh);
            M[2] = __funnelshift_r(W2, W3, sh);
        }
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            U[i] = __shfl_up_sync(0xffffffffu, M[i], LPR);
            D[i] = __shfl_down_sync(0xffffffffu, M[i], LPR);
        }
        if (L.row == 0 || L.row == 15) {                 // rows ipy-1 and ipy+16: aligned like the lane's own row
            const unsigned* ew = reinterpret_cast<const unsigned*>(
                reinterpret_cast<const uint8_t*>(aw) + (L.row == 0 ? -(ptrdiff_t)pitch : (ptrdiff_t)pitch));
            const unsigned W0 = __ldg(ew), W1 = __ldg(ew + 1), W2 = __ldg(ew + 2), W3 = o >= 2 ? __ldg(ew + 3) : 0u;
            const unsigned E0 = __funnelshift_r(W0, W1, sh), E1 = __funnelshift_r(W1, W2, sh), E2 = __funnelshift_r(W2, W3, sh);
            if (L.row == 0) {
                U[0] = E0, U[1] = E1, U[2] = E2;
            } else {
                D[0] = E0, D[1] = E1, D[2] = E2;
            }
        }
#pragma unroll
        for (int k = 0; k < NPX + 3; ++k) {
            up[k] = (int)((U[k >> 2] >> (8 * (k & 3))) & 0xffu);
            mid[k] = (int)((M[k >> 2] >> (8 * (k & 3))) & 0xffu);
            dn[k] = (int)((D[k >> 2] >> (8 * (k & 3))) & 0xffu);
        }
    } else {
        load_row<NPX + 3>(img, cols, rows, pitch, x - 1, y - 1, up);
        load_row<NPX + 3>(img, cols, rows, pitch, x - 1, y, mid);
        load_row<NPX + 3>(img, cols, rows, pitch, x - 1, y + 1, dn);
    }
    // Scharr derivative at (y, x+k), k = 0..NPX, separably: dx = [3 10 3]^T (x) [-1 0 1], dy = [-1 0 1]^T (x) [3 10 3]
    int sv[NPX + 3], tv[NPX + 3];
#pragma unroll
    for (int j = 0; j < NPX + 3; ++j) {
        sv[j] = 3 * (up[j] + dn[j]) + 10 * mid[j];
        tv[j] = dn[j] - up[j];
    }
    int dxc[NPX + 1], dyc[NPX + 1], dxn[NPX + 1], dyn[NPX + 1];
#pragma unroll
    for (int k = 0; k <= NPX; ++k) {
        dxc[k] = sv[k + 2] - sv[k];
        dyc[k] = 3 * (tv[k] + tv[k + 2]) + 10 * tv[k + 1];
    }
    if (!inside) {                                       // zero outside the level (cv2's derivative image has no border)
        const bool yin = (y >= 0) && (y < rows);
#pragma unroll
        for (int k = 0; k <= NPX; ++k) {
            const bool in = yin && (x + k >= 0) && (x + k < cols);
            dxc[k] = in ? dxc[k] : 0;
            dyc[k] = in ? dyc[k] : 0;
        }
    }
#pragma unroll
    for (int k = 0; k <= NPX; ++k) {                     // the row below, from lane + LPR
        dxn[k] = __shfl_down_sync(0xffffffffu, dxc[k], LPR);
        dyn[k] = __shfl_down_sync(0xffffffffu, dyc[k], LPR);
    }
    // lanes that only supply a row (npx == 0) interpolate their derivatives with zero weights: Ix = Iy = 0, which is what
    // the iteration sums need from an inactive slot; the same for the one slot past a 7-pixel lane
    const bool lane_on = L.npx > 0;
    const int v00 = lane_on ? w00 : 0, v01 = lane_on ? w01 : 0, v10 = lane_on ? w10 : 0, v11 = lane_on ? w11 : 0;
    sA11 = sA12 = sA22 = 0;
#pragma unroll
    for (int k = 0; k < NPX; ++k) {
        const int iv = (mid[k + 1] * w00 + mid[k + 2] * w01 + dn[k + 1] * w10 + dn[k + 2] * w11 + (1 << (LK_W_BITS - 6))) >>
                       (LK_W_BITS - 5);
        int ixv = (dxc[k] * v00 + dxc[k + 1] * v01 + dxn[k] * v10 + dxn[k + 1] * v11 + (1 << (LK_W_BITS - 1))) >> LK_W_BITS;
        int iyv = (dyc[k] * v00 + dyc[k + 1] * v01 + dyn[k] * v10 + dyn[k + 1] * v11 + (1 << (LK_W_BITS - 1))) >> LK_W_BITS;
        if (k == NPX - 1) {
            const bool act = L.npx == NPX || !lane_on;   // (inactive lanes are already zero)
            ixv = act ? ixv : 0;
            iyv = act ? iyv : 0;
        }
        tI[k] = (short)iv;
        tIx[k] = (short)ixv;
        tIy[k] = (short)iyv;
        sA11 += ixv * ixv;
        sA12 += ixv * iyv;
        sA22 += iyv * iyv;
    }
}

// Structure tensor -> (A11, A12, A22, 1/D); false when cv2 rejects the level (minEig / determinant test).
//   minEig = (A22 + A11 - sqrt((A11 - A22)^2 + 4 A12^2)) / (2 * 15 * 15) = lambda_min / 225, compared with the threshold
// in double.  lambda_min = D / lambda_max >= D / (A11 + A22), so D > 450 * (2 thr + 4e-5) * (A11 + A22) puts the exact
// value above twice the threshold; the float rounding of D (<= 6e-8 s^2, s = A11 + A22 <= 7056) and of the minEig
// formula (<= 5.3e-10 s) cannot bring it back below it.  Only borderline levels evaluate the IEEE sqrt and division.
__device__ __forceinline__ bool lk_tensor_f(float A11, float A12, float A22, double min_eig_thr, float eig_accept,
                                            float& Dinv) {
    const float D = __fsub_rn(__fmul_rn(A11, A22), __fmul_rn(A12, A12));
    if (D < 1.1920929e-07f) return false;
    const float s = __fadd_rn(A22, A11);
    if (!(D > __fmul_rn(eig_accept, s))) {
        const float dif = __fsub_rn(A11, A22);
        const float rad = __fadd_rn(__fmul_rn(dif, dif), __fmul_rn(__fmul_rn(4.f, A12), A12));
        const float min_eig = __fdiv_rn(__fsub_rn(s, __fsqrt_rn(rad)), (float)(2 * AVB_WIN * AVB_WIN));
        if ((double)min_eig < min_eig_thr) return false;
    }
    Dinv = __fdiv_rn(1.f, D);
    return true;
}
__device__ __forceinline__ bool lk_tensor(long long q11, long long q12, long long q22, double min_eig_thr, float eig_accept,
                                          float& A11, float& A12, float& A22, float& Dinv) {
    A11 = __fmul_rn(__ll2float_rn(q11), 9.5367431640625e-07f);
    A12 = __fmul_rn(__ll2float_rn(q12), 9.5367431640625e-07f);
    A22 = __fmul_rn(__ll2float_rn(q22), 9.5367431640625e-07f);
    return lk_tensor_f(A11, A12, A22, min_eig_thr, eig_accept, Dinv);
}

// cv2's two loop exits, bit for bit, without the double-precision round trip on the common path:
//   converged   <=> (double)dx*dx + (double)dy*dy <= eps^2          (criteria.epsilon squared, in double)
//   oscillating <=> j > 0 and |dx + prev dx| < 0.01 and |dy + prev dy| < 0.01   (float sums compared with the double 0.01)
// The float32 estimate of |delta|^2 is within 2e-7 relative of the exact value; only inside a 1e-5 band around
// eps^2 is the exact double expression evaluated.  (double)f < 0.01 <=> f <= 0.01f because 0.01f is the float just
// below 0.01.
__device__ __forceinline__ bool lk_converged(float dx, float dy, double eps2, float eps2_lo, float eps2_hi) {
    const float f = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
    if (f < eps2_lo) return true;
    if (f > eps2_hi) return false;
    return __dadd_rn(__dmul_rn((double)dx, (double)dx), __dmul_rn((double)dy, (double)dy)) <= eps2;
}
__device__ __forceinline__ bool lk_oscillating(float dx, float dy, float pdx, float pdy) {
    return fabsf(__fadd_rn(dx, pdx)) <= 0.01f && fabsf(__fadd_rn(dy, pdy)) <= 0.01f;
}

// ---- WPF = 1: one warp per point, template in registers ------------------------------------------------
// Tracks (px0, py0) from pyramid A to pyramid B starting at guess (gx, gy).  All arguments and results are
// warp-uniform.  Returns status (cv2's status byte); (ox, oy) = nextPts[i].
__device__ __forceinline__ bool lk_track_warp(const PyrView& A, const PyrView& B, const Geom& g, float px0, float py0,
                                              float gx, float gy, const LKParams& prm, float& ox, float& oy) {
    constexpr int LPR = LKMap<1>::LPR, NPX = LKMap<1>::NPX;
    const LKStatic L = lk_static<1>();
    bool status = true;
    float nx = 0.f, ny = 0.f;       // nextPts[i]
    int flip = 0;

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

        short tI[NPX], tIx[NPX], tIy[NPX];
        int sA[3];
        lk_template<NPX, LPR>(pyr_level(A, g, level), cols, rows, pitch, ipx, ipy, w00, w01, w10, w11, L, tI, tIx, tIy, sA[0],
                              sA[1], sA[2]);
        // Exact sums of the structure tensor.  q11, q22 <= 225 * 4080^2 < 2^32: one unsigned REDUX each (a lane's partial
        // is < 2^28).  |q12| <= sqrt(q11 q22), so one signed REDUX is exact whenever both diagonals are below 2^31.
        const unsigned q11 = __reduce_add_sync(0xffffffffu, (unsigned)sA[0]);
        const unsigned q22 = __reduce_add_sync(0xffffffffu, (unsigned)sA[2]);
        float A12;
        if ((q11 | q22) < 0x80000000u) {
            A12 = __fmul_rn(__int2float_rn(__reduce_add_sync(0xffffffffu, sA[1])), 9.5367431640625e-07f);
        } else {
            const int one[1] = {sA[1]};
            long long q12[1];
            team_sum_exact<1, 1>(one, q12, nullptr, flip);
            A12 = __fmul_rn(__ll2float_rn(q12[0]), 9.5367431640625e-07f);
        }
        const float A11 = __fmul_rn(__uint2float_rn(q11), 9.5367431640625e-07f);
        const float A22 = __fmul_rn(__uint2float_rn(q22), 9.5367431640625e-07f);
        float D;
        if (!lk_tensor_f(A11, A12, A22, prm.min_eig, prm.eig_accept, D)) {
            if (level == 0) status = false;
            continue;
        }
        const bool narrow = q11 <= LK_NARROW_MAX && q22 <= LK_NARROW_MAX;     // b1, b2 fit in int32 (see LK_NARROW_MAX)

        // ---- iterations on image B ------------------------------------------------------------------
        float cx = __fsub_rn(nx, (float)AVB_HALF), cy = __fsub_rn(ny, (float)AVB_HALF);
        float pdx = 0.f, pdy = 0.f;
        const uint8_t* imgB = pyr_level(B, g, level);
        // window tests as single unsigned compares: in range <=> -WIN <= inx < cols; footprint (16 rows x 17 columns)
        // inside the level <=> 0 <= inx <= cols - 17, 0 <= iny <= rows - 16
        const unsigned x_rng = (unsigned)(cols + AVB_WIN), y_rng = (unsigned)(rows + AVB_WIN);
        const unsigned x_in = (unsigned)max(cols - 16, 0), y_in = (unsigned)max(rows - 15, 0);
        const uint8_t* lane_base = imgB + L.row * pitch + L.c0;      // this lane's first sample for a window at (0, 0)
        // this lane's 9 bytes of row (iny + row) from column inx + c0, packed: P0 = b0..b3, P1 = b4..b7, P2 = b8; N = the
        // row below (from lane + 2).  Kept across iterations: sub-pixel updates leave the integer window position where
        // it was on most iterations after the first, and the samples are then still the right ones.
        unsigned P0 = 0, P1 = 0, P2 = 0, N0 = 0, N1 = 0, N2 = 0;
        int pinx = INT_MIN, piny = INT_MIN;
        for (int j = 0; j < prm.max_iter; ++j) {
            const int inx = __float2int_rd(cx), iny = __float2int_rd(cy);
            if ((unsigned)(inx + AVB_WIN) >= x_rng || (unsigned)(iny + AVB_WIN) >= y_rng) {
                if (level == 0) status = false;
                break;
            }
            lk_weights(__fsub_rn(cx, (float)inx), __fsub_rn(cy, (float)iny), w00, w01, w10, w11);
            if (inx == pinx && iny == piny) {                        // warp-uniform
            } else {
            if ((unsigned)inx < x_in && (unsigned)iny < y_in) {      // warp-uniform: footprint inside the level
                // three aligned 32-bit loads + funnel shifts instead of nine byte loads (rows are 16-byte aligned; the
                // <= 3 bytes read past the footprint stay inside the row pitch / the padded allocation)
                const uint8_t* a = lane_base + (iny * pitch + inx);
                const unsigned sh = ((unsigned)(size_t)a & 3u) * 8u;
                const unsigned* aw = reinterpret_cast<const unsigned*>((size_t)a & ~(size_t)3);
                const unsigned W0 = __ldg(aw), W1 = __ldg(aw + 1), W2 = __ldg(aw + 2);
                P0 = __funnelshift_r(W0, W1, sh);
                P1 = __funnelshift_r(W1, W2, sh);
                P2 = (W2 >> sh) & 0xffu;
            } else {
                int cur[NPX + 1];
                load_row<NPX + 1>(imgB, cols, rows, pitch, inx + L.c0, iny + L.row, cur);
                P0 = (unsigned)cur[0] | ((unsigned)cur[1] << 8) | ((unsigned)cur[2] << 16) | ((unsigned)cur[3] << 24);
                P1 = (unsigned)cur[4] | ((unsigned)cur[5] << 8) | ((unsigned)cur[6] << 16) | ((unsigned)cur[7] << 24);
                P2 = (unsigned)cur[8];
            }
            N0 = __shfl_down_sync(0xffffffffu, P0, LPR);            // the row below, from lane + 2
            N1 = __shfl_down_sync(0xffffffffu, P1, LPR);
            N2 = __shfl_down_sync(0xffffffffu, P2, LPR);
            pinx = inx;
            piny = iny;
            }
            // bilinear sample k = w00 b_k + w01 b_{k+1} + w10 n_k + w11 n_{k+1}: two 2-way 16x8-bit dot products
            // (dp2a) on byte pairs; odd k read the pairs from the words shifted by one byte
            const unsigned Q0 = __funnelshift_r(P0, P1, 8), Q1 = __funnelshift_r(P1, P2, 8);
            const unsigned M0 = __funnelshift_r(N0, N1, 8), M1 = __funnelshift_r(N1, N2, 8);
            const unsigned Wt = ((unsigned)w00 & 0xffffu) | ((unsigned)w01 << 16);
            const unsigned Wb = ((unsigned)w10 & 0xffffu) | ((unsigned)w11 << 16);     // w11 may be -1
            const int rnd = 1 << (LK_W_BITS - 6);
            int jv[NPX];
            jv[0] = dp2a_lo_su(Wt, P0, dp2a_lo_su(Wb, N0, rnd)) >> (LK_W_BITS - 5);
            jv[1] = dp2a_lo_su(Wt, Q0, dp2a_lo_su(Wb, M0, rnd)) >> (LK_W_BITS - 5);
            jv[2] = dp2a_hi_su(Wt, P0, dp2a_hi_su(Wb, N0, rnd)) >> (LK_W_BITS - 5);
            jv[3] = dp2a_hi_su(Wt, Q0, dp2a_hi_su(Wb, M0, rnd)) >> (LK_W_BITS - 5);
            jv[4] = dp2a_lo_su(Wt, P1, dp2a_lo_su(Wb, N1, rnd)) >> (LK_W_BITS - 5);
            jv[5] = dp2a_lo_su(Wt, Q1, dp2a_lo_su(Wb, M1, rnd)) >> (LK_W_BITS - 5);
            jv[6] = dp2a_hi_su(Wt, P1, dp2a_hi_su(Wb, N1, rnd)) >> (LK_W_BITS - 5);
            jv[7] = dp2a_hi_su(Wt, Q1, dp2a_hi_su(Wb, M1, rnd)) >> (LK_W_BITS - 5);
            int sb[2] = {0, 0};
#pragma unroll
            for (int k = 0; k < NPX; ++k) {
                const int diff = jv[k] - (int)tI[k];
                sb[0] += diff * (int)tIx[k];            // tIx/tIy are zero for inactive slots
                sb[1] += diff * (int)tIy[k];
            }
            float b1, b2;
            if (narrow) {               // one REDUX per sum: the true totals fit in int32
                b1 = __fmul_rn(__int2float_rn(__reduce_add_sync(0xffffffffu, sb[0])), 9.5367431640625e-07f);
                b2 = __fmul_rn(__int2float_rn(__reduce_add_sync(0xffffffffu, sb[1])), 9.5367431640625e-07f);
            } else {
                long long qb[2];
                team_sum_exact<1, 2>(sb, qb, nullptr, flip);
                b1 = __fmul_rn(__ll2float_rn(qb[0]), 9.5367431640625e-07f);
                b2 = __fmul_rn(__ll2float_rn(qb[1]), 9.5367431640625e-07f);
            }
            const float dx = __fmul_rn(__fsub_rn(__fmul_rn(A12, b2), __fmul_rn(A22, b1)), D);
            const float dy = __fmul_rn(__fsub_rn(__fmul_rn(A12, b1), __fmul_rn(A11, b2)), D);
            cx = __fadd_rn(cx, dx);
            cy = __fadd_rn(cy, dy);
            nx = __fadd_rn(cx, (float)AVB_HALF);
            ny = __fadd_rn(cy, (float)AVB_HALF);
            if (lk_converged(dx, dy, prm.eps2, prm.eps2_lo, prm.eps2_hi)) break;
            if (j > 0 && lk_oscillating(dx, dy, pdx, pdy)) {
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

// ---- WPF = 4: four warps per point (latency mapping) ----------------------------------------------------
// The template of a level depends only on prevPt, never on the iterations, so the four warps first build the
// templates of ALL levels concurrently (warp w -> levels w, w+4; 16x2 lane mapping, results to shared memory), then
// walk the levels together with the 5x6 lane mapping: 3 pixels per thread, both bilinear rows fetched directly
// (no shuffle on the critical path), one REDUX per sum (60 pixels x 8160 x 4080 < 2^31) and one block barrier per
// iteration.
struct LKShared {
    short I[AVB_MAX_LEVELS][AVB_WIN][16];      // [level][row][col]; col 15 is padding
    short Ix[AVB_MAX_LEVELS][AVB_WIN][16];
    short Iy[AVB_MAX_LEVELS][AVB_WIN][16];
    float A11[AVB_MAX_LEVELS], A12[AVB_MAX_LEVELS], A22[AVB_MAX_LEVELS], Dinv[AVB_MAX_LEVELS];
    int flag[AVB_MAX_LEVELS];                   // bits 0-1: 0 usable, 1 window outside the level, 2 minEig / det reject; bit 2: narrow sums
    alignas(16) int red[2][2][4];               // [flip][sum][warp]
};

__device__ __forceinline__ bool lk_track_coop(const PyrView& A, const PyrView& B, const Geom& g, float px0, float py0,
                                              float gx, float gy, const LKParams& prm, LKShared* sh, int& flip, float& ox,
                                              float& oy) {
    const int warp = (threadIdx.x >> 5) & 3;
    __syncthreads();                            // the previous pass is done with the shared templates
    {
        constexpr int LPR = LKMap<1>::LPR, NPX = LKMap<1>::NPX;
        const LKStatic L = lk_static<1>();
        for (int level = warp; level < prm.nlev; level += 4) {
            const float sc = __int_as_float((127 - level) << 23);
            const int cols = g.lv[level].w, rows = g.lv[level].h, pitch = g.lv[level].pitch;
            const float px = __fsub_rn(__fmul_rn(px0, sc), (float)AVB_HALF), py = __fsub_rn(__fmul_rn(py0, sc), (float)AVB_HALF);
            const int ipx = __float2int_rd(px), ipy = __float2int_rd(py);
            if (ipx < -AVB_WIN || ipx >= cols || ipy < -AVB_WIN || ipy >= rows) {
                if ((threadIdx.x & 31) == 0) sh->flag[level] = 1;
                continue;
            }
            int w00, w01, w10, w11;
            lk_weights(__fsub_rn(px, (float)ipx), __fsub_rn(py, (float)ipy), w00, w01, w10, w11);
            short tI[NPX], tIx[NPX], tIy[NPX];
            int sA[3];
            lk_template<NPX, LPR>(pyr_level(A, g, level), cols, rows, pitch, ipx, ipy, w00, w01, w10, w11, L, tI, tIx, tIy,
                                  sA[0], sA[1], sA[2]);
            if (L.row < AVB_WIN) {
#pragma unroll
                for (int k = 0; k < NPX; ++k) {
                    if (k < L.npx) {
                        sh->I[level][L.row][L.c0 + k] = tI[k];
                        sh->Ix[level][L.row][L.c0 + k] = tIx[k];
                        sh->Iy[level][L.row][L.c0 + k] = tIy[k];
                    }
                }
            }
            // exact sums as in lk_track_warp: one REDUX per sum whenever the totals fit in 32 bits
            const unsigned q11 = __reduce_add_sync(0xffffffffu, (unsigned)sA[0]);
            const unsigned q22 = __reduce_add_sync(0xffffffffu, (unsigned)sA[2]);
            float A12;
            if ((q11 | q22) < 0x80000000u) {
                A12 = __fmul_rn(__int2float_rn(__reduce_add_sync(0xffffffffu, sA[1])), 9.5367431640625e-07f);
            } else {
                const int one[1] = {sA[1]};
                long long q12[1];
                int f0 = 0;
                team_sum_exact<1, 1>(one, q12, nullptr, f0);
                A12 = __fmul_rn(__ll2float_rn(q12[0]), 9.5367431640625e-07f);
            }
            const float A11 = __fmul_rn(__uint2float_rn(q11), 9.5367431640625e-07f);
            const float A22 = __fmul_rn(__uint2float_rn(q22), 9.5367431640625e-07f);
            float D = 0.f;
            const bool ok = lk_tensor_f(A11, A12, A22, prm.min_eig, prm.eig_accept, D);
            if ((threadIdx.x & 31) == 0) {
                sh->A11[level] = A11;
                sh->A12[level] = A12;
                sh->A22[level] = A22;
                sh->Dinv[level] = D;
                // bit 2: the residual sums of this level fit in int32 (LK_NARROW_MAX)
                sh->flag[level] = (ok ? 0 : 2) | ((q11 <= LK_NARROW_MAX && q22 <= LK_NARROW_MAX) ? 4 : 0);
            }
        }
    }
    __syncthreads();

    const LKStatic L = lk_static<4>();
    bool status = true;
    float nx = 0.f, ny = 0.f;
    for (int level = prm.nlev - 1; level >= 0; --level) {
        const float sc = __int_as_float((127 - level) << 23);
        if (level == prm.nlev - 1) {
            nx = __fmul_rn(gx, sc);
            ny = __fmul_rn(gy, sc);
        } else {
            nx = __fmul_rn(nx, 2.f);
            ny = __fmul_rn(ny, 2.f);
        }
        const int lflag = sh->flag[level];
        if (lflag & 3) {
            if (level == 0) status = false;
            continue;
        }
        const bool narrow = (lflag & 4) != 0;
        const int cols = g.lv[level].w, rows = g.lv[level].h, pitch = g.lv[level].pitch;
        const float A11 = sh->A11[level], A12 = sh->A12[level], A22 = sh->A22[level], D = sh->Dinv[level];
        int tI[3] = {0, 0, 0}, tIx[3] = {0, 0, 0}, tIy[3] = {0, 0, 0};
        if (L.npx) {
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                tI[k] = sh->I[level][L.row][L.c0 + k];
                tIx[k] = sh->Ix[level][L.row][L.c0 + k];
                tIy[k] = sh->Iy[level][L.row][L.c0 + k];
            }
        }
        float cx = __fsub_rn(nx, (float)AVB_HALF), cy = __fsub_rn(ny, (float)AVB_HALF);
        float pdx = 0.f, pdy = 0.f;
        const uint8_t* imgB = pyr_level(B, g, level);
        // window tests as single unsigned compares: in range <=> -WIN <= inx < cols; inside <=> 0 <= inx, inx + 16 < cols
        const unsigned x_rng = (unsigned)(cols + AVB_WIN), y_rng = (unsigned)(rows + AVB_WIN);
        const unsigned x_in = (unsigned)max(cols - 16, 0), y_in = (unsigned)max(rows - 16, 0);
        const uint8_t* lane_base = imgB + L.row * pitch + L.c0;      // this lane's first sample for a window at (0, 0)
        unsigned P = 0, N = 0;              // this lane's 4 bytes of window row L.row and of the row below it, packed
        int pinx = INT_MIN, piny = INT_MIN; // integer window position they were fetched for
        for (int j = 0; j < prm.max_iter; ++j) {
            const int inx = __float2int_rd(cx), iny = __float2int_rd(cy);
            if ((unsigned)(inx + AVB_WIN) >= x_rng || (unsigned)(iny + AVB_WIN) >= y_rng) {
                if (level == 0) status = false;
                break;
            }
            int sb1 = 0, sb2 = 0;
            if (L.npx) {
                // Sub-pixel updates leave the integer window position where it was on most iterations after the first:
                // the packed samples are then still the right ones (team-uniform test; no address, no load, no latency)
                if (inx == pinx && iny == piny) {
                } else if ((unsigned)inx < x_in && (unsigned)iny < y_in) {  // team-uniform: window inside the level
                    // two aligned 32-bit loads per row + one funnel shift (rows are 16-byte aligned, so both rows share
                    // the byte phase; the <= 3 bytes read past the sample stay inside the row pitch / padded allocation)
                    const uint8_t* a = lane_base + iny * pitch + inx;
                    const unsigned sh8 = ((unsigned)(size_t)a & 3u) * 8u;
                    const unsigned* aw = reinterpret_cast<const unsigned*>((size_t)a & ~(size_t)3);
                    const unsigned* bw = reinterpret_cast<const unsigned*>(reinterpret_cast<const uint8_t*>(aw) + pitch);
                    const unsigned A0 = __ldg(aw), A1 = __ldg(aw + 1), B0 = __ldg(bw), B1 = __ldg(bw + 1);
                    P = __funnelshift_r(A0, A1, sh8);
                    N = __funnelshift_r(B0, B1, sh8);
                } else {                                                     // REFLECT_101 on both axes
                    const uint8_t* r0 = imgB + refl101(iny + L.row, rows) * pitch;
                    const uint8_t* r1 = imgB + refl101(iny + L.row + 1, rows) * pitch;
                    const int x0 = inx + L.c0;
                    const int c0 = refl101(x0, cols), c1 = refl101(x0 + 1, cols), c2 = refl101(x0 + 2, cols),
                              c3 = refl101(x0 + 3, cols);
                    P = (unsigned)__ldg(r0 + c0) | ((unsigned)__ldg(r0 + c1) << 8) | ((unsigned)__ldg(r0 + c2) << 16) |
                        ((unsigned)__ldg(r0 + c3) << 24);
                    N = (unsigned)__ldg(r1 + c0) | ((unsigned)__ldg(r1 + c1) << 8) | ((unsigned)__ldg(r1 + c2) << 16) |
                        ((unsigned)__ldg(r1 + c3) << 24);
                }
                pinx = inx;
                piny = iny;
                int w00, w01, w10, w11;
                lk_weights(__fsub_rn(cx, (float)inx), __fsub_rn(cy, (float)iny), w00, w01, w10, w11);
                // bilinear sample k = w00 b_k + w01 b_{k+1} + w10 n_k + w11 n_{k+1} as two 2-way 16x8-bit dot products
                const unsigned Wt = ((unsigned)w00 & 0xffffu) | ((unsigned)w01 << 16);
                const unsigned Wb = ((unsigned)w10 & 0xffffu) | ((unsigned)w11 << 16);     // w11 may be -1
                const int rnd = 1 << (LK_W_BITS - 6);
                const int j0 = dp2a_lo_su(Wt, P, dp2a_lo_su(Wb, N, rnd)) >> (LK_W_BITS - 5);
                const int j1 = dp2a_lo_su(Wt, P >> 8, dp2a_lo_su(Wb, N >> 8, rnd)) >> (LK_W_BITS - 5);
                const int j2 = dp2a_hi_su(Wt, P, dp2a_hi_su(Wb, N, rnd)) >> (LK_W_BITS - 5);
                const int d0 = j0 - tI[0], d1 = j1 - tI[1], d2 = j2 - tI[2];
                sb1 = d0 * tIx[0] + d1 * tIx[1] + d2 * tIx[2];
                sb2 = d0 * tIy[0] + d1 * tIy[1] + d2 * tIy[2];
            }
            sb1 = __reduce_add_sync(0xffffffffu, sb1);
            sb2 = __reduce_add_sync(0xffffffffu, sb2);
            if ((threadIdx.x & 31) == 0) {
                sh->red[flip][0][warp] = sb1;
                sh->red[flip][1][warp] = sb2;
            }
            __syncthreads();
            const int4 r1 = *reinterpret_cast<const int4*>(sh->red[flip][0]);
            const int4 r2 = *reinterpret_cast<const int4*>(sh->red[flip][1]);
            flip ^= 1;
            float b1, b2;
            if (narrow) {               // the totals fit in int32: wrap-around of the partial sums cancels
                b1 = __fmul_rn(__int2float_rn((r1.x + r1.y) + (r1.z + r1.w)), 9.5367431640625e-07f);
                b2 = __fmul_rn(__int2float_rn((r2.x + r2.y) + (r2.z + r2.w)), 9.5367431640625e-07f);
            } else {
                const long long q1 = ((long long)r1.x + (long long)r1.y) + ((long long)r1.z + (long long)r1.w);
                const long long q2 = ((long long)r2.x + (long long)r2.y) + ((long long)r2.z + (long long)r2.w);
                b1 = __fmul_rn(__ll2float_rn(q1), 9.5367431640625e-07f);
                b2 = __fmul_rn(__ll2float_rn(q2), 9.5367431640625e-07f);
            }
            const float dx = __fmul_rn(__fsub_rn(__fmul_rn(A12, b2), __fmul_rn(A22, b1)), D);
            const float dy = __fmul_rn(__fsub_rn(__fmul_rn(A12, b1), __fmul_rn(A11, b2)), D);
            cx = __fadd_rn(cx, dx);
            cy = __fadd_rn(cy, dy);
            nx = __fadd_rn(cx, (float)AVB_HALF);
            ny = __fadd_rn(cy, (float)AVB_HALF);
            if (lk_converged(dx, dy, prm.eps2, prm.eps2_lo, prm.eps2_hi)) break;
            if (j > 0 && lk_oscillating(dx, dy, pdx, pdy)) {
                nx = __fsub_rn(nx, __fmul_rn(dx, 0.5f));
                ny = __fsub_rn(ny, __fmul_rn(dy, 0.5f));
                break;
            }
            pdx = dx;
            pdy = dy;
        }
    }
    if (status) {
        const int fx = __float2int_rd(__fsub_rn(nx, (float)AVB_HALF)), fy = __float2int_rd(__fsub_rn(ny, (float)AVB_HALF));
        if (fx < -AVB_WIN || fx >= g.lv[0].w || fy < -AVB_WIN || fy >= g.lv[0].h) status = false;
    }
    ox = nx;
    oy = ny;
    return status;
}

template <int WPF>
__device__ __forceinline__ bool lk_track_team(const PyrView& A, const PyrView& B, const Geom& g, float px0, float py0,
                                              float gx, float gy, const LKParams& prm, LKShared* sh, int& flip, float& ox,
                                              float& oy) {
    if constexpr (WPF == 1)
        return lk_track_warp(A, B, g, px0, py0, gx, gy, prm, ox, oy);
    else
        return lk_track_coop(A, B, g, px0, py0, gx, gy, prm, sh, flip, ox, oy);
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

// One feature through [temporal LK ->] stereo forward LK -> stereo backward LK -> filters, with ONE inlined copy
// of the LK body (the three passes run through the same code; three copies made the kernel 99 KB and the
// instruction fetch its top stall).  Team-uniform.
//   temporal pass  FeatureTracker.track_features steps 3-6 (feature_tracker.py:85-123): LK prev cam0 -> cur cam0
//                  from the gyro-predicted guess, then the image-bounds cull (> W-1 rule, B6)
//   stereo passes  StereoMatcher.stereo_match (stereo_matcher.py:33-115): the cam0 model is used for both cameras
//                  (B3); backward-LK status is ignored (B5); the epipolar error keeps only the x-term (B4)
struct ChainResult {
    bool tracked;                   // temporal LK ok and inside the image (always true without the temporal pass)
    bool matched;                   // stereo inlier
    float cx, cy;                   // cam0 position in the current frame
    float x1, y1;                   // cam1 position
    double u0, v0, u1, v1;          // matched only: normalized coordinates as FeaturePublisher.publish undistorts them
                                    // (cam0 point with the cam0 model, cam1 point with the cam1 model,
                                    // feature_publisher.py:104-113), precomputed here where the work is spread over
                                    // the whole GPU instead of the one CTA per stream of k_finish
};

template <int WPF>
__device__ __forceinline__ ChainResult feature_chain(const Geom& g, const DevState& d, int s, int parity, bool temporal,
                                                     float x, float y, float gx, float gy, LKShared* sh) {
    int flip = 0;
    LKParams prm;
    prm.nlev = g.nlev;
    prm.max_iter = g.max_iter;
    prm.min_eig = g.min_eig;
    prm.eps2 = g.eps2;
    prm.eps2_lo = g.eps2_lo;
    prm.eps2_hi = g.eps2_hi;
    prm.eig_accept = g.eig_accept;
    ChainResult r;
    r.tracked = true;
    r.matched = false;
    r.cx = x;
    r.cy = y;
    r.x1 = 0.f;
    r.y1 = 0.f;
    r.u0 = r.v0 = r.u1 = r.v1 = 0.0;
    float projy = 0.f;
    float ax = x, ay = y, bx = gx, by = gy;
#pragma unroll 1
    for (int pass = temporal ? 0 : 1; pass < 3; ++pass) {
        int slotA, slotB;
        if (pass == 0) {
            slotA = SLOT(0, parity ^ 1);
            slotB = SLOT(0, parity);
        } else if (pass == 1) {
            slotA = SLOT(0, parity);
            slotB = SLOT(1, parity);
            double ux, uy, pu, pv;                              // infinite-depth prediction (stereo_matcher.py:49-61)
            undistort_pt(g.cam0, (double)r.cx, (double)r.cy, g.R01, ux, uy);
            distort_pt(g.cam0, (double)(float)ux, (double)(float)uy, pu, pv);
            ax = r.cx;
            ay = r.cy;
            bx = (float)pu;
            by = (float)pv;
            projy = by;
        } else {
            slotA = SLOT(1, parity);
            slotB = SLOT(0, parity);
            ax = r.x1;
            ay = r.y1;
            bx = r.cx;
            by = r.cy;
        }
        const PyrView A = pyr_view(d, g, s, slotA), B = pyr_view(d, g, s, slotB);
        float ox, oy;
        const bool st = lk_track_team<WPF>(A, B, g, ax, ay, bx, by, prm, sh, flip, ox, oy);
        if (pass == 0) {
            r.cx = ox;
            r.cy = oy;
            r.tracked = st && !(ox < 0.f || ox > (float)(g.W - 1) || oy < 0.f || oy > (float)(g.H - 1));
            if (!r.tracked) break;
        } else if (pass == 1) {
            r.x1 = ox;
            r.y1 = oy;
            if (!st) break;
        } else {
            const float ex = __fsub_rn(r.cx, ox), ey = __fsub_rn(r.cy, oy);
            const float err = __fsqrt_rn(__fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey)));
            const float disp = fabsf(__fsub_rn(projy, r.y1));
            bool ok = (err < 3.f) && (disp < 20.f);
            ok = ok && !(r.x1 < 0.f || r.x1 >= (float)g.W || r.y1 < 0.f || r.y1 >= (float)g.H);
            if (ok) {
                // Three undistortions are due: (cx, cy) and (x1, y1) through the cam0 model for the epipolar test (B3)
                // and (x1, y1) through the cam1 model for the publisher.  Lanes 0, 1, 2 take one each (one FP64
                // chain for the warp instead of three), the results come back by shuffle.
                const int l = threadIdx.x & 31;
                const bool c1 = l == 2;
                CamModel cm;
                cm.fx = c1 ? g.cam1.fx : g.cam0.fx;
                cm.fy = c1 ? g.cam1.fy : g.cam0.fy;
                cm.cx = c1 ? g.cam1.cx : g.cam0.cx;
                cm.cy = c1 ? g.cam1.cy : g.cam0.cy;
                cm.k1 = c1 ? g.cam1.k1 : g.cam0.k1;
                cm.k2 = c1 ? g.cam1.k2 : g.cam0.k2;
                cm.p1 = c1 ? g.cam1.p1 : g.cam0.p1;
                cm.p2 = c1 ? g.cam1.p2 : g.cam0.p2;
                double ux, uy;
                undistort_pt(cm, (double)(l == 0 ? r.cx : r.x1), (double)(l == 0 ? r.cy : r.y1), nullptr, ux, uy);
                const double a0 = __shfl_sync(0xffffffffu, ux, 0), b0 = __shfl_sync(0xffffffffu, uy, 0);
                const double a1 = __shfl_sync(0xffffffffu, ux, 1);
                const double u0x = (double)(float)a0, u0y = (double)(float)b0, u1x = (double)(float)a1;
                const double l0 = g.E[0] * u0x + g.E[1] * u0y + g.E[2];
                const double l1 = g.E[3] * u0x + g.E[4] * u0y + g.E[5];
                const double epi = fabs(u1x * l0) / sqrt(l0 * l0 + l1 * l1);
                ok = !(epi > g.epi_thr);
                r.u0 = a0;
                r.v0 = b0;
                r.u1 = __shfl_sync(0xffffffffu, ux, 2);
                r.v1 = __shfl_sync(0xffffffffu, uy, 2);
            }
            r.matched = ok;
        }
    }
    return r;
}
