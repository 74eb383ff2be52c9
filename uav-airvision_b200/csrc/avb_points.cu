// K3/K4 kernels: one warp per point.  The per-frame kernels work directly on the grid-ordered tables
// (slot i of cell c is point c*gmax + i; slots >= count[c] exit at once), so no compaction pass sits
// between the reference's stages: "select()" (image_processing/utils.py:10-11) becomes a mask.
#include "avb_lk.cuh"

#define WARPS_PER_BLOCK 4

__device__ __forceinline__ LKParams make_lk(const Geom& g) {
    LKParams p;
    p.nlev = g.nlev;
    p.max_iter = g.max_iter;
    p.min_eig = g.min_eig;
    p.eps2 = g.eps2;
    return p;
}

// FeatureTracker.track_features, steps 3-8 (feature_tracker.py:85-133) for every previous feature:
// gyro prediction (K R K^-1) -> temporal LK -> image-bounds cull (> W-1 rule, B6) -> stereo match.
__global__ void __launch_bounds__(32 * WARPS_PER_BLOCK) k_track(Geom g, DevState d, int parity) {
    const int s = blockIdx.y;
    const int wi = blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (wi >= g.NMAX) return;
    const GridTable prev = d.grid[parity ^ 1];
    const size_t base = (size_t)s * g.NMAX;
    const int cell = wi / g.gmax, slot = wi - cell * g.gmax;
    if (slot >= prev.count[s * g.NC + cell]) {
        if (lane == 0) d.t_cell[base + wi] = -1;
        return;
    }
    const float2 p = prev.p0[base + wi];
    const double* H = frame_H(d, g, s, parity);
    const double hx = H[0] * (double)p.x + H[1] * (double)p.y + H[2];
    const double hy = H[3] * (double)p.x + H[4] * (double)p.y + H[5];
    const double hz = H[6] * (double)p.x + H[7] * (double)p.y + H[8];
    const float gx = (float)(hx / hz), gy = (float)(hy / hz);

    const LKParams prm = make_lk(g);
    const PyrView Pprev = pyr_view(d, g, s, SLOT(0, parity ^ 1)), Pcur = pyr_view(d, g, s, SLOT(0, parity)),
                  P1 = pyr_view(d, g, s, SLOT(1, parity));
    float cx, cy;
    bool keep = lk_track_warp(Pprev, Pcur, g, p.x, p.y, gx, gy, prm, cx, cy);
    keep = keep && !(cx < 0.f || cx > (float)(g.W - 1) || cy < 0.f || cy > (float)(g.H - 1));
    int* cnt = d.counters + s * 8;
    int new_cell = -1;
    float x1 = 0.f, y1 = 0.f;
    if (keep) {
        const bool ok = stereo_match_warp(Pcur, P1, g, prm, cx, cy, x1, y1);
        if (ok) new_cell = (int)__fdiv_rn(cy, (float)g.gh) * g.cols + (int)__fdiv_rn(cx, (float)g.gw);
        if (lane == 0) {
            atomicAdd(&cnt[1], 1);
            if (ok) atomicAdd(&cnt[2], 1);
        }
    }
    if (lane == 0) {
        atomicAdd(&cnt[0], 1);
        d.t_p0[base + wi] = make_float2(cx, cy);
        d.t_p1[base + wi] = make_float2(x1, y1);
        d.t_cell[base + wi] = new_cell;
    }
}

// stereo_match of the new-feature candidates (feature_adder.py:79-80)
__global__ void __launch_bounds__(32 * WARPS_PER_BLOCK) k_stereo_candidates(Geom g, DevState d, int parity) {
    const int s = blockIdx.y;
    const int wi = blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
    if (wi >= g.NMAX) return;
    const int cell = wi / g.gmax, j = wi - cell * g.gmax;
    if (j >= d.c_count[s * g.NC + cell]) return;
    const size_t idx = (size_t)s * g.NMAX + wi;
    int resp, x, y;
    kp_decode(d.c_key[idx], g.W, resp, x, y);
    const LKParams prm = make_lk(g);
    const PyrView P0 = pyr_view(d, g, s, SLOT(0, parity)), P1 = pyr_view(d, g, s, SLOT(1, parity));
    float x1, y1;
    const bool ok = stereo_match_warp(P0, P1, g, prm, (float)x, (float)y, x1, y1);
    if ((threadIdx.x & 31) == 0) {
        d.c_p1[idx] = make_float2(x1, y1);
        d.c_ok[idx] = ok ? 1 : 0;
    }
}

// frame 0: stereo_match of EVERY FAST keypoint before ranking (feature_initializer.py:52-55, B15)
__global__ void __launch_bounds__(32 * WARPS_PER_BLOCK) k_stereo_buckets(Geom g, DevState d, int parity) {
    const int s = blockIdx.z, cell = blockIdx.y;
    const int j = blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
    const int n = min(d.kp_count[s * g.NC + cell], g.KPC);
    if (j >= n) return;
    const size_t idx = ((size_t)s * g.NC + cell) * g.KPC + j;
    int resp, x, y;
    kp_decode(d.kp_key[idx], g.W, resp, x, y);
    const LKParams prm = make_lk(g);
    const PyrView P0 = pyr_view(d, g, s, SLOT(0, parity)), P1 = pyr_view(d, g, s, SLOT(1, parity));
    float x1, y1;
    const bool ok = stereo_match_warp(P0, P1, g, prm, (float)x, (float)y, x1, y1);
    if ((threadIdx.x & 31) == 0) {
        d.kp_p1[idx] = make_float2(x1, y1);
        d.kp_ok[idx] = ok ? 1 : 0;
    }
}

// ---- flat lists for the per-stage entry points -------------------------------------------------
__global__ void __launch_bounds__(32 * WARPS_PER_BLOCK) k_klt_points(Geom g, DevState d, int s, int slot_from, int slot_to,
                                                                    const float2* prev, const float2* guess, int n, float2* out,
                                                                    uint8_t* status) {
    const int wi = blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
    if (wi >= n) return;
    const LKParams prm = make_lk(g);
    const PyrView A = pyr_view(d, g, s, slot_from), B = pyr_view(d, g, s, slot_to);
    float ox, oy;
    const bool st = lk_track_warp(A, B, g, prev[wi].x, prev[wi].y, guess[wi].x, guess[wi].y, prm, ox, oy);
    if ((threadIdx.x & 31) == 0) {
        out[wi] = make_float2(ox, oy);
        status[wi] = st ? 1 : 0;
    }
}

__global__ void __launch_bounds__(32 * WARPS_PER_BLOCK) k_stereo_points(Geom g, DevState d, int s, int parity, const float2* p0,
                                                                       int n, float2* p1, uint8_t* ok) {
    const int wi = blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
    if (wi >= n) return;
    const LKParams prm = make_lk(g);
    const PyrView P0 = pyr_view(d, g, s, SLOT(0, parity)), P1 = pyr_view(d, g, s, SLOT(1, parity));
    float x1, y1;
    const bool r = stereo_match_warp(P0, P1, g, prm, p0[wi].x, p0[wi].y, x1, y1);
    if ((threadIdx.x & 31) == 0) {
        p1[wi] = make_float2(x1, y1);
        ok[wi] = r ? 1 : 0;
    }
}

__global__ void k_undistort(CamModel cam, const double* xy, int n, const double* R, int has_R, int f32_io, int distort,
                            double* out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double x = xy[2 * i], y = xy[2 * i + 1], ox, oy;
    if (f32_io) {
        x = (double)(float)x;
        y = (double)(float)y;
    }
    if (distort)
        distort_pt(cam, x, y, ox, oy);
    else
        undistort_pt(cam, x, y, has_R ? R : nullptr, ox, oy);
    if (f32_io) {
        ox = (double)(float)ox;
        oy = (double)(float)oy;
    }
    out[2 * i] = ox;
    out[2 * i + 1] = oy;
}

void launch_track(const Geom& g, const DevState& d, int parity, cudaStream_t st) {
    dim3 grid((g.NMAX + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK, g.S);
    k_track<<<grid, 32 * WARPS_PER_BLOCK, 0, st>>>(g, d, parity);
}
void launch_stereo_candidates(const Geom& g, const DevState& d, int parity, cudaStream_t st) {
    dim3 grid((g.NMAX + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK, g.S);
    k_stereo_candidates<<<grid, 32 * WARPS_PER_BLOCK, 0, st>>>(g, d, parity);
}
void launch_stereo_buckets(const Geom& g, const DevState& d, int parity, cudaStream_t st) {
    dim3 grid((g.KPC + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK, g.NC, g.S);
    k_stereo_buckets<<<grid, 32 * WARPS_PER_BLOCK, 0, st>>>(g, d, parity);
}
void launch_klt_points(const Geom& g, const DevState& d, int s, int slot_from, int slot_to, const float2* prev,
                       const float2* guess, int n, float2* out, uint8_t* status, cudaStream_t st) {
    if (n <= 0) return;
    k_klt_points<<<(n + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK, 32 * WARPS_PER_BLOCK, 0, st>>>(g, d, s, slot_from, slot_to, prev,
                                                                                                 guess, n, out, status);
}
void launch_stereo_points(const Geom& g, const DevState& d, int s, int parity, const float2* p0, int n, float2* p1,
                          uint8_t* ok, cudaStream_t st) {
    if (n <= 0) return;
    k_stereo_points<<<(n + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK, 32 * WARPS_PER_BLOCK, 0, st>>>(g, d, s, parity, p0, n, p1, ok);
}
void launch_undistort(const CamModel& cam, const double* xy, int n, const double* R, int has_R, int f32_io, int distort,
                      double* out, cudaStream_t st) {
    if (n <= 0) return;
    k_undistort<<<(n + 127) / 128, 128, 0, st>>>(cam, xy, n, R, has_R, f32_io, distort, out);
}
