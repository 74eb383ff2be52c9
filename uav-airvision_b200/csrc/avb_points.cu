// K3/K4 kernels: one warp per point.  The per-frame kernels work directly on the grid-ordered tables
// (slot i of cell c is point c*gmax + i; slots >= count[c] exit at once), so no compaction pass sits
// between the reference's stages: "select()" (image_processing/utils.py:10-11) becomes a mask.
#include "avb_lk.cuh"

#define WARPS_PER_BLOCK 4
#ifndef AVB_WPF1_BLOCKS
#define AVB_WPF1_BLOCKS 5            // resident CTAs per SM the 1-warp LK kernels are compiled for: 96 registers, no
                                     // spills.  Measured at 64 streams (frames/s): 8 CTAs (64 regs, 244 B spilled)
                                     // 50,981; 6 (80 regs) 52,243; 5 (96 regs) 53,253; 4 (127 regs) 52,544.  Again
                                     // after the template rewrite (r01h): 8 -> 58,182; 6 -> 59,130; 5 -> 61,541;
                                     // 4 -> 59,260 (build variants with -DAVB_WPF1_BLOCKS=n via AVB_EXTRA_NVCC)
#endif

// Team decomposition of a 128-thread block: WPF = 1 -> four features per block (one warp each), WPF = 4 -> one
// feature per block.  Returns the feature index; `sh` points at the team's shared scratch.
template <int WPF>
__device__ __forceinline__ int team_index(int bx) {
    return WPF == 1 ? bx * WARPS_PER_BLOCK + (threadIdx.x >> 5) : bx;
}

// FeatureTracker.track_features, steps 3-8 (feature_tracker.py:85-133) for every previous feature:
// gyro prediction (K R K^-1) -> temporal LK -> image-bounds cull (> W-1 rule, B6) -> stereo match.
template <int WPF, bool RANSAC>
__global__ void __launch_bounds__(32 * WARPS_PER_BLOCK, WPF == 1 ? AVB_WPF1_BLOCKS : 4) k_track(const __grid_constant__ Geom g, const __grid_constant__ DevState d,
                                                                int parity) {
    __shared__ LKShared sh;
    pdl_wait();
    pdl_launch_dependents();
    const int s = blockIdx.y;
    const int wi = team_index<WPF>(blockIdx.x);
    const bool lead = WPF == 1 ? (threadIdx.x & 31) == 0 : threadIdx.x == 0;
    if (wi >= g.NMAX) return;
    const GridTable prev = d.grid[parity ^ 1];
    const size_t base = (size_t)s * g.NMAX;
    const int cell = wi / g.gmax, slot = wi - cell * g.gmax;
    if (slot >= prev.count[s * g.NC + cell]) {
        if (lead) d.t_cell[base + wi] = -1;
        return;
    }
    const float2 p = prev.p0[base + wi];
    const double* H = frame_H(d, g, s, parity);
    const double hx = H[0] * (double)p.x + H[1] * (double)p.y + H[2];
    const double hy = H[3] * (double)p.x + H[4] * (double)p.y + H[5];
    const double hz = H[6] * (double)p.x + H[7] * (double)p.y + H[8];
    const float gx = (float)(hx / hz), gy = (float)(hy / hz);

    const ChainResult r = feature_chain<WPF>(g, d, s, parity, true, p.x, p.y, gx, gy, &sh);
    const bool matched = r.matched;
    if (lead) {
        int* cnt = d.counters + s * 8;
        int new_cell = -1;
        if (r.tracked) {
            if (r.matched) new_cell = (int)__fdiv_rn(r.cy, (float)g.gh) * g.cols + (int)__fdiv_rn(r.cx, (float)g.gw);
            atomicAdd(&cnt[1], 1);
            if (r.matched) atomicAdd(&cnt[2], 1);
        }
        atomicAdd(&cnt[0], 1);
        d.t_p0[base + wi] = make_float2(r.cx, r.cy);
        d.t_p1[base + wi] = make_float2(r.x1, r.y1);
        d.t_cell[base + wi] = new_cell;
        if (new_cell >= 0) d.t_und[base + wi] = make_double4(r.u0, r.v0, r.u1, r.v1);
    }
    if (RANSAC && matched && (WPF == 1 || threadIdx.x < 32)) {      // a template flag: the plain instantiation keeps its 96 registers without spills
        // k_ransac's inputs that do not depend on other features are made here, where the work is spread over the GPU
        // (that kernel runs one CTA per stream: at 2000 points the FP64 undistortions were half of its 41 us): the
        // previous position of each camera, undistorted and rotated by the gyro prediction of that camera
        // (oracle/ransac.py: undistort(prev, R_p_c)).  Lane 0 cam0, lane 1 cam1; the current positions are t_und.
        const int l = threadIdx.x & 31;
        const bool c1 = l == 1;
        CamModel cm;
        cm.fx = c1 ? g.cam1.fx : g.cam0.fx;
        cm.fy = c1 ? g.cam1.fy : g.cam0.fy;
        cm.cx = c1 ? g.cam1.cx : g.cam0.cx;
        cm.cy = c1 ? g.cam1.cy : g.cam0.cy;
        cm.k1 = c1 ? g.cam1.k1 : g.cam0.k1;
        cm.k2 = c1 ? g.cam1.k2 : g.cam0.k2;
        cm.p1 = c1 ? g.cam1.p1 : g.cam0.p1;
        cm.p2 = c1 ? g.cam1.p2 : g.cam0.p2;
        const float2 q = c1 ? prev.p1[base + wi] : p;
        double ux, uy;
        undistort_pt(cm, (double)q.x, (double)q.y, H + 9 + (c1 ? 9 : 0), ux, uy);
        const float x1u = __shfl_sync(0xffffffffu, (float)ux, 1), y1u = __shfl_sync(0xffffffffu, (float)uy, 1);
        if (l == 0) d.r_prev[base + wi] = make_float4((float)ux, (float)uy, x1u, y1u);
    }
}

// stereo_match of the new-feature candidates (feature_adder.py:79-80)
// round < 0: every candidate in one launch (latency mode).  Throughput mode splits the work: the adder keeps only the
// first `gmin` stereo inliers of a cell's list (feature_adder.py:102-108, B8), so round 0 matches list positions
// < gmin and round 1 the rest, and only in cells where round 0 left a deficit (a position that failed or a list
// shorter than it looks) -- typically a few percent of the cells.  Unmatched tail positions keep a stale c_ok, which
// cannot matter: with gmin inliers ahead of them they rank >= gmin and are never adopted nor counted.
template <int WPF>
__global__ void __launch_bounds__(32 * WARPS_PER_BLOCK, WPF == 1 ? AVB_WPF1_BLOCKS : 4) k_stereo_candidates(const __grid_constant__ Geom g,
                                                                            const __grid_constant__ DevState d, int parity,
                                                                            int round) {
    __shared__ LKShared sh;
    pdl_wait();
    pdl_launch_dependents();
    const int s = blockIdx.y;
    const int t = team_index<WPF>(blockIdx.x);              // teams are dealt densely over the positions of this round
    const bool lead = WPF == 1 ? (threadIdx.x & 31) == 0 : threadIdx.x == 0;
    const int per_cell = round < 0 ? g.gmax : (round == 0 ? g.gmin : g.gmax - g.gmin);
    if (t >= g.NC * per_cell) return;
    const int cell = t / per_cell, j = t - cell * per_cell + (round == 1 ? g.gmin : 0);
    if (j >= d.c_count[s * g.NC + cell]) return;
    const int wi = cell * g.gmax + j;
    const size_t idx = (size_t)s * g.NMAX + wi;
    if (round == 1) {
        int inl = 0;
        for (int k = 0; k < g.gmin; ++k) inl += d.c_ok[idx - j + k];   // count > j >= gmin: all gmin positions were matched
        if (inl == g.gmin) return;
    }
    int resp, x, y;
    kp_decode(d.c_key[idx], g.W, resp, x, y);
    const ChainResult r = feature_chain<WPF>(g, d, s, parity, false, (float)x, (float)y, 0.f, 0.f, &sh);
    if (lead) {
        d.c_p1[idx] = make_float2(r.x1, r.y1);
        d.c_ok[idx] = r.matched ? 1 : 0;
        if (r.matched) d.c_und[idx] = make_double4(r.u0, r.v0, r.u1, r.v1);
    }
}

// Speculative stereo_match of the spec_k strongest keypoints of every cell (k_select mode 2) on the side stream beside
// k_track.  Same arithmetic as k_stereo_candidates (both lane mappings give bit-identical results).  WPF = 4: the chain of
// one candidate is what this branch's length is made of (all candidates are resident at once), and four warps walk it
// faster than one.
template <int WPF>
__global__ void __launch_bounds__(32 * WARPS_PER_BLOCK, WPF == 1 ? AVB_WPF1_BLOCKS : 4) k_spec_match(const __grid_constant__ Geom g,
                                                                                    const __grid_constant__ DevState d, int parity) {
    __shared__ LKShared sh;
    const int s = blockIdx.y;
    const int t = team_index<WPF>(blockIdx.x);
    const bool lead = WPF == 1 ? (threadIdx.x & 31) == 0 : threadIdx.x == 0;
    if (t >= g.NC * g.spec_k) return;
    // rank-major: the teams of every cell's strongest candidate come first, so that if the launch does not fit the GPU at
    // once, the candidates most likely to be adopted are not the ones left waiting
    const int j = t / g.NC, cell = t - j * g.NC;
    if (j >= d.s_n[s * g.NC + cell]) return;
    const size_t idx = ((size_t)s * g.NC + cell) * g.spec_k + j;
    int resp, x, y;
    kp_decode(d.s_key[idx], g.W, resp, x, y);
    const ChainResult r = feature_chain<WPF>(g, d, s, parity, false, (float)x, (float)y, 0.f, 0.f, &sh);
    if (lead) {
        d.s_p1[idx] = make_float2(r.x1, r.y1);
        d.s_ok[idx] = r.matched ? 1 : 0;
        if (r.matched) d.s_und[idx] = make_double4(r.u0, r.v0, r.u1, r.v1);
    }
}

// frame 0: stereo_match of EVERY FAST keypoint before ranking (feature_initializer.py:52-55, B15)
__global__ void __launch_bounds__(32 * WARPS_PER_BLOCK, 8) k_stereo_buckets(const __grid_constant__ Geom g,
                                                                         const __grid_constant__ DevState d, int parity) {
    const int s = blockIdx.z, cell = blockIdx.y;
    const int j = blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
    const int n = min(d.kp_count[s * g.NC + cell], g.KPC);
    if (j >= n) return;
    const size_t idx = ((size_t)s * g.NC + cell) * g.KPC + j;
    int resp, x, y;
    kp_decode(d.kp_key[idx], g.W, resp, x, y);
    const ChainResult r = feature_chain<1>(g, d, s, parity, false, (float)x, (float)y, 0.f, 0.f, nullptr);
    if ((threadIdx.x & 31) == 0) {
        d.kp_p1[idx] = make_float2(r.x1, r.y1);
        d.kp_ok[idx] = r.matched ? 1 : 0;
    }
}

// ---- flat lists for the per-stage entry points -------------------------------------------------
__global__ void __launch_bounds__(32 * WARPS_PER_BLOCK) k_klt_points(const __grid_constant__ Geom g, const __grid_constant__ DevState d,
                                                                    int s, int slot_from, int slot_to, const float2* prev,
                                                                    const float2* guess, int n, float2* out, uint8_t* status,
                                                                    int wpf) {
    __shared__ LKShared sh;
    const int wi = wpf == 1 ? blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5) : blockIdx.x;
    if (wi >= n) return;
    LKParams prm;
    prm.nlev = g.nlev;
    prm.max_iter = g.max_iter;
    prm.min_eig = g.min_eig;
    prm.eps2 = g.eps2;
    prm.eps2_lo = g.eps2_lo;
    prm.eps2_hi = g.eps2_hi;
    prm.eig_accept = g.eig_accept;
    const PyrView A = pyr_view(d, g, s, slot_from), B = pyr_view(d, g, s, slot_to);
    float ox, oy;
    int flip = 0;
    bool st;
    if (wpf == 1)
        st = lk_track_team<1>(A, B, g, prev[wi].x, prev[wi].y, guess[wi].x, guess[wi].y, prm, nullptr, flip, ox, oy);
    else
        st = lk_track_team<4>(A, B, g, prev[wi].x, prev[wi].y, guess[wi].x, guess[wi].y, prm, &sh, flip, ox, oy);
    if ((wpf == 1 ? (threadIdx.x & 31) : threadIdx.x) == 0) {
        out[wi] = make_float2(ox, oy);
        status[wi] = st ? 1 : 0;
    }
}

template <int WPF>
__global__ void __launch_bounds__(32 * WARPS_PER_BLOCK) k_stereo_points(const __grid_constant__ Geom g,
                                                                       const __grid_constant__ DevState d, int s, int parity,
                                                                       const float2* p0, int n, float2* p1, uint8_t* ok) {
    __shared__ LKShared sh;
    const int wi = team_index<WPF>(blockIdx.x);
    if (wi >= n) return;
    const ChainResult r = feature_chain<WPF>(g, d, s, parity, false, p0[wi].x, p0[wi].y, 0.f, 0.f, &sh);
    if ((WPF == 1 ? (threadIdx.x & 31) : threadIdx.x) == 0) {
        p1[wi] = make_float2(r.x1, r.y1);
        ok[wi] = r.matched ? 1 : 0;
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

static inline int teams_grid(int n, int wpf) { return wpf == 1 ? (n + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK : n; }

void launch_track(const Geom& g, const DevState& d, int parity, cudaStream_t st) {
    dim3 grid(teams_grid(g.NMAX, g.wpf), g.S);
    if (g.wpf == 1) {
        if (g.ransac) launch_k(k_track<1, true>, grid, dim3(32 * WARPS_PER_BLOCK), 0, st, g_avb_pdl != 0, g, d, parity);
        else launch_k(k_track<1, false>, grid, dim3(32 * WARPS_PER_BLOCK), 0, st, g_avb_pdl != 0, g, d, parity);
    } else {
        if (g.ransac) launch_k(k_track<4, true>, grid, dim3(32 * WARPS_PER_BLOCK), 0, st, g_avb_pdl != 0, g, d, parity);
        else launch_k(k_track<4, false>, grid, dim3(32 * WARPS_PER_BLOCK), 0, st, g_avb_pdl != 0, g, d, parity);
    }
}
// Two rounds save ~40 % of the matching work but cost a second chain of latency: worth it only when the candidates
// fill the GPU several times over (throughput-bound), not when they fit in a wave or two.
int avb_candidate_rounds(const Geom& g) {
    return (g.wpf == 1 && g.gmin < g.gmax && (long long)g.S * g.NMAX > 2 * 148 * AVB_WPF1_BLOCKS * WARPS_PER_BLOCK) ? 2 : 1;
}

void launch_stereo_candidates(const Geom& g, const DevState& d, int parity, cudaStream_t st) {
    dim3 grid(teams_grid(g.NMAX, g.wpf), g.S);
    if (g.cand_rounds == 2) {
        launch_k(k_stereo_candidates<1>, dim3(teams_grid(g.NC * g.gmin, 1), g.S), dim3(32 * WARPS_PER_BLOCK), 0, st, g_avb_pdl != 0,
                 g, d, parity, 0);
        launch_k(k_stereo_candidates<1>, dim3(teams_grid(g.NC * (g.gmax - g.gmin), 1), g.S), dim3(32 * WARPS_PER_BLOCK), 0, st,
                 g_avb_pdl != 0, g, d, parity, 1);
    } else if (g.wpf == 1) {
        launch_k(k_stereo_candidates<1>, grid, dim3(32 * WARPS_PER_BLOCK), 0, st, g_avb_pdl != 0, g, d, parity, -1);
    } else {
        launch_k(k_stereo_candidates<4>, grid, dim3(32 * WARPS_PER_BLOCK), 0, st, g_avb_pdl != 0, g, d, parity, -1);
    }
}
void launch_spec_match(const Geom& g, const DevState& d, int parity, cudaStream_t st) {
    if (g.spec_wpf == 4)
        k_spec_match<4><<<dim3(teams_grid(g.NC * g.spec_k, 4), g.S), 32 * WARPS_PER_BLOCK, 0, st>>>(g, d, parity);
    else
        k_spec_match<1><<<dim3(teams_grid(g.NC * g.spec_k, 1), g.S), 32 * WARPS_PER_BLOCK, 0, st>>>(g, d, parity);
}
void launch_stereo_buckets(const Geom& g, const DevState& d, int parity, cudaStream_t st) {
    dim3 grid((g.KPC + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK, g.NC, g.S);
    k_stereo_buckets<<<grid, 32 * WARPS_PER_BLOCK, 0, st>>>(g, d, parity);
}
void launch_klt_points(const Geom& g, const DevState& d, int s, int slot_from, int slot_to, const float2* prev,
                       const float2* guess, int n, float2* out, uint8_t* status, cudaStream_t st) {
    if (n <= 0) return;
    const int wpf = n <= 2048 ? g.wpf : 1;
    k_klt_points<<<teams_grid(n, wpf), 32 * WARPS_PER_BLOCK, 0, st>>>(g, d, s, slot_from, slot_to, prev, guess, n, out, status, wpf);
}
void launch_stereo_points(const Geom& g, const DevState& d, int s, int parity, const float2* p0, int n, float2* p1,
                          uint8_t* ok, cudaStream_t st) {
    if (n <= 0) return;
    const int wpf = n <= 2048 ? g.wpf : 1;
    if (wpf == 1)
        k_stereo_points<1><<<teams_grid(n, 1), 32 * WARPS_PER_BLOCK, 0, st>>>(g, d, s, parity, p0, n, p1, ok);
    else
        k_stereo_points<4><<<teams_grid(n, 4), 32 * WARPS_PER_BLOCK, 0, st>>>(g, d, s, parity, p0, n, p1, ok);
}
void launch_undistort(const CamModel& cam, const double* xy, int n, const double* R, int has_R, int f32_io, int distort,
                      double* out, cudaStream_t st) {
    if (n <= 0) return;
    k_undistort<<<(n + 127) / 128, 128, 0, st>>>(cam, xy, n, R, has_R, f32_io, distort, out);
}

// Lazy module loading (the CUDA 12 default) loads a kernel on its first launch: ~0.2 ms each, which frame 0 of a stream
// would pay for the kernels only it uses.  cudaFuncGetAttributes loads the function now (called from avb_create).
int avb_preload_points() {
    cudaFuncAttributes a;
    cudaError_t e = cudaSuccess;
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, k_track<1, false>);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, k_track<4, false>);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, k_track<1, true>);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, k_track<4, true>);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, k_stereo_candidates<1>);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, k_stereo_candidates<4>);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, k_spec_match<1>);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, k_spec_match<4>);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, k_stereo_buckets);
    return e == cudaSuccess ? 0 : -1;
}
