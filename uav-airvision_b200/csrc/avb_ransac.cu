// k_ransac: undistortion + two-point RANSAC outlier rejection between the previous and the current frame, both
// cameras, fused in one kernel (the "K4" of the north star).  NOT on the reference's path: uav-airvision hard-codes the
// tracker's RANSAC masks to all-ones (feature_tracker.py:135-136); with avb_config.ransac = 0 this kernel is never
// launched and the front end reproduces the reference.  With ransac = 1 it implements the algorithm the reference
// descends from (twoPointRansac of the stereo MSCKF-VIO image processor) exactly as
// oracle/ransac.py restates it: same counter-based draws, same float64 operation order, same summation trees, so
// the masks are compared bit for bit.
//
// One 512-thread CTA per stream; threads 0-255 work on camera 0, 256-511 on camera 1 (named barriers keep the two
// halves independent), then the halves meet and a feature survives iff both cameras call it an inlier
// (feature_tracker.py:141).  Per camera:
//   1. ordered compaction of the stereo-matched survivors of k_track (t_cell >= 0) -> point list in table order
//   2. per point: undistort the previous position with R_p_c (gyro-compensated) and the current position, f64
//      arithmetic rounded to f32 like cv2.undistortPoints does for float32 input
//   3. rescale to mean norm sqrt(2), pre-reject |d| > 50 px, pure-rotation shortcut
//   4. 7 hypotheses (success probability 0.99, inlier ratio 0.7): two drawn constraints, closed-form 2x2 solve;
//      every thread scores its points against all 7 models at once (7 x N scoring matrix, one pass)
//   5. the first hypothesis with the largest inlier set (>= 0.2 N) gives the mask
#include "avb_lk.cuh"

#define RS_HALF 256
#define RS_WARPS (RS_HALF / 32)
#define RS_MAX_HYP 8

struct RansacScratch {             // per (stream, camera), NMAX entries each
    int* idx;                      // compacted table slots of the points
    float4* und;                   // undistorted (prev rotated, cur): x1 y1 x2 y2, f32
    int* raw_idx;                  // positions of the points that pass the pre-rejection
    uint8_t* bits;                 // bit 7: raw; bits 0..6: inlier of hypothesis h; finally 0/1 mask
};

struct RsShared {
    double wsum[RS_WARPS];
    int wcnt[RS_WARPS];
    int hcnt[RS_WARPS][RS_MAX_HYP];
    double model[RS_MAX_HYP][3];
};

__device__ __forceinline__ void bar_half(int cam) { asm volatile("bar.sync %0, %1;" ::"r"(cam + 1), "r"(RS_HALF) : "memory"); }

__device__ __forceinline__ unsigned long long splitmix64(unsigned long long z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

// Sum over all 256 threads of a half in the order oracle/ransac.py::tree_sum spells out.
__device__ __forceinline__ double half_tree_sum(double part, RsShared& sh, int cam, int wq, int lane) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part = __dadd_rn(part, __shfl_xor_sync(0xffffffffu, part, o));
    bar_half(cam);                 // previous readers of wsum are done
    if (lane == 0) sh.wsum[wq] = part;
    bar_half(cam);
    double total = sh.wsum[0];
#pragma unroll
    for (int w = 1; w < RS_WARPS; ++w) total = __dadd_rn(total, sh.wsum[w]);
    return total;
}

// Ordered compaction: out[rank] = k for every k in [0, n) with flag(k), ranks in ascending k.  Returns the count.
template <typename F>
__device__ __forceinline__ int half_compact(int n, F flag, int* out, RsShared& sh, int cam, int t, int wq, int lane) {
    int run = 0;
    for (int b0 = 0; b0 < n; b0 += RS_HALF) {
        const int k = b0 + t;
        const bool v = k < n && flag(k);
        const unsigned b = __ballot_sync(0xffffffffu, v);
        bar_half(cam);
        if (lane == 0) sh.wcnt[wq] = __popc(b);
        bar_half(cam);
        int before = 0, all = 0;
#pragma unroll
        for (int w = 0; w < RS_WARPS; ++w) {
            const int c = sh.wcnt[w];
            before += w < wq ? c : 0;
            all += c;
        }
        if (v) out[run + before + __popc(b & ((1u << lane) - 1u))] = k;
        run += all;
    }
    return run;
}

__device__ __forceinline__ void constraint_row(const float4 u, double sf, double& c0, double& c1, double& c2, double& dist) {
    const double x1 = __dmul_rn((double)u.x, sf), y1 = __dmul_rn((double)u.y, sf);
    const double x2 = __dmul_rn((double)u.z, sf), y2 = __dmul_rn((double)u.w, sf);
    const double dx = __dsub_rn(x1, x2), dy = __dsub_rn(y1, y2);
    dist = __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
    c0 = dy;
    c1 = -dx;
    c2 = __dsub_rn(__dmul_rn(x1, y2), __dmul_rn(y1, x2));
}

// [p q] s = -r for the two drawn rows, closed-form inverse (oracle/ransac.py::solve2)
__device__ __forceinline__ void solve2(double p0, double p1, double q0, double q1, double r0, double r1, double& s0, double& s1) {
    const double det = __dsub_rn(__dmul_rn(p0, q1), __dmul_rn(q0, p1));
    const double a = -r0, b = -r1;
    s0 = __ddiv_rn(__dsub_rn(__dmul_rn(q1, a), __dmul_rn(q0, b)), det);
    s1 = __ddiv_rn(__dsub_rn(__dmul_rn(p0, b), __dmul_rn(p1, a)), det);
}

// One camera's RANSAC on 256 threads (`bar` = the named barrier of this thread group).  fetch(k, prev, cur) yields
// point k of the list; the 0/1 inlier mask is left in bits[0..n).
struct RansacParams {
    double thr_px;                 // cfg.ransac_threshold
    int iters, seed, frame_index, cam_key;
};

// und_of(k): point k undistorted, f32: (previous position rotated by R_p_c, current position).
template <typename UndOf>
__device__ __forceinline__ void ransac_camera(const CamModel& cm, int n, UndOf und_of, float4* und, int* raw_idx,
                                              uint8_t* bits, RsShared& sh, int bar, int t, const RansacParams& prm) {
    const int lane = t & 31, wq = t >> 5;
    if (n <= 0) return;
    // 2. undistorted points (f64 inside, f32 results), partial sums of the point norms
    double part = 0.0;
    for (int k = t; k < n; k += RS_HALF) {
        const float4 u = und_of(k);
        und[k] = u;
        const double n1 = __dsqrt_rn(__dadd_rn(__dmul_rn((double)u.x, (double)u.x), __dmul_rn((double)u.y, (double)u.y)));
        const double n2 = __dsqrt_rn(__dadd_rn(__dmul_rn((double)u.z, (double)u.z), __dmul_rn((double)u.w, (double)u.w)));
        part = __dadd_rn(part, __dadd_rn(n1, n2));
    }
    const double norm_sum = half_tree_sum(part, sh, bar, wq, lane);
    const double sf = __dmul_rn(__ddiv_rn(__dmul_rn(2.0, (double)n), norm_sum), 1.4142135623730951);
    const double unit = __dmul_rn(__ddiv_rn(2.0, __dadd_rn(cm.fx, cm.fy)), sf);
    const double reject = __dmul_rn(50.0, unit);
    const double thr = __dmul_rn(prm.thr_px, unit);

    // 3. pre-rejection, mean displacement
    part = 0.0;
    int my_raw = 0;
    for (int k = t; k < n; k += RS_HALF) {
        double c0, c1, c2, dist;
        constraint_row(und[k], sf, c0, c1, c2, dist);
        const bool raw = !(dist > reject);
        bits[k] = raw ? 0x80 : 0;
        my_raw += raw;
        part = __dadd_rn(part, raw ? dist : 0.0);
    }
    const double dist_sum = half_tree_sum(part, sh, bar, wq, lane);
    my_raw = __reduce_add_sync(0xffffffffu, my_raw);
    bar_half(bar);
    if (lane == 0) sh.wcnt[wq] = my_raw;
    bar_half(bar);
    int n_raw = 0;
#pragma unroll
    for (int w = 0; w < RS_WARPS; ++w) n_raw += sh.wcnt[w];
    bar_half(bar);

    if (n_raw < 3) {
        for (int k = t; k < n; k += RS_HALF) bits[k] = 0;
        return;
    }
    if (__ddiv_rn(dist_sum, (double)n_raw) < unit) {                // (almost) pure rotation
        for (int k = t; k < n; k += RS_HALF) {
            double c0, c1, c2, dist;
            constraint_row(und[k], sf, c0, c1, c2, dist);
            bits[k] = ((bits[k] & 0x80) && !(dist > thr)) ? 1 : 0;
        }
        return;
    }
    // 4. hypotheses: thread h draws its pair and solves its model
    half_compact(n, [&](int k) { return (bits[k] & 0x80) != 0; }, raw_idx, sh, bar, t, wq, lane);
    bar_half(bar);
    const int iters = prm.iters;
    if (t < iters) {
        const unsigned long long key = ((unsigned long long)(prm.seed & 0xFFFFFF) << 40) |
                                       ((unsigned long long)(prm.frame_index & 0xFFFFFFF) << 12) |
                                       ((unsigned long long)(prm.cam_key & 0xF) << 8) | (unsigned long long)(t & 0xFF);
        const unsigned long long r = splitmix64(key);
        const unsigned long long nr = (unsigned long long)n_raw;
        const int i1 = (int)(((r >> 32) * nr) >> 32);
        const int diff = 1 + (int)(((r & 0xFFFFFFFFull) * (nr - 1)) >> 32);
        const int i2 = i1 + diff < n_raw ? i1 + diff : i1 + diff - n_raw;
        double a0, b0, c0, a1, b1, c1, dist;
        constraint_row(und[raw_idx[i1]], sf, a0, b0, c0, dist);
        constraint_row(und[raw_idx[i2]], sf, a1, b1, c1, dist);
        const double la = __dadd_rn(fabs(a0), fabs(a1)), lb = __dadd_rn(fabs(b0), fabs(b1)), lc = __dadd_rn(fabs(c0), fabs(c1));
        const int fixed = (la <= lb && la <= lc) ? 0 : (lb <= lc ? 1 : 2);        // first minimum
        double s0, s1;
        if (fixed == 0) {
            solve2(b0, b1, c0, c1, a0, a1, s0, s1);
            sh.model[t][0] = 1.0; sh.model[t][1] = s0; sh.model[t][2] = s1;
        } else if (fixed == 1) {
            solve2(a0, a1, c0, c1, b0, b1, s0, s1);
            sh.model[t][0] = s0; sh.model[t][1] = 1.0; sh.model[t][2] = s1;
        } else {
            solve2(a0, a1, b0, b1, c0, c1, s0, s1);
            sh.model[t][0] = s0; sh.model[t][1] = s1; sh.model[t][2] = 1.0;
        }
    }
    bar_half(bar);
    // 5. score every point against every model: the (hypotheses x points) matrix in one pass
    int cnt[RS_MAX_HYP];
#pragma unroll
    for (int h = 0; h < RS_MAX_HYP; ++h) cnt[h] = 0;
    for (int k = t; k < n; k += RS_HALF) {
        double c0, c1, c2, dist;
        constraint_row(und[k], sf, c0, c1, c2, dist);
        const unsigned raw = bits[k] & 0x80;
        unsigned m = raw;
#pragma unroll
        for (int h = 0; h < RS_MAX_HYP - 1; ++h) {
            const double err = __dadd_rn(__dadd_rn(__dmul_rn(c0, sh.model[h][0]), __dmul_rn(c1, sh.model[h][1])),
                                         __dmul_rn(c2, sh.model[h][2]));
            const bool inl = raw && h < iters && fabs(err) < thr;
            cnt[h] += inl;
            m |= inl ? (1u << h) : 0u;
        }
        bits[k] = (uint8_t)m;
    }
#pragma unroll
    for (int h = 0; h < RS_MAX_HYP - 1; ++h) {
        const int c = __reduce_add_sync(0xffffffffu, cnt[h]);
        if (lane == 0) sh.hcnt[wq][h] = c;
    }
    bar_half(bar);
    int best = -1, best_cnt = 0;
    const double min_cnt = __dmul_rn(0.2, (double)n);
    for (int h = 0; h < iters; ++h) {
        int c = 0;
#pragma unroll
        for (int w = 0; w < RS_WARPS; ++w) c += sh.hcnt[w][h];
        if ((double)c < min_cnt) continue;
        if (c > best_cnt) {                                         // the first hypothesis with the largest set
            best_cnt = c;
            best = h;
        }
    }
    for (int k = t; k < n; k += RS_HALF) bits[k] = (best >= 0 && ((bits[k] >> best) & 1)) ? 1 : 0;
}

__global__ void __launch_bounds__(2 * RS_HALF) k_ransac(const __grid_constant__ Geom g, const __grid_constant__ DevState d,
                                                        RansacScratch sc0, RansacScratch sc1, int parity) {
    __shared__ RsShared shm[2];
    __shared__ int n_pts[2];
    pdl_wait();
    pdl_launch_dependents();
    const int s = blockIdx.x, tid = threadIdx.x, cam = tid >> 8, t = tid & (RS_HALF - 1), lane = tid & 31, wq = t >> 5;
    RsShared& sh = shm[cam];
    const size_t base = (size_t)s * g.NMAX;
    const RansacScratch sc = cam ? sc1 : sc0;
    int* idx = sc.idx + base;

    // 1. the point list: stereo-matched survivors in table order
    const int* tc = d.t_cell + base;
    const int n = half_compact(g.NMAX, [&](int k) { return tc[k] >= 0; }, idx, sh, cam, t, wq, lane);
    if (t == 0) n_pts[cam] = n;
    bar_half(cam);

    RansacParams prm;
    prm.thr_px = g.ransac_thr;
    prm.iters = g.ransac_iters;
    prm.seed = g.ransac_seed;
    prm.frame_index = d.frame_index[s];
    prm.cam_key = cam;
    // the undistortions themselves were made by k_track at the end of every feature's LK chain (r_prev: previous
    // positions through R_p_c; t_und: the current ones, the values the publisher uses), rounded to f32 here as
    // cv2.undistortPoints rounds them for float32 input
    const float4* rp = d.r_prev + base;
    const double4* tu = d.t_und + base;
    ransac_camera(cam ? g.cam1 : g.cam0, n,
                  [&](int k) {
                      const int wi = idx[k];
                      const float4 pr = rp[wi];
                      const double4 cu = tu[wi];
                      return cam ? make_float4(pr.z, pr.w, (float)cu.z, (float)cu.w) : make_float4(pr.x, pr.y, (float)cu.x, (float)cu.y);
                  },
                  sc.und + base, sc.raw_idx + base, sc.bits + base, sh, cam, t, prm);
    __syncthreads();

    // both cameras must agree (feature_tracker.py:141); losers leave the table before the grid is rebuilt
    const int n_all = n_pts[0];
    int kept = 0;
    for (int k = tid; k < n_all; k += 2 * RS_HALF) {
        const bool ok = sc0.bits[base + k] && sc1.bits[base + k];
        if (!ok) d.t_cell[base + sc0.idx[base + k]] = -1;
        kept += ok;
    }
    kept = __reduce_add_sync(0xffffffffu, kept);
    if (lane == 0 && kept) atomicAdd(&d.counters[s * 8 + 5], kept);
}

// Flat point lists of one camera (per-stage entry point avb_two_point_ransac).
__global__ void __launch_bounds__(RS_HALF) k_ransac_points(CamModel cm, const double* R, const float2* prev, const float2* cur, int n,
                                                           float4* und, int* raw_idx, uint8_t* bits, RansacParams prm) {
    __shared__ RsShared sh;
    ransac_camera(cm, n,
                  [&](int k) {
                      const float2 a = prev[k], b = cur[k];
                      double ax, ay, bx, by;
                      undistort_pt(cm, (double)a.x, (double)a.y, R, ax, ay);
                      undistort_pt(cm, (double)b.x, (double)b.y, nullptr, bx, by);
                      return make_float4((float)ax, (float)ay, (float)bx, (float)by);
                  },
                  und, raw_idx, bits, sh, 0, threadIdx.x, prm);
}

void launch_ransac(const Geom& g, const DevState& d, int parity, cudaStream_t st) {
    RansacScratch a = {d.r_idx, d.r_und, d.r_raw, d.r_bits};
    const size_t off = (size_t)g.S * g.NMAX;
    RansacScratch b = {d.r_idx + off, d.r_und + off, d.r_raw + off, d.r_bits + off};
    launch_k(k_ransac, dim3(g.S), dim3(2 * RS_HALF), 0, st, g_avb_pdl != 0, g, d, a, b, parity);
}

void launch_ransac_points(const Geom& g, const CamModel& cm, const double* R, const float2* prev, const float2* cur, int n,
                          float4* und, int* raw_idx, uint8_t* bits, int frame_index, int cam_key, int seed, double thr_px,
                          cudaStream_t st) {
    RansacParams prm;
    prm.thr_px = thr_px;
    prm.iters = g.ransac_iters;
    prm.seed = seed;
    prm.frame_index = frame_index;
    prm.cam_key = cam_key;
    k_ransac_points<<<1, RS_HALF, 0, st>>>(cm, R, prev, cur, n, und, raw_idx, bits, prm);
}
