// Grid bookkeeping kernels: the reference's per-cell Python list logic (Appendix B of SURVEY.md) on device.
//   k_select       FeatureAdder mask + sieve + per-cell cap (feature_adder.py:59-77); on frame 0 the per-cell
//                  top-`min` ranking of FeatureInitializer (feature_initializer.py:79-80)
//   k_finish       tracker re-bin (feature_tracker.py:139-155) + adder top-`min` insert (feature_adder.py:102-108)
//                  + pruner (feature_pruner.py:13-19) + id assignment in cell-major order (B10)
//                  + FeaturePublisher.publish (feature_publisher.py:90-121), one CTA per stream
// Every ranking is max-by-key: key = (response << 24 | inverted scan index) reproduces Python's stable
// sorted(..., reverse=True) over detections in scan order (B9); lifetime ranking is stable by list position.
#include "avb_lk.cuh"

__global__ void k_clear_frame(Geom g, DevState d) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < g.S * g.NC) d.kp_count[i] = 0;
    if (i < g.S * 8) d.counters[i] = 0;
}

void launch_clear_frame(const Geom& g, const DevState& d, cudaStream_t st) {
    const int n = g.S * (g.NC > 8 ? g.NC : 8);
    k_clear_frame<<<(n + 255) / 256, 256, 0, st>>>(g, d);
}

__device__ __forceinline__ unsigned long long warp_max_u64(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long t = __shfl_xor_sync(0xffffffffu, v, o);
        v = t > v ? t : v;
    }
    return v;
}

__global__ void k_finish(const __grid_constant__ Geom g, const __grid_constant__ DevState d, int parity, int first_frame);

// One CTA per grid cell.  Dynamic smem: keys[KPC] (u32).
//   1. live features whose 7x7 mask patch can reach this cell are collected (a handful), the cell's FAST keys are
//      masked against them (frame 0: keys of stereo outliers are dropped instead)
//   2. top-k by key: every warp extracts the k largest keys of its strided share with warp-max rounds (no block
//      barrier), then warp 0 extracts the k largest of the 8k finalists.  A key packs (response, inverted scan
//      index), the low word of a finalist is its bucket index, so ties cannot occur.
#define SEL_WARPS 8
#define SEL_MAXF 128               // live features that can touch one cell: gmax per cell x the 9 cells around, generously
__global__ void __launch_bounds__(32 * SEL_WARPS) k_select(const __grid_constant__ Geom g, const __grid_constant__ DevState d,
                                                          int parity, int first_frame) {
    extern __shared__ unsigned smem_u32[];
    unsigned* keys = smem_u32;
    __shared__ unsigned feat[SEL_MAXF];
    __shared__ int n_feat;
    __shared__ unsigned long long fin[SEL_WARPS * AVB_MAX_CAP];

    pdl_wait();
    pdl_launch_dependents();
    const int cell = blockIdx.x, s = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const size_t kbase = ((size_t)s * g.NC + cell) * g.KPC;
    const int n = min(d.kp_count[s * g.NC + cell], g.KPC);
    const int k = first_frame ? g.gmin : g.gmax;

    if (tid == 0) n_feat = 0;
    __syncthreads();
    if (!first_frame) {
        const int cx0 = (cell % g.cols) * g.gw, cy0 = (cell / g.cols) * g.gh;
        for (int i = tid; i < g.NMAX; i += 32 * SEL_WARPS) {
            if (d.t_cell[(size_t)s * g.NMAX + i] >= 0) {
                const float2 p = d.t_p0[(size_t)s * g.NMAX + i];
                const int fx = (int)p.x, fy = (int)p.y;             // int() truncation (B7)
                // mask[y-3:y+4, x-3:x+4] = 0 with a negative slice start selects nothing (B7)
                if (fx >= 3 && fy >= 3 && fx >= cx0 - 3 && fx < cx0 + g.gw + 3 && fy >= cy0 - 3 && fy < cy0 + g.gh + 3) {
                    const int pos = atomicAdd(&n_feat, 1);
                    if (pos < SEL_MAXF) feat[pos] = ((unsigned)fy << 16) | (unsigned)fx;
                }
            }
        }
    }
    __syncthreads();
    const int nf_all = n_feat, nf = min(nf_all, SEL_MAXF);
    for (int i = tid; i < n; i += 32 * SEL_WARPS) {
        unsigned key = d.kp_key[kbase + i];
        if (first_frame) {
            if (!d.kp_ok[kbase + i]) key = 0;
        } else {
            int resp, x, y;
            kp_decode(key, g.W, resp, x, y);
            bool masked = false;
            for (int j = 0; j < nf; ++j) {
                const int fx = (int)(feat[j] & 0xffffu), fy = (int)(feat[j] >> 16);
                masked |= (abs(x - fx) <= 3) && (abs(y - fy) <= 3);
            }
            if (nf_all > SEL_MAXF) {                                // overflow of the local list: fall back to the full table
                for (int j = 0; j < g.NMAX && !masked; ++j) {
                    if (d.t_cell[(size_t)s * g.NMAX + j] < 0) continue;
                    const float2 p = d.t_p0[(size_t)s * g.NMAX + j];
                    const int fx = (int)p.x, fy = (int)p.y;
                    masked = fx >= 3 && fy >= 3 && (abs(x - fx) <= 3) && (abs(y - fy) <= 3);
                }
            }
            if (masked) key = 0;
        }
        keys[i] = key;
    }
    __syncthreads();

    // per-warp top-k of the keys i = warp*32 + lane (mod 256); an exhausted warp reports zeros
    {
        unsigned long long last = ~0ull;
        for (int r = 0; r < k; ++r) {
            unsigned long long best = 0;
            for (int i = warp * 32 + lane; i < n; i += 32 * SEL_WARPS) {
                const unsigned long long v = ((unsigned long long)keys[i] << 32) | (unsigned)i;
                if (v < last && v > best) best = v;
            }
            best = warp_max_u64(best);
            if ((best >> 32) == 0) best = 0;
            if (lane == 0) fin[warp * AVB_MAX_CAP + r] = best;
            last = best;
        }
    }
    __syncthreads();
    if (warp == 0) {
        unsigned long long last = ~0ull;
        int found = 0;
        for (int r = 0; r < k; ++r) {
            unsigned long long best = 0;
            for (int i = lane; i < SEL_WARPS * k; i += 32) {
                const unsigned long long v = fin[(i / k) * AVB_MAX_CAP + (i % k)];
                if (v < last && v > best) best = v;
            }
            best = warp_max_u64(best);
            if ((best >> 32) == 0) break;
            last = best;
            if (lane == 0) {
                const size_t o = (size_t)s * g.NMAX + cell * g.gmax + r;
                const int src = (int)(best & 0xffffffffu);
                d.c_key[o] = (unsigned)(best >> 32);
                d.c_src[o] = src;
                if (first_frame) {
                    d.c_p1[o] = d.kp_p1[kbase + src];
                    d.c_ok[o] = 1;
                }
            }
            ++found;
        }
        if (lane == 0) {
            d.c_count[s * g.NC + cell] = found;
            atomicAdd(&d.counters[s * 8 + 3], n);
            atomicAdd(&d.counters[s * 8 + 4], found);
        }
    }
}

int avb_set_smem_limits(size_t select_bytes, size_t grid_bytes) {
    cudaError_t e = cudaSuccess;
    if (select_bytes > 48 * 1024)
        e = cudaFuncSetAttribute(k_select, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)select_bytes);
    if (e == cudaSuccess && grid_bytes > 48 * 1024)
        e = cudaFuncSetAttribute(k_finish, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)grid_bytes);
    return e == cudaSuccess ? 0 : -1;
}

void launch_select(const Geom& g, const DevState& d, int parity, int first_frame, cudaStream_t st) {
    const size_t smem = (size_t)g.KPC * sizeof(unsigned);
    launch_k(k_select, dim3(g.NC, g.S), dim3(32 * SEL_WARPS), smem, st, g_avb_pdl && !first_frame, g, d, parity, first_frame);
}

// k_finish: one 1024-thread CTA per stream closes the frame.
//   phase 1  stable counting sort of the tracked survivors by their new cell, in the previous frame's flattened grid
//            order (feature_tracker.py:139-155, B16)
//   phase 2  one warp per cell: append the first `gmin` stereo inliers of the cell's candidate list
//            (feature_adder.py:102-108, B8), prune to `gmax` by lifetime with a stable rank (feature_pruner.py:13-19, B9)
//   phase 3  exclusive scans over the cells -> ids in cell-major order (B10); one thread per (feature, camera) undistorts
//            to normalized coordinates with that camera's model and fills the packed result block
//            (feature_publisher.py:90-121; dtype quirk B11)
// Dynamic smem: cellof[NMAX] | seg[NMAX] | slife[NMAX] (int).
#define FIN_THREADS 1024
__global__ void __launch_bounds__(FIN_THREADS) k_finish(const __grid_constant__ Geom g, const __grid_constant__ DevState d, int parity,
                                                        int first_frame) {
    extern __shared__ int smem_i32[];
    int* cellof = smem_i32;
    int* seg = smem_i32 + g.NMAX;
    int* slife = smem_i32 + 2 * g.NMAX;
    __shared__ int cnt[AVB_MAX_CELLS + 1];      // survivors per cell, then the cell's final feature count
    __shared__ int off[AVB_MAX_CELLS + 1];      // segment start of a cell in seg[]
    __shared__ int off_cnt[AVB_MAX_CELLS + 1];  // exclusive scan of the final counts
    __shared__ int off_new[AVB_MAX_CELLS + 1];  // exclusive scan of the new-feature counts
    __shared__ int nnew[AVB_MAX_CELLS];

    pdl_wait();
    pdl_launch_dependents();
    const int s = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned lt = (1u << lane) - 1u;
    const GridTable prev = d.grid[parity ^ 1], cur = d.grid[parity];
    const size_t base = (size_t)s * g.NMAX;

    for (int c = tid; c <= g.NC; c += FIN_THREADS) cnt[c] = 0;
    __syncthreads();
    for (int i = tid; i < g.NMAX; i += FIN_THREADS) {
        const int c = first_frame ? -1 : d.t_cell[base + i];
        cellof[i] = c;
        if (c >= 0) atomicAdd(&cnt[c], 1);
    }
    __syncthreads();
    if (warp == 0) {                    // exclusive scan of the survivor counts
        int run = 0;
        for (int b0 = 0; b0 < g.NC; b0 += 32) {
            const int c = b0 + lane;
            const int v = c < g.NC ? cnt[c] : 0;
            int inc = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += t;
            }
            if (c < g.NC) off[c] = run + inc - v;
            run += __shfl_sync(0xffffffffu, inc, 31);
        }
    }
    __syncthreads();

    for (int c = warp; c < g.NC; c += FIN_THREADS / 32) {
        const int nt = cnt[c], o0 = off[c];
        if (nt) {                       // stable placement: survivors of this cell in table order
            int placed = 0;
            for (int b0 = 0; b0 < g.NMAX && placed < nt; b0 += 32) {
                const int wi = b0 + lane;
                const bool v = wi < g.NMAX && cellof[wi] == c;
                const unsigned b = __ballot_sync(0xffffffffu, v);
                if (v) {
                    const int pos = o0 + placed + __popc(b & lt);
                    seg[pos] = wi;
                    slife[pos] = prev.life[base + wi];
                }
                placed += __popc(b);
            }
        }
        // new features: the first `gmin` stereo inliers of the candidate list (already in descending key order)
        const int nc = d.c_count[s * g.NC + c];
        const bool okj = lane < nc && d.c_ok[base + c * g.gmax + lane];
        const unsigned bn = __ballot_sync(0xffffffffu, okj);
        const int nrank = __popc(bn & lt);
        const bool is_new = okj && nrank < g.gmin;
        const int nn = min(__popc(bn), g.gmin);
        const int total = nt + nn;
        const bool prune = total > g.gmax;
        __syncwarp();

        const size_t obase = base + (size_t)c * g.gmax;
        for (int i = lane; i < nt; i += 32) {
            int pos = i;
            const int li = slife[o0 + i];
            if (prune) {                // stable sort by lifetime, descending (feature_pruner.py:18)
                int rank = 0;
                for (int j = 0; j < nt; ++j) {
                    const int lj = slife[o0 + j];
                    rank += (lj > li) || (lj == li && j < i);
                }
                pos = rank;
            }
            if (pos < g.gmax) {
                const int wi = seg[o0 + i];
                cur.ids[obase + pos] = prev.ids[base + wi];
                cur.life[obase + pos] = li + 1;
                cur.p0[obase + pos] = d.t_p0[base + wi];
                cur.p1[obase + pos] = d.t_p1[base + wi];
                cur.fresh[obase + pos] = 0;
            }
        }
        if (is_new) {                   // lifetime 1 ranks after every tracked feature (lifetime >= 2)
            const int pos = nt + nrank;
            if (pos < g.gmax) {
                int resp, x, y;
                kp_decode(d.c_key[base + c * g.gmax + lane], g.W, resp, x, y);
                cur.ids[obase + pos] = -1;
                cur.life[obase + pos] = 1;
                cur.p0[obase + pos] = make_float2((float)x, (float)y);
                cur.p1[obase + pos] = d.c_p1[base + c * g.gmax + lane];
                cur.fresh[obase + pos] = 1;
                d.new_rank[obase + pos] = (uint8_t)nrank;
            }
        }
        __syncwarp();
        if (lane == 0) {
            const int fc = min(total, g.gmax);
            cur.count[s * g.NC + c] = fc;
            cnt[c] = fc;                // from here on: the cell's final count
            nnew[c] = nn;
        }
    }
    __syncthreads();

    if (warp == 0) {                    // exclusive scans over the cells
        int run_c = 0, run_n = 0;
        for (int b0 = 0; b0 < g.NC; b0 += 32) {
            const int c = b0 + lane;
            const int vc = c < g.NC ? cnt[c] : 0;
            const int vn = c < g.NC ? nnew[c] : 0;
            int ic = vc, in = vn;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int tc = __shfl_up_sync(0xffffffffu, ic, o), tn = __shfl_up_sync(0xffffffffu, in, o);
                if (lane >= o) {
                    ic += tc;
                    in += tn;
                }
            }
            if (c < g.NC) {
                off_cnt[c] = run_c + ic - vc;
                off_new[c] = run_n + in - vn;
            }
            run_c += __shfl_sync(0xffffffffu, ic, 31);
            run_n += __shfl_sync(0xffffffffu, in, 31);
        }
        if (lane == 0) {
            off_cnt[g.NC] = run_c;
            off_new[g.NC] = run_n;
        }
    }
    __syncthreads();
    const int total = off_cnt[g.NC], total_new = off_new[g.NC];

    int any_fresh = 0;
    for (int i = tid; i < g.NMAX; i += FIN_THREADS) {
        const int c = i / g.gmax, j = i - c * g.gmax;
        if (j < cnt[c]) any_fresh |= cur.fresh[base + i];
    }
    const int has_new = __syncthreads_or(any_fresh);

    uint8_t* ob = d.out + (size_t)s * out_stride_bytes(g.NMAX);
    const long long next_id = d.next_id[s];
    for (int w = tid; w < 2 * g.NMAX; w += FIN_THREADS) {
        const int i = w >> 1, cam = w & 1;
        const int c = i / g.gmax, j = i - c * g.gmax;
        if (j >= cnt[c]) continue;
        const int pos = off_cnt[c] + j;
        double* m = out_meas(ob, g.NMAX) + 4 * pos;
        if (cam == 0) {
            long long id = cur.ids[base + i];
            if (cur.fresh[base + i]) {
                id = next_id + off_new[c] + d.new_rank[base + i];
                cur.ids[base + i] = id;
            }
            const float2 p0 = cur.p0[base + i];
            double u0, v0;
            undistort_pt(g.cam0, (double)p0.x, (double)p0.y, nullptr, u0, v0);
            if (!has_new) {             // all-float32 point list -> cv2 returns float32 (B11)
                u0 = (double)(float)u0;
                v0 = (double)(float)v0;
            }
            out_ids(ob)[pos] = id;
            m[0] = u0;
            m[1] = v0;
            out_cell(ob, g.NMAX)[pos] = c;
            out_life(ob, g.NMAX)[pos] = cur.life[base + i];
            out_p0(ob, g.NMAX)[2 * pos] = p0.x;
            out_p0(ob, g.NMAX)[2 * pos + 1] = p0.y;
        } else {
            const float2 p1 = cur.p1[base + i];
            double u1, v1;
            undistort_pt(g.cam1, (double)p1.x, (double)p1.y, nullptr, u1, v1);
            m[2] = (double)(float)u1;
            m[3] = (double)(float)v1;
            out_p1(ob, g.NMAX)[2 * pos] = p1.x;
            out_p1(ob, g.NMAX)[2 * pos + 1] = p1.y;
        }
    }
    if (tid == 0) {
        avb_frame_header* h = reinterpret_cast<avb_frame_header*>(ob);
        const int* cn = d.counters + s * 8;
        h->n_features = total;
        h->next_feature_id = next_id + total_new;
        h->before_tracking = cn[0];
        h->after_tracking = cn[1];
        h->after_matching = cn[2];
        h->after_ransac = g.ransac ? cn[5] : cn[2];    // reference: all-ones stub (B2) -> same as after_matching
        h->has_new = has_new;
        h->n_fast = cn[3];
        h->n_candidates = cn[4];
        h->frame_index = d.frame_index[s];
        d.next_id[s] = next_id + total_new;
        d.frame_index[s] += 1;
    }
}

void launch_finish(const Geom& g, const DevState& d, int parity, int first_frame, cudaStream_t st) {
    launch_k(k_finish, dim3(g.S), dim3(FIN_THREADS), (size_t)g.NMAX * 3 * sizeof(int), st, g_avb_pdl != 0, g, d, parity, first_frame);
}
