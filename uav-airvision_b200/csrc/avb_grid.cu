// Grid bookkeeping kernels: the reference's per-cell Python list logic (Appendix B of SURVEY.md) on device.
//   k_select       FeatureAdder mask + sieve + per-cell cap (feature_adder.py:59-77); on frame 0 the per-cell
//                  top-`min` ranking of FeatureInitializer (feature_initializer.py:79-80)
//   k_grid_update  tracker re-bin (feature_tracker.py:139-155) + adder top-`min` insert (feature_adder.py:102-108)
//                  + pruner (feature_pruner.py:13-19), one warp per cell
//   k_publish      id assignment in cell-major order (B10) + FeaturePublisher.publish (feature_publisher.py:90-121)
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

__global__ void k_grid_update(Geom g, DevState d, int parity, int first_frame);

// One CTA per grid cell.  Dynamic smem: keys[KPC] (u32) | feat[NMAX] (packed int16 x,y of tracked survivors).
__global__ void __launch_bounds__(256) k_select(Geom g, DevState d, int parity, int first_frame) {
    extern __shared__ unsigned smem_u32[];
    unsigned* keys = smem_u32;
    unsigned* feat = smem_u32 + g.KPC;
    __shared__ int n_feat;
    __shared__ unsigned long long red[8];

    const int cell = blockIdx.x, s = blockIdx.y, tid = threadIdx.x;
    const size_t kbase = ((size_t)s * g.NC + cell) * g.KPC;
    const int n = min(d.kp_count[s * g.NC + cell], g.KPC);

    if (tid == 0) n_feat = 0;
    __syncthreads();
    if (!first_frame) {
        // live features of the WHOLE frame (a 7x7 mask patch may reach across a cell border)
        for (int i = tid; i < g.NMAX; i += 256) {
            if (d.t_cell[(size_t)s * g.NMAX + i] >= 0) {
                const float2 p = d.t_p0[(size_t)s * g.NMAX + i];
                const int fx = (int)p.x, fy = (int)p.y;             // int() truncation (B7)
                // mask[y-3:y+4, x-3:x+4] = 0 with a negative slice start selects nothing (B7)
                if (fx >= 3 && fy >= 3) feat[atomicAdd(&n_feat, 1)] = ((unsigned)fy << 16) | (unsigned)fx;
            }
        }
    }
    __syncthreads();
    const int nf = n_feat;
    for (int i = tid; i < n; i += 256) {
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
            if (masked) key = 0;
        }
        keys[i] = key;
    }
    __syncthreads();

    const int k = first_frame ? g.gmin : g.gmax;
    unsigned long long last = ~0ull;
    int found = 0;
    for (int r = 0; r < k; ++r) {
        unsigned long long best = 0;
        for (int i = tid; i < n; i += 256) {
            const unsigned long long v = ((unsigned long long)keys[i] << 32) | (unsigned)i;
            if (v < last && v > best) best = v;
        }
        best = warp_max_u64(best);
        if ((tid & 31) == 0) red[tid >> 5] = best;
        __syncthreads();
        best = red[0];
#pragma unroll
        for (int w = 1; w < 8; ++w) best = red[w] > best ? red[w] : best;
        __syncthreads();
        if ((best >> 32) == 0) break;
        last = best;
        if (tid == 0) {
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
    if (tid == 0) {
        d.c_count[s * g.NC + cell] = found;
        atomicAdd(&d.counters[s * 8 + 3], n);
        atomicAdd(&d.counters[s * 8 + 4], found);
    }
}

int avb_set_smem_limits(size_t select_bytes, size_t grid_bytes) {
    cudaError_t e = cudaSuccess;
    if (select_bytes > 48 * 1024)
        e = cudaFuncSetAttribute(k_select, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)select_bytes);
    if (e == cudaSuccess && grid_bytes > 48 * 1024)
        e = cudaFuncSetAttribute(k_grid_update, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)grid_bytes);
    return e == cudaSuccess ? 0 : -1;
}

void launch_select(const Geom& g, const DevState& d, int parity, int first_frame, cudaStream_t st) {
    const size_t smem = ((size_t)g.KPC + g.NMAX) * sizeof(unsigned);
    k_select<<<dim3(g.NC, g.S), 256, smem, st>>>(g, d, parity, first_frame);
}

// One warp per grid cell.  Dynamic smem: list[NMAX] (int) | llife[NMAX] (int).
__global__ void __launch_bounds__(32) k_grid_update(Geom g, DevState d, int parity, int first_frame) {
    extern __shared__ int smem_i32[];
    int* list = smem_i32;
    int* llife = smem_i32 + g.NMAX;

    const int c = blockIdx.x, s = blockIdx.y, lane = threadIdx.x;
    const unsigned lt = (1u << lane) - 1u;
    const GridTable prev = d.grid[parity ^ 1], cur = d.grid[parity];
    const size_t base = (size_t)s * g.NMAX;

    // tracked survivors that landed in this cell, in the previous frame's flattened grid order (B16)
    int nt = 0;
    if (!first_frame) {
        for (int b0 = 0; b0 < g.NMAX; b0 += 32) {
            const int wi = b0 + lane;
            const bool v = wi < g.NMAX && d.t_cell[base + wi] == c;
            const unsigned b = __ballot_sync(0xffffffffu, v);
            if (v) {
                const int pos = nt + __popc(b & lt);
                list[pos] = wi;
                llife[pos] = prev.life[base + wi];
            }
            nt += __popc(b);
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
        if (prune) {                    // stable sort by lifetime, descending (feature_pruner.py:18)
            const int li = llife[i];
            int rank = 0;
            for (int j = 0; j < nt; ++j) rank += (llife[j] > li) || (llife[j] == li && j < i);
            pos = rank;
        }
        if (pos < g.gmax) {
            const int wi = list[i];
            cur.ids[obase + pos] = prev.ids[base + wi];
            cur.life[obase + pos] = llife[i] + 1;
            cur.p0[obase + pos] = d.t_p0[base + wi];
            cur.p1[obase + pos] = d.t_p1[base + wi];
            cur.fresh[obase + pos] = 0;
        }
    }
    if (is_new) {                       // lifetime 1 ranks after every tracked feature (lifetime >= 2)
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
    if (lane == 0) {
        cur.count[s * g.NC + c] = min(total, g.gmax);
        d.n_new[s * g.NC + c] = nn;
    }
}

void launch_grid_update(const Geom& g, const DevState& d, int parity, int first_frame, cudaStream_t st) {
    k_grid_update<<<dim3(g.NC, g.S), 32, (size_t)g.NMAX * 2 * sizeof(int), st>>>(g, d, parity, first_frame);
}

// One CTA per stream: ids for the new features, undistortion to normalized coordinates, packed result block.
__global__ void __launch_bounds__(256) k_publish(Geom g, DevState d, int parity) {
    __shared__ int off_cnt[AVB_MAX_CELLS + 1];
    __shared__ int off_new[AVB_MAX_CELLS + 1];
    const int s = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
    const GridTable cur = d.grid[parity];
    const size_t base = (size_t)s * g.NMAX;

    if (tid < 32) {                     // exclusive scans over the cells
        int run_c = 0, run_n = 0;
        for (int b0 = 0; b0 < g.NC; b0 += 32) {
            const int c = b0 + lane;
            const int vc = c < g.NC ? cur.count[s * g.NC + c] : 0;
            const int vn = c < g.NC ? d.n_new[s * g.NC + c] : 0;
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
    for (int i = tid; i < g.NMAX; i += 256) {
        const int c = i / g.gmax, j = i - c * g.gmax;
        if (j < cur.count[s * g.NC + c]) any_fresh |= cur.fresh[base + i];
    }
    const int has_new = __syncthreads_or(any_fresh);

    uint8_t* ob = d.out + (size_t)s * out_stride_bytes(g.NMAX);
    const long long next_id = d.next_id[s];
    for (int i = tid; i < g.NMAX; i += 256) {
        const int c = i / g.gmax, j = i - c * g.gmax;
        if (j >= cur.count[s * g.NC + c]) continue;
        const int pos = off_cnt[c] + j;
        long long id = cur.ids[base + i];
        if (cur.fresh[base + i]) {
            id = next_id + off_new[c] + d.new_rank[base + i];
            cur.ids[base + i] = id;
        }
        const float2 p0 = cur.p0[base + i], p1 = cur.p1[base + i];
        double u0, v0, u1, v1;
        undistort_pt(g.cam0, (double)p0.x, (double)p0.y, nullptr, u0, v0);
        undistort_pt(g.cam1, (double)p1.x, (double)p1.y, nullptr, u1, v1);
        if (!has_new) {                 // all-float32 point list -> cv2 returns float32 (B11)
            u0 = (double)(float)u0;
            v0 = (double)(float)v0;
        }
        out_ids(ob)[pos] = id;
        double* m = out_meas(ob, g.NMAX) + 4 * pos;
        m[0] = u0;
        m[1] = v0;
        m[2] = (double)(float)u1;
        m[3] = (double)(float)v1;
        out_cell(ob, g.NMAX)[pos] = c;
        out_life(ob, g.NMAX)[pos] = cur.life[base + i];
        out_p0(ob, g.NMAX)[2 * pos] = p0.x;
        out_p0(ob, g.NMAX)[2 * pos + 1] = p0.y;
        out_p1(ob, g.NMAX)[2 * pos] = p1.x;
        out_p1(ob, g.NMAX)[2 * pos + 1] = p1.y;
    }
    if (tid == 0) {
        avb_frame_header* h = reinterpret_cast<avb_frame_header*>(ob);
        const int* cnt = d.counters + s * 8;
        h->n_features = total;
        h->next_feature_id = next_id + total_new;
        h->before_tracking = cnt[0];
        h->after_tracking = cnt[1];
        h->after_matching = cnt[2];
        h->after_ransac = cnt[2];       // RANSAC is an all-ones stub in the reference (B2)
        h->has_new = has_new;
        h->n_fast = cnt[3];
        h->n_candidates = cnt[4];
        h->frame_index = d.frame_index[s];
        d.next_id[s] = next_id + total_new;
        d.frame_index[s] += 1;
    }
}

void launch_publish(const Geom& g, const DevState& d, int parity, cudaStream_t st) {
    k_publish<<<g.S, 256, 0, st>>>(g, d, parity);
}
