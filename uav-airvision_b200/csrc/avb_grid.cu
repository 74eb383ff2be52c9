// Grid bookkeeping kernels: the reference's per-cell Python list logic (Appendix B of SURVEY.md) on device.
//   k_select       FeatureAdder mask + sieve + per-cell cap (feature_adder.py:59-77); on frame 0 the per-cell
//                  top-`min` ranking of FeatureInitializer (feature_initializer.py:79-80)
//   k_finish       tracker re-bin (feature_tracker.py:139-155) + adder top-`min` insert (feature_adder.py:102-108)
//                  + pruner (feature_pruner.py:13-19) + id assignment in cell-major order (B10)
//                  + FeaturePublisher.publish (feature_publisher.py:90-121), one CTA per stream
// Every ranking is max-by-key: key = (response << 24 | inverted scan index) reproduces Python's stable
// sorted(..., reverse=True) over detections in scan order (B9); lifetime ranking is stable by list position.
#include <algorithm>

#include "avb_lk.cuh"

// FAST bucket counts.  Not part of the frame chain: between frames the counts ARE zero -- the last reader of a cell's count
// (k_select in mode 0 / 1, one CTA per cell) zeroes it, as k_finish does with the per-frame counters -- so k_fast appends
// without a clearing launch ahead of it.  This kernel serves the stage API (avb_fast_detect), which has no k_select.
__global__ void k_clear_frame(Geom g, DevState d) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < g.S * g.NC) d.kp_count[i] = 0;
}

void launch_clear_frame(const Geom& g, const DevState& d, cudaStream_t st) {
    const int n = g.S * g.NC;
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
// mode 0: steady state (FeatureAdder); 1: first frame (FeatureInitializer ranking of the stereo inliers);
// mode 2: SPECULATIVE list (few streams, Geom::spec_k): the spec_k strongest keypoints of the cell, unmasked -- whatever
//         the mask removes later, the adder's candidates are the first unmasked ones of this order, so they are in this
//         list unless more than spec_k - gmax keypoints above them get masked.  Their stereo match (k_spec_match) runs
//         BESIDE k_track instead of behind it; mode 0 then copies the result of every candidate it finds in the list
//         and marks the others (c_ok = 2) for k_stereo_candidates, which matches only those.
#define SEL_WARPS 8
#define SEL_MAXF 128               // live features that can touch one cell: gmax per cell x the 9 cells around, generously
template <bool SPEC>                // SPEC: the speculative look-up + in-kernel matching of the misses is compiled in (115 registers
                                    // against 32: the many-stream launches, which never speculate, keep the small kernel)
__global__ void __launch_bounds__(32 * SEL_WARPS) k_select(const __grid_constant__ Geom g, const __grid_constant__ DevState d,
                                                          int parity, int mode) {
    extern __shared__ unsigned smem_u32[];
    unsigned* keys = smem_u32;
    __shared__ unsigned feat[SEL_MAXF];
    __shared__ int n_feat;
    __shared__ unsigned long long fin[SEL_WARPS * AVB_MAX_CAP];
    __shared__ unsigned skeys[32 * SEL_WARPS];
    __shared__ unsigned miss_key[AVB_MAX_CAP];  // candidates the speculative list does not hold: matched here, one warp each
    __shared__ int miss_slot[AVB_MAX_CAP];
    __shared__ int n_miss;

    if (mode != 2) {                // the speculative list depends on FAST alone and is not a link of the PDL chain
        pdl_wait();
        pdl_launch_dependents();
    }
    const bool first_frame = mode == 1;
    const int cell = blockIdx.x, s = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const size_t kbase = ((size_t)s * g.NC + cell) * g.KPC;
    // every global load that does not depend on another one is issued here, together: the kernel is a chain of memory
    // round trips otherwise (it ran 9.4 us for a few hundred bytes of work)
    const int n_raw = d.kp_count[s * g.NC + cell];
    const unsigned key_raw = tid < g.KPC ? d.kp_key[kbase + tid] : 0u;          // valid when tid < n
    const uint8_t ok_raw = (first_frame && tid < g.KPC) ? d.kp_ok[kbase + tid] : (uint8_t)1;
    const int k = mode == 2 ? g.spec_k : (first_frame ? g.gmin : g.gmax);

    if (tid == 0) n_feat = n_miss = 0;
    __syncthreads();
    // every thread's load of the count is behind the barrier: the final selection of the frame (not the speculative list,
    // which runs first) leaves the bucket empty for the next frame's k_fast
    if (mode != 2 && tid == 0) d.kp_count[s * g.NC + cell] = 0;
    if (mode == 0) {
        const int cx0 = (cell % g.cols) * g.gw, cy0 = (cell / g.cols) * g.gh;
        for (int i0 = 0; i0 < g.NMAX; i0 += 2 * 32 * SEL_WARPS) {
            const int ia = i0 + tid, ib = ia + 32 * SEL_WARPS;
            const size_t fb = (size_t)s * g.NMAX;
            const int ca = ia < g.NMAX ? d.t_cell[fb + ia] : -1, cb = ib < g.NMAX ? d.t_cell[fb + ib] : -1;
            const float2 pa = ia < g.NMAX ? d.t_p0[fb + ia] : make_float2(0.f, 0.f);
            const float2 pb = ib < g.NMAX ? d.t_p0[fb + ib] : make_float2(0.f, 0.f);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int c = h ? cb : ca;
                const float2 p = h ? pb : pa;
                if (c >= 0) {
                    const int fx = (int)p.x, fy = (int)p.y;         // int() truncation (B7)
                    // mask[y-3:y+4, x-3:x+4] = 0 with a negative slice start selects nothing (B7)
                    if (fx >= 3 && fy >= 3 && fx >= cx0 - 3 && fx < cx0 + g.gw + 3 && fy >= cy0 - 3 && fy < cy0 + g.gh + 3) {
                        const int pos = atomicAdd(&n_feat, 1);
                        if (pos < SEL_MAXF) feat[pos] = ((unsigned)fy << 16) | (unsigned)fx;
                    }
                }
            }
        }
    }
    __syncthreads();
    const int n = min(n_raw, g.KPC);
    const int nf_all = n_feat, nf = min(nf_all, SEL_MAXF);
    auto masked_key = [&](unsigned key, uint8_t ok) -> unsigned {
        if (first_frame) return ok ? key : 0u;
        if (mode != 0) return key;
        int resp, x, y;
        kp_decode(key, g.W, resp, x, y);
        bool masked = false;
        for (int j = 0; j < nf; ++j) {
            const int fx = (int)(feat[j] & 0xffffu), fy = (int)(feat[j] >> 16);
            masked |= (abs(x - fx) <= 3) && (abs(y - fy) <= 3);
        }
        if (nf_all > SEL_MAXF) {                                    // overflow of the local list: fall back to the full table
            for (int j = 0; j < g.NMAX && !masked; ++j) {
                if (d.t_cell[(size_t)s * g.NMAX + j] < 0) continue;
                const float2 p = d.t_p0[(size_t)s * g.NMAX + j];
                const int fx = (int)p.x, fy = (int)p.y;
                masked = fx >= 3 && fy >= 3 && (abs(x - fx) <= 3) && (abs(y - fy) <= 3);
            }
        }
        return masked ? 0u : key;
    };
    const size_t sbase = ((size_t)s * g.NC + cell) * (size_t)max(g.spec_k, 1);
    const bool lookup = SPEC && mode == 0 && g.spec_k > 0;
    const int sn = lookup ? d.s_n[s * g.NC + cell] : 0;

    if (n <= 32 * SEL_WARPS) {
        // The usual case: one key per thread.  Rank = number of larger keys (keys of distinct keypoints are distinct: a
        // key holds the scan position), read from shared memory: no reduction rounds, no serial tail; the thread that
        // owns rank r writes list entry r and does its own look-up in the speculative list.
        const unsigned key = tid < n ? masked_key(key_raw, ok_raw) : 0u;
        skeys[tid] = key;
        const int live = __syncthreads_count(key != 0u);
        int rank = 0;
        for (int j = 0; j < n; ++j) rank += skeys[j] > key;
        if (key != 0u && rank < k) {
            if (mode == 2) {
                d.s_key[sbase + rank] = key;
            } else {
                const size_t o = (size_t)s * g.NMAX + cell * g.gmax + rank;
                d.c_key[o] = key;
                d.c_src[o] = tid;
                if (first_frame) {
                    d.c_p1[o] = d.kp_p1[kbase + tid];
                    d.c_ok[o] = 1;
                } else if (lookup) {
                    int hit = -1;
                    for (int j = 0; j < sn; ++j)
                        if (d.s_key[sbase + j] == key) hit = j;
                    if (hit >= 0) {                                 // matched speculatively: take the result
                        const uint8_t ok = d.s_ok[sbase + hit];
                        d.c_p1[o] = d.s_p1[sbase + hit];
                        d.c_ok[o] = ok;
                        if (ok) d.c_und[o] = d.s_und[sbase + hit];
                    } else {                                        // not in the list: matched below
                        const int m = atomicAdd(&n_miss, 1);
                        miss_key[m] = key;
                        miss_slot[m] = rank;
                    }
                }
            }
        }
        if (tid == 0) {
            const int found = min(live, k);
            if (mode == 2) {
                d.s_n[s * g.NC + cell] = found;
            } else {
                d.c_count[s * g.NC + cell] = found;
                atomicAdd(&d.counters[s * 8 + 3], n);
                atomicAdd(&d.counters[s * 8 + 4], found);
            }
        }
    } else {

    // More keypoints in the cell than threads (dense noise): top-k by reduction rounds.  Every warp extracts the k largest
    // keys of its strided share with warp-max rounds, then warp 0 extracts the k largest of the 8k finalists.
    for (int i = tid; i < n; i += 32 * SEL_WARPS) keys[i] = masked_key(d.kp_key[kbase + i], first_frame ? d.kp_ok[kbase + i] : (uint8_t)1);
    __syncthreads();
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
        const unsigned skey = (lookup && lane < sn) ? d.s_key[sbase + lane] : 0u;  // the speculative list: one entry per lane
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
            const unsigned key = (unsigned)(best >> 32);
            const unsigned hit = lookup ? __ballot_sync(0xffffffffu, skey == key) : 0u;
            if (lane == 0) {
                if (mode == 2) {
                    d.s_key[sbase + r] = key;
                } else {
                    const size_t o = (size_t)s * g.NMAX + cell * g.gmax + r;
                    const int src = (int)(best & 0xffffffffu);
                    d.c_key[o] = key;
                    d.c_src[o] = src;
                    if (first_frame) {
                        d.c_p1[o] = d.kp_p1[kbase + src];
                        d.c_ok[o] = 1;
                    } else if (lookup) {
                        if (hit) {
                            const size_t si = sbase + (__ffs(hit) - 1);
                            const uint8_t ok = d.s_ok[si];
                            d.c_p1[o] = d.s_p1[si];
                            d.c_ok[o] = ok;
                            if (ok) d.c_und[o] = d.s_und[si];
                        } else {
                            const int m = atomicAdd(&n_miss, 1);
                            miss_key[m] = key;
                            miss_slot[m] = r;
                        }
                    }
                }
            }
            ++found;
        }
        if (lane == 0) {
            if (mode == 2) {
                d.s_n[s * g.NC + cell] = found;
            } else {
                d.c_count[s * g.NC + cell] = found;
                atomicAdd(&d.counters[s * 8 + 3], n);
                atomicAdd(&d.counters[s * 8 + 4], found);
            }
        }
    }
    }
    if (!SPEC || !lookup) return;
    // Candidates the speculative list did not hold (more than spec_k - gmax stronger keypoints were masked: rare) are
    // stereo-matched here, one warp each (stereo_matcher.py:33-115 through the one-warp mapping)
    __syncthreads();
    if constexpr (SPEC) {
        for (int m = warp; m < n_miss; m += SEL_WARPS) {
            int resp, x, y;
            kp_decode(miss_key[m], g.W, resp, x, y);
            const ChainResult r = feature_chain<1>(g, d, s, parity, false, (float)x, (float)y, 0.f, 0.f, nullptr);
            if (lane == 0) {
                const size_t o = (size_t)s * g.NMAX + cell * g.gmax + miss_slot[m];
                d.c_p1[o] = make_float2(r.x1, r.y1);
                d.c_ok[o] = r.matched ? 1 : 0;
                if (r.matched) d.c_und[o] = make_double4(r.u0, r.v0, r.u1, r.v1);
            }
        }
    }
}

int avb_set_smem_limits(size_t select_bytes, size_t grid_bytes) {
    cudaError_t e = cudaSuccess;
    if (select_bytes > 44 * 1024)      // the kernels also hold ~4 KB of static shared memory
        e = cudaFuncSetAttribute(k_select<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)select_bytes);
    if (e == cudaSuccess && select_bytes > 44 * 1024)
        e = cudaFuncSetAttribute(k_select<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)select_bytes);
    if (e == cudaSuccess && grid_bytes > 16 * 1024)      // k_finish also holds ~29 KB of static shared memory
        e = cudaFuncSetAttribute(k_finish, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)grid_bytes);
    return e == cudaSuccess ? 0 : -1;
}

void launch_select(const Geom& g, const DevState& d, int parity, int first_frame, cudaStream_t st) {
    const size_t smem = (size_t)g.KPC * sizeof(unsigned);
    if (g.spec_k > 0 && !first_frame)
        launch_k(k_select<true>, dim3(g.NC, g.S), dim3(32 * SEL_WARPS), smem, st, g_avb_pdl != 0, g, d, parity, 0);
    else
        launch_k(k_select<false>, dim3(g.NC, g.S), dim3(32 * SEL_WARPS), smem, st, g_avb_pdl && !first_frame, g, d, parity, first_frame ? 1 : 0);
}
void launch_spec_select(const Geom& g, const DevState& d, cudaStream_t st) {
    const size_t smem = (size_t)g.KPC * sizeof(unsigned);
    launch_k(k_select<false>, dim3(g.NC, g.S), dim3(32 * SEL_WARPS), smem, st, false, g, d, 0, 2);
}

// k_finish: one 1024-thread CTA per stream closes the frame.
//   phase 1  per-cell survivor histogram (tracked features by their NEW cell) and per-cell count of new features
//            (the first `gmin` stereo inliers of the cell's candidate list, feature_adder.py:102-108, B8).  Everything
//            phase 3 needs to know about OTHER cells follows from these two numbers per cell: final count
//            min(nt + nn, gmax), id offsets (B10), and whether the frame holds fresh features (dtype quirk B11).
//   phase 2  exclusive scans over the cells
//   phase 3  one warp per cell, one lane per output slot: stable placement of the survivors in the previous frame's
//            flattened grid order (feature_tracker.py:139-155, B16), append the new features, prune to `gmax` by
//            lifetime with a stable rank (feature_pruner.py:13-19, B9), then straight from registers: id assignment,
//            undistortion with the owning camera's model and the packed result block
//            (feature_publisher.py:90-121) plus the grid table of the new frame.
// The global loads of a phase are independent of each other (one memory round trip per phase); the values a slot
// needs travel in registers from the gather to the result block.
// Dynamic smem: cellof[NMAX] | seg[NMAX] | slife[NMAX] (int).
#define FIN_THREADS 1024
__global__ void __launch_bounds__(FIN_THREADS) k_finish(const __grid_constant__ Geom g, const __grid_constant__ DevState d, int parity,
                                                        int first_frame) {
    extern __shared__ int smem_i32[];
    int* cellof = smem_i32;
    int* seg = smem_i32 + g.NMAX;
    int* slife = smem_i32 + 2 * g.NMAX;         // lifetime by table slot
    __shared__ int cnt[AVB_MAX_CELLS + 1];      // tracked survivors per (new) cell
    __shared__ int nnew[AVB_MAX_CELLS + 1];     // new features of the cell (before pruning)
    __shared__ int off[AVB_MAX_CELLS + 1];      // segment start of a cell in seg[]
    __shared__ int off_cnt[AVB_MAX_CELLS + 1];  // exclusive scan of the final counts
    __shared__ int off_new[AVB_MAX_CELLS + 1];  // exclusive scan of the new-feature counts
    __shared__ unsigned cmask[AVB_MAX_CELLS];   // stereo inliers among the cell's candidates (bit = list position)
    __shared__ int s_fresh;                     // any new feature survives pruning
    __shared__ int route[FIN_THREADS / 32][32];

    pdl_wait();
    pdl_launch_dependents();
#ifdef AVB_DEBUG_CLOCKS
    long long dbg_t[5];
    dbg_t[0] = clock64();
#endif
    const int s = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned lt = (1u << lane) - 1u;
    const GridTable prev = d.grid[parity ^ 1], cur = d.grid[parity];
    const size_t base = (size_t)s * g.NMAX;

    // Every global load of phase 1 is issued before anything waits on one (the kernel is a chain of memory round trips:
    // the first slice of the table and of the candidate lists -- all there is for up to 1024 slots / 32 cells per CTA --
    // the id counter and the frame index travel together)
    const long long next_id = d.next_id[(parity ^ 1) * g.S + s];
    const int c_first = (!first_frame && tid < g.NMAX) ? d.t_cell[base + tid] : -1;
    const int li_first = (!first_frame && tid < g.NMAX) ? prev.life[base + tid] : 0;
    const int nc_first = warp < g.NC ? d.c_count[s * g.NC + warp] : 0;
    const uint8_t ok_first = (warp < g.NC && lane < g.gmax) ? d.c_ok[base + warp * g.gmax + lane] : (uint8_t)0;
    for (int c = tid; c <= g.NC; c += FIN_THREADS) cnt[c] = 0;
    if (tid == 0) s_fresh = 0;
    __syncthreads();
    for (int i = tid; i < g.NMAX; i += FIN_THREADS) {
        const int c = i == tid ? c_first : (first_frame ? -1 : d.t_cell[base + i]);
        const int li = i == tid ? li_first : (first_frame ? 0 : prev.life[base + i]);
        cellof[i] = c;
        slife[i] = li;
        if (c >= 0) atomicAdd(&cnt[c], 1);
    }
    for (int c = warp; c < g.NC; c += FIN_THREADS / 32) {       // new features per cell
        const int nc = c == warp ? nc_first : d.c_count[s * g.NC + c];          // c_ok past the count is stale but masked
        const uint8_t okv = c == warp ? ok_first : (lane < g.gmax ? d.c_ok[base + c * g.gmax + lane] : (uint8_t)0);
        const unsigned bn = __ballot_sync(0xffffffffu, lane < nc && okv);
        if (lane == 0) {
            nnew[c] = min(__popc(bn), g.gmin);
            cmask[c] = bn;
        }
    }
    __syncthreads();
#ifdef AVB_DEBUG_CLOCKS
    dbg_t[1] = clock64();
#endif
    if (warp == 0) {                    // exclusive scans over the cells
        int run_s = 0, run_c = 0, run_n = 0, fresh = 0;
        for (int b0 = 0; b0 < g.NC; b0 += 32) {
            const int c = b0 + lane;
            const int vs = c < g.NC ? cnt[c] : 0;
            const int vn = c < g.NC ? nnew[c] : 0;
            const int vc = min(vs + vn, g.gmax);
            fresh |= (vn > 0 && vs < g.gmax) ? 1 : 0;           // a lifetime-1 feature ranks after every tracked one
            int is = vs, ic = vc, in = vn;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int ts = __shfl_up_sync(0xffffffffu, is, o), tc = __shfl_up_sync(0xffffffffu, ic, o),
                          tn = __shfl_up_sync(0xffffffffu, in, o);
                if (lane >= o) {
                    is += ts;
                    ic += tc;
                    in += tn;
                }
            }
            if (c < g.NC) {
                off[c] = run_s + is - vs;
                off_cnt[c] = run_c + ic - vc;
                off_new[c] = run_n + in - vn;
            }
            run_s += __shfl_sync(0xffffffffu, is, 31);
            run_c += __shfl_sync(0xffffffffu, ic, 31);
            run_n += __shfl_sync(0xffffffffu, in, 31);
        }
        fresh = __any_sync(0xffffffffu, fresh);
        if (lane == 0) {
            off_cnt[g.NC] = run_c;
            off_new[g.NC] = run_n;
            s_fresh = fresh;
        }
    }
    __syncthreads();
    const int total = off_cnt[g.NC], total_new = off_new[g.NC], has_new = s_fresh;
    uint8_t* ob = d.out + (size_t)s * out_stride_bytes(g.NMAX);
#ifdef AVB_DEBUG_CLOCKS
    dbg_t[2] = clock64();
#endif

    // the cells are dealt out to the CTAs of the stream (gridDim.y), one warp each; phases 1-2 are cheap and repeated
    // by every CTA so that no inter-CTA synchronisation exists
    for (int c = blockIdx.y * (FIN_THREADS / 32) + warp; c < g.NC; c += gridDim.y * (FIN_THREADS / 32)) {
        const int nt = cnt[c], o0 = off[c], nn = nnew[c];
        // candidate side first: its loads do not depend on the placement below
        const size_t cb = base + (size_t)c * g.gmax;
        const unsigned bn = cmask[c];
        const bool okj = (bn >> lane) & 1u;
        const int nrank = __popc(bn & lt);
        const bool is_new = okj && nrank < g.gmin;
        unsigned ckey = 0;
        float2 cp1 = make_float2(0.f, 0.f);
        if (is_new) {
            ckey = d.c_key[cb + lane];
            cp1 = d.c_p1[cb + lane];
        }
#ifdef AVB_DEBUG_CLOCKS
        long long q0 = clock64(), q1 = 0, q2 = 0, q3 = 0, q4 = 0;
#endif
        if (nt) {                       // stable placement: survivors of this cell in table order
            int placed = 0;
            for (int b0 = 0; b0 < g.NMAX && placed < nt; b0 += 32) {
                const int wi = b0 + lane;
                const bool v = wi < g.NMAX && cellof[wi] == c;
                const unsigned b = __ballot_sync(0xffffffffu, v);
                if (v) seg[o0 + placed + __popc(b & lt)] = wi;
                placed += __popc(b);
            }
        }
        __syncwarp();
        const int total_c = nt + nn, fc = min(total_c, g.gmax);
        const bool prune = total_c > g.gmax;

        // where does each source land?  tracked survivor i -> pos (rank by lifetime when pruning), new feature -> nt + nrank
        // gmax <= 32: one lane per output slot; sources are routed through shared memory (tracked) or shuffles (new)
#ifdef AVB_DEBUG_CLOCKS
        q1 = clock64();
#endif
        route[warp][lane] = -1;         // output slot -> table slot of the tracked feature that lands there
        __syncwarp();
        for (int i0 = 0; i0 < nt; i0 += 32) {
            const int i = i0 + lane;
            const int wi = i < nt ? seg[o0 + i] : 0;
            const int li = i < nt ? slife[wi] : 0;
            int pos = i;
            if (prune) {                // stable sort by lifetime, descending (feature_pruner.py:18)
                int rank = 0;
                if (nt <= 32) {         // the usual case: lifetimes travel by shuffle
                    for (int j = 0; j < nt; ++j) {
                        const int lj = __shfl_sync(0xffffffffu, li, j);
                        rank += (lj > li) || (lj == li && j < i);
                    }
                } else {
                    for (int j = 0; j < nt; ++j) {
                        const int lj = slife[seg[o0 + j]];
                        rank += (lj > li) || (lj == li && j < i);
                    }
                }
                pos = rank;
            }
            if (i < nt && pos < g.gmax) route[warp][pos] = wi;
        }
        __syncwarp();
        const int my_src = route[warp][lane];
        // new features occupy slots nt + nrank (lifetime 1 ranks after every tracked feature, lifetime >= 2)
        int new_from = -1;              // candidate lane whose feature lands in this output slot
        {
            const int want = lane - nt;                     // rank among the new ones this slot would hold
            if (want >= 0 && want < nn && lane < g.gmax) {
                // the lane with is_new and nrank == want: the (want+1)-th set bit of bn
                unsigned m = bn;
                for (int k = 0; k < want; ++k) m &= m - 1;
                new_from = __ffs(m) - 1;
            }
        }
        const unsigned key_in = __shfl_sync(0xffffffffu, ckey, new_from < 0 ? 0 : new_from);
        const float cpx = __shfl_sync(0xffffffffu, cp1.x, new_from < 0 ? 0 : new_from);
        const float cpy = __shfl_sync(0xffffffffu, cp1.y, new_from < 0 ? 0 : new_from);

#ifdef AVB_DEBUG_CLOCKS
        q2 = clock64();
#endif
        if (lane < fc) {
            const size_t o = cb + lane;
            const int pos = off_cnt[c] + lane;
            long long id;
            int life;
            float2 p0, p1;
            const bool fresh = my_src < 0;
            double4 un = make_double4(0.0, 0.0, 0.0, 0.0);       // normalized coordinates precomputed by the LK kernels
            if (!fresh) {
                id = prev.ids[base + my_src];
                life = slife[my_src] + 1;
                p0 = d.t_p0[base + my_src];
                p1 = d.t_p1[base + my_src];
                un = d.t_und[base + my_src];
            } else {
                if (!first_frame) un = d.c_und[cb + new_from];
                int resp, x, y;
                kp_decode(key_in, g.W, resp, x, y);
                id = next_id + off_new[c] + (lane - nt);    // ids in cell-major order (B10)
                life = 1;
                p0 = make_float2((float)x, (float)y);
                p1 = make_float2(cpx, cpy);
            }
            cur.ids[o] = id;
            cur.life[o] = life;
            cur.p0[o] = p0;
            cur.p1[o] = p1;
            cur.fresh[o] = fresh ? 1 : 0;
#ifdef AVB_DEBUG_CLOCKS
            q3 = clock64();
#endif
            double u0, v0, u1, v1;
            if (first_frame) {          // frame 0 comes through k_stereo_buckets, which keeps no normalized coordinates
                undistort_pt(g.cam0, (double)p0.x, (double)p0.y, nullptr, u0, v0);
                undistort_pt(g.cam1, (double)p1.x, (double)p1.y, nullptr, u1, v1);
            } else {                    // ChainResult::u0..v1
                u0 = un.x;
                v0 = un.y;
                u1 = un.z;
                v1 = un.w;
            }
            if (!has_new) {             // all-float32 point list -> cv2 returns float32 (B11)
                u0 = (double)(float)u0;
                v0 = (double)(float)v0;
            }
#ifdef AVB_DEBUG_CLOCKS
            q4 = clock64();
            if (tid == 0 && c == 0) {
                d.counters[s * 8 + 6] = (int)(q0 - dbg_t[2]) | ((int)(q2 - q0) << 16);
                dbg_t[3] = q4;
            }
#endif
            double* m = out_meas(ob, g.NMAX) + 4 * pos;
            out_ids(ob)[pos] = id;
            m[0] = u0;
            m[1] = v0;
            m[2] = (double)(float)u1;
            m[3] = (double)(float)v1;
            out_cell(ob, g.NMAX)[pos] = c;
            out_life(ob, g.NMAX)[pos] = life;
            out_p0(ob, g.NMAX)[2 * pos] = p0.x;
            out_p0(ob, g.NMAX)[2 * pos + 1] = p0.y;
            out_p1(ob, g.NMAX)[2 * pos] = p1.x;
            out_p1(ob, g.NMAX)[2 * pos + 1] = p1.y;
        }
        if (lane == 0) cur.count[s * g.NC + c] = fc;
#ifdef AVB_DEBUG_CLOCKS
        if (tid == 0 && c == 0) d.counters[s * 8 + 7] = (int)(clock64() - dbg_t[3]) | ((int)(dbg_t[3] - q2) << 16);
#endif
    }
#ifdef AVB_DEBUG_CLOCKS
    dbg_t[3] = clock64();
    __syncthreads();
    dbg_t[4] = clock64();
#endif
    if (tid == 0 && blockIdx.y == 0) {
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
#ifdef AVB_DEBUG_CLOCKS                 // phase durations in SM cycles instead of the counters (debug builds only)
        h->before_tracking = (int)(dbg_t[1] - dbg_t[0]);
        h->after_tracking = (int)(dbg_t[2] - dbg_t[1]);
        h->after_matching = (int)(dbg_t[3] - dbg_t[2]);
        h->after_ransac = (int)(dbg_t[4] - dbg_t[3]);
        h->n_fast = d.counters[s * 8 + 6];
        h->n_candidates = d.counters[s * 8 + 7];
#endif
        d.next_id[parity * g.S + s] = next_id + total_new;
        d.frame_index[s] += 1;
#pragma unroll
        for (int i = 0; i < 8; ++i) d.counters[s * 8 + i] = 0;     // published: zero for the next frame's kernels
    }
}

void launch_finish(const Geom& g, const DevState& d, int parity, int first_frame, cudaStream_t st) {
    const int per_cta = FIN_THREADS / 32;
    const int nct = std::min((g.NC + per_cta - 1) / per_cta, 8);
    launch_k(k_finish, dim3(g.S, nct), dim3(FIN_THREADS), (size_t)g.NMAX * 3 * sizeof(int), st, g_avb_pdl != 0, g, d, parity, first_frame);
}

// Lazy module loading (the CUDA 12 default) loads a kernel on its first launch: ~0.2 ms each, which frame 0 of a stream
// would pay for the kernels only it uses.  cudaFuncGetAttributes loads the function now (called from avb_create).
int avb_preload_grid() {
    cudaFuncAttributes a;
    cudaError_t e = cudaSuccess;
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, k_clear_frame);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, k_select<false>);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, k_select<true>);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, k_finish);
    return e == cudaSuccess ? 0 : -1;
}
