// K2: FAST-9/16 corner score + 3x3 non-maximum suppression + per-cell bucketing, one pass over cam0.
// Replaces cv2.FastFeatureDetector_create(thr).detect (reference: pipeline.py:23-25,
// feature_initializer.py:52, feature_adder.py:64).  Semantics:
//   best  = max over the 16 arcs of 9 contiguous ring pixels of min(ring - c) / min(c - ring)
//   corner <=> best > thr;  response = best - 1;  tested only for 3 <= x <= W-4, 3 <= y <= H-4
//   keypoint <=> response strictly greater than the 8 neighbours' responses (non-corners = 0)
// The score map never goes to HBM: each CTA scores a (64+2) x (32+2) patch in shared memory and
// suppresses its 64x32 interior.  Keypoints are appended to the bucket of their grid cell as a
// 32-bit key (response << 24 | inverted scan index) so every later ranking is a plain integer max
// that equals the reference's stable sort by response (Appendix B9).
#include "avb_common.cuh"

#define FT_W 64
#define FT_H 32
#define FB_W 96                     // TMA box: 16 left halo (TMA start column must be a 16-byte multiple) + FT_W + 16
#define FB_X 16
#define FB_H 40                     // FT_H + 8
#define SC_PITCH 68

__device__ __forceinline__ void f_mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void f_mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void f_mbar_wait(uint64_t* bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"((unsigned)__cvta_generic_to_shared(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void f_tma_load_3d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int x, int y, int z) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
            (unsigned)__cvta_generic_to_shared(smem_dst)),
        "l"(map), "r"((unsigned)__cvta_generic_to_shared(bar)), "r"(x), "r"(y), "r"(z)
        : "memory");
}

__device__ __forceinline__ unsigned run9(unsigned m) {      // any 9 contiguous set bits on the 16-ring?
    m |= m << 16;
    unsigned a = m & (m >> 1);
    a &= a >> 2;
    a &= a >> 4;
    a &= m >> 8;
    return a & 0xFFFFu;
}

// Ring of pixel p (pointer into the shared tile): the 16 Bresenham-circle samples in cv2's order.
__device__ __forceinline__ void fast_ring(const uint8_t* p, int (&v)[16]) {
    v[0] = p[3 * FB_W];       v[1] = p[3 * FB_W + 1];   v[2] = p[2 * FB_W + 2];   v[3] = p[FB_W + 3];
    v[4] = p[3];              v[5] = p[-FB_W + 3];      v[6] = p[-2 * FB_W + 2];  v[7] = p[-3 * FB_W + 1];
    v[8] = p[-3 * FB_W];      v[9] = p[-3 * FB_W - 1];  v[10] = p[-2 * FB_W - 2]; v[11] = p[-FB_W - 3];
    v[12] = p[-3];            v[13] = p[FB_W - 3];      v[14] = p[2 * FB_W - 2];  v[15] = p[3 * FB_W - 1];
}

// Corner test: 0 = no corner, 1 = 9 contiguous ring pixels brighter than c + thr, 2 = darker than c - thr
// (both at once is impossible: 18 > 16).  A 9-arc always contains two ADJACENT cardinal samples (0, 4, 8, 12), so four
// loads reject most pixels; the survivors build the two 16-bit masks with one subtract + one funnel shift per
// sample and polarity (the sign bit of (c + thr) - v is "brighter", of v - (c - thr) "darker").
__device__ __forceinline__ int fast_is_corner(const uint8_t* p, int thr) {
    const int hi = p[0] + thr, lo = p[0] - thr;
    {
        const int a = p[3 * FB_W], b = p[3], c = p[-3 * FB_W], e = p[-3];
        const unsigned ba = a > hi, bb = b > hi, bc = c > hi, be = e > hi;
        const unsigned da = a < lo, db = b < lo, dc = c < lo, de = e < lo;
        const unsigned anyb = (ba & bb) | (bb & bc) | (bc & be) | (be & ba);
        const unsigned anyd = (da & db) | (db & dc) | (dc & de) | (de & da);
        if (!(anyb | anyd)) return 0;
    }
    int v[16];
    fast_ring(p, v);
    unsigned br = 0, dk = 0;
#pragma unroll
    for (int i = 15; i >= 0; --i) {             // bit i of the mask = sample i
        br = __funnelshift_l((unsigned)(hi - v[i]), br, 1);
        dk = __funnelshift_l((unsigned)(v[i] - lo), dk, 1);
    }
    return run9(br) ? 1 : (run9(dk) ? 2 : 0);
}

// cv2 response of a pixel known to be a corner of polarity `pol` (1 bright, 2 dark):
// max over the 16 arcs of 9 contiguous ring pixels of min |ring - c| in that polarity, minus 1.
__device__ __forceinline__ int fast_score(const uint8_t* p, int pol) {
    const int c = p[0];
    int v[16], dd[16], m2[16], m4[16], m8[16];
    fast_ring(p, v);
#pragma unroll
    for (int i = 0; i < 16; ++i) dd[i] = pol == 2 ? (c - v[i]) : (v[i] - c);
#pragma unroll
    for (int i = 0; i < 16; ++i) m2[i] = min(dd[i], dd[(i + 1) & 15]);
#pragma unroll
    for (int i = 0; i < 16; ++i) m4[i] = min(m2[i], m2[(i + 2) & 15]);
#pragma unroll
    for (int i = 0; i < 16; ++i) m8[i] = min(m4[i], m4[(i + 4) & 15]);
    int best = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) best = max(best, min(m8[i], dd[(i + 8) & 15]));
    return best - 1;                // best > thr for a corner, so this is >= thr
}

#define FAST_POS ((FT_H + 2) * (FT_W + 2))

// Three passes over the tile: (A) cheap corner test on every pixel of the tile + 1-pixel ring, corners compacted
// into a shared list; (B) the expensive score only for listed corners, one per thread (no lane idles through
// someone else's score); (C) strict 3x3 non-maximum suppression and bucketing.
__global__ void __launch_bounds__(256) k_fast(const __grid_constant__ CUtensorMap map0, const __grid_constant__ Geom g,
                                              const __grid_constant__ DevState d) {
    __shared__ __align__(128) uint8_t tile[FB_H][FB_W];
    __shared__ uint8_t sc[FT_H + 2][SC_PITCH];
    __shared__ unsigned short clist[FAST_POS];
    __shared__ int ccount;
    __shared__ __align__(8) uint64_t bar;

    const int tid = threadIdx.x;
    const int s = blockIdx.z;
    const int X0 = FT_W * blockIdx.x, Y0 = FT_H * blockIdx.y;

    if (tid == 0) {
        f_mbar_init(&bar, 1);
        ccount = 0;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        f_mbar_expect_tx(&bar, FB_W * FB_H);
        f_tma_load_3d(&tile[0][0], &map0, &bar, X0 - FB_X, Y0 - 4, s * 2);   // image = s*2 + cam0
    }
    f_mbar_wait(&bar, 0);

    for (int i = tid; i < FAST_POS; i += 256) {
        const int py = i / (FT_W + 2) - 1, px = i % (FT_W + 2) - 1;
        const int X = X0 + px, Y = Y0 + py;
        int pol = 0;
        if (X >= 3 && X <= g.W - 4 && Y >= 3 && Y <= g.H - 4) pol = fast_is_corner(&tile[py + 4][px + FB_X], g.fast_thr);
        sc[py + 1][px + 1] = 0;
        if (pol) clist[atomicAdd(&ccount, 1)] = (unsigned short)(i | (pol << 14));
    }
    __syncthreads();
    const int nc = ccount;
    for (int j = tid; j < nc; j += 256) {
        const int e = clist[j], i = e & 0x3fff, pol = e >> 14;
        const int py = i / (FT_W + 2) - 1, px = i % (FT_W + 2) - 1;
        sc[py + 1][px + 1] = (uint8_t)fast_score(&tile[py + 4][px + FB_X], pol);
    }
    __syncthreads();

    // (C) strict 3x3 NMS, only at the listed corners that lie inside the tile proper
    for (int j = tid; j < nc; j += 256) {
        const int i = clist[j] & 0x3fff;
        const int py = i / (FT_W + 2) - 1, px = i % (FT_W + 2) - 1;
        if (py < 0 || py >= FT_H || px < 0 || px >= FT_W) continue;
        const int v = sc[py + 1][px + 1];
        const bool kp = v > sc[py][px] && v > sc[py][px + 1] && v > sc[py][px + 2] && v > sc[py + 1][px] &&
                        v > sc[py + 1][px + 2] && v > sc[py + 2][px] && v > sc[py + 2][px + 1] && v > sc[py + 2][px + 2];
        if (kp) {
            const int X = X0 + px, Y = Y0 + py;
            const int cell = (Y / g.gh) * g.cols + (X / g.gw);
            const int pos = atomicAdd(&d.kp_count[s * g.NC + cell], 1);
            if (pos < g.KPC) d.kp_key[((size_t)s * g.NC + cell) * g.KPC + pos] = kp_make_key(v, X, Y, g.W);
        }
    }
}

void launch_fast(const Geom& g, const DevState& d, const PyrMaps& maps, int parity, cudaStream_t st) {
    dim3 grid((g.W + FT_W - 1) / FT_W, (g.H + FT_H - 1) / FT_H, g.S);
    k_fast<<<grid, 256, 0, st>>>(maps.fast0[parity], g, d);
}
