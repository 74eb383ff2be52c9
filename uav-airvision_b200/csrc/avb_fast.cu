// K2: FAST-9/16 corner score + 3x3 non-maximum suppression + per-cell bucketing, one pass over cam0.
// Replaces cv2.FastFeatureDetector_create(thr).detect (reference: pipeline.py:23-25,
// feature_initializer.py:52, feature_adder.py:64).  Semantics:
//   best  = max over the 16 arcs of 9 contiguous ring pixels of min(ring - c) / min(c - ring)
//   corner <=> best > thr;  response = best - 1;  tested only for 3 <= x <= W-4, 3 <= y <= H-4
//   keypoint <=> response strictly greater than the 8 neighbours' responses (non-corners = 0)
// The score map never goes to HBM: each CTA scores a (64+2) x (16+2) patch in shared memory and
// suppresses its 64x16 interior.  Keypoints are appended to the bucket of their grid cell as a
// 32-bit key (response << 24 | inverted scan index) so every later ranking is a plain integer max
// that equals the reference's stable sort by response (Appendix B9).
#include "avb_common.cuh"

#define FT_W 64
#define FT_H 16
#define FB_W 96                     // TMA box: 16 left halo (TMA start column must be a 16-byte multiple) + FT_W + 16
#define FB_X 16
#define FB_H 24                     // FT_H + 8
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

// Returns the cv2 response (best - 1) or 0.  p points at the centre pixel inside the shared tile.
__device__ __forceinline__ int fast_response(const uint8_t* p, int thr) {
    const int c = p[0];
    int v[16];
    v[0] = p[3 * FB_W];       v[1] = p[3 * FB_W + 1];   v[2] = p[2 * FB_W + 2];   v[3] = p[FB_W + 3];
    v[4] = p[3];              v[5] = p[-FB_W + 3];      v[6] = p[-2 * FB_W + 2];  v[7] = p[-3 * FB_W + 1];
    v[8] = p[-3 * FB_W];      v[9] = p[-3 * FB_W - 1];  v[10] = p[-2 * FB_W - 2]; v[11] = p[-FB_W - 3];
    v[12] = p[-3];            v[13] = p[FB_W - 3];      v[14] = p[2 * FB_W - 2];  v[15] = p[3 * FB_W - 1];
    unsigned br = 0, dk = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        br |= (unsigned)(v[i] > c + thr) << i;
        dk |= (unsigned)(v[i] < c - thr) << i;
    }
    if (!(run9(br) | run9(dk))) return 0;
    int best = 0;
#pragma unroll
    for (int pol = 0; pol < 2; ++pol) {
        int dd[16], m2[16], m4[16], m8[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) dd[i] = pol ? (c - v[i]) : (v[i] - c);
#pragma unroll
        for (int i = 0; i < 16; ++i) m2[i] = min(dd[i], dd[(i + 1) & 15]);
#pragma unroll
        for (int i = 0; i < 16; ++i) m4[i] = min(m2[i], m2[(i + 2) & 15]);
#pragma unroll
        for (int i = 0; i < 16; ++i) m8[i] = min(m4[i], m4[(i + 4) & 15]);
#pragma unroll
        for (int i = 0; i < 16; ++i) best = max(best, min(m8[i], dd[(i + 8) & 15]));
    }
    return best - 1;                // best > thr here, so this is >= thr
}

__global__ void __launch_bounds__(256) k_fast(const __grid_constant__ CUtensorMap map0, Geom g, DevState d) {
    __shared__ __align__(128) uint8_t tile[FB_H][FB_W];
    __shared__ uint8_t sc[FT_H + 2][SC_PITCH];
    __shared__ __align__(8) uint64_t bar;

    const int tid = threadIdx.x;
    const int s = blockIdx.z;
    const int X0 = FT_W * blockIdx.x, Y0 = FT_H * blockIdx.y;

    if (tid == 0) {
        f_mbar_init(&bar, 1);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        f_mbar_expect_tx(&bar, FB_W * FB_H);
        f_tma_load_3d(&tile[0][0], &map0, &bar, X0 - FB_X, Y0 - 4, s * 2);   // image = s*2 + cam0
    }
    f_mbar_wait(&bar, 0);

    // responses on the tile plus a one-pixel ring
    for (int i = tid; i < (FT_H + 2) * (FT_W + 2); i += 256) {
        const int py = i / (FT_W + 2) - 1, px = i % (FT_W + 2) - 1;
        const int X = X0 + px, Y = Y0 + py;
        int r = 0;
        if (X >= 3 && X <= g.W - 4 && Y >= 3 && Y <= g.H - 4) r = fast_response(&tile[py + 4][px + FB_X], g.fast_thr);
        sc[py + 1][px + 1] = (uint8_t)r;
    }
    __syncthreads();

    for (int i = tid; i < FT_H * FT_W; i += 256) {
        const int py = i / FT_W, px = i % FT_W;
        const int v = sc[py + 1][px + 1];
        bool kp = v > 0;
        if (kp) {
            kp = v > sc[py][px] && v > sc[py][px + 1] && v > sc[py][px + 2] && v > sc[py + 1][px] &&
                 v > sc[py + 1][px + 2] && v > sc[py + 2][px] && v > sc[py + 2][px + 1] && v > sc[py + 2][px + 2];
        }
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
