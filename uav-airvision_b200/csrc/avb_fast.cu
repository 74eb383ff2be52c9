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
#include <algorithm>

#include "avb_common.cuh"

#define FT_W 64
// Tile height TH and threads per CTA NT are template parameters of k_fast:
//   <64, 26>   many streams: (26 + 2) rows x 9 eight-pixel groups = 252 items = four full rounds of 64 threads; small CTAs =
//              many independent barrier domains per SM (measured at 64 streams: 256 threads x 32 rows 212 us, 128 x 26
//              181 us, 64 x 26 175 us, 64 x 12 198 us)
//   <256, 16>  a few streams: every CTA's corner test is ONE round ((16 + 2) x 9 = 162 items of 256 threads) and one
//              752x480 image makes 360 CTAs = one wave; the halo overhead does not matter when the GPU is mostly idle
#define FB_W 96                     // TMA box: 16 left halo (TMA start column must be a 16-byte multiple) + FT_W + 16
#define FB_X 16
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

// Ring of pixel p (pointer into the shared tile): the 16 Bresenham-circle samples in cv2's order.
__device__ __forceinline__ void fast_ring(const uint8_t* p, int (&v)[16]) {
    v[0] = p[3 * FB_W];       v[1] = p[3 * FB_W + 1];   v[2] = p[2 * FB_W + 2];   v[3] = p[FB_W + 3];
    v[4] = p[3];              v[5] = p[-FB_W + 3];      v[6] = p[-2 * FB_W + 2];  v[7] = p[-3 * FB_W + 1];
    v[8] = p[-3 * FB_W];      v[9] = p[-3 * FB_W - 1];  v[10] = p[-2 * FB_W - 2]; v[11] = p[-FB_W - 3];
    v[12] = p[-3];            v[13] = p[FB_W - 3];      v[14] = p[2 * FB_W - 2];  v[15] = p[3 * FB_W - 1];
}

// ---- pass A: the corner TEST, four pixels per 32-bit word ---------------------------------------------------------------
// A thread owns 8 adjacent pixels of one tile row (two words).  The 16 ring samples of four adjacent centres are 4-byte
// windows of 7 tile rows: each row is fetched once as four aligned words (two 64-bit shared loads) and every window is one
// byte permute with a constant selector.  "Brighter than c + thr" / "darker than c - thr" are evaluated on all four bytes
// of a word at once, the answer living in bit 7 of each byte (the other bits are don't-care all the way through):
//     v > h  <=>  (v8 & ~h8) | (~(v8 ^ h8) & carry7(v7 + ~h7))          v8/h8 top bits, v7/h7 the low seven bits
// with h = min(c + thr, 255) resp. l = max(c - thr, 0) per byte (saturation never creates a corner: nothing exceeds 255 or
// undercuts 0).  "9 contiguous ring samples" is then a bit-sliced AND/OR network over the 16 flag words (40 three-input
// logic ops per polarity).  A 9-arc always holds two ADJACENT cardinal samples (0, 4, 8, 12), so after those four the
// rest is skipped for a word none of whose pixels can still be a corner.
__device__ __forceinline__ unsigned fast_gt(unsigned v, unsigned h, unsigned s) { return (v & ~h) | (~(v ^ h) & s); }

__device__ __forceinline__ unsigned any9(const unsigned (&m)[16]) {         // bit 7 of byte j: pixel j has 9 contiguous flags
    unsigned a3[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a3[i] = m[i] & m[(i + 1) & 15] & m[(i + 2) & 15];
    unsigned any = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) any |= a3[i] & a3[(i + 3) & 15] & a3[(i + 6) & 15];
    return any;
}

// window of ring offset DX over the words L | M | R of one tile row (M holds the four centres' columns)
template <int DX>
__device__ __forceinline__ unsigned ring_window(unsigned L, unsigned M, unsigned R) {
    if (DX == 0) return M;
    if (DX < 0) return __byte_perm(L, M, DX == -1 ? 0x6543 : (DX == -2 ? 0x5432 : 0x4321));
    return __byte_perm(M, R, DX == 1 ? 0x4321 : (DX == 2 ? 0x5432 : 0x6543));
}

// cv2 response of a pixel known to be a corner of polarity `pol` (1 bright, 2 dark):
// max over the 16 arcs of 9 contiguous ring pixels of min |ring - c| in that polarity, minus 1.
__device__ __forceinline__ int fast_score(const uint8_t* p, int pol) {
    const int c = p[0];
    int v[16], dd[16], m3[16];
    fast_ring(p, v);
#pragma unroll
    for (int i = 0; i < 16; ++i) dd[i] = pol == 2 ? (c - v[i]) : (v[i] - c);
#pragma unroll
    for (int i = 0; i < 16; ++i) m3[i] = __vimin3_s32(dd[i], dd[(i + 1) & 15], dd[(i + 2) & 15]);
    int best = 0;
#pragma unroll
    for (int i = 0; i < 16; i += 2) {
        const int a = __vimin3_s32(m3[i], m3[(i + 3) & 15], m3[(i + 6) & 15]);
        const int b = __vimin3_s32(m3[i + 1], m3[(i + 4) & 15], m3[(i + 7) & 15]);
        best = __vimax3_s32(best, a, b);
    }
    return best - 1;                // best > thr for a corner, so this is >= thr
}

#define FG 9                        // 8-pixel groups per tile row: tile columns 12 .. 83 cover px = -1 .. 64


struct FastDivs {
    FastDiv per_img, tiles_x, gw, gh;
};

// Persistent: a CTA walks over 64 x TH tiles (tile t, t + grid, ...), the TMA copy of the next tile in flight while this
// one is processed.  Passes per tile: (A) the corner test on every pixel of the tile + 1-pixel ring, four pixels per
// word, corners compacted into a shared list (one shared atomic per warp and round); (B) the exact score only for listed
// corners, one per thread (no lane idles through someone else's score); (C1) strict 3x3 non-maximum suppression, the
// keypoints compacted again; (C2) bucketing by grid cell, every lane busy.
template <int FAST_NT, int FT_H>
__global__ void __launch_bounds__(FAST_NT, FAST_NT == 64 ? 12 : 3) k_fast(const __grid_constant__ CUtensorMap map0, const __grid_constant__ Geom g,
                                                                       const __grid_constant__ DevState d, const __grid_constant__ FastDivs dv,
                                                                       int tiles_x, int tiles_y, int n_tiles) {
    constexpr int FB_H = FT_H + 8;                                  // TMA box height
    constexpr int FB_HP = (FB_H + 3) & ~3;                          // rows per shared stage: FB_W * FB_HP is a multiple of 128 bytes
    constexpr int FAST_POS = (FT_H + 2) * (FT_W + 2), FAST_ITEMS = FG * (FT_H + 2);
    constexpr int FAST_MAXK = (FT_W / 2) * (FT_H / 2);              // strict 3x3 NMS: no two keypoints are 8-neighbours
    __shared__ __align__(128) uint8_t tile[2][FB_HP][FB_W];        // FB_HP: every stage starts on a 128-byte boundary (TMA)
    __shared__ __align__(16) uint8_t sc[2][FT_H + 2][SC_PITCH];
    __shared__ unsigned short clist[FAST_POS];
    __shared__ unsigned klist[FAST_MAXK];
    __shared__ int ccount[2], kcount[2];
    __shared__ __align__(8) uint64_t bar[2];

    const int tid = threadIdx.x, lane = tid & 31;
    const int per_img = tiles_x * tiles_y;
    auto place = [&](int t, int& s, int& bx, int& by) {
        s = fdiv(t, dv.per_img);
        const int r = t - s * per_img;
        by = fdiv(r, dv.tiles_x);
        bx = r - by * tiles_x;
    };
    auto fetch = [&](int t, int st) {                               // thread 0 only
        int s, bx, by;
        place(t, s, bx, by);
        f_mbar_expect_tx(&bar[st], FB_W * FB_H);
        f_tma_load_3d(&tile[st][0][0], &map0, &bar[st], FT_W * bx - FB_X, FT_H * by - 4, s * 2);   // image = s*2 + cam0
    };

    if (tid == 0) {
        f_mbar_init(&bar[0], 1);
        f_mbar_init(&bar[1], 1);
        ccount[0] = ccount[1] = kcount[0] = kcount[1] = 0;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    for (int i = tid; i < 2 * (FT_H + 2) * SC_PITCH / 4; i += FAST_NT) reinterpret_cast<unsigned*>(&sc[0][0][0])[i] = 0u;
    __syncthreads();
    if (tid == 0 && (int)blockIdx.x < n_tiles) fetch(blockIdx.x, 0);

    const unsigned thr4 = (unsigned)g.fast_thr * 0x01010101u;
    int it = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
        const int st = it & 1;
        // every thread is done with the previous tile (its score pass read tile[st ^ 1]): refill that stage now
        if (tid == 0 && t + (int)gridDim.x < n_tiles) fetch(t + gridDim.x, st ^ 1);
        int s, bx, by;
        place(t, s, bx, by);
        const int X0 = FT_W * bx, Y0 = FT_H * by;
        // positions px (tile-relative column) that may hold a corner: the tile and its 1-pixel ring, 3 <= X <= W - 4
        const int pmin = max(-1, 3 - X0), pmax = min(FT_W, g.W - 4 - X0);
        f_mbar_wait(&bar[st], (it >> 1) & 1);

        for (int base = 0; base < FAST_ITEMS; base += FAST_NT) {        // warp-uniform trip count (the append below scans the warp)
            const int item = base + tid;
            const int row = item / FG, grp = item - row * FG;     // row 0 .. FT_H+1 <-> py = row - 1
            const int Y = Y0 + row - 1;
            const int tx = 12 + 8 * grp;                            // tile column of the first of this thread's 8 pixels
            unsigned cm[2] = {0u, 0u}, cb[2] = {0u, 0u};            // corner flags / "bright" flags, bit 7 of byte j = pixel j
            if (item < FAST_ITEMS && Y >= 3 && Y <= g.H - 4) {
                // rows Y-3 .. Y+3 of the tile, columns tx-4 .. tx+11 as four words each
                unsigned w[7][4];
#pragma unroll
                for (int r = 0; r < 7; ++r) {
                    const uint2 a = *reinterpret_cast<const uint2*>(&tile[st][row + r][tx - 4]);   // centre row = tile row (row + 3)
                    const uint2 b = *reinterpret_cast<const uint2*>(&tile[st][row + r][tx + 4]);
                    w[r][0] = a.x, w[r][1] = a.y, w[r][2] = b.x, w[r][3] = b.y;
                }
#pragma unroll
                for (int h = 0; h < 2; ++h) {                       // the two words of centres
                    const unsigned c4 = w[3][1 + h];
                    const unsigned hi = __vaddus4(c4, thr4), lo = __vsubus4(c4, thr4);
                    const unsigned nhi7 = ~hi & 0x7f7f7f7fu, K = (lo & 0x7f7f7f7fu) + 0x7f7f7f7fu;
                    unsigned B[16], D[16];
#define RING(I, R, DX)                                                                     \
    {                                                                                      \
        const unsigned v = ring_window<DX>(w[R][h], w[R][1 + h], w[R][2 + h]);             \
        const unsigned v7 = v & 0x7f7f7f7fu;                                               \
        B[I] = fast_gt(v, hi, v7 + nhi7);                                                  \
        D[I] = fast_gt(lo, v, K - v7);                                                     \
    }
                    RING(0, 6, 0) RING(4, 3, 3) RING(8, 0, 0) RING(12, 3, -3)
                    const unsigned pre = ((B[0] & B[4]) | (B[4] & B[8]) | (B[8] & B[12]) | (B[12] & B[0]) |
                                          (D[0] & D[4]) | (D[4] & D[8]) | (D[8] & D[12]) | (D[12] & D[0])) & 0x80808080u;
                    if (!pre) continue;
                    RING(1, 6, 1) RING(2, 5, 2) RING(3, 4, 3) RING(5, 2, 3) RING(6, 1, 2) RING(7, 0, 1)
                    RING(9, 0, -1) RING(10, 1, -2) RING(11, 2, -3) RING(13, 4, -3) RING(14, 5, -2) RING(15, 6, -1)
#undef RING
                    const unsigned bb = any9(B) & 0x80808080u, dk = any9(D) & 0x80808080u;
                    unsigned m = bb | dk;
                    const int pb = tx + 4 * h - FB_X;               // px of byte 0
                    if (m && (pb < pmin || pb + 3 > pmax)) {        // words on the rim of the tile / of the image
                        const int jlo = max(pmin - pb, 0), jhi = min(pmax - pb, 3);
                        m = jlo > jhi ? 0u : (m & (0xffffffffu << (8 * jlo)) & (0xffffffffu >> (8 * (3 - jhi))));
                    }
                    cm[h] = m;
                    cb[h] = bb;
                }
            }
            // append this warp's corners: one ballot per pixel slot (no data-dependent loop), one shared atomic per warp
            // and round
            if (__any_sync(0xffffffffu, (cm[0] | cm[1]) != 0u)) {
                unsigned bal[8];
                int total = 0;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    bal[k] = __ballot_sync(0xffffffffu, (cm[k >> 2] >> (8 * (k & 3) + 7)) & 1u);
                    total += __popc(bal[k]);
                }
                int wbase = 0;
                if (lane == 0) wbase = atomicAdd(&ccount[st], total);
                wbase = __shfl_sync(0xffffffffu, wbase, 0);
                const unsigned lt = (1u << lane) - 1u;
                const int e0 = row * (FT_W + 2) + tx - FB_X + 1;    // list entry of pixel slot 0
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    if ((bal[k] >> lane) & 1u) {
                        const int pol = (cb[k >> 2] >> (8 * (k & 3) + 7)) & 1u ? 1 : 2;
                        clist[wbase + __popc(bal[k] & lt)] = (unsigned short)((e0 + k) | (pol << 14));
                    }
                    wbase += __popc(bal[k]);
                }
            }
        }
        __syncthreads();
        const int nc = ccount[st];
        for (int j = tid; j < nc; j += FAST_NT) {
            const int e = clist[j], i = e & 0x3fff, pol = e >> 14;
            const int py = i / (FT_W + 2) - 1, px = i % (FT_W + 2) - 1;
            sc[st][py + 1][px + 1] = (uint8_t)fast_score(&tile[st][py + 4][px + FB_X], pol);
        }
        // the other score map (last read by the previous tile's NMS, before the barrier above) is cleared for the next tile
        for (int i = tid; i < (FT_H + 2) * SC_PITCH / 4; i += FAST_NT) reinterpret_cast<unsigned*>(&sc[st ^ 1][0][0])[i] = 0u;
        if (tid == 0) ccount[st ^ 1] = kcount[st ^ 1] = 0;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy reads of the tile before its refill
        __syncthreads();

        // (C1) strict 3x3 NMS at the listed corners that lie inside the tile proper; keypoints compacted
        for (int jb = 0; jb < nc; jb += FAST_NT) {
            const int j = jb + tid;
            bool kp = false;
            int py = 0, px = 0, v = 0;
            if (j < nc) {
                const int i = clist[j] & 0x3fff;
                py = i / (FT_W + 2) - 1, px = i % (FT_W + 2) - 1;
                if (py >= 0 && py < FT_H && px >= 0 && px < FT_W) {
                    const uint8_t(*q)[SC_PITCH] = sc[st];
                    v = q[py + 1][px + 1];
                    kp = v > q[py][px] && v > q[py][px + 1] && v > q[py][px + 2] && v > q[py + 1][px] &&
                         v > q[py + 1][px + 2] && v > q[py + 2][px] && v > q[py + 2][px + 1] && v > q[py + 2][px + 2];
                }
            }
            const unsigned bal = __ballot_sync(0xffffffffu, kp);
            if (bal) {
                int wbase = 0;
                if (lane == 0) wbase = atomicAdd(&kcount[st], __popc(bal));
                wbase = __shfl_sync(0xffffffffu, wbase, 0);
                if (kp) klist[wbase + __popc(bal & ((1u << lane) - 1u))] = ((unsigned)py << 16) | ((unsigned)px << 8) | (unsigned)v;
            }
        }
        __syncthreads();
        // (C2) bucket by grid cell
        const int nk = kcount[st];
        for (int j = tid; j < nk; j += FAST_NT) {
            const unsigned e = klist[j];
            const int X = X0 + (int)((e >> 8) & 0xffu), Y = Y0 + (int)(e >> 16), v = (int)(e & 0xffu);
            const int cell = fdiv(Y, dv.gh) * g.cols + fdiv(X, dv.gw);
            const int pos = atomicAdd(&d.kp_count[s * g.NC + cell], 1);
            if (pos < g.KPC) d.kp_key[((size_t)s * g.NC + cell) * g.KPC + pos] = kp_make_key(v, X, Y, g.W);
        }
        __syncthreads();                // clist, klist, sc[st] and tile[st ^ 1]'s refill slot are free for the next tile
    }
}

void avb_fast_box(int variant, int* w, int* h) {      // shape of the TMA box the fast0 descriptors of a variant need
    *w = FB_W;
    *h = (variant ? 16 : 26) + 8;
}

void launch_fast(const Geom& g, const DevState& d, const PyrMaps& maps, int parity, cudaStream_t st) {
    const int tx = (g.W + FT_W - 1) / FT_W;
    // many streams (every SM gets a dozen small tiles' worth): 64 threads x 26 rows; else 256 threads x 16 rows
    const int v = (long)tx * ((g.H + 25) / 26) * g.S >= 148 * 12 ? 0 : 1;
    const int th = v ? 16 : 26, ty = (g.H + th - 1) / th, nt = tx * ty * g.S;
    FastDivs dv;
    dv.per_img = make_fdiv(tx * ty, nt);
    dv.tiles_x = make_fdiv(tx, tx * ty);
    dv.gw = make_fdiv(g.gw, g.W);
    dv.gh = make_fdiv(g.gh, g.H);
    if (v)
        k_fast<256, 16><<<std::min(nt, 148 * 3), 256, 0, st>>>(maps.fast0[parity][v], g, d, dv, tx, ty, nt);
    else
        k_fast<64, 26><<<std::min(nt, 148 * 12), 64, 0, st>>>(maps.fast0[parity][v], g, d, dv, tx, ty, nt);
}

// Lazy module loading (the CUDA 12 default) loads a kernel on its first launch: ~0.2 ms each, which frame 0 of a stream
// would pay for the kernels only it uses.  cudaFuncGetAttributes loads the function now (called from avb_create).
int avb_preload_fast() {
    cudaFuncAttributes a;
    cudaError_t e = cudaSuccess;
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, k_fast<64, 26>);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, k_fast<256, 16>);
    return e == cudaSuccess ? 0 : -1;
}
