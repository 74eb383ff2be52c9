// K1: pyramid build.  dst(x,y) = (sum_{i,j} k_i k_j src(2x+i, 2y+j) + 128) >> 8, k = [1 4 6 4 1],
// BORDER_REFLECT_101, dst size ((w+1)/2, (h+1)/2)   (cv2.pyrDown as calcOpticalFlowPyrLK uses it; reference call
// sites feature_tracker.py:102, stereo_matcher.py:64,70; the reference's own PyramidBuilder is a no-op,
// pyramid_builder.py:30-48).
//
// k_pyr_down   one level, persistent: a CTA walks over 64x32 tiles of the destination level (tile t, t + grid, ...).  The
//              (2*64+4) x (2*32+4) source footprint of a tile is staged into shared memory by ONE TMA tensor copy
//              (cp.async.bulk.tensor.3d, box 160 x 68 x 1 over the (x, y, image) view; out-of-image elements arrive as
//              zeros and are never read because taps are reflected first), double-buffered: the copy of the NEXT tile is
//              in flight while this one is filtered.  The TMA start column must be a multiple of 16 bytes (measured on
//              B200: any other inner coordinate raises "illegal instruction"), so the box starts 16 columns left of the
//              tile instead of 2.  Horizontal 5-tap pass -> u16 shared buffer -> vertical pass -> the finished 64x32
//              tile in shared memory -> ONE TMA tensor store (cp.async.bulk.tensor.3d.global.shared::cta, clipped at the
//              level's width and height by the descriptor), also double-buffered.
// k_pyr_pair   the LAST TWO levels in one launch (they are tiny and launch-latency bound): a CTA owns a 16x8 tile of
//              level l+1 and the 32x16 block of level l under it; it recomputes the 2-pixel halo of level l it needs
//              from a 96 x 44 TMA box of level l-1.
#include <algorithm>

#include "avb_common.cuh"

#define PT_W 64
// rows of a destination tile: template parameter TH of k_pyr_down, 16 or 32.  The per-tile overhead (barriers, fences, TMA
// issue: ~90 of ~150 instructions per warp at TH = 16) is paid once per 64 x 32 pixels at TH = 32, which is what a launch
// over many streams wants; a few streams want many small tiles (latency).
#define PB_W 160                    // box width: 16 left halo (alignment) + 2*PT_W + 16, multiple of 16 bytes
#define PB_X 16                     // columns between the box start and the tile's first source column
#define HB_PITCH 68                 // u16 pitch of the horizontal-pass buffer: rows 8-byte aligned for 64-bit accesses

#define QT_W 16                     // pair kernel: tile of level l+1
#define QT_H 8
#define QM_W 36                     // level-l region: 2*QT_W + 4
#define QM_H 20
#define QB_W 96                     // box of level l-1: 10 left halo (16-aligned start) + 2*QM_W + 4 = 86 -> 96
#define QB_X 10
#define QB_H 44                     // 2*QM_H + 4

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
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
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int x, int y, int z) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
            (unsigned)__cvta_generic_to_shared(smem_dst)),
        "l"(map), "r"((unsigned)__cvta_generic_to_shared(bar)), "r"(x), "r"(y), "r"(z)
        : "memory");
}

__device__ __forceinline__ int dp4a_uu(unsigned a, unsigned b, int c) {      // sum of 4 unsigned byte products + c
    int d;
    asm("dp4a.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, const void* smem_src, int x, int y, int z) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(map),
                 "r"((unsigned)__cvta_generic_to_shared(smem_src)), "r"(x), "r"(y), "r"(z)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}

// Persistent grid = min(tiles, 148 x CTAs per SM).  In flight per SM: 6 CTAs x 4 boxes x 5.6 KB (TH = 16) or 4 x 3 x 10.9 KB
// (TH = 32) = 130-138 KB, what the bandwidth-latency product of HBM asks for (~66 KB per SM at 6.5 TB/s x 1.5 us) with
// headroom.  Measured at 64 streams: the pipeline depth does not matter beyond two boxes (the kernel is issue-bound).
template <int TH>
__global__ void __launch_bounds__(256) k_pyr_down(const __grid_constant__ CUtensorMap src_map, const __grid_constant__ CUtensorMap dst_map,
                                                  const __grid_constant__ Geom g, int level /*dst*/, int parity, int tiles_x,
                                                  int tiles_y, int n_tiles) {
    constexpr int PT_H = TH, PB_H = 2 * TH + 4;
    constexpr int NS = TH == 16 ? 4 : 3;                           // source boxes in flight per CTA (5.6 / 10.9 KB each)
    __shared__ __align__(128) uint8_t tile[NS][PB_H][PB_W];
    __shared__ __align__(128) uint8_t otile[2][PT_H][PT_W];
    __shared__ __align__(16) unsigned short hbuf[PB_H][HB_PITCH];
    __shared__ __align__(8) uint64_t bar[NS];

    const int tid = threadIdx.x;
    const LevelGeom ls = g.lv[level - 1], ld = g.lv[level];

    // This CTA's tiles: a contiguous range [t0, t1) of the (image, tile row, tile column) order, so that the position of
    // the next tile follows from the current one by increments (no division per tile).
    const int t0 = (int)(((long long)n_tiles * blockIdx.x) / gridDim.x), t1 = (int)(((long long)n_tiles * (blockIdx.x + 1)) / gridDim.x);
    struct Pos {
        int bx, by, im;
    };
    auto first = [&](int t) {
        Pos p;
        const int per_img = tiles_x * tiles_y;
        p.im = t / per_img;
        const int r = t - p.im * per_img;
        p.by = r / tiles_x;
        p.bx = r - p.by * tiles_x;
        return p;
    };
    auto advance = [&](Pos& p) {
        if (++p.bx == tiles_x) {
            p.bx = 0;
            if (++p.by == tiles_y) {
                p.by = 0;
                ++p.im;
            }
        }
    };
    auto images = [&](const Pos& p, int& img_src, int& img_dst) {
        const int s = p.im >> 1, cam = p.im & 1;
        img_dst = s * SLOTS_PER_STREAM + SLOT(cam, parity);
        img_src = (level == 1) ? (s * 2 + cam) : img_dst;
    };
    auto fetch = [&](const Pos& p, int st) {                        // thread 0 only
        int is, id;
        images(p, is, id);
        mbar_expect_tx(&bar[st], PB_W * PB_H);
        tma_load_3d(&tile[st][0][0], &src_map, &bar[st], 2 * PT_W * p.bx - PB_X, 2 * PT_H * p.by - 2, is);
    };

    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < NS; ++i) mbar_init(&bar[i], 1);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    pdl_wait();
    pdl_launch_dependents();
    __syncthreads();
    Pos cur = first(t0), ahead = cur;                               // `ahead`: the next tile to fetch (thread 0)
    int t_ahead = t0;
    if (tid == 0) {
        for (int i = 0; i < NS - 1 && t_ahead < t1; ++i, ++t_ahead) {
            fetch(ahead, i);
            advance(ahead);
        }
    }

    // thread roles, fixed over the tiles: horizontal pass -> 4 adjacent outputs of box row hr (+16, +32 in later rounds),
    // vertical pass -> 4 adjacent destination pixels of tile rows ry, ry + 16, ..
    const int hr = tid >> 4, hx = (tid & 15) * 4;
    const int ry = tid >> 4, cx = (tid & 15) * 4;
    int it = 0;
    for (int t = t0; t < t1; ++t, ++it, advance(cur)) {
        const int st = it % NS;
        // the stage refilled now was last read by the horizontal pass of the previous tile, which every thread left
        // through a barrier: the copy of tile t + NS - 1 runs under the arithmetic of this and the next tiles
        if (tid == 0 && t_ahead < t1) {
            fetch(ahead, (it + NS - 1) % NS);
            advance(ahead);
            ++t_ahead;
        }
        const int bx = cur.bx, by = cur.by;
        int img_src, img_dst;
        images(cur, img_src, img_dst);
        const int x0 = 2 * PT_W * bx - PB_X, y0 = 2 * PT_H * by - 2;
        mbar_wait(&bar[st], (it / NS) & 1);

        // BORDER_REFLECT_101 in x: the two columns left of column 0 / right of column w-1 that the 5 taps can touch arrive
        // as zeros from the TMA copy; tiles on the image border write the mirrored pixels there, so that the horizontal
        // pass below needs no per-tap index arithmetic anywhere.
        const int dx0 = PT_W * bx, dy0 = PT_H * by;
        const bool left = bx == 0, right = 2 * (dx0 + PT_W - 1) + 2 >= ls.w;
        if (left || right) {                                        // CTA-uniform
            for (int i = tid; i < PB_H * 4; i += 256) {
                const int r = i >> 2, j = i & 3;                    // j: 0,1 left columns -2,-1; 2,3 right columns w, w+1
                uint8_t* row = tile[st][r];
                if (j < 2) {
                    if (left) row[PB_X - 2 + j] = row[PB_X + 2 - j];
                } else if (right) {
                    const int c = ls.w + (j - 2);                   // source column to fabricate: w or w+1 -> w-2 or w-3
                    if (c - x0 < PB_W) row[c - x0] = row[2 * ls.w - 2 - c - x0];
                }
            }
            __syncthreads();
        }

        // horizontal pass: a thread makes 4 adjacent outputs of one box row from 16 source bytes (aligned 32-bit shared
        // loads, the [1 4 6 4 1] taps as byte dot products): out_k = sum_i w_i b[2k - 2 + i]
        const bool clip_x = dx0 + PT_W > ld.w;                      // CTA-uniform: outputs beyond the level's width exist
#pragma unroll
        for (int rr = 0; rr < PB_H; rr += 16) {
            const int r = rr + hr;
            if (rr + 16 > PB_H && r >= PB_H) break;
            const unsigned* q = reinterpret_cast<const unsigned*>(&tile[st][r][2 * hx + PB_X - 4]);  // bytes c-4 .. c+11, c = 2x + PB_X
            const unsigned w0 = q[0], w1 = q[1], w2 = q[2], w3 = q[3];
            int o0 = dp4a_uu(w0, 0x04010000u, dp4a_uu(w1, 0x00010406u, 0));      // b[c-2] + 4 b[c-1] | 6 b[c] + 4 b[c+1] + b[c+2]
            int o1 = dp4a_uu(w1, 0x04060401u, (int)(w2 & 0xffu));                // b[c] .. b[c+3] | b[c+4]
            int o2 = dp4a_uu(w1, 0x04010000u, dp4a_uu(w2, 0x00010406u, 0));
            int o3 = dp4a_uu(w2, 0x04060401u, (int)(w3 & 0xffu));
            if (clip_x) {                                           // outputs at and beyond the level's width stay zero
                const int lim = ld.w - (dx0 + hx);
                o0 = lim > 0 ? o0 : 0;
                o1 = lim > 1 ? o1 : 0;
                o2 = lim > 2 ? o2 : 0;
                o3 = lim > 3 ? o3 : 0;
            }
            *reinterpret_cast<uint2*>(&hbuf[r][hx]) = make_uint2((unsigned)o0 | ((unsigned)o1 << 16), (unsigned)o2 | ((unsigned)o3 << 16));
        }
        // the tensor store issued from this output buffer two tiles ago must have read its source before it is overwritten
        if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        __syncthreads();

        // vertical pass: thread -> 4 consecutive destination pixels of one row, two per register: a column sum is at most
        // 16 * 16 * 255 + 128 < 2^16, so the packed halves never carry into each other
        const bool inner_y = by > 0 && 2 * (dy0 + PT_H - 1) + 2 < ls.h;   // CTA-uniform: no tap leaves the level vertically
#pragma unroll
        for (int vr = 0; vr < PT_H; vr += 16) {
            const int oy = vr + ry, dy = dy0 + oy;
            if (dy >= ld.h) break;                                  // rows below the level are clipped by the tensor store
            int r0, r1, r2, r3, r4;
            if (inner_y) {
                r2 = 2 * oy + 2;
                r0 = r2 - 2, r1 = r2 - 1, r3 = r2 + 1, r4 = r2 + 2;
            } else {
                const int sy = 2 * dy;
                r0 = refl101(sy - 2, ls.h) - y0, r1 = refl101(sy - 1, ls.h) - y0, r2 = sy - y0;
                r3 = refl101(sy + 1, ls.h) - y0, r4 = refl101(sy + 2, ls.h) - y0;
            }
            const uint2 a0 = *reinterpret_cast<const uint2*>(&hbuf[r0][cx]), a1 = *reinterpret_cast<const uint2*>(&hbuf[r1][cx]);
            const uint2 a2 = *reinterpret_cast<const uint2*>(&hbuf[r2][cx]), a3 = *reinterpret_cast<const uint2*>(&hbuf[r3][cx]);
            const uint2 a4 = *reinterpret_cast<const uint2*>(&hbuf[r4][cx]);
            const unsigned lo = ((a0.x + a4.x + 4u * (a1.x + a3.x) + 6u * a2.x + 0x00800080u) >> 8) & 0x00ff00ffu;
            const unsigned hi = ((a0.y + a4.y + 4u * (a1.y + a3.y) + 6u * a2.y + 0x00800080u) >> 8) & 0x00ff00ffu;
            *reinterpret_cast<unsigned*>(&otile[it & 1][oy][cx]) = __byte_perm(lo, hi, 0x6420);     // bytes lo.0, lo.2, hi.0, hi.2
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the tile, written through the generic proxy, for the TMA
        __syncthreads();                                            // ... and hbuf is free for the next tile
        if (tid == 0) tma_store_3d(&dst_map, &otile[it & 1][0][0], dx0, dy0, img_dst);
    }
    if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// levels `level` and `level + 1` from level `level - 1`
__global__ void __launch_bounds__(256) k_pyr_pair(const __grid_constant__ CUtensorMap src_map, const __grid_constant__ Geom g,
                                                  const __grid_constant__ DevState d, int level, int parity) {
    __shared__ __align__(128) uint8_t tile[QB_H][QB_W];
    __shared__ unsigned short hbuf[QB_H][QM_W + 2];
    __shared__ uint8_t mid[QM_H][QM_W + 4];
    __shared__ unsigned short hbuf2[QM_H][QT_W + 2];
    __shared__ __align__(8) uint64_t bar;

    const int tid = threadIdx.x;
    const int s = blockIdx.z >> 1, cam = blockIdx.z & 1;
    const int slot = SLOT(cam, parity);
    const int img = (level == 1) ? (s * 2 + cam) : (s * SLOTS_PER_STREAM + slot);
    const LevelGeom l0 = g.lv[level - 1], l1 = g.lv[level], l2 = g.lv[level + 1];
    const int X2 = QT_W * blockIdx.x, Y2 = QT_H * blockIdx.y;       // tile origin in level+1
    const int X1 = 2 * X2 - 2, Y1 = 2 * Y2 - 2;                     // region origin in level
    const int x0 = 2 * X1 - 2 - QB_X, y0 = 2 * Y1 - 2;              // box origin in level-1 (x0 is a multiple of 16)

    if (tid == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    pdl_wait();
    pdl_launch_dependents();
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(&bar, QB_W * QB_H);
        tma_load_3d(&tile[0][0], &src_map, &bar, x0, y0, img);
    }
    mbar_wait(&bar, 0);

    // level: horizontal pass over the box rows, at region columns X1 .. X1+QM_W-1 that exist in the level
    for (int i = tid; i < QB_H * QM_W; i += 256) {
        const int r = i / QM_W, x = i % QM_W;
        const int X = X1 + x;
        int v = 0;
        if (X >= 0 && X < l1.w) {
            const uint8_t* row = tile[r];
            const int sx = 2 * X;
            const int c0 = refl101(sx - 2, l0.w) - x0, c1 = refl101(sx - 1, l0.w) - x0, c2 = sx - x0;
            const int c3 = refl101(sx + 1, l0.w) - x0, c4 = refl101(sx + 2, l0.w) - x0;
            v = row[c0] + 4 * row[c1] + 6 * row[c2] + 4 * row[c3] + row[c4];
        }
        hbuf[r][x] = (unsigned short)v;
    }
    __syncthreads();
    // level: vertical pass -> mid (the region of `level` this CTA needs); its inner 32x16 block goes to memory
    uint8_t* out1 = pyr_slot(d, g, s, slot) + l1.off;
    for (int i = tid; i < QM_H * QM_W; i += 256) {
        const int r = i / QM_W, x = i % QM_W;
        const int X = X1 + x, Y = Y1 + r;
        int v = 0;
        if (X >= 0 && X < l1.w && Y >= 0 && Y < l1.h) {
            const int sy = 2 * Y;
            const int r0 = refl101(sy - 2, l0.h) - y0, r1 = refl101(sy - 1, l0.h) - y0, r2 = sy - y0;
            const int r3 = refl101(sy + 1, l0.h) - y0, r4 = refl101(sy + 2, l0.h) - y0;
            v = (hbuf[r0][x] + 4 * hbuf[r1][x] + 6 * hbuf[r2][x] + 4 * hbuf[r3][x] + hbuf[r4][x] + 128) >> 8;
            if (x >= 2 && x < QM_W - 2 && r >= 2 && r < QM_H - 2) out1[(size_t)Y * l1.pitch + X] = (uint8_t)v;
        }
        mid[r][x] = (uint8_t)v;
    }
    __syncthreads();
    // level+1: horizontal pass over the region rows
    for (int i = tid; i < QM_H * QT_W; i += 256) {
        const int r = i / QT_W, x = i % QT_W;
        const int X = X2 + x;
        int v = 0;
        if (X < l2.w) {
            const uint8_t* row = mid[r];
            const int sx = 2 * X;
            const int c0 = refl101(sx - 2, l1.w) - X1, c1 = refl101(sx - 1, l1.w) - X1, c2 = sx - X1;
            const int c3 = refl101(sx + 1, l1.w) - X1, c4 = refl101(sx + 2, l1.w) - X1;
            v = row[c0] + 4 * row[c1] + 6 * row[c2] + 4 * row[c3] + row[c4];
        }
        hbuf2[r][x] = (unsigned short)v;
    }
    __syncthreads();
    if (tid < QT_W * QT_H) {
        const int r = tid / QT_W, x = tid % QT_W;
        const int X = X2 + x, Y = Y2 + r;
        if (X < l2.w && Y < l2.h) {
            const int sy = 2 * Y;
            const int r0 = refl101(sy - 2, l1.h) - Y1, r1 = refl101(sy - 1, l1.h) - Y1, r2 = sy - Y1;
            const int r3 = refl101(sy + 1, l1.h) - Y1, r4 = refl101(sy + 2, l1.h) - Y1;
            const int v = hbuf2[r0][x] + 4 * hbuf2[r1][x] + 6 * hbuf2[r2][x] + 4 * hbuf2[r3][x] + hbuf2[r4][x];
            (pyr_slot(d, g, s, slot) + l2.off)[(size_t)Y * l2.pitch + X] = (uint8_t)((v + 128) >> 8);
        }
    }
}

// The pair kernel trades redundant halo work for one launch less: right for a few streams (launch-latency bound), wrong
// when the one-level kernel already fills the GPU at the pair's first level (many streams: it ran at 6 % of the HBM
// rate there).
int avb_pyramid_pair_level(const Geom& g) {
    const int built = g.nlev - 1;                       // levels 1..built
    if (built < 2) return 0;
    const int l = built - 1;
    const long ctas = (long)((g.lv[l].w + PT_W - 1) / PT_W) * ((g.lv[l].h + 15) / 16) * 2 * g.S;
    return ctas >= 148 * 4 ? 0 : l;
}

void avb_pyramid_boxes(int variant, int* src_w, int* src_h, int* dst_w, int* dst_h) {   // shapes the k_pyr_down descriptors need
    const int th = variant ? 32 : 16;
    *src_w = PB_W, *src_h = 2 * th + 4, *dst_w = PT_W, *dst_h = th;
}

// tile height per level: 32 rows once 16-row tiles alone would give every SM eight CTAs' worth of tiles
static int pyr_tile_variant(const Geom& g, int l) {
    const long t16 = (long)((g.lv[l].w + PT_W - 1) / PT_W) * ((g.lv[l].h + 15) / 16) * 2 * g.S;
    return t16 >= 148 * 8 ? 1 : 0;
}

void launch_pyramid(const Geom& g, const DevState& d, const PyrMaps& maps, int parity, cudaStream_t st) {
    const int built = g.nlev - 1;                       // levels 1..built
    const int pair_at = g.pyr_pair_level;               // the pair kernel builds levels pair_at and pair_at + 1 (0: not used)
    for (int l = 1; l <= built; ++l) {
        if (l == pair_at) {
            dim3 grid((g.lv[l + 1].w + QT_W - 1) / QT_W, (g.lv[l + 1].h + QT_H - 1) / QT_H, 2 * g.S);
            launch_k(k_pyr_pair, grid, dim3(256), 0, st, g_avb_pdl && l > 1, l == 1 ? maps.pair0[parity] : maps.pair, g, d, l, parity);
            break;
        }
        const int v = pyr_tile_variant(g, l), th = v ? 32 : 16;
        const int tx = (g.lv[l].w + PT_W - 1) / PT_W, ty = (g.lv[l].h + th - 1) / th, nt = tx * ty * 2 * g.S;
        const CUtensorMap& src = l == 1 ? maps.l0[parity][v] : maps.lv[l - 1][v];
        if (v)
            launch_k(k_pyr_down<32>, dim3(std::min(nt, 148 * 4)), dim3(256), 0, st, g_avb_pdl && l > 1, src, maps.dst[l][v], g, l, parity,
                     tx, ty, nt);
        else
            launch_k(k_pyr_down<16>, dim3(std::min(nt, 148 * 6)), dim3(256), 0, st, g_avb_pdl && l > 1, src, maps.dst[l][v], g, l, parity,
                     tx, ty, nt);
    }
}

int avb_pyramid_launches(const Geom& g) {
    const int built = g.nlev - 1;
    return g.pyr_pair_level ? built - 1 : built;
}

// Lazy module loading (the CUDA 12 default) loads a kernel on its first launch: ~0.2 ms each, which frame 0 of a stream
// would pay for the kernels only it uses.  cudaFuncGetAttributes loads the function now (called from avb_create).
int avb_preload_pyramid() {
    cudaFuncAttributes a;
    cudaError_t e = cudaSuccess;
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, k_pyr_down<16>);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, k_pyr_down<32>);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, k_pyr_pair);
    return e == cudaSuccess ? 0 : -1;
}
