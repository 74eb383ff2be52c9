// K1: pyramid build.  One launch per level: dst(x,y) = (sum_{i,j} k_i k_j src(2x+i, 2y+j) + 128) >> 8,
// k = [1 4 6 4 1], BORDER_REFLECT_101, dst size ((w+1)/2, (h+1)/2)   (cv2.pyrDown as
// calcOpticalFlowPyrLK uses it; reference call sites feature_tracker.py:102, stereo_matcher.py:64,70;
// the reference's own PyramidBuilder is a no-op, pyramid_builder.py:30-48).
//
// Each CTA produces a 64x32 tile of the destination level.  The (2*64+4) x (2*32+4) source
// footprint is staged into shared memory by ONE TMA tensor copy (cp.async.bulk.tensor.3d, box
// 160 x 68 x 1 over the (x, y, image) view of the pyramid arena; out-of-image elements arrive as
// zeros and are never read because taps are reflected first).  The TMA start column must be a
// multiple of 16 bytes (measured on B200: any other inner coordinate raises "illegal instruction"),
// so the box starts 16 columns left of the tile instead of 2.  Horizontal 5-tap pass -> u16
// shared buffer -> vertical pass; every thread emits 16 output pixels with one 128-bit store.
#include "avb_common.cuh"

#define PT_W 64
#define PT_H 32
#define PB_W 160                    // box width: 16 left halo (alignment) + 2*PT_W + 16, multiple of 16 bytes
#define PB_X 16                     // columns between the box start and the tile's first source column
#define PB_H 68                     // box height (2*PT_H + 4)
#define HB_PITCH 72                 // u16 pitch of the horizontal-pass buffer (bank spread)

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

__global__ void __launch_bounds__(128) k_pyr_down(const __grid_constant__ CUtensorMap src_map, Geom g, DevState d,
                                                  int level /*dst*/, int parity) {
    __shared__ __align__(128) uint8_t tile[PB_H][PB_W];
    __shared__ __align__(16) unsigned short hbuf[PB_H][HB_PITCH];
    __shared__ __align__(8) uint64_t bar;

    const int tid = threadIdx.x;
    const int s = blockIdx.z >> 1, cam = blockIdx.z & 1;
    const int slot = SLOT(cam, parity);
    const int img = (level == 1) ? (s * 2 + cam) : (s * SLOTS_PER_STREAM + slot);
    const LevelGeom ls = g.lv[level - 1], ld = g.lv[level];
    const int x0 = 2 * PT_W * blockIdx.x - PB_X, y0 = 2 * PT_H * blockIdx.y - 2;

    if (tid == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(&bar, PB_W * PB_H);
        tma_load_3d(&tile[0][0], &src_map, &bar, x0, y0, img);
    }
    mbar_wait(&bar, 0);

    // horizontal pass at the even source columns of this tile
    const int dx0 = PT_W * blockIdx.x, dy0 = PT_H * blockIdx.y;
    for (int i = tid; i < PB_H * PT_W; i += 128) {
        const int r = i / PT_W, x = i % PT_W;
        const int sx = 2 * (dx0 + x);
        int v = 0;
        if (dx0 + x < ld.w) {
            const uint8_t* row = tile[r];
            const int c0 = refl101(sx - 2, ls.w) - x0, c1 = refl101(sx - 1, ls.w) - x0, c2 = sx - x0;
            const int c3 = refl101(sx + 1, ls.w) - x0, c4 = refl101(sx + 2, ls.w) - x0;
            v = row[c0] + 4 * row[c1] + 6 * row[c2] + 4 * row[c3] + row[c4];
        }
        hbuf[r][x] = (unsigned short)v;
    }
    __syncthreads();

    // vertical pass: thread -> 16 consecutive destination pixels of one row
    const int ry = tid >> 2, cx = (tid & 3) * 16;
    const int dy = dy0 + ry, dx = dx0 + cx;
    if (dy < ld.h && dx < ld.pitch) {
        const int sy = 2 * dy;
        const int r0 = refl101(sy - 2, ls.h) - y0, r1 = refl101(sy - 1, ls.h) - y0, r2 = sy - y0;
        const int r3 = refl101(sy + 1, ls.h) - y0, r4 = refl101(sy + 2, ls.h) - y0;
        unsigned out[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            unsigned w = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int x = cx + q * 4 + k;
                const int v = hbuf[r0][x] + 4 * hbuf[r1][x] + 6 * hbuf[r2][x] + 4 * hbuf[r3][x] + hbuf[r4][x];
                w |= (unsigned)((v + 128) >> 8) << (8 * k);
            }
            out[q] = w;
        }
        uint8_t* dst = pyr_slot(d, g, s, slot) + ld.off + (size_t)dy * ld.pitch + dx;
        *reinterpret_cast<uint4*>(dst) = make_uint4(out[0], out[1], out[2], out[3]);
    }
}

void launch_pyramid(const Geom& g, const DevState& d, const PyrMaps& maps, int parity, cudaStream_t st) {
    for (int l = 1; l < g.nlev; ++l) {
        dim3 grid((g.lv[l].w + PT_W - 1) / PT_W, (g.lv[l].h + PT_H - 1) / PT_H, 2 * g.S);
        k_pyr_down<<<grid, 128, 0, st>>>(l == 1 ? maps.l0[parity] : maps.lv[l - 1], g, d, l, parity);
    }
}
