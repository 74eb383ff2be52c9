// Minimal TMA probe: which tensor-map shapes load correctly on this box?  usage: tma_probe rank boxw boxh x0 y0
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
template <int RANK>
__global__ void k(const __grid_constant__ CUtensorMap map, uint8_t* out, int bw, int bh, int x0, int y0, int z) {
    extern __shared__ __align__(128) uint8_t tile[];
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(&bar)), "r"(1));
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(&bar)), "r"(bw * bh) : "memory");
        if (RANK == 3)
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"((unsigned)__cvta_generic_to_shared(tile)), "l"(&map), "r"((unsigned)__cvta_generic_to_shared(&bar)), "r"(x0), "r"(y0), "r"(z) : "memory");
        else
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"((unsigned)__cvta_generic_to_shared(tile)), "l"(&map), "r"((unsigned)__cvta_generic_to_shared(&bar)), "r"(x0), "r"(y0) : "memory");
    }
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"((unsigned)__cvta_generic_to_shared(&bar)), "r"(0) : "memory");
    for (int i = threadIdx.x; i < bw * bh; i += blockDim.x) out[i] = tile[i];
}
int main(int argc, char** argv) {
    int rank = atoi(argv[1]), bw = atoi(argv[2]), bh = atoi(argv[3]), x0 = atoi(argv[4]), y0 = atoi(argv[5]);
    const int W = 752, H = 480, N = 2;
    std::vector<uint8_t> h((size_t)W * H * N);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (uint8_t)((i * 7 + i / W) & 0xff);
    uint8_t *d, *o;
    cudaMalloc(&d, h.size()); cudaMalloc(&o, 65536);
    cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    CUtensorMap m;
    cuuint64_t dims[3] = {W, H, N}; cuuint64_t str[2] = {W, (cuuint64_t)W * H};
    cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, 1}; cuuint32_t es[3] = {1, 1, 1};
    CUresult r = ((EncodeTiledFn)fn)(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, rank, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("rank %d box %dx%d at (%d,%d): encode=%d ", rank, bw, bh, x0, y0, (int)r);
    if (r) { printf("\n"); return 1; }
    if (rank == 3) k<3><<<1, 128, bw * bh>>>(m, o, bw, bh, x0, y0, 1); else k<2><<<1, 128, bw * bh>>>(m, o, bw, bh, x0, y0, 0);
    cudaError_t e = cudaDeviceSynchronize();
    printf("run=%s ", cudaGetErrorString(e));
    if (e == cudaSuccess) {
        std::vector<uint8_t> g(bw * bh); cudaMemcpy(g.data(), o, g.size(), cudaMemcpyDeviceToHost);
        int bad = 0; size_t zoff = rank == 3 ? (size_t)W * H : 0;
        for (int y = 0; y < bh; ++y) for (int x = 0; x < bw; ++x) {
            int X = x0 + x, Y = y0 + y; uint8_t exp = (X < 0 || X >= W || Y < 0 || Y >= H) ? 0 : h[zoff + (size_t)Y * W + X];
            bad += g[y * bw + x] != exp;
        }
        printf("mismatches=%d", bad);
    }
    printf("\n");
    return 0;
}
