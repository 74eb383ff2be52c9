// HBM-resident sequence store + the per-frame gather that feeds a multi-stream context from it.
//
// The reference sweeps every sequence over 7 time offsets (run.bat:4-12), one process per run: each run re-reads and
// re-decodes the same PNGs from its own start time (main.py:12-13, streaming/dataset.py:206-214).  On a B200 a whole
// EuRoC sequence fits in HBM hundreds of times over (MH_01: 3682 stereo frames x 722 kB = 2.7 GB of 180 GB), so a
// sequence is decoded and uploaded ONCE, and the S lock-stepped offset runs of a context read their current frames
// straight from the store: per step one gather kernel (2S image pointers as kernel parameters -> the context's input
// block) replaces S host copies + H2D transfers.
#include "avb_common.cuh"

#include <cstdio>
#include <cstring>
#include <new>

struct avb_store {
    int device = 0, W = 0, H = 0, n = 0;
    size_t image_bytes = 0;     // W*H padded to 256 B
    uint8_t* d = nullptr;       // [n][2][image_bytes]
    uint8_t* h_stage = nullptr; // pinned bounce block for one stereo frame (strided / pageable input)
};

extern "C" int avb_store_create(int device, int width, int height, int n_frames, avb_store** out) {
    if (!out || width <= 0 || height <= 0 || n_frames <= 0 || (width & 15)) return AVB_E_INVALID;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return AVB_E_NO_DEVICE;
    }
    if (device < 0 || device >= ndev) return AVB_E_INVALID;
    if (cudaSetDevice(device) != cudaSuccess) return AVB_E_CUDA;
    avb_store* s = new (std::nothrow) avb_store;
    if (!s) return AVB_E_CUDA;
    s->device = device;
    s->W = width;
    s->H = height;
    s->n = n_frames;
    s->image_bytes = (((size_t)width * height) + 255) & ~(size_t)255;
    if (cudaMalloc(&s->d, (size_t)n_frames * 2 * s->image_bytes) != cudaSuccess ||
        cudaMallocHost(&s->h_stage, 2 * s->image_bytes) != cudaSuccess) {
        cudaGetLastError();
        if (s->d) cudaFree(s->d);
        delete s;
        return AVB_E_CUDA;
    }
    *out = s;
    return AVB_OK;
}

extern "C" void avb_store_destroy(avb_store* s) {
    if (!s) return;
    cudaSetDevice(s->device);
    if (s->d) cudaFree(s->d);
    if (s->h_stage) cudaFreeHost(s->h_stage);
    delete s;
}

extern "C" int avb_store_num_frames(const avb_store* s) { return s ? s->n : 0; }
extern "C" size_t avb_store_bytes(const avb_store* s) { return s ? (size_t)s->n * 2 * s->image_bytes : 0; }

extern "C" int avb_store_upload(avb_store* s, int k, const uint8_t* img0, const uint8_t* img1, int stride) {
    if (!s || !img0 || !img1 || k < 0 || k >= s->n || stride < s->W) return AVB_E_INVALID;
    if (cudaSetDevice(s->device) != cudaSuccess) return AVB_E_CUDA;
    for (int cam = 0; cam < 2; ++cam) {
        const uint8_t* src = cam ? img1 : img0;
        uint8_t* dst = s->h_stage + cam * s->image_bytes;
        for (int y = 0; y < s->H; ++y) memcpy(dst + (size_t)y * s->W, src + (size_t)y * stride, s->W);
    }
    // the two images of a frame are adjacent in the store: one transfer
    if (cudaMemcpy(s->d + (size_t)k * 2 * s->image_bytes, s->h_stage, 2 * s->image_bytes, cudaMemcpyHostToDevice) != cudaSuccess) {
        cudaGetLastError();
        return AVB_E_CUDA;
    }
    return AVB_OK;
}

extern "C" const uint8_t* avb_store_image(const avb_store* s, int k, int cam) {
    if (!s || k < 0 || k >= s->n || cam < 0 || cam > 1) return nullptr;
    return s->d + ((size_t)k * 2 + cam) * s->image_bytes;
}

// ---- gather: 2S device-resident images -> the [stream][cam] image section of an input block -------------------------
// HBM-bound copy: reads W*H and writes W*H per image with 128-bit accesses, grid sized to fill the SMs.
#define AVB_GATHER_MAX 256
struct GatherTab {
    const uint4* src[AVB_GATHER_MAX];
};

__global__ void __launch_bounds__(256) k_gather_frames(const __grid_constant__ GatherTab tab, uint8_t* dst_base, size_t ib,
                                                       int first) {
    const uint4* __restrict__ src = tab.src[blockIdx.y];
    uint4* __restrict__ dst = reinterpret_cast<uint4*>(dst_base + (size_t)(first + blockIdx.y) * ib);
    const size_t n = ib / sizeof(uint4);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        dst[i] = __ldg(src + i);
}

// images[i] = device pointer of image i (i = stream*2 + cam), each W*H dense bytes, 16-byte aligned.
cudaError_t launch_gather_frames(const Geom& g, uint8_t* d_in, const uint8_t* const* images, cudaStream_t st) {
    const size_t ib = (size_t)g.W * g.H;
    const int total = 2 * g.S;
    const int vec = (int)(ib / sizeof(uint4));
    for (int first = 0; first < total; first += AVB_GATHER_MAX) {
        const int cnt = total - first < AVB_GATHER_MAX ? total - first : AVB_GATHER_MAX;
        GatherTab tab;
        for (int i = 0; i < cnt; ++i) tab.src[i] = reinterpret_cast<const uint4*>(images[first + i]);
        for (int i = cnt; i < AVB_GATHER_MAX; ++i) tab.src[i] = nullptr;
        int gx = (148 * 8 + cnt - 1) / cnt;                  // ~8 CTAs per SM over the whole launch
        const int gx_max = (vec + 256 * 4 - 1) / (256 * 4);  // at least 4 vectors per thread
        if (gx > gx_max) gx = gx_max;
        if (gx < 1) gx = 1;
        k_gather_frames<<<dim3(gx, cnt), 256, 0, st>>>(tab, d_in, ib, first);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}
