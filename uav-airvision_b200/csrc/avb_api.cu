// C-ABI of libavb (include/avb.h): context, memory, CUDA-graph frame path, per-stage entry points.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "avb_common.cuh"

static thread_local std::string g_create_error;
int g_avb_pdl = 1;

// Host helper thread for pageable caller images: the two images of a stereo frame are staged into pinned memory by two
// cores at once (a 361 kB memcpy from cold memory takes 25-35 us).  Polls for about a millisecond after a job, then
// sleeps on a condition variable; the poster spins on an atomic for the few microseconds it may have to wait.
struct CopyWorker {
    std::thread th;
    std::mutex m;
    std::condition_variable cv;
    const uint8_t* src = nullptr;
    uint8_t* dst = nullptr;
    int W = 0, H = 0, stride = 0;
    std::atomic<int> state{0};      // 0 idle, 1 job posted, 2 job done
    bool quit = false;

    static void copy(uint8_t* dst, const uint8_t* src, int W, int H, int stride) {
        if (stride == W)
            memcpy(dst, src, (size_t)W * H);
        else
            for (int y = 0; y < H; ++y) memcpy(dst + (size_t)y * W, src + (size_t)y * stride, W);
    }
    void run() {
        for (;;) {
            // a stream of frames posts every few hundred microseconds: poll that long before going to sleep (waking a
            // sleeping thread costs more than the copy it is there to hide)
            bool have = false;
            for (int spin = 0; spin < 40000 && !have; ++spin) {
                have = state.load(std::memory_order_acquire) == 1;
#if defined(__x86_64__)
                if (!have) __builtin_ia32_pause();
#endif
            }
            if (!have) {
                std::unique_lock<std::mutex> lk(m);
                cv.wait(lk, [this] { return quit || state.load(std::memory_order_acquire) == 1; });
                if (quit) return;
            }
            copy(dst, src, W, H, stride);
            state.store(2, std::memory_order_release);
        }
    }
    void post(uint8_t* d, const uint8_t* s, int w, int h, int st) {
        if (!th.joinable()) th = std::thread([this] { run(); });
        if (state.load(std::memory_order_acquire) != 0) wait();        // a job left behind by an error return
        {
            std::lock_guard<std::mutex> lk(m);
            dst = d, src = s, W = w, H = h, stride = st;
            state.store(1, std::memory_order_release);
        }
        cv.notify_one();
    }
    void wait() {
        while (state.load(std::memory_order_acquire) != 2) {
#if defined(__x86_64__)
            __builtin_ia32_pause();
#endif
        }
        state.store(0, std::memory_order_relaxed);
    }
    ~CopyWorker() {
        if (th.joinable()) {
            {
                std::lock_guard<std::mutex> lk(m);
                quit = true;
            }
            cv.notify_one();
            th.join();
        }
    }
};

struct avb_ctx {
    avb_config cfg;
    Geom g;
    DevState d;
    PyrMaps maps;
    cudaStream_t st = nullptr, st_side = nullptr, st_rot = nullptr;   // st_rot: the rotation section's copy (avb_process_submitted)
    int submitted = 0;              // avb_submit_images done, avb_process_submitted due: 1 split-graph path, 2 general path
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_pyr = nullptr, ev_t0 = nullptr, ev_t1 = nullptr;
    uint8_t* h_in = nullptr;        // pinned: one input block (images + H)
    uint8_t* h_out[2] = {nullptr, nullptr};   // pinned: S result blocks per frame parity (the results of frame k stay
                                    // readable while frame k+1 runs: a driver launches k+1 first and reads k behind it)
    uint8_t* d_out[2] = {nullptr, nullptr};   // what k_finish of a parity writes: the mapped host block (zc_out) or the device mirror
    bool zc_out = false;            // k_finish writes the result blocks straight into h_out (mapped): no D2H copy node
    size_t out_stride = 0;
    int parity = 1;                 // parity of the current frame; the first frame lands in parity 0
    bool first_frame = true;
    bool have_frame = false;
    bool rot_copy_pending = false;  // an enqueued (not waited-for) gather frame's rotation copy may not have run yet
    cudaGraphExec_t graph[2] = {nullptr, nullptr};   // steady-state frame, per parity, host-input variant
    cudaGraphExec_t graph_dev[2] = {nullptr, nullptr}; // same, device-input variant (no H2D of images)
    // host-image path (avb_process_frame): the cam0-only work (FAST, speculative list) as a graph of its own, launched
    // as soon as the cam0 images have arrived, while the cam1 images are still on the bus; and the rest of the chain
    cudaGraphExec_t graph_cam0[2] = {nullptr, nullptr}, graph_rest[2] = {nullptr, nullptr};
    cudaEvent_t ev_cam0 = nullptr, ev_rot = nullptr, ev_side_done = nullptr;
    std::vector<void*> allocs;
    // scratch for the per-stage entry points
    float2 *s_a = nullptr, *s_b = nullptr, *s_c = nullptr;
    uint8_t* s_st = nullptr;
    double *s_da = nullptr, *s_db = nullptr, *s_R = nullptr;
    int s_cap = 0;
    std::string err;
    float last_ms = 0.f;
    static constexpr int PIN_CACHE = 8;   // is_page_locked: recently seen caller buffers
    uintptr_t pin_base[PIN_CACHE] = {};
    size_t pin_len[PIN_CACHE] = {};
    bool pin_yes[PIN_CACHE] = {};
    unsigned pin_next = 0;
    CopyWorker copier;
};

static int fail(avb_ctx* c, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (c) c->err = buf; else g_create_error = buf;
    return code;
}

#define CK(call)                                                                                      \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess) return fail(c, AVB_E_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

template <typename T>
static cudaError_t dalloc(avb_ctx* c, T** p, size_t count, bool zero = true) {
    void* v = nullptr;
    cudaError_t e = cudaMalloc(&v, std::max<size_t>(count * sizeof(T), 256));
    if (e != cudaSuccess) return e;
    c->allocs.push_back(v);
    if (zero) e = cudaMemset(v, 0, std::max<size_t>(count * sizeof(T), 256));
    *p = static_cast<T*>(v);
    return e;
}

extern "C" int avb_abi_version(void) { return AVB_ABI_VERSION; }

extern "C" const char* avb_last_error(const avb_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int make_map(avb_ctx* c, EncodeTiledFn enc, CUtensorMap* m, void* base, int w, int h, int nimg, size_t pitch,
                    size_t img_stride, int bw, int bh) {
    cuuint64_t dims[3] = {(cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)nimg};
    cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)img_stride};
    cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(c, AVB_E_CUDA, "cuTensorMapEncodeTiled failed (%d) for %dx%dx%d pitch %zu", (int)r, w, h, nimg, pitch);
    return AVB_OK;
}

static int build_graphs(avb_ctx* c);

extern "C" int avb_create(const avb_config* cfg, avb_ctx** out) {
    avb_ctx* c = nullptr;
    if (!cfg || !out) return fail(c, AVB_E_INVALID, "null argument");
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(c, AVB_E_NO_DEVICE, "no CUDA device: libavb has no CPU fallback");
    if (cfg->device < 0 || cfg->device >= ndev) return fail(c, AVB_E_INVALID, "device %d out of range", cfg->device);
    if (cfg->win_size != AVB_WIN) return fail(c, AVB_E_INVALID, "only patch_size 15 is compiled (got %d)", cfg->win_size);
    if (cfg->width % 16 || cfg->width < 64 || cfg->height < 64 || cfg->width > 4096 || cfg->height > 4095)
        return fail(c, AVB_E_INVALID, "width must be a multiple of 16, 64 <= size <= 4096 (got %dx%d)", cfg->width, cfg->height);
    if (cfg->max_level < 0 || cfg->max_level >= AVB_MAX_LEVELS) return fail(c, AVB_E_INVALID, "pyramid_levels must be 0..%d", AVB_MAX_LEVELS - 1);
    if ((cfg->width >> cfg->max_level) < 32 || (cfg->height >> cfg->max_level) < 32)
        return fail(c, AVB_E_INVALID, "coarsest pyramid level would be smaller than 32 px");
    if (cfg->grid_row < 1 || cfg->grid_col < 1 || cfg->grid_row * cfg->grid_col > AVB_MAX_CELLS)
        return fail(c, AVB_E_INVALID, "grid must have 1..%d cells", AVB_MAX_CELLS);
    if (cfg->grid_max_feature_num < 1 || cfg->grid_max_feature_num > AVB_MAX_CAP || cfg->grid_min_feature_num < 0 ||
        cfg->grid_min_feature_num > cfg->grid_max_feature_num)
        return fail(c, AVB_E_INVALID, "need 0 <= grid_min <= grid_max <= %d", AVB_MAX_CAP);
    if (cfg->num_streams < 1 || cfg->num_streams > 4096) return fail(c, AVB_E_INVALID, "num_streams must be 1..4096");
    if (cfg->fast_threshold < 1 || cfg->fast_threshold > 254) return fail(c, AVB_E_INVALID, "fast_threshold must be 1..254");
    if (cfg->ransac != 0 && cfg->ransac != 1) return fail(c, AVB_E_INVALID, "ransac must be 0 or 1");
    if (cfg->ransac && !(cfg->ransac_threshold > 0.0)) return fail(c, AVB_E_INVALID, "ransac_threshold must be positive");
    if ((size_t)cfg->width * cfg->height >= (1u << 24)) return fail(c, AVB_E_INVALID, "image too large for the 24-bit scan index");

    c = new avb_ctx();
    c->cfg = *cfg;
    cudaError_t e = cudaSetDevice(cfg->device);
    if (e != cudaSuccess) {
        fail(nullptr, AVB_E_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
        delete c;
        return AVB_E_CUDA;
    }

    Geom& g = c->g;
    memset(&g, 0, sizeof g);
    g.W = cfg->width;
    g.H = cfg->height;
    g.nlev = cfg->max_level + 1;
    g.S = cfg->num_streams;
    size_t off = 0;
    for (int l = 0; l < g.nlev; ++l) {
        g.lv[l].w = l ? (g.lv[l - 1].w + 1) / 2 : g.W;
        g.lv[l].h = l ? (g.lv[l - 1].h + 1) / 2 : g.H;
        g.lv[l].pitch = l ? ((g.lv[l].w + 15) & ~15) : g.W;
        g.lv[l].off = l ? off : 0;
        if (l) off += ((size_t)g.lv[l].pitch * g.lv[l].h + 255) & ~(size_t)255;
    }
    g.slot_bytes = std::max<size_t>(off, 256);
    g.rows = cfg->grid_row;
    g.cols = cfg->grid_col;
    g.NC = g.rows * g.cols;
    g.gh = (g.H + g.rows - 1) / g.rows;       // int(np.ceil(h / grid_row))  (B12)
    g.gw = (g.W + g.cols - 1) / g.cols;
    g.gmin = cfg->grid_min_feature_num;
    g.gmax = cfg->grid_max_feature_num;
    g.NMAX = g.NC * g.gmax;
    // strict 3x3 NMS: no two keypoints are 8-neighbours -> at most ceil(gw/2)*ceil(gh/2) per cell
    g.KPC = (((g.gw + 1) / 2) * ((g.gh + 1) / 2) + 3) & ~3;
    g.fast_thr = cfg->fast_threshold;
    // LK lane mapping: few features in flight (one or a handful of streams) -> 4 warps per feature to shorten the
    // dependent chain; many -> 1 warp per feature for throughput.  AVB_WPF=1|4 overrides (experiments).
    // Both LK kernels are latency-bound per feature, so what counts is the number of WAVES: the 4-warp mapping keeps
    // 148 SMs x 4 CTAs = 592 teams resident (~120 registers), the 1-warp mapping 148 x 20 = 2960 (96 registers).  Measured at C3 (2000
    // slots, one stream): k_track 228 us with 4 warps (3.4 waves) vs 150 us with 1 warp (one wave).
    g.wpf = ((long long)g.S * g.NMAX <= 592) ? 4 : 1;
    if (const char* e = getenv("AVB_WPF")) g.wpf = (atoi(e) == 4) ? 4 : 1;
    if (const char* e = getenv("AVB_PDL")) g_avb_pdl = atoi(e) ? 1 : 0;
    g.max_iter = std::min(std::max(cfg->max_iteration, 0), 100);
    g.min_eig = cfg->min_eig_threshold;
    const double eps = std::min(std::max(cfg->track_precision, 0.0), 10.0);
    g.eps2 = eps * eps;
    g.eps2_lo = (float)(g.eps2 * 0.99999);      // formed here: in the kernels the compiler re-did this FP64 product every iteration
    g.eps2_hi = (float)(g.eps2 * 1.00001);
    // a threshold <= 0 never takes the shortcut (the float minEig can round below zero)
    g.eig_accept = g.min_eig > 0.0 ? (float)(2.0 * AVB_WIN * AVB_WIN * (2.0 * g.min_eig + 4e-5)) : INFINITY;
    g.cam0 = {cfg->cam0_intrinsics[0], cfg->cam0_intrinsics[1], cfg->cam0_intrinsics[2], cfg->cam0_intrinsics[3],
              cfg->cam0_distortion[0], cfg->cam0_distortion[1], cfg->cam0_distortion[2], cfg->cam0_distortion[3]};
    g.cam1 = {cfg->cam1_intrinsics[0], cfg->cam1_intrinsics[1], cfg->cam1_intrinsics[2], cfg->cam1_intrinsics[3],
              cfg->cam1_distortion[0], cfg->cam1_distortion[1], cfg->cam1_distortion[2], cfg->cam1_distortion[3]};
    memcpy(g.R01, cfg->R_cam0_to_cam1, sizeof g.R01);
    memcpy(g.E, cfg->essential, sizeof g.E);
    g.epi_thr = cfg->stereo_threshold * (4.0 / (2 * g.cam0.fx + 2 * g.cam0.fy));
    g.ransac = cfg->ransac;
    g.ransac_seed = cfg->ransac_seed;
    g.ransac_iters = (int)ceil(log(1.0 - 0.99) / log(1.0 - 0.7 * 0.7));     // success probability 0.99, inlier ratio 0.7
    g.ransac_thr = cfg->ransac_threshold;
    // Launch-shape choices that depend on how many streams share a launch; the environment overrides exist so that the
    // parity tests can force the many-stream shapes (per-level pyramid kernel, two candidate rounds) on a small context.
    g.pyr_pair_level = avb_pyramid_pair_level(g);
    if (const char* e = getenv("AVB_PYR_PAIR")) g.pyr_pair_level = atoi(e) ? (g.nlev - 1 >= 2 ? g.nlev - 2 : 0) : 0;
    g.cand_rounds = avb_candidate_rounds(g);
    if (const char* e = getenv("AVB_CAND_ROUNDS")) g.cand_rounds = (atoi(e) == 2 && g.wpf == 1 && g.gmin < g.gmax) ? 2 : 1;
    // Few streams (the 4-warp, latency-bound mapping): the candidates' stereo matches run speculatively beside k_track.
    // 16 list entries per cell cover gmax candidates unless the mask removes more than 16 - gmax stronger keypoints.
    g.spec_k = g.wpf == 4 ? std::min(std::max(16, g.gmax), std::min(g.KPC, 32)) : 0;
    if (const char* e = getenv("AVB_SPEC_K")) g.spec_k = g.wpf == 4 ? std::min(std::max(atoi(e), 0), std::min(g.KPC, 32)) : 0;
    g.spec_wpf = 1;
    if (const char* e = getenv("AVB_SPEC_WPF")) g.spec_wpf = atoi(e) == 4 ? 4 : 1;
    if (g.NMAX > 8192) {
        delete c;
        return fail(nullptr, AVB_E_INVALID, "grid_num*grid_max = %d exceeds 8192", g.NMAX);
    }

#define CKC(call)                                                                                   \
    do {                                                                                            \
        cudaError_t e2_ = (call);                                                                   \
        if (e2_ != cudaSuccess) {                                                                   \
            fail(nullptr, AVB_E_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e2_), __FILE__, __LINE__); \
            avb_destroy(c);                                                                         \
            return AVB_E_CUDA;                                                                      \
        }                                                                                           \
    } while (0)

    DevState& d = c->d;
    memset(&d, 0, sizeof d);
    const size_t S = g.S, NM = g.NMAX, NC = g.NC;
    const size_t inb = in_block_bytes(g);
    CKC(cudaStreamCreateWithFlags(&c->st, cudaStreamNonBlocking));
    CKC(cudaStreamCreateWithFlags(&c->st_side, cudaStreamNonBlocking));
    CKC(cudaStreamCreateWithFlags(&c->st_rot, cudaStreamNonBlocking));
    CKC(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    CKC(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
    CKC(cudaEventCreateWithFlags(&c->ev_pyr, cudaEventDisableTiming));
    CKC(cudaEventCreateWithFlags(&c->ev_cam0, cudaEventDisableTiming));
    CKC(cudaEventCreateWithFlags(&c->ev_rot, cudaEventDisableTiming));
    CKC(cudaEventCreateWithFlags(&c->ev_side_done, cudaEventDisableTiming));
    CKC(cudaEventCreate(&c->ev_t0));
    CKC(cudaEventCreate(&c->ev_t1));
    CKC(dalloc(c, &d.in[0], inb));
    CKC(dalloc(c, &d.in[1], inb));
    CKC(dalloc(c, &d.pyr, S * SLOTS_PER_STREAM * g.slot_bytes + 256));   // + slack: LK reads whole aligned words
    CKC(dalloc(c, &d.kp_key, S * NC * g.KPC));
    CKC(dalloc(c, &d.kp_count, S * NC));
    CKC(dalloc(c, &d.kp_p1, S * NC * g.KPC));
    CKC(dalloc(c, &d.kp_ok, S * NC * g.KPC));
    for (int p = 0; p < 2; ++p) {
        CKC(dalloc(c, &d.grid[p].ids, S * NM));
        CKC(dalloc(c, &d.grid[p].life, S * NM));
        CKC(dalloc(c, &d.grid[p].p0, S * NM));
        CKC(dalloc(c, &d.grid[p].p1, S * NM));
        CKC(dalloc(c, &d.grid[p].fresh, S * NM));
        CKC(dalloc(c, &d.grid[p].count, S * NC));
    }
    CKC(dalloc(c, &d.t_p0, S * NM));
    CKC(dalloc(c, &d.t_p1, S * NM));
    CKC(dalloc(c, &d.t_cell, S * NM));
    CKC(dalloc(c, &d.t_und, S * NM));
    CKC(dalloc(c, &d.c_und, S * NM));
    CKC(dalloc(c, &d.c_key, S * NM));
    CKC(dalloc(c, &d.c_src, S * NM));
    CKC(dalloc(c, &d.c_p1, S * NM));
    CKC(dalloc(c, &d.c_ok, S * NM));
    CKC(dalloc(c, &d.c_count, S * NC));
    if (g.spec_k > 0) {
        const size_t SK = S * NC * (size_t)g.spec_k;
        CKC(dalloc(c, &d.s_key, SK));
        CKC(dalloc(c, &d.s_p1, SK));
        CKC(dalloc(c, &d.s_ok, SK));
        CKC(dalloc(c, &d.s_und, SK));
        CKC(dalloc(c, &d.s_n, S * NC));
    }
    CKC(dalloc(c, &d.n_new, S * NC));
    CKC(dalloc(c, &d.next_id, 2 * S));
    CKC(dalloc(c, &d.counters, S * 8));
    CKC(dalloc(c, &d.frame_index, S));
    if (g.ransac) {
        CKC(dalloc(c, &d.r_idx, 2 * S * NM));
        CKC(dalloc(c, &d.r_und, 2 * S * NM));
        CKC(dalloc(c, &d.r_raw, 2 * S * NM));
        CKC(dalloc(c, &d.r_bits, 2 * S * NM));
        CKC(dalloc(c, &d.r_prev, S * NM));
    }
    c->out_stride = out_stride_bytes(g.NMAX);
    // A small result (one or a few streams) is written by k_finish directly into mapped pinned host memory: the posted
    // PCIe writes overlap the kernel's tail and the D2H copy node (launch + completion latency of a DMA for ~19 kB)
    // disappears from the frame.  Many streams: one bulk DMA of the device mirror is cheaper than scattered stores.
    c->zc_out = S * c->out_stride <= (size_t)64 * 1024;
    if (const char* e = getenv("AVB_ZC_OUT")) c->zc_out = atoi(e) != 0;
    for (int q = 0; q < 2; ++q) {
        CKC(cudaHostAlloc((void**)&c->h_out[q], S * c->out_stride, cudaHostAllocMapped));
        memset(c->h_out[q], 0, S * c->out_stride);
        if (c->zc_out) {
            void* dp = nullptr;
            CKC(cudaHostGetDevicePointer(&dp, c->h_out[q], 0));
            c->d_out[q] = static_cast<uint8_t*>(dp);
        } else if (q == 0) {
            CKC(dalloc(c, &c->d_out[0], S * c->out_stride));   // one device mirror: its D2H copy is stream-ordered before the next k_finish
        } else {
            c->d_out[1] = c->d_out[0];
        }
    }
    d.out = c->d_out[0];
    CKC(cudaHostAlloc((void**)&c->h_in, inb, cudaHostAllocDefault));
    memset(c->h_in, 0, inb);
    avb_fill_rotations(c, c->h_in, nullptr, nullptr);      // identity until the caller provides rotations

    // dynamic shared memory of the bookkeeping kernels
    const size_t sel_smem = (size_t)g.KPC * 4;
    if (sel_smem > 200 * 1024) {
        avb_destroy(c);
        return fail(nullptr, AVB_E_INVALID, "grid cell too large for the selection kernel (%zu B shared)", sel_smem);
    }
    if (avb_set_smem_limits(sel_smem, (size_t)g.NMAX * 12) != 0) {
        avb_destroy(c);
        return fail(nullptr, AVB_E_CUDA, "cudaFuncSetAttribute(MaxDynamicSharedMemorySize) failed");
    }

    // TMA descriptors
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CKC(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) {
        avb_destroy(c);
        return fail(nullptr, AVB_E_CUDA, "cuTensorMapEncodeTiled not available from the driver");
    }
    EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(fn);
    const size_t img_bytes = (size_t)g.W * g.H;
    {
        const int built = g.nlev - 1, pair_at = built >= 2 ? built - 1 : 0;
        int r = AVB_OK;
        for (int v = 0; v < 2 && r == AVB_OK; ++v) {              // k_pyr_down: 16-row and 32-row tiles
            int pbw = 0, pbh = 0, ptw = 0, pth = 0;
            avb_pyramid_boxes(v, &pbw, &pbh, &ptw, &pth);
            for (int p = 0; p < 2 && r == AVB_OK; ++p)
                r = make_map(c, enc, &c->maps.l0[p][v], d.in[p], g.W, g.H, g.S * 2, g.W, img_bytes, pbw, pbh);
            for (int l = 1; l < g.nlev - 1 && r == AVB_OK; ++l)
                r = make_map(c, enc, &c->maps.lv[l][v], d.pyr + g.lv[l].off, g.lv[l].w, g.lv[l].h, g.S * SLOTS_PER_STREAM,
                             g.lv[l].pitch, g.slot_bytes, pbw, pbh);
            for (int l = 1; l < g.nlev && r == AVB_OK; ++l)      // store views: clipped at the level's width and height
                r = make_map(c, enc, &c->maps.dst[l][v], d.pyr + g.lv[l].off, g.lv[l].w, g.lv[l].h, g.S * SLOTS_PER_STREAM,
                             g.lv[l].pitch, g.slot_bytes, ptw, pth);
        }
        for (int p = 0; p < 2 && r == AVB_OK; ++p) {
            r = make_map(c, enc, &c->maps.pair0[p], d.in[p], g.W, g.H, g.S * 2, g.W, img_bytes, 96, 44);
            for (int v = 0; v < 2 && r == AVB_OK; ++v) {
                int fbw = 0, fbh = 0;
                avb_fast_box(v, &fbw, &fbh);
                r = make_map(c, enc, &c->maps.fast0[p][v], d.in[p], g.W, g.H, g.S * 2, g.W, img_bytes, fbw, fbh);
            }
        }
        if (pair_at >= 2 && r == AVB_OK) {
            const int l = pair_at - 1;
            r = make_map(c, enc, &c->maps.pair, d.pyr + g.lv[l].off, g.lv[l].w, g.lv[l].h, g.S * SLOTS_PER_STREAM, g.lv[l].pitch,
                         g.slot_bytes, 96, 44);
        }
        if (r != AVB_OK) {
            g_create_error = c->err;
            avb_destroy(c);
            return r;
        }
    }
    CKC(cudaStreamSynchronize(c->st));
    CKC(cudaDeviceSynchronize());
    if (avb_preload_fast() || avb_preload_grid() || avb_preload_points() || avb_preload_pyramid()) {
        avb_destroy(c);
        return fail(nullptr, AVB_E_CUDA, "loading the kernels failed: %s", cudaGetErrorString(cudaGetLastError()));
    }
    if (cfg->use_graph) {
        int r = build_graphs(c);
        if (r != AVB_OK) {
            g_create_error = c->err;
            avb_destroy(c);
            return r;
        }
    }
    *out = c;
    return AVB_OK;
}

extern "C" void avb_destroy(avb_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->cfg.device);
    if (c->st) cudaStreamSynchronize(c->st);
    for (int p = 0; p < 2; ++p) {
        if (c->graph[p]) cudaGraphExecDestroy(c->graph[p]);
        if (c->graph_dev[p]) cudaGraphExecDestroy(c->graph_dev[p]);
        if (c->graph_cam0[p]) cudaGraphExecDestroy(c->graph_cam0[p]);
        if (c->graph_rest[p]) cudaGraphExecDestroy(c->graph_rest[p]);
    }
    for (void* p : c->allocs) cudaFree(p);
    for (void* p : {(void*)c->s_a, (void*)c->s_b, (void*)c->s_c, (void*)c->s_st, (void*)c->s_da, (void*)c->s_db, (void*)c->s_R})
        if (p) cudaFree(p);
    if (c->h_in) cudaFreeHost(c->h_in);
    for (uint8_t* h : c->h_out)
        if (h) cudaFreeHost(h);
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    if (c->ev_join) cudaEventDestroy(c->ev_join);
    if (c->ev_pyr) cudaEventDestroy(c->ev_pyr);
    for (cudaEvent_t e : {c->ev_cam0, c->ev_rot, c->ev_side_done})
        if (e) cudaEventDestroy(e);
    if (c->ev_t0) cudaEventDestroy(c->ev_t0);
    if (c->ev_t1) cudaEventDestroy(c->ev_t1);
    if (c->st) cudaStreamDestroy(c->st);
    if (c->st_side) cudaStreamDestroy(c->st_side);
    if (c->st_rot) cudaStreamDestroy(c->st_rot);
    delete c;
}

extern "C" int avb_capacity(const avb_ctx* c) { return c ? c->g.NMAX : 0; }
extern "C" int avb_num_cells(const avb_ctx* c) { return c ? c->g.NC : 0; }
extern "C" int avb_get_geometry(const avb_ctx* c, int* w, int* h, int* S) {
    if (!c) return AVB_E_INVALID;
    if (w) *w = c->g.W;
    if (h) *h = c->g.H;
    if (S) *S = c->g.S;
    return AVB_OK;
}
extern "C" uint8_t* avb_input_staging(avb_ctx* c) { return c ? c->h_in : nullptr; }
extern "C" void* avb_cuda_stream(avb_ctx* c) { return c ? (void*)c->st : nullptr; }

extern "C" int avb_reset(avb_ctx* c) {
    if (!c) return AVB_E_INVALID;
    CK(cudaSetDevice(c->cfg.device));
    CK(cudaStreamSynchronize(c->st_side));
    CK(cudaStreamSynchronize(c->st_rot));
    CK(cudaStreamSynchronize(c->st));
    c->submitted = 0;                   // a submitted, unfinished frame is dropped with the rest of the state
    const Geom& g = c->g;
    for (int p = 0; p < 2; ++p) CK(cudaMemsetAsync(c->d.grid[p].count, 0, (size_t)g.S * g.NC * sizeof(int), c->st));
    CK(cudaMemsetAsync(c->d.next_id, 0, (size_t)2 * g.S * sizeof(long long), c->st));
    CK(cudaMemsetAsync(c->d.frame_index, 0, (size_t)g.S * sizeof(int), c->st));
    CK(cudaMemsetAsync(c->d.counters, 0, (size_t)g.S * 8 * sizeof(int), c->st));
    CK(cudaMemsetAsync(c->d.kp_count, 0, (size_t)g.S * g.NC * sizeof(int), c->st));
    CK(cudaStreamSynchronize(c->st));
    c->parity = 1;
    c->first_frame = true;
    c->have_frame = false;
    return AVB_OK;
}

// ---- the frame ------------------------------------------------------------------------------------

// The steady-state chain in two parts for the host-image path.  Part 1 needs the cam0 images only.
// DevState as the kernels of parity p see it (k_finish writes that parity's result block)
static DevState dev_state(const avb_ctx* c, int p) {
    DevState d = c->d;
    d.out = c->d_out[p];
    return d;
}

static void enqueue_cam0_part(avb_ctx* c, int p, cudaStream_t st) {
    launch_fast(c->g, c->d, c->maps, p, st);
    if (c->g.spec_k > 0) launch_spec_select(c->g, c->d, st);
}
// Part 2 (captured on c->st): everything else.  It meets part 1 through ev_side_done, an event OUTSIDE the graph (recorded
// on the side stream behind part 1, before this graph is launched): first where the speculative matches start, else
// where k_select reads the FAST buckets.
static void enqueue_rest_part(avb_ctx* c, int p) {
    const Geom& g = c->g;
    const DevState d = dev_state(c, p);
    const bool spec = g.spec_k > 0;
    launch_pyramid(g, d, c->maps, p, c->st);
    cudaEventRecord(c->ev_fork, c->st);
    cudaStreamWaitEvent(c->st_side, c->ev_fork, 0);                            // the side stream joins the capture
    cudaStreamWaitEvent(c->st_side, c->ev_side_done, cudaEventWaitExternal);
    if (spec) launch_spec_match(g, d, p, c->st_side);
    cudaEventRecord(c->ev_join, c->st_side);
    launch_track(g, d, p, c->st);
    if (g.ransac) launch_ransac(g, d, p, c->st);
    cudaStreamWaitEvent(c->st, c->ev_join, 0);
    launch_select(g, d, p, 0, c->st);
    if (!spec) launch_stereo_candidates(g, d, p, c->st);
    launch_finish(g, d, p, 0, c->st);
}

// Enqueues the kernel chain of one frame on c->st (FAST runs on a forked branch).  Inputs of parity p must
// already be (or be ordered before this on c->st) in d.in[p].
static void enqueue_chain(avb_ctx* c, int p, bool first) {
    const Geom& g = c->g;
    const DevState d = dev_state(c, p);
    cudaEventRecord(c->ev_fork, c->st);
    cudaStreamWaitEvent(c->st_side, c->ev_fork, 0);
    launch_fast(g, d, c->maps, p, c->st_side);          // the buckets are empty: k_select left them so (avb_grid.cu)
    const bool spec = g.spec_k > 0 && !first;
    if (spec) launch_spec_select(g, d, c->st_side);
    if (!spec) cudaEventRecord(c->ev_join, c->st_side);
    launch_pyramid(g, d, c->maps, p, c->st);
    if (spec) {                         // the speculative matches read this frame's pyramids of both cameras
        cudaEventRecord(c->ev_pyr, c->st);
        cudaStreamWaitEvent(c->st_side, c->ev_pyr, 0);
        launch_spec_match(g, d, p, c->st_side);
        cudaEventRecord(c->ev_join, c->st_side);
    }
    if (first) {
        cudaStreamWaitEvent(c->st, c->ev_join, 0);
        launch_stereo_buckets(g, d, p, c->st);
        launch_select(g, d, p, 1, c->st);
    } else {
        launch_track(g, d, p, c->st);
        if (g.ransac) launch_ransac(g, d, p, c->st);
        cudaStreamWaitEvent(c->st, c->ev_join, 0);
        launch_select(g, d, p, 0, c->st);
        if (!spec) launch_stereo_candidates(g, d, p, c->st);    // with speculation k_select itself matches what the list misses
    }
    launch_finish(g, d, p, first ? 1 : 0, c->st);
}

// Serialised, instrumented variant of the steady-state chain: one event after every stage on c->st.
// Stage order: 0 input copy | 1 FAST (+ speculative list) | 2 pyramid | 3 track | 4 select | 5 stereo(new) |
// 6 finish (grid update + publish) | 7 speculative stereo matches (beside k_track in the real chain; 0 when off) |
// 8 result copy.  Used by bench.py for the per-kernel roofline; never by the hot path.
extern "C" int avb_profile_frame_device(avb_ctx* c, const uint8_t* d_block, float* stage_ms /*[9]*/) {
    if (!c || !d_block || !stage_ms) return AVB_E_INVALID;
    if (c->first_frame) return fail(c, AVB_E_STATE, "profile the steady state: process frame 0 first");
    CK(cudaSetDevice(c->cfg.device));
    const Geom& g = c->g;
    const int p = c->parity ^ 1;
    const DevState d = dev_state(c, p);
    cudaEvent_t ev[11];
    for (auto& e : ev) CK(cudaEventCreate(&e));
    CK(cudaStreamSynchronize(c->st));
    CK(cudaEventRecord(c->ev_t0, c->st));
    CK(cudaEventRecord(ev[0], c->st));
    CK(cudaMemcpyAsync(c->d.in[p], d_block, in_block_bytes(g), cudaMemcpyDeviceToDevice, c->st));
    CK(cudaEventRecord(ev[1], c->st));
    launch_fast(g, d, c->maps, p, c->st);
    if (g.spec_k > 0) launch_spec_select(g, d, c->st);
    CK(cudaEventRecord(ev[2], c->st));
    launch_pyramid(g, d, c->maps, p, c->st);
    CK(cudaEventRecord(ev[3], c->st));
    if (g.spec_k > 0) launch_spec_match(g, d, p, c->st);
    CK(cudaEventRecord(ev[10], c->st));
    launch_track(g, d, p, c->st);
    if (g.ransac) launch_ransac(g, d, p, c->st);       // counted with the track stage
    CK(cudaEventRecord(ev[4], c->st));
    launch_select(g, d, p, 0, c->st);
    CK(cudaEventRecord(ev[5], c->st));
    if (g.spec_k == 0) launch_stereo_candidates(g, d, p, c->st);
    CK(cudaEventRecord(ev[6], c->st));
    launch_finish(g, d, p, 0, c->st);
    CK(cudaEventRecord(ev[7], c->st));
    CK(cudaEventRecord(ev[8], c->st));
    if (!c->zc_out) CK(cudaMemcpyAsync(c->h_out[p], c->d_out[p], (size_t)g.S * c->out_stride, cudaMemcpyDeviceToHost, c->st));
    CK(cudaEventRecord(ev[9], c->st));
    CK(cudaEventRecord(c->ev_t1, c->st));
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(c->st));
    static const int from[9] = {0, 1, 2, 10, 4, 5, 6, 3, 8}, to[9] = {1, 2, 3, 4, 5, 6, 7, 10, 9};
    for (int i = 0; i < 9; ++i) cudaEventElapsedTime(&stage_ms[i], ev[from[i]], ev[to[i]]);
    for (auto& e : ev) cudaEventDestroy(e);
    c->parity = p;
    return AVB_OK;
}

extern "C" int avb_kernels_per_frame(const avb_ctx* c) {
    if (!c) return 0;
    // fast, pyramid launches, track, [ransac], select, stereo_candidates (two rounds in throughput mode), finish;
    // with speculation: + the speculative list and its matches, - stereo_candidates
    return 1 + avb_pyramid_launches(c->g) + 4 + (c->g.ransac ? 1 : 0) + (c->g.cand_rounds == 2 ? 1 : 0) + (c->g.spec_k > 0 ? 1 : 0);
}

static int build_graphs(avb_ctx* c) {
    const Geom& g = c->g;
    const size_t inb = in_block_bytes(g);
    for (int variant = 0; variant < 2; ++variant) {
        for (int p = 0; p < 2; ++p) {
            cudaGraph_t graph = nullptr;
            CK(cudaStreamBeginCapture(c->st, cudaStreamCaptureModeThreadLocal));
            if (variant == 0)   // device variant: the block was placed by a D2D copy ordered before the launch
                cudaMemcpyAsync(c->d.in[p], c->h_in, inb, cudaMemcpyHostToDevice, c->st);
            enqueue_chain(c, p, false);
            if (!c->zc_out) cudaMemcpyAsync(c->h_out[p], c->d_out[p], (size_t)g.S * c->out_stride, cudaMemcpyDeviceToHost, c->st);
            CK(cudaStreamEndCapture(c->st, &graph));
            cudaGraphExec_t exec = nullptr;
            CK(cudaGraphInstantiate(&exec, graph, 0));
            CK(cudaGraphDestroy(graph));
            (variant == 0 ? c->graph : c->graph_dev)[p] = exec;
        }
    }
    for (int p = 0; p < 2; ++p) {
        cudaGraph_t graph = nullptr;
        cudaGraphExec_t exec = nullptr;
        CK(cudaStreamBeginCapture(c->st_side, cudaStreamCaptureModeThreadLocal));
        enqueue_cam0_part(c, p, c->st_side);
        CK(cudaStreamEndCapture(c->st_side, &graph));
        CK(cudaGraphInstantiate(&exec, graph, 0));
        CK(cudaGraphDestroy(graph));
        c->graph_cam0[p] = exec;
        CK(cudaStreamBeginCapture(c->st, cudaStreamCaptureModeThreadLocal));
        enqueue_rest_part(c, p);
        if (!c->zc_out) cudaMemcpyAsync(c->h_out[p], c->d_out[p], (size_t)g.S * c->out_stride, cudaMemcpyDeviceToHost, c->st);
        CK(cudaStreamEndCapture(c->st, &graph));
        CK(cudaGraphInstantiate(&exec, graph, 0));
        CK(cudaGraphDestroy(graph));
        c->graph_rest[p] = exec;
    }
    return AVB_OK;
}

extern "C" size_t avb_input_block_bytes(const avb_ctx* c) { return c ? in_block_bytes(c->g) : 0; }
extern "C" size_t avb_input_rotation_offset(const avb_ctx* c) { return c ? in_images_bytes(c->g) : 0; }
extern "C" size_t avb_input_rotation_stride(const avb_ctx* c) { return c ? AVB_ROT_DOUBLES * sizeof(double) : 0; }

extern "C" int avb_fill_rotations(const avb_ctx* c, uint8_t* block, const double* R_p_c0, const double* R_p_c1) {
    // H = K R_p_c K^-1 (feature_tracker.py:166-171) in double: (K @ R) @ inv(K), inv(K) in closed form
    if (!c || !block) return AVB_E_INVALID;
    const Geom& g = c->g;
    double* sec = reinterpret_cast<double*>(block + in_images_bytes(g));
    const double fx = g.cam0.fx, fy = g.cam0.fy, cx = g.cam0.cx, cy = g.cam0.cy;
    const double K[9] = {fx, 0, cx, 0, fy, cy, 0, 0, 1};
    const double Ki[9] = {1.0 / fx, 0, -cx / fx, 0, 1.0 / fy, -cy / fy, 0, 0, 1};
    auto mul = [](const double* A, const double* B, double* C) {
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) C[i * 3 + j] = A[i * 3] * B[j] + A[i * 3 + 1] * B[3 + j] + A[i * 3 + 2] * B[6 + j];
    };
    for (int s = 0; s < g.S; ++s) {
        double* h = sec + (size_t)s * AVB_ROT_DOUBLES;
        if (!R_p_c0) {
            for (int i = 0; i < AVB_ROT_DOUBLES; ++i) h[i] = ((i % 9) % 4 == 0) ? 1.0 : 0.0;
            continue;
        }
        const double* R = R_p_c0 + (size_t)s * 9;
        double KR[9];
        mul(K, R, KR);
        mul(KR, Ki, h);
        memcpy(h + 9, R, 9 * sizeof(double));
        if (R_p_c1) {
            memcpy(h + 18, R_p_c1 + (size_t)s * 9, 9 * sizeof(double));
        } else {                        // the same gyro rotation seen from cam1: R01 R0 R01^T
            double T[9], R01t[9];
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j) R01t[i * 3 + j] = g.R01[j * 3 + i];
            mul(g.R01, R, T);
            mul(T, R01t, h + 18);
        }
    }
    return AVB_OK;
}

// variant 0: the whole block lies in the pinned staging (one H2D node inside the graph); 1: the block was placed in
// d.in[p] by copies ordered before this on c->st; 2: like 1, and ev_t0 was already recorded before those copies.
static int run_frame(avb_ctx* c, int variant, bool wait) {
    const Geom& g = c->g;
    if (c->submitted) return fail(c, AVB_E_STATE, "images submitted: avb_process_submitted must finish that frame first");
    const int p = c->parity ^ 1;
    const size_t inb = in_block_bytes(g);
    if (variant != 2) CK(cudaEventRecord(c->ev_t0, c->st));
    if (!c->first_frame && c->cfg.use_graph) {
        CK(cudaGraphLaunch((variant == 0 ? c->graph : c->graph_dev)[p], c->st));
    } else {
        if (variant == 0) CK(cudaMemcpyAsync(c->d.in[p], c->h_in, inb, cudaMemcpyHostToDevice, c->st));
        enqueue_chain(c, p, c->first_frame);
        if (!c->zc_out) CK(cudaMemcpyAsync(c->h_out[p], c->d_out[p], (size_t)g.S * c->out_stride, cudaMemcpyDeviceToHost, c->st));
        CK(cudaGetLastError());
    }
    CK(cudaEventRecord(c->ev_t1, c->st));
    c->parity = p;
    c->first_frame = false;
    c->have_frame = true;
    if (wait) CK(cudaStreamSynchronize(c->st));       // device time of the frame: on demand, avb_last_frame_ms
    return AVB_OK;
}

// Is [p, p + n) page-locked host memory?  The driver query costs about a microsecond per image, and callers keep their
// frames in a few large page-locked allocations (or recycle buffers), so the last few answers are remembered per
// context as address ranges (the whole allocation when the driver reports it).  A stale "yes" (buffer freed and
// reallocated pageable) is harmless: cudaMemcpyAsync accepts pageable memory, it merely stages the copy itself.
typedef CUresult (*GetAddressRangeFn)(CUdeviceptr*, size_t*, CUdeviceptr);
static bool is_page_locked(avb_ctx* c, const void* p, size_t n) {
    const uintptr_t a0 = reinterpret_cast<uintptr_t>(p);
    for (int i = 0; i < avb_ctx::PIN_CACHE; ++i)
        if (c->pin_len[i] && a0 >= c->pin_base[i] && a0 + n <= c->pin_base[i] + c->pin_len[i]) return c->pin_yes[i];
    cudaPointerAttributes a;
    bool yes = false;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess)
        cudaGetLastError();
    else
        yes = a.type == cudaMemoryTypeHost;
    uintptr_t base = a0;
    size_t len = n;
    if (yes) {
        static GetAddressRangeFn range_fn = [] {
            void* fn = nullptr;
            cudaDriverEntryPointQueryResult q;
            if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &q) != cudaSuccess ||
                q != cudaDriverEntryPointSuccess) {
                cudaGetLastError();
                fn = nullptr;
            }
            return reinterpret_cast<GetAddressRangeFn>(fn);
        }();
        CUdeviceptr b = 0;
        size_t sz = 0;
        if (range_fn && range_fn(&b, &sz, (CUdeviceptr)a0) == CUDA_SUCCESS && b <= a0 && a0 + n <= b + sz) {
            base = (uintptr_t)b;
            len = sz;
        }
    }
    const int slot = c->pin_next++ % avb_ctx::PIN_CACHE;
    c->pin_base[slot] = base;
    c->pin_len[slot] = len;
    c->pin_yes[slot] = yes;
    return yes;
}

// Intake of one frame's host images (pipelined: each image is staged into pinned memory -- or, when the caller's buffer
// is itself page-locked and dense, taken from where it lies -- and its H2D copy is enqueued before the next image is
// touched, so the PCIe transfer of one image overlaps the host copy of the next).  The cam0 copy goes out before anything
// else: the GPU is idle until it arrives and every host call ahead of it is on the frame's critical path.  Nothing here
// needs the frame's gyro rotations, so a caller may compute them (IMUProcessor.integrate_imu_data) while the copies
// fly: avb_submit_images, then avb_process_submitted(R).
static int submit_images(avb_ctx* c, const uint8_t* const* img0, const uint8_t* const* img1, int stride) {
    const Geom& g = c->g;
    if (c->submitted) return fail(c, AVB_E_STATE, "images already submitted: call avb_process_submitted first");
    if (!(img0 && img1)) return fail(c, AVB_E_INVALID, "null image arrays");
    if (stride < g.W) return fail(c, AVB_E_INVALID, "stride %d < width %d", stride, g.W);
    const int p = c->parity ^ 1;
    const size_t ib = (size_t)g.W * g.H;
    CK(cudaEventRecord(c->ev_t0, c->st));
    if (g.S == 1 && !c->first_frame && c->cfg.use_graph && c->graph_cam0[p]) {
        // One stream, steady state: the cam0-only part of the chain (FAST, speculative list) starts as soon as the cam0
        // image has arrived, while the cam1 image is still on the bus (7 us at 752x480); the rest of the chain follows
        // the cam1 copy and meets that part through ev_side_done.
        const uint8_t *src0 = img0[0], *src1 = img1[0];
        if (!src0 || !src1) return fail(c, AVB_E_INVALID, "null image pointer (stream 0)");
        const bool pin0 = stride == g.W && is_page_locked(c, src0, ib), pin1 = stride == g.W && is_page_locked(c, src1, ib);
        int rc = AVB_OK;
        auto ck = [&](cudaError_t e, const char* what) {
            if (e != cudaSuccess && rc == AVB_OK) rc = fail(c, AVB_E_CUDA, "%s: %s", what, cudaGetErrorString(e));
        };
        // Pageable (or strided) images are staged into the pinned block by two cores at once, each image split in a top
        // and a bottom half: cam0 is what the GPU waits for first (FAST runs behind its copy), so both cores work on it
        // before either touches cam1 (one core per image put cam0 on the bus 10 us later).
        const int h_top = g.H / 2, h_bot = g.H - h_top;
        auto stage = [&](uint8_t* dst, const uint8_t* src) {
            c->copier.post(dst + (size_t)h_top * g.W, src + (size_t)h_top * stride, g.W, h_bot, stride);
            CopyWorker::copy(dst, src, g.W, h_top, stride);
            c->copier.wait();
        };
        if (pin0) {
            ck(cudaMemcpyAsync(c->d.in[p], src0, ib, cudaMemcpyHostToDevice, c->st), "H2D cam0");
        } else {
            stage(c->h_in, src0);
            ck(cudaMemcpyAsync(c->d.in[p], c->h_in, ib, cudaMemcpyHostToDevice, c->st), "H2D cam0");
        }
        ck(cudaEventRecord(c->ev_cam0, c->st), "event");
        ck(cudaStreamWaitEvent(c->st_side, c->ev_t0, 0), "wait");               // not before the previous frame is done with its tables
        ck(cudaStreamWaitEvent(c->st_side, c->ev_cam0, 0), "wait");
        ck(cudaGraphLaunch(c->graph_cam0[p], c->st_side), "graph (cam0 part)");
        ck(cudaEventRecord(c->ev_side_done, c->st_side), "event");
        if (pin1) {
            ck(cudaMemcpyAsync(c->d.in[p] + ib, src1, ib, cudaMemcpyHostToDevice, c->st), "H2D cam1");
        } else {
            stage(c->h_in + ib, src1);
            ck(cudaMemcpyAsync(c->d.in[p] + ib, c->h_in + ib, ib, cudaMemcpyHostToDevice, c->st), "H2D cam1");
        }
        if (rc != AVB_OK) {
            cudaStreamSynchronize(c->st_side);
            cudaStreamSynchronize(c->st);
            return rc;
        }
        c->submitted = 1;
        return AVB_OK;
    }
    for (int s = 0; s < g.S; ++s) {
        const uint8_t* src0 = img0[s];
        const uint8_t* src1 = img1[s];
        if (!src0 || !src1) {
            cudaStreamSynchronize(c->st);
            return fail(c, AVB_E_INVALID, "null image pointer (stream %d)", s);
        }
        const size_t off0 = ((size_t)s * 2) * ib, off1 = off0 + ib;
        const bool pin0 = stride == g.W && is_page_locked(c, src0, ib), pin1 = stride == g.W && is_page_locked(c, src1, ib);
        if (!pin1) c->copier.post(c->h_in + off1, src1, g.W, g.H, stride);     // the helper stages cam1 meanwhile
        cudaError_t e;
        if (pin0) {
            e = cudaMemcpyAsync(c->d.in[p] + off0, src0, ib, cudaMemcpyHostToDevice, c->st);
        } else {
            CopyWorker::copy(c->h_in + off0, src0, g.W, g.H, stride);
            e = cudaMemcpyAsync(c->d.in[p] + off0, c->h_in + off0, ib, cudaMemcpyHostToDevice, c->st);
        }
        if (!pin1) c->copier.wait();
        if (e == cudaSuccess)
            e = cudaMemcpyAsync(c->d.in[p] + off1, pin1 ? src1 : c->h_in + off1, ib, cudaMemcpyHostToDevice, c->st);
        if (e != cudaSuccess) {
            cudaStreamSynchronize(c->st);
            return fail(c, AVB_E_CUDA, "H2D (stream %d): %s", s, cudaGetErrorString(e));
        }
    }
    c->submitted = 2;
    return AVB_OK;
}

// The rest of a submitted frame: the 216-byte-per-stream rotation section travels on a stream of its own (a copy this
// small is all latency; in line behind the images it would add ~2.5 us to the chain), then the kernels, then the wait.
static int finish_submitted(avb_ctx* c, const double* R_p_c0, const double* R_p_c1) {
    const Geom& g = c->g;
    if (!c->submitted) return fail(c, AVB_E_STATE, "no images submitted");
    const int mode = c->submitted;
    c->submitted = 0;
    const int p = c->parity ^ 1;
    const size_t ro = in_images_bytes(g);
    avb_fill_rotations(c, c->h_in, R_p_c0, R_p_c1);
    cudaError_t e = cudaStreamWaitEvent(c->st_rot, c->ev_t0, 0);               // not before the previous frame is done with d.in[p]
    if (e == cudaSuccess) e = cudaMemcpyAsync(c->d.in[p] + ro, c->h_in + ro, in_block_bytes(g) - ro, cudaMemcpyHostToDevice, c->st_rot);
    if (e == cudaSuccess) e = cudaEventRecord(c->ev_rot, c->st_rot);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(c->st, c->ev_rot, 0);
    if (e == cudaSuccess && mode == 1) {
        e = cudaGraphLaunch(c->graph_rest[p], c->st);
        if (e == cudaSuccess) e = cudaEventRecord(c->ev_t1, c->st);
    }
    if (e != cudaSuccess) {
        cudaStreamSynchronize(c->st_rot);
        cudaStreamSynchronize(c->st_side);
        cudaStreamSynchronize(c->st);
        return fail(c, AVB_E_CUDA, "frame submission: %s", cudaGetErrorString(e));
    }
    if (mode == 1) {
        c->parity = p;
        c->have_frame = true;
        CK(cudaStreamSynchronize(c->st));
        return AVB_OK;
    }
    return run_frame(c, 2, true);
}

extern "C" int avb_submit_images(avb_ctx* c, const uint8_t* const* img0, const uint8_t* const* img1, int stride) {
    if (!c) return AVB_E_INVALID;
    CK(cudaSetDevice(c->cfg.device));
    return submit_images(c, img0, img1, stride);
}

extern "C" int avb_process_submitted(avb_ctx* c, const double* R_p_c0, const double* R_p_c1) {
    if (!c) return AVB_E_INVALID;
    CK(cudaSetDevice(c->cfg.device));
    return finish_submitted(c, R_p_c0, R_p_c1);
}

extern "C" int avb_process_frame(avb_ctx* c, const uint8_t* const* img0, const uint8_t* const* img1, int stride,
                                 const double* R_p_c0, const double* R_p_c1) {
    if (!c) return AVB_E_INVALID;
    CK(cudaSetDevice(c->cfg.device));
    if (!(img0 && img1)) {              // the caller filled the pinned staging block in place
        if (c->submitted) return fail(c, AVB_E_STATE, "images already submitted: call avb_process_submitted first");
        avb_fill_rotations(c, c->h_in, R_p_c0, R_p_c1);
        return run_frame(c, 0, true);
    }
    const int rc = submit_images(c, img0, img1, stride);
    return rc != AVB_OK ? rc : finish_submitted(c, R_p_c0, R_p_c1);
}

extern "C" int avb_enqueue_frame_device(avb_ctx* c, const uint8_t* d_block) {
    if (!c || !d_block) return AVB_E_INVALID;
    CK(cudaSetDevice(c->cfg.device));
    const int p = c->parity ^ 1;
    CK(cudaMemcpyAsync(c->d.in[p], d_block, in_block_bytes(c->g), cudaMemcpyDeviceToDevice, c->st));
    return run_frame(c, 1, false);
}

// Frames that already live in HBM (an avb_store, or any device buffers): one gather kernel places them in the input
// block, the rotation section travels from the pinned staging on the side stream, then the device-input graph runs.
cudaError_t launch_gather_frames(const Geom& g, uint8_t* d_in, const uint8_t* const* images, cudaStream_t st);

static int gather_frame(avb_ctx* c, const uint8_t* const* d_images, const double* R_p_c0, const double* R_p_c1, bool wait) {
    if (!c || !d_images) return AVB_E_INVALID;
    const Geom& g = c->g;
    CK(cudaSetDevice(c->cfg.device));
    for (int i = 0; i < 2 * g.S; ++i)
        if (!d_images[i] || (reinterpret_cast<uintptr_t>(d_images[i]) & 15))
            return fail(c, AVB_E_INVALID, "image %d (stream %d cam %d): null or not 16-byte aligned", i, i / 2, i & 1);
    const int p = c->parity ^ 1;
    // One frame may be in flight while the next is enqueued (avb_enqueue_frame_gather).  The rotation section travels
    // from ONE pinned staging area: before it is rewritten, the copy that carried the previous frame's rotations must
    // have executed (it waits for the frame before that, so this blocks only a caller that runs more than one frame
    // ahead -- and then it is what keeps frame k from reading the rotations of frame k+1).
    if (c->rot_copy_pending) {
        CK(cudaEventSynchronize(c->ev_join));
        c->rot_copy_pending = false;
    }
    CK(cudaEventRecord(c->ev_t0, c->st));
    avb_fill_rotations(c, c->h_in, R_p_c0, R_p_c1);
    const size_t ro = in_images_bytes(g);
    CK(cudaStreamWaitEvent(c->st_side, c->ev_t0, 0));
    CK(cudaMemcpyAsync(c->d.in[p] + ro, c->h_in + ro, in_block_bytes(g) - ro, cudaMemcpyHostToDevice, c->st_side));
    CK(cudaEventRecord(c->ev_join, c->st_side));
    c->rot_copy_pending = !wait;
    CK(launch_gather_frames(g, c->d.in[p], d_images, c->st));
    CK(cudaStreamWaitEvent(c->st, c->ev_join, 0));
    return run_frame(c, 2, wait);
}

extern "C" int avb_process_frame_gather(avb_ctx* c, const uint8_t* const* d_images, const double* R_p_c0,
                                        const double* R_p_c1) {
    return gather_frame(c, d_images, R_p_c0, R_p_c1, true);
}

extern "C" int avb_enqueue_frame_gather(avb_ctx* c, const uint8_t* const* d_images, const double* R_p_c0,
                                        const double* R_p_c1) {
    return gather_frame(c, d_images, R_p_c0, R_p_c1, false);
}

extern "C" int avb_sync(avb_ctx* c) {
    if (!c) return AVB_E_INVALID;
    CK(cudaStreamSynchronize(c->st));
    c->rot_copy_pending = false;
    return AVB_OK;
}

extern "C" int avb_process_frame_device(avb_ctx* c, const uint8_t* d_block) {
    int r = avb_enqueue_frame_device(c, d_block);
    if (r != AVB_OK) return r;
    return avb_sync(c);
}

extern "C" int avb_last_frame_ms(avb_ctx* c, float* ms) {
    if (!c || !ms) return AVB_E_INVALID;
    if (c->have_frame) {
        CK(cudaEventSynchronize(c->ev_t1));
        CK(cudaEventElapsedTime(&c->last_ms, c->ev_t0, c->ev_t1));
    }
    *ms = c->last_ms;
    return AVB_OK;
}

static int result_at(avb_ctx* c, int s, int parity, const avb_frame_header** hdr, const int64_t** ids, const double** meas) {
    if (!c || s < 0 || s >= c->g.S) return AVB_E_INVALID;
    if (!c->have_frame) return fail(c, AVB_E_STATE, "no frame processed yet");
    uint8_t* b = c->h_out[parity] + (size_t)s * c->out_stride;
    if (hdr) *hdr = reinterpret_cast<const avb_frame_header*>(b);
    if (ids) *ids = reinterpret_cast<const int64_t*>(out_ids(b));
    if (meas) *meas = out_meas(b, c->g.NMAX);
    return AVB_OK;
}
extern "C" int avb_get_result(avb_ctx* c, int s, const avb_frame_header** hdr, const int64_t** ids, const double** meas) {
    return result_at(c, s, c ? c->parity : 0, hdr, ids, meas);
}
extern "C" int avb_get_result_prev(avb_ctx* c, int s, const avb_frame_header** hdr, const int64_t** ids, const double** meas) {
    return result_at(c, s, c ? c->parity ^ 1 : 0, hdr, ids, meas);
}

extern "C" int avb_get_features(avb_ctx* c, int s, int32_t* cell, int32_t* lifetime, float* cam0_xy, float* cam1_xy) {
    if (!c || s < 0 || s >= c->g.S) return AVB_E_INVALID;
    if (!c->have_frame) return fail(c, AVB_E_STATE, "no frame processed yet");
    uint8_t* b = c->h_out[c->parity] + (size_t)s * c->out_stride;
    const size_t n = (size_t) reinterpret_cast<const avb_frame_header*>(b)->n_features;
    const int nm = c->g.NMAX;
    if (cell) memcpy(cell, out_cell(b, nm), n * 4);
    if (lifetime) memcpy(lifetime, out_life(b, nm), n * 4);
    if (cam0_xy) memcpy(cam0_xy, out_p0(b, nm), n * 8);
    if (cam1_xy) memcpy(cam1_xy, out_p1(b, nm), n * 8);
    return AVB_OK;
}

// ---- per-stage entry points --------------------------------------------------------------------------

static int api_slot(const avb_ctx* c, int slot) {       // 0 cur cam0, 1 cur cam1, 2 prev cam0, 3 prev cam1
    const int cam = slot & 1, par = (slot & 2) ? (c->parity ^ 1) : c->parity;
    return SLOT(cam, par);
}

extern "C" int avb_advance(avb_ctx* c) {
    if (!c) return AVB_E_INVALID;
    c->parity ^= 1;
    return AVB_OK;
}

extern "C" int avb_upload_stereo(avb_ctx* c, int s, const uint8_t* img0, const uint8_t* img1, int stride) {
    if (!c || s < 0 || s >= c->g.S || !img0 || !img1) return AVB_E_INVALID;
    const Geom& g = c->g;
    if (stride < g.W) return fail(c, AVB_E_INVALID, "stride %d < width %d", stride, g.W);
    CK(cudaSetDevice(c->cfg.device));
    const size_t ib = (size_t)g.W * g.H;
    for (int cam = 0; cam < 2; ++cam) {
        uint8_t* dst = c->d.in[c->parity] + ((size_t)s * 2 + cam) * ib;
        CK(cudaMemcpy2DAsync(dst, g.W, cam ? img1 : img0, stride, g.W, g.H, cudaMemcpyHostToDevice, c->st));
    }
    CK(cudaStreamSynchronize(c->st));
    return AVB_OK;
}

extern "C" int avb_build_pyramids(avb_ctx* c) {
    if (!c) return AVB_E_INVALID;
    CK(cudaSetDevice(c->cfg.device));
    launch_pyramid(c->g, c->d, c->maps, c->parity, c->st);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(c->st));
    return AVB_OK;
}

extern "C" int avb_download_level(avb_ctx* c, int s, int slot, int level, uint8_t* out, int* w, int* h) {
    if (!c || s < 0 || s >= c->g.S || slot < 0 || slot > 3 || level < 0 || level >= c->g.nlev) return AVB_E_INVALID;
    const Geom& g = c->g;
    CK(cudaSetDevice(c->cfg.device));
    const int sl = api_slot(c, slot);
    const uint8_t* src = level ? pyr_slot(c->d, g, s, sl) + g.lv[level].off : level0_ptr(c->d, g, s, sl);
    if (w) *w = g.lv[level].w;
    if (h) *h = g.lv[level].h;
    if (out) {
        CK(cudaMemcpy2DAsync(out, g.lv[level].w, src, g.lv[level].pitch, g.lv[level].w, g.lv[level].h, cudaMemcpyDeviceToHost, c->st));
        CK(cudaStreamSynchronize(c->st));
    }
    return AVB_OK;
}

extern "C" int avb_fast_detect(avb_ctx* c, int s, const uint8_t* mask, int32_t* xs, int32_t* ys, int32_t* responses, int* n) {
    if (!c || s < 0 || s >= c->g.S || !n) return AVB_E_INVALID;
    const Geom& g = c->g;
    CK(cudaSetDevice(c->cfg.device));
    launch_clear_frame(g, c->d, c->st);
    launch_fast(g, c->d, c->maps, c->parity, c->st);
    CK(cudaGetLastError());
    std::vector<int> counts(g.NC);
    std::vector<unsigned> keys((size_t)g.NC * g.KPC);
    CK(cudaMemcpyAsync(counts.data(), c->d.kp_count + (size_t)s * g.NC, g.NC * sizeof(int), cudaMemcpyDeviceToHost, c->st));
    CK(cudaMemcpyAsync(keys.data(), c->d.kp_key + (size_t)s * g.NC * g.KPC, keys.size() * sizeof(unsigned), cudaMemcpyDeviceToHost, c->st));
    launch_clear_frame(g, c->d, c->st);    // the frame chain expects empty buckets
    CK(cudaStreamSynchronize(c->st));
    std::vector<unsigned long long> all;   // (scan index << 8) | response
    for (int cell = 0; cell < g.NC; ++cell) {
        if (counts[cell] > g.KPC) return fail(c, AVB_E_CAPACITY, "FAST bucket overflow in cell %d (%d > %d)", cell, counts[cell], g.KPC);
        for (int j = 0; j < counts[cell]; ++j) {
            int r, x, y;
            kp_decode(keys[(size_t)cell * g.KPC + j], g.W, r, x, y);
            if (mask && !mask[(size_t)y * g.W + x]) continue;
            all.push_back(((unsigned long long)(y * g.W + x) << 8) | (unsigned)r);
        }
    }
    std::sort(all.begin(), all.end());
    const int cap = *n;
    *n = (int)all.size();
    if ((int)all.size() > cap) return fail(c, AVB_E_CAPACITY, "%zu keypoints > capacity %d", all.size(), cap);
    for (size_t i = 0; i < all.size(); ++i) {
        const unsigned lin = (unsigned)(all[i] >> 8);
        if (xs) xs[i] = lin % g.W;
        if (ys) ys[i] = lin / g.W;
        if (responses) responses[i] = (int)(all[i] & 0xff);
    }
    return AVB_OK;
}

static int ensure_scratch(avb_ctx* c, int n) {
    if (n <= c->s_cap) return AVB_OK;
    const int cap = std::max(n, 4096);
    for (void* p : {(void*)c->s_a, (void*)c->s_b, (void*)c->s_c, (void*)c->s_st, (void*)c->s_da, (void*)c->s_db})
        if (p) cudaFree(p);
    c->s_a = c->s_b = c->s_c = nullptr;
    c->s_st = nullptr;
    c->s_da = c->s_db = nullptr;
    c->s_cap = 0;
    CK(cudaMalloc(&c->s_a, cap * sizeof(float2)));
    CK(cudaMalloc(&c->s_b, cap * sizeof(float2)));
    CK(cudaMalloc(&c->s_c, cap * sizeof(float2)));
    CK(cudaMalloc(&c->s_st, cap));
    CK(cudaMalloc(&c->s_da, cap * 2 * sizeof(double)));
    CK(cudaMalloc(&c->s_db, cap * 2 * sizeof(double)));
    if (!c->s_R) CK(cudaMalloc(&c->s_R, 9 * sizeof(double)));
    c->s_cap = cap;
    return AVB_OK;
}

extern "C" int avb_klt_track(avb_ctx* c, int s, int slot_from, int slot_to, const float* prev_xy, const float* guess_xy, int n,
                             float* out_xy, uint8_t* status) {
    if (!c || s < 0 || s >= c->g.S || slot_from < 0 || slot_from > 3 || slot_to < 0 || slot_to > 3 || n < 0) return AVB_E_INVALID;
    if (n == 0) return AVB_OK;
    if (!prev_xy || !guess_xy || !out_xy || !status) return AVB_E_INVALID;
    CK(cudaSetDevice(c->cfg.device));
    int r = ensure_scratch(c, n);
    if (r != AVB_OK) return r;
    CK(cudaMemcpyAsync(c->s_a, prev_xy, (size_t)n * 8, cudaMemcpyHostToDevice, c->st));
    CK(cudaMemcpyAsync(c->s_b, guess_xy, (size_t)n * 8, cudaMemcpyHostToDevice, c->st));
    launch_klt_points(c->g, c->d, s, api_slot(c, slot_from), api_slot(c, slot_to), c->s_a, c->s_b, n, c->s_c, c->s_st, c->st);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out_xy, c->s_c, (size_t)n * 8, cudaMemcpyDeviceToHost, c->st));
    CK(cudaMemcpyAsync(status, c->s_st, (size_t)n, cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->st));
    return AVB_OK;
}

extern "C" int avb_stereo_match(avb_ctx* c, int s, const float* cam0_xy, int n, float* cam1_xy, uint8_t* inlier) {
    if (!c || s < 0 || s >= c->g.S || n < 0) return AVB_E_INVALID;
    if (n == 0) return AVB_OK;
    if (!cam0_xy || !cam1_xy || !inlier) return AVB_E_INVALID;
    CK(cudaSetDevice(c->cfg.device));
    int r = ensure_scratch(c, n);
    if (r != AVB_OK) return r;
    CK(cudaMemcpyAsync(c->s_a, cam0_xy, (size_t)n * 8, cudaMemcpyHostToDevice, c->st));
    launch_stereo_points(c->g, c->d, s, c->parity, c->s_a, n, c->s_c, c->s_st, c->st);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(cam1_xy, c->s_c, (size_t)n * 8, cudaMemcpyDeviceToHost, c->st));
    CK(cudaMemcpyAsync(inlier, c->s_st, (size_t)n, cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->st));
    return AVB_OK;
}

static int undist_common(avb_ctx* c, const double* intr, const double* dist, const double* xy, int n, const double* R,
                         int f32_io, int distort, double* out) {
    if (!c || !intr || !dist || n < 0) return AVB_E_INVALID;
    if (n == 0) return AVB_OK;
    if (!xy || !out) return AVB_E_INVALID;
    CK(cudaSetDevice(c->cfg.device));
    int r = ensure_scratch(c, n);
    if (r != AVB_OK) return r;
    const CamModel cam = {intr[0], intr[1], intr[2], intr[3], dist[0], dist[1], dist[2], dist[3]};
    CK(cudaMemcpyAsync(c->s_da, xy, (size_t)n * 16, cudaMemcpyHostToDevice, c->st));
    if (R) CK(cudaMemcpyAsync(c->s_R, R, 9 * sizeof(double), cudaMemcpyHostToDevice, c->st));
    launch_undistort(cam, c->s_da, n, c->s_R, R ? 1 : 0, f32_io, distort, c->s_db, c->st);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, c->s_db, (size_t)n * 16, cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->st));
    return AVB_OK;
}

extern "C" int avb_undistort_points(avb_ctx* c, const double* intr, const double* dist, const double* xy, int n,
                                    const double* R, int f32_io, double* out_xy) {
    return undist_common(c, intr, dist, xy, n, R, f32_io, 0, out_xy);
}
extern "C" int avb_distort_points(avb_ctx* c, const double* intr, const double* dist, const double* xy, int n, int f32_io,
                                  double* out_xy) {
    return undist_common(c, intr, dist, xy, n, nullptr, f32_io, 1, out_xy);
}

extern "C" int avb_two_point_ransac(avb_ctx* c, const double* intr, const double* dist, const float* prev_xy, const float* cur_xy,
                                    int n, const double* R_p_c, double threshold_px, int seed, int frame_index, int cam,
                                    uint8_t* inlier) {
    if (!c || !intr || !dist || n < 0 || !(threshold_px > 0.0)) return AVB_E_INVALID;
    if (n == 0) return AVB_OK;
    if (!prev_xy || !cur_xy || !inlier) return AVB_E_INVALID;
    CK(cudaSetDevice(c->cfg.device));
    int r = ensure_scratch(c, n);
    if (r != AVB_OK) return r;
    const CamModel cm = {intr[0], intr[1], intr[2], intr[3], dist[0], dist[1], dist[2], dist[3]};
    const double ident[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    CK(cudaMemcpyAsync(c->s_a, prev_xy, (size_t)n * 8, cudaMemcpyHostToDevice, c->st));
    CK(cudaMemcpyAsync(c->s_b, cur_xy, (size_t)n * 8, cudaMemcpyHostToDevice, c->st));
    CK(cudaMemcpyAsync(c->s_R, R_p_c ? R_p_c : ident, 9 * sizeof(double), cudaMemcpyHostToDevice, c->st));
    CK(cudaStreamSynchronize(c->st));       // `ident` lives on this stack frame
    launch_ransac_points(c->g, cm, c->s_R, c->s_a, c->s_b, n, reinterpret_cast<float4*>(c->s_da), reinterpret_cast<int*>(c->s_db),
                         c->s_st, frame_index, cam, seed, threshold_px, c->st);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(inlier, c->s_st, (size_t)n, cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->st));
    return AVB_OK;
}

extern "C" int avb_time_pyramid(avb_ctx* c, int iters, float* ms_avg) {
    if (!c || iters < 1 || !ms_avg) return AVB_E_INVALID;
    CK(cudaSetDevice(c->cfg.device));
    CK(cudaStreamSynchronize(c->st));
    CK(cudaEventRecord(c->ev_t0, c->st));
    for (int i = 0; i < iters; ++i) launch_pyramid(c->g, c->d, c->maps, c->parity, c->st);
    CK(cudaEventRecord(c->ev_t1, c->st));
    CK(cudaStreamSynchronize(c->st));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, c->ev_t0, c->ev_t1);
    *ms_avg = ms / iters;
    return AVB_OK;
}
