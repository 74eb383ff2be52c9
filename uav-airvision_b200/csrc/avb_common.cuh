// Shared device-side layout for libavb (sm_100a only).  See DESIGN.md "Data layout in HBM".
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/avb.h"

#define AVB_WIN 15                 // LK window (cfg.patch_size); the lane mapping is built around 15x15
#define AVB_HALF 7
#define AVB_MAX_CAP 32             // grid_max_feature_num upper bound (one warp ballots a cell's candidates)
#define AVB_MAX_CELLS 1024
#define AVB_ROT_DOUBLES 27         // per stream and frame: H (9) | cam0_R_p_c (9) | cam1_R_p_c (9)

// Image slots of a stream: slot = cam * 2 + parity (parity = frame index & 1).  Level 0 of a slot is the
// uploaded image itself inside the input block of that parity; levels >= 1 live in the pyramid arena.
#define SLOTS_PER_STREAM 4
#define SLOT(cam, parity) ((cam) * 2 + (parity))

struct LevelGeom {
    int w, h, pitch;
    unsigned long long off;        // byte offset inside one pyramid slot (levels >= 1)
};

// One image pyramid as the LK code sees it: level 0 is the uploaded image, levels >= 1 sit at
// arena + Geom::lv[level].off.
struct PyrView {
    const uint8_t* l0;
    const uint8_t* arena;
};

struct CamModel {
    double fx, fy, cx, cy;
    double k1, k2, p1, p2;
};

// Geometry + constants, identical for all streams of a context (kernel parameter, by value).
struct Geom {
    int W, H, nlev;
    LevelGeom lv[AVB_MAX_LEVELS];
    unsigned long long slot_bytes; // bytes of one pyramid slot (all levels), multiple of 256
    int S;
    int rows, cols, NC;            // grid
    int gh, gw;                    // cell size in px (ceil)
    int gmin, gmax;                // per-cell min/max feature counts; gmax = slot capacity of a cell
    int NMAX;                      // NC * gmax
    int KPC;                       // FAST bucket capacity of one cell
    int wpf;                       // warps per feature in the LK kernels: 4 = latency mapping, 1 = throughput mapping
    int pyr_pair_level;            // k_pyr_pair builds levels pyr_pair_level and +1 in one launch (0: one k_pyr_down per level)
    int cand_rounds;               // new-feature stereo matching in 1 launch or 2 dense rounds (avb_points.cu)
    int spec_k;                    // > 0 (few streams): the spec_k strongest FAST keypoints of every cell are stereo-matched
                                   // SPECULATIVELY beside k_track; k_select then only looks its candidates up (avb_points.cu)
    int spec_wpf;                  // warps per speculative candidate (1 or 4)
    int fast_thr;
    int max_iter;
    double min_eig;
    double eps2;                   // (track_precision)^2 in double, like criteria.epsilon
    float eps2_lo, eps2_hi;        // (float)(eps2 * (1 -+ 1e-5)): band outside which the float estimate of |delta|^2 decides
    float eig_accept;              // 450 * (2 min_eig + 4e-5): D > eig_accept * (A11 + A22) proves minEig >= threshold (avb_lk.cuh)
    CamModel cam0, cam1;
    double R01[9];                 // R_cam0_to_cam1
    double E[9];                   // essential
    double epi_thr;                // stereo_threshold * 4/(2fx+2fy)
    int ransac;                    // 1: k_ransac runs between k_track and k_select (not the reference's behaviour)
    int ransac_seed, ransac_iters; // counter-based draws; ceil(log(1-0.99)/log(1-0.7^2)) = 7 hypotheses
    double ransac_thr;             // cfg.ransac_threshold (px)
};

// Feature table in grid order: cell c owns slots [c*gmax, c*gmax + count[c]).
struct GridTable {
    long long* ids;                // [S][NMAX]
    int* life;                     // [S][NMAX]
    float2* p0;                    // [S][NMAX]  cam0 pixel
    float2* p1;                    // [S][NMAX]  cam1 pixel
    uint8_t* fresh;                // [S][NMAX]  1 = created this frame
    int* count;                    // [S][NC]
};

struct FrameOut {                  // per-stream block in pinned host + mirror on device
    avb_frame_header hdr;
    // followed by: int64 ids[NMAX]; double meas[NMAX*4]; int32 cell[NMAX]; int32 life[NMAX];
    //              float p0[NMAX*2]; float p1[NMAX*2]
};

__host__ __device__ inline size_t out_stride_bytes(int nmax) {
    size_t b = sizeof(avb_frame_header) + (size_t)nmax * (8 + 32 + 4 + 4 + 8 + 8);
    return (b + 255) & ~(size_t)255;
}
__host__ __device__ inline long long* out_ids(uint8_t* base) { return (long long*)(base + sizeof(avb_frame_header)); }
__host__ __device__ inline double* out_meas(uint8_t* base, int nmax) { return (double*)(base + sizeof(avb_frame_header) + (size_t)nmax * 8); }
__host__ __device__ inline int* out_cell(uint8_t* base, int nmax) { return (int*)(base + sizeof(avb_frame_header) + (size_t)nmax * 40); }
__host__ __device__ inline int* out_life(uint8_t* base, int nmax) { return (int*)(base + sizeof(avb_frame_header) + (size_t)nmax * 44); }
__host__ __device__ inline float* out_p0(uint8_t* base, int nmax) { return (float*)(base + sizeof(avb_frame_header) + (size_t)nmax * 48); }
__host__ __device__ inline float* out_p1(uint8_t* base, int nmax) { return (float*)(base + sizeof(avb_frame_header) + (size_t)nmax * 56); }

// All device buffers of a context (kernel parameter, by value).
struct DevState {
    uint8_t* in[2];                // input block per parity: [S][2 cams][H*W] images, then [S][27] doubles
                                   // (H = K0 R0 K0^-1 | R0 = cam0_R_p_c | R1 = cam1_R_p_c)
    uint8_t* pyr;                  // [S][4 slots][slot_bytes], levels 1..L
    // FAST buckets
    unsigned* kp_key;              // [S][NC][KPC]  (response << 24) | (0xFFFFFF - (y*W + x))
    int* kp_count;                 // [S][NC]
    float2* kp_p1;                 // [S][NC][KPC]  frame-0 stereo result
    uint8_t* kp_ok;                // [S][NC][KPC]
    // feature tables (ping/pong)
    GridTable grid[2];
    // tracking scratch, indexed like the previous grid table
    float2* t_p0;                  // [S][NMAX] tracked cam0 position in the current frame
    float2* t_p1;                  // [S][NMAX]
    int* t_cell;                   // [S][NMAX] new cell, -1 = lost
    double4* t_und;                // [S][NMAX] normalized coordinates (u0 v0 u1 v1) of the tracked feature
    // new-feature candidates, [cell][gmax], in descending key order
    unsigned* c_key;               // [S][NMAX]
    int* c_src;                    // [S][NMAX] index into the cell's FAST bucket
    float2* c_p1;                  // [S][NMAX]
    uint8_t* c_ok;                 // [S][NMAX]
    double4* c_und;                // [S][NMAX] normalized coordinates of the candidate (valid when c_ok)
    int* c_count;                  // [S][NC]
    // speculative matches (Geom::spec_k > 0): the spec_k largest keys of each cell's FAST bucket, unmasked, and their stereo result
    unsigned* s_key;               // [S][NC][spec_k], descending
    float2* s_p1;                  // [S][NC][spec_k]
    uint8_t* s_ok;                 // [S][NC][spec_k]
    double4* s_und;                // [S][NC][spec_k]
    int* s_n;                      // [S][NC]
    int* n_new;                    // [S][NC]  new features given ids this frame (before pruning)
    long long* next_id;            // [2 parities][S]: read from the previous frame's parity, written to this frame's
    int* counters;                 // [S][8]: before_tracking, after_tracking, after_matching, n_fast, n_cand, after_ransac
    // two-point RANSAC scratch, [2 cams][S][NMAX] (allocated only when Geom::ransac)
    int* r_idx;
    float4* r_und;
    float4* r_prev;                // [S][NMAX] previous positions, gyro-rotated and undistorted (cam0 xy, cam1 zw): written by k_track
    int* r_raw;
    uint8_t* r_bits;
    uint8_t* out;                  // [S][out_stride] device mirror of the result block
    int* frame_index;              // [S]
};

__host__ __device__ inline uint8_t* pyr_slot(const DevState& d, const Geom& g, int s, int slot) {
    return d.pyr + ((size_t)s * SLOTS_PER_STREAM + slot) * g.slot_bytes;
}
__host__ __device__ inline uint8_t* level0_ptr(const DevState& d, const Geom& g, int s, int slot) {
    return d.in[slot & 1] + ((size_t)s * 2 + (slot >> 1)) * ((size_t)g.W * g.H);
}
__host__ __device__ inline size_t in_images_bytes(const Geom& g) {
    return (((size_t)g.S * 2 * g.W * g.H) + 255) & ~(size_t)255;
}
__host__ __device__ inline size_t in_block_bytes(const Geom& g) {
    return in_images_bytes(g) + (((size_t)g.S * AVB_ROT_DOUBLES * sizeof(double)) + 255 & ~(size_t)255);
}
__host__ __device__ inline const double* frame_H(const DevState& d, const Geom& g, int s, int parity) {
    return reinterpret_cast<const double*>(d.in[parity] + in_images_bytes(g)) + (size_t)s * AVB_ROT_DOUBLES;
}

__device__ __forceinline__ PyrView pyr_view(const DevState& d, const Geom& g, int s, int slot) {
    PyrView v;
    v.l0 = level0_ptr(d, g, s, slot);
    v.arena = pyr_slot(d, g, s, slot);
    return v;
}
__device__ __forceinline__ const uint8_t* pyr_level(const PyrView& v, const Geom& g, int level) {
    return level ? v.arena + g.lv[level].off : v.l0;
}

__device__ __forceinline__ int refl101(int i, int n) {
    i = i < 0 ? -i : i;
    return i >= n ? 2 * n - 2 - i : i;
}

// floor(t / d) by a host-made reciprocal (magic = ceil(2^32 / d), exact while t * d < 2^32; 0 = divide): tile and cell
// decompositions run once per tile and per keypoint in every thread, and an emulated integer division is ~20 instructions
struct FastDiv {
    unsigned d, magic;
};
__device__ __forceinline__ int fdiv(int t, const FastDiv& f) { return f.magic ? (int)__umulhi((unsigned)t, f.magic) : t / (int)f.d; }
inline FastDiv make_fdiv(int d, long long max_t) {
    FastDiv f;
    f.d = (unsigned)d;
    f.magic = (d > 1 && max_t * d < (1ll << 32)) ? (unsigned)(((1ull << 32) + d - 1) / d) : 0u;
    return f;
}

// FAST keypoint key: larger key = earlier in the reference's ranking
// (response desc, then row-major scan order; stable sorted(..., reverse=True), Appendix B9).
__host__ __device__ inline unsigned kp_make_key(int resp, int x, int y, int W) {
    return ((unsigned)resp << 24) | (0xFFFFFFu - (unsigned)(y * W + x));
}
__host__ __device__ inline void kp_decode(unsigned key, int W, int& resp, int& x, int& y) {
    resp = (int)(key >> 24);
    unsigned lin = 0xFFFFFFu - (key & 0xFFFFFFu);
    y = (int)(lin / (unsigned)W);
    x = (int)(lin % (unsigned)W);
}

// ---- programmatic dependent launch (sm_90+) -------------------------------------------------------------
// The frame is a chain of small dependent kernels; with the ProgrammaticStreamSerialization attribute the next
// kernel's CTAs are scheduled while the current one drains and sit in griddepcontrol.wait until it has completed and
// flushed, which hides most of the ~3 us launch gap per kernel boundary.  Every kernel of the chain executes
// pdl_wait() before it touches memory and pdl_launch_dependents() right after (a no-op without the attribute).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

#ifdef __CUDACC__
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl,
                            Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif
extern int g_avb_pdl;              // 1: chain kernels are launched with the programmatic attribute (AVB_PDL=0 disables)

// ---- launchers implemented in the stage files -------------------------------------------
// TMA views (x, y, image).  Level 0: one map per parity over the input block, image = s*2 + cam.
// Levels >= 1: one map per level over the arena, image = s*4 + slot.  `fast0` has the FAST box shape.
struct PyrMaps {
    // k_pyr_down, per tile-height variant (0: 16 rows, 1: 32 rows)
    CUtensorMap l0[2][2];           // [parity][variant] source level 0
    CUtensorMap lv[AVB_MAX_LEVELS][2];  // source level l >= 1
    CUtensorMap dst[AVB_MAX_LEVELS][2]; // destination level l >= 1 (store box 64 x rows)
    CUtensorMap pair0[2];           // source level 0, box of k_pyr_pair (used when the pair kernel builds levels 1+2)
    CUtensorMap pair;               // source level nlev-3 >= 1, box of k_pyr_pair
    CUtensorMap fast0[2][2];        // [parity][k_fast variant]
};

void launch_pyramid(const Geom& g, const DevState& d, const PyrMaps& maps, int parity, cudaStream_t st);
int  avb_pyramid_launches(const Geom& g);
void avb_pyramid_boxes(int variant, int* src_w, int* src_h, int* dst_w, int* dst_h);
void launch_fast(const Geom& g, const DevState& d, const PyrMaps& maps, int parity, cudaStream_t st);
void avb_fast_box(int variant, int* w, int* h);
void launch_track(const Geom& g, const DevState& d, int parity, cudaStream_t st);
void launch_ransac(const Geom& g, const DevState& d, int parity, cudaStream_t st);
void launch_ransac_points(const Geom& g, const CamModel& cm, const double* R, const float2* prev, const float2* cur, int n,
                          float4* und, int* raw_idx, uint8_t* bits, int frame_index, int cam_key, int seed, double thr_px,
                          cudaStream_t st);
void launch_select(const Geom& g, const DevState& d, int parity, int first_frame, cudaStream_t st);
void launch_stereo_candidates(const Geom& g, const DevState& d, int parity, cudaStream_t st);
void launch_spec_select(const Geom& g, const DevState& d, cudaStream_t st);
void launch_spec_match(const Geom& g, const DevState& d, int parity, cudaStream_t st);
int  avb_candidate_rounds(const Geom& g);    // default for Geom::cand_rounds. 1: every candidate in one launch; 2: positions < gmin first, the rest on demand
int  avb_pyramid_pair_level(const Geom& g);  // default for Geom::pyr_pair_level
void launch_stereo_buckets(const Geom& g, const DevState& d, int parity, cudaStream_t st);
void launch_finish(const Geom& g, const DevState& d, int parity, int first_frame, cudaStream_t st);
void launch_clear_frame(const Geom& g, const DevState& d, cudaStream_t st);
int  avb_set_smem_limits(size_t select_bytes, size_t grid_bytes);
int  avb_preload_fast();
int  avb_preload_grid();
int  avb_preload_points();
int  avb_preload_pyramid();   // opt in to > 48 KB dynamic shared memory
// flat point lists (per-stage entry points)
void launch_klt_points(const Geom& g, const DevState& d, int s, int slot_from, int slot_to,
                       const float2* prev, const float2* guess, int n, float2* out, uint8_t* status,
                       cudaStream_t st);
void launch_stereo_points(const Geom& g, const DevState& d, int s, int parity, const float2* p0, int n,
                          float2* p1, uint8_t* ok, cudaStream_t st);
void launch_undistort(const CamModel& cam, const double* xy, int n, const double* R, int has_R, int f32_io,
                      int distort, double* out, cudaStream_t st);
