/*
 * avb.h -- C-ABI of the B200-native stereo image front end (libavb.so).
 *
 * Drop-in boundary for ONE path of BUBLET/uav-airvision: the per-frame image front end
 * `ImageProcessingPipeline.stereo_callback` (reference: src/image_processing/pipeline.py:46-150)
 * and the cv2 calls its stage classes make.  Plain pointers and sizes only; every function
 * returns 0 on success or a negative AVB_E_* code (text via avb_last_error).  There is no CPU
 * fallback: without a CUDA device avb_create fails.
 *
 * A context owns S independent streams (one ImageProcessingPipeline instance each) that are
 * processed in lock-step by the same kernel launches (blockIdx.z = stream).
 *
 * Reference interface each entry point replaces (file:line under /root/reference/src):
 *   avb_create / avb_destroy      ImageProcessingPipeline.__init__      image_processing/pipeline.py:15-40
 *   avb_upload_stereo             stereo_msg.cam{0,1}_msg.image intake  image_processing/pipeline.py:52-55
 *   avb_build_pyramids            PyramidBuilder.create_image_pyramids  image_processing/pyramid_builder.py:22-48
 *                                 (+ the pyramid cv2.calcOpticalFlowPyrLK builds internally)
 *   avb_fast_detect               detector.detect(img[, mask])          image_processing/feature_initializer.py:52,
 *                                                                       image_processing/feature_adder.py:64
 *   avb_klt_track                 cv2.calcOpticalFlowPyrLK              image_processing/feature_tracker.py:102-108,
 *                                                                       image_processing/stereo_matcher.py:64-74
 *   avb_stereo_match              StereoMatcher.stereo_match            image_processing/stereo_matcher.py:33-115
 *   avb_undistort_points          CameraModel.undistort_points          image_processing/camera_model.py:24-47
 *   avb_distort_points            CameraModel.distort_points            image_processing/camera_model.py:49-75
 *   avb_process_frame             ImageProcessingPipeline.stereo_callback  image_processing/pipeline.py:46-150
 *                                 (FeatureInitializer / FeatureTracker / FeatureAdder / FeaturePruner /
 *                                  FeaturePublisher fused into one CUDA-graph launch per frame)
 *   avb_get_features              pipeline.prev_features (state read-back)  image_processing/pipeline.py:145-148
 *   avb_two_point_ransac          (none: all-ones stub)                 image_processing/feature_tracker.py:135-136
 *   avb_store_* / avb_process_frame_gather
 *                                 EuRoCDataset image readers of a run.bat sweep   streaming/dataset.py:101-117, 206-214
 */
#ifndef AVB_H_
#define AVB_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AVB_ABI_VERSION 5   /* 3: + avb_store_*, avb_process_frame_gather, avb_enqueue_frame_gather; 4: + avb_get_result_prev
                             * (result blocks double-buffered by frame parity); 5: + avb_submit_images /
                             * avb_process_submitted; structs unchanged since 2 */

#define AVB_OK              0
#define AVB_E_INVALID      -1   /* bad argument / unsupported configuration */
#define AVB_E_CUDA         -2   /* CUDA runtime error (see avb_last_error)   */
#define AVB_E_NO_DEVICE    -3   /* no CUDA device: there is no CPU fallback  */
#define AVB_E_CAPACITY     -4   /* more points than the context was sized for */
#define AVB_E_STATE        -5   /* call order violated (e.g. track before any frame) */

#define AVB_MAX_LEVELS      6   /* pyramid images incl. level 0 */

typedef struct avb_config {
    int32_t width, height;            /* image size (both cameras)                          */
    int32_t max_level;                /* cfg.pyramid_levels: levels 0..max_level            */
    int32_t win_size;                 /* cfg.patch_size; only 15 is compiled                */
    int32_t max_iteration;            /* cfg.max_iteration                                  */
    int32_t fast_threshold;           /* cfg.fast_threshold                                 */
    int32_t grid_row, grid_col;       /* cfg.grid_row / grid_col                            */
    int32_t grid_min_feature_num;     /* cfg.grid_min_feature_num                           */
    int32_t grid_max_feature_num;     /* cfg.grid_max_feature_num (<= 32)                   */
    int32_t num_streams;              /* S independent pipelines processed per launch       */
    int32_t device;                   /* CUDA device ordinal                                */
    int32_t use_graph;                /* 1: replay the per-frame kernel chain as a CUDA graph */
    int32_t ransac;                   /* 0 = reference parity (all-ones stub, feature_tracker.py:135-136);
                                         1 = two-point RANSAC between k_track and the grid rebuild (config C3) */
    int32_t ransac_seed;              /* key of the counter-based draws (with stream frame index, camera, hypothesis) */
    int32_t reserved0;                /* keeps the doubles below 8-byte aligned; must be 0   */
    double  track_precision;          /* cfg.track_precision (LK epsilon, px)               */
    double  min_eig_threshold;        /* cv2 default 1e-4                                   */
    double  stereo_threshold;         /* cfg.stereo_threshold                               */
    double  ransac_threshold;         /* cfg.ransac_threshold (used only when ransac=1)     */
    double  cam0_intrinsics[4];       /* fx fy cx cy                                        */
    double  cam0_distortion[4];       /* k1 k2 p1 p2 (radtan)                               */
    double  cam1_intrinsics[4];
    double  cam1_distortion[4];
    double  R_cam0_to_cam1[9];        /* R_cam1_imu^T R_cam0_imu, row-major (stereo_matcher.py:49) */
    double  essential[9];             /* skew(t01) R0to1, row-major       (stereo_matcher.py:90-91) */
} avb_config;

/* Per-frame, per-stream result header followed by the arrays (all in one pinned block). */
typedef struct avb_frame_header {
    int64_t n_features;               /* len(feature_msg.features)                          */
    int64_t next_feature_id;          /* pipeline.next_feature_id after this frame          */
    int32_t before_tracking, after_tracking, after_matching, after_ransac;  /* num_features */
    int32_t has_new;                  /* 1: the frame holds freshly created features (u0,v0 f64 quirk B11) */
    int32_t n_fast;                   /* FAST keypoints after NMS (before the mask)         */
    int32_t n_candidates;             /* new-feature candidates sent to stereo matching     */
    int32_t frame_index;
} avb_frame_header;

typedef struct avb_ctx avb_ctx;

int  avb_abi_version(void);
const char* avb_last_error(const avb_ctx* ctx);       /* ctx may be NULL (creation errors) */

int  avb_create(const avb_config* cfg, avb_ctx** out);
void avb_destroy(avb_ctx* ctx);
int  avb_capacity(const avb_ctx* ctx);                /* max features per stream = grid_num*grid_max */
int  avb_num_cells(const avb_ctx* ctx);
int  avb_get_geometry(const avb_ctx* ctx, int* width, int* height, int* num_streams);
int  avb_reset(avb_ctx* ctx);                         /* back to first_frame = True for all streams */

/* One frame's input for all streams is a single "input block": S*2 images of width*height bytes in
 * [stream][cam] order, padded to 256 B, then per stream 27 doubles: H = K0 R_p_c0 K0^-1 (gyro prediction,
 * feature_tracker.py:159-177), R_p_c0 and R_p_c1 (the two rotations IMUProcessor.integrate_imu_data returns,
 * imu_processor.py:55-66; read by the RANSAC kernel only), each 3x3 row-major.
 * avb_input_staging returns the context's pinned block, which the caller may fill in place (zero-copy
 * intake); avb_process_frame(…, img0 = img1 = NULL) consumes it as is. */
uint8_t* avb_input_staging(avb_ctx* ctx);
size_t   avb_input_block_bytes(const avb_ctx* ctx);
size_t   avb_input_rotation_offset(const avb_ctx* ctx);
size_t   avb_input_rotation_stride(const avb_ctx* ctx);   /* bytes per stream in the rotation section (27 doubles) */
/* Writes the rotation section for every stream into `block` (host memory, block layout above).
 * R_p_c0 / R_p_c1 = S*9 doubles each (cam0_R_p_c, cam1_R_p_c of IMUProcessor.integrate_imu_data);
 * R_p_c0 NULL = identity for both; R_p_c1 NULL = the same gyro rotation seen from cam1,
 * R_cam0_to_cam1 R_p_c0 R_cam0_to_cam1^T. */
int      avb_fill_rotations(const avb_ctx* ctx, uint8_t* block, const double* R_p_c0, const double* R_p_c1);

/* ---- hot path -------------------------------------------------------------------------- */

/* One stereo frame for every stream.  img0[s]/img1[s]: host uint8, row stride = stride bytes
 * (NULL arrays: take the pinned staging).  R_p_c0 / R_p_c1: S*9 doubles each, the camera rotations from the
 * IMU integration (NULL rules of avb_fill_rotations; identity on frame 0).  Blocks until the results are in
 * host memory. */
int  avb_process_frame(avb_ctx* ctx, const uint8_t* const* img0, const uint8_t* const* img1,
                       int stride, const double* R_p_c0, const double* R_p_c1);

/* avb_process_frame in two halves.  avb_submit_images starts the intake (host staging, H2D copies and, for a single
 * stream, the cam0-only kernels: FAST and the speculative candidate list); nothing in it needs the frame's gyro
 * rotations, so the caller can integrate the IMU window (IMUProcessor.integrate_imu_data, imu_processor.py:28-67;
 * pipeline.py:52-55 does it at the top of stereo_callback) while the images are on the bus.  avb_process_submitted
 * then takes the rotations, runs the rest of the frame and blocks until the results are in host memory.  The image
 * buffers must stay valid until it returns.  AVB_E_STATE: submit twice, finish without submit, or any other frame
 * entry point in between. */
int  avb_submit_images(avb_ctx* ctx, const uint8_t* const* img0, const uint8_t* const* img1, int stride);
int  avb_process_submitted(avb_ctx* ctx, const double* R_p_c0, const double* R_p_c1);

/* Same, the whole input block already resident in device memory (layout above). */
int  avb_process_frame_device(avb_ctx* ctx, const uint8_t* d_block);

/* Asynchronous variant of the device-resident path: enqueue only; avb_sync waits.  Frames of one
 * context are ordered on its stream, so K enqueues + one sync time K dependent frames. */
int  avb_enqueue_frame_device(avb_ctx* ctx, const uint8_t* d_block);
int  avb_sync(avb_ctx* ctx);

/* ---- HBM-resident sequence store (sweeps of time-offset runs) ---------------------------
 * Replaces, for sweeps, the per-run image readers of the reference: every `python main.py --path <seq>
 * --offset <s>` of run.bat:4-12 re-reads and re-decodes its frames (streaming/dataset.py:101-117, 206-214).
 * Here a sequence is decoded and uploaded once; the S offset runs of a context take their current frames
 * straight from device memory.  Frame k holds cam0 then cam1, width*height dense bytes each. */
typedef struct avb_store avb_store;
int  avb_store_create(int device, int width, int height, int n_frames, avb_store** out);
void avb_store_destroy(avb_store* store);
int  avb_store_num_frames(const avb_store* store);
size_t avb_store_bytes(const avb_store* store);        /* device bytes held */
/* Host images (uint8, row stride in bytes) -> frame k of the store.  Blocking; done once per sequence frame. */
int  avb_store_upload(avb_store* store, int k, const uint8_t* img0, const uint8_t* img1, int stride);
/* Device pointer of image `cam` of frame k (NULL when out of range). */
const uint8_t* avb_store_image(const avb_store* store, int k, int cam);

/* One stereo frame for every stream from images that already live in device memory: d_images[s*2 + cam] is a
 * dense width*height uint8 image, 16-byte aligned (e.g. avb_store_image).  A gather kernel places them in the
 * input block; rotations as in avb_process_frame.  Blocks until the results are in host memory. */
int  avb_process_frame_gather(avb_ctx* ctx, const uint8_t* const* d_images, const double* R_p_c0,
                              const double* R_p_c1);
/* Enqueue-only variant (avb_sync waits): the caller prepares the next step (IMU windows of every stream) while
 * this one runs.  One frame in flight per context: call avb_sync before any result read.  A second enqueue without
 * avb_sync is safe but not useful: it waits until the previous frame's rotation section has left the (single) pinned
 * staging area, and the result block holds the latest frame only. */
int  avb_enqueue_frame_gather(avb_ctx* ctx, const uint8_t* const* d_images, const double* R_p_c0,
                              const double* R_p_c1);

/* Results of the last frame for stream s (pointers into pinned host memory, valid until the
 * frame after next: two blocks, alternating by frame parity): ids[n], meas[n*4] = u0 v0 u1 v1 (normalized coords; u0,v0 carry the
 * f64 result, u1,v1 the f32-rounded one, as the reference publishes them). */
int  avb_get_result(avb_ctx* ctx, int s, const avb_frame_header** hdr,
                    const int64_t** ids, const double** meas);
/* The same for the frame BEFORE the last one enqueued.  The result blocks are double-buffered by frame parity, so a
 * sweep driver may enqueue step k+1 (avb_enqueue_frame_gather) as soon as step k has completed (avb_sync) and read
 * step k's results here while k+1 runs (the reference's runs have no such coupling: every run is a process of its
 * own, run.bat:4-12).  Valid until the frame after next is enqueued. */
int  avb_get_result_prev(avb_ctx* ctx, int s, const avb_frame_header** hdr,
                         const int64_t** ids, const double** meas);

/* State read-back for stream s in grid order (pipeline.prev_features after the roll):
 * cell[n], lifetime[n], cam0_xy[n*2], cam1_xy[n*2] (pixels, f32). Any pointer may be NULL. */
int  avb_get_features(avb_ctx* ctx, int s, int32_t* cell, int32_t* lifetime,
                      float* cam0_xy, float* cam1_xy);

/* ---- per-stage entry points (stage classes / tests; operate on stream s) ---------------- */

/* Image slots below: 0 = current cam0, 1 = current cam1, 2 = previous cam0, 3 = previous cam1.
 * avb_upload_stereo overwrites the CURRENT images of stream s; avb_advance makes the current frame
 * the previous one for all streams (what the roll at pipeline.py:145-148 does) without processing. */
int  avb_upload_stereo(avb_ctx* ctx, int s, const uint8_t* img0, const uint8_t* img1, int stride);
int  avb_advance(avb_ctx* ctx);
int  avb_build_pyramids(avb_ctx* ctx);                /* levels 1..max_level of the current cam0 + cam1 images, all streams */
/* Copies pyramid level `level` of an image slot to host (dense, w*h bytes). */
int  avb_download_level(avb_ctx* ctx, int s, int slot, int level, uint8_t* out, int* w, int* h);
/* FAST-9/16 + NMS on current cam0.  mask (optional, width*height u8, 0 = drop) is a post-filter.
 * Returns the keypoints in row-major scan order: xs[n], ys[n], responses[n]; *n in: capacity, out: count. */
int  avb_fast_detect(avb_ctx* ctx, int s, const uint8_t* mask, int32_t* xs, int32_t* ys,
                     int32_t* responses, int* n);
/* calcOpticalFlowPyrLK(USE_INITIAL_FLOW) between two image slots. */
int  avb_klt_track(avb_ctx* ctx, int s, int slot_from, int slot_to, const float* prev_xy,
                   const float* guess_xy, int n, float* out_xy, uint8_t* status);
int  avb_stereo_match(avb_ctx* ctx, int s, const float* cam0_xy, int n, float* cam1_xy,
                      uint8_t* inlier);
/* Point undistortion / distortion with an explicit radtan model: intrinsics4 = fx fy cx cy,
 * distortion4 = k1 k2 p1 p2.  R: optional 3x3 row-major matrix applied after undistortion (cv2's R, or
 * P*R when new intrinsics are wanted).  f32_io = 1 rounds input and result to float32, as cv2 does for
 * float32 input. */
int  avb_undistort_points(avb_ctx* ctx, const double* intrinsics4, const double* distortion4,
                          const double* xy, int n, const double* R, int f32_io, double* out_xy);
int  avb_distort_points(avb_ctx* ctx, const double* intrinsics4, const double* distortion4,
                        const double* xy, int n, int f32_io, double* out_xy);

/* Two-point RANSAC between the previous and the current frame of ONE camera (k_ransac on a flat list).  Not a
 * reference interface: the reference's masks are an all-ones stub (feature_tracker.py:135-136); this is the stage
 * BASELINE config C3 names, defined by oracle/ransac.py.  prev_xy / cur_xy: n pixel positions (f32 pairs);
 * R_p_c: the camera's gyro rotation (9 doubles, NULL = identity); draws are keyed by (seed, frame_index, cam). */
int  avb_two_point_ransac(avb_ctx* ctx, const double* intrinsics4, const double* distortion4, const float* prev_xy,
                          const float* cur_xy, int n, const double* R_p_c, double threshold_px, int seed,
                          int frame_index, int cam, uint8_t* inlier);

/* ---- measurement hooks ------------------------------------------------------------------ */

/* Device time (ms, CUDA events on the launching stream) of the last avb_process_frame*. */
int  avb_last_frame_ms(avb_ctx* ctx, float* ms);
/* Number of kernels this library launches per steady-state frame (for bench.py gpu_launches). */
int  avb_kernels_per_frame(const avb_ctx* ctx);
/* Raw CUDA stream handle (cudaStream_t) so callers can bracket work with their own events. */
void* avb_cuda_stream(avb_ctx* ctx);
/* Instrumented steady-state frame from a device-resident input block: kernels run serialised with a CUDA
 * event after every stage.  stage_ms[9] = input copy, FAST, pyramid (all levels), track, select,
 * stereo match of new candidates, finish (grid update + publish), 0 (reserved), result copy.  Advances the stream
 * like a frame. */
int  avb_profile_frame_device(avb_ctx* ctx, const uint8_t* d_block, float* stage_ms);
/* Per-kernel timing: run the last frame's pyramid kernel `iters` times back to back on the
 * context stream and return the average device ms (bench.py roofline for the HBM-bound stage). */
int  avb_time_pyramid(avb_ctx* ctx, int iters, float* ms_avg);

#ifdef __cplusplus
}
#endif
#endif /* AVB_H_ */
