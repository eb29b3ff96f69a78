/* vpb200.h - C ABI of libvpb200.so: the B200 (sm_100a) enhance-chain behind video-processor's
 * Upscaler / SelectiveBilateral / AdaptiveSharpening / TemporalConsistency / FrameBuffer API.
 *
 * The reference (skandvarma/video-processor) is C++ over OpenCV and has no FFI of its own; every
 * entry point below names the reference interface it replaces (paths relative to the reference
 * repository).  The C++ classes in include/vp/ (same names, signatures and Config field order as the
 * reference headers) are thin wrappers over this ABI; INTEGRATION.md shows the binding a maintainer
 * adds to the reference tree.
 *
 * Conventions
 *   - frames are 8-bit BGR interleaved (CV_8UC3), row-major, `pitch`/`step` in bytes (>= 3*width);
 *   - every function returns VP_OK (0) or a negative vp_status; nothing throws or aborts;
 *     vp_last_error() returns a human-readable description of the last failure on that context;
 *   - a vp_ctx is bound to one CUDA device and is used by one caller thread at a time (the
 *     reference's "processing thread", src/pipeline.cpp:351-383, src/main.cpp:222-259);
 *   - device-pointer entry points are asynchronous on `stream` (a cudaStream_t passed as void*,
 *     NULL = the context's own compute stream); host-pointer entry points are synchronous unless
 *     named submit/wait.  The stage-wise entry points keep their parameter tables and scratch planes per
 *     context, not per stream: use ONE stream per vp_ctx for them (a call with new parameters on a second
 *     stream may rewrite tables a kernel still running on the first one reads); contexts are cheap, take one
 *     per stream;
 *   - there is NO CPU fallback: without a usable sm_100 device vp_create() fails.
 */
#ifndef VPB200_H
#define VPB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VPB200_ABI_VERSION 5

typedef enum vp_status {
    VP_OK = 0,
    VP_ERR_INVALID_ARG = -1,   /* null pointer, non-positive size, pitch < 3*width ...           */
    VP_ERR_UNSUPPORTED = -2,   /* configuration outside the built path             */
    VP_ERR_CUDA = -3,          /* a CUDA runtime call failed; vp_last_error() carries the string   */
    VP_ERR_NO_DEVICE = -4,     /* no CUDA device / not sm_100                                      */
    VP_ERR_STATE = -5,         /* call sequence error (e.g. vp_wait without vp_submit)             */
    VP_ERR_NOMEM = -6
} vp_status;

/* SelectiveBilateral::Config (include/selective_bilateral.h:20-40).  process() dispatches as the reference does
 * (src/selective_bilateral.cpp:57-82): use_multiscale -> applyMultiscaleBilateral (:138-237, the config's diameter at
 * every scale, sigmas x (1 + 0.5 i), pyrDown / pyrUp, detail-weighted coarse-to-fine blend); else selective ->
 * applySelectiveBilateral (:121-136: createDetailMask :239-313 + applyJointBilateral :373-439); else the plain
 * applyBilateralFilter (:84-119).  The two detail-mask branches run with REPAIRED semantics: the literal code
 * calls cv::sqrt on a CV_8U Mat (:276, undefined behaviour); here local_var is converted to float32 first and a zero
 * maximum normalises to 0 instead of NaN.  Fields after sigma_space default to 0 = plain. */
enum { VP_STAGE_PRE = 0, VP_STAGE_POST = 1 };   /* SelectiveBilateral::ProcessingStage */
typedef struct vp_bilateral_config {
    int32_t stage;             /* VP_STAGE_PRE / VP_STAGE_POST                                    */
    int32_t adaptive_params;   /* calculateAdaptiveParams, src/selective_bilateral.cpp:315-371    */
    int32_t diameter;
    int32_t reserved_;         /* keeps the doubles 8-byte aligned on every ABI                   */
    double sigma_color;        /* double in the reference's Config                                */
    double sigma_space;
    int32_t selective;         /* Config::selective                                               */
    int32_t use_multiscale;    /* Config::use_multiscale                                          */
    int32_t num_scales;        /* Config::num_scales, clamped to 1..5 as initialize() does (:40-43) */
    int32_t reserved2_;
    double detail_threshold;   /* Config::detail_threshold (0..255 scale, :291)                   */
    double edge_preserve;      /* Config::edge_preserve (:392)                                    */
} vp_bilateral_config;

/* AdaptiveSharpening::Config (include/adaptive_sharpening.h:14-24).  adaptive_sigma = 0: createEdgeMask +
 * applyUnsharpMask (src/adaptive_sharpening.cpp:107-355), the north-star branch.  adaptive_sigma = 1:
 * calculateTextureMap + calculateAdaptiveSigma + applyVariableSigmaUnsharpMask (:357-592) with REPAIRED semantics
 * (the literal code calls cv::sqrt on a CV_8U Mat, :386; local_var is converted to float32 first, degenerate
 * ranges give 0 instead of NaN - DESIGN.md section 4f). */
typedef struct vp_sharpen_config {
    float strength;
    float edge_strength;
    float smooth_strength;
    float edge_threshold;
    float sigma;
    int32_t kernel_size;       /* odd, 3..7 */
    int32_t preserve_tone;
    int32_t adaptive_sigma;    /* Config::adaptive_sigma */
} vp_sharpen_config;

/* Temporal stage.  VP_TEMPORAL_MAIN_BLEND = src/main.cpp:269-292 (addWeighted(cur,w,prev,1-w) on
 * the history of unblended frames).  VP_TEMPORAL_BLEND_FRAMES = TemporalConsistency::blendFrames
 * (src/temporal_consistency.cpp:355-444) over the current frame and every buffered unblended frame - up to
 * buffer_size of them: m_frame_buffer keeps the last buffer_size inputs (:162-165) and m_flow_buffer reaches
 * buffer_size entries once the new flow is pushed (:97; it is trimmed to buffer_size-1 only after the blend,
 * :167-169), so the loop of :127 runs over buffer_size previous frames.  Masks == 1, float32 accumulator. */
enum { VP_TEMPORAL_NONE = 0, VP_TEMPORAL_MAIN_BLEND = 1, VP_TEMPORAL_BLEND_FRAMES = 2 };
typedef struct vp_temporal_config {
    int32_t mode;
    int32_t buffer_size;       /* TemporalConsistency::Config::buffer_size (1..3)                */
    float blend_strength;      /* TemporalConsistency::Config::blend_strength                    */
    float main_weight;         /* current-frame weight of the main.cpp blend (0.6 / 0.7)         */
    int32_t scene_detect;      /* VP_TEMPORAL_BLEND_FRAMES only: TemporalConsistency::detectSceneChange
                                  (src/temporal_consistency.cpp:283-323: 64-bin histogram correlation + mean
                                  absolute difference of the gray frames) - a cut clears the history and passes
                                  the frame through (:100-112).  Needs a whole-frame reduction before the blend,
                                  so the blend runs as its own kernel instead of in the POST epilogue.          */
    float scene_change_threshold;   /* TemporalConsistency::Config::scene_change_threshold (100)  */
    /* VP_TEMPORAL_BLEND_FRAMES only: the motion-compensated front end of TemporalConsistency::process
     * (src/temporal_consistency.cpp:94-150) - Farneback flow prev->cur (:211-216), every buffered frame warped by
     * cv::remap with its own stored flow (:261-273) and weighted by that flow's reliability mask (:325-353).  Implies
     * scene_detect.  0 = the north-star blend of unwarped frames with unit masks. */
    int32_t use_flow;
    float motion_threshold;    /* Config::motion_threshold (15)                                   */
    double pyr_scale;          /* Config::pyr_scale (0.5)                                         */
    double poly_sigma;         /* Config::poly_sigma (1.2)                                        */
    int32_t levels;            /* Config::levels (3)                                              */
    int32_t winsize;           /* Config::winsize (15), odd                                       */
    int32_t iterations;        /* Config::iterations (3)                                          */
    int32_t poly_n;            /* Config::poly_n (5), <= 7                                        */
    int32_t flags;             /* Config::flags: only 0 is built (no initial flow, box window)    */
    int32_t reserved3_;
} vp_temporal_config;

/* The chain of Upscaler::upscale (src/upscaler.cpp:772-829) with step 2 = cv::resize(INTER_CUBIC)
 * (src/test_enhancements.cpp:346-349). */
typedef struct vp_chain_config {
    int32_t src_w, src_h;      /* input frame size                                               */
    int32_t dst_w, dst_h;      /* Upscaler::initialize(target_width, target_height)              */
    int32_t use_pre;           /* m_bilateral_pre   (src/upscaler.cpp:777-784)                   */
    int32_t use_sharpen;       /* m_sharpening      (:800-807)                                   */
    int32_t use_post;          /* m_bilateral_post  (:810-817)                                   */
    int32_t bicubic_enhance;   /* 1: the spatial part is CPUImpl's BICUBIC wrapper instead - GaussianBlur 3x3
                                  sigma 0.5 -> cv::resize(INTER_CUBIC) -> enhanceBicubicResult (bilateral 5/30/30,
                                  Canny 50/150, dilate 2x2, 5-point sharpen, select), src/upscaler.cpp:75-148;
                                  use_pre / use_sharpen / use_post are ignored                       */
    vp_bilateral_config pre;
    vp_sharpen_config sharpen;
    vp_bilateral_config post;
    vp_temporal_config temporal;
} vp_chain_config;

typedef struct vp_ctx vp_ctx;

/* ---- library ------------------------------------------------------------------------------- */
int vp_abi_version(void);
/* number of usable sm_100 devices (Upscaler::isGPUAvailable, src/upscaler.cpp:912-921) */
int vp_device_count(void);
/* fills `cfg` with the in-scope form of Upscaler::initializeEnhancements (src/upscaler.cpp:666-758) */
int vp_default_chain_config(vp_chain_config* cfg, int src_w, int src_h, int dst_w, int dst_h);

/* ---- context ------------------------------------------------------------------------------- */
/* Upscaler::initialize + initializeEnhancements (src/upscaler.cpp:569-758).  `cfg` may be NULL for
 * a context used only through the stage-wise entry points. */
int vp_create(vp_ctx** out, int device, const vp_chain_config* cfg);
void vp_destroy(vp_ctx* ctx);
const char* vp_last_error(const vp_ctx* ctx);   /* ctx may be NULL: last vp_create failure */
/* number of kernels this library launched on the context so far (bench.py's gpu_launches) */
uint64_t vp_launch_count(const vp_ctx* ctx);
int vp_synchronize(vp_ctx* ctx);                /* waits for every stream the context owns */
/* Per-kernel device timing for the roofline report: while enabled every kernel launch is bracketed by
 * CUDA events on its stream.  vp_profile_read sums them per kernel kind (0 stats, 1 bilateral PRE,
 * 2 bicubic+sharpen, 3 bilateral POST (+blend), 4 blend, 5 stage-wise bilateral) and clears the log. */
#define VP_PROFILE_KINDS 6
int vp_profile_enable(vp_ctx* ctx, int on);
int vp_profile_read(vp_ctx* ctx, double* total_ms, uint64_t* launches, int nkinds);

/* ---- the chain: Upscaler::upscale(const cv::Mat&, cv::Mat&) (src/upscaler.cpp:760-861) ------ */
/* Device frames.  d_dst == NULL runs the frame through everything except the temporal blend and
 * the output store: the warm-up frame of a multi-GPU segment (multi-GPU). */
int vp_process_device(vp_ctx* ctx, const uint8_t* d_src, size_t src_pitch,
                      uint8_t* d_dst, size_t dst_pitch, void* stream);
/* file / --fast mode: n device frames known ahead (src/main.cpp:627-629 `simulate_realtime=false`).
 * Consecutive frames cycle through up to four internal stream lanes with their own intermediates, so one
 * frame's reduction / prologue latency overlaps the others' kernels; the temporal stage still sees frames
 * in order (an event orders frame i's blend after frame i-1's).  Asynchronous on `stream`: the other
 * lanes fork from it and join it again, so work queued on `stream` before / after the call is ordered
 * before / after the whole batch.  d_dst[i] == NULL marks a warm-up frame; d_dst[i] and d_dst[i+1]
 * must not alias. */
int vp_set_batch_lanes(vp_ctx* ctx, int lanes);     /* 1..4 stream lanes for the batch call (default 4) */
/* CUDA-graph replay (default on): a batch of >= 8 frames whose arguments - frame pointers, pitches, stream, temporal ring
 * position - repeat is captured on its second sighting and afterwards launched as ONE instantiated graph (file mode cycles
 * through a fixed ring of frame buffers; src/main.cpp:222-355).  Only chains whose frames are plain kernel launches are
 * captured (the north-star chain and its bicubic-only / no-temporal forms); results are bit-identical either way.
 * vp_batch_graph_replays counts the batches that went out as a graph. */
int vp_set_batch_graphs(vp_ctx* ctx, int on);
uint64_t vp_batch_graph_replays(const vp_ctx* ctx);
int vp_process_batch_device(vp_ctx* ctx, int n, const uint8_t* const* d_src, size_t src_pitch,
                            uint8_t* const* d_dst, size_t dst_pitch, void* stream);
/* Host frames, synchronous (pinned staging inside when the buffers are not pinned). */
int vp_process_host(vp_ctx* ctx, const uint8_t* src, size_t src_step, uint8_t* dst, size_t dst_step);
/* Host frames, pipelined: up to vp_pipeline_depth() frames in flight; H2D and D2H have a stream each
 * and the frames in flight cycle through the compute lanes of the batch call, so copies overlap
 * compute and consecutive frames overlap each other while the temporal stage still sees them in
 * order (replaces FrameBuffer's deep copy, src/frame_buffer.cpp:31-37, and the per-op upload /
 * download of src/processor.cpp:98-124).  Device entry points called while host frames are in flight
 * are ordered behind them.  vp_wait returns the oldest submitted frame.  dst == NULL at submit =
 * warm-up frame (no output; vp_wait still pairs with it).  BOTH buffers belong to the library from
 * vp_submit_host until the matching vp_wait returns: a pinned `src` is read by DMA after the call returns
 * (do not overwrite, release or recycle it - e.g. the capture Mat - before vp_wait), `dst` is written until then.
 * A failed vp_wait leaves the frame in flight (call again or destroy the context); a failed vp_submit_host
 * queues nothing and clears the temporal history. */
int vp_pipeline_depth(const vp_ctx* ctx);
int vp_submit_host(vp_ctx* ctx, const uint8_t* src, size_t src_step, uint8_t* dst, size_t dst_step);
int vp_wait(vp_ctx* ctx);
/* ---- file / --fast mode over several devices (src/main.cpp:222-355 processing loop with simulate_realtime = false,
 * :627-629) ----------------------------------------------------------------------------------------------------
 * The clip's nframes frames are cut into ndev contiguous segments [floor(k n / G), floor((k+1) n / G)); device
 * devices[k] runs segment k on its own host thread and vp_ctx, frames pipelined through vp_submit_host / vp_wait,
 * after first feeding the temporal warm-up frames its first output depends on (0 without a temporal stage, 1 for
 * VP_TEMPORAL_MAIN_BLEND, buffer_size for VP_TEMPORAL_BLEND_FRAMES) with no output.  No data moves between devices.
 * Outputs are bit-identical to a one-device run of the same clip.  Returns the first failing segment's status;
 * vp_segments_last_error() describes it (per calling thread).  `report` (ndev entries) may be NULL. */
typedef struct vp_segment_report {
    int32_t device;
    int32_t first_frame;       /* frames [first_frame, start) are warm-up frames (no output)                       */
    int32_t start, stop;       /* frames [start, stop) are this segment's output                                   */
    int32_t status;            /* vp_status of the segment                                                         */
    int32_t reserved_;
    double seconds;            /* host wall time of the segment's frame loop (context creation excluded)           */
} vp_segment_report;
/* the partition itself (what vp_run_segments uses): frames [*first, *start) warm-up, [*start, *stop) output */
int vp_plan_segment(int nframes, int nsegments, int k, const vp_temporal_config* temporal, int* first, int* start, int* stop);
/* host frames given as pointer arrays (pinned memory makes the copies true DMA; pageable memory is staged) */
int vp_run_segments(const vp_chain_config* cfg, const int* devices, int ndev, int nframes,
                    const uint8_t* const* src, size_t src_step, uint8_t* const* dst, size_t dst_step,
                    vp_segment_report* report);
/* the same with a decoder / encoder on either side: `source` fills a pinned buffer with frame `frame`, `sink` receives
 * each output frame.  Both are called from the segment's own thread - concurrently across segments, in increasing frame
 * order within one - and return 0 on success (anything else aborts the run).  sink may be NULL. */
typedef int (*vp_frame_source)(void* user, int frame, uint8_t* buf, size_t step);
typedef int (*vp_frame_sink)(void* user, int frame, const uint8_t* buf, size_t step);
int vp_run_segments_cb(const vp_chain_config* cfg, const int* devices, int ndev, int nframes,
                       vp_frame_source source, vp_frame_sink sink, void* user, vp_segment_report* report);
const char* vp_segments_last_error(void);
/* TemporalConsistency::reset (src/temporal_consistency.cpp:52-59); also main.cpp's history clear */
int vp_reset_temporal(vp_ctx* ctx);
/* (d, sigma_color, sigma_space) calculateAdaptiveParams selected for the LAST processed frame;
 * which = VP_STAGE_PRE / VP_STAGE_POST.  Synchronises the context. */
int vp_get_stage_params(vp_ctx* ctx, int which, int* d, double* sigma_color, double* sigma_space);

/* pinned host memory for FrameBuffer slots / cv::Mat data (FrameBuffer, include/frame_buffer.h) */
int vp_host_alloc(void** p, size_t bytes);
int vp_host_free(void* p);

/* ---- stage-wise entry points (device pointers; used by the C++ stage classes and parity tests) */
/* cv::resize(src, dst, Size(dst_w,dst_h), 0, 0, INTER_CUBIC) - src/upscaler.cpp:81,
 * src/test_enhancements.cpp:348.  dst_w >= src_w and dst_h >= src_h. */
int vp_bicubic(vp_ctx* ctx, const uint8_t* d_src, size_t src_pitch, int src_w, int src_h,
               uint8_t* d_dst, size_t dst_pitch, int dst_w, int dst_h, void* stream);
/* cv::bilateralFilter(src, dst, d, sigma_color, sigma_space) - src/selective_bilateral.cpp:110 */
int vp_bilateral(vp_ctx* ctx, const uint8_t* d_src, size_t src_pitch, uint8_t* d_dst, size_t dst_pitch,
                 int w, int h, int d, double sigma_color, double sigma_space, void* stream);
/* SelectiveBilateral::process (src/selective_bilateral.cpp:57-82) in the mode `cfg` selects: plain (adaptive
 * parameters from the frame's own mean/std-dev, then the bilateral filter), selective or multi-scale. */
int vp_selective_bilateral(vp_ctx* ctx, const vp_bilateral_config* cfg, const uint8_t* d_src, size_t src_pitch,
                           uint8_t* d_dst, size_t dst_pitch, int w, int h, void* stream);
/* SelectiveBilateral::createDetailMask (src/selective_bilateral.cpp:239-313, repaired semantics): float32 mask,
 * pitch in bytes */
int vp_detail_mask(vp_ctx* ctx, double detail_threshold, const uint8_t* d_src, size_t src_pitch,
                   float* d_mask, size_t mask_pitch, int w, int h, void* stream);
/* cv::pyrDown / cv::pyrUp on CV_8UC3 (src/selective_bilateral.cpp:148-150, :188); pyrUp takes dst = 2*src or 2*src-1 */
int vp_pyr_down(vp_ctx* ctx, const uint8_t* d_src, size_t src_pitch, int src_w, int src_h,
                uint8_t* d_dst, size_t dst_pitch, void* stream);
int vp_pyr_up(vp_ctx* ctx, const uint8_t* d_src, size_t src_pitch, int src_w, int src_h,
              uint8_t* d_dst, size_t dst_pitch, int dst_w, int dst_h, void* stream);
/* cv::meanStdDev sums (src/selective_bilateral.cpp:321-322): d_sums[0..2] = sum x per channel,
 * d_sums[3..5] = sum x^2 per channel (uint64, device memory, overwritten). */
int vp_mean_stddev_sums(vp_ctx* ctx, const uint8_t* d_src, size_t src_pitch, int w, int h,
                        uint64_t* d_sums, void* stream);
/* AdaptiveSharpening::createEdgeMask (src/adaptive_sharpening.cpp:107-230): float32 mask, pitch in bytes */
int vp_edge_mask(vp_ctx* ctx, const vp_sharpen_config* cfg, const uint8_t* d_src, size_t src_pitch,
                 float* d_mask, size_t mask_pitch, int w, int h, void* stream);
/* AdaptiveSharpening::applyUnsharpMask (src/adaptive_sharpening.cpp:232-355) with a caller-supplied mask */
int vp_unsharp(vp_ctx* ctx, const vp_sharpen_config* cfg, const uint8_t* d_src, size_t src_pitch,
               const float* d_mask, size_t mask_pitch, uint8_t* d_dst, size_t dst_pitch, int w, int h, void* stream);
/* AdaptiveSharpening::process (src/adaptive_sharpening.cpp:51-105), either branch (cfg->adaptive_sigma) */
int vp_sharpen(vp_ctx* ctx, const vp_sharpen_config* cfg, const uint8_t* d_src, size_t src_pitch,
               uint8_t* d_dst, size_t dst_pitch, int w, int h, void* stream);
/* AdaptiveSharpening::calculateTextureMap + calculateAdaptiveSigma (src/adaptive_sharpening.cpp:357-440, repaired
 * semantics): the blurred float32 sigma map, pitch in bytes */
int vp_sigma_map(vp_ctx* ctx, const uint8_t* d_src, size_t src_pitch, float* d_sigma, size_t sigma_pitch,
                 int w, int h, void* stream);
/* cv::calcOpticalFlowFarneback(gray(prev), gray(cur), flow, ...) as TemporalConsistency::calculateOpticalFlow calls it
 * (src/temporal_consistency.cpp:211-216), parameters from `cfg` (pyr_scale .. flags).  The flow is returned as two
 * float32 planes (x and y displacement), pitch in bytes. */
int vp_optical_flow(vp_ctx* ctx, const vp_temporal_config* cfg, const uint8_t* d_prev, const uint8_t* d_cur, size_t pitch,
                    float* d_flow_x, float* d_flow_y, size_t flow_pitch, int w, int h, void* stream);
/* TemporalConsistency::warpFrame (src/temporal_consistency.cpp:261-273): cv::remap(frame, (x, y) + flow, INTER_LINEAR,
 * BORDER_REPLICATE) */
int vp_warp_flow(vp_ctx* ctx, const uint8_t* d_src, size_t src_pitch, const float* d_flow_x, const float* d_flow_y,
                 size_t flow_pitch, uint8_t* d_dst, size_t dst_pitch, int w, int h, void* stream);
/* TemporalConsistency::calculateFlowReliabilityMask (src/temporal_consistency.cpp:325-353): float32 mask, pitch in bytes */
int vp_flow_reliability(vp_ctx* ctx, const float* d_flow_x, const float* d_flow_y, size_t flow_pitch, float motion_threshold,
                        float* d_mask, size_t mask_pitch, int w, int h, void* stream);
/* cv::addWeighted(a, alpha, b, beta, 0, dst) on CV_8UC3 - src/main.cpp:286-290 */
int vp_add_weighted(vp_ctx* ctx, const uint8_t* d_a, size_t a_pitch, double alpha,
                    const uint8_t* d_b, size_t b_pitch, double beta,
                    uint8_t* d_dst, size_t dst_pitch, int w, int h, void* stream);
/* TemporalConsistency::blendFrames (src/temporal_consistency.cpp:355-444): frames[0] = current,
 * frames[i] = i-th previous; nframes in 1..4; all frames share `pitch`. */
int vp_temporal_blend(vp_ctx* ctx, const uint8_t* const* d_frames, int nframes, size_t pitch,
                      float blend_strength, uint8_t* d_dst, size_t dst_pitch, int w, int h, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VPB200_H */
