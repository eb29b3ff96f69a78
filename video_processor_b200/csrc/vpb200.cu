// vpb200.cu - context, stage scheduling and the C ABI (include/vpb200.h) of libvpb200.so.
//
// The chain is Upscaler::upscale's REAL_ESRGAN branch (reference src/upscaler.cpp:772-829) with the
// bicubic substitute of src/test_enhancements.cpp:346-349:
//     [k_stats] -> k_bilateral(PRE) -> k_upsharp(bicubic + AdaptiveSharpening, +stats of its output)
//               -> k_bilateral(POST, + fused temporal blend and state store)
// Kernel boundaries sit exactly where the reference needs a whole-frame reduction
// (calculateAdaptiveParams' cv::meanStdDev, src/selective_bilateral.cpp:321-322); the parameter
// choice itself happens on the device so no stage waits for the host.
#include "../../include/vpb200.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <cstddef>
#include <cstdio>
#include <cmath>
#include <cstring>
#include <deque>
#include <new>
#include <string>
#include <vector>

#include "vp_bicubic.cuh"
#include "vp_bicubic_int.cuh"
#include "vp_bilateral.cuh"
#include "vp_blend.cuh"
#include "vp_canny.cuh"
#include "vp_detail.cuh"
#include "vp_flow.cuh"
#include "vp_scene.cuh"
#include "vp_stats.cuh"
#include "vp_tables.h"
#include "vp_upsharp.cuh"
#include "vp_varsharp.cuh"

using namespace vp;

// layout the ctypes / FFI stubs rely on (tests/test_host_logic.py::test_struct_layout_matches_header)
static_assert(sizeof(vp_bilateral_config) == 64 && offsetof(vp_bilateral_config, detail_threshold) == 48, "vp_bilateral_config layout");
static_assert(sizeof(vp_temporal_config) == 72 && offsetof(vp_temporal_config, pyr_scale) == 32, "vp_temporal_config layout");
static_assert(sizeof(vp_sharpen_config) == 32, "vp_sharpen_config layout");
static_assert(offsetof(vp_chain_config, pre) == 32 && offsetof(vp_chain_config, post) == 128, "vp_chain_config layout");

namespace {

thread_local std::string g_create_error;

struct ResizeCache {
    int sw = 0, sh = 0, dw = 0, dh = 0;
    int* xofs = nullptr; int4* xcoef = nullptr; int* yofs = nullptr; float4* ycoef = nullptr;
    int max_rows[2] = {0, 0}, max_cols[2] = {0, 0};   // [0] halo 0, [1] halo US_HALO (k_upsharp tiles)
    int bc_rows = 0, bc_cols = 0;                     // k_bicubic tiles
    int yperiod = 0;                                  // 2 / 4 when the row tables are periodic with an integer ratio (k_bicubic<S>)
    int int_ratio = 0;                                // 2 / 4 when BOTH axes have that integer ratio and the tables are the
                                                      // periodic ones k_bicubic_int<S> assumes (verified entry by entry)
    float xk[4][4] = {}, yk[4][4] = {};               // its per-phase taps
};

struct HostSlot {
    uint8_t *d_in = nullptr, *d_out = nullptr;
    uint8_t *h_in = nullptr, *h_out = nullptr;        // pinned staging (allocated on first need)
    cudaEvent_t ev_h2d = nullptr, ev_comp = nullptr, ev_d2h = nullptr;
    uint8_t* user_dst = nullptr; size_t user_dst_step = 0;
    bool staged_out = false;
};

}  // namespace

struct MsScratch {                  // SelectiveBilateral's selective / multi-scale branches: one frame size's buffers
    int w = 0, h = 0, n = 0; bool p0 = false;
    uint8_t* lvl[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};    // pyrDown levels 1..n-1 (level 0 is the caller's frame)
    uint8_t* proc[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};   // filtered levels (proc[0] only when the caller's dst is not used)
    size_t pitch[5] = {0, 0, 0, 0, 0};
    int lw[5] = {0, 0, 0, 0, 0}, lh[5] = {0, 0, 0, 0, 0};
    float* mask = nullptr; size_t mpitch = 0;
    vp::DetailMax* mx = nullptr;
    uint32_t* dv = nullptr; int dvp = 0;   // k_detail_max's per-pixel plane (level-0 size; smaller levels reuse it)
};

struct FlowLevel {                  // one pyramid level of the Farneback engine
    int w = 0, h = 0, p = 0; size_t ps = 0;      // size, row pitch and plane stride in floats
    int ksize = 3; double sigma = 0;             // pre-smoothing of the full-resolution image for this level
    float* I = nullptr;                          // level image
    float* R[3] = {nullptr, nullptr, nullptr};   // polynomial expansions (5 planes) of the last 2 (stand-alone) / 3 (chain) frames
    float* flow = nullptr;                       // 2 planes
    vp::LinTap* px = nullptr; vp::LinTap* py = nullptr;   // cv::resize(INTER_LINEAR) taps, full resolution -> this level
    vp::LinTap* ux = nullptr; vp::LinTap* uy = nullptr;   // ... and the next coarser level's flow -> this level
};
struct FlowEngine {                 // cv::calcOpticalFlowFarneback for one frame size and parameter set
    int w = 0, h = 0, nlev = 0;     // nlev = levels + 1 pyramid entries, [0] = full resolution
    vp_temporal_config prm{};
    FlowLevel lv[8];
    float* gray = nullptr; float* blur_a = nullptr; float* blur_b = nullptr;   // full-resolution float planes
    float* tmp = nullptr;           // polyexp row moments (3 planes)
    float* M = nullptr;             // 5 planes
    vp::PolyTaps poly{};
};
struct FlowTemporal {               // TemporalConsistency::process with the flow front end: device state
    FlowEngine eng;
    int cur = 0;                    // which R[] holds the newest frame's expansions (ring of three: frame t+1 is expanded
                                    // on its own lane while the iterations between t-1 and t still read the other two)
    cudaEvent_t ev_expand = nullptr;                         // last flow_expand done (its scratch planes are shared)
    cudaEvent_t ev_solve[3] = {nullptr, nullptr, nullptr};   // [s]: the iterations whose newer frame sits in R[s] are done
    bool have_prev = false;
    uint8_t* raw_prev = nullptr; size_t rpitch = 0;          // the previous unblended input (:158)
    uint8_t* warped[3] = {nullptr, nullptr, nullptr};        // remap(frame t-1-i, its flow): fixed once computed
    float* mask[3] = {nullptr, nullptr, nullptr};            // that flow's reliability mask
    float* mtmp = nullptr;
    int head = 0, nvalid = 0;
};

struct VsScratch {                  // AdaptiveSharpening, adaptive_sigma branch: one frame size's buffers
    int w = 0, h = 0;
    float* emask = nullptr; float* sigma = nullptr; size_t fpitch = 0;   // edge mask / sigma map planes
    vp::SigmaStats* st = nullptr; vp::SigmaLevels* lv = nullptr;
    uint8_t* var = nullptr; size_t vpitch = 0;                           // local-variance plane (k_texture_minmax -> k_sigma_map)
    uint8_t* up = nullptr; size_t upitch = 0;                            // chain only: the bicubic output
};

struct Lane {                       // per-frame intermediates; the lanes let consecutive frames overlap
    cudaStream_t st = nullptr;      // lanes 1..3: their own stream (lane 0 runs on the caller's / the compute stream)
    uint8_t* d_pre = nullptr;       // PRE-filtered frame, src resolution
    uint8_t* d_sharp = nullptr;     // bicubic+sharpen output, dst resolution
    unsigned long long* d_sums = nullptr;   // [0..5] PRE input, [6..11] POST input, [12..17] stage-wise scratch
    int* d_sel = nullptr;           // [0] PRE class, [1] POST class, [2] stage-wise
    unsigned* d_counter = nullptr;  // CTA ticket counters of the three sel producers
    cudaEvent_t ev_done = nullptr;
    uint8_t* d_map = nullptr;       // Canny map (bicubic_enhance chains): 0 none, 1 weak, 2 edge
    unsigned char* d_dirty = nullptr;   // hysteresis tile flags (2 x ntiles) followed by the 2 `again` counters
    MsScratch ms_pre, ms_post;      // selective / multi-scale SelectiveBilateral stages of the chain
    uint8_t* d_post = nullptr;      // their POST output (the plain POST filter writes the caller's frame directly)
    VsScratch vs;                   // adaptive_sigma sharpening of the chain
};

struct vp_ctx {
    int device = 0;
    Lane lanes[4];
    // CUDA-graph replay of vp_process_batch_device: a batch whose arguments (frame pointers, pitches, temporal ring
    // position) were seen before is launched as one instantiated graph instead of 1-4 kernel launches per frame from the
    // host - file mode cycles through a fixed ring of frame buffers, and the bicubic-only chains are launch-bound otherwise
    struct BatchGraph { uint64_t key = 0; cudaGraphExec_t exec = nullptr; int head = 0, nvalid = 0, last_lane = 0; uint64_t launches = 0, used = 0; };
    std::vector<BatchGraph> graphs;
    std::vector<uint64_t> graph_seen;               // keys run eagerly once (tables warmed); the next sighting is captured
    bool graph_ok = true; uint64_t graph_clock = 0, graph_replays = 0;
    cudaEvent_t ev_cap_batch = nullptr, ev_cap_done[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t stage_ev = nullptr; cudaStream_t stage_last = nullptr; bool stage_any = false;   // StageOrder
    int nlanes = 2;                 // lanes vp_process_batch_device cycles through (vp_set_batch_lanes)
    cudaEvent_t ev_batch = nullptr;
    int last_lane = 0;
    // vp_submit_host cycles through the lanes as well: frames in flight overlap on the device, the temporal stages stay in
    // order through host_prev_done; the device entry points join the lanes before they touch the temporal state again
    int host_lane = 0; cudaEvent_t host_prev_done = nullptr;
    size_t map_pitch = 0; int hy_tiles_x = 0, hy_tiles_y = 0, hy_grid = 0; size_t dirty_bytes = 0;
    BilCand* d_cand_be = nullptr; int cand_be_rad[3] = {0, 0, 0};
    // scene-change detection (temporal.scene_detect)
    SceneAccum* d_scene_acc = nullptr; SceneState* d_scene_state = nullptr; int* d_use = nullptr; bool scene_reset = true;
    cudaStream_t s_compute = nullptr, s_h2d = nullptr, s_d2h = nullptr;
    bool has_chain = false;
    vp_chain_config cfg{};
    std::string err;
    uint64_t launches = 0;
    int sm_count = 148;

    // chain buffers
    size_t pre_pitch = 0;
    size_t frame_pitch = 0;                                 // pitch of dst-sized intermediates / states
    size_t sharp4_pitch = 0;                                // pitch of a lane's d_sharp when it holds 32-bit (B, G, R, 1) pixels
    size_t pre4_pitch = 0;                                  // the same for a lane's d_pre (PRE -> k_upsharp)
    bool lane_bgrx = false, pre_bgrx = false;                                 // plain chain on a driver with tensor maps: k_upsharp -> POST goes as words
    uint8_t* d_state[4] = {nullptr, nullptr, nullptr, nullptr};
    int nring = 0, head = 0, nvalid = 0;                    // temporal ring
    BilCand* d_cand_pre = nullptr; BilCand* d_cand_post = nullptr; BilCand* d_cand_tmp = nullptr;
    int cand_pre_rad[3] = {0, 0, 0}, cand_post_rad[3] = {0, 0, 0};
    // cached stage-wise tables
    int tmp_d = -1; double tmp_sc = 0, tmp_ss = 0; int tmp_r = 0;
    vp_bilateral_config tmp_sel_cfg{}; bool tmp_sel_valid = false; int tmp_sel_rad[3] = {0, 0, 0};
    BilCand* d_cand_sel = nullptr;
    // per-scale tables of the multi-scale branch: chain PRE / POST stages and the stage-wise entry point
    BilCand* d_cand_ms_pre = nullptr; BilCand* d_cand_ms_post = nullptr; BilCand* d_cand_ms_tmp = nullptr;
    int ms_pre_rad[5] = {0, 0, 0, 0, 0}, ms_post_rad[5] = {0, 0, 0, 0, 0}, ms_tmp_rad[5] = {0, 0, 0, 0, 0};
    FlowTemporal ftc;               // temporal.use_flow
    FlowEngine flow_tmp;            // stage-wise vp_optical_flow
    float* flow_tmp_mask = nullptr; size_t flow_tmp_mask_n = 0;
    int* d_mask_varies = nullptr;   // reliability mask: set when any raw value differs from 1.0f
    MsScratch ms_tmp;               // stage-wise scratch
    VsScratch vs_tmp;
    ResizeCache rc_chain, rc_tmp;
    float* d_siglut = nullptr; float siglut_thr = -1e30f; bool siglut_valid = false;
    float* d_siglut_tmp = nullptr; float siglut_tmp_thr = 0; bool siglut_tmp_valid = false;

    // optional per-kernel timing (bench.py's roofline leg)
    bool prof_on = false;
    struct ProfRec { int kind; cudaEvent_t e0, e1; };
    std::vector<ProfRec> prof_recs;
    std::vector<cudaEvent_t> prof_pool;

    // host pipeline
    std::vector<HostSlot> slots;
    std::deque<int> inflight;
    int next_slot = 0;
};

namespace {

int fail(vp_ctx* c, int code, const std::string& msg) {
    if (c) c->err = msg; else g_create_error = msg;
    return code;
}

enum { K_STATS = 0, K_BIL_PRE = 1, K_UPSHARP = 2, K_BIL_POST = 3, K_BLEND = 4, K_BIL_OTHER = 5, K_NKINDS = 6 };

// The stage-wise entry points take a stream per call but share per-context tables and scratch (candidate tables, resize
// tables, the sigmoid LUT, mask / pyramid / flow scratch).  Successive stage-wise calls on DIFFERENT streams of one context
// are therefore ordered on the device: a call first makes its stream wait for the previous call's last kernel, so a table
// re-upload or a scratch reuse can never overtake a kernel that still reads the old contents.  (Frees synchronise the
// device by themselves.)  Calls on one stream cost nothing extra.
struct StageOrder {
    vp_ctx* c; cudaStream_t st;
    StageOrder(vp_ctx* c_, cudaStream_t s);
    ~StageOrder();
};

struct ProfScope {   // brackets one launch with CUDA events on its stream when profiling is on
    vp_ctx* c; int kind; cudaStream_t st; cudaEvent_t e0 = nullptr, e1 = nullptr;
    static cudaEvent_t get(vp_ctx* c) {
        cudaEvent_t e = nullptr;
        if (!c->prof_pool.empty()) { e = c->prof_pool.back(); c->prof_pool.pop_back(); }
        else cudaEventCreate(&e);
        return e;
    }
    ProfScope(vp_ctx* c_, int k, cudaStream_t s) : c(c_), kind(k), st(s) {
        if (c->prof_on) { e0 = get(c); e1 = get(c); cudaEventRecord(e0, st); }
    }
    ~ProfScope() {
        if (e0) { cudaEventRecord(e1, st); c->prof_recs.push_back({kind, e0, e1}); }
    }
};

StageOrder::StageOrder(vp_ctx* c_, cudaStream_t s) : c(c_), st(s) {
    if (c->stage_any && c->stage_last != st) cudaStreamWaitEvent(st, c->stage_ev, 0);
}
StageOrder::~StageOrder() {
    if (!c->stage_ev && cudaEventCreateWithFlags(&c->stage_ev, cudaEventDisableTiming) != cudaSuccess) { c->stage_ev = nullptr; return; }
    if (cudaEventRecord(c->stage_ev, st) == cudaSuccess) { c->stage_last = st; c->stage_any = true; }
}

#define VP_CUDA(ctx, call)                                                                          \
    do {                                                                                            \
        cudaError_t e_ = (call);                                                                    \
        if (e_ != cudaSuccess)                                                                      \
            return fail((ctx), VP_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));    \
    } while (0)

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }


int check_frame(vp_ctx* c, const void* p, size_t pitch, int w, int h, const char* what) {
    if (!p) return fail(c, VP_ERR_INVALID_ARG, std::string(what) + ": null pointer");
    if (w <= 0 || h <= 0) return fail(c, VP_ERR_INVALID_ARG, std::string(what) + ": empty frame");
    if (pitch < (size_t)w * 3) return fail(c, VP_ERR_INVALID_ARG, std::string(what) + ": pitch < 3*width");
    return VP_OK;
}

// ---- table uploads ------------------------------------------------------------------------------
void fill_cand(BilCand& bc, const vp_host::BilateralTables& t) {
    std::memset(&bc, 0, sizeof(bc));
    bc.head.radius = t.radius;
    bc.head.d = t.d;
    for (int i = -t.radius; i <= t.radius; ++i) {
        int xm = 0;
        while ((xm + 1) * (xm + 1) + i * i <= t.radius * t.radius) ++xm;
        bc.head.xmax[i + t.radius] = xm;
    }
    std::memcpy(bc.head.space_w, t.space_w.data(), t.space_w.size() * sizeof(float));
    std::memcpy(bc.head.color_w, t.color_w.data(), 768 * sizeof(float));
    if (t.radius <= BIL_CLS_RMAX) {            // class tables: the tap weight space_w * color_w[n] formed here, once
        const int D = 2 * t.radius + 1;
        for (int i = -t.radius; i <= t.radius; ++i)
            for (int j = -t.radius; j <= t.radius; ++j) {
                const int cls = bil_class_of(i * i + j * j);
                if (i * i + j * j > t.radius * t.radius || cls < 0) continue;
                const volatile float sw = t.space_w[(size_t)(i + t.radius) * D + (j + t.radius)];
                for (int n = 0; n < 768; ++n) {
                    const volatile float prod = sw * bc.head.color_w[n];      // one IEEE single multiply, as the kernel's FMUL
                    bc.class_w[cls][n] = prod;
                }
            }
    }
}

int build_selective_cands(vp_ctx* c, const vp_bilateral_config& cfg, BilCand* d_cand, int radius[3], cudaStream_t st) {
    std::vector<BilCand> h(3);
    int rm = 0;
    for (int cls = 0; cls < 3; ++cls) {
        int d; double sc, ss;
        if (cfg.adaptive_params)
            vp_host::adaptive_params_for_class(cls, cfg.stage == VP_STAGE_POST, cfg.diameter, cfg.sigma_color, cfg.sigma_space, &d, &sc, &ss);
        else { d = cfg.diameter; sc = cfg.sigma_color; ss = cfg.sigma_space; }
        vp_host::BilateralTables t = vp_host::bilateral_tables(d, sc, ss);
        if (t.radius > BIL_RMAX)
            return fail(c, VP_ERR_UNSUPPORTED, "bilateral radius > 7 (diameter > 15) is not built");
        fill_cand(h[cls], t);
        radius[cls] = t.radius;
        rm = t.radius > rm ? t.radius : rm;
    }
    (void)rm;
    VP_CUDA(c, cudaMemcpyAsync(d_cand, h.data(), 3 * sizeof(BilCand), cudaMemcpyHostToDevice, st));
    VP_CUDA(c, cudaStreamSynchronize(st));
    return VP_OK;
}

void free_resize(ResizeCache& rc) {
    cudaFree(rc.xofs); cudaFree(rc.xcoef); cudaFree(rc.yofs); cudaFree(rc.ycoef);
    rc = ResizeCache();
}

int tile_span(const std::vector<int>& ofs, int dst_len, int tile, int halo) {
    int best = 0;
    for (int t0 = 0; t0 < dst_len; t0 += tile) {
        const int a = std::max(0, t0 - halo), b = std::min(dst_len - 1, t0 + tile - 1 + halo);
        best = std::max(best, ofs[b] + 3 - ofs[a] + 1);   // unclamped window (border columns are materialised)
    }
    return best;
}

int ensure_resize(vp_ctx* c, ResizeCache& rc, int sw, int sh, int dw, int dh, cudaStream_t st) {
    if (rc.sw == sw && rc.sh == sh && rc.dw == dw && rc.dh == dh && rc.xofs) return VP_OK;
    VP_CUDA(c, cudaStreamSynchronize(st));
    free_resize(rc);
    const vp_host::CubicTable tx = vp_host::cubic_table(sw, dw), ty = vp_host::cubic_table(sh, dh);
    std::vector<float> yb((size_t)dh * 4);
    for (size_t i = 0; i < yb.size(); ++i) yb[i] = (float)ty.coef[i] * (1.0f / 4194304.0f);
    VP_CUDA(c, cudaMalloc(&rc.xofs, (size_t)dw * sizeof(int)));
    VP_CUDA(c, cudaMalloc(&rc.xcoef, (size_t)dw * sizeof(int4)));
    std::vector<int> xc((size_t)dw * 4);
    for (size_t i = 0; i < xc.size(); ++i) xc[i] = tx.coef[i];
    VP_CUDA(c, cudaMalloc(&rc.yofs, (size_t)dh * sizeof(int)));
    VP_CUDA(c, cudaMalloc(&rc.ycoef, (size_t)dh * sizeof(float4)));
    VP_CUDA(c, cudaMemcpyAsync(rc.xofs, tx.ofs.data(), (size_t)dw * sizeof(int), cudaMemcpyHostToDevice, st));
    VP_CUDA(c, cudaMemcpyAsync(rc.xcoef, xc.data(), (size_t)dw * sizeof(int4), cudaMemcpyHostToDevice, st));
    VP_CUDA(c, cudaMemcpyAsync(rc.yofs, ty.ofs.data(), (size_t)dh * sizeof(int), cudaMemcpyHostToDevice, st));
    VP_CUDA(c, cudaMemcpyAsync(rc.ycoef, yb.data(), (size_t)dh * sizeof(float4), cudaMemcpyHostToDevice, st));
    VP_CUDA(c, cudaStreamSynchronize(st));
    for (int k = 0; k < 2; ++k) {
        const int halo = k ? US_HALO : 0;
        rc.max_cols[k] = tile_span(tx.ofs, dw, US_TW, halo);
        rc.max_rows[k] = tile_span(ty.ofs, dh, US_TH, halo);
    }
    rc.bc_cols = tile_span(tx.ofs, dw, BC_TW, 0);
    rc.bc_rows = tile_span(ty.ofs, dh, BC_TH, 0);
    // integer vertical ratio S: the tap origin advances exactly at y % S == S/2 and the taps repeat with period S
    rc.yperiod = 0;
    for (int S : {2, 4}) {
        if (dh != S * sh) continue;
        bool ok = true;
        for (int y = 0; y < dh && ok; ++y) {
            if (y + 1 < dh && ty.ofs[y + 1] - ty.ofs[y] != (((y + 1) % S) == S / 2 ? 1 : 0)) ok = false;
            if (y >= S && std::memcmp(&yb[(size_t)y * 4], &yb[(size_t)(y - S) * 4], 4 * sizeof(float)) != 0) ok = false;
        }
        if (ok) rc.yperiod = S;
    }
    // integer ratio S on both axes: tap origin of output x is x/S - 2 + (x % S >= S/2) and the taps repeat with period S
    rc.int_ratio = 0;
    for (int S : {2, 4}) {
        if (dw != S * sw || dh != S * sh || rc.yperiod != S) continue;
        bool ok = true;
        for (int x = 0; x < dw && ok; ++x) {
            if (tx.ofs[x] != x / S - 2 + ((x % S) >= S / 2 ? 1 : 0)) ok = false;
            if (x >= S && std::memcmp(&xc[(size_t)x * 4], &xc[(size_t)(x - S) * 4], 4 * sizeof(int)) != 0) ok = false;
            for (int t = 0; t < 4; ++t) if (std::abs(xc[(size_t)x * 4 + t]) > 4096) ok = false;     // exact in the float32 pass
        }
        for (int y = 0; y < dh && ok; ++y)
            if (ty.ofs[y] != y / S - 2 + ((y % S) >= S / 2 ? 1 : 0)) ok = false;
        if (!ok) continue;
        rc.int_ratio = S;
        for (int p = 0; p < S; ++p)
            for (int t = 0; t < 4; ++t) { rc.xk[p][t] = (float)xc[(size_t)p * 4 + t]; rc.yk[p][t] = yb[(size_t)p * 4 + t]; }
    }
    rc.sw = sw; rc.sh = sh; rc.dw = dw; rc.dh = dh;
    return VP_OK;
}

int ensure_siglut(vp_ctx* c, float** d_lut, float* cached_thr, bool* valid, float thr, cudaStream_t st) {
    if (*valid && *cached_thr == thr) return VP_OK;
    if (!*d_lut) VP_CUDA(c, cudaMalloc(d_lut, 256 * sizeof(float)));
    float lut[256];
    vp_host::sigmoid_lut(thr, lut);
    VP_CUDA(c, cudaMemcpyAsync(*d_lut, lut, sizeof(lut), cudaMemcpyHostToDevice, st));
    VP_CUDA(c, cudaStreamSynchronize(st));
    *cached_thr = thr; *valid = true;
    return VP_OK;
}

// ---- tensor maps (2-D TMA boxes) -------------------------------------------------------------------------------------
// cuTensorMapEncodeTiled is a host-side pure function of the driver; it is fetched through the runtime so the library
// carries no link-time dependency on libcuda (it must still load, and refuse, on a machine without a driver).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return (EncodeTiledFn)p;
    }();
    return fn;
}
// rows of `w_elems` elements (1, 2 or 4 bytes each), `h` rows `pitch` bytes apart (a multiple of 16, base 16-byte aligned)
bool make_tmap_2d(CUtensorMap* tm, int elem_bytes, const void* base, uint64_t w_elems, uint64_t h, uint64_t pitch, uint32_t box_w, uint32_t box_h) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn || ((uintptr_t)base & 15) || (pitch & 15) || box_w > 256 || box_h > 256) return false;
    const cuuint64_t gdim[2] = {w_elems, h}, gstride[1] = {pitch};
    const cuuint32_t box[2] = {box_w, box_h}, estr[2] = {1, 1};
    const CUtensorMapDataType dt = elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_UINT32 : elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_UINT16 : CU_TENSOR_MAP_DATA_TYPE_UINT8;
    return fn(tm, dt, 2, const_cast<void*>(base), gdim, gstride,
              box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// ---- launches -----------------------------------------------------------------------------------
int launch_stats(vp_ctx* c, const uint8_t* src, size_t pitch, int w, int h, unsigned long long* sums, const SelOut& fin, cudaStream_t st) {
    ProfScope ps(c, K_STATS, st);
    // every thread folds at most ~4000 16-pixel groups (uint32 partial sums): size the grid for that
    const long long groups48 = ((long long)3 * w * h + 47) / 48;
    long long grid = std::min<long long>((groups48 + 255) / 256, (long long)c->sm_count * 4);
    grid = std::max<long long>(grid, (groups48 + 256LL * 2000 - 1) / (256LL * 2000));
    grid = std::max<long long>(grid, 1);
    k_stats<<<(unsigned)grid, 256, 0, st>>>(src, pitch, w, h, sums, fin);
    VP_CUDA(c, cudaGetLastError());
    c->launches++;
    return VP_OK;
}

int launch_bilateral(vp_ctx* c, BilArgs& a, const int radius[3], cudaStream_t st, int kind = K_BIL_OTHER) {
    ProfScope ps(c, kind, st);
    int rmax = 0;
    for (int i = 0; i < 3; ++i) { a.radius[i] = radius[i]; rmax = std::max(rmax, radius[i]); }
    a.ns = rmax <= 2 ? 4 : (rmax <= 3 ? 2 : 1);     // small windows: taller tiles amortise the per-CTA table/halo cost
    if (const char* e = std::getenv("VPB200_BIL_NS")) {      // development: "a,b,c" = strips per thread for r <= 2, r == 3, r >= 4
        int v[3] = {4, 2, 1};
        if (std::sscanf(e, "%d,%d,%d", &v[0], &v[1], &v[2]) == 3) a.ns = std::min(8, std::max(1, rmax <= 2 ? v[0] : (rmax <= 3 ? v[1] : v[2])));
    }
    const dim3 grid((a.w + BIL_TW - 1) / BIL_TW, (a.h + BIL_TH * a.ns - 1) / (BIL_TH * a.ns));
    size_t smem = 0;                                // the device picks one of the candidates: size for the largest layout
    for (int i = 0; i < 3; ++i) smem = std::max(smem, bil_smem_bytes(radius[i], a.ns));
    a.use_tm = 1;                                   // one box per candidate radius
    for (int i = 0; i < 3 && a.use_tm; ++i) {
        if (i > 0 && radius[i] == radius[i - 1]) { a.tm[i] = a.tm[i - 1]; continue; }
        if (a.src_bgrx)                             // 32-bit pixels: the box is the tile itself
            a.use_tm = make_tmap_2d(&a.tm[i], 4, a.src, (uint64_t)a.w, a.h, a.spitch, (uint32_t)bil_tile_pitch_x(radius[i]),
                                    (uint32_t)(BIL_TH * a.ns + 2 * radius[i]));
        else
            a.use_tm = make_tmap_2d(&a.tm[i], 1, a.src, (uint64_t)a.w * 3, a.h, a.spitch, (uint32_t)bil_raw_pitch(radius[i]) - 16,
                                    (uint32_t)(BIL_TH * a.ns + 2 * radius[i]));
    }
    if (a.src_bgrx && !a.use_tm) return fail(c, VP_ERR_STATE, "internal: 32-bit lane buffer without tensor maps");
    if (a.dst_bgrx) {
        if (a.src_bgrx || a.blend.mode != 0 || a.blend.state || a.scene || a.emap || !a.dst || (a.dpitch & 15) || ((uintptr_t)a.dst & 15))
            return fail(c, VP_ERR_STATE, "internal: 32-bit bilateral output is the bare PRE stage only");
        k_bilateral<false, true><<<grid, BIL_NT, smem, st>>>(a);
    } else if (a.src_bgrx) k_bilateral<true><<<grid, BIL_NT, smem, st>>>(a);
    else k_bilateral<false><<<grid, BIL_NT, smem, st>>>(a);
    VP_CUDA(c, cudaGetLastError());
    c->launches++;
    return VP_OK;
}

// ---- SelectiveBilateral's selective / multi-scale branches (vp_detail.cuh) ------------------------------------
inline bool bil_has_modes(const vp_bilateral_config& b) { return b.use_multiscale || b.selective; }
inline int bil_num_scales(const vp_bilateral_config& b) { return std::max(1, std::min(b.num_scales, 5)); }   // initialize(), :40-43

void free_ms_scratch(MsScratch& S) {
    for (int i = 0; i < 5; ++i) { cudaFree(S.lvl[i]); cudaFree(S.proc[i]); }
    cudaFree(S.mask); cudaFree(S.mx); cudaFree(S.dv);
    S = MsScratch();
}

int ensure_ms_scratch(vp_ctx* c, MsScratch& S, int w, int h, int n, bool p0) {
    if (S.w == w && S.h == h && S.n == n && S.p0 == p0 && S.mask) return VP_OK;
    VP_CUDA(c, cudaDeviceSynchronize());
    free_ms_scratch(S);
    S.w = w; S.h = h; S.n = n; S.p0 = p0;
    int lw = w, lh = h;
    for (int i = 0; i < n; ++i) {
        S.lw[i] = lw; S.lh[i] = lh; S.pitch[i] = align_up((size_t)lw * 3, 256);
        if (i > 0) VP_CUDA(c, cudaMalloc(&S.lvl[i], S.pitch[i] * lh));
        if (i > 0 || p0) VP_CUDA(c, cudaMalloc(&S.proc[i], S.pitch[i] * lh));
        lw = (lw + 1) / 2; lh = (lh + 1) / 2;
    }
    S.mpitch = align_up((size_t)w * 4, 256);
    VP_CUDA(c, cudaMalloc(&S.mask, S.mpitch * h));
    VP_CUDA(c, cudaMalloc(&S.mx, sizeof(DetailMax)));
    S.dvp = (int)align_up((size_t)w, 64);
    VP_CUDA(c, cudaMalloc(&S.dv, (size_t)S.dvp * h * sizeof(uint32_t)));
    return VP_OK;
}

// one BilCand per scale: the config's diameter, sigmas x (1 + 0.5 i) (src/selective_bilateral.cpp:157-160)
int build_multiscale_cands(vp_ctx* c, const vp_bilateral_config& cfg, BilCand** d_cand, int rad[5], cudaStream_t st) {
    const int n = bil_num_scales(cfg);
    std::vector<BilCand> hc(5);
    for (int i = 0; i < n; ++i) {
        vp_host::BilateralTables t = vp_host::bilateral_tables(cfg.diameter, cfg.sigma_color * (1.0 + 0.5 * i), cfg.sigma_space * (1.0 + 0.5 * i));
        if (t.radius > BIL_RMAX) return fail(c, VP_ERR_UNSUPPORTED, "bilateral radius > 7 (diameter > 15) is not built");
        fill_cand(hc[i], t);
        rad[i] = t.radius;
    }
    if (!*d_cand) VP_CUDA(c, cudaMalloc(d_cand, 5 * sizeof(BilCand)));
    VP_CUDA(c, cudaMemcpyAsync(*d_cand, hc.data(), 5 * sizeof(BilCand), cudaMemcpyHostToDevice, st));
    VP_CUDA(c, cudaStreamSynchronize(st));
    return VP_OK;
}

int launch_detail_mask(vp_ctx* c, const uint8_t* src, size_t pitch, int w, int h, double detail_threshold,
                       DetailMax* mx, uint32_t* dv, int dvp, float* mask, size_t mpitch, cudaStream_t st) {
    ProfScope ps(c, K_BIL_OTHER, st);
    VP_CUDA(c, cudaMemsetAsync(mx, 0, sizeof(DetailMax), st));
    const dim3 grid((w + DT_TW - 1) / DT_TW, (h + DT_TH - 1) / DT_TH);
    k_detail_max<<<grid, DT_NT, 0, st>>>(src, pitch, w, h, mx, dv, dvp);
    DetailMaskArgs a{};
    a.dv = dv; a.dvp = dvp; a.w = w; a.h = h; a.mx = mx;
    a.thr = detail_threshold / 255.0;
    const std::vector<double> g = vp_host::gaussian_kernel_f64(5, 1.0);   // literal Size(5,5), 1.0 at :306
    for (int i = 0; i < 5; ++i) a.gf[i] = (float)g[i];
    a.mask = mask; a.mpitch = mpitch;
    k_detail_mask<<<grid, DT_NT, 0, st>>>(a);
    VP_CUDA(c, cudaGetLastError());
    c->launches += 2;
    return VP_OK;
}

struct StatsRes { unsigned long long* sums; unsigned* counter; int* sel; };   // calculateAdaptiveParams' reduction slots

// SelectiveBilateral::process in the selective or multi-scale mode (src/selective_bilateral.cpp:57-82).
// `cand3` / `rad3`: the plain filter's candidates (selective mode); `cand_ms` / `rad_ms`: one candidate per scale.
int run_bilateral_modes(vp_ctx* c, MsScratch& S, const vp_bilateral_config& cfg, const BilCand* cand3, const int rad3[3],
                        const BilCand* cand_ms, const int* rad_ms, const StatsRes& sr,
                        const uint8_t* src, size_t spitch, uint8_t* dst, size_t dpitch, int w, int h, cudaStream_t st, int kind) {
    const dim3 eb(256);
    int rc;
    if (cfg.use_multiscale) {
        const int n = bil_num_scales(cfg);
        rc = ensure_ms_scratch(c, S, w, h, n, false); if (rc) return rc;
        const uint8_t* lvl[5]; size_t lp[5]; uint8_t* proc[5]; size_t pp[5];
        lvl[0] = src; lp[0] = spitch; proc[0] = dst; pp[0] = dpitch;
        for (int i = 1; i < n; ++i) { lvl[i] = S.lvl[i]; lp[i] = S.pitch[i]; proc[i] = S.proc[i]; pp[i] = S.pitch[i]; }
        for (int i = 1; i < n; ++i) {                                            // :146-151
            ProfScope ps(c, K_BIL_OTHER, st);
            const dim3 grid(((S.lw[i] + 1) / 2 + 31) / 32, ((S.lh[i] + 1) / 2 + 7) / 8);           // a thread per 2x2 outputs
            k_pyr_down<<<grid, eb, 0, st>>>(lvl[i - 1], lp[i - 1], S.lw[i - 1], S.lh[i - 1], S.lvl[i], S.pitch[i], S.lw[i], S.lh[i]);
            VP_CUDA(c, cudaGetLastError());
            c->launches++;
        }
        for (int i = 0; i < n; ++i) {                                            // :155-182
            BilArgs a{};
            a.src = lvl[i]; a.spitch = lp[i]; a.dst = proc[i]; a.dpitch = pp[i]; a.w = S.lw[i]; a.h = S.lh[i];
            a.cand = cand_ms + i;
            const int rad[3] = {rad_ms[i], rad_ms[i], rad_ms[i]};
            rc = launch_bilateral(c, a, rad, st, i == 0 ? kind : K_BIL_OTHER); if (rc) return rc;
        }
        for (int i = n - 1; i > 0; --i) {                                        // :185-230
            rc = launch_detail_mask(c, proc[i - 1], pp[i - 1], S.lw[i - 1], S.lh[i - 1], cfg.detail_threshold, S.mx, S.dv, S.dvp, S.mask, S.mpitch, st);
            if (rc) return rc;
            MsBlendArgs b{proc[i - 1], pp[i - 1], S.lw[i - 1], S.lh[i - 1], proc[i], pp[i], S.lw[i], S.lh[i], S.mask, S.mpitch};
            ProfScope ps(c, K_BIL_OTHER, st);
            const dim3 grid(((S.lw[i - 1] + 1) / 2 + 31) / 32, ((S.lh[i - 1] + 1) / 2 + 7) / 8);    // a thread per 2x2 block
            const bool a2 = ((uintptr_t)b.fine | b.fpitch) % 2 == 0 && ((uintptr_t)b.mask | b.mpitch) % 8 == 0;
            if (a2) k_ms_blend<true><<<grid, eb, 0, st>>>(b);
            else k_ms_blend<false><<<grid, eb, 0, st>>>(b);
            VP_CUDA(c, cudaGetLastError());
            c->launches++;
        }
        return VP_OK;
    }
    // selective: detail mask of the input, the plain (adaptive) filter, the joint blend (:121-136, :373-439)
    rc = ensure_ms_scratch(c, S, w, h, 1, true); if (rc) return rc;
    rc = launch_detail_mask(c, src, spitch, w, h, cfg.detail_threshold, S.mx, S.dv, S.dvp, S.mask, S.mpitch, st); if (rc) return rc;
    if (cfg.adaptive_params) {
        VP_CUDA(c, cudaMemsetAsync(sr.sums, 0, 6 * sizeof(unsigned long long), st));
        rc = launch_stats(c, src, spitch, w, h, sr.sums, SelOut{sr.counter, sr.sel, (double)w * h}, st); if (rc) return rc;
    }
    BilArgs a{};
    a.src = src; a.spitch = spitch; a.dst = S.proc[0]; a.dpitch = S.pitch[0]; a.w = w; a.h = h;
    a.cand = cand3; a.sel = cfg.adaptive_params ? sr.sel : nullptr;
    rc = launch_bilateral(c, a, rad3, st, kind); if (rc) return rc;
    JointBlendArgs j{src, spitch, S.proc[0], S.pitch[0], S.mask, S.mpitch, dst, dpitch, w, h, cfg.edge_preserve - 1.0};
    ProfScope ps(c, K_BIL_OTHER, st);
    const dim3 grid((w + 31) / 32, (h + 7) / 8);
    k_joint_blend<<<grid, eb, 0, st>>>(j);
    VP_CUDA(c, cudaGetLastError());
    c->launches++;
    return VP_OK;
}

// ---- Farneback engine (vp_flow.cuh) ---------------------------------------------------------------------------
void free_flow_engine(FlowEngine& E) {
    for (FlowLevel& L : E.lv) { cudaFree(L.I); cudaFree(L.R[0]); cudaFree(L.R[1]); cudaFree(L.R[2]); cudaFree(L.flow); cudaFree(L.px); cudaFree(L.py); cudaFree(L.ux); cudaFree(L.uy); }
    cudaFree(E.gray); cudaFree(E.blur_a); cudaFree(E.blur_b); cudaFree(E.tmp); cudaFree(E.M);
    E = FlowEngine();
}
void free_flow_temporal(FlowTemporal& T) {
    free_flow_engine(T.eng);
    cudaFree(T.raw_prev); cudaFree(T.mtmp);
    for (int i = 0; i < 3; ++i) { cudaFree(T.warped[i]); cudaFree(T.mask[i]); if (T.ev_solve[i]) cudaEventDestroy(T.ev_solve[i]); }
    if (T.ev_expand) cudaEventDestroy(T.ev_expand);
    T = FlowTemporal();
}

// FarnebackPrepareGaussian (optflowgf.cpp): float taps g, x g, x^2 g and the four entries of inv(G) the expansion uses
int poly_taps(vp_ctx* c, int n, double sigma, PolyTaps& t) {
    if (n < 1 || n > 7) return fail(c, VP_ERR_UNSUPPORTED, "Farneback poly_n must be 1..7");
    if (sigma < 1.1920929e-07) sigma = n * 0.3;
    std::vector<float> g(2 * n + 1), xg(2 * n + 1), xxg(2 * n + 1);
    double s = 0.0;
    for (int x = -n; x <= n; ++x) { g[x + n] = (float)std::exp(-x * x / (2 * sigma * sigma)); s += g[x + n]; }
    s = 1.0 / s;
    for (int x = -n; x <= n; ++x) {
        g[x + n] = (float)(g[x + n] * s);
        xg[x + n] = (float)(x * g[x + n]);
        xxg[x + n] = (float)(x * x * g[x + n]);
    }
    double G[6][6] = {};
    for (int y = -n; y <= n; ++y)
        for (int x = -n; x <= n; ++x) {
            const float gg = g[y + n] * g[x + n];
            G[0][0] += gg;
            G[1][1] += gg * x * x;
            G[3][3] += gg * x * x * x * x;
            G[5][5] += gg * x * x * y * y;
        }
    G[2][2] = G[0][3] = G[0][4] = G[3][0] = G[4][0] = G[1][1];
    G[4][4] = G[3][3];
    G[3][4] = G[4][3] = G[5][5];
    double A[6][12];
    for (int i = 0; i < 6; ++i) for (int j = 0; j < 12; ++j) A[i][j] = j < 6 ? G[i][j] : (j - 6 == i ? 1.0 : 0.0);
    for (int i = 0; i < 6; ++i) {                      // Gauss-Jordan with partial pivoting (G is SPD, 6x6)
        int piv = i;
        for (int r = i + 1; r < 6; ++r) if (std::fabs(A[r][i]) > std::fabs(A[piv][i])) piv = r;
        if (std::fabs(A[piv][i]) < 1e-300) return fail(c, VP_ERR_UNSUPPORTED, "Farneback: singular moment matrix");
        if (piv != i) for (int j = 0; j < 12; ++j) std::swap(A[i][j], A[piv][j]);
        const double d = A[i][i];
        for (int j = 0; j < 12; ++j) A[i][j] /= d;
        for (int r = 0; r < 6; ++r) if (r != i) { const double f = A[r][i]; if (f != 0.0) for (int j = 0; j < 12; ++j) A[r][j] -= f * A[i][j]; }
    }
    t.n = n;
    for (int k = 0; k <= n; ++k) { t.g[k] = g[n + k]; t.xg[k] = xg[n + k]; t.xxg[k] = xxg[n + k]; }
    t.ig11 = A[1][7]; t.ig03 = A[0][9]; t.ig33 = A[3][9]; t.ig55 = A[5][11];
    return VP_OK;
}

bool same_flow_params(const vp_temporal_config& a, const vp_temporal_config& b) {
    return a.pyr_scale == b.pyr_scale && a.poly_sigma == b.poly_sigma && a.levels == b.levels && a.winsize == b.winsize &&
           a.iterations == b.iterations && a.poly_n == b.poly_n && a.flags == b.flags;
}

int ensure_flow_engine(vp_ctx* c, FlowEngine& E, int w, int h, const vp_temporal_config& prm, int nR = 2) {
    if (E.w == w && E.h == h && E.gray && same_flow_params(E.prm, prm)) return VP_OK;
    if (prm.flags != 0) return fail(c, VP_ERR_UNSUPPORTED, "Farneback flags other than 0 are not built");
    if (prm.winsize < 1 || !(prm.winsize & 1) || prm.winsize > 63) return fail(c, VP_ERR_UNSUPPORTED, "Farneback winsize must be odd, 1..63");
    VP_CUDA(c, cudaFuncSetAttribute(k_box_solve<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)box_smem_bytes(31)));
    VP_CUDA(c, cudaFuncSetAttribute(k_box_solve<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)box_smem_bytes(7, true)));
    if (prm.iterations < 1 || prm.levels < 0 || prm.levels > 7 || !(prm.pyr_scale > 0.0 && prm.pyr_scale < 1.0))
        return fail(c, VP_ERR_INVALID_ARG, "Farneback: iterations >= 1, levels 0..7, 0 < pyr_scale < 1");
    VP_CUDA(c, cudaDeviceSynchronize());
    free_flow_engine(E);
    int rc = poly_taps(c, prm.poly_n, prm.poly_sigma, E.poly); if (rc) return rc;
    E.w = w; E.h = h; E.prm = prm;
    // levels actually used (FarnebackOpticalFlowImpl::calc): stop before a level smaller than 32 pixels
    int levels = 0; double scale = 1.0;
    for (; levels < prm.levels; ++levels) {
        scale *= prm.pyr_scale;
        if (w * scale < 32 || h * scale < 32) break;
    }
    E.nlev = levels + 1;
    auto planes = [&](float** p, int n, const FlowLevel& L) { return cudaMalloc(p, (size_t)n * L.ps * sizeof(float)); };
    for (int k = 0; k < E.nlev; ++k) {
        FlowLevel& L = E.lv[k];
        scale = 1.0;
        for (int i = 0; i < k; ++i) scale *= prm.pyr_scale;
        L.sigma = (1.0 / scale - 1) * 0.5;
        L.ksize = std::max((int)std::lrint(L.sigma * 5) | 1, 3);
        if (L.ksize > 19) return fail(c, VP_ERR_UNSUPPORTED, "Farneback: pyramid pre-smoothing kernel > 19 (pyr_scale^levels too small)");
        L.w = (int)std::lrint(w * scale); L.h = (int)std::lrint(h * scale);
        L.p = (int)align_up((size_t)L.w, 64); L.ps = (size_t)L.p * L.h;
        VP_CUDA(c, planes(&L.I, 1, L)); VP_CUDA(c, planes(&L.flow, 2, L));
        for (int r = 0; r < nR; ++r) VP_CUDA(c, planes(&L.R[r], 5, L));
    }
    auto taps = [&](LinTap** t, int n, int sn) {          // tabulated on the device: the very arithmetic the kernels used inline
        cudaError_t e = cudaMalloc(t, (size_t)n * sizeof(LinTap));
        if (e != cudaSuccess) return e;
        k_linear_taps<<<(n + 255) / 256, 256>>>(n, (double)sn / (double)n, sn, *t);
        return cudaGetLastError();
    };
    for (int k = 0; k < E.nlev; ++k) {
        FlowLevel& L = E.lv[k];
        VP_CUDA(c, taps(&L.px, L.w, w)); VP_CUDA(c, taps(&L.py, L.h, h));
        if (k + 1 < E.nlev) { VP_CUDA(c, taps(&L.ux, L.w, E.lv[k + 1].w)); VP_CUDA(c, taps(&L.uy, L.h, E.lv[k + 1].h)); }
    }
    VP_CUDA(c, cudaDeviceSynchronize());
    const FlowLevel& L0 = E.lv[0];
    VP_CUDA(c, planes(&E.gray, 1, L0)); VP_CUDA(c, planes(&E.blur_a, 1, L0)); VP_CUDA(c, planes(&E.blur_b, 1, L0));
    VP_CUDA(c, planes(&E.tmp, 3, L0)); VP_CUDA(c, planes(&E.M, 5, L0));
    return VP_OK;
}

inline dim3 grid2d(int w, int h) { return dim3((w + 31) / 32, (h + 7) / 8); }

GaussTaps gauss_taps(int ksize, double sigma) {
    GaussTaps t{};
    t.ksize = ksize;
    const std::vector<double> g = vp_host::gaussian_kernel_f64(ksize, sigma);
    for (int i = 0; i < ksize; ++i) t.k[i] = (float)g[i];
    return t;
}

// gray -> per level: GaussianBlur(full resolution) -> resize(INTER_LINEAR) -> FarnebackPolyExp, into R[which]
int flow_expand(vp_ctx* c, FlowEngine& E, const uint8_t* bgr, size_t pitch, int which, cudaStream_t st) {
    ProfScope ps(c, K_BLEND, st);
    const FlowLevel& L0 = E.lv[0];
    k_gray_f32<<<grid2d(E.w, E.h), 256, 0, st>>>(bgr, pitch, E.gray, L0.p, E.w, E.h);
    c->launches++;
    for (int k = 0; k < E.nlev; ++k) {
        FlowLevel& L = E.lv[k];
        const GaussTaps t = gauss_taps(L.ksize, L.sigma);
        const float* img = L.I;
        if (L.w == E.w && L.h == E.h) {        // full resolution: the resize is the identity
            if (t.ksize == 3) {
                k_gauss3x3_f32<<<dim3((E.w + 31) / 32, (E.h + 31) / 32), 256, 0, st>>>(E.gray, L0.p, E.blur_b, L0.p, E.w, E.h, t);
            } else {
                k_gauss_f32<true><<<grid2d(E.w, E.h), 256, 0, st>>>(E.gray, L0.p, E.blur_a, L0.p, E.w, E.h, t);
                k_gauss_f32<false><<<grid2d(E.w, E.h), 256, 0, st>>>(E.blur_a, L0.p, E.blur_b, L0.p, E.w, E.h, t);
            }
            img = E.blur_b;
        } else {                               // blur only where the resize samples: row pass at 2 columns per output column ...
            const dim3 gr = grid2d(2 * L.w, E.h), gc = grid2d(L.w, L.h);      // ... column pass at 2 rows
#define VP_PYR(KS)                                                                                                  \
            do {                                                                                                    \
                k_gauss_rows_at<KS><<<gr, 256, 0, st>>>(E.gray, L0.p, E.w, E.h, E.blur_a, L0.p, L.w, t, L.px);      \
                k_gauss_cols_resize<KS><<<gc, 256, 0, st>>>(E.blur_a, L0.p, E.w, E.h, L.I, L.p, L.w, L.h, t, L.px, L.py);\
            } while (0)
            if (t.ksize == 3) VP_PYR(3); else if (t.ksize == 9) VP_PYR(9); else if (t.ksize == 19) VP_PYR(19); else VP_PYR(0);
#undef VP_PYR
        }
        const int ip = img == L.I ? L.p : L0.p;
        if (E.poly.n == 5) {
            k_polyexp_v<5><<<grid2d(L.w, L.h), 256, 0, st>>>(img, ip, E.tmp, L.p, L.ps, L.w, L.h, E.poly);
            k_polyexp_h<5><<<grid2d(L.w, L.h), 256, 0, st>>>(E.tmp, L.p, L.ps, L.R[which], L.p, L.ps, L.w, L.h, E.poly);
        } else {
            k_polyexp_v<0><<<grid2d(L.w, L.h), 256, 0, st>>>(img, ip, E.tmp, L.p, L.ps, L.w, L.h, E.poly);
            k_polyexp_h<0><<<grid2d(L.w, L.h), 256, 0, st>>>(E.tmp, L.p, L.ps, L.R[which], L.p, L.ps, L.w, L.h, E.poly);
        }
        c->launches += (img == E.blur_b && t.ksize == 3) ? 3 : 4;
    }
    VP_CUDA(c, cudaGetLastError());
    return VP_OK;
}

// coarse-to-fine Farneback iterations between R[prev] and R[cur]; the result is lv[0].flow
int flow_solve(vp_ctx* c, FlowEngine& E, int prev, int cur, cudaStream_t st) {
    ProfScope ps(c, K_BLEND, st);
    const int m = E.prm.winsize / 2;
    const double scale = 1.0 / ((double)E.prm.winsize * E.prm.winsize);
    for (int k = E.nlev - 1; k >= 0; --k) {
        FlowLevel& L = E.lv[k];
        const dim3 g = grid2d(L.w, L.h);
        if (k == E.nlev - 1) VP_CUDA(c, cudaMemsetAsync(L.flow, 0, 2 * L.ps * sizeof(float), st));
        else {
            const FlowLevel& U = E.lv[k + 1];
            k_resize_linear_f32<2><<<g, 256, 0, st>>>(U.flow, U.p, U.ps, U.w, U.h, L.flow, L.p, L.ps, L.w, L.h, (float)(1.0 / E.prm.pyr_scale), L.ux, L.uy);
            c->launches++;
        }
        k_update_matrices<<<g, 256, 0, st>>>(L.R[prev], L.R[cur], L.p, L.ps, L.flow, L.p, L.ps, E.M, L.p, L.ps, L.w, L.h);
        c->launches++;
        for (int i = 0; i < E.prm.iterations; ++i) {
            const dim3 gb((L.w + BOX_T - 1) / BOX_T, (L.h + BOX_T - 1) / BOX_T);
            if (m == 7) k_box_solve<7><<<gb, 256, box_smem_bytes(m, true), st>>>(E.M, L.p, L.ps, L.flow, L.p, L.ps, L.w, L.h, m, scale);
            else k_box_solve<0><<<gb, 256, box_smem_bytes(m), st>>>(E.M, L.p, L.ps, L.flow, L.p, L.ps, L.w, L.h, m, scale);
            c->launches++;
            if (i < E.prm.iterations - 1) {
                k_update_matrices<<<g, 256, 0, st>>>(L.R[prev], L.R[cur], L.p, L.ps, L.flow, L.p, L.ps, E.M, L.p, L.ps, L.w, L.h);
                c->launches++;
            }
        }
    }
    VP_CUDA(c, cudaGetLastError());
    return VP_OK;
}

// calculateFlowReliabilityMask: raw mask then cv::GaussianBlur(15x15, sigma 5) (:347); tmp = one float plane
int launch_reliability(vp_ctx* c, const float* fx, const float* fy, int fp, float thr, float* mask, int mp, float* tmp, int w, int h, cudaStream_t st) {
    ProfScope ps(c, K_BLEND, st);
    const GaussTaps t = gauss_taps(15, 5.0);
    // a frame without motion above the threshold has a mask of exactly 1.0f everywhere; its blur is then a constant - the
    // value the same float32 tap sequence gives on a constant plane - and k_gauss_f32 only stores it (flag set by k_reliability)
    if (!c->d_mask_varies) VP_CUDA(c, cudaMalloc(&c->d_mask_varies, sizeof(int)));
    VP_CUDA(c, cudaMemsetAsync(c->d_mask_varies, 0, sizeof(int), st));
    volatile float c1 = 0.f, c2 = 0.f;
    for (int i = 0; i < t.ksize; ++i) { volatile float p = 1.0f * t.k[i]; c1 = i == 0 ? p : c1 + p; }
    for (int i = 0; i < t.ksize; ++i) { volatile float p = c1 * t.k[i]; c2 = i == 0 ? p : c2 + p; }
    k_reliability<<<grid2d(w, h), 256, 0, st>>>(fx, fy, fp, mask, mp, w, h, thr, c->d_mask_varies);
    k_gauss_f32<true, 15><<<grid2d(w, h), 256, 0, st>>>(mask, mp, tmp, mp, w, h, t, c->d_mask_varies, c1);
    k_gauss_f32<false, 15><<<grid2d(w, h), 256, 0, st>>>(tmp, mp, mask, mp, w, h, t, c->d_mask_varies, c2);
    VP_CUDA(c, cudaGetLastError());
    c->launches += 3;
    return VP_OK;
}

int ensure_flow_temporal(vp_ctx* c, FlowTemporal& T, int w, int h, const vp_temporal_config& prm) {
    if (T.raw_prev && T.eng.w == w && T.eng.h == h && same_flow_params(T.eng.prm, prm)) return VP_OK;
    VP_CUDA(c, cudaDeviceSynchronize());
    free_flow_temporal(T);
    int rc = ensure_flow_engine(c, T.eng, w, h, prm, 3); if (rc) return rc;
    VP_CUDA(c, cudaEventCreateWithFlags(&T.ev_expand, cudaEventDisableTiming));
    for (int i = 0; i < 3; ++i) VP_CUDA(c, cudaEventCreateWithFlags(&T.ev_solve[i], cudaEventDisableTiming));
    T.rpitch = align_up((size_t)w * 3, 256);
    VP_CUDA(c, cudaMalloc(&T.raw_prev, T.rpitch * h));
    const FlowLevel& L0 = T.eng.lv[0];
    for (int i = 0; i < 3; ++i) {
        VP_CUDA(c, cudaMalloc(&T.warped[i], T.rpitch * h));
        VP_CUDA(c, cudaMalloc(&T.mask[i], L0.ps * sizeof(float)));
    }
    VP_CUDA(c, cudaMalloc(&T.mtmp, L0.ps * sizeof(float)));
    return VP_OK;
}

// TemporalConsistency::process (src/temporal_consistency.cpp:61-172) with the flow front end, all on the device.
// The warped copy of a buffered frame and its reliability mask depend only on that frame and on the flow computed when
// its successor arrived (:127-138), so both are computed once and kept in a ring instead of being redone every frame.
// `use_dev` is k_scene_decide's verdict for this frame (how many buffered frames may be used; 0 = pass through).
// Part 1, independent of the previous frame's blend: the new frame's pyramid and polynomial expansion.  It may run (on this
// frame's lane) while the previous frame's iterations are still going: it waits only for the previous expansion (shared
// scratch planes) and for the iterations that last read the R slot it overwrites.
int flow_temporal_expand(vp_ctx* c, FlowTemporal& T, const uint8_t* cur, size_t cur_pitch, cudaStream_t st) {
    const int nxt = (T.cur + 1) % 3;
    VP_CUDA(c, cudaStreamWaitEvent(st, T.ev_expand, 0));
    VP_CUDA(c, cudaStreamWaitEvent(st, T.ev_solve[(nxt + 1) % 3], 0));
    int rc = flow_expand(c, T.eng, cur, cur_pitch, nxt, st); if (rc) return rc;
    VP_CUDA(c, cudaEventRecord(T.ev_expand, st));
    return VP_OK;
}
// Part 2, in temporal order (after the previous frame is complete): iterations, warp, mask, blend.
int run_flow_temporal(vp_ctx* c, FlowTemporal& T, const vp_temporal_config& t, const uint8_t* cur, size_t cur_pitch,
                      uint8_t* dst, size_t dst_pitch, const int* use_dev, cudaStream_t st) {
    FlowEngine& E = T.eng;
    const int w = E.w, h = E.h;
    const FlowLevel& L0 = E.lv[0];
    const int nxt = (T.cur + 1) % 3;
    int rc = VP_OK;
    if (T.have_prev) {
        rc = flow_solve(c, E, T.cur, nxt, st); if (rc) return rc;                       // :95 flow prev -> cur
        VP_CUDA(c, cudaEventRecord(T.ev_solve[nxt], st));
        const int slot = T.head;
        {
            ProfScope ps(c, K_BLEND, st);
            k_warp<<<grid2d(w, h), 256, 0, st>>>(T.raw_prev, T.rpitch, L0.flow, L0.flow + L0.ps, L0.p, T.warped[slot], T.rpitch, w, h);   // :133
            VP_CUDA(c, cudaGetLastError());
            c->launches++;
        }
        rc = launch_reliability(c, L0.flow, L0.flow + L0.ps, L0.p, t.motion_threshold, T.mask[slot], L0.p, T.mtmp, w, h, st);             // :137
        if (rc) return rc;
        T.head = (T.head + 1) % 3;
        T.nvalid = std::min(T.nvalid + 1, std::min(std::max(t.buffer_size, 1), 3));
    }
    if (dst) {
        if (!T.have_prev) {
            VP_CUDA(c, cudaMemcpy2DAsync(dst, dst_pitch, cur, cur_pitch, (size_t)w * 3, h, cudaMemcpyDeviceToDevice, st));   // :79-86
        } else {
            MaskedBlendArgs a{};
            a.cur = cur; a.cpitch = cur_pitch; a.ppitch = T.rpitch; a.mp = L0.p;
            a.nprev = T.nvalid;
            for (int i = 0; i < T.nvalid; ++i) { const int s = (T.head - 1 - i + 6) % 3; a.prev[i] = T.warped[s]; a.mask[i] = T.mask[s]; }
            a.use_dev = use_dev;
            for (int i = 0; i < 4; ++i) { volatile float arg = -(float)i / 2.0f; a.fw[i] = (float)std::exp((double)arg); }
            a.blend_strength = t.blend_strength;
            a.dst = dst; a.dpitch = dst_pitch; a.w = w; a.h = h;
            ProfScope ps(c, K_BLEND, st);
            const bool words = w % 4 == 0 && ((uintptr_t)cur | cur_pitch | (uintptr_t)dst | dst_pitch | T.rpitch) % 4 == 0 &&
                               ((uintptr_t)T.mask[0] | (uintptr_t)T.mask[1] | (uintptr_t)T.mask[2]) % 16 == 0 && L0.p % 4 == 0;
            if (words) k_blend_masked4<<<grid2d(w / 4, h), 256, 0, st>>>(a);
            else k_blend_masked<<<grid2d(w, h), 256, 0, st>>>(a);
            VP_CUDA(c, cudaGetLastError());
            c->launches++;
        }
    }
    VP_CUDA(c, cudaMemcpy2DAsync(T.raw_prev, T.rpitch, cur, cur_pitch, (size_t)w * 3, h, cudaMemcpyDeviceToDevice, st));     // :158
    T.cur = nxt;
    T.have_prev = true;
    return VP_OK;
}

void free_vs_scratch(VsScratch& S) {
    cudaFree(S.emask); cudaFree(S.sigma); cudaFree(S.st); cudaFree(S.lv); cudaFree(S.up); cudaFree(S.var);
    S = VsScratch();
}

int ensure_vs_scratch(vp_ctx* c, VsScratch& S, int w, int h, bool with_up) {
    if (S.w == w && S.h == h && S.emask && (!with_up || S.up)) return VP_OK;
    VP_CUDA(c, cudaDeviceSynchronize());
    free_vs_scratch(S);
    S.w = w; S.h = h;
    S.fpitch = align_up((size_t)w * 4, 256);
    VP_CUDA(c, cudaMalloc(&S.emask, S.fpitch * h));
    VP_CUDA(c, cudaMalloc(&S.sigma, S.fpitch * h));
    VP_CUDA(c, cudaMalloc(&S.st, sizeof(SigmaStats)));
    VP_CUDA(c, cudaMalloc(&S.lv, sizeof(SigmaLevels)));
    S.vpitch = align_up((size_t)w, 256);
    VP_CUDA(c, cudaMalloc(&S.var, S.vpitch * h));
    if (with_up) {
        S.upitch = align_up((size_t)w * 3, 256);
        VP_CUDA(c, cudaMalloc(&S.up, S.upitch * h));
    }
    return VP_OK;
}

// calculateTextureMap + calculateAdaptiveSigma (src/adaptive_sharpening.cpp:357-440) -> sigma plane, and the extremes of
// that plane in S.st for k_sigma_levels
int launch_sigma_map(vp_ctx* c, SigmaStats* st_dev, uint8_t* var, size_t vpitch, const uint8_t* src, size_t pitch, int w, int h,
                     float* sigma, size_t spitch, cudaStream_t st) {
    static const SigmaStats init{0x7fffffff, 0, 0x7f800000, 0};
    ProfScope ps(c, K_UPSHARP, st);
    VP_CUDA(c, cudaMemcpyAsync(st_dev, &init, sizeof(init), cudaMemcpyHostToDevice, st));
    const dim3 grid((w + DT_TW - 1) / DT_TW, (h + DT_TH - 1) / DT_TH);
    k_texture_minmax<<<grid, DT_NT, 0, st>>>(src, pitch, w, h, st_dev, var, vpitch);
    SigmaMapArgs a{};
    a.var = var; a.vpitch = vpitch; a.w = w; a.h = h; a.st = st_dev;
    const std::vector<double> g = vp_host::gaussian_kernel_f64(5, 1.0);   // literal Size(5,5), 1.0 at :431
    for (int i = 0; i < 5; ++i) a.gf[i] = (float)g[i];
    a.sigma = sigma; a.spitch = spitch;
    k_sigma_map<<<grid, DT_NT, 0, st>>>(a);
    VP_CUDA(c, cudaGetLastError());
    c->launches += 2;
    return VP_OK;
}

int launch_upsharp(vp_ctx* c, UpsharpArgs& a, bool do_resize, bool do_sharpen, cudaStream_t st);
int fill_sharpen_args(vp_ctx* c, UpsharpArgs& a, const vp_sharpen_config& s);

// AdaptiveSharpening::process with adaptive_sigma = true (src/adaptive_sharpening.cpp:70-90) on a frame that is already
// at its final size: edge mask (k_upsharp's mask-only form), sigma map, levels, variable-sigma unsharp mask
int run_varsharp(vp_ctx* c, VsScratch& S, const vp_sharpen_config& cfg, const float* siglut, const uint8_t* src, size_t spitch,
                 uint8_t* dst, size_t dpitch, int w, int h, unsigned long long* sums, const SelOut& fin, cudaStream_t st) {
    if (w < 8 || h < 8) return fail(c, VP_ERR_UNSUPPORTED, "AdaptiveSharpening needs frames of at least 8x8");
    UpsharpArgs m{};
    int rc = fill_sharpen_args(c, m, cfg); if (rc) return rc;
    m.sig_lut = siglut;
    m.src = src; m.spitch = spitch; m.sw = w; m.sh = h; m.dst = nullptr; m.dpitch = 0; m.dw = w; m.dh = h;
    m.max_src_rows = US_TH + 2 * US_HALO; m.max_src_cols = US_TW + 2 * US_HALO;
    m.mask_out = S.emask; m.mask_out_pitch = S.fpitch;
    rc = launch_upsharp(c, m, false, true, st); if (rc) return rc;
    rc = launch_sigma_map(c, S.st, S.var, S.vpitch, src, spitch, w, h, S.sigma, S.fpitch, st); if (rc) return rc;
    ProfScope ps(c, K_UPSHARP, st);
    k_sigma_levels<<<1, 32, 0, st>>>(S.st, cfg.kernel_size, S.lv);
    VarSharpArgs a{};
    a.src = src; a.spitch = spitch; a.w = w; a.h = h;
    a.emask = S.emask; a.epitch = S.fpitch; a.sigma = S.sigma; a.gpitch = S.fpitch; a.lv = S.lv;
    a.gk = cfg.kernel_size;
    a.strength = cfg.strength; a.edge_strength = cfg.edge_strength; a.smooth_strength = cfg.smooth_strength;
    a.preserve_tone = cfg.preserve_tone;
    a.dst = dst; a.dpitch = dpitch; a.sums = sums; a.fin = fin;
    const dim3 grid((w + VS_TW - 1) / VS_TW, (h + VS_TH - 1) / VS_TH);
    if (a.gk == 3) k_varsharp<3><<<grid, VS_NT, 0, st>>>(a);
    else if (a.gk == 5) k_varsharp<5><<<grid, VS_NT, 0, st>>>(a);
    else k_varsharp<7><<<grid, VS_NT, 0, st>>>(a);
    VP_CUDA(c, cudaGetLastError());
    c->launches += 2;
    return VP_OK;
}

int fill_sharpen_args(vp_ctx* c, UpsharpArgs& a, const vp_sharpen_config& s) {
    if (s.kernel_size != 3 && s.kernel_size != 5 && s.kernel_size != 7)
        return fail(c, VP_ERR_UNSUPPORTED, "AdaptiveSharpening kernel_size must be 3, 5 or 7");
    a.strength = s.strength; a.edge_strength = s.edge_strength; a.smooth_strength = s.smooth_strength;
    a.preserve_tone = s.preserve_tone;
    a.gk = s.kernel_size;
    vp_host::gaussian_kernel_q8(s.kernel_size, (double)s.sigma, a.gq);
    const std::vector<double> g = vp_host::gaussian_kernel_f64(5, 1.5);   // literal Size(5,5), 1.5 at :222
    for (int i = 0; i < 5; ++i) a.gf[i] = (float)g[i];
    return VP_OK;
}

int launch_upsharp(vp_ctx* c, UpsharpArgs& a, bool do_resize, bool do_sharpen, cudaStream_t st) {
    if (do_sharpen && (a.dw < 8 || a.dh < 8))
        return fail(c, VP_ERR_UNSUPPORTED, "AdaptiveSharpening needs frames of at least 8x8");
    const UpsharpSmem L = upsharp_layout(do_resize, do_sharpen, a.max_src_rows, a.max_src_cols);
    if (L.total > 200 * 1024) return fail(c, VP_ERR_UNSUPPORTED, "resize ratio needs more shared memory than a CTA has");
    const dim3 grid((a.dw + US_TW - 1) / US_TW, (a.dh + US_TH - 1) / US_TH);
    if (a.src_bgrx) {
        a.src_pitch_x = (a.max_src_cols + 3 + 3) & ~3;          // up to 3 columns left of the window (box origin on 4 pixels)
        if (!do_resize || a.src_pitch_x * a.max_src_rows * 4 > L.raw_pitch * a.max_src_rows + L.src_pitch * a.max_src_rows * 4 ||
            !make_tmap_2d(&a.tm_src, 4, a.src, (uint64_t)a.sw, a.sh, a.spitch, (uint32_t)a.src_pitch_x, (uint32_t)a.max_src_rows))
            return fail(c, VP_ERR_STATE, "internal: 32-bit source window without tensor maps");
        a.use_tm_src = 1;
    } else
        a.use_tm_src = make_tmap_2d(&a.tm_src, 1, a.src, (uint64_t)a.sw * 3, a.sh, a.spitch, (uint32_t)L.raw_pitch - 16, (uint32_t)a.max_src_rows);
    if (a.out_bgrx) {
        a.use_tm_dst = make_tmap_2d(&a.tm_dst, 4, a.dst, (uint64_t)a.dw, a.dh, a.dpitch, US_TW, US_TH);
        if (!a.use_tm_dst) return fail(c, VP_ERR_STATE, "internal: 32-bit lane buffer without tensor maps");
    } else
        a.use_tm_dst = a.dst && make_tmap_2d(&a.tm_dst, 1, a.dst, (uint64_t)a.dw * 3, a.dh, a.dpitch, US_TW * 3, US_TH);
    ProfScope ps(c, K_UPSHARP, st);
    if (a.out_bgrx && !do_resize) return fail(c, VP_ERR_STATE, "internal: 32-bit lane output is built for the resizing forms only");
    if (a.out_bgrx) {
        if (!do_sharpen) k_upsharp<true, false, 5, true><<<grid, US_NT, L.total, st>>>(a);
        else if (a.gk == 3) k_upsharp<true, true, 3, true><<<grid, US_NT, L.total, st>>>(a);
        else if (a.gk == 5) k_upsharp<true, true, 5, true><<<grid, US_NT, L.total, st>>>(a);
        else k_upsharp<true, true, 7, true><<<grid, US_NT, L.total, st>>>(a);
    } else if (!do_sharpen) k_upsharp<true, false, 5><<<grid, US_NT, L.total, st>>>(a);
    else if (do_resize) {
        if (a.gk == 3) k_upsharp<true, true, 3><<<grid, US_NT, L.total, st>>>(a);
        else if (a.gk == 5) k_upsharp<true, true, 5><<<grid, US_NT, L.total, st>>>(a);
        else k_upsharp<true, true, 7><<<grid, US_NT, L.total, st>>>(a);
    } else {
        if (a.gk == 3) k_upsharp<false, true, 3><<<grid, US_NT, L.total, st>>>(a);
        else if (a.gk == 5) k_upsharp<false, true, 5><<<grid, US_NT, L.total, st>>>(a);
        else k_upsharp<false, true, 7><<<grid, US_NT, L.total, st>>>(a);
    }
    VP_CUDA(c, cudaGetLastError());
    c->launches++;
    return VP_OK;
}

int launch_bicubic(vp_ctx* c, const uint8_t* src, size_t spitch, int sw, int sh, uint8_t* dst, size_t dpitch, int dw, int dh,
                   const ResizeCache& rcache, cudaStream_t st) {
    if (rcache.int_ratio) {
        // integer ratio on both axes (BASELINE C1 / C5): the float4-window kernel, half the instructions per pixel
        BicubicIntArgs b{};
        b.src = src; b.spitch = spitch; b.sw = sw; b.sh = sh; b.dst = dst; b.dpitch = dpitch; b.dw = dw; b.dh = dh;
        for (int p = 0; p < 4; ++p)
            for (int t = 0; t < 4; ++t) { b.xk[p][t] = make_float2(rcache.xk[p][t], rcache.xk[p][t]); b.yk[p][t] = make_float2(rcache.yk[p][t], rcache.yk[p][t]); }
        if (rcache.int_ratio == 2)
            b.use_tm = make_tmap_2d(&b.tm_src, 1, src, (uint64_t)sw * 3, sh, spitch, BicubicIntGeom<2>::WB, BicubicIntGeom<2>::WR) &&
                       make_tmap_2d(&b.tm_dst, 2, dst, (uint64_t)dw * 3 / 2, dh, dpitch, BI_TW * 3 / 2, bi_th(2));
        else
            b.use_tm = make_tmap_2d(&b.tm_src, 1, src, (uint64_t)sw * 3, sh, spitch, BicubicIntGeom<4>::WB, BicubicIntGeom<4>::WR) &&
                       make_tmap_2d(&b.tm_dst, 2, dst, (uint64_t)dw * 3 / 2, dh, dpitch, BI_TW * 3 / 2, bi_th(4));
        const int ntx = (dw + BI_TW - 1) / BI_TW, ntiles = ntx * ((dh + bi_th(rcache.int_ratio) - 1) / bi_th(rcache.int_ratio));
        const int grid = (ntiles + BI_K - 1) / BI_K;
        ProfScope ps(c, K_UPSHARP, st);
        if (rcache.int_ratio == 2) k_bicubic_int<2><<<grid, BI_NT, BicubicIntGeom<2>::TOTAL, st>>>(b, ntx, ntiles);
        else k_bicubic_int<4><<<grid, BI_NT, BicubicIntGeom<4>::TOTAL, st>>>(b, ntx, ntiles);
        VP_CUDA(c, cudaGetLastError());
        c->launches++;
        return VP_OK;
    }
    BicubicArgs a{};
    a.src = src; a.spitch = spitch; a.sw = sw; a.sh = sh; a.dst = dst; a.dpitch = dpitch; a.dw = dw; a.dh = dh;
    a.rt = ResizeTables{rcache.xofs, rcache.xcoef, rcache.yofs, rcache.ycoef};
    a.max_src_rows = rcache.bc_rows; a.max_src_cols = rcache.bc_cols;
    const BicubicSmem L = bicubic_layout(a.max_src_rows, a.max_src_cols);
    if (L.total > 200 * 1024) return fail(c, VP_ERR_UNSUPPORTED, "resize ratio needs more shared memory than a CTA has");
    const dim3 grid((dw + BC_TW - 1) / BC_TW, (dh + BC_TH - 1) / BC_TH);
    ProfScope ps(c, K_UPSHARP, st);
    a.yperiod = rcache.yperiod;
    if (a.yperiod == 2) k_bicubic<2><<<grid, BC_NT, L.total, st>>>(a);
    else if (a.yperiod == 4) k_bicubic<4><<<grid, BC_NT, L.total, st>>>(a);
    else k_bicubic<0><<<grid, BC_NT, L.total, st>>>(a);
    VP_CUDA(c, cudaGetLastError());
    c->launches++;
    return VP_OK;
}

int launch_blend(vp_ctx* c, BlendKArgs& a, cudaStream_t st) {
    const long long work = ((long long)a.row_bytes * a.h + 15) / 16;
    const unsigned grid = (unsigned)std::min<long long>((work + 255) / 256, (long long)c->sm_count * 16);
    ProfScope ps(c, K_BLEND, st);
    k_blend<<<std::max(grid, 1u), 256, 0, st>>>(a);
    VP_CUDA(c, cudaGetLastError());
    c->launches++;
    return VP_OK;
}

void blend_weights(const vp_temporal_config& t, int nprev, BlendArgs& b) {
    b.mode = t.mode;
    b.nprev = nprev;
    if (t.mode == VP_TEMPORAL_MAIN_BLEND) {
        volatile float cw = t.main_weight;
        volatile float pw = 1.0f - cw;
        b.w[0] = cw; b.w[1] = pw; b.w[2] = 0.f; b.w[3] = 0.f; b.inv_wsum = 1.f;
    } else {
        // frame_weights[i] = std::exp(-(float)i / 2.0f) (float overload, correctly rounded here through double)
        volatile float wsum = 0.f;
        for (int i = 0; i < 4; ++i) {
            volatile float arg = -(float)i / 2.0f;
            volatile float fw = (float)std::exp((double)arg);
            volatile float bf = i == 0 ? 1.0f : t.blend_strength;
            volatile float m = fw * 1.0f;
            volatile float wt = m * bf;
            b.w[i] = wt;
            if (i <= nprev) wsum = wsum + wt;
        }
        volatile float inv = 1.0f / wsum;
        b.inv_wsum = inv;
    }
}

void destroy_slots(vp_ctx* c) {
    for (HostSlot& s : c->slots) {
        cudaFree(s.d_in); cudaFree(s.d_out);
        if (s.h_in) cudaFreeHost(s.h_in);
        if (s.h_out) cudaFreeHost(s.h_out);
        if (s.ev_h2d) cudaEventDestroy(s.ev_h2d);
        if (s.ev_comp) cudaEventDestroy(s.ev_comp);
        if (s.ev_d2h) cudaEventDestroy(s.ev_d2h);
    }
    c->slots.clear();
}

bool host_ptr_is_pinned(const void* p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost;
}

int alloc_enhance_buffers(vp_ctx* c, Lane& L) {
    const vp_chain_config& g = c->cfg;
    if (!g.bicubic_enhance || L.d_map) return VP_OK;
    if (!L.d_pre) VP_CUDA(c, cudaMalloc(&L.d_pre, std::max(c->pre_pitch, c->pre4_pitch) * g.src_h));
    VP_CUDA(c, cudaMalloc(&L.d_map, c->map_pitch * g.dst_h));
    VP_CUDA(c, cudaMalloc(&L.d_dirty, c->dirty_bytes));
    return VP_OK;
}

constexpr int PIPE_DEPTH = 3;

}  // namespace

// =================================================================================================
// C ABI
// =================================================================================================
extern "C" {

int vp_abi_version(void) { return VPB200_ABI_VERSION; }

int vp_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    int ok = 0;
    for (int i = 0; i < n; ++i) {
        int major = 0;
        if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, i) == cudaSuccess && major == 10) ok++;
    }
    return ok;
}

int vp_default_chain_config(vp_chain_config* cfg, int src_w, int src_h, int dst_w, int dst_h) {
    if (!cfg) return VP_ERR_INVALID_ARG;
    std::memset(cfg, 0, sizeof(*cfg));
    cfg->src_w = src_w; cfg->src_h = src_h; cfg->dst_w = dst_w; cfg->dst_h = dst_h;
    cfg->use_pre = cfg->use_sharpen = cfg->use_post = 1;
    cfg->pre = vp_bilateral_config{VP_STAGE_PRE, 1, 7, 0, 30.0, 30.0};          // src/upscaler.cpp:668-682
    cfg->sharpen = vp_sharpen_config{0.8f, 1.2f, 0.4f, 30.0f, 1.5f, 5, 1};      // :691-702
    cfg->post = vp_bilateral_config{VP_STAGE_POST, 1, 5, 0, 20.0, 20.0};        // :711-725
    // :734-748, main.cpp:282; scene_detect and the flow front end (use_flow) are opt-in, their parameters carry the reference's values
    cfg->temporal = vp_temporal_config{VP_TEMPORAL_BLEND_FRAMES, 3, 0.6f, 0.6f, 0, 100.0f, 0, 15.0f, 0.5, 1.2, 3, 15, 3, 5, 0, 0};
    return VP_OK;
}

const char* vp_last_error(const vp_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }
uint64_t vp_launch_count(const vp_ctx* ctx) { return ctx ? ctx->launches : 0; }

void vp_destroy(vp_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    destroy_slots(c);
    for (auto& r : c->prof_recs) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
    for (auto& e : c->prof_pool) cudaEventDestroy(e);
    for (Lane& L : c->lanes) {
        cudaFree(L.d_pre); cudaFree(L.d_sharp); cudaFree(L.d_sums); cudaFree(L.d_sel); cudaFree(L.d_counter);
        cudaFree(L.d_map); cudaFree(L.d_dirty); cudaFree(L.d_post);
        free_ms_scratch(L.ms_pre); free_ms_scratch(L.ms_post); free_vs_scratch(L.vs);
        if (L.ev_done) cudaEventDestroy(L.ev_done);
        if (L.st) cudaStreamDestroy(L.st);
    }
    if (c->ev_batch) cudaEventDestroy(c->ev_batch);
    if (c->stage_ev) cudaEventDestroy(c->stage_ev);
    for (auto& g : c->graphs) cudaGraphExecDestroy(g.exec);
    if (c->ev_cap_batch) cudaEventDestroy(c->ev_cap_batch);
    for (auto& e : c->ev_cap_done) if (e) cudaEventDestroy(e);
    for (auto& p : c->d_state) cudaFree(p);
    cudaFree(c->d_cand_be); cudaFree(c->d_scene_acc); cudaFree(c->d_scene_state); cudaFree(c->d_use);
    cudaFree(c->d_cand_pre); cudaFree(c->d_cand_post); cudaFree(c->d_cand_tmp); cudaFree(c->d_cand_sel);
    cudaFree(c->d_cand_ms_pre); cudaFree(c->d_cand_ms_post); cudaFree(c->d_cand_ms_tmp);
    free_flow_temporal(c->ftc); free_flow_engine(c->flow_tmp); cudaFree(c->flow_tmp_mask); cudaFree(c->d_mask_varies);
    free_ms_scratch(c->ms_tmp); free_vs_scratch(c->vs_tmp);
    free_resize(c->rc_chain); free_resize(c->rc_tmp);
    cudaFree(c->d_siglut); cudaFree(c->d_siglut_tmp);
    if (c->s_compute) cudaStreamDestroy(c->s_compute);
    if (c->s_h2d) cudaStreamDestroy(c->s_h2d);
    if (c->s_d2h) cudaStreamDestroy(c->s_d2h);
    cudaGetLastError();
    delete c;
}

static int create_impl(vp_ctx* c, const vp_chain_config* cfg) {
    VP_CUDA(c, cudaSetDevice(c->device));
    VP_CUDA(c, cudaStreamCreateWithFlags(&c->s_compute, cudaStreamNonBlocking));
    VP_CUDA(c, cudaStreamCreateWithFlags(&c->s_h2d, cudaStreamNonBlocking));
    VP_CUDA(c, cudaStreamCreateWithFlags(&c->s_d2h, cudaStreamNonBlocking));
    VP_CUDA(c, cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, c->device));
    {
        size_t bil_max = 0;                         // every (candidate radius, strips per thread) pair launch_bilateral can pick
        for (int rmax = 1; rmax <= BIL_RMAX; ++rmax)
            for (int r = 1; r <= rmax; ++r) bil_max = std::max(bil_max, bil_smem_bytes(r, std::getenv("VPB200_BIL_NS") ? 8 : (rmax <= 2 ? 4 : (rmax <= 3 ? 2 : 1))));
        VP_CUDA(c, cudaFuncSetAttribute(k_bilateral<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bil_max));
        VP_CUDA(c, cudaFuncSetAttribute(k_bilateral<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bil_max));
        VP_CUDA(c, cudaFuncSetAttribute(k_bilateral<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bil_max));
    }
    VP_CUDA(c, cudaFuncSetAttribute(k_bicubic<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    VP_CUDA(c, cudaFuncSetAttribute(k_bicubic<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    VP_CUDA(c, cudaFuncSetAttribute(k_bicubic<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    VP_CUDA(c, cudaFuncSetAttribute(k_bicubic_int<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, BicubicIntGeom<2>::TOTAL));
    VP_CUDA(c, cudaFuncSetAttribute(k_bicubic_int<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, BicubicIntGeom<4>::TOTAL));
    VP_CUDA(c, cudaFuncSetAttribute(k_upsharp<true, false, 5, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    VP_CUDA(c, cudaFuncSetAttribute(k_upsharp<true, true, 3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    VP_CUDA(c, cudaFuncSetAttribute(k_upsharp<true, true, 5, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    VP_CUDA(c, cudaFuncSetAttribute(k_upsharp<true, true, 7, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    VP_CUDA(c, cudaFuncSetAttribute(k_upsharp<true, false, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    VP_CUDA(c, cudaFuncSetAttribute(k_upsharp<true, true, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    VP_CUDA(c, cudaFuncSetAttribute(k_upsharp<true, true, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    VP_CUDA(c, cudaFuncSetAttribute(k_upsharp<true, true, 7>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    VP_CUDA(c, cudaFuncSetAttribute(k_upsharp<false, true, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    VP_CUDA(c, cudaFuncSetAttribute(k_upsharp<false, true, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    VP_CUDA(c, cudaFuncSetAttribute(k_upsharp<false, true, 7>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    VP_CUDA(c, cudaMalloc(&c->lanes[0].d_sums, 18 * sizeof(unsigned long long)));
    VP_CUDA(c, cudaMemset(c->lanes[0].d_sums, 0, 18 * sizeof(unsigned long long)));
    VP_CUDA(c, cudaMalloc(&c->lanes[0].d_sel, 4 * sizeof(int)));
    VP_CUDA(c, cudaMemset(c->lanes[0].d_sel, 0, 4 * sizeof(int)));
    VP_CUDA(c, cudaMalloc(&c->lanes[0].d_counter, 4 * sizeof(unsigned)));
    VP_CUDA(c, cudaMemset(c->lanes[0].d_counter, 0, 4 * sizeof(unsigned)));
    VP_CUDA(c, cudaMalloc(&c->d_cand_tmp, sizeof(BilCand)));
    VP_CUDA(c, cudaMalloc(&c->d_cand_sel, 3 * sizeof(BilCand)));
    if (!cfg) return VP_OK;

    c->cfg = *cfg;
    const vp_chain_config& g = c->cfg;
    if (g.src_w <= 0 || g.src_h <= 0 || g.dst_w <= 0 || g.dst_h <= 0)
        return fail(c, VP_ERR_INVALID_ARG, "vp_create: non-positive frame size");
    if (g.dst_w < g.src_w || g.dst_h < g.src_h)
        return fail(c, VP_ERR_UNSUPPORTED, "vp_create: the chain upscales (dst >= src); downscaling is not built");
    if (g.temporal.mode < 0 || g.temporal.mode > 2) return fail(c, VP_ERR_INVALID_ARG, "vp_create: temporal.mode");
    if (g.temporal.mode == VP_TEMPORAL_BLEND_FRAMES && (g.temporal.buffer_size < 1 || g.temporal.buffer_size > 3))
        return fail(c, VP_ERR_UNSUPPORTED, "vp_create: temporal.buffer_size must be 1..3");
    c->pre_pitch = align_up((size_t)g.src_w * 3, 256);
    c->pre4_pitch = align_up((size_t)g.src_w * 4, 256);
    c->frame_pitch = align_up((size_t)g.dst_w * 3, 256);
    if (g.use_pre && !g.bicubic_enhance) {
        VP_CUDA(c, cudaMalloc(&c->lanes[0].d_pre, std::max(c->pre_pitch, c->pre4_pitch) * g.src_h));
        VP_CUDA(c, cudaMalloc(&c->d_cand_pre, 3 * sizeof(BilCand)));
        int rc = build_selective_cands(c, g.pre, c->d_cand_pre, c->cand_pre_rad, c->s_compute);
        if (rc) return rc;
        if (g.pre.use_multiscale) { rc = build_multiscale_cands(c, g.pre, &c->d_cand_ms_pre, c->ms_pre_rad, c->s_compute); if (rc) return rc; }
    }
    if (g.use_post && !g.bicubic_enhance) {
        VP_CUDA(c, cudaMalloc(&c->d_cand_post, 3 * sizeof(BilCand)));
        int rc = build_selective_cands(c, g.post, c->d_cand_post, c->cand_post_rad, c->s_compute);
        if (rc) return rc;
        if (g.post.use_multiscale) { rc = build_multiscale_cands(c, g.post, &c->d_cand_ms_post, c->ms_post_rad, c->s_compute); if (rc) return rc; }
        if (bil_has_modes(g.post) || (g.temporal.mode == VP_TEMPORAL_BLEND_FRAMES && g.temporal.use_flow))
            VP_CUDA(c, cudaMalloc(&c->lanes[0].d_post, c->frame_pitch * g.dst_h));
    }
    c->sharp4_pitch = align_up((size_t)g.dst_w * 4, 256);
    // 32-bit pixels between k_upsharp and the plain POST bilateral (the POST halo tile then needs no expansion pass)
    c->lane_bgrx = encode_tiled_fn() != nullptr && g.use_post && !bil_has_modes(g.post) && !g.bicubic_enhance &&
                   g.dst_w > 2 * BIL_RMAX && g.dst_h > 2 * BIL_RMAX && !std::getenv("VPB200_NO_BGRX");
    c->pre_bgrx = encode_tiled_fn() != nullptr && g.use_pre && !bil_has_modes(g.pre) && !g.bicubic_enhance &&
                  g.src_w > 2 * BIL_RMAX && g.src_h > 2 * BIL_RMAX && !std::getenv("VPB200_NO_BGRX") && !std::getenv("VPB200_NO_PRE_BGRX");
    VP_CUDA(c, cudaMalloc(&c->lanes[0].d_sharp, std::max(c->frame_pitch, c->sharp4_pitch) * g.dst_h));
    c->nring = g.temporal.mode == VP_TEMPORAL_MAIN_BLEND ? 2
             : g.temporal.mode == VP_TEMPORAL_BLEND_FRAMES ? std::max(1, g.temporal.buffer_size) + 1 : 0;
    const bool flow_tc = g.temporal.mode == VP_TEMPORAL_BLEND_FRAMES && g.temporal.use_flow;
    if (flow_tc) {                       // the flow front end keeps its own state (FlowTemporal)
        c->nring = 1;
        int rc = ensure_flow_temporal(c, c->ftc, g.dst_w, g.dst_h, g.temporal); if (rc) return rc;
    } else {
        for (int i = 0; i < c->nring; ++i) VP_CUDA(c, cudaMalloc(&c->d_state[i], c->frame_pitch * g.dst_h));
    }
    if (g.dst_w != g.src_w || g.dst_h != g.src_h) {
        int rc = ensure_resize(c, c->rc_chain, g.src_w, g.src_h, g.dst_w, g.dst_h, c->s_compute);
        if (rc) return rc;
    }
    if (g.temporal.mode == VP_TEMPORAL_BLEND_FRAMES && (g.temporal.scene_detect || flow_tc)) {
        VP_CUDA(c, cudaMalloc(&c->d_scene_acc, sizeof(SceneAccum)));
        VP_CUDA(c, cudaMemset(c->d_scene_acc, 0, sizeof(SceneAccum)));
        VP_CUDA(c, cudaMalloc(&c->d_scene_state, sizeof(SceneState)));
        VP_CUDA(c, cudaMemset(c->d_scene_state, 0, sizeof(SceneState)));
        VP_CUDA(c, cudaMalloc(&c->d_use, sizeof(int)));
        VP_CUDA(c, cudaMemset(c->d_use, 0, sizeof(int)));
    }
    if (g.bicubic_enhance) {
        c->map_pitch = align_up((size_t)g.dst_w, 256);
        c->hy_tiles_x = (g.dst_w + HY_TW - 1) / HY_TW; c->hy_tiles_y = (g.dst_h + HY_TH - 1) / HY_TH;
        c->dirty_bytes = align_up((size_t)2 * c->hy_tiles_x * c->hy_tiles_y, 16) + 16;
        int per_sm = 0;
        VP_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_canny_hyst, HY_NT, 0));
        c->hy_grid = std::max(1, std::min(per_sm * c->sm_count, c->hy_tiles_x * c->hy_tiles_y));
        VP_CUDA(c, cudaMalloc(&c->d_cand_be, 3 * sizeof(BilCand)));
        vp_bilateral_config fixed{VP_STAGE_PRE, 0, 5, 0, 30.0, 30.0};      // cv::bilateralFilter(image, blurred, 5, 30, 30)
        int rc = build_selective_cands(c, fixed, c->d_cand_be, c->cand_be_rad, c->s_compute);
        if (rc) return rc;
        rc = alloc_enhance_buffers(c, c->lanes[0]); if (rc) return rc;
    }
    if (g.use_sharpen && !g.bicubic_enhance) {
        UpsharpArgs tmp{};
        int rc = fill_sharpen_args(c, tmp, g.sharpen);
        if (rc) return rc;
        rc = ensure_siglut(c, &c->d_siglut, &c->siglut_thr, &c->siglut_valid, g.sharpen.edge_threshold, c->s_compute);
        if (rc) return rc;
    }
    c->nlanes = 4;
    c->has_chain = true;
    return VP_OK;
}

int vp_create(vp_ctx** out, int device, const vp_chain_config* cfg) {
    if (!out) return fail(nullptr, VP_ERR_INVALID_ARG, "vp_create: out is null");
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(nullptr, VP_ERR_NO_DEVICE, "vp_create: no CUDA device (this library has no CPU fallback)");
    }
    if (device < 0 || device >= n) return fail(nullptr, VP_ERR_INVALID_ARG, "vp_create: device index out of range");
    int major = 0;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device);
    if (major != 10) return fail(nullptr, VP_ERR_NO_DEVICE, "vp_create: device is not sm_100 (kernels are built for sm_100a only)");
    vp_ctx* c = new (std::nothrow) vp_ctx();
    if (!c) return fail(nullptr, VP_ERR_NOMEM, "vp_create: out of host memory");
    c->device = device;
    const int rc = create_impl(c, cfg);
    if (rc != VP_OK) {
        g_create_error = c->err;
        vp_destroy(c);
        return rc;
    }
    *out = c;
    return VP_OK;
}

int vp_profile_enable(vp_ctx* c, int on) {
    if (!c) return VP_ERR_INVALID_ARG;
    c->prof_on = on != 0;
    return VP_OK;
}

int vp_profile_read(vp_ctx* c, double* total_ms, uint64_t* launches, int nkinds) {
    if (!c || !total_ms || !launches || nkinds < 1) return VP_ERR_INVALID_ARG;
    VP_CUDA(c, cudaSetDevice(c->device));
    for (int k = 0; k < nkinds; ++k) { total_ms[k] = 0.0; launches[k] = 0; }
    for (auto& r : c->prof_recs) {
        VP_CUDA(c, cudaEventSynchronize(r.e1));
        float ms = 0.f;
        VP_CUDA(c, cudaEventElapsedTime(&ms, r.e0, r.e1));
        if (r.kind < nkinds) { total_ms[r.kind] += ms; launches[r.kind]++; }
        c->prof_pool.push_back(r.e0); c->prof_pool.push_back(r.e1);
    }
    c->prof_recs.clear();
    return VP_OK;
}

int vp_synchronize(vp_ctx* c) {
    if (!c) return VP_ERR_INVALID_ARG;
    VP_CUDA(c, cudaSetDevice(c->device));
    VP_CUDA(c, cudaStreamSynchronize(c->s_h2d));
    for (int l = 1; l < 4; ++l)
        if (c->lanes[l].st) VP_CUDA(c, cudaStreamSynchronize(c->lanes[l].st));
    VP_CUDA(c, cudaStreamSynchronize(c->s_compute));
    VP_CUDA(c, cudaStreamSynchronize(c->s_d2h));
    return VP_OK;
}

int vp_reset_temporal(vp_ctx* c) {
    if (!c) return VP_ERR_INVALID_ARG;
    c->nvalid = 0;
    c->head = 0;
    c->scene_reset = true;
    c->ftc.have_prev = false; c->ftc.nvalid = 0; c->ftc.head = 0;
    return VP_OK;
}

// ---- the chain -----------------------------------------------------------------------------------
// temporal bookkeeping shared by both chain forms: which ring slots hold t-1, t-2 and where t is kept
static void temporal_args(vp_ctx* c, bool temporal, BlendArgs& bl) {
    const vp_chain_config& g = c->cfg;
    if (!temporal) return;
    // TemporalConsistency::process blends the current frame with EVERY buffered frame: m_frame_buffer holds the last
    // buffer_size inputs (:162-165) and m_flow_buffer is trimmed to buffer_size-1 only after the blend (:167-169), so with
    // the flow pushed at :97 the loop bound min(flows, frames) of :127 reaches buffer_size
    const int maxprev = g.temporal.mode == VP_TEMPORAL_MAIN_BLEND ? 1 : g.temporal.buffer_size;
    const int nprev = std::min(c->nvalid, maxprev);
    blend_weights(g.temporal, nprev, bl);
    for (int i = 0; i < nprev; ++i) bl.prev[i] = c->d_state[(c->head - 1 - i + 2 * c->nring) % c->nring];
    bl.ppitch = c->frame_pitch;
    bl.state = c->nring > 1 ? c->d_state[c->head] : nullptr;
    bl.spitch = c->frame_pitch;
}
static void temporal_advance(vp_ctx* c, bool temporal) {
    if (temporal && c->nring > 1) {
        c->head = (c->head + 1) % c->nring;
        c->nvalid = std::min(c->nvalid + 1, c->nring - 1);
    }
}

// CPUImpl's BICUBIC wrapper (src/upscaler.cpp:75-148) + the temporal stage:
//   k_gauss3 -> k_bicubic -> k_canny_nms -> k_canny_hyst (cooperative) -> k_bilateral(5,30,30) with the
//   edge-select epilogue and the fused temporal blend
static int process_frame_enhance(vp_ctx* c, Lane& L, cudaStream_t st, cudaEvent_t temporal_dep, bool temporal,
                                 const uint8_t* d_src, size_t src_pitch, uint8_t* d_dst, size_t dst_pitch) {
    const vp_chain_config& g = c->cfg;
    int q3[7];
    vp_host::gaussian_kernel_q8(3, 0.5, q3);
    {
        ProfScope ps(c, K_STATS, st);
        const long long px = (long long)g.src_w * g.src_h;
        const unsigned grid = (unsigned)std::min<long long>((px + 255) / 256, (long long)c->sm_count * 16);
        k_gauss3<<<grid, 256, 0, st>>>(d_src, src_pitch, L.d_pre, c->pre_pitch, g.src_w, g.src_h, q3[0], q3[1]);
        VP_CUDA(c, cudaGetLastError());
        c->launches++;
    }
    int rc;
    if (g.dst_w != g.src_w || g.dst_h != g.src_h) {
        rc = launch_bicubic(c, L.d_pre, c->pre_pitch, g.src_w, g.src_h, L.d_sharp, c->frame_pitch, g.dst_w, g.dst_h, c->rc_chain, st);
        if (rc) return rc;
    } else {
        VP_CUDA(c, cudaMemcpy2DAsync(L.d_sharp, c->frame_pitch, L.d_pre, c->pre_pitch, (size_t)g.src_w * 3, g.src_h, cudaMemcpyDeviceToDevice, st));
    }
    {
        ProfScope ps(c, K_BLEND, st);
        CannyArgs a{L.d_sharp, c->frame_pitch, g.dst_w, g.dst_h, L.d_map, c->map_pitch, 50, 150};
        const dim3 grid((g.dst_w + CN_TW - 1) / CN_TW, (g.dst_h + CN_TH - 1) / CN_TH);
        k_canny_nms<<<grid, CN_NT, 0, st>>>(a);
        VP_CUDA(c, cudaGetLastError());
        c->launches++;
        const int ntiles = c->hy_tiles_x * c->hy_tiles_y;
        VP_CUDA(c, cudaMemsetAsync(L.d_dirty, 1, ntiles, st));
        VP_CUDA(c, cudaMemsetAsync(L.d_dirty + ntiles, 0, c->dirty_bytes - ntiles, st));
        HystArgs hy{L.d_map, c->map_pitch, g.dst_w, g.dst_h, c->hy_tiles_x, c->hy_tiles_y, L.d_dirty,
                    reinterpret_cast<unsigned*>(L.d_dirty + c->dirty_bytes - 16)};
        void* params[] = {&hy};
        VP_CUDA(c, cudaLaunchCooperativeKernel((void*)k_canny_hyst, dim3(c->hy_grid), dim3(HY_NT), params, 0, st));
        c->launches++;
    }
    BlendArgs bl{};
    temporal_args(c, temporal, bl);
    if (temporal && temporal_dep) VP_CUDA(c, cudaStreamWaitEvent(st, temporal_dep, 0));
    BilArgs a{};
    a.src = L.d_sharp; a.spitch = c->frame_pitch; a.dst = d_dst; a.dpitch = dst_pitch; a.w = g.dst_w; a.h = g.dst_h;
    a.cand = c->d_cand_be; a.sel = nullptr;
    a.emap = L.d_map; a.epitch = c->map_pitch;
    a.blend = bl;
    rc = launch_bilateral(c, a, c->cand_be_rad, st, K_BIL_POST); if (rc) return rc;
    temporal_advance(c, temporal);
    return VP_OK;
}

// One frame through the chain on stream `st`, using lane `ln`'s intermediates.  `temporal_dep` (may be null)
// is waited on right before the kernel that reads the temporal state (the previous frame's last kernel when
// consecutive frames run on different streams).
static int process_frame(vp_ctx* c, int ln, cudaStream_t st, cudaEvent_t temporal_dep, const uint8_t* d_src, size_t src_pitch,
                         uint8_t* d_dst, size_t dst_pitch) {
    if (!c) return VP_ERR_INVALID_ARG;
    if (!c->has_chain) return fail(c, VP_ERR_STATE, "vp_process_device: context was created without a chain config");
    const vp_chain_config& g = c->cfg;
    int rc = check_frame(c, d_src, src_pitch, g.src_w, g.src_h, "vp_process_device(src)");
    if (rc) return rc;
    if (d_dst && dst_pitch < (size_t)g.dst_w * 3) return fail(c, VP_ERR_INVALID_ARG, "vp_process_device(dst): pitch < 3*width");
    Lane& L = c->lanes[ln];
    c->last_lane = ln;
    const bool temporal = g.temporal.mode != VP_TEMPORAL_NONE && c->nring > 0;
    if (!d_dst && !temporal) return VP_OK;   // warm-up frame of a chain without temporal state: nothing to do

    if (g.bicubic_enhance) return process_frame_enhance(c, L, st, temporal_dep, temporal, d_src, src_pitch, d_dst, dst_pitch);
    // selective / multi-scale SelectiveBilateral stages run as their own kernel sequence (run_bilateral_modes)
    const bool pre_modes = g.use_pre && bil_has_modes(g.pre), post_modes = g.use_post && bil_has_modes(g.post);
    const bool plain_post = g.use_post && !post_modes;
    const bool need_pre_stats = g.use_pre && g.pre.adaptive_params && !pre_modes;
    const bool need_post_stats = plain_post && g.post.adaptive_params;
    if (need_pre_stats || need_post_stats)
        VP_CUDA(c, cudaMemsetAsync(L.d_sums, 0, 12 * sizeof(unsigned long long), st));

    const uint8_t* cur = d_src; size_t cur_pitch = src_pitch;
    bool cur_bgrx = false;                    // `cur` holds 32-bit (B, G, R, 1) pixels (k_upsharp's out_bgrx form)
    bool pre_is_bgrx = false;                 // the PRE stage wrote 32-bit pixels for k_upsharp
    // 1. SelectiveBilateral PRE (src resolution)
    if (pre_modes) {
        rc = run_bilateral_modes(c, L.ms_pre, g.pre, c->d_cand_pre, c->cand_pre_rad, c->d_cand_ms_pre, c->ms_pre_rad,
                                 StatsRes{L.d_sums, L.d_counter, L.d_sel}, cur, cur_pitch, L.d_pre, c->pre_pitch, g.src_w, g.src_h, st, K_BIL_PRE);
        if (rc) return rc;
        cur = L.d_pre; cur_pitch = c->pre_pitch;
    } else if (g.use_pre) {
        if (need_pre_stats) {
            rc = launch_stats(c, cur, cur_pitch, g.src_w, g.src_h, L.d_sums, SelOut{L.d_counter, L.d_sel, (double)g.src_w * g.src_h}, st);
            if (rc) return rc;
        }
        BilArgs a{};
        a.src = cur; a.spitch = cur_pitch; a.dst = L.d_pre; a.dpitch = c->pre_pitch; a.w = g.src_w; a.h = g.src_h;
        a.cand = c->d_cand_pre; a.sel = need_pre_stats ? L.d_sel : nullptr;
        // the next stage is k_upsharp with a resize: hand it 32-bit pixels (its source window then needs no expansion pass)
        const bool next_is_upsharp = (g.dst_w != g.src_w || g.dst_h != g.src_h) && !(g.use_sharpen && g.sharpen.adaptive_sigma) &&
                                     (g.use_sharpen || need_post_stats);
        if (next_is_upsharp && c->pre_bgrx) { a.dst_bgrx = 1; a.dpitch = c->pre4_pitch; pre_is_bgrx = true; }
        rc = launch_bilateral(c, a, c->cand_pre_rad, st, K_BIL_PRE); if (rc) return rc;
        cur = L.d_pre; cur_pitch = a.dpitch;
    }
    // 2+3. bicubic resize + AdaptiveSharpening (dst resolution)
    const bool do_resize = (g.dst_w != g.src_w || g.dst_h != g.src_h);
    const bool last_is_upsharp = !g.use_post && !temporal;     // (a POST stage of either kind reads the lane buffer)
    if (g.use_sharpen && g.sharpen.adaptive_sigma) {
        // adaptive_sigma branch: bicubic on its own, then the four-kernel variable-sigma stage (run_varsharp)
        rc = ensure_vs_scratch(c, L.vs, g.dst_w, g.dst_h, do_resize); if (rc) return rc;
        if (do_resize) {
            rc = launch_bicubic(c, cur, cur_pitch, g.src_w, g.src_h, L.vs.up, L.vs.upitch, g.dst_w, g.dst_h, c->rc_chain, st);
            if (rc) return rc;
            cur = L.vs.up; cur_pitch = L.vs.upitch;
        }
        uint8_t* out = last_is_upsharp ? d_dst : L.d_sharp;
        const size_t opitch = last_is_upsharp ? dst_pitch : c->frame_pitch;
        rc = run_varsharp(c, L.vs, g.sharpen, c->d_siglut, cur, cur_pitch, out, opitch, g.dst_w, g.dst_h,
                          need_post_stats ? L.d_sums + 6 : nullptr,
                          need_post_stats ? SelOut{L.d_counter + 1, L.d_sel + 1, (double)g.dst_w * g.dst_h} : SelOut{nullptr, nullptr, 0.0}, st);
        if (rc) return rc;
        cur = out; cur_pitch = opitch;
    } else if (do_resize && !g.use_sharpen && !need_post_stats) {
        uint8_t* out = last_is_upsharp ? d_dst : L.d_sharp;
        const size_t opitch = last_is_upsharp ? dst_pitch : c->frame_pitch;
        rc = launch_bicubic(c, cur, cur_pitch, g.src_w, g.src_h, out, opitch, g.dst_w, g.dst_h, c->rc_chain, st);
        if (rc) return rc;
        cur = out; cur_pitch = opitch;
    } else if (do_resize || g.use_sharpen) {
        UpsharpArgs a{};
        a.src = cur; a.spitch = cur_pitch; a.sw = g.src_w; a.sh = g.src_h;
        a.src_bgrx = pre_is_bgrx ? 1 : 0;
        a.dst = last_is_upsharp ? d_dst : L.d_sharp; a.dpitch = last_is_upsharp ? dst_pitch : c->frame_pitch;
        a.dw = g.dst_w; a.dh = g.dst_h;
        if (!last_is_upsharp && plain_post && c->lane_bgrx && do_resize) { a.out_bgrx = 1; a.dpitch = c->sharp4_pitch; cur_bgrx = true; }
        const int k = g.use_sharpen ? 1 : 0, halo = g.use_sharpen ? US_HALO : 0;
        if (do_resize) {
            a.rt = ResizeTables{c->rc_chain.xofs, c->rc_chain.xcoef, c->rc_chain.yofs, c->rc_chain.ycoef};
            a.max_src_rows = c->rc_chain.max_rows[k]; a.max_src_cols = c->rc_chain.max_cols[k];
            a.yperiod = c->rc_chain.yperiod;
        } else {
            a.max_src_rows = US_TH + 2 * halo; a.max_src_cols = US_TW + 2 * halo;
        }
        if (g.use_sharpen) { rc = fill_sharpen_args(c, a, g.sharpen); if (rc) return rc; a.sig_lut = c->d_siglut; }
        a.sums = need_post_stats ? L.d_sums + 6 : nullptr;
        if (need_post_stats) a.fin = SelOut{L.d_counter + 1, L.d_sel + 1, (double)g.dst_w * g.dst_h};
        rc = launch_upsharp(c, a, do_resize, g.use_sharpen != 0, st); if (rc) return rc;
        cur = a.dst; cur_pitch = a.dpitch;
    } else if (need_post_stats) {
        rc = launch_stats(c, cur, cur_pitch, g.dst_w, g.dst_h, L.d_sums + 6, SelOut{L.d_counter + 1, L.d_sel + 1, (double)g.dst_w * g.dst_h}, st);
        if (rc) return rc;
    }
    if (post_modes) {
        // the temporal stage (if any) then runs as the stand-alone blend, exactly as for a chain without POST
        uint8_t* out = temporal ? L.d_post : d_dst;
        const size_t opitch = temporal ? c->frame_pitch : dst_pitch;
        rc = run_bilateral_modes(c, L.ms_post, g.post, c->d_cand_post, c->cand_post_rad, c->d_cand_ms_post, c->ms_post_rad,
                                 StatsRes{L.d_sums + 6, L.d_counter + 1, L.d_sel + 1}, cur, cur_pitch, out, opitch, g.dst_w, g.dst_h, st, K_BIL_POST);
        if (rc) return rc;
        cur = out; cur_pitch = opitch;
    }
    if (temporal && g.temporal.mode == VP_TEMPORAL_BLEND_FRAMES && g.temporal.use_flow) {
        // TemporalConsistency::process with its flow front end (run_flow_temporal): the unblended frame has to exist in
        // memory first, so the plain POST filter runs without its blend epilogue
        if (plain_post) {
            BilArgs a{};
            a.src = cur; a.spitch = cur_pitch; a.dst = L.d_post; a.dpitch = c->frame_pitch; a.w = g.dst_w; a.h = g.dst_h;
            a.src_bgrx = cur_bgrx ? 1 : 0;
            a.cand = c->d_cand_post; a.sel = need_post_stats ? L.d_sel + 1 : nullptr;
            rc = launch_bilateral(c, a, c->cand_post_rad, st, K_BIL_POST); if (rc) return rc;
            cur = L.d_post; cur_pitch = c->frame_pitch;
        }
        rc = flow_temporal_expand(c, c->ftc, cur, cur_pitch, st); if (rc) return rc;
        if (temporal_dep) VP_CUDA(c, cudaStreamWaitEvent(st, temporal_dep, 0));
        {
            SceneAccumArgs sa{cur, cur_pitch, c->ftc.have_prev ? c->ftc.raw_prev : nullptr, c->ftc.rpitch, g.dst_w, g.dst_h, c->d_scene_acc};
            ProfScope ps(c, K_BLEND, st);
            k_scene_accum<<<c->sm_count * 8, 256, 0, st>>>(sa);
            k_scene_decide<<<1, 64, 0, st>>>(c->d_scene_acc, c->d_scene_state, c->d_use, (double)g.dst_w * g.dst_h,
                                             g.temporal.scene_change_threshold, g.temporal.buffer_size, (c->scene_reset || !c->ftc.have_prev) ? 1 : 0);
            VP_CUDA(c, cudaGetLastError());
            c->launches += 2;
            c->scene_reset = false;
        }
        return run_flow_temporal(c, c->ftc, g.temporal, cur, cur_pitch, d_dst, dst_pitch, c->d_use, st);
    }
    BlendArgs bl{};
    temporal_args(c, temporal, bl);
    if (temporal && temporal_dep) VP_CUDA(c, cudaStreamWaitEvent(st, temporal_dep, 0));
    const bool scene = temporal && c->d_scene_acc != nullptr;
    if (scene) {
        // detectSceneChange needs the whole unblended frame before any pixel may be blended: the producer
        // accumulates the histogram / gray difference, k_scene_decide rules, k_blend blends (or passes through)
        const uint8_t* blend_cur = cur; size_t blend_cur_pitch = cur_pitch;
        if (plain_post) {
            BilArgs a{};
            a.src = cur; a.spitch = cur_pitch; a.dst = nullptr; a.dpitch = 0; a.w = g.dst_w; a.h = g.dst_h;
            a.src_bgrx = cur_bgrx ? 1 : 0;
            a.cand = c->d_cand_post; a.sel = need_post_stats ? L.d_sel + 1 : nullptr;
            a.blend = bl;                         // state store + previous frame for the gray difference
            a.scene = c->d_scene_acc;
            rc = launch_bilateral(c, a, c->cand_post_rad, st, K_BIL_POST); if (rc) return rc;
            blend_cur = bl.state; blend_cur_pitch = c->frame_pitch;
        } else {
            if (cur == d_src && (g.dst_w != g.src_w || g.dst_h != g.src_h))
                return fail(c, VP_ERR_STATE, "internal: chain produced no frame");
            SceneAccumArgs sa{cur, cur_pitch, bl.nprev > 0 ? bl.prev[0] : nullptr, c->frame_pitch, g.dst_w, g.dst_h, c->d_scene_acc};
            ProfScope ps(c, K_BLEND, st);
            k_scene_accum<<<c->sm_count * 8, 256, 0, st>>>(sa);
            VP_CUDA(c, cudaGetLastError());
            c->launches++;
        }
        k_scene_decide<<<1, 64, 0, st>>>(c->d_scene_acc, c->d_scene_state, c->d_use, (double)g.dst_w * g.dst_h,
                                         g.temporal.scene_change_threshold, g.temporal.buffer_size, c->scene_reset ? 1 : 0);
        VP_CUDA(c, cudaGetLastError());
        c->launches++;
        c->scene_reset = false;
        if (d_dst || !plain_post) {
            BlendKArgs a{};
            a.cur = blend_cur; a.cpitch = blend_cur_pitch; a.dst = d_dst; a.dpitch = dst_pitch; a.row_bytes = g.dst_w * 3; a.h = g.dst_h;
            a.blend = bl;
            if (plain_post) a.blend.state = nullptr;      // POST already stored the unblended frame
            a.use_dev = c->d_use;
            for (int n = 1; n <= 3; ++n) { BlendArgs t{}; blend_weights(g.temporal, n, t); a.inv_wsum_n[n - 1] = t.inv_wsum; }
            rc = launch_blend(c, a, st); if (rc) return rc;
        }
    } else
    // 4+5. SelectiveBilateral POST with the temporal blend in its epilogue
    if (plain_post) {
        BilArgs a{};
        a.src = cur; a.spitch = cur_pitch; a.dst = d_dst; a.dpitch = dst_pitch; a.w = g.dst_w; a.h = g.dst_h;
        a.src_bgrx = cur_bgrx ? 1 : 0;
        a.cand = c->d_cand_post; a.sel = need_post_stats ? L.d_sel + 1 : nullptr;
        a.blend = bl;
        rc = launch_bilateral(c, a, c->cand_post_rad, st, K_BIL_POST); if (rc) return rc;
    } else if (temporal) {
        BlendKArgs a{};
        a.cur = cur; a.cpitch = cur_pitch; a.dst = d_dst; a.dpitch = dst_pitch; a.row_bytes = g.dst_w * 3; a.h = g.dst_h;
        a.blend = bl;
        rc = launch_blend(c, a, st); if (rc) return rc;
    } else if (d_dst && cur != d_dst) {
        // no kernel has written the caller's frame: every stage is off (cur == d_src) or the PRE filter was the last one
        // (cur == the lane's d_pre; sizes are equal here, a resize always writes d_dst or feeds a later stage)
        VP_CUDA(c, cudaMemcpy2DAsync(d_dst, dst_pitch, cur, cur_pitch, (size_t)g.dst_w * 3, g.dst_h, cudaMemcpyDeviceToDevice, st));
    }
    temporal_advance(c, temporal);
    return VP_OK;
}

// after pipelined host frames: order `st0` behind every lane they used
static int join_host_lanes(vp_ctx* c, cudaStream_t st0) {
    if (!c->host_prev_done) return VP_OK;
    for (int l = 0; l < 4; ++l)
        if (c->lanes[l].ev_done && (l == 0 || c->lanes[l].st)) VP_CUDA(c, cudaStreamWaitEvent(st0, c->lanes[l].ev_done, 0));
    c->host_prev_done = nullptr; c->host_lane = 0;
    return VP_OK;
}

int vp_process_device(vp_ctx* c, const uint8_t* d_src, size_t src_pitch, uint8_t* d_dst, size_t dst_pitch, void* stream) {
    if (!c) return VP_ERR_INVALID_ARG;
    VP_CUDA(c, cudaSetDevice(c->device));
    { int rc = join_host_lanes(c, stream ? (cudaStream_t)stream : c->s_compute); if (rc) return rc; }
    return process_frame(c, 0, stream ? (cudaStream_t)stream : c->s_compute, nullptr, d_src, src_pitch, d_dst, dst_pitch);
}


static int ensure_lane(vp_ctx* c, int ln) {
    Lane& L = c->lanes[ln];
    if (L.st) return VP_OK;
    const vp_chain_config& g = c->cfg;
    VP_CUDA(c, cudaStreamCreateWithFlags(&L.st, cudaStreamNonBlocking));
    if (g.use_pre && !g.bicubic_enhance) VP_CUDA(c, cudaMalloc(&L.d_pre, std::max(c->pre_pitch, c->pre4_pitch) * g.src_h));
    VP_CUDA(c, cudaMalloc(&L.d_sharp, std::max(c->frame_pitch, c->sharp4_pitch) * g.dst_h));
    if (g.use_post && !g.bicubic_enhance && (bil_has_modes(g.post) || (g.temporal.mode == VP_TEMPORAL_BLEND_FRAMES && g.temporal.use_flow)))
        VP_CUDA(c, cudaMalloc(&L.d_post, c->frame_pitch * g.dst_h));
    VP_CUDA(c, cudaMalloc(&L.d_sums, 18 * sizeof(unsigned long long)));
    VP_CUDA(c, cudaMalloc(&L.d_sel, 4 * sizeof(int)));
    VP_CUDA(c, cudaMemset(L.d_sel, 0, 4 * sizeof(int)));
    VP_CUDA(c, cudaMalloc(&L.d_counter, 4 * sizeof(unsigned)));
    VP_CUDA(c, cudaMemset(L.d_counter, 0, 4 * sizeof(unsigned)));
    { int rc = alloc_enhance_buffers(c, L); if (rc) return rc; }
    for (Lane& X : c->lanes)
        if (!X.ev_done) VP_CUDA(c, cudaEventCreateWithFlags(&X.ev_done, cudaEventDisableTiming));
    if (!c->ev_batch) VP_CUDA(c, cudaEventCreateWithFlags(&c->ev_batch, cudaEventDisableTiming));
    return VP_OK;
}

int vp_set_batch_lanes(vp_ctx* c, int lanes) {
    if (!c || lanes < 1 || lanes > 4) return VP_ERR_INVALID_ARG;
    c->nlanes = lanes;
    return VP_OK;
}

static void drop_batch_graphs(vp_ctx* c) {
    for (auto& g : c->graphs) cudaGraphExecDestroy(g.exec);
    c->graphs.clear(); c->graph_seen.clear();
}
int vp_set_batch_graphs(vp_ctx* c, int on) {
    if (!c) return VP_ERR_INVALID_ARG;
    VP_CUDA(c, cudaSetDevice(c->device));
    c->graph_ok = on != 0;
    if (!on) drop_batch_graphs(c);
    return VP_OK;
}
uint64_t vp_batch_graph_replays(const vp_ctx* c) { return c ? c->graph_replays : 0; }

// the frames of one batch over the stream lanes (eagerly, or while `st0` is being captured into a graph)
static int batch_enqueue(vp_ctx* c, int n, int NL, cudaStream_t st0, const uint8_t* const* d_src, size_t src_pitch, uint8_t* const* d_dst,
                         size_t dst_pitch) {
    int rc = VP_OK;
    // lane 1 forks from, and later joins, the caller's stream: everything is ordered after what is already
    // queued there, and events the caller records on it bracket the whole batch
    if (NL > 1) {
        VP_CUDA(c, cudaEventRecord(c->ev_batch, st0));
        for (int l = 1; l < NL; ++l) VP_CUDA(c, cudaStreamWaitEvent(c->lanes[l].st, c->ev_batch, 0));
    }
    cudaEvent_t prev_done = nullptr;
    for (int i = 0; i < n; ++i) {
        const int ln = i % NL;
        cudaStream_t st = ln == 0 ? st0 : c->lanes[ln].st;
        rc = process_frame(c, ln, st, prev_done, d_src[i], src_pitch, d_dst[i], dst_pitch);
        if (rc) return rc;
        VP_CUDA(c, cudaEventRecord(c->lanes[ln].ev_done, st));
        prev_done = c->lanes[ln].ev_done;
    }
    for (int l = 1; l < NL; ++l) VP_CUDA(c, cudaStreamWaitEvent(st0, c->lanes[l].ev_done, 0));   // join
    return VP_OK;
}

// chains whose frames are nothing but kernel launches on the lane streams (no lazily built tables, no host
// synchronisation, no cooperative launch): the north-star chain and its bicubic-only / no-temporal forms
static bool batch_graphable(const vp_ctx* c) {
    const vp_chain_config& g = c->cfg;
    if (!c->graph_ok || c->prof_on || g.bicubic_enhance || c->d_scene_acc) return false;
    if (g.temporal.mode != VP_TEMPORAL_NONE && g.temporal.use_flow) return false;
    if (g.use_pre && (g.pre.selective || g.pre.use_multiscale)) return false;
    if (g.use_post && (g.post.selective || g.post.use_multiscale)) return false;
    if (g.use_sharpen && g.sharpen.adaptive_sigma) return false;
    return true;
}

int vp_process_batch_device(vp_ctx* c, int n, const uint8_t* const* d_src, size_t src_pitch, uint8_t* const* d_dst, size_t dst_pitch,
                            void* stream) {
    if (!c || n < 0 || (n && (!d_src || !d_dst))) return VP_ERR_INVALID_ARG;
    if (!c->has_chain) return fail(c, VP_ERR_STATE, "vp_process_batch_device: context was created without a chain config");
    if (n == 0) return VP_OK;
    VP_CUDA(c, cudaSetDevice(c->device));
    int rc = VP_OK;
    const int NL = std::min(c->nlanes, n);
    for (int l = 1; l < NL; ++l) { rc = ensure_lane(c, l); if (rc) return rc; }
    if (NL == 1 && !c->lanes[0].ev_done) VP_CUDA(c, cudaEventCreateWithFlags(&c->lanes[0].ev_done, cudaEventDisableTiming));
    cudaStream_t st0 = stream ? (cudaStream_t)stream : c->s_compute;
    rc = join_host_lanes(c, st0); if (rc) return rc;
    if (NL > 1 && !c->ev_batch) VP_CUDA(c, cudaEventCreateWithFlags(&c->ev_batch, cudaEventDisableTiming));

    if (n >= 8 && batch_graphable(c)) {
        // FNV-1a over everything a frame's launches depend on besides the (fixed) configuration and lane buffers
        uint64_t key = 1469598103934665603ull;
        auto mix = [&](uint64_t v) { for (int b = 0; b < 8; ++b) { key ^= (v >> (8 * b)) & 0xff; key *= 1099511628211ull; } };
        mix((uint64_t)n); mix((uint64_t)NL); mix(src_pitch); mix(dst_pitch); mix((uint64_t)c->head); mix((uint64_t)c->nvalid);
        mix((uint64_t)(uintptr_t)st0);
        for (int i = 0; i < n; ++i) { mix((uint64_t)(uintptr_t)d_src[i]); mix((uint64_t)(uintptr_t)d_dst[i]); }
        for (auto& g : c->graphs)
            if (g.key == key) {
                VP_CUDA(c, cudaGraphLaunch(g.exec, st0));
                c->head = g.head; c->nvalid = g.nvalid; c->last_lane = g.last_lane; c->launches += g.launches;
                g.used = ++c->graph_clock; c->graph_replays++;
                return VP_OK;
            }
        const bool seen = std::find(c->graph_seen.begin(), c->graph_seen.end(), key) != c->graph_seen.end();
        if (!seen) {
            if (c->graph_seen.size() >= 64) c->graph_seen.erase(c->graph_seen.begin());
            c->graph_seen.push_back(key);
        } else {
            // second sighting: capture the batch (capture-only events: an event recorded inside a capture cannot be waited
            // on by later eager work), instantiate, replay from now on
            cudaEvent_t saved_batch = c->ev_batch, saved_done[4];
            if (!c->ev_cap_batch) VP_CUDA(c, cudaEventCreateWithFlags(&c->ev_cap_batch, cudaEventDisableTiming));
            for (int l = 0; l < NL; ++l) {
                saved_done[l] = c->lanes[l].ev_done;
                if (!c->ev_cap_done[l]) VP_CUDA(c, cudaEventCreateWithFlags(&c->ev_cap_done[l], cudaEventDisableTiming));
            }
            const int head0 = c->head, nvalid0 = c->nvalid, lane0 = c->last_lane; const uint64_t launches0 = c->launches;
            cudaGraph_t graph = nullptr; cudaGraphExec_t exec = nullptr;
            bool ok = cudaStreamBeginCapture(st0, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
            if (ok) {
                c->ev_batch = c->ev_cap_batch;
                for (int l = 0; l < NL; ++l) c->lanes[l].ev_done = c->ev_cap_done[l];
                const int crc = batch_enqueue(c, n, NL, st0, d_src, src_pitch, d_dst, dst_pitch);
                c->ev_batch = saved_batch;
                for (int l = 0; l < NL; ++l) c->lanes[l].ev_done = saved_done[l];
                ok = cudaStreamEndCapture(st0, &graph) == cudaSuccess && crc == VP_OK && graph != nullptr;
                if (ok) ok = cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess;
                if (graph) cudaGraphDestroy(graph);
            }
            if (ok) {
                vp_ctx::BatchGraph g;
                g.key = key; g.exec = exec; g.head = c->head; g.nvalid = c->nvalid; g.last_lane = c->last_lane;
                g.launches = c->launches - launches0; g.used = ++c->graph_clock;
                if (c->graphs.size() >= 4) {         // least recently used out
                    size_t lru = 0;
                    for (size_t i = 1; i < c->graphs.size(); ++i) if (c->graphs[i].used < c->graphs[lru].used) lru = i;
                    cudaGraphExecDestroy(c->graphs[lru].exec);
                    c->graphs.erase(c->graphs.begin() + lru);
                }
                c->graphs.push_back(g);
                VP_CUDA(c, cudaGraphLaunch(exec, st0));
                c->graph_replays++;
                return VP_OK;
            }
            // capture refused (e.g. the caller's stream is itself being captured): stay eager from now on
            cudaGetLastError();
            c->graph_ok = false;
            c->head = head0; c->nvalid = nvalid0; c->last_lane = lane0; c->launches = launches0;
        }
    }
    return batch_enqueue(c, n, NL, st0, d_src, src_pitch, d_dst, dst_pitch);
}

int vp_get_stage_params(vp_ctx* c, int which, int* d, double* sigma_color, double* sigma_space) {
    if (!c || which < 0 || which > 1) return VP_ERR_INVALID_ARG;
    if (!c->has_chain) return fail(c, VP_ERR_STATE, "vp_get_stage_params: no chain");
    int rc = vp_synchronize(c); if (rc) return rc;
    int sel[2];
    VP_CUDA(c, cudaMemcpy(sel, c->lanes[c->last_lane].d_sel, sizeof(sel), cudaMemcpyDeviceToHost));
    const vp_bilateral_config& b = which == VP_STAGE_PRE ? c->cfg.pre : c->cfg.post;
    int dd = b.diameter; double sc = b.sigma_color, ss = b.sigma_space;
    if (b.adaptive_params)
        vp_host::adaptive_params_for_class(sel[which], b.stage == VP_STAGE_POST, b.diameter, b.sigma_color, b.sigma_space, &dd, &sc, &ss);
    if (d) *d = dd;
    if (sigma_color) *sigma_color = sc;
    if (sigma_space) *sigma_space = ss;
    return VP_OK;
}

// ---- host frames: pinned, pipelined -------------------------------------------------------------------
int vp_host_alloc(void** p, size_t bytes) {
    if (!p || !bytes) return VP_ERR_INVALID_ARG;
    if (cudaHostAlloc(p, bytes, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); *p = nullptr; return VP_ERR_NOMEM; }
    return VP_OK;
}
int vp_host_free(void* p) {
    if (!p) return VP_OK;
    if (cudaFreeHost(p) != cudaSuccess) { cudaGetLastError(); return VP_ERR_CUDA; }
    return VP_OK;
}

int vp_pipeline_depth(const vp_ctx* c) { return c ? PIPE_DEPTH : 0; }

static int ensure_slots(vp_ctx* c) {
    if (!c->slots.empty()) return VP_OK;
    const vp_chain_config& g = c->cfg;
    c->slots.resize(PIPE_DEPTH);
    for (HostSlot& s : c->slots) {
        VP_CUDA(c, cudaMalloc(&s.d_in, c->pre_pitch * g.src_h));
        VP_CUDA(c, cudaMalloc(&s.d_out, c->frame_pitch * g.dst_h));
        VP_CUDA(c, cudaEventCreateWithFlags(&s.ev_h2d, cudaEventDisableTiming));
        VP_CUDA(c, cudaEventCreateWithFlags(&s.ev_comp, cudaEventDisableTiming));
        VP_CUDA(c, cudaEventCreateWithFlags(&s.ev_d2h, cudaEventDisableTiming));
    }
    return VP_OK;
}

int vp_submit_host(vp_ctx* c, const uint8_t* src, size_t src_step, uint8_t* dst, size_t dst_step) {
    if (!c) return VP_ERR_INVALID_ARG;
    if (!c->has_chain) return fail(c, VP_ERR_STATE, "vp_submit_host: context was created without a chain config");
    const vp_chain_config& g = c->cfg;
    int rc = check_frame(c, src, src_step, g.src_w, g.src_h, "vp_submit_host(src)");
    if (rc) return rc;
    if (dst && dst_step < (size_t)g.dst_w * 3) return fail(c, VP_ERR_INVALID_ARG, "vp_submit_host(dst): step < 3*width");
    if ((int)c->inflight.size() >= PIPE_DEPTH) return fail(c, VP_ERR_STATE, "vp_submit_host: pipeline full, call vp_wait first");
    VP_CUDA(c, cudaSetDevice(c->device));
    rc = ensure_slots(c); if (rc) return rc;
    HostSlot& s = c->slots[c->next_slot];
    const size_t in_row = (size_t)g.src_w * 3, out_row = (size_t)g.dst_w * 3;

    const uint8_t* hsrc = src; size_t hstep = src_step;
    if (!host_ptr_is_pinned(src)) {   // stage through pinned memory so the copy is a real async DMA
        if (!s.h_in) VP_CUDA(c, cudaHostAlloc((void**)&s.h_in, in_row * g.src_h, cudaHostAllocDefault));
        for (int y = 0; y < g.src_h; ++y) std::memcpy(s.h_in + (size_t)y * in_row, src + (size_t)y * src_step, in_row);
        hsrc = s.h_in; hstep = in_row;
    }
    VP_CUDA(c, cudaMemcpy2DAsync(s.d_in, c->pre_pitch, hsrc, hstep, in_row, g.src_h, cudaMemcpyHostToDevice, c->s_h2d));
    VP_CUDA(c, cudaEventRecord(s.ev_h2d, c->s_h2d));
    // one lane per frame in flight: the next frame's PRE / upscale / POST (and flow expansion) overlap this frame's tail;
    // process_frame orders the temporal stage behind the previous frame through host_prev_done
    const int NL = std::min(c->nlanes, PIPE_DEPTH);
    for (int l = 1; l < NL; ++l) { rc = ensure_lane(c, l); if (rc) return rc; }
    if (!c->lanes[0].ev_done) VP_CUDA(c, cudaEventCreateWithFlags(&c->lanes[0].ev_done, cudaEventDisableTiming));
    const int ln = c->host_lane % NL;
    cudaStream_t st = ln == 0 ? c->s_compute : c->lanes[ln].st;
    if (!c->host_prev_done && ln != 0) return fail(c, VP_ERR_STATE, "vp_submit_host: lane order lost");
    if (!c->host_prev_done && NL > 1) {            // first pipelined frame: the other lanes start behind what s_compute holds
        if (!c->ev_batch) VP_CUDA(c, cudaEventCreateWithFlags(&c->ev_batch, cudaEventDisableTiming));
        VP_CUDA(c, cudaEventRecord(c->ev_batch, c->s_compute));
        for (int l = 1; l < NL; ++l) VP_CUDA(c, cudaStreamWaitEvent(c->lanes[l].st, c->ev_batch, 0));
    }
    VP_CUDA(c, cudaStreamWaitEvent(st, s.ev_h2d, 0));
    rc = process_frame(c, ln, st, c->host_prev_done, s.d_in, c->pre_pitch, dst ? s.d_out : nullptr, c->frame_pitch);
    if (rc) {
        // the slot is not queued, so nothing would wait for its upload before the next submit reuses d_in / h_in: drain the
        // copy stream and the lane here, and drop the temporal history (the failed frame may have advanced it half way)
        cudaStreamSynchronize(c->s_h2d); cudaStreamSynchronize(st); cudaGetLastError();
        const std::string keep = c->err;
        vp_reset_temporal(c);
        c->err = keep;
        return rc;
    }
    VP_CUDA(c, cudaEventRecord(c->lanes[ln].ev_done, st));
    c->host_prev_done = c->lanes[ln].ev_done;
    c->host_lane = (ln + 1) % NL;
    VP_CUDA(c, cudaEventRecord(s.ev_comp, st));
    s.user_dst = dst; s.user_dst_step = dst_step; s.staged_out = false;
    if (dst) {
        VP_CUDA(c, cudaStreamWaitEvent(c->s_d2h, s.ev_comp, 0));
        uint8_t* hdst = dst; size_t hdstep = dst_step;
        if (!host_ptr_is_pinned(dst)) {
            if (!s.h_out) VP_CUDA(c, cudaHostAlloc((void**)&s.h_out, out_row * g.dst_h, cudaHostAllocDefault));
            hdst = s.h_out; hdstep = out_row; s.staged_out = true;
        }
        VP_CUDA(c, cudaMemcpy2DAsync(hdst, hdstep, s.d_out, c->frame_pitch, out_row, g.dst_h, cudaMemcpyDeviceToHost, c->s_d2h));
        VP_CUDA(c, cudaEventRecord(s.ev_d2h, c->s_d2h));
    }
    c->inflight.push_back(c->next_slot);
    c->next_slot = (c->next_slot + 1) % PIPE_DEPTH;
    return VP_OK;
}

int vp_wait(vp_ctx* c) {
    if (!c) return VP_ERR_INVALID_ARG;
    if (c->inflight.empty()) return fail(c, VP_ERR_STATE, "vp_wait: nothing submitted");
    VP_CUDA(c, cudaSetDevice(c->device));
    HostSlot& s = c->slots[c->inflight.front()];
    // the frame leaves the queue only once it is known to be complete: a CUDA error here keeps it (and its slot) in flight
    VP_CUDA(c, cudaEventSynchronize(s.user_dst ? s.ev_d2h : s.ev_comp));
    c->inflight.pop_front();
    if (s.user_dst && s.staged_out) {
        const size_t out_row = (size_t)c->cfg.dst_w * 3;
        for (int y = 0; y < c->cfg.dst_h; ++y)
            std::memcpy(s.user_dst + (size_t)y * s.user_dst_step, s.h_out + (size_t)y * out_row, out_row);
    }
    return VP_OK;
}

int vp_process_host(vp_ctx* c, const uint8_t* src, size_t src_step, uint8_t* dst, size_t dst_step) {
    if (!c) return VP_ERR_INVALID_ARG;
    if (!c->inflight.empty()) return fail(c, VP_ERR_STATE, "vp_process_host: frames still in flight, call vp_wait first");
    int rc = vp_submit_host(c, src, src_step, dst, dst_step);
    if (rc) return rc;
    return vp_wait(c);
}

// ---- stage-wise entry points -----------------------------------------------------------------------
int vp_bicubic(vp_ctx* c, const uint8_t* d_src, size_t src_pitch, int src_w, int src_h,
               uint8_t* d_dst, size_t dst_pitch, int dst_w, int dst_h, void* stream) {
    if (!c) return VP_ERR_INVALID_ARG;
    int rc = check_frame(c, d_src, src_pitch, src_w, src_h, "vp_bicubic(src)"); if (rc) return rc;
    rc = check_frame(c, d_dst, dst_pitch, dst_w, dst_h, "vp_bicubic(dst)"); if (rc) return rc;
    if (dst_w < src_w || dst_h < src_h) return fail(c, VP_ERR_UNSUPPORTED, "vp_bicubic: downscaling is not built");
    VP_CUDA(c, cudaSetDevice(c->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : c->s_compute;
    StageOrder stage_order(c, st);
    ResizeCache& rcache = (c->has_chain && c->rc_chain.sw == src_w && c->rc_chain.sh == src_h &&
                           c->rc_chain.dw == dst_w && c->rc_chain.dh == dst_h && c->rc_chain.xofs) ? c->rc_chain : c->rc_tmp;
    rc = ensure_resize(c, rcache, src_w, src_h, dst_w, dst_h, st); if (rc) return rc;
    return launch_bicubic(c, d_src, src_pitch, src_w, src_h, d_dst, dst_pitch, dst_w, dst_h, rcache, st);
}

int vp_bilateral(vp_ctx* c, const uint8_t* d_src, size_t src_pitch, uint8_t* d_dst, size_t dst_pitch,
                 int w, int h, int d, double sigma_color, double sigma_space, void* stream) {
    if (!c) return VP_ERR_INVALID_ARG;
    int rc = check_frame(c, d_src, src_pitch, w, h, "vp_bilateral(src)"); if (rc) return rc;
    rc = check_frame(c, d_dst, dst_pitch, w, h, "vp_bilateral(dst)"); if (rc) return rc;
    VP_CUDA(c, cudaSetDevice(c->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : c->s_compute;
    StageOrder stage_order(c, st);
    if (!(c->tmp_d == d && c->tmp_sc == sigma_color && c->tmp_ss == sigma_space)) {
        vp_host::BilateralTables t = vp_host::bilateral_tables(d, sigma_color, sigma_space);
        if (t.radius > BIL_RMAX) return fail(c, VP_ERR_UNSUPPORTED, "vp_bilateral: radius > 7 (diameter > 15) is not built");
        BilCand bc; fill_cand(bc, t);
        VP_CUDA(c, cudaMemcpyAsync(c->d_cand_tmp, &bc, sizeof(bc), cudaMemcpyHostToDevice, st));
        VP_CUDA(c, cudaStreamSynchronize(st));
        c->tmp_d = d; c->tmp_sc = sigma_color; c->tmp_ss = sigma_space; c->tmp_r = t.radius;
    }
    BilArgs a{};
    a.src = d_src; a.spitch = src_pitch; a.dst = d_dst; a.dpitch = dst_pitch; a.w = w; a.h = h;
    a.cand = c->d_cand_tmp;
    const int rad[3] = {c->tmp_r, c->tmp_r, c->tmp_r};
    return launch_bilateral(c, a, rad, st);
}

int vp_selective_bilateral(vp_ctx* c, const vp_bilateral_config* cfg, const uint8_t* d_src, size_t src_pitch,
                           uint8_t* d_dst, size_t dst_pitch, int w, int h, void* stream) {
    if (!c || !cfg) return VP_ERR_INVALID_ARG;
    int rc = check_frame(c, d_src, src_pitch, w, h, "vp_selective_bilateral(src)"); if (rc) return rc;
    rc = check_frame(c, d_dst, dst_pitch, w, h, "vp_selective_bilateral(dst)"); if (rc) return rc;
    VP_CUDA(c, cudaSetDevice(c->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : c->s_compute;
    StageOrder stage_order(c, st);
    if (!c->tmp_sel_valid || std::memcmp(&c->tmp_sel_cfg, cfg, sizeof(*cfg)) != 0) {
        VP_CUDA(c, cudaStreamSynchronize(st));
        c->tmp_sel_valid = false;
        rc = build_selective_cands(c, *cfg, c->d_cand_sel, c->tmp_sel_rad, st); if (rc) return rc;
        if (cfg->use_multiscale) { rc = build_multiscale_cands(c, *cfg, &c->d_cand_ms_tmp, c->ms_tmp_rad, st); if (rc) return rc; }
        c->tmp_sel_cfg = *cfg; c->tmp_sel_valid = true;
    }
    unsigned long long* sums = c->lanes[0].d_sums + 12;
    if (bil_has_modes(*cfg))
        return run_bilateral_modes(c, c->ms_tmp, *cfg, c->d_cand_sel, c->tmp_sel_rad, c->d_cand_ms_tmp, c->ms_tmp_rad,
                                   StatsRes{sums, c->lanes[0].d_counter + 2, c->lanes[0].d_sel + 2}, d_src, src_pitch, d_dst, dst_pitch, w, h, st, K_BIL_OTHER);
    if (cfg->adaptive_params) {
        VP_CUDA(c, cudaMemsetAsync(sums, 0, 6 * sizeof(unsigned long long), st));
        rc = launch_stats(c, d_src, src_pitch, w, h, sums, SelOut{c->lanes[0].d_counter + 2, c->lanes[0].d_sel + 2, (double)w * h}, st); if (rc) return rc;
    }
    BilArgs a{};
    a.src = d_src; a.spitch = src_pitch; a.dst = d_dst; a.dpitch = dst_pitch; a.w = w; a.h = h;
    a.cand = c->d_cand_sel; a.sel = cfg->adaptive_params ? c->lanes[0].d_sel + 2 : nullptr;
    return launch_bilateral(c, a, c->tmp_sel_rad, st);
}

int vp_detail_mask(vp_ctx* c, double detail_threshold, const uint8_t* d_src, size_t src_pitch, float* d_mask, size_t mask_pitch,
                   int w, int h, void* stream) {
    if (!c || !d_mask) return VP_ERR_INVALID_ARG;
    int rc = check_frame(c, d_src, src_pitch, w, h, "vp_detail_mask(src)"); if (rc) return rc;
    if (mask_pitch < (size_t)w * 4) return fail(c, VP_ERR_INVALID_ARG, "vp_detail_mask: mask pitch < 4*width");
    VP_CUDA(c, cudaSetDevice(c->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : c->s_compute;
    StageOrder stage_order(c, st);
    rc = ensure_ms_scratch(c, c->ms_tmp, w, h, 1, true); if (rc) return rc;
    return launch_detail_mask(c, d_src, src_pitch, w, h, detail_threshold, c->ms_tmp.mx, c->ms_tmp.dv, c->ms_tmp.dvp, d_mask, mask_pitch, st);
}

int vp_pyr_down(vp_ctx* c, const uint8_t* d_src, size_t src_pitch, int src_w, int src_h, uint8_t* d_dst, size_t dst_pitch, void* stream) {
    if (!c) return VP_ERR_INVALID_ARG;
    const int dw = (src_w + 1) / 2, dh = (src_h + 1) / 2;
    int rc = check_frame(c, d_src, src_pitch, src_w, src_h, "vp_pyr_down(src)"); if (rc) return rc;
    rc = check_frame(c, d_dst, dst_pitch, dw, dh, "vp_pyr_down(dst)"); if (rc) return rc;
    VP_CUDA(c, cudaSetDevice(c->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : c->s_compute;
    StageOrder stage_order(c, st);
    ProfScope ps(c, K_BIL_OTHER, st);
    k_pyr_down<<<dim3(((dw + 1) / 2 + 31) / 32, ((dh + 1) / 2 + 7) / 8), 256, 0, st>>>(d_src, src_pitch, src_w, src_h, d_dst, dst_pitch, dw, dh);
    VP_CUDA(c, cudaGetLastError());
    c->launches++;
    return VP_OK;
}

int vp_pyr_up(vp_ctx* c, const uint8_t* d_src, size_t src_pitch, int src_w, int src_h, uint8_t* d_dst, size_t dst_pitch,
              int dst_w, int dst_h, void* stream) {
    if (!c) return VP_ERR_INVALID_ARG;
    int rc = check_frame(c, d_src, src_pitch, src_w, src_h, "vp_pyr_up(src)"); if (rc) return rc;
    rc = check_frame(c, d_dst, dst_pitch, dst_w, dst_h, "vp_pyr_up(dst)"); if (rc) return rc;
    if ((dst_w != 2 * src_w && dst_w != 2 * src_w - 1) || (dst_h != 2 * src_h && dst_h != 2 * src_h - 1))
        return fail(c, VP_ERR_UNSUPPORTED, "vp_pyr_up: dst must be 2*src or 2*src-1 per axis");
    VP_CUDA(c, cudaSetDevice(c->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : c->s_compute;
    StageOrder stage_order(c, st);
    ProfScope ps(c, K_BIL_OTHER, st);
    k_pyr_up<<<dim3((dst_w + 31) / 32, (dst_h + 7) / 8), 256, 0, st>>>(d_src, src_pitch, src_w, src_h, d_dst, dst_pitch, dst_w, dst_h);
    VP_CUDA(c, cudaGetLastError());
    c->launches++;
    return VP_OK;
}

int vp_mean_stddev_sums(vp_ctx* c, const uint8_t* d_src, size_t src_pitch, int w, int h, uint64_t* d_sums, void* stream) {
    if (!c || !d_sums) return VP_ERR_INVALID_ARG;
    int rc = check_frame(c, d_src, src_pitch, w, h, "vp_mean_stddev_sums(src)"); if (rc) return rc;
    VP_CUDA(c, cudaSetDevice(c->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : c->s_compute;
    StageOrder stage_order(c, st);
    VP_CUDA(c, cudaMemsetAsync(d_sums, 0, 6 * sizeof(uint64_t), st));
    return launch_stats(c, d_src, src_pitch, w, h, reinterpret_cast<unsigned long long*>(d_sums), SelOut{nullptr, nullptr, 0.0}, st);
}

static int sharpen_common(vp_ctx* c, const vp_sharpen_config* cfg, const uint8_t* d_src, size_t src_pitch,
                          const float* mask_in, size_t mask_in_pitch, float* mask_out, size_t mask_out_pitch,
                          uint8_t* d_dst, size_t dst_pitch, int w, int h, void* stream, const char* who) {
    if (!c || !cfg) return VP_ERR_INVALID_ARG;
    int rc = check_frame(c, d_src, src_pitch, w, h, who); if (rc) return rc;
    if (d_dst && dst_pitch < (size_t)w * 3) return fail(c, VP_ERR_INVALID_ARG, std::string(who) + ": dst pitch < 3*width");
    if ((mask_in && mask_in_pitch < (size_t)w * 4) || (mask_out && mask_out_pitch < (size_t)w * 4))
        return fail(c, VP_ERR_INVALID_ARG, std::string(who) + ": mask pitch < 4*width");
    VP_CUDA(c, cudaSetDevice(c->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : c->s_compute;
    StageOrder stage_order(c, st);
    UpsharpArgs a{};
    rc = fill_sharpen_args(c, a, *cfg); if (rc) return rc;
    rc = ensure_siglut(c, &c->d_siglut_tmp, &c->siglut_tmp_thr, &c->siglut_tmp_valid, cfg->edge_threshold, st); if (rc) return rc;
    if (cfg->adaptive_sigma && d_dst && !mask_in && !mask_out) {      // vp_sharpen, adaptive_sigma branch
        rc = ensure_vs_scratch(c, c->vs_tmp, w, h, false); if (rc) return rc;
        return run_varsharp(c, c->vs_tmp, *cfg, c->d_siglut_tmp, d_src, src_pitch, d_dst, dst_pitch, w, h, nullptr, SelOut{nullptr, nullptr, 0.0}, st);
    }
    a.sig_lut = c->d_siglut_tmp;
    a.src = d_src; a.spitch = src_pitch; a.sw = w; a.sh = h;
    a.dst = d_dst; a.dpitch = dst_pitch; a.dw = w; a.dh = h;
    a.max_src_rows = US_TH + 2 * US_HALO; a.max_src_cols = US_TW + 2 * US_HALO;
    a.mask_in = mask_in; a.mask_in_pitch = mask_in_pitch; a.mask_out = mask_out; a.mask_out_pitch = mask_out_pitch;
    return launch_upsharp(c, a, false, true, st);
}

int vp_edge_mask(vp_ctx* c, const vp_sharpen_config* cfg, const uint8_t* d_src, size_t src_pitch,
                 float* d_mask, size_t mask_pitch, int w, int h, void* stream) {
    if (!d_mask) return VP_ERR_INVALID_ARG;
    return sharpen_common(c, cfg, d_src, src_pitch, nullptr, 0, d_mask, mask_pitch, nullptr, 0, w, h, stream, "vp_edge_mask");
}

int vp_unsharp(vp_ctx* c, const vp_sharpen_config* cfg, const uint8_t* d_src, size_t src_pitch,
               const float* d_mask, size_t mask_pitch, uint8_t* d_dst, size_t dst_pitch, int w, int h, void* stream) {
    if (!d_mask || !d_dst) return VP_ERR_INVALID_ARG;
    return sharpen_common(c, cfg, d_src, src_pitch, d_mask, mask_pitch, nullptr, 0, d_dst, dst_pitch, w, h, stream, "vp_unsharp");
}

int vp_sharpen(vp_ctx* c, const vp_sharpen_config* cfg, const uint8_t* d_src, size_t src_pitch,
               uint8_t* d_dst, size_t dst_pitch, int w, int h, void* stream) {
    if (!d_dst) return VP_ERR_INVALID_ARG;
    return sharpen_common(c, cfg, d_src, src_pitch, nullptr, 0, nullptr, 0, d_dst, dst_pitch, w, h, stream, "vp_sharpen");
}

int vp_sigma_map(vp_ctx* c, const uint8_t* d_src, size_t src_pitch, float* d_sigma, size_t sigma_pitch, int w, int h, void* stream) {
    if (!c || !d_sigma) return VP_ERR_INVALID_ARG;
    int rc = check_frame(c, d_src, src_pitch, w, h, "vp_sigma_map(src)"); if (rc) return rc;
    if (sigma_pitch < (size_t)w * 4) return fail(c, VP_ERR_INVALID_ARG, "vp_sigma_map: sigma pitch < 4*width");
    VP_CUDA(c, cudaSetDevice(c->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : c->s_compute;
    StageOrder stage_order(c, st);
    rc = ensure_vs_scratch(c, c->vs_tmp, w, h, false); if (rc) return rc;
    return launch_sigma_map(c, c->vs_tmp.st, c->vs_tmp.var, c->vs_tmp.vpitch, d_src, src_pitch, w, h, d_sigma, sigma_pitch, st);
}

int vp_optical_flow(vp_ctx* c, const vp_temporal_config* cfg, const uint8_t* d_prev, const uint8_t* d_cur, size_t pitch,
                    float* d_flow_x, float* d_flow_y, size_t flow_pitch, int w, int h, void* stream) {
    if (!c || !cfg || !d_flow_x || !d_flow_y) return VP_ERR_INVALID_ARG;
    int rc = check_frame(c, d_prev, pitch, w, h, "vp_optical_flow(prev)"); if (rc) return rc;
    rc = check_frame(c, d_cur, pitch, w, h, "vp_optical_flow(cur)"); if (rc) return rc;
    if (flow_pitch < (size_t)w * 4) return fail(c, VP_ERR_INVALID_ARG, "vp_optical_flow: flow pitch < 4*width");
    VP_CUDA(c, cudaSetDevice(c->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : c->s_compute;
    StageOrder stage_order(c, st);
    FlowEngine& E = c->flow_tmp;
    rc = ensure_flow_engine(c, E, w, h, *cfg); if (rc) return rc;
    rc = flow_expand(c, E, d_prev, pitch, 0, st); if (rc) return rc;
    rc = flow_expand(c, E, d_cur, pitch, 1, st); if (rc) return rc;
    rc = flow_solve(c, E, 0, 1, st); if (rc) return rc;
    const FlowLevel& L0 = E.lv[0];
    VP_CUDA(c, cudaMemcpy2DAsync(d_flow_x, flow_pitch, L0.flow, (size_t)L0.p * 4, (size_t)w * 4, h, cudaMemcpyDeviceToDevice, st));
    VP_CUDA(c, cudaMemcpy2DAsync(d_flow_y, flow_pitch, L0.flow + L0.ps, (size_t)L0.p * 4, (size_t)w * 4, h, cudaMemcpyDeviceToDevice, st));
    return VP_OK;
}

int vp_warp_flow(vp_ctx* c, const uint8_t* d_src, size_t src_pitch, const float* d_flow_x, const float* d_flow_y, size_t flow_pitch,
                 uint8_t* d_dst, size_t dst_pitch, int w, int h, void* stream) {
    if (!c || !d_flow_x || !d_flow_y) return VP_ERR_INVALID_ARG;
    int rc = check_frame(c, d_src, src_pitch, w, h, "vp_warp_flow(src)"); if (rc) return rc;
    rc = check_frame(c, d_dst, dst_pitch, w, h, "vp_warp_flow(dst)"); if (rc) return rc;
    if (flow_pitch < (size_t)w * 4 || (flow_pitch & 3)) return fail(c, VP_ERR_INVALID_ARG, "vp_warp_flow: flow pitch");
    VP_CUDA(c, cudaSetDevice(c->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : c->s_compute;
    StageOrder stage_order(c, st);
    ProfScope ps(c, K_BLEND, st);
    k_warp<<<grid2d(w, h), 256, 0, st>>>(d_src, src_pitch, d_flow_x, d_flow_y, (int)(flow_pitch / 4), d_dst, dst_pitch, w, h);
    VP_CUDA(c, cudaGetLastError());
    c->launches++;
    return VP_OK;
}

int vp_flow_reliability(vp_ctx* c, const float* d_flow_x, const float* d_flow_y, size_t flow_pitch, float motion_threshold,
                        float* d_mask, size_t mask_pitch, int w, int h, void* stream) {
    if (!c || !d_flow_x || !d_flow_y || !d_mask || w <= 0 || h <= 0) return VP_ERR_INVALID_ARG;
    if (flow_pitch < (size_t)w * 4 || (flow_pitch & 3) || mask_pitch < (size_t)w * 4 || (mask_pitch & 3))
        return fail(c, VP_ERR_INVALID_ARG, "vp_flow_reliability: pitch");
    VP_CUDA(c, cudaSetDevice(c->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : c->s_compute;
    StageOrder stage_order(c, st);
    const size_t need = mask_pitch / 4 * (size_t)h;
    if (c->flow_tmp_mask_n < need) {
        VP_CUDA(c, cudaStreamSynchronize(st));
        cudaFree(c->flow_tmp_mask); c->flow_tmp_mask = nullptr; c->flow_tmp_mask_n = 0;
        VP_CUDA(c, cudaMalloc(&c->flow_tmp_mask, need * sizeof(float)));
        c->flow_tmp_mask_n = need;
    }
    if (flow_pitch != mask_pitch) return fail(c, VP_ERR_UNSUPPORTED, "vp_flow_reliability: flow and mask planes must share a pitch");
    return launch_reliability(c, d_flow_x, d_flow_y, (int)(flow_pitch / 4), motion_threshold, d_mask, (int)(mask_pitch / 4), c->flow_tmp_mask, w, h, st);
}

int vp_add_weighted(vp_ctx* c, const uint8_t* d_a, size_t a_pitch, double alpha, const uint8_t* d_b, size_t b_pitch, double beta,
                    uint8_t* d_dst, size_t dst_pitch, int w, int h, void* stream) {
    if (!c) return VP_ERR_INVALID_ARG;
    int rc = check_frame(c, d_a, a_pitch, w, h, "vp_add_weighted(a)"); if (rc) return rc;
    rc = check_frame(c, d_b, b_pitch, w, h, "vp_add_weighted(b)"); if (rc) return rc;
    rc = check_frame(c, d_dst, dst_pitch, w, h, "vp_add_weighted(dst)"); if (rc) return rc;
    VP_CUDA(c, cudaSetDevice(c->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : c->s_compute;
    StageOrder stage_order(c, st);
    BlendKArgs a{};
    a.cur = d_a; a.cpitch = a_pitch; a.dst = d_dst; a.dpitch = dst_pitch; a.row_bytes = 3 * w; a.h = h;
    a.blend.mode = VP_TEMPORAL_MAIN_BLEND; a.blend.nprev = 1; a.blend.prev[0] = d_b; a.blend.ppitch = b_pitch;
    a.blend.w[0] = (float)alpha; a.blend.w[1] = (float)beta;
    return launch_blend(c, a, st);
}

int vp_temporal_blend(vp_ctx* c, const uint8_t* const* d_frames, int nframes, size_t pitch, float blend_strength,
                      uint8_t* d_dst, size_t dst_pitch, int w, int h, void* stream) {
    if (!c || !d_frames) return VP_ERR_INVALID_ARG;
    if (nframes < 1 || nframes > 4) return fail(c, VP_ERR_UNSUPPORTED, "vp_temporal_blend: 1..4 frames");
    int rc = VP_OK;
    for (int i = 0; i < nframes; ++i) { rc = check_frame(c, d_frames[i], pitch, w, h, "vp_temporal_blend(frame)"); if (rc) return rc; }
    rc = check_frame(c, d_dst, dst_pitch, w, h, "vp_temporal_blend(dst)"); if (rc) return rc;
    VP_CUDA(c, cudaSetDevice(c->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : c->s_compute;
    StageOrder stage_order(c, st);
    BlendKArgs a{};
    a.cur = d_frames[0]; a.cpitch = pitch; a.dst = d_dst; a.dpitch = dst_pitch; a.row_bytes = 3 * w; a.h = h;
    vp_temporal_config t{VP_TEMPORAL_BLEND_FRAMES, 3, blend_strength, 0.6f, 0, 100.0f};
    blend_weights(t, nframes - 1, a.blend);
    for (int i = 1; i < nframes; ++i) a.blend.prev[i - 1] = d_frames[i];
    a.blend.ppitch = pitch;
    if (nframes == 1) a.blend.mode = VP_TEMPORAL_NONE;
    return launch_blend(c, a, st);
}

}  // extern "C"
