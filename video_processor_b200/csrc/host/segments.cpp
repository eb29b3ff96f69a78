// segments.cpp - file / --fast mode over several B200s: vp_plan_segment, vp_run_segments, vp_run_segments_cb
// (include/vpb200.h) and Upscaler::upscaleSegments (include/upscaler.h).
//
// The reference's --fast loop (src/main.cpp:222-355 with simulate_realtime = false, :627-629) walks the clip
// frame by frame on one device.  Frames are independent except for the temporal stage, which looks back a
// fixed number of UNBLENDED frames (src/main.cpp:277: one; TemporalConsistency with buffer_size b: b,
// src/temporal_consistency.cpp:127 with :97 and :162-169), so the clip is cut into contiguous segments, one per
// device, each preceded by that many warm-up frames (fed with no output).  One host thread and one vp_ctx per
// device, frames pipelined through vp_submit_host / vp_wait; there is no exchange between devices, hence no
// collective.  The result is bit-identical to the one-device run (tests/test_gpu_segments.py).
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstring>
#include <iostream>
#include <string>
#include <thread>
#include <vector>

#include "upscaler.h"
#include "vpb200.h"

namespace {

thread_local std::string g_segments_error;

int warmup_frames(const vp_temporal_config& t) {
    if (t.mode == VP_TEMPORAL_NONE) return 0;
    if (t.mode == VP_TEMPORAL_MAIN_BLEND) return 1;
    return t.buffer_size > 0 ? t.buffer_size : 0;
}

struct Plan { int first, start, stop; };

Plan plan(int nframes, int nseg, int k, const vp_temporal_config& t) {
    const int start = (int)((long long)k * nframes / nseg), stop = (int)((long long)(k + 1) * nframes / nseg);
    const int w = std::min(warmup_frames(t), start);
    return Plan{start - w, start, stop};
}

struct Job {
    const vp_chain_config* cfg;
    int device;
    Plan p;
    // pointer-array form
    const uint8_t* const* src = nullptr; size_t src_step = 0;
    uint8_t* const* dst = nullptr; size_t dst_step = 0;
    // callback form
    vp_frame_source source = nullptr; vp_frame_sink sink = nullptr; void* user = nullptr;
    int status = VP_OK;
    std::string err;
    double seconds = 0.0;
};

void run_job(Job* j, std::atomic<int>* abort_all) {
    if (j->p.stop <= j->p.first) return;
    vp_ctx* ctx = nullptr;
    int rc = vp_create(&ctx, j->device, j->cfg);
    if (rc != VP_OK) { j->status = rc; j->err = vp_last_error(nullptr); abort_all->store(1); return; }
    const auto t0 = std::chrono::steady_clock::now();
    const int depth = vp_pipeline_depth(ctx);
    const bool cb = j->source != nullptr;
    const size_t in_step = (size_t)j->cfg->src_w * 3, out_step = (size_t)j->cfg->dst_w * 3;
    std::vector<uint8_t*> hin, hout;
    auto fail = [&](int code, const char* what) { j->status = code; j->err = what; abort_all->store(1); };
    if (cb) {
        // a slot's buffers are reused only after vp_wait has returned that slot's frame: depth slots suffice
        for (int s = 0; s < depth && j->status == VP_OK; ++s) {
            void *a = nullptr, *b = nullptr;
            if (vp_host_alloc(&a, in_step * j->cfg->src_h) != VP_OK || vp_host_alloc(&b, out_step * j->cfg->dst_h) != VP_OK)
                fail(VP_ERR_NOMEM, "vp_run_segments_cb: pinned allocation failed");
            hin.push_back((uint8_t*)a); hout.push_back((uint8_t*)b);
        }
    }
    int inflight = 0, oldest = j->p.first;
    bool wait_failed = false;
    auto drain_one = [&]() {               // collects the oldest frame in flight; delivers it unless the job already failed
        rc = vp_wait(ctx);
        if (rc != VP_OK) { fail(rc, vp_last_error(ctx)); wait_failed = true; return; }
        if (cb && j->status == VP_OK && oldest >= j->p.start && j->sink) {
            const int s = (oldest - j->p.first) % depth;
            if (j->sink(j->user, oldest, hout[s], out_step) != 0) fail(VP_ERR_STATE, "vp_run_segments_cb: the sink reported an error");
        }
        ++oldest; --inflight;
    };
    for (int i = j->p.first; i < j->p.stop && j->status == VP_OK && !abort_all->load(); ++i) {
        if (inflight == depth) drain_one();
        if (j->status != VP_OK) break;
        const bool out = i >= j->p.start;
        if (cb) {
            const int s = (i - j->p.first) % depth;
            if (j->source(j->user, i, hin[s], in_step) != 0) { fail(VP_ERR_STATE, "vp_run_segments_cb: the source reported an error"); break; }
            rc = vp_submit_host(ctx, hin[s], in_step, out ? hout[s] : nullptr, out_step);
        } else {
            rc = vp_submit_host(ctx, j->src[i], j->src_step, out ? j->dst[i] : nullptr, j->dst_step);
        }
        if (rc != VP_OK) { fail(rc, vp_last_error(ctx)); break; }
        ++inflight;
    }
    // frames already submitted are collected even after a failure: their copies target the caller's memory
    while (inflight > 0 && !wait_failed) drain_one();
    vp_synchronize(ctx);
    j->seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    for (uint8_t* p : hin) vp_host_free(p);
    for (uint8_t* p : hout) vp_host_free(p);
    vp_destroy(ctx);
}

int run_all(std::vector<Job>& jobs, vp_segment_report* report) {
    std::atomic<int> abort_all{0};
    std::vector<std::thread> th;
    for (size_t k = 1; k < jobs.size(); ++k) th.emplace_back(run_job, &jobs[k], &abort_all);
    run_job(&jobs[0], &abort_all);            // segment 0 runs on the caller's thread
    for (std::thread& t : th) t.join();
    int rc = VP_OK;
    for (size_t k = 0; k < jobs.size(); ++k) {
        if (report) {
            report[k].device = jobs[k].device; report[k].first_frame = jobs[k].p.first; report[k].start = jobs[k].p.start;
            report[k].stop = jobs[k].p.stop; report[k].status = jobs[k].status; report[k].reserved_ = 0; report[k].seconds = jobs[k].seconds;
        }
        if (jobs[k].status != VP_OK && rc == VP_OK) {
            rc = jobs[k].status;
            g_segments_error = "segment " + std::to_string(k) + " (device " + std::to_string(jobs[k].device) + "): " + jobs[k].err;
        }
    }
    return rc;
}

int check_common(const vp_chain_config* cfg, const int* devices, int ndev, int nframes) {
    if (!cfg || !devices || ndev < 1 || nframes < 0) { g_segments_error = "vp_run_segments: null config / device list, ndev < 1 or nframes < 0"; return VP_ERR_INVALID_ARG; }
    const int have = vp_device_count();
    if (have <= 0) { g_segments_error = "vp_run_segments: no sm_100 device (this library has no CPU fallback)"; return VP_ERR_NO_DEVICE; }
    for (int k = 0; k < ndev; ++k)
        if (devices[k] < 0) { g_segments_error = "vp_run_segments: negative device index"; return VP_ERR_INVALID_ARG; }
    return VP_OK;
}

}  // namespace

extern "C" {

const char* vp_segments_last_error(void) { return g_segments_error.c_str(); }

int vp_plan_segment(int nframes, int nsegments, int k, const vp_temporal_config* temporal, int* first, int* start, int* stop) {
    if (nframes < 0 || nsegments < 1 || k < 0 || k >= nsegments || !temporal) return VP_ERR_INVALID_ARG;
    const Plan p = plan(nframes, nsegments, k, *temporal);
    if (first) *first = p.first;
    if (start) *start = p.start;
    if (stop) *stop = p.stop;
    return VP_OK;
}

int vp_run_segments(const vp_chain_config* cfg, const int* devices, int ndev, int nframes,
                    const uint8_t* const* src, size_t src_step, uint8_t* const* dst, size_t dst_step,
                    vp_segment_report* report) {
    int rc = check_common(cfg, devices, ndev, nframes);
    if (rc) return rc;
    if (nframes && (!src || !dst)) { g_segments_error = "vp_run_segments: null frame arrays"; return VP_ERR_INVALID_ARG; }
    for (int i = 0; i < nframes; ++i)
        if (!src[i] || !dst[i]) { g_segments_error = "vp_run_segments: null frame pointer"; return VP_ERR_INVALID_ARG; }
    std::vector<Job> jobs(ndev);
    for (int k = 0; k < ndev; ++k) {
        jobs[k].cfg = cfg; jobs[k].device = devices[k]; jobs[k].p = plan(nframes, ndev, k, cfg->temporal);
        jobs[k].src = src; jobs[k].src_step = src_step; jobs[k].dst = dst; jobs[k].dst_step = dst_step;
    }
    return run_all(jobs, report);
}

int vp_run_segments_cb(const vp_chain_config* cfg, const int* devices, int ndev, int nframes,
                       vp_frame_source source, vp_frame_sink sink, void* user, vp_segment_report* report) {
    int rc = check_common(cfg, devices, ndev, nframes);
    if (rc) return rc;
    if (!source) { g_segments_error = "vp_run_segments_cb: null source"; return VP_ERR_INVALID_ARG; }
    std::vector<Job> jobs(ndev);
    for (int k = 0; k < ndev; ++k) {
        jobs[k].cfg = cfg; jobs[k].device = devices[k]; jobs[k].p = plan(nframes, ndev, k, cfg->temporal);
        jobs[k].source = source; jobs[k].sink = sink; jobs[k].user = user;
    }
    return run_all(jobs, report);
}

}  // extern "C"

// ---- C++ face: Upscaler::upscaleSegments -----------------------------------------------------------------------
bool Upscaler::upscaleSegments(const std::vector<cv::Mat>& inputs, std::vector<cv::Mat>& outputs, const std::vector<int>& devices) {
    if (!m_initialized) { std::cerr << "Upscaler not initialized" << std::endl; return false; }
    if (inputs.empty()) { outputs.clear(); return true; }
    std::vector<int> dev = devices;
    if (dev.empty()) for (int d = 0; d < vp_device_count(); ++d) dev.push_back(d);
    if (dev.empty()) { std::cerr << "Upscaler::upscaleSegments: no B200 visible (there is no CPU fallback)" << std::endl; return false; }
    const cv::Mat& f0 = inputs[0];
    for (const cv::Mat& m : inputs)
        if (m.empty() || m.type() != CV_8UC3 || m.cols != f0.cols || m.rows != f0.rows || m.step != f0.step) {
            std::cerr << "Upscaler::upscaleSegments: every frame must be a CV_8UC3 Mat of one size and step" << std::endl;
            return false;
        }
    vp_chain_config c{};
    if (!buildChainConfig(f0.cols, f0.rows, &c)) return false;
    outputs.resize(inputs.size());
    std::vector<const uint8_t*> src(inputs.size());
    std::vector<uint8_t*> dst(inputs.size());
    for (size_t i = 0; i < inputs.size(); ++i) {
        vp::createPinned(outputs[i], m_target_height, m_target_width, CV_8UC3);
        src[i] = inputs[i].data; dst[i] = outputs[i].data;
    }
    const int rc = vp_run_segments(&c, dev.data(), (int)dev.size(), (int)inputs.size(), src.data(), f0.step, dst.data(), outputs[0].step, nullptr);
    if (rc != VP_OK) { std::cerr << "Upscaler::upscaleSegments: " << vp_segments_last_error() << std::endl; return false; }
    return true;
}
