// stages.cpp - SelectiveBilateral / AdaptiveSharpening / TemporalConsistency (include/*.h) over the
// stage-wise entry points of the C ABI.  Error policy follows the reference: bool returns, nothing
// throws; on a stage failure the input is copied through and false is returned
// (src/selective_bilateral.cpp:77-81, src/adaptive_sharpening.cpp:100-104).
#include <cstring>
#include <mutex>

#include "adaptive_sharpening.h"
#include "host_common.h"
#include "selective_bilateral.h"
#include "temporal_consistency.h"
#include "vp/imgproc.h"

using vp_host_api::StageRunner;

// ------------------------------------------------------------------------------------------------
// SelectiveBilateral
// ------------------------------------------------------------------------------------------------
struct SelectiveBilateral::Impl { StageRunner run; bool warned = false; };

SelectiveBilateral::SelectiveBilateral() : m_impl(new Impl) {}
SelectiveBilateral::SelectiveBilateral(const Config& config) : m_config(config), m_impl(new Impl) {}
SelectiveBilateral::~SelectiveBilateral() = default;

bool SelectiveBilateral::initialize() {
    if (m_config.use_multiscale) m_config.num_scales = std::max(1, std::min(m_config.num_scales, 5));
    std::string err;
    if (!m_impl->run.create(0, &err)) {
        std::cerr << "SelectiveBilateral: no B200 context (" << err << "); this build has no CPU path" << std::endl;
        m_initialized = false;
        return false;
    }
    m_initialized = true;
    return true;
}

void SelectiveBilateral::setConfig(const Config& config) { m_config = config; }
SelectiveBilateral::Config SelectiveBilateral::getConfig() const { return m_config; }

bool SelectiveBilateral::process(const cv::Mat& input, cv::Mat& output) {
    if (!m_initialized) {
        std::cerr << "Selective bilateral filtering module not initialized" << std::endl;
        return false;
    }
    if (input.empty()) {
        std::cerr << "Empty input image" << std::endl;
        return false;
    }
    if ((m_config.use_multiscale || m_config.selective) && !m_impl->warned) {
        std::cerr << "SelectiveBilateral: the multiscale/selective branches run with repaired semantics (the reference "
                     "calls cv::sqrt on a CV_8U Mat, src/selective_bilateral.cpp:276; see DESIGN.md)" << std::endl;
        m_impl->warned = true;
    }
    StageRunner& r = m_impl->run;
    vp_bilateral_config c{};
    c.stage = m_config.stage == POST_PROCESSING ? VP_STAGE_POST : VP_STAGE_PRE;
    c.adaptive_params = m_config.adaptive_params;
    c.diameter = m_config.diameter;
    c.sigma_color = m_config.sigma_color;
    c.sigma_space = m_config.sigma_space;
    c.selective = m_config.selective;
    c.use_multiscale = m_config.use_multiscale;
    c.num_scales = m_config.num_scales;
    c.detail_threshold = m_config.detail_threshold;
    c.edge_preserve = m_config.edge_preserve;
    bool ok = vp_host_api::is_bgr8(input) && r.upload(input, input.cols, input.rows, 3);
    ok = ok && vp_selective_bilateral(r.ctx(), &c, r.d_in(), r.in_pitch(), r.d_out(), r.out_pitch(), input.cols, input.rows, r.stream()) == VP_OK;
    ok = ok && r.download(output, input.cols, input.rows, CV_8UC3, 3);
    if (!ok) {
        std::cerr << "Error in selective bilateral filtering: " << r.error() << std::endl;
        input.copyTo(output);
    }
    return ok;
}

// ------------------------------------------------------------------------------------------------
// AdaptiveSharpening
// ------------------------------------------------------------------------------------------------
struct AdaptiveSharpening::Impl { StageRunner run; bool warned = false; };

AdaptiveSharpening::AdaptiveSharpening() : m_impl(new Impl) {}
AdaptiveSharpening::AdaptiveSharpening(const Config& config) : m_config(config), m_impl(new Impl) {}
AdaptiveSharpening::~AdaptiveSharpening() = default;

bool AdaptiveSharpening::initialize() {
    std::string err;
    if (!m_impl->run.create(0, &err)) {
        std::cerr << "AdaptiveSharpening: no B200 context (" << err << "); this build has no CPU path" << std::endl;
        m_initialized = false;
        return false;
    }
    m_initialized = true;
    return true;
}

void AdaptiveSharpening::setConfig(const Config& config) { m_config = config; }
AdaptiveSharpening::Config AdaptiveSharpening::getConfig() const { return m_config; }

bool AdaptiveSharpening::process(const cv::Mat& input, cv::Mat& output) {
    if (!m_initialized) {
        std::cerr << "Adaptive sharpening module not initialized" << std::endl;
        return false;
    }
    if (input.empty()) {
        std::cerr << "Empty input image" << std::endl;
        return false;
    }
    if (m_config.adaptive_sigma && !m_impl->warned) {
        std::cerr << "AdaptiveSharpening: the adaptive_sigma branch runs with repaired semantics (the reference calls "
                     "cv::sqrt on a CV_8U Mat, src/adaptive_sharpening.cpp:386; see DESIGN.md)" << std::endl;
        m_impl->warned = true;
    }
    StageRunner& r = m_impl->run;
    vp_sharpen_config c{m_config.strength, m_config.edge_strength, m_config.smooth_strength, m_config.edge_threshold,
                        m_config.sigma, m_config.kernel_size, m_config.preserve_tone ? 1 : 0, m_config.adaptive_sigma ? 1 : 0};
    bool ok = vp_host_api::is_bgr8(input) && r.upload(input, input.cols, input.rows, 3);
    ok = ok && vp_sharpen(r.ctx(), &c, r.d_in(), r.in_pitch(), r.d_out(), r.out_pitch(), input.cols, input.rows, r.stream()) == VP_OK;
    ok = ok && r.download(output, input.cols, input.rows, CV_8UC3, 3);
    if (!ok) {
        std::cerr << "Error in adaptive sharpening: " << r.error() << std::endl;
        input.copyTo(output);
    }
    return ok;
}

// ------------------------------------------------------------------------------------------------
// TemporalConsistency: a chain context with every spatial stage off - buffer policy, scene-change reset
// and blendFrames all run on the device (src/temporal_consistency.cpp:61-172, :283-323, :355-444)
// ------------------------------------------------------------------------------------------------
struct TemporalConsistency::Impl {
    vp_ctx* ctx = nullptr;
    vp_chain_config cfg{};
    bool available = false;
    ~Impl() { if (ctx) vp_destroy(ctx); }
};

TemporalConsistency::TemporalConsistency() : m_impl(new Impl) {}
TemporalConsistency::TemporalConsistency(const Config& config) : m_config(config), m_impl(new Impl) {}
TemporalConsistency::~TemporalConsistency() = default;

bool TemporalConsistency::initialize() {
    m_impl->available = vp_device_count() > 0;
    if (!m_impl->available) {
        std::cerr << "TemporalConsistency: no B200 context (vp_create: no CUDA device); this build has no CPU path" << std::endl;
        m_initialized = false;
        return false;
    }
    reset();
    m_initialized = true;
    return true;
}

void TemporalConsistency::reset() {
    std::lock_guard<std::mutex> lock(m_buffer_mutex);
    if (m_impl->ctx) vp_reset_temporal(m_impl->ctx);
}

void TemporalConsistency::setConfig(const Config& config) { m_config = config; }
TemporalConsistency::Config TemporalConsistency::getConfig() const { return m_config; }

bool TemporalConsistency::process(const cv::Mat& input, cv::Mat& output) {
    if (!m_initialized) {
        std::cerr << "Temporal consistency module not initialized" << std::endl;
        if (!input.empty()) input.copyTo(output);
        return false;
    }
    if (input.empty()) {
        std::cerr << "Empty input frame" << std::endl;
        return false;
    }
    std::lock_guard<std::mutex> lock(m_buffer_mutex);
    vp_chain_config c{};
    vp_default_chain_config(&c, input.cols, input.rows, input.cols, input.rows);
    c.use_pre = c.use_sharpen = c.use_post = 0;
    c.temporal.mode = VP_TEMPORAL_BLEND_FRAMES;
    c.temporal.buffer_size = std::min(std::max(m_config.buffer_size, 1), 3);
    c.temporal.blend_strength = m_config.blend_strength;
    c.temporal.scene_detect = 1;
    c.temporal.scene_change_threshold = m_config.scene_change_threshold;
    // the motion-compensated front end, :94-150 (flow, warp, reliability masks)
    c.temporal.use_flow = 1;
    c.temporal.motion_threshold = m_config.motion_threshold;
    c.temporal.pyr_scale = m_config.pyr_scale;
    c.temporal.poly_sigma = m_config.poly_sigma;
    c.temporal.levels = m_config.levels;
    c.temporal.winsize = m_config.winsize;
    c.temporal.iterations = m_config.iterations;
    c.temporal.poly_n = m_config.poly_n;
    c.temporal.flags = m_config.flags;
    Impl& s = *m_impl;
    if (!s.ctx || std::memcmp(&c, &s.cfg, sizeof(c)) != 0) {
        if (s.ctx) { vp_destroy(s.ctx); s.ctx = nullptr; }
        if (vp_create(&s.ctx, 0, &c) != VP_OK) {
            std::cerr << "Error in temporal consistency: " << vp_last_error(nullptr) << std::endl;
            input.copyTo(output);
            return false;
        }
        s.cfg = c;
    }
    vp::createPinned(output, input.rows, input.cols, CV_8UC3);
    if (!vp_host_api::is_bgr8(input) ||
        vp_process_host(s.ctx, input.data, input.step, output.data, output.step) != VP_OK) {
        std::cerr << "Error in temporal consistency: " << vp_last_error(s.ctx) << std::endl;
        input.copyTo(output);
        return false;
    }
    return true;
}

// ------------------------------------------------------------------------------------------------
// vp::resizeCubic - stand-in for cv::resize(..., INTER_CUBIC) in OpenCV-less builds
// ------------------------------------------------------------------------------------------------
namespace vp {
bool resizeCubic(const cv::Mat& src, cv::Mat& dst, cv::Size dsize) {
    static thread_local StageRunner runner;       // one context per calling thread
    if (src.empty() || !vp_host_api::is_bgr8(src) || dsize.width < src.cols || dsize.height < src.rows) return false;
    std::string err;
    if (!runner.ready() && !runner.create(0, &err)) {
        std::cerr << "vp::resizeCubic: no B200 context (" << err << ")" << std::endl;
        return false;
    }
    bool ok = runner.upload(src, dsize.width, dsize.height, 3);
    ok = ok && vp_bicubic(runner.ctx(), runner.d_in(), runner.in_pitch(), src.cols, src.rows, runner.d_out(), runner.out_pitch(),
                          dsize.width, dsize.height, runner.stream()) == VP_OK;
    return ok && runner.download(dst, dsize.width, dsize.height, CV_8UC3, 3);
}
}  // namespace vp
