// stages.cpp - SelectiveBilateral / AdaptiveSharpening / TemporalConsistency (include/*.h) over the
// stage-wise entry points of the C ABI.  Error policy follows the reference: bool returns, nothing
// throws; on a stage failure the input is copied through and false is returned
// (src/selective_bilateral.cpp:77-81, src/adaptive_sharpening.cpp:100-104).
#include <deque>
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
        std::cerr << "SelectiveBilateral: multiscale/selective modes run the plain bilateral branch "
                     "(the reference's own bodies are undefined behaviour, see DESIGN.md)" << std::endl;
        m_impl->warned = true;
    }
    StageRunner& r = m_impl->run;
    vp_bilateral_config c{};
    c.stage = m_config.stage == POST_PROCESSING ? VP_STAGE_POST : VP_STAGE_PRE;
    c.adaptive_params = m_config.adaptive_params;
    c.diameter = m_config.diameter;
    c.sigma_color = m_config.sigma_color;
    c.sigma_space = m_config.sigma_space;
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
        std::cerr << "AdaptiveSharpening: adaptive_sigma runs the unsharp-mask branch (the reference's "
                     "variable-sigma body is undefined behaviour, see DESIGN.md)" << std::endl;
        m_impl->warned = true;
    }
    StageRunner& r = m_impl->run;
    vp_sharpen_config c{m_config.strength, m_config.edge_strength, m_config.smooth_strength, m_config.edge_threshold,
                        m_config.sigma, m_config.kernel_size, m_config.preserve_tone ? 1 : 0};
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
// TemporalConsistency
// ------------------------------------------------------------------------------------------------
struct TemporalConsistency::Impl {
    StageRunner run;
    std::deque<uint8_t*> frames;          // device copies of the previous unblended inputs, newest last
    size_t pitch = 0;
    int w = 0, h = 0;
    void clear() { for (uint8_t* p : frames) cudaFree(p); frames.clear(); }
    ~Impl() { clear(); }
};

TemporalConsistency::TemporalConsistency() : m_impl(new Impl) {}
TemporalConsistency::TemporalConsistency(const Config& config) : m_config(config), m_impl(new Impl) {}
TemporalConsistency::~TemporalConsistency() = default;

bool TemporalConsistency::initialize() {
    std::string err;
    if (!m_impl->run.create(0, &err)) {
        std::cerr << "TemporalConsistency: no B200 context (" << err << "); this build has no CPU path" << std::endl;
        m_initialized = false;
        return false;
    }
    reset();
    m_initialized = true;
    return true;
}

void TemporalConsistency::reset() {
    std::lock_guard<std::mutex> lock(m_buffer_mutex);
    m_impl->clear();
}

void TemporalConsistency::setConfig(const Config& config) { m_config = config; }
TemporalConsistency::Config TemporalConsistency::getConfig() const { return m_config; }

bool TemporalConsistency::process(const cv::Mat& input, cv::Mat& output) {
    if (!m_initialized) {
        std::cerr << "Temporal consistency module not initialized" << std::endl;
        return false;
    }
    if (input.empty()) {
        std::cerr << "Empty input image" << std::endl;
        return false;
    }
    std::lock_guard<std::mutex> lock(m_buffer_mutex);
    Impl& s = *m_impl;
    StageRunner& r = s.run;
    if (!vp_host_api::is_bgr8(input) || !r.upload(input, input.cols, input.rows, 3)) {
        input.copyTo(output);
        return false;
    }
    if (s.w != input.cols || s.h != input.rows) { s.clear(); s.w = input.cols; s.h = input.rows; s.pitch = r.in_pitch(); }
    const int buffer = std::min(std::max(m_config.buffer_size, 1), 3);
    const int nprev = std::min<int>((int)s.frames.size(), buffer - 1);
    const uint8_t* ptrs[3] = {r.d_in(), nullptr, nullptr};
    for (int i = 0; i < nprev; ++i) ptrs[1 + i] = s.frames[s.frames.size() - 1 - i];
    bool ok = vp_temporal_blend(r.ctx(), ptrs, 1 + nprev, s.pitch, m_config.blend_strength, r.d_out(), r.out_pitch(),
                                input.cols, input.rows, r.stream()) == VP_OK;
    ok = ok && r.download(output, input.cols, input.rows, CV_8UC3, 3);
    // keep the unblended input (src/temporal_consistency.cpp:158-159), trim to buffer_size (:162-169)
    uint8_t* keep = nullptr;
    if ((int)s.frames.size() >= buffer) { keep = s.frames.front(); s.frames.pop_front(); }
    if (!keep && cudaMalloc(&keep, s.pitch * input.rows) != cudaSuccess) { cudaGetLastError(); keep = nullptr; }
    if (keep) {
        cudaMemcpy2DAsync(keep, s.pitch, r.d_in(), s.pitch, (size_t)input.cols * 3, input.rows, cudaMemcpyDeviceToDevice, (cudaStream_t)r.stream());
        cudaStreamSynchronize((cudaStream_t)r.stream());
        s.frames.push_back(keep);
    }
    if (!ok) {
        std::cerr << "Error in temporal consistency: " << r.error() << std::endl;
        input.copyTo(output);
    }
    return ok;
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
