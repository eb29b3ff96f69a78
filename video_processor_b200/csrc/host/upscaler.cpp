// upscaler.cpp - Upscaler (include/upscaler.h) over the fused chain of the C ABI.
// Mirrors the control flow of the reference's Upscaler (src/upscaler.cpp:552-1003): constructor,
// initialize(), upscale() with the per-stage toggles, algorithm names compared by callers
// (src/main.cpp:231-232), quality downgrade hook.  The chain context is created on the first frame
// (the input size is only known then) and rebuilt when the size, the toggles or a stage Config change.
#include "upscaler.h"

#include <cstring>

#include "adaptive_sharpening.h"
#include "host_common.h"
#include "selective_bilateral.h"
#include "temporal_consistency.h"

struct Upscaler::Chain {
    vp_ctx* ctx = nullptr;
    vp_chain_config cfg{};
    int device = 0;
    ~Chain() { if (ctx) vp_destroy(ctx); }
};

Upscaler::Upscaler(Algorithm algorithm, bool use_gpu)
    : m_algorithm(algorithm), m_use_gpu(use_gpu), m_initialized(false), m_target_width(0), m_target_height(0),
      m_use_selective_bilateral(true), m_use_adaptive_sharpening(true), m_use_temporal_consistency(true) {
    if (m_use_gpu && !isGPUAvailable())
        std::cout << "GPU acceleration requested but no B200 is visible; initialize() will fail (this build has no CPU path)." << std::endl;
}

Upscaler::~Upscaler() = default;

bool Upscaler::isGPUAvailable() { return vp_device_count() > 0; }
bool Upscaler::isUsingGPU() const { return m_use_gpu && m_initialized; }

std::string Upscaler::getAlgorithmName() const {
    switch (m_algorithm) {
        case NEAREST: return "Nearest Neighbor";
        case BILINEAR: return "Bilinear";
        case BICUBIC: return "Bicubic";
        case LANCZOS: return "Lanczos";
        case SUPER_RES: return "Standard Super-Res";
        case REAL_ESRGAN: return "RealESRGAN";
        default: return "Unknown";
    }
}

bool Upscaler::initialize(int target_width, int target_height) {
    if (target_width <= 0 || target_height <= 0) {
        std::cerr << "Invalid target resolution: " << target_width << "x" << target_height << std::endl;
        return false;
    }
    m_target_width = target_width;
    m_target_height = target_height;
    m_chain.reset();
    m_bilateral_pre.reset(); m_sharpening.reset(); m_bilateral_post.reset(); m_temporal_consistency.reset();
    m_use_selective_bilateral = m_use_adaptive_sharpening = m_use_temporal_consistency = true;
    return initializeImpl();
}

bool Upscaler::initializeImpl() {
    m_chain.reset();
    m_initialized = false;
    if (m_target_width <= 0 || m_target_height <= 0) {
        std::cerr << "Invalid target resolution" << std::endl;
        return false;
    }
    if (m_algorithm != BICUBIC && m_algorithm != REAL_ESRGAN) {
        std::cerr << "Upscaler: algorithm '" << getAlgorithmName() << "' is not on the B200 enhance-chain path "
                     "(built: Bicubic, RealESRGAN chain)" << std::endl;
        return false;
    }
    if (!m_use_gpu || !isGPUAvailable()) {
        std::cerr << "Upscaler: a B200 is required (use_gpu=" << m_use_gpu << ", devices=" << vp_device_count()
                  << "); there is no CPU fallback" << std::endl;
        return false;
    }
    if (m_algorithm == REAL_ESRGAN && !initializeEnhancements()) return false;
    std::cout << "Using B200 upscaling with " << getAlgorithmName() << std::endl;
    m_initialized = true;
    return true;
}

// The stage objects carry the Configs of Upscaler::initializeEnhancements (src/upscaler.cpp:666-758);
// they are exposed through the getters for tuning and read back when the chain is (re)built.
bool Upscaler::initializeEnhancements() {
    m_bilateral_pre = std::make_unique<SelectiveBilateral>(SelectiveBilateral::Config{
        SelectiveBilateral::PRE_PROCESSING, m_use_gpu, true, 7, 30.0, 30.0, true, 30.0, 1.5, 2.0, true, 3});
    m_sharpening = std::make_unique<AdaptiveSharpening>(AdaptiveSharpening::Config{
        0.8f, 1.2f, 0.4f, 30.0f, 1.5f, 5, true, m_use_gpu, true});
    m_bilateral_post = std::make_unique<SelectiveBilateral>(SelectiveBilateral::Config{
        SelectiveBilateral::POST_PROCESSING, m_use_gpu, true, 5, 20.0, 20.0, true, 30.0, 1.2, 1.5, true, 2});
    m_temporal_consistency = std::make_unique<TemporalConsistency>(TemporalConsistency::Config{
        3, 0.6f, 15.0f, 100.0f, m_use_gpu, 0.5, 3, 15, 3, 5, 1.2, 0});
    return true;
}

bool Upscaler::buildChainConfig(int src_w, int src_h, vp_chain_config* out) const {
    vp_chain_config& c = *out;
    c = vp_chain_config{};
    if (vp_default_chain_config(&c, src_w, src_h, m_target_width, m_target_height) != VP_OK) return false;
    if (m_algorithm == BICUBIC) {
        // CPUImpl::upscale for BICUBIC (src/upscaler.cpp:75-84): pre-blur, cv::resize(INTER_CUBIC),
        // enhanceBicubicResult; the temporal smoothing of main.cpp:269-292 stays with the caller
        c.use_pre = c.use_sharpen = c.use_post = 0;
        c.bicubic_enhance = 1;
        c.temporal.mode = VP_TEMPORAL_NONE;
    } else {
        const SelectiveBilateral::Config pre = m_bilateral_pre->getConfig(), post = m_bilateral_post->getConfig();
        const AdaptiveSharpening::Config sh = m_sharpening->getConfig();
        const TemporalConsistency::Config tc = m_temporal_consistency->getConfig();
        c.use_pre = c.use_post = m_use_selective_bilateral ? 1 : 0;
        c.use_sharpen = m_use_adaptive_sharpening ? 1 : 0;
        c.pre = vp_bilateral_config{pre.stage == SelectiveBilateral::POST_PROCESSING ? VP_STAGE_POST : VP_STAGE_PRE,
                                    pre.adaptive_params, pre.diameter, 0, pre.sigma_color, pre.sigma_space,
                                    pre.selective, pre.use_multiscale, pre.num_scales, 0, pre.detail_threshold, pre.edge_preserve};
        c.post = vp_bilateral_config{post.stage == SelectiveBilateral::POST_PROCESSING ? VP_STAGE_POST : VP_STAGE_PRE,
                                     post.adaptive_params, post.diameter, 0, post.sigma_color, post.sigma_space,
                                     post.selective, post.use_multiscale, post.num_scales, 0, post.detail_threshold, post.edge_preserve};
        c.sharpen = vp_sharpen_config{sh.strength, sh.edge_strength, sh.smooth_strength, sh.edge_threshold, sh.sigma,
                                      sh.kernel_size, sh.preserve_tone ? 1 : 0, sh.adaptive_sigma ? 1 : 0};
        c.temporal.mode = m_use_temporal_consistency ? VP_TEMPORAL_BLEND_FRAMES : VP_TEMPORAL_NONE;
        c.temporal.buffer_size = std::min(std::max(tc.buffer_size, 1), 3);
        c.temporal.blend_strength = tc.blend_strength;
        c.temporal.scene_detect = 1;                       // TemporalConsistency::process always checks for a cut (:96-105)
        c.temporal.scene_change_threshold = tc.scene_change_threshold;
        c.temporal.use_flow = 1;                           // ... and warps every buffered frame along its Farneback flow (:94-150)
        c.temporal.motion_threshold = tc.motion_threshold;
        c.temporal.pyr_scale = tc.pyr_scale; c.temporal.poly_sigma = tc.poly_sigma;
        c.temporal.levels = tc.levels; c.temporal.winsize = tc.winsize; c.temporal.iterations = tc.iterations;
        c.temporal.poly_n = tc.poly_n; c.temporal.flags = tc.flags;
    }
    return true;
}

bool Upscaler::ensureChain(int src_w, int src_h) {
    vp_chain_config c{};
    if (!buildChainConfig(src_w, src_h, &c)) return false;
    if (m_chain && m_chain->device == m_device && std::memcmp(&m_chain->cfg, &c, sizeof(c)) == 0) return true;
    if (m_chain && m_chain->ctx) {          // drain frames in flight before the context is replaced
        while (vp_wait(m_chain->ctx) == VP_OK) {}
    }
    auto chain = std::make_unique<Chain>();
    chain->cfg = c;
    chain->device = m_device;
    if (vp_create(&chain->ctx, m_device, &c) != VP_OK) {
        std::cerr << "Upscaler: cannot create the B200 chain: " << vp_last_error(nullptr) << std::endl;
        return false;
    }
    m_chain = std::move(chain);
    return true;
}

bool Upscaler::upscale(const cv::Mat& input, cv::Mat& output) {
    if (!m_initialized) {
        std::cerr << "Upscaler not initialized" << std::endl;
        return false;
    }
    if (input.empty()) {
        std::cerr << "Input frame is empty" << std::endl;
        return false;
    }
    if (!vp_host_api::is_bgr8(input)) {
        std::cerr << "Upscaler: input must be CV_8UC3" << std::endl;
        return false;
    }
    if (!ensureChain(input.cols, input.rows)) return false;
    vp::createPinned(output, m_target_height, m_target_width, CV_8UC3);
    const int rc = vp_process_host(m_chain->ctx, input.data, input.step, output.data, output.step);
    if (rc != VP_OK) {
        std::cerr << "Error in enhanced upscaling: " << vp_last_error(m_chain->ctx) << std::endl;
        return false;
    }
    return true;
}

int Upscaler::pipelineDepth() const { return m_chain && m_chain->ctx ? vp_pipeline_depth(m_chain->ctx) : 3; }

bool Upscaler::submit(const cv::Mat& input, cv::Mat& output) {
    if (!m_initialized || input.empty() || !vp_host_api::is_bgr8(input)) return false;
    if (!ensureChain(input.cols, input.rows)) return false;
    vp::createPinned(output, m_target_height, m_target_width, CV_8UC3);
    return vp_submit_host(m_chain->ctx, input.data, input.step, output.data, output.step) == VP_OK;
}

bool Upscaler::wait() { return m_chain && m_chain->ctx && vp_wait(m_chain->ctx) == VP_OK; }

bool Upscaler::warmup(const cv::Mat& input) {
    if (!m_initialized || input.empty() || !vp_host_api::is_bgr8(input)) return false;
    if (!ensureChain(input.cols, input.rows)) return false;
    return vp_submit_host(m_chain->ctx, input.data, input.step, nullptr, 0) == VP_OK && vp_wait(m_chain->ctx) == VP_OK;
}

void Upscaler::setAlgorithm(Algorithm algorithm) {
    if (m_algorithm != algorithm) {
        m_algorithm = algorithm;
        if (m_initialized) initializeImpl();
    }
}

bool Upscaler::setUseGPU(bool use_gpu) {
    if (use_gpu && !isGPUAvailable()) {
        std::cerr << "GPU acceleration requested but not available" << std::endl;
        return false;
    }
    if (m_use_gpu != use_gpu) {
        m_use_gpu = use_gpu;
        if (m_initialized) initializeImpl();
    }
    return true;
}

// src/upscaler.cpp:994-1003: SUPER_RES -> BICUBIC when a frame takes > 1.5x its budget
bool Upscaler::adjustQualityForPerformance(double processing_time, double target_time) {
    if (m_algorithm == SUPER_RES && processing_time > target_time * 1.5) {
        std::cout << "Performance warning: Super-resolution taking too long (" << processing_time
                  << "ms). Switching to Bicubic upscaling." << std::endl;
        m_algorithm = BICUBIC;
        initializeImpl();
        return true;
    }
    return false;
}
