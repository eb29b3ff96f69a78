// misc.cpp - Processor, VideoEnhancer and FrameBuffer (include/processor.h, video_enhancer.h,
// frame_buffer.h).  Host control code; device work happens in the classes that own a vp_ctx.
#include <chrono>

#include "adaptive_sharpening.h"
#include "frame_buffer.h"
#include "host_common.h"
#include "processor.h"
#include "selective_bilateral.h"
#include "upscaler.h"
#include "video_enhancer.h"

// ------------------------------------------------------------------------------------------------
// Processor: ordered plugin chain (reference src/processor.cpp:32-58, :173-193)
// ------------------------------------------------------------------------------------------------
class Processor::ProcessorImpl {
public:
    // out = in; for each enabled op: op(out, tmp), out = tmp   (what CPUProcessorImpl::process does)
    bool run(const cv::Mat& input, cv::Mat& output, std::vector<Operation>& ops) {
        cv::Mat cur = input;                // refcounted view; ops must not write their input
        for (Operation& op : ops) {
            if (!op.enabled) continue;
            cv::Mat next;
            op.func(cur, next);
            if (next.empty()) return false;
            cur = next;
        }
        if (cur.data == input.data) input.copyTo(output); else output = cur;
        return true;
    }
};

Processor::Processor(bool use_gpu)
    : m_use_gpu(use_gpu), m_initialized(false), m_last_processing_time(0.0), m_impl(new ProcessorImpl) {}
Processor::~Processor() = default;

bool Processor::isGPUAvailable() { return vp_device_count() > 0; }
bool Processor::isUsingGPU() const { return m_use_gpu; }
double Processor::getLastProcessingTime() const { return m_last_processing_time; }

bool Processor::initialize() {
    if (m_use_gpu && !isGPUAvailable()) {
        std::cerr << "Processor: GPU requested but no B200 is visible" << std::endl;
        m_initialized = false;
        return false;
    }
    m_initialized = true;
    return true;
}

bool Processor::process(const cv::Mat& input, cv::Mat& output) {
    if (!m_initialized) {
        std::cerr << "Processor not initialized" << std::endl;
        return false;
    }
    if (input.empty()) {
        std::cerr << "Input frame is empty" << std::endl;
        return false;
    }
    std::lock_guard<std::mutex> lock(m_mutex);
    const auto t0 = std::chrono::steady_clock::now();
    const bool ok = m_impl->run(input, output, m_operations);
    m_last_processing_time = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    return ok;
}

Processor& Processor::addOperation(const std::string& name, ProcessFunction func) {
    std::lock_guard<std::mutex> lock(m_mutex);
    m_operations.emplace_back(name, func);
    return *this;
}

void Processor::enableOperation(const std::string& name, bool enabled) {
    std::lock_guard<std::mutex> lock(m_mutex);
    for (Operation& op : m_operations)
        if (op.name == name) { op.enabled = enabled; return; }
    std::cerr << "Operation not found: " << name << std::endl;
}

bool Processor::setUseGPU(bool use_gpu) {
    if (use_gpu && !isGPUAvailable()) {
        std::cerr << "GPU acceleration requested but not available" << std::endl;
        return false;
    }
    m_use_gpu = use_gpu;
    return true;
}

Processor& Processor::addDefaultPreProcessing() {
    auto filter = std::make_shared<SelectiveBilateral>(SelectiveBilateral::Config{
        SelectiveBilateral::PRE_PROCESSING, true, true, 7, 30.0, 30.0, false, 30.0, 1.5, 2.0, false, 3});
    filter->initialize();
    return addOperation("denoise", [filter](const cv::Mat& in, cv::Mat& out) { filter->process(in, out); });
}

Processor& Processor::addDefaultPostProcessing() {
    auto sharpen = std::make_shared<AdaptiveSharpening>(AdaptiveSharpening::Config{
        0.8f, 1.2f, 0.4f, 30.0f, 1.5f, 5, true, true, false});
    sharpen->initialize();
    return addOperation("sharpen", [sharpen](const cv::Mat& in, cv::Mat& out) { sharpen->process(in, out); });
}

Processor& Processor::addEnhanceChain(Upscaler* upscaler) {
    return addOperation("enhance_chain", [upscaler](const cv::Mat& in, cv::Mat& out) { upscaler->upscale(in, out); });
}

// ------------------------------------------------------------------------------------------------
// VideoEnhancer facade
// ------------------------------------------------------------------------------------------------
struct VideoEnhancer::Impl {
    vp_ctx* ctx = nullptr;
    vp_chain_config cfg{};
    ~Impl() { if (ctx) vp_destroy(ctx); }
};

VideoEnhancer::VideoEnhancer(EnhancementLevel level) : m_level(level), m_initialized(false), m_impl(new Impl) {}
VideoEnhancer::~VideoEnhancer() = default;
VideoEnhancer::EnhancementLevel VideoEnhancer::getLevel() const { return m_level; }

void VideoEnhancer::setLevel(EnhancementLevel level) {
    if (level != m_level) {
        m_level = level;
        if (m_impl->ctx) { vp_destroy(m_impl->ctx); m_impl->ctx = nullptr; }
    }
}

bool VideoEnhancer::initialize() {
    if (vp_device_count() <= 0) {
        std::cerr << "VideoEnhancer: a B200 is required; there is no CPU fallback" << std::endl;
        m_initialized = false;
        return false;
    }
    m_initialized = true;
    return true;
}

bool VideoEnhancer::enhance(const cv::Mat& input, cv::Mat& output) {
    if (!m_initialized || input.empty()) return false;
    if (m_level == NONE) { input.copyTo(output); return true; }
    if (!vp_host_api::is_bgr8(input)) return false;
    vp_chain_config c{};
    vp_default_chain_config(&c, input.cols, input.rows, input.cols, input.rows);
    c.use_sharpen = 1;
    c.use_pre = m_level >= MEDIUM;
    c.use_post = m_level >= STRONG;
    c.temporal.mode = m_level >= YOUTUBE ? VP_TEMPORAL_BLEND_FRAMES : VP_TEMPORAL_NONE;
    c.temporal.scene_detect = 1;
    if (!m_impl->ctx || std::memcmp(&c, &m_impl->cfg, sizeof(c)) != 0) {
        if (m_impl->ctx) { vp_destroy(m_impl->ctx); m_impl->ctx = nullptr; }
        if (vp_create(&m_impl->ctx, 0, &c) != VP_OK) {
            std::cerr << "VideoEnhancer: " << vp_last_error(nullptr) << std::endl;
            return false;
        }
        m_impl->cfg = c;
    }
    vp::createPinned(output, input.rows, input.cols, CV_8UC3);
    if (vp_process_host(m_impl->ctx, input.data, input.step, output.data, output.step) != VP_OK) {
        std::cerr << "VideoEnhancer: " << vp_last_error(m_impl->ctx) << std::endl;
        input.copyTo(output);
        return false;
    }
    return true;
}

// ------------------------------------------------------------------------------------------------
// FrameBuffer: bounded ring, pinned slots (reference src/frame_buffer.cpp:13-110)
// ------------------------------------------------------------------------------------------------
FrameBuffer::FrameBuffer(size_t capacity) : m_capacity(capacity), m_head(0), m_tail(0), m_size(0) {
    m_frames.resize(capacity);
}

size_t FrameBuffer::nextIndex(size_t current) const { return (current + 1) % m_capacity; }
size_t FrameBuffer::size() const { return m_size.load(); }
bool FrameBuffer::empty() const { return size() == 0; }
bool FrameBuffer::full() const { return size() >= m_capacity; }
size_t FrameBuffer::capacity() const { return m_capacity; }

bool FrameBuffer::pushFrame(const cv::Mat& frame, bool blocking) {
    if (frame.empty()) {
        std::cerr << "Warning: Attempting to push empty frame to buffer" << std::endl;
        return false;
    }
    std::unique_lock<std::mutex> lock(m_mutex);
    if (full()) {
        if (!blocking) return false;
        m_not_full.wait(lock, [this]() { return !full(); });
    }
    // deep copy into the slot (src/frame_buffer.cpp:31-37): the slot is page-locked - with the Mat shim and, through the
    // pinned cv::MatAllocator, with real OpenCV - and copyTo's create() keeps that storage while the geometry matches
    vp::createPinned(m_frames[m_head], frame.rows, frame.cols, frame.type());
    frame.copyTo(m_frames[m_head]);
    m_head = nextIndex(m_head);
    m_size++;
    m_not_empty.notify_one();
    return true;
}

bool FrameBuffer::popFrame(cv::Mat& frame, bool blocking) {
    std::unique_lock<std::mutex> lock(m_mutex);
    if (empty()) {
        if (!blocking) return false;
        m_not_empty.wait(lock, [this]() { return !empty(); });
    }
    frame = m_frames[m_tail];             // refcount hand-off, no pixel copy
    m_frames[m_tail] = cv::Mat();
    m_tail = nextIndex(m_tail);
    m_size--;
    m_not_full.notify_one();
    return true;
}

void FrameBuffer::clear() {
    std::unique_lock<std::mutex> lock(m_mutex);
    for (cv::Mat& f : m_frames) f.release();
    m_head = m_tail = 0;
    m_size = 0;
    m_not_full.notify_all();
}
