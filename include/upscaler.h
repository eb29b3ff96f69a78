// upscaler.h - Upscaler with the reference's interface (reference include/upscaler.h:26-161), so it
// drops into src/pipeline.cpp:97-104,382 and src/main.cpp:792-800,259 unchanged.
//
// What runs underneath is libvpb200's fused CUDA chain (include/vpb200.h):
//   REAL_ESRGAN : SelectiveBilateral(PRE) -> cv::resize(INTER_CUBIC) -> AdaptiveSharpening ->
//                 SelectiveBilateral(POST) -> TemporalConsistency blend   (src/upscaler.cpp:772-829,
//                 the resize standing in for the ONNX model exactly as src/test_enhancements.cpp:346-349)
//   BICUBIC     : CPUImpl's wrapper - GaussianBlur 3x3 -> cv::resize(INTER_CUBIC) -> enhanceBicubicResult
//                 (src/upscaler.cpp:75-148); a bare bit-exact resize is vp::resizeCubic (vp/imgproc.h)
// NEAREST / BILINEAR / LANCZOS / SUPER_RES are not on the hot path this library accelerates:
// initialize() reports that and returns false.  There is no CPU fallback: use_gpu=true without a
// B200 makes initialize() fail, and isUsingGPU() stays truthful.
#pragma once
#include <memory>
#include <string>
#include <vector>
#include "vp/mat.h"

struct vp_chain_config;

class SelectiveBilateral;
class AdaptiveSharpening;
class TemporalConsistency;

class Upscaler {
public:
    enum Algorithm { NEAREST, BILINEAR, BICUBIC, LANCZOS, SUPER_RES, REAL_ESRGAN };

    explicit Upscaler(Algorithm algorithm = BILINEAR, bool use_gpu = true);
    ~Upscaler();

    bool initialize(int target_width, int target_height);
    bool upscale(const cv::Mat& input, cv::Mat& output);
    void setAlgorithm(Algorithm algorithm);
    bool setUseGPU(bool use_gpu);
    bool isUsingGPU() const;
    std::string getAlgorithmName() const;
    static bool isGPUAvailable();
    bool adjustQualityForPerformance(double processing_time, double target_time);
    int getTargetWidth() const { return m_target_width; }
    int getTargetHeight() const { return m_target_height; }

    void setUseSelectiveBilateral(bool enable) { m_use_selective_bilateral = enable; }
    void setUseAdaptiveSharpening(bool enable) { m_use_adaptive_sharpening = enable; }
    void setUseTemporalConsistency(bool enable) { m_use_temporal_consistency = enable; }
    bool isUsingSelectiveBilateral() const { return m_use_selective_bilateral; }
    bool isUsingAdaptiveSharpening() const { return m_use_adaptive_sharpening; }
    bool isUsingTemporalConsistency() const { return m_use_temporal_consistency; }

    SelectiveBilateral* getBilateralPreProcessor() { return m_bilateral_pre.get(); }
    AdaptiveSharpening* getAdaptiveSharpening() { return m_sharpening.get(); }
    SelectiveBilateral* getBilateralPostProcessor() { return m_bilateral_post.get(); }
    TemporalConsistency* getTemporalConsistency() { return m_temporal_consistency.get(); }

    // --- additions for file/--fast mode (not in the reference) ---------------------------------
    // Pipelined form of upscale(): up to pipelineDepth() frames in flight, H2D / kernels / D2H on
    // an H2D stream, the compute lanes and a D2H stream.  `input` AND `output` must stay alive and untouched until the
    // matching wait() returns (a pinned input Mat is read by DMA after submit() returns - do not reuse the capture Mat).
    int pipelineDepth() const;
    bool submit(const cv::Mat& input, cv::Mat& output);
    bool wait();
    // Runs a frame through the chain for its temporal side effects only (multi-GPU segment warm-up).
    bool warmup(const cv::Mat& input);
    void setDevice(int device) { m_device = device; }
    // file/--fast mode over several B200s (the loop of src/main.cpp:222-355 with the whole clip known ahead): the frames are
    // cut into contiguous segments, one per device, each preceded by its temporal warm-up frames; outputs come back in clip
    // order and are bit-identical to calling upscale() frame by frame on one device.  `devices` empty = every visible B200.
    // A thin wrapper over vp_run_segments (include/vpb200.h); this object's own temporal state is not touched.
    bool upscaleSegments(const std::vector<cv::Mat>& inputs, std::vector<cv::Mat>& outputs, const std::vector<int>& devices = {});

private:
    struct Chain;
    Algorithm m_algorithm;
    bool m_use_gpu;
    bool m_initialized;
    int m_target_width;
    int m_target_height;
    int m_device = 0;
    std::unique_ptr<Chain> m_chain;       // vp_ctx of the fused chain, (re)built lazily per input size / toggles

    bool initializeImpl();
    bool initializeEnhancements();
    bool ensureChain(int src_w, int src_h);
    bool buildChainConfig(int src_w, int src_h, vp_chain_config* out) const;

    std::unique_ptr<SelectiveBilateral> m_bilateral_pre;
    std::unique_ptr<AdaptiveSharpening> m_sharpening;
    std::unique_ptr<SelectiveBilateral> m_bilateral_post;
    std::unique_ptr<TemporalConsistency> m_temporal_consistency;

    bool m_use_selective_bilateral;
    bool m_use_adaptive_sharpening;
    bool m_use_temporal_consistency;
};
