// adaptive_sharpening.h - AdaptiveSharpening with the reference's interface
// (reference include/adaptive_sharpening.h:12-126), executed by libvpb200's fused edge-mask +
// unsharp-mask + tone-preservation kernel.
//
// adaptive_sigma = false: createEdgeMask + applyUnsharpMask (src/adaptive_sharpening.cpp:107-355), one fused kernel.
// adaptive_sigma = true: calculateTextureMap + calculateAdaptiveSigma + applyVariableSigmaUnsharpMask (:357-592) with
// REPAIRED semantics - the reference calls cv::sqrt on a CV_8U Mat (:386, undefined behaviour); here local_var is
// converted to float32 first (oracle/ref_chain.py:sharpen_variable_sigma).
#pragma once
#include <memory>
#include "vp/mat.h"

class AdaptiveSharpening {
public:
    struct Config {                       // field order is ABI
        float strength = 0.8f;
        float edge_strength = 1.2f;
        float smooth_strength = 0.4f;
        float edge_threshold = 30.0f;
        float sigma = 1.5f;
        int kernel_size = 5;
        bool preserve_tone = true;
        bool use_gpu = true;
        bool adaptive_sigma = true;
    };

    AdaptiveSharpening();
    explicit AdaptiveSharpening(const Config& config);
    ~AdaptiveSharpening();

    bool initialize();
    bool process(const cv::Mat& input, cv::Mat& output);
    void setConfig(const Config& config);
    Config getConfig() const;

private:
    struct Impl;
    Config m_config;
    bool m_initialized = false;
    std::unique_ptr<Impl> m_impl;
};
