// selective_bilateral.h - SelectiveBilateral with the reference's interface
// (reference include/selective_bilateral.h:12-153), executed by libvpb200's CUDA bilateral kernel.
//
// process() dispatches as the reference does (src/selective_bilateral.cpp:57-82): use_multiscale ->
// applyMultiscaleBilateral (:138-237), else selective -> applySelectiveBilateral (:121-136, :373-439), else
// applyBilateralFilter (:84-119) with calculateAdaptiveParams (:315-371).  The first two run with REPAIRED
// semantics: createDetailMask calls cv::sqrt on a CV_8U Mat in the reference (:276, undefined behaviour); here
// local_var is converted to float32 first (oracle/ref_chain.py:create_detail_mask).
#pragma once
#include <memory>
#include "vp/mat.h"

class SelectiveBilateral {
public:
    enum FilteringStage { PRE_PROCESSING, POST_PROCESSING };

    struct Config {                       // field order is ABI: callers brace-initialise positionally
        FilteringStage stage = PRE_PROCESSING;
        bool use_gpu = true;
        bool adaptive_params = true;
        int diameter = 7;
        double sigma_color = 30.0;
        double sigma_space = 30.0;
        bool selective = true;
        double detail_threshold = 30.0;
        double texture_boost = 1.5;
        double edge_preserve = 2.0;
        bool use_multiscale = true;
        int num_scales = 3;
    };

    SelectiveBilateral();
    explicit SelectiveBilateral(const Config& config);
    ~SelectiveBilateral();

    bool initialize();
    bool process(const cv::Mat& input, cv::Mat& output);
    void setConfig(const Config& config);
    Config getConfig() const;

private:
    struct Impl;
    Config m_config;
    bool m_initialized = false;
    std::unique_ptr<Impl> m_impl;         // vp_ctx + device staging, created by initialize()
};
