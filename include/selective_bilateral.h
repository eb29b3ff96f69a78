// selective_bilateral.h - SelectiveBilateral with the reference's interface
// (reference include/selective_bilateral.h:12-153), executed by libvpb200's CUDA bilateral kernel.
//
// Built branch: SelectiveBilateral::applyBilateralFilter (src/selective_bilateral.cpp:84-119) with
// calculateAdaptiveParams (:315-371).  use_multiscale / selective are accepted but take the same
// branch: in the reference their bodies call cv::sqrt on a CV_8U Mat (:276, undefined behaviour) and
// its own GPU multiscale path ends in `return applyBilateralFilter(input, output)` through the
// exception fallback (:164-176, :232-235).  DESIGN.md lists this under "reference defects".
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
