// temporal_consistency.h - TemporalConsistency with the reference's interface
// (reference include/temporal_consistency.h:16-141).
//
// Built: process() as the reference writes it (src/temporal_consistency.cpp:61-172) - the frame-buffer policy, the
// scene-change reset (detectSceneChange :283-323), cv::calcOpticalFlowFarneback with the Config's parameters (:174-224),
// the cv::remap warp of every buffered frame with its own stored flow (:127-135), the reliability masks (:325-353) and
// blendFrames (:355-444, with the float32 accumulator the code evidently intends; the literal code writes Vec3f into a
// CV_8UC3 Mat, :367 vs :408).  All of it runs on the device.  Limits: buffer_size <= 3, Farneback
// `flags` must be 0 (the reference passes 0; OPTFLOW_USE_INITIAL_FLOW / OPTFLOW_FARNEBACK_GAUSSIAN are not built).
#pragma once
#include <memory>
#include <mutex>
#include "vp/mat.h"

class TemporalConsistency {
public:
    struct Config {                       // field order is ABI
        int buffer_size = 3;
        float blend_strength = 0.6f;
        float motion_threshold = 15.0f;
        float scene_change_threshold = 100.0f;
        bool use_gpu = true;
        double pyr_scale = 0.5;
        int levels = 3;
        int winsize = 15;
        int iterations = 3;
        int poly_n = 5;
        double poly_sigma = 1.2;
        int flags = 0;
    };

    TemporalConsistency();
    explicit TemporalConsistency(const Config& config);
    ~TemporalConsistency();

    bool initialize();
    bool process(const cv::Mat& input, cv::Mat& output);
    void reset();
    void setConfig(const Config& config);
    Config getConfig() const;

private:
    struct Impl;
    Config m_config;
    bool m_initialized = false;
    std::mutex m_buffer_mutex;            // the reference guards its deques the same way (:77)
    std::unique_ptr<Impl> m_impl;
};
