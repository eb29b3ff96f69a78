// video_enhancer.h - VideoEnhancer facade (reference include/video_enhancer.h:7-24).
//
// The reference's src/video_enhancer.cpp is in no build target, has no callers and does not compile
// (:434 uses an undeclared `lut`); its colour-grading body is out of scope.  The class is kept as the
// entry point BASELINE.json names: enhance() runs the same-size enhancement stages on the B200 chain,
//   NONE: copy   LIGHT: AdaptiveSharpening   MEDIUM: + SelectiveBilateral(PRE)
//   STRONG: + SelectiveBilateral(POST)       YOUTUBE: + TemporalConsistency blend
#pragma once
#include <memory>
#include <string>
#include "vp/mat.h"

class VideoEnhancer {
public:
    enum EnhancementLevel { NONE, LIGHT, MEDIUM, STRONG, YOUTUBE };

    VideoEnhancer(EnhancementLevel level = MEDIUM);
    ~VideoEnhancer();

    bool initialize();
    bool enhance(const cv::Mat& input, cv::Mat& output);
    void setLevel(EnhancementLevel level);
    EnhancementLevel getLevel() const;

private:
    struct Impl;
    EnhancementLevel m_level;
    bool m_initialized;
    std::unique_ptr<Impl> m_impl;
};
