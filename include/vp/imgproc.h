// vp/imgproc.h - the one cv:: free function the enhance-chain callers use besides the classes:
// cv::resize(src, dst, Size, 0, 0, cv::INTER_CUBIC) (reference src/test_enhancements.cpp:346-349,
// src/main.cpp:265).  With OpenCV present callers keep using cv::resize; in OpenCV-less builds this is the
// stand-in, executed by libvpb200's bit-exact bicubic kernel.
#pragma once
#include "vp/mat.h"

namespace vp {
bool resizeCubic(const cv::Mat& src, cv::Mat& dst, cv::Size dsize);
}
