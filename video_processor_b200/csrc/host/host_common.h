// host_common.h - shared plumbing of the C++ API classes (include/*.h) over the C ABI (include/vpb200.h).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <iostream>
#include <string>

#include "vp/mat.h"
#include "vpb200.h"

namespace vp_host_api {

// One vp_ctx used through the stage-wise entry points, plus device staging for a same-size or
// resizing stage: host Mat -> device -> stage kernel(s) -> device -> host Mat.
class StageRunner {
public:
    ~StageRunner() { destroy(); }
    bool ready() const { return ctx_ != nullptr; }
    bool create(int device, std::string* err) {
        destroy();
        const int rc = vp_create(&ctx_, device, nullptr);
        if (rc != VP_OK) { if (err) *err = vp_last_error(nullptr); ctx_ = nullptr; return false; }
        device_ = device;
        cudaSetDevice(device_);
        // copies and stage kernels share this stream, so an upload from pageable memory (whose DMA may
        // still be in flight when cudaMemcpy returns) is ordered before the kernel that reads it
        if (cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking) != cudaSuccess) { destroy(); return false; }
        return true;
    }
    void destroy() {
        if (ctx_) {
            cudaSetDevice(device_);
            if (stream_) { cudaStreamSynchronize(stream_); cudaStreamDestroy(stream_); }
            cudaFree(d_in_); cudaFree(d_out_); vp_destroy(ctx_);
        }
        stream_ = nullptr;
        ctx_ = nullptr; d_in_ = d_out_ = nullptr; in_bytes_ = out_bytes_ = 0;
    }
    vp_ctx* ctx() { return ctx_; }
    void* stream() { return stream_; }
    const char* error() const { return vp_last_error(ctx_); }

    // uploads `input` (CV_8UC3) and makes sure an output buffer of out_w x out_h x out_elem exists
    bool upload(const cv::Mat& input, int out_w, int out_h, size_t out_elem) {
        cudaSetDevice(device_);
        in_pitch_ = align256((size_t)input.cols * 3);
        out_pitch_ = align256((size_t)out_w * out_elem);
        if (!reserve(&d_in_, &in_bytes_, in_pitch_ * input.rows)) return false;
        if (!reserve(&d_out_, &out_bytes_, out_pitch_ * out_h)) return false;
        return cudaMemcpy2DAsync(d_in_, in_pitch_, input.data, input.step, (size_t)input.cols * 3, input.rows,
                                 cudaMemcpyHostToDevice, stream_) == cudaSuccess;
    }
    bool download(cv::Mat& output, int w, int h, int type, size_t elem) {
        cudaSetDevice(device_);
        vp::createPinned(output, h, w, type);
        if (cudaMemcpy2DAsync(output.data, output.step, d_out_, out_pitch_, (size_t)w * elem, h,
                              cudaMemcpyDeviceToHost, stream_) != cudaSuccess) return false;
        return cudaStreamSynchronize(stream_) == cudaSuccess;
    }
    uint8_t* d_in() { return d_in_; }
    uint8_t* d_out() { return d_out_; }
    size_t in_pitch() const { return in_pitch_; }
    size_t out_pitch() const { return out_pitch_; }

private:
    static size_t align256(size_t v) { return (v + 255) / 256 * 256; }
    bool reserve(uint8_t** p, size_t* have, size_t need) {
        if (*have >= need) return true;
        cudaFree(*p); *p = nullptr; *have = 0;
        if (cudaMalloc(p, need) != cudaSuccess) { cudaGetLastError(); return false; }
        *have = need;
        return true;
    }
    vp_ctx* ctx_ = nullptr;
    int device_ = 0;
    cudaStream_t stream_ = nullptr;
    uint8_t *d_in_ = nullptr, *d_out_ = nullptr;
    size_t in_bytes_ = 0, out_bytes_ = 0, in_pitch_ = 0, out_pitch_ = 0;
};

inline bool is_bgr8(const cv::Mat& m) { return m.type() == CV_8UC3; }

}  // namespace vp_host_api
