// processor.h - Processor, the reference's ordered op-chain plugin API (reference
// include/processor.h:17-122, src/processor.cpp:173-300).  Host-side control only: operations are
// std::function plugins run in order under a mutex; the B200 chain registers itself as one fused
// operation through addEnhanceChain().  The reference's "GPU" implementation uploads and downloads
// around every op (src/processor.cpp:98-124) without computing on the device; here device work
// happens inside the ops that own a vp_ctx.
#pragma once
#include <functional>
#include <memory>
#include <mutex>
#include <string>
#include <vector>
#include "vp/mat.h"

class Upscaler;

class Processor {
public:
    using ProcessFunction = std::function<void(const cv::Mat&, cv::Mat&)>;

    struct Operation {
        std::string name;
        ProcessFunction func;
        bool enabled;
        Operation(const std::string& n, ProcessFunction f) : name(n), func(f), enabled(true) {}
    };

    class ProcessorImpl;

    explicit Processor(bool use_gpu = true);
    ~Processor();

    bool initialize();
    bool process(const cv::Mat& input, cv::Mat& output);
    Processor& addOperation(const std::string& name, ProcessFunction func);
    Processor& addDefaultPreProcessing();    // "denoise": SelectiveBilateral (PRE)
    Processor& addDefaultPostProcessing();   // "sharpen": AdaptiveSharpening
    void enableOperation(const std::string& name, bool enabled);
    bool setUseGPU(bool use_gpu);
    bool isUsingGPU() const;
    static bool isGPUAvailable();
    double getLastProcessingTime() const;

    // addition: registers `upscaler->upscale` as the operation "enhance_chain" (upscaler must outlive this)
    Processor& addEnhanceChain(Upscaler* upscaler);

private:
    std::vector<Operation> m_operations;
    bool m_use_gpu;
    bool m_initialized;
    mutable std::mutex m_mutex;
    double m_last_processing_time;
    std::unique_ptr<ProcessorImpl> m_impl;
};
