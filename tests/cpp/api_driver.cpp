// api_driver.cpp - exercises the C++ API (include/*.h) the way the reference's callers do:
//   pipeline.cpp:351-383 (FrameBuffer::popFrame -> Upscaler::upscale), test_enhancements.cpp:333-377
//   (stage objects composed by hand), Processor plugin chain, VideoEnhancer facade.
// usage: api_driver <mode> <in.bin> <w> <h> <nframes> <tw> <th> <out.bin>
// Reads nframes raw BGR8 frames, writes the produced frames; tests/test_gpu_cpp_api.py compares them
// with the oracle.  Exit code 0 = every API call returned what the mode expects.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "adaptive_sharpening.h"
#include "frame_buffer.h"
#include "processor.h"
#include "selective_bilateral.h"
#include "temporal_consistency.h"
#include "upscaler.h"
#include "video_enhancer.h"
#include "vp/imgproc.h"

static int fail(const char* what) { std::fprintf(stderr, "api_driver: FAILED: %s\n", what); return 1; }

int main(int argc, char** argv) {
    if (argc < 2) return fail("usage");
    const std::string mode = argv[1];

    if (mode == "nogpu") {   // CPU-only box: everything must refuse cleanly, nothing may crash or fall back
        Upscaler up(Upscaler::REAL_ESRGAN, true);
        if (up.initialize(64, 64)) return fail("Upscaler::initialize succeeded without a GPU");
        if (up.isUsingGPU()) return fail("isUsingGPU() is not truthful");
        cv::Mat in(8, 8, CV_8UC3), out;
        std::memset(in.data, 7, in.step * 8);
        if (up.upscale(in, out)) return fail("upscale succeeded uninitialised");
        SelectiveBilateral sb; if (sb.initialize()) return fail("SelectiveBilateral::initialize succeeded without a GPU");
        AdaptiveSharpening as; if (as.initialize()) return fail("AdaptiveSharpening::initialize succeeded without a GPU");
        TemporalConsistency tc; if (tc.initialize()) return fail("TemporalConsistency::initialize succeeded without a GPU");
        VideoEnhancer ve; if (ve.initialize()) return fail("VideoEnhancer::initialize succeeded without a GPU");
        Upscaler bl(Upscaler::BILINEAR, true); if (bl.initialize(64, 64)) return fail("BILINEAR is not a built path");
        if (Upscaler(Upscaler::BICUBIC).getAlgorithmName() != "Bicubic") return fail("algorithm name");
        FrameBuffer fb(2);
        if (!fb.pushFrame(in, false) || !fb.pushFrame(in, false) || fb.pushFrame(in, false)) return fail("FrameBuffer capacity");
        cv::Mat got;
        if (!fb.popFrame(got, false) || got.rows != 8 || got.data == in.data || std::memcmp(got.data, in.data, 8 * 24)) return fail("FrameBuffer pop");
        fb.clear();
        if (!fb.empty() || fb.popFrame(got, false)) return fail("FrameBuffer clear");
        std::puts("nogpu ok");
        return 0;
    }

    if (argc < 9) return fail("usage");
    const int w = std::atoi(argv[3]), h = std::atoi(argv[4]), n = std::atoi(argv[5]);
    const int tw = std::atoi(argv[6]), th = std::atoi(argv[7]);
    std::vector<cv::Mat> frames;
    {
        FILE* f = std::fopen(argv[2], "rb");
        if (!f) return fail("open input");
        for (int i = 0; i < n; ++i) {
            cv::Mat m(h, w, CV_8UC3);
            if (std::fread(m.data, 1, (size_t)w * h * 3, f) != (size_t)w * h * 3) return fail("short input");
            frames.push_back(m);
        }
        std::fclose(f);
    }
    std::vector<cv::Mat> outs;

    if (mode == "chain" || mode == "bicubic") {
        // capture thread -> FrameBuffer -> processing loop, as pipeline.cpp does
        Upscaler up(mode == "chain" ? Upscaler::REAL_ESRGAN : Upscaler::BICUBIC, true);
        if (!up.initialize(tw, th)) return fail("Upscaler::initialize");
        if (!up.isUsingGPU() || up.getTargetWidth() != tw) return fail("Upscaler state");
        FrameBuffer raw(3);
        std::thread capture([&]() { for (auto& f : frames) raw.pushFrame(f, true); });
        for (int i = 0; i < n; ++i) {
            cv::Mat in, out;
            if (!raw.popFrame(in, true)) return fail("popFrame");
            if (!up.upscale(in, out)) return fail("upscale");
            if (out.cols != tw || out.rows != th || out.type() != CV_8UC3) return fail("output geometry");
            outs.push_back(out.clone());
        }
        capture.join();
        cv::Mat empty, out;
        if (up.upscale(empty, out)) return fail("empty input accepted");
    } else if (mode == "resize") {
        for (int i = 0; i < n; ++i) {
            cv::Mat out;
            if (!vp::resizeCubic(frames[i], out, cv::Size(tw, th))) return fail("resizeCubic");
            outs.push_back(out.clone());
        }
    } else if (mode == "pipelined") {
        Upscaler up(Upscaler::REAL_ESRGAN, true);
        if (!up.initialize(tw, th)) return fail("Upscaler::initialize");
        outs.resize(n);
        int inflight = 0;
        for (int i = 0; i < n; ++i) {
            if (inflight == up.pipelineDepth()) { if (!up.wait()) return fail("wait"); --inflight; }
            if (!up.submit(frames[i], outs[i])) return fail("submit");
            ++inflight;
        }
        while (inflight--) if (!up.wait()) return fail("wait (drain)");
    } else if (mode == "segments") {
        // file/--fast mode over several devices: VP_SEGMENT_DEVICES = "0,0,0" or "0,1" (default: every visible device)
        Upscaler up(Upscaler::REAL_ESRGAN, true);
        if (!up.initialize(tw, th)) return fail("Upscaler::initialize");
        std::vector<int> devs;
        if (const char* e = std::getenv("VP_SEGMENT_DEVICES"))
            for (const char* p = e; *p; ++p) if (*p >= '0' && *p <= '9') devs.push_back(*p - '0');
        if (!up.upscaleSegments(frames, outs, devs)) return fail("upscaleSegments");
        if ((int)outs.size() != n) return fail("upscaleSegments output count");
    } else if (mode == "toggles") {
        // keys 1-4 of test_enhancements.cpp: stages switched off one by one between frames
        Upscaler up(Upscaler::REAL_ESRGAN, true);
        if (!up.initialize(tw, th)) return fail("Upscaler::initialize");
        up.setUseTemporalConsistency(false);
        for (int i = 0; i < n; ++i) {
            if (i == 1) up.setUseSelectiveBilateral(false);
            if (i == 2) up.setUseAdaptiveSharpening(false);
            cv::Mat out;
            if (!up.upscale(frames[i], out)) return fail("upscale");
            outs.push_back(out.clone());
        }
    } else if (mode == "stages") {
        // test_enhancements.cpp:333-377 composition with in-scope configs
        SelectiveBilateral pre(SelectiveBilateral::Config{SelectiveBilateral::PRE_PROCESSING, true, true, 7, 30.0, 30.0, false, 30.0, 1.5, 2.0, false, 3});
        SelectiveBilateral post(SelectiveBilateral::Config{SelectiveBilateral::POST_PROCESSING, true, true, 5, 20.0, 20.0, false, 30.0, 1.2, 1.5, false, 2});
        AdaptiveSharpening sharp(AdaptiveSharpening::Config{0.8f, 1.2f, 0.4f, 30.0f, 1.5f, 5, true, true, false});
        TemporalConsistency temporal(TemporalConsistency::Config{3, 0.6f, 15.0f, 100.0f, true, 0.5, 3, 15, 3, 5, 1.2, 0});
        if (!pre.initialize() || !post.initialize() || !sharp.initialize() || !temporal.initialize())
            return fail("stage initialize");
        for (int i = 0; i < n; ++i) {
            cv::Mat a, b, c, d, e;
            if (!pre.process(frames[i], a) || !vp::resizeCubic(a, b, cv::Size(tw, th)) || !sharp.process(b, c) || !post.process(c, d) || !temporal.process(d, e))
                return fail("stage process");
            outs.push_back(e.clone());
        }
        temporal.reset();
    } else if (mode == "processor") {
        Upscaler up(Upscaler::REAL_ESRGAN, true);
        if (!up.initialize(tw, th)) return fail("Upscaler::initialize");
        Processor proc(true);
        int calls = 0;
        proc.addEnhanceChain(&up).addOperation("count", [&calls](const cv::Mat& in, cv::Mat& out) { ++calls; in.copyTo(out); })
            .addOperation("disabled", [](const cv::Mat&, cv::Mat& out) { out.release(); });
        proc.enableOperation("disabled", false);
        if (!proc.initialize()) return fail("Processor::initialize");
        for (int i = 0; i < n; ++i) {
            cv::Mat out;
            if (!proc.process(frames[i], out)) return fail("Processor::process");
            outs.push_back(out.clone());
        }
        if (calls != n || proc.getLastProcessingTime() <= 0.0) return fail("Processor bookkeeping");
    } else if (mode == "enhancer") {
        VideoEnhancer ve(VideoEnhancer::YOUTUBE);
        if (!ve.initialize()) return fail("VideoEnhancer::initialize");
        for (int i = 0; i < n; ++i) {
            cv::Mat out;
            if (!ve.enhance(frames[i], out)) return fail("enhance");
            outs.push_back(out.clone());
        }
    } else {
        return fail("unknown mode");
    }

    FILE* f = std::fopen(argv[8], "wb");
    if (!f) return fail("open output");
    for (auto& m : outs)
        for (int y = 0; y < m.rows; ++y) std::fwrite(m.ptr(y), 1, (size_t)m.cols * 3, f);
    std::fclose(f);
    std::printf("%s ok: %zu frames\n", mode.c_str(), outs.size());
    return 0;
}
