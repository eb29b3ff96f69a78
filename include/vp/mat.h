// vp/mat.h - the cv::Mat the API headers use.
//
// With OpenCV installed (<opencv2/core.hpp> on the include path) this header is just that include and
// the classes in include/*.h take real cv::Mat objects, exactly as the reference does
// (include/upscaler.h:3 `#include <opencv2/opencv.hpp>`).  This image ships no OpenCV C++ headers or
// libraries, so otherwise a minimal, source-compatible subset is provided: enough for the call sites
// of pipeline.cpp / main.cpp / test_enhancements.cpp on the enhance-chain path (rows, cols, type,
// step, data, create, clone, copyTo, release, empty, size, refcounted copies, external-data ctor).
// Large buffers are page-locked (vp_host_alloc) so frames move by asynchronous DMA without a
// staging copy - the role cv::cuda::HostMem plays for real OpenCV.
#pragma once
#if __has_include(<opencv2/core.hpp>) && !defined(VP_FORCE_MAT_SHIM)
#include <opencv2/core.hpp>
#define VP_HAVE_OPENCV 1
namespace vp {
// page-locked storage for the Mats the library creates (csrc/host/mat_alloc.cpp): what keeps FrameBuffer slots and
// output frames on the asynchronous-DMA path when the build uses real OpenCV (reference: src/frame_buffer.cpp:31-37)
cv::MatAllocator* pinnedMatAllocator();
inline void createPinned(cv::Mat& m, int rows, int cols, int type) {
    if (!m.empty() && m.rows == rows && m.cols == cols && m.type() == type && m.allocator == pinnedMatAllocator()) return;
    m.release();
    m.allocator = pinnedMatAllocator();
    m.create(rows, cols, type);
}
}  // namespace vp
#else
#define VP_HAVE_OPENCV 0
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <memory>

#ifndef CV_8U
#define CV_CN_SHIFT 3
#define CV_8U 0
#define CV_32F 5
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn) - 1) << CV_CN_SHIFT))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#endif

namespace cv {

struct Size {
    int width = 0, height = 0;
    Size() = default;
    Size(int w, int h) : width(w), height(h) {}
    bool operator==(const Size& o) const { return width == o.width && height == o.height; }
    bool operator!=(const Size& o) const { return !(*this == o); }
};

namespace detail {
// implemented in libvpb200 (csrc/host/mat_alloc.cpp): pinned for frame-sized buffers, malloc otherwise
std::shared_ptr<uint8_t> allocate(size_t bytes);
}

class Mat {
public:
    int rows = 0, cols = 0;
    uint8_t* data = nullptr;
    size_t step = 0;                       // bytes per row

    Mat() = default;
    Mat(int r, int c, int type) { create(r, c, type); }
    Mat(Size s, int type) { create(s.height, s.width, type); }
    // external data, not owned (cv::Mat(rows, cols, type, void* data, size_t step))
    Mat(int r, int c, int type, void* ext, size_t ext_step = 0)
        : rows(r), cols(c), data(static_cast<uint8_t*>(ext)), type_(type) {
        step = ext_step ? ext_step : (size_t)c * elemSize();
    }

    void create(int r, int c, int type) {
        if (r == rows && c == cols && type == type_ && data && owner_) return;
        release();
        if (r <= 0 || c <= 0) return;
        rows = r; cols = c; type_ = type;
        step = (size_t)c * elemSize();
        owner_ = detail::allocate(step * (size_t)r);
        data = owner_.get();
    }
    void create(Size s, int type) { create(s.height, s.width, type); }
    void release() { owner_.reset(); data = nullptr; rows = cols = 0; step = 0; }

    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    int type() const { return type_; }
    int depth() const { return type_ & ((1 << CV_CN_SHIFT) - 1); }
    int channels() const { return (type_ >> CV_CN_SHIFT) + 1; }
    size_t elemSize() const { return (size_t)channels() * (depth() == CV_32F ? 4 : 1); }
    size_t total() const { return (size_t)rows * cols; }
    Size size() const { return Size(cols, rows); }
    bool isContinuous() const { return step == (size_t)cols * elemSize(); }
    uint8_t* ptr(int y = 0) { return data + (size_t)y * step; }
    const uint8_t* ptr(int y = 0) const { return data + (size_t)y * step; }

    void copyTo(Mat& dst) const {
        if (empty()) { dst.release(); return; }
        dst.create(rows, cols, type_);
        const size_t rb = (size_t)cols * elemSize();
        for (int y = 0; y < rows; ++y) std::memcpy(dst.data + (size_t)y * dst.step, data + (size_t)y * step, rb);
    }
    Mat clone() const { Mat m; copyTo(m); return m; }

private:
    int type_ = CV_8UC3;
    std::shared_ptr<uint8_t> owner_;       // refcounted storage: copies share pixels like cv::Mat
};

}  // namespace cv
namespace vp {
// the shim pins every frame-sized Mat (detail::allocate), so this is plain create()
inline void createPinned(cv::Mat& m, int rows, int cols, int type) { m.create(rows, cols, type); }
}  // namespace vp
#endif
