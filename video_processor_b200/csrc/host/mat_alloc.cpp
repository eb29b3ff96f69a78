// mat_alloc.cpp - storage for the cv::Mat subset of include/vp/mat.h: frame-sized buffers are
// page-locked so cudaMemcpyAsync from/to them is a true asynchronous DMA (what cv::cuda::HostMem
// does for real OpenCV); small ones use the C heap.  Page-locking is slow (milliseconds), so released
// frame buffers go back to a size-keyed pool and FrameBuffer slots / per-frame output Mats recycle
// them.  This allocates HOST memory only - it is not a compute fallback.
#include <cstdlib>
#include <map>
#include <mutex>

#include "vp/mat.h"
#include "vpb200.h"

#if !VP_HAVE_OPENCV
namespace {

struct PinnedPool {
    std::mutex m;
    std::multimap<size_t, void*> free_blocks;
    size_t cached = 0;
    static constexpr size_t kCap = size_t(2) << 30;     // keep at most 2 GiB of idle pinned memory
};
PinnedPool& pool() { static PinnedPool* p = new PinnedPool; return *p; }   // leaked on purpose: outlives every Mat

void* pinned_get(size_t bytes) {
    PinnedPool& P = pool();
    {
        std::lock_guard<std::mutex> lk(P.m);
        auto it = P.free_blocks.find(bytes);
        if (it != P.free_blocks.end()) {
            void* p = it->second;
            P.free_blocks.erase(it);
            P.cached -= bytes;
            return p;
        }
    }
    void* p = nullptr;
    return vp_host_alloc(&p, bytes) == VP_OK ? p : nullptr;
}

void pinned_put(void* p, size_t bytes) {
    PinnedPool& P = pool();
    {
        std::lock_guard<std::mutex> lk(P.m);
        if (P.cached + bytes <= PinnedPool::kCap) {
            P.free_blocks.emplace(bytes, p);
            P.cached += bytes;
            return;
        }
    }
    vp_host_free(p);
}

}  // namespace

namespace cv {
namespace detail {

std::shared_ptr<uint8_t> allocate(size_t bytes) {
    if (bytes >= (256u << 10) && vp_device_count() > 0) {
        if (void* p = pinned_get(bytes))
            return std::shared_ptr<uint8_t>(static_cast<uint8_t*>(p), [bytes](uint8_t* q) { pinned_put(q, bytes); });
    }
    uint8_t* p = static_cast<uint8_t*>(std::malloc(bytes ? bytes : 1));
    return std::shared_ptr<uint8_t>(p, [](uint8_t* q) { std::free(q); });
}

}  // namespace detail
}  // namespace cv
#endif
