// mat_alloc.cpp - page-locked storage for frame-sized cv::Mat buffers, so cudaMemcpyAsync from/to them is a true
// asynchronous DMA (the reference's FrameBuffer deep-copies every frame into its slots, src/frame_buffer.cpp:31-37; here
// that copy lands in pinned memory and is the only host copy).  Page-locking is slow (milliseconds), so released frame
// buffers go back to a size-keyed pool and FrameBuffer slots / per-frame output Mats recycle them.
//   * without OpenCV (this image): the storage of the cv::Mat subset of include/vp/mat.h - large Mats are pinned, small
//     ones use the C heap;
//   * built against real OpenCV (VP_HAVE_OPENCV): a cv::MatAllocator over the same pool.  vp::createPinned() installs it
//     on the Mats the library itself creates - FrameBuffer slots and the output frames of Upscaler / the stage classes -
//     so those stay pinned in a drop-in build too; Mats the caller allocates are whatever the caller made them (pageable
//     ones are staged through pinned slots by vp_submit_host / vp_process_host).  That branch follows OpenCV's
//     StdMatAllocator (modules/core/src/matrix.cpp) and could not be compiled in this image (no OpenCV headers).
// This allocates HOST memory only - it is not a compute fallback.
#include <cstdlib>
#include <map>
#include <mutex>

#include "vp/mat.h"
#include "vpb200.h"

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

#if !VP_HAVE_OPENCV
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
#else
namespace {

class PinnedMatAllocator : public cv::MatAllocator {
public:
    cv::UMatData* allocate(int dims, const int* sizes, int type, void* data0, size_t* step, cv::AccessFlag /*flags*/,
                           cv::UMatUsageFlags /*usage*/) const override {
        size_t total = CV_ELEM_SIZE(type);
        for (int i = dims - 1; i >= 0; --i) {
            if (step) {
                if (data0 && step[i] != CV_AUTOSTEP) { CV_Assert(total <= step[i]); total = step[i]; }
                else step[i] = total;
            }
            total *= (size_t)sizes[i];
        }
        cv::UMatData* u = new cv::UMatData(this);
        if (data0) {
            u->data = u->origdata = static_cast<uchar*>(data0);
            u->flags |= cv::UMatData::USER_ALLOCATED;
        } else {
            void* p = (total >= (256u << 10) && vp_device_count() > 0) ? pinned_get(total) : nullptr;
            if (p) u->userdata = const_cast<PinnedMatAllocator*>(this);   // marks a pool block
            else p = cv::fastMalloc(total);
            u->data = u->origdata = static_cast<uchar*>(p);
        }
        u->size = total;
        return u;
    }
    bool allocate(cv::UMatData* u, cv::AccessFlag, cv::UMatUsageFlags) const override { return u != nullptr; }
    void deallocate(cv::UMatData* u) const override {
        if (!u) return;
        CV_Assert(u->urefcount == 0 && u->refcount == 0);
        if (!(u->flags & cv::UMatData::USER_ALLOCATED)) {
            if (u->userdata == this) pinned_put(u->origdata, u->size);
            else cv::fastFree(u->origdata);
            u->origdata = nullptr;
        }
        delete u;
    }
};

}  // namespace

namespace vp {
cv::MatAllocator* pinnedMatAllocator() {
    static PinnedMatAllocator* a = new PinnedMatAllocator;     // leaked on purpose: outlives every Mat that names it
    return a;
}
}  // namespace vp
#endif
