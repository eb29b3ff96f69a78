// Minimal stand-in carrying OpenCV 4.x declarations (cv::MatAllocator, cv::UMatData, cv::Mat::allocator) - used ONLY by
// tests/test_host_logic.py to syntax-check the VP_HAVE_OPENCV branch of csrc/host/mat_alloc.cpp; this image has no OpenCV.
#pragma once
#include <cstddef>
#include <cstdlib>
#include <cassert>
typedef unsigned char uchar;
#define CV_Assert(x) assert(x)
#define CV_ELEM_SIZE(t) ((size_t)((((t) >> 3) + 1) * (((t)&7) == 5 ? 4 : 1)))
#define CV_AUTOSTEP 0x7fffffff
#define CV_8UC3 16
namespace cv {
enum AccessFlag { ACCESS_READ = 1 };
enum UMatUsageFlags { USAGE_DEFAULT = 0 };
struct UMatData;
class MatAllocator { public: MatAllocator() {} virtual ~MatAllocator() {}
  virtual UMatData* allocate(int dims, const int* sizes, int type, void* data, size_t* step, AccessFlag flags, UMatUsageFlags usageFlags) const = 0;
  virtual bool allocate(UMatData* data, AccessFlag accessflags, UMatUsageFlags usageFlags) const = 0;
  virtual void deallocate(UMatData* data) const = 0; };
struct UMatData { enum MemoryFlag { USER_ALLOCATED = 32 }; UMatData(const MatAllocator* a) : prevAllocator(a), currAllocator(a) {}
  const MatAllocator* prevAllocator; const MatAllocator* currAllocator; int urefcount = 0, refcount = 0; uchar* data = nullptr; uchar* origdata = nullptr; size_t size = 0; int flags = 0; void* handle = nullptr; void* userdata = nullptr; };
inline int operator|(int a, UMatData::MemoryFlag b) { return a | (int)b; }
inline void* fastMalloc(size_t n) { return std::malloc(n); }
inline void fastFree(void* p) { std::free(p); }
class Mat { public: int rows = 0, cols = 0; uchar* data = nullptr; MatAllocator* allocator = nullptr; struct { size_t v; operator size_t() const { return v; } } step;
  bool empty() const { return !data; } int type() const { return 16; } void release() {} void create(int, int, int) {} void copyTo(Mat&) const {} };
}
