// frame_buffer.h - FrameBuffer with the reference's interface (reference include/frame_buffer.h:17-97):
// a bounded MPMC ring of frames between the capture and processing threads.
//
// Difference underneath (BASELINE.json north star): slot storage is page-locked host memory, so the
// deep copy pushFrame makes (reference src/frame_buffer.cpp:31-37) lands in memory that
// Upscaler::upscale / submit can DMA from with cudaMemcpyAsync on a side stream, without another
// staging copy; popFrame hands the slot's refcounted Mat out (zero copy, :62-65) and the slot gets a
// recycled pinned buffer back instead of a fresh allocation.
#pragma once
#include <atomic>
#include <condition_variable>
#include <cstddef>
#include <memory>
#include <mutex>
#include <vector>
#include "vp/mat.h"

class FrameBuffer {
public:
    explicit FrameBuffer(size_t capacity = 10);
    FrameBuffer(const FrameBuffer&) = delete;
    FrameBuffer& operator=(const FrameBuffer&) = delete;
    FrameBuffer(FrameBuffer&&) = delete;
    FrameBuffer& operator=(FrameBuffer&&) = delete;

    bool pushFrame(const cv::Mat& frame, bool blocking = true);
    bool popFrame(cv::Mat& frame, bool blocking = true);
    size_t size() const;
    bool empty() const;
    bool full() const;
    size_t capacity() const;
    void clear();

private:
    size_t m_capacity;
    std::vector<cv::Mat> m_frames;
    size_t m_head;
    size_t m_tail;
    std::atomic<size_t> m_size;
    mutable std::mutex m_mutex;
    std::condition_variable m_not_empty;
    std::condition_variable m_not_full;
    size_t nextIndex(size_t current) const;
};
