// vp_blend.cuh - standalone temporal blend / cv::addWeighted on CV_8UC3 rows
// (src/main.cpp:269-292, TemporalConsistency::blendFrames src/temporal_consistency.cpp:355-444).
// Used when the post-bilateral stage (whose epilogue normally carries the blend) is disabled, and by
// the stage-wise entry points vp_add_weighted / vp_temporal_blend.  Pure streaming: 16 bytes per
// thread per step when every pointer/pitch is 16-byte aligned.
#pragma once
#include "vp_bilateral.cuh"

namespace vp {

struct BlendKArgs {
    const uint8_t* cur; size_t cpitch;
    uint8_t* dst; size_t dpitch;      // may be null: state-only update
    int row_bytes, h;
    BlendArgs blend;                  // prev[], state, weights
    // scene detection: k_scene_decide's verdict caps the number of previous frames (0 = pass-through);
    // inv_wsum_n[n-1] is 1/(w0+..+wn) for n previous frames.  Null = use blend.nprev as given.
    const int* use_dev;
    float inv_wsum_n[3];
};

__global__ void __launch_bounds__(256) k_blend(const BlendKArgs a) {
    BlendArgs b = a.blend;
    if (a.use_dev) {
        b.nprev = min(b.nprev, *a.use_dev);
        if (b.nprev > 0) b.inv_wsum = a.inv_wsum_n[b.nprev - 1];
    }
    const bool doblend = (b.mode != 0 && b.nprev > 0 && a.dst);
    bool al = aligned16(a.cur, a.cpitch);
    if (a.dst) al = al && aligned16(a.dst, a.dpitch);
    if (b.state) al = al && aligned16(b.state, b.spitch);
    if (doblend) {
        al = al && aligned16(b.prev[0], b.ppitch);
        if (b.nprev > 1) al = al && aligned16(b.prev[1], b.ppitch);
        if (b.nprev > 2) al = al && aligned16(b.prev[2], b.ppitch);
    }
    const int nv = al ? a.row_bytes >> 4 : 0;
    const long long nvec = (long long)nv * a.h;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
        const int y = (int)(i / nv);
        const size_t off = (size_t)(i - (long long)y * nv) * 16;
        const uint4 c = *reinterpret_cast<const uint4*>(a.cur + (size_t)y * a.cpitch + off);
        if (b.state) *reinterpret_cast<uint4*>(b.state + (size_t)y * b.spitch + off) = c;
        if (!a.dst) continue;
        uint4 o = c;
        if (doblend) {
            const uint4 p0 = *reinterpret_cast<const uint4*>(b.prev[0] + (size_t)y * b.ppitch + off);
            const uint4 p1 = b.nprev > 1 ? *reinterpret_cast<const uint4*>(b.prev[1] + (size_t)y * b.ppitch + off) : p0;
            const uint4 p2 = b.nprev > 2 ? *reinterpret_cast<const uint4*>(b.prev[2] + (size_t)y * b.ppitch + off) : p0;
            uint32_t r[4];
            r[0] = blend_word(b, c.x, p0.x, p1.x, p2.x); r[1] = blend_word(b, c.y, p0.y, p1.y, p2.y);
            r[2] = blend_word(b, c.z, p0.z, p1.z, p2.z); r[3] = blend_word(b, c.w, p0.w, p1.w, p2.w);
            o = make_uint4(r[0], r[1], r[2], r[3]);
        }
        *reinterpret_cast<uint4*>(a.dst + (size_t)y * a.dpitch + off) = o;
    }
    // byte tail of each row (or everything when unaligned)
    const int t0 = nv * 16, tn = a.row_bytes - t0;
    const long long ntail = (long long)tn * a.h;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < ntail; i += (long long)gridDim.x * blockDim.x) {
        const int y = (int)(i / tn);
        const size_t off = (size_t)t0 + (size_t)(i - (long long)y * tn);
        const int c = a.cur[(size_t)y * a.cpitch + off];
        if (b.state) b.state[(size_t)y * b.spitch + off] = (uint8_t)c;
        if (!a.dst) continue;
        int o = c;
        if (doblend) {
            const int p0 = b.prev[0][(size_t)y * b.ppitch + off];
            const int p1 = b.nprev > 1 ? b.prev[1][(size_t)y * b.ppitch + off] : p0;
            const int p2 = b.nprev > 2 ? b.prev[2][(size_t)y * b.ppitch + off] : p0;
            o = (int)blend_byte(b, (float)c, (float)p0, (float)p1, (float)p2);
        }
        a.dst[(size_t)y * a.dpitch + off] = (uint8_t)o;
    }
}

}  // namespace vp
