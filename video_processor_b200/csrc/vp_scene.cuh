// vp_scene.cuh - TemporalConsistency::detectSceneChange (reference src/temporal_consistency.cpp:283-323) and
// the buffer reset it triggers (:100-112), decided on the device so the frame loop never waits for the host.
//   combined = (1 - corr(hist64(prev_gray), hist64(cur_gray))) * 100 + mean|prev_gray - cur_gray| * 0.5 > threshold
// The 64-bin histogram of the current frame and the sum of absolute gray differences are accumulated by the
// kernel that produces the frame (k_bilateral's epilogue) or by k_scene_accum; k_scene_decide then normalises
// (float32, as cv::normalize), correlates (double, as cv::compareHist) and publishes how many previous frames
// the blend may use.  Follows oracle/restate.py:scene_metrics (pinned against cv2).
#pragma once
#include "vp_common.cuh"

namespace vp {

constexpr int SC_COPIES = 128;                    // privatised global histograms: same-line L2 atomics serialise, spread them

struct SceneAccum {                               // device memory, zeroed by k_scene_decide after use
    unsigned hist[SC_COPIES][64];
    unsigned long long mad;
};
struct SceneState {                               // device memory, persistent across frames
    int nframes;                                  // frames currently in the (conceptual) buffer
    float prev_hist[64];                          // raw counts of the previous frame
};

__device__ __forceinline__ int gray_of(uint32_t bgr) {
    return (9798 * (int)((bgr >> 16) & 0xff) + 19235 * (int)((bgr >> 8) & 0xff) + 3735 * (int)(bgr & 0xff) + (1 << 14)) >> 15;
}

// one CTA, 64 threads
__global__ void k_scene_decide(SceneAccum* __restrict__ acc, SceneState* __restrict__ st, int* __restrict__ use_out,
                               double npx, float threshold, int buffer_size, int reset) {
    __shared__ float s_cur[64], s_prev[64];
    const int b = threadIdx.x;
    unsigned cnt = 0;
#pragma unroll
    for (int k = 0; k < SC_COPIES; ++k) { cnt += acc->hist[k][b]; acc->hist[k][b] = 0; }
    s_cur[b] = (float)cnt;
    s_prev[b] = st->prev_hist[b];
    __syncthreads();
    if (b == 0) {
        int nframes = reset ? 0 : st->nframes;
        int use = 0;
        if (nframes > 0) {
            float hn[2][64];
            for (int w = 0; w < 2; ++w) {             // cv::normalize(h, h, 0, 1, NORM_MINMAX) on CV_32F
                const float* h = w ? s_cur : s_prev;
                float mn = h[0], mx = h[0];
                for (int i = 1; i < 64; ++i) { mn = fminf(mn, h[i]); mx = fmaxf(mx, h[i]); }
                const double d = (double)mx - (double)mn;
                const double scale = d > 2.220446049250313e-16 ? 1.0 / d : 0.0;
                const float fs = (float)scale, fsh = (float)(0.0 - (double)mn * scale);
                for (int i = 0; i < 64; ++i) hn[w][i] = __fadd_rn(__fmul_rn(h[i], fs), fsh);
            }
            double s1 = 0, s2 = 0, s11 = 0, s22 = 0, s12 = 0;     // cv::compareHist(HISTCMP_CORREL)
            for (int i = 0; i < 64; ++i) {
                const double a = hn[0][i], c = hn[1][i];
                s1 += a; s2 += c; s11 += a * a; s22 += c * c; s12 += a * c;
            }
            const double num = s12 - s1 * s2 / 64.0;
            const double den = (s11 - s1 * s1 / 64.0) * (s22 - s2 * s2 / 64.0);
            const double corr = fabs(den) > 2.220446049250313e-16 ? num / sqrt(den) : 1.0;
            const double mad = (double)acc->mad / npx;
            const double combined = (1.0 - corr) * 100.0 + mad * 0.5;
            if (combined > (double)threshold) nframes = 0;          // cut: history cleared, frame passes through
            else use = min(nframes, buffer_size);                   // every buffered frame has a flow slot (:127, :167-169)
        }
        st->nframes = min(nframes + 1, buffer_size);
        acc->mad = 0ull;
        *use_out = use;
    }
    st->prev_hist[b] = s_cur[b];
}

// stand-alone accumulation (chains whose last spatial stage is not k_bilateral): streams cur and prev
struct SceneAccumArgs {
    const uint8_t* cur; size_t cpitch;
    const uint8_t* prev; size_t ppitch;             // may be null (no previous frame yet)
    int w, h;
    SceneAccum* acc;
};

__global__ void __launch_bounds__(256) k_scene_accum(const SceneAccumArgs a) {
    __shared__ unsigned s_hist[64];
    __shared__ unsigned long long s_mad;
    if (threadIdx.x < 64) s_hist[threadIdx.x] = 0;
    if (threadIdx.x == 0) s_mad = 0ull;
    __syncthreads();
    unsigned mad = 0;
    const long long total = (long long)a.w * a.h;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int y = (int)(i / a.w), x = (int)(i - (long long)y * a.w);
        const uint8_t* p = a.cur + (size_t)y * a.cpitch + (size_t)x * 3;
        const int g = gray_of((uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16));
        atomicAdd(&s_hist[g >> 2], 1u);
        if (a.prev) {
            const uint8_t* q = a.prev + (size_t)y * a.ppitch + (size_t)x * 3;
            mad += abs(g - gray_of((uint32_t)q[0] | ((uint32_t)q[1] << 8) | ((uint32_t)q[2] << 16)));
        }
    }
    {   // per-thread sums fit 32 bits only for short strides: widen before the warp reduction
        unsigned long long m = mad;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m += __shfl_xor_sync(0xffffffffu, m, o);
        if ((threadIdx.x & 31) == 0 && m) atomicAdd(&s_mad, m);
    }
    __syncthreads();
    if (threadIdx.x < 64 && s_hist[threadIdx.x]) atomicAdd(&a.acc->hist[blockIdx.x % SC_COPIES][threadIdx.x], s_hist[threadIdx.x]);
    if (threadIdx.x == 0 && s_mad) atomicAdd(&a.acc->mad, s_mad);
}

}  // namespace vp
