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

// A CTA walks whole rows (no per-pixel division), a thread takes four pixels as three words when the rows are word aligned,
// and a warp issues one shared-memory atomic per distinct bin (neighbouring pixels mostly share their bin).
__global__ void __launch_bounds__(256) k_scene_accum(const SceneAccumArgs a) {
    __shared__ unsigned s_hist[64];
    __shared__ unsigned long long s_mad;
    if (threadIdx.x < 64) s_hist[threadIdx.x] = 0;
    if (threadIdx.x == 0) s_mad = 0ull;
    __syncthreads();
    unsigned mad = 0;
    const bool words = ((uintptr_t)a.cur | a.cpitch) % 4 == 0 && (!a.prev || ((uintptr_t)a.prev | a.ppitch) % 4 == 0);
    const int ngroups = (a.w + 3) / 4;
    auto load4 = [&](const uint8_t* row, int gx, int npx, uint32_t (&px)[4]) {
        if (words && npx == 4) {
            const uint32_t* p = reinterpret_cast<const uint32_t*>(row) + 3 * gx;
            const uint32_t w0 = __ldg(p), w1 = __ldg(p + 1), w2 = __ldg(p + 2);
            px[0] = w0 & 0xffffffu; px[1] = (w0 >> 24) | ((w1 & 0xffffu) << 8); px[2] = (w1 >> 16) | ((w2 & 0xffu) << 16); px[3] = w2 >> 8;
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint8_t* p = row + (size_t)(4 * gx + min(j, npx - 1)) * 3;
                px[j] = (uint32_t)__ldg(p) | ((uint32_t)__ldg(p + 1) << 8) | ((uint32_t)__ldg(p + 2) << 16);
            }
        }
    };
    for (int y = blockIdx.x; y < a.h; y += gridDim.x) {
        const uint8_t* crow = a.cur + (size_t)y * a.cpitch;
        const uint8_t* prow = a.prev ? a.prev + (size_t)y * a.ppitch : nullptr;
        for (int gx = threadIdx.x; gx < ngroups; gx += blockDim.x) {
            const int npx = min(4, a.w - 4 * gx);
            uint32_t c[4], q[4] = {0, 0, 0, 0};
            load4(crow, gx, npx, c);
            if (prow) load4(prow, gx, npx, q);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int g = gray_of(c[j]);
                const int bin = j < npx ? g >> 2 : -1;
                const unsigned same = __match_any_sync(__activemask(), bin);
                if (bin >= 0 && (int)(__ffs(same) - 1) == (int)(threadIdx.x & 31)) atomicAdd(&s_hist[bin], (unsigned)__popc(same));
                if (prow && j < npx) mad += abs(g - gray_of(q[j]));
            }
        }
    }
    {   // widen before the warp reduction
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
