// vp_flow.cuh - TemporalConsistency's motion-compensated front end:
//   calculateOpticalFlow          src/temporal_consistency.cpp:174-225  (cv::calcOpticalFlowFarneback :211-216)
//   warpFrame                     src/temporal_consistency.cpp:227-281  (cv::remap INTER_LINEAR / BORDER_REPLICATE :273)
//   calculateFlowReliabilityMask  src/temporal_consistency.cpp:325-353
//   blendFrames with masks        src/temporal_consistency.cpp:355-444
// The Farneback kernels follow OpenCV's published algorithm (modules/video/src/optflowgf.cpp, restated and pinned
// against cv2 in oracle/restate.py:farneback) including its float / double mix: the polynomial expansion's row pass
// and the 15x15 box sums accumulate in double (the 2x2 solve cancels catastrophically in float on edges).
//
// Planes are planar float32 (`ps` = plane stride in floats, `p` = row pitch in floats): R = 5 polynomial coefficients,
// M = 5 structure-tensor entries, flow = 2.  Every kernel is a 32x8 thread block over output pixels; these are
// streaming / small-stencil passes whose operands sit in L2 between launches (HBM-bound at worst).
#pragma once
#include "vp_bilateral.cuh"   // f2u8
#include "vp_common.cuh"
#include "vp_scene.cuh"       // gray_of

namespace vp {

#define VP_FLOW_XY                                                                          \
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5); \
    if (x >= w || y >= h) return

// BGR8 -> gray float32 (:74 cv::cvtColor, then convertTo(CV_32F) inside calcOpticalFlowFarneback)
__global__ void __launch_bounds__(256) k_gray_f32(const uint8_t* __restrict__ src, size_t spitch, float* __restrict__ dst, int p, int w, int h) {
    VP_FLOW_XY;
    const uint8_t* q = src + (size_t)y * spitch + (size_t)x * 3;
    dst[(size_t)y * p + x] = (float)gray_of((uint32_t)__ldg(q) | ((uint32_t)__ldg(q + 1) << 8) | ((uint32_t)__ldg(q + 2) << 16));
}

struct GaussTaps { int ksize; float k[19]; };

// one axis of cv::resize(INTER_LINEAR): source index and fraction behind destination index d
__device__ __forceinline__ void linear_tap(int d, double scale, int sn, int& s0, float& f) {
    float fx = (float)(((double)d + 0.5) * scale - 0.5);
    int sx = (int)floorf(fx);
    fx = __fsub_rn(fx, (float)sx);
    if (sx < 0) { fx = 0.f; sx = 0; }
    if (sx >= sn - 1) { fx = 0.f; sx = sn - 1; }
    s0 = sx; f = fx;
}

// the taps of a whole axis, tabulated once per engine (the double arithmetic above costs more than the resize itself)
struct LinTap { int s0; float f; };
__global__ void k_linear_taps(int n, double scale, int sn, LinTap* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    LinTap t;
    linear_tap(i, scale, sn, t.s0, t.f);
    out[i] = t;
}

// separable float32 Gaussian, BORDER_REFLECT_101, accumulation left to right without FMA (cv::GaussianBlur on CV_32F)
// `varies` (optional): a device flag that is zero when the input plane is known to be the constant 1.0f everywhere (the
// reliability mask of a frame without fast motion); the result is then `const_out`, the value the very same float
// sequence produces on a constant plane (computed by the host), and the taps are skipped.
// KS > 0: kernel size known at compile time (3 = full-resolution pyramid level, 15 = reliability mask), taps unrolled
template <bool ROWS, int KS = 0>
__global__ void __launch_bounds__(256) k_gauss_f32(const float* __restrict__ src, int sp, float* __restrict__ dst, int dp, int w, int h, const GaussTaps t,
                                                   const int* __restrict__ varies = nullptr, float const_out = 0.f) {
    VP_FLOW_XY;
    if (varies && *varies == 0) {          // constant plane: the row pass's output is never read, the column pass stores the constant
        if (!ROWS) dst[(size_t)y * dp + x] = const_out;
        return;
    }
    if constexpr (KS > 0) {
        constexpr int R = KS / 2;
        const int c = ROWS ? x : y, n = ROWS ? w : h;
        float acc = 0.f;
        if (c >= R && c + R < n) {
            const float* p = ROWS ? src + (size_t)y * sp + (x - R) : src + (size_t)(y - R) * sp + x;
#pragma unroll
            for (int i = 0; i < KS; ++i) {
                const float v = ROWS ? p[i] : p[(size_t)i * sp];
                acc = i == 0 ? __fmul_rn(v, t.k[0]) : __fadd_rn(acc, __fmul_rn(v, t.k[i]));
            }
        } else {
#pragma unroll
            for (int i = 0; i < KS; ++i) {
                const float v = ROWS ? src[(size_t)y * sp + reflect101(x + i - R, w)] : src[(size_t)reflect101(y + i - R, h) * sp + x];
                acc = i == 0 ? __fmul_rn(v, t.k[0]) : __fadd_rn(acc, __fmul_rn(v, t.k[i]));
            }
        }
        dst[(size_t)y * dp + x] = acc;
        return;
    }
    const int r = t.ksize >> 1;
    float acc = 0.f;
    const int c = ROWS ? x : y, n = ROWS ? w : h;
    if (c >= r && c + r < n) {                    // interior: no border map per tap
        const float* p = ROWS ? src + (size_t)y * sp + (x - r) : src + (size_t)(y - r) * sp + x;
        const size_t step = ROWS ? 1 : (size_t)sp;
        for (int i = 0; i < t.ksize; ++i) {
            const float v = p[i * step];
            acc = i == 0 ? __fmul_rn(v, t.k[0]) : __fadd_rn(acc, __fmul_rn(v, t.k[i]));
        }
    } else {
        for (int i = 0; i < t.ksize; ++i) {
            const float v = ROWS ? src[(size_t)y * sp + reflect101(x + i - r, w)] : src[(size_t)reflect101(y + i - r, h) * sp + x];
            acc = i == 0 ? __fmul_rn(v, t.k[0]) : __fadd_rn(acc, __fmul_rn(v, t.k[i]));
        }
    }
    dst[(size_t)y * dp + x] = acc;
}

// The 3x3 blur of the full-resolution level as one kernel: a thread owns four vertically adjacent pixels, forms the six row
// results they need in registers (the very float sequence of the row pass: left to right, no FMA, reflect-101) and runs the
// column pass over them - no intermediate plane, 18 loads for four outputs.
__global__ void __launch_bounds__(256) k_gauss3x3_f32(const float* __restrict__ src, int sp, float* __restrict__ dst, int dp, int w, int h,
                                                      const GaussTaps t) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y0 = (blockIdx.y * 8 + (threadIdx.x >> 5)) * 4;
    if (x >= w || y0 >= h) return;
    const int xl = reflect101(x - 1, w), xr = reflect101(x + 1, w);
    float r[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        const float* p = src + (size_t)reflect101(min(y0 + j - 1, h), h) * sp;     // rows beyond h feed outputs that are not stored
        r[j] = __fadd_rn(__fadd_rn(__fmul_rn(__ldg(p + xl), t.k[0]), __fmul_rn(__ldg(p + x), t.k[1])), __fmul_rn(__ldg(p + xr), t.k[2]));
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (y0 + k >= h) break;
        dst[(size_t)(y0 + k) * dp + x] = __fadd_rn(__fadd_rn(__fmul_rn(r[k], t.k[0]), __fmul_rn(r[k + 1], t.k[1])), __fmul_rn(r[k + 2], t.k[2]));
    }
}

// Pyramid levels below full resolution only need the blurred image where cv::resize(INTER_LINEAR) samples it: two source
// columns and two source rows per output pixel.  k_gauss_rows_at runs the row pass at those columns only (compact plane
// of 2*w' columns), k_gauss_cols_resize runs the column pass at the four taps and interpolates - the same arithmetic per
// sample as the full blur followed by the resize, on a fraction of the samples.
template <int KS = 0>
__global__ void __launch_bounds__(256) k_gauss_rows_at(const float* __restrict__ src, int sp, int sw, int sh, float* __restrict__ dst, int dp,
                                                       int dw, const GaussTaps t, const LinTap* __restrict__ tx) {
    const int w = 2 * dw, h = sh;
    VP_FLOW_XY;
    const int x0 = __ldg(&tx[x >> 1].s0);
    const int sx = (x & 1) ? min(x0 + 1, sw - 1) : x0;
    const int ks = KS > 0 ? KS : t.ksize, r = ks >> 1;
    const float* row = src + (size_t)y * sp;
    float acc = 0.f;
    if (sx >= r && sx + r < sw) {
        const float* p = row + sx - r;
#pragma unroll
        for (int i = 0; i < ks; ++i) acc = i == 0 ? __fmul_rn(p[0], t.k[0]) : __fadd_rn(acc, __fmul_rn(p[i], t.k[i]));
    } else {
#pragma unroll
        for (int i = 0; i < ks; ++i) {
            const float v = row[reflect101(sx + i - r, sw)];
            acc = i == 0 ? __fmul_rn(v, t.k[0]) : __fadd_rn(acc, __fmul_rn(v, t.k[i]));
        }
    }
    dst[(size_t)y * dp + x] = acc;
}
template <int KS = 0>
__global__ void __launch_bounds__(256) k_gauss_cols_resize(const float* __restrict__ src, int sp, int sw, int sh, float* __restrict__ dst, int dp,
                                                           int w, int h, const GaussTaps t, const LinTap* __restrict__ tx, const LinTap* __restrict__ ty) {
    VP_FLOW_XY;
    const float fx = __ldg(&tx[x].f);
    const int y0 = __ldg(&ty[y].s0);
    const float fy = __ldg(&ty[y].f);
    const int y1 = min(y0 + 1, sh - 1);
    const int ks = KS > 0 ? KS : t.ksize, r = ks >> 1;
    float v[2][2];
#pragma unroll
    for (int a = 0; a < 2; ++a) {
        const int yy = a ? y1 : y0;
        float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
        for (int i = 0; i < ks; ++i) {
            const float* p = src + (size_t)reflect101(yy + i - r, sh) * sp + 2 * x;
            const float2 q = *reinterpret_cast<const float2*>(p);
            acc0 = i == 0 ? __fmul_rn(q.x, t.k[0]) : __fadd_rn(acc0, __fmul_rn(q.x, t.k[i]));
            acc1 = i == 0 ? __fmul_rn(q.y, t.k[0]) : __fadd_rn(acc1, __fmul_rn(q.y, t.k[i]));
        }
        v[a][0] = acc0; v[a][1] = acc1;
    }
    const float a0 = __fsub_rn(1.f, fx), b0 = __fsub_rn(1.f, fy);
    const float h0 = __fadd_rn(__fmul_rn(v[0][0], a0), __fmul_rn(v[0][1], fx));
    const float h1 = __fadd_rn(__fmul_rn(v[1][0], a0), __fmul_rn(v[1][1], fx));
    dst[(size_t)y * dp + x] = __fadd_rn(__fmul_rn(h0, b0), __fmul_rn(h1, fy));
}

// cv::resize(INTER_LINEAR) on CV_32F planes (nch planes, plane strides sps / dps), optional scale of the result
// (`flow *= 1 / pyr_scale`).  Horizontal interpolation of both rows first, then vertical, all float32 without FMA.
template <int NCH>
__global__ void __launch_bounds__(256) k_resize_linear_f32(const float* __restrict__ src, int sp, size_t sps, int sw, int sh,
                                                           float* __restrict__ dst, int dp, size_t dps, int w, int h, float mul,
                                                           const LinTap* __restrict__ tx, const LinTap* __restrict__ ty) {
    VP_FLOW_XY;
    const LinTap ax = tx[x], ay = ty[y];
    const int x0 = ax.s0, y0 = ay.s0;
    const float fx = ax.f, fy = ay.f;
    const int x1 = min(x0 + 1, sw - 1), y1 = min(y0 + 1, sh - 1);
    const float a0 = __fsub_rn(1.f, fx), b0 = __fsub_rn(1.f, fy);
    const int r0 = y0 * sp, r1 = y1 * sp, o = y * dp + x;          // planes are far below 2^31 elements
    float v00[NCH], v01[NCH], v10[NCH], v11[NCH];
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        const float* s = src + c * sps;
        v00[c] = __ldg(s + r0 + x0); v01[c] = __ldg(s + r0 + x1); v10[c] = __ldg(s + r1 + x0); v11[c] = __ldg(s + r1 + x1);
    }
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        const float h0 = __fadd_rn(__fmul_rn(v00[c], a0), __fmul_rn(v01[c], fx));
        const float h1 = __fadd_rn(__fmul_rn(v10[c], a0), __fmul_rn(v11[c], fx));
        const float v = __fadd_rn(__fmul_rn(h0, b0), __fmul_rn(h1, fy));
        dst[c * dps + o] = mul == 1.f ? v : __fmul_rn(v, mul);
    }
}

struct PolyTaps { int n; float g[8], xg[8], xxg[8]; double ig11, ig03, ig33, ig55; };   // FarnebackPrepareGaussian, taps 0..n

// FarnebackPolyExp, vertical part: three float32 moment rows per pixel (rows clamp at the border)
template <int N = 0>
__global__ void __launch_bounds__(256) k_polyexp_v(const float* __restrict__ src, int sp, float* __restrict__ tmp, int tp, size_t tps, int w, int h, const PolyTaps t) {
    VP_FLOW_XY;
    const int n = N > 0 ? N : t.n;
    const float s0 = src[(size_t)y * sp + x];
    float r0 = __fmul_rn(s0, t.g[0]), r1 = 0.f, r2 = 0.f;
#pragma unroll
    for (int k = 1; k <= n; ++k) {
        const float a = src[(size_t)max(y - k, 0) * sp + x], b = src[(size_t)min(y + k, h - 1) * sp + x];
        const float p = __fadd_rn(a, b);
        r0 = __fadd_rn(r0, __fmul_rn(t.g[k], p));
        r1 = __fadd_rn(r1, __fmul_rn(t.xg[k], __fsub_rn(b, a)));
        r2 = __fadd_rn(r2, __fmul_rn(t.xxg[k], p));
    }
    const size_t o = (size_t)y * tp + x;
    tmp[o] = r0; tmp[tps + o] = r1; tmp[2 * tps + o] = r2;
}
// horizontal part: double accumulators over float32 partial products exactly as the C++ mixes them; columns replicate
template <int N = 0>
__global__ void __launch_bounds__(256) k_polyexp_h(const float* __restrict__ tmp, int tp, size_t tps, float* __restrict__ R, int rp, size_t rps, int w, int h, const PolyTaps t) {
    VP_FLOW_XY;
    const int n = N > 0 ? N : t.n;
    const float* r0 = tmp + (size_t)y * tp;
    const float* r1 = r0 + tps;
    const float* r2 = r0 + 2 * tps;
    double b1 = (double)__fmul_rn(r0[x], t.g[0]), b2 = 0, b3 = (double)__fmul_rn(r1[x], t.g[0]), b4 = 0, b5 = (double)__fmul_rn(r2[x], t.g[0]), b6 = 0;
#pragma unroll
    for (int k = 1; k <= n; ++k) {
        const int xp = min(x + k, w - 1), xm = max(x - k, 0);
        const double tg = (double)__fadd_rn(r0[xp], r0[xm]);
        b1 += tg * (double)t.g[k];
        b4 += tg * (double)t.xxg[k];
        b2 += (double)__fmul_rn(__fsub_rn(r0[xp], r0[xm]), t.xg[k]);
        b3 += (double)__fmul_rn(__fadd_rn(r1[xp], r1[xm]), t.g[k]);
        b6 += (double)__fmul_rn(__fsub_rn(r1[xp], r1[xm]), t.xg[k]);
        b5 += (double)__fmul_rn(__fadd_rn(r2[xp], r2[xm]), t.g[k]);
    }
    const size_t o = (size_t)y * rp + x;
    R[rps + o] = (float)(b2 * t.ig11);
    R[o] = (float)(b3 * t.ig11);
    R[3 * rps + o] = (float)(b1 * t.ig03 + b4 * t.ig33);
    R[2 * rps + o] = (float)(b1 * t.ig03 + b5 * t.ig33);
    R[4 * rps + o] = (float)(b6 * t.ig55);
}

// FarnebackUpdateMatrices: the second frame's expansion sampled at x + flow (bilinear), structure tensor M
__global__ void __launch_bounds__(256) k_update_matrices(const float* __restrict__ R0, const float* __restrict__ R1, int rp, size_t rps,
                                                         const float* __restrict__ flow, int fp, size_t fps,
                                                         float* __restrict__ M, int mp, size_t mps, int w, int h) {
    VP_FLOW_XY;
    const size_t o = (size_t)y * rp + x;
    const float dx = flow[(size_t)y * fp + x], dy = flow[fps + (size_t)y * fp + x];
    float fx = __fadd_rn((float)x, dx), fy = __fadd_rn((float)y, dy);
    const int x1 = (int)floorf(fx), y1 = (int)floorf(fy);
    fx = __fsub_rn(fx, (float)x1); fy = __fsub_rn(fy, (float)y1);
    float r2, r3, r4, r5, r6;
    if ((unsigned)x1 < (unsigned)(w - 1) && (unsigned)y1 < (unsigned)(h - 1)) {
        const float a00 = __fmul_rn(__fsub_rn(1.f, fx), __fsub_rn(1.f, fy)), a01 = __fmul_rn(fx, __fsub_rn(1.f, fy));
        const float a10 = __fmul_rn(__fsub_rn(1.f, fx), fy), a11 = __fmul_rn(fx, fy);
        const size_t q = (size_t)y1 * rp + x1;
        float v[5];
#pragma unroll
        for (int c = 0; c < 5; ++c) {
            const float* p = R1 + c * rps + q;
            float s = __fmul_rn(a00, p[0]);
            s = __fadd_rn(s, __fmul_rn(a01, p[1]));
            s = __fadd_rn(s, __fmul_rn(a10, p[rp]));
            v[c] = __fadd_rn(s, __fmul_rn(a11, p[rp + 1]));
        }
        r2 = v[0]; r3 = v[1];
        r4 = __fmul_rn(__fadd_rn(R0[2 * rps + o], v[2]), 0.5f);
        r5 = __fmul_rn(__fadd_rn(R0[3 * rps + o], v[3]), 0.5f);
        r6 = __fmul_rn(__fadd_rn(R0[4 * rps + o], v[4]), 0.25f);
    } else {
        r2 = r3 = 0.f;
        r4 = R0[2 * rps + o]; r5 = R0[3 * rps + o]; r6 = __fmul_rn(R0[4 * rps + o], 0.5f);
    }
    r2 = __fmul_rn(__fsub_rn(R0[o], r2), 0.5f);
    r3 = __fmul_rn(__fsub_rn(R0[rps + o], r3), 0.5f);
    r2 = __fadd_rn(r2, __fadd_rn(__fmul_rn(r4, dy), __fmul_rn(r6, dx)));
    r3 = __fadd_rn(r3, __fadd_rn(__fmul_rn(r6, dy), __fmul_rn(r5, dx)));
    if ((unsigned)(x - 5) >= (unsigned)(w - 10) || (unsigned)(y - 5) >= (unsigned)(h - 10)) {
        const float border[5] = {0.14f, 0.14f, 0.4472f, 0.4472f, 0.4472f};
        float scale = x < 5 ? border[x] : 1.f;
        scale = __fmul_rn(scale, x >= w - 5 ? border[w - x - 1] : 1.f);
        scale = __fmul_rn(scale, y < 5 ? border[y] : 1.f);
        scale = __fmul_rn(scale, y >= h - 5 ? border[h - y - 1] : 1.f);
        r2 = __fmul_rn(r2, scale); r3 = __fmul_rn(r3, scale); r4 = __fmul_rn(r4, scale); r5 = __fmul_rn(r5, scale); r6 = __fmul_rn(r6, scale);
    }
    const size_t m = (size_t)y * mp + x;
    M[m] = __fadd_rn(__fmul_rn(r4, r4), __fmul_rn(r6, r6));
    M[mps + m] = __fmul_rn(__fadd_rn(r4, r5), r6);
    M[2 * mps + m] = __fadd_rn(__fmul_rn(r5, r5), __fmul_rn(r6, r6));
    M[3 * mps + m] = __fadd_rn(__fmul_rn(r4, r2), __fmul_rn(r6, r3));
    M[4 * mps + m] = __fadd_rn(__fmul_rn(r6, r2), __fmul_rn(r5, r3));
}

// FarnebackUpdateFlow_Blur: (2m+1)^2 box sums of M in double (rows / columns replicate), then the 2x2 solve.
// One CTA per 32x32 tile, one plane of M at a time: (1) the plane's (32+2m)^2 halo tile is staged in shared memory with
// coalesced, independent loads; (2) a thread per (tile row, 8-column segment) forms the row sums with a running double sum
// (one add, one subtract per step - OpenCV's own scheme); (3) every thread owns four vertically adjacent pixels and slides
// the column window over the row sums.  ~10 shared-memory accesses per pixel and plane instead of 2 x (2m+1); the five sums
// stay in registers, M is read ~2x, nothing is spilled to HBM.  Shared pitches are odd so lanes walking rows hit distinct banks.
constexpr int BOX_T = 32, BOX_SEG = 8;
__host__ __device__ inline size_t box_smem_bytes(int m, bool two_tiles = false) {
    const int ch = BOX_T + 2 * m;
    return (size_t)ch * (BOX_T + 1) * sizeof(double) + (size_t)(two_tiles ? 2 : 1) * ch * (ch | 1) * sizeof(float);
}
__device__ __forceinline__ void cp_async_f32(float* smem, const float* g) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
// MT > 0: window half-width known at compile time (the reference's winsize 15 -> MT = 7), loops unrolled, constant divisors
template <int MT>
__global__ void __launch_bounds__(256) k_box_solve(const float* __restrict__ M, int mp, size_t mps, float* __restrict__ flow, int fp, size_t fps,
                                                   int w, int h, int m_rt, double scale) {
    extern __shared__ double s_r[];                       // [ch][BOX_T + 1] row sums, then the float halo tile [ch][cp]
    const int m = MT > 0 ? MT : m_rt;
    const int ch = BOX_T + 2 * m, cp = ch | 1, rp = BOX_T + 1, win = 2 * m;
    float* s_m = reinterpret_cast<float*>(s_r + ch * rp);
    const int x0 = blockIdx.x * BOX_T, y0 = blockIdx.y * BOX_T;
    const int lx = threadIdx.x & 31, r0 = (threadIdx.x >> 5) * 4;
    double acc[4][5];
    // staging addresses are the same for the five planes: with the window known at compile time a warp owns tile rows
    // wrp, wrp+8, ... and a lane the columns lane and lane+32, all clamped once (rows / columns replicate)
    constexpr int NRW = MT > 0 ? (BOX_T + 2 * MT + 7) / 8 : 1;
    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
    size_t roff[NRW];
    int cx0 = 0, cx1 = 0;
    if (MT > 0) {
        static_assert(MT == 0 || BOX_T + 2 * MT <= 64, "two columns per lane");
        cx0 = min(max(x0 - m + lane, 0), w - 1);
        cx1 = min(max(x0 - m + lane + 32, 0), w - 1);
#pragma unroll
        for (int k = 0; k < NRW; ++k) roff[k] = (size_t)min(max(y0 - m + wrp + 8 * k, 0), h - 1) * mp;
    }
    // MT > 0: two halo tiles; plane c + 1 streams in (cp.async, no registers) while plane c is summed
    auto stage_async = [&](int c) {
        const float* p = M + c * mps;
        float* d = s_m + (c & 1) * ch * cp;
#pragma unroll
        for (int k = 0; k < NRW; ++k) {
            const int ty = wrp + 8 * k;
            if (ty < ch) {
                cp_async_f32(d + ty * cp + lane, p + roff[k] + cx0);
                if (lane + 32 < ch) cp_async_f32(d + ty * cp + lane + 32, p + roff[k] + cx1);
            }
        }
        cp_async_commit();
    };
    if (MT > 0) stage_async(0);
#pragma unroll
    for (int c = 0; c < 5; ++c) {
        const float* p = M + c * mps;
        const float* tile = s_m + (MT > 0 ? (c & 1) * ch * cp : 0);
        if (MT > 0) {
            if (c + 1 < 5) { stage_async(c + 1); cp_async_wait<1>(); } else cp_async_wait<0>();
        } else {
            for (int i = threadIdx.x; i < ch * ch; i += 256) {
                const int ty = i / ch, tx = i - ty * ch;
                s_m[ty * cp + tx] = p[(size_t)min(max(y0 - m + ty, 0), h - 1) * mp + min(max(x0 - m + tx, 0), w - 1)];
            }
        }
        __syncthreads();
        for (int task = threadIdx.x; task < ch * (BOX_T / BOX_SEG); task += 256) {
            const int seg = task / ch, row = task - seg * ch;
            const float* q = tile + row * cp + seg * BOX_SEG;          // first column of the first window
            double s = 0.0;
            float head[BOX_SEG - 1];                                   // the values the window drops again: read once
#pragma unroll
            for (int k = 0; k <= win; ++k) {
                const float v = q[k];
                if (k < BOX_SEG - 1) head[k] = v;
                s += (double)v;
            }
            double* o = s_r + row * rp + seg * BOX_SEG;
            o[0] = s;
#pragma unroll
            for (int t = 1; t < BOX_SEG; ++t) {
                s += (double)q[t + win];
                s -= (double)(MT > 0 && MT * 2 + 1 >= BOX_SEG - 1 ? head[t - 1] : q[t - 1]);
                o[t] = s;
            }
        }
        __syncthreads();
        const double* v = s_r + r0 * rp + lx;                          // window rows r0 .. r0 + 2m of the tile's halo rows
        double a = 0.0, top[3];
#pragma unroll
        for (int k = 0; k <= win; ++k) {
            const double u = v[k * rp];
            if (k < 3) top[k] = u;
            a += u;
        }
        acc[0][c] = a * scale;
#pragma unroll
        for (int t = 1; t < 4; ++t) {
            a += v[(t + win) * rp];
            a -= (MT > 0 ? top[t - 1] : v[(t - 1) * rp]);
            acc[t][c] = a * scale;
        }
    }
    const int x = x0 + lx;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const int y = y0 + r0 + t;
        if (x >= w || y >= h) continue;
        const double* s = acc[t];
        const double idet = 1.0 / (s[0] * s[2] - s[1] * s[1] + 1e-3);
        flow[(size_t)y * fp + x] = (float)((s[0] * s[4] - s[1] * s[3]) * idet);
        flow[fps + (size_t)y * fp + x] = (float)((s[2] * s[3] - s[1] * s[4]) * idet);
    }
}

// cv::remap(frame, map = (x, y) + flow, INTER_LINEAR, BORDER_REPLICATE) on BGR8, bit-exact: coordinates rounded to
// 1/32 pixel, taps clamped into the image, integer weights summing to 2^15, (sum + 2^14) >> 15
__global__ void __launch_bounds__(256) k_warp(const uint8_t* __restrict__ src, size_t spitch, const float* __restrict__ flow_x,
                                              const float* __restrict__ flow_y, int fp, uint8_t* __restrict__ dst, size_t dpitch, int w, int h) {
    VP_FLOW_XY;
    const float mx = __fadd_rn((float)x, flow_x[(size_t)y * fp + x]), my = __fadd_rn((float)y, flow_y[(size_t)y * fp + x]);
    const int sx = __float2int_rn(__fmul_rn(mx, 32.f)), sy = __float2int_rn(__fmul_rn(my, 32.f));
    const int fx = sx & 31, fy = sy & 31;
    const int ix = min(max(sx >> 5, -32768), 32767), iy = min(max(sy >> 5, -32768), 32767);
    const int x0 = min(max(ix, 0), w - 1), x1 = min(max(ix + 1, 0), w - 1), y0 = min(max(iy, 0), h - 1), y1 = min(max(iy + 1, 0), h - 1);
    const int w00 = 32 * (32 - fx) * (32 - fy), w01 = 32 * fx * (32 - fy), w10 = 32 * (32 - fx) * fy, w11 = 32 * fx * fy;
    const uint8_t* r0 = src + (size_t)y0 * spitch;
    const uint8_t* r1 = src + (size_t)y1 * spitch;
    uint8_t* o = dst + (size_t)y * dpitch + (size_t)x * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const int v = __ldg(r0 + 3 * x0 + c) * w00 + __ldg(r0 + 3 * x1 + c) * w01 + __ldg(r1 + 3 * x0 + c) * w10 + __ldg(r1 + 3 * x1 + c) * w11;
        o[c] = (uint8_t)min(max((v + (1 << 14)) >> 15, 0), 255);
    }
}

// calculateFlowReliabilityMask before its blur (:331-344): 1 where |f| <= threshold, else clamp(expf(-(|f| - thr) / 10), 0, 1)
__global__ void __launch_bounds__(256) k_reliability(const float* __restrict__ flow_x, const float* __restrict__ flow_y, int fp,
                                                     float* __restrict__ mask, int mp, int w, int h, float thr, int* __restrict__ varies) {
    VP_FLOW_XY;
    const float fx = flow_x[(size_t)y * fp + x], fy = flow_y[(size_t)y * fp + x];
    const float mag = __fsqrt_rn(__fadd_rn(__fmul_rn(fx, fx), __fmul_rn(fy, fy)));
    float r = 1.f;
    if (mag > thr) r = fminf(fmaxf(expf(__fdiv_rn(-__fsub_rn(mag, thr), 10.0f)), 0.f), 1.f);
    mask[(size_t)y * mp + x] = r;
    if (r != 1.f) *varies = 1;          // benign race: every writer stores the same value
}

// blendFrames with reliability masks (:380-436): weight_i = e^{-i/2} * mask_i * blend_factor_i in float32, accumulation
// in frame order, `pixel /= weight` (multiplication by 1.f / weight), convertTo RNE.  frames[0] = current (mask 1).
struct MaskedBlendArgs {
    const uint8_t* cur; size_t cpitch;
    const uint8_t* prev[3]; size_t ppitch;      // warped previous frames
    const float* mask[3]; int mp;               // their reliability masks (pitch in floats)
    int nprev;
    const int* use_dev;                         // scene verdict: caps nprev (0 = pass the frame through); may be null
    float fw[4];                                // e^{-i/2} (float), blend_strength
    float blend_strength;
    uint8_t* dst; size_t dpitch;
    int w, h;
};
__global__ void __launch_bounds__(256) k_blend_masked(const MaskedBlendArgs a) {
    const int w = a.w, h = a.h;
    VP_FLOW_XY;
    int n = a.nprev;
    if (a.use_dev) n = min(n, *a.use_dev);
    const uint8_t* c = a.cur + (size_t)y * a.cpitch + (size_t)x * 3;
    uint8_t* o = a.dst + (size_t)y * a.dpitch + (size_t)x * 3;
    if (n <= 0) { o[0] = c[0]; o[1] = c[1]; o[2] = c[2]; return; }
    const float w0 = __fmul_rn(__fmul_rn(a.fw[0], 1.0f), 1.0f);
    float acc[3] = {__fmul_rn(w0, (float)c[0]), __fmul_rn(w0, (float)c[1]), __fmul_rn(w0, (float)c[2])};
    float wsum = w0;
    for (int i = 0; i < n; ++i) {
        const float wt = __fmul_rn(__fmul_rn(a.fw[i + 1], a.mask[i][(size_t)y * a.mp + x]), a.blend_strength);
        const uint8_t* p = a.prev[i] + (size_t)y * a.ppitch + (size_t)x * 3;
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) acc[ch] = __fadd_rn(acc[ch], __fmul_rn(wt, (float)p[ch]));
        wsum = __fadd_rn(wsum, wt);
    }
    const float inv = __fdiv_rn(1.0f, wsum);
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) o[ch] = (uint8_t)f2u8(wsum > 0.f ? __fmul_rn(acc[ch], inv) : (float)c[ch]);
}

// the same arithmetic on four pixels per thread (12 bytes = three words per frame, one float4 per mask): for frames whose
// rows are word aligned and whose width is a multiple of four
__global__ void __launch_bounds__(256) k_blend_masked4(const MaskedBlendArgs a) {
    const int x4 = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x4 * 4 >= a.w || y >= a.h) return;
    int n = a.nprev;
    if (a.use_dev) n = min(n, *a.use_dev);
    const uint32_t* c = reinterpret_cast<const uint32_t*>(a.cur + (size_t)y * a.cpitch) + x4 * 3;
    uint32_t* o = reinterpret_cast<uint32_t*>(a.dst + (size_t)y * a.dpitch) + x4 * 3;
    const uint32_t cw[3] = {__ldg(c), __ldg(c + 1), __ldg(c + 2)};
    if (n <= 0) { o[0] = cw[0]; o[1] = cw[1]; o[2] = cw[2]; return; }
    const float w0 = __fmul_rn(__fmul_rn(a.fw[0], 1.0f), 1.0f);
    float acc[12], wsum[4] = {w0, w0, w0, w0};
#pragma unroll
    for (int b = 0; b < 12; ++b) acc[b] = __fmul_rn(w0, (float)((cw[b >> 2] >> (8 * (b & 3))) & 0xffu));
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        if (i < n) {
            const float4 m = __ldg(reinterpret_cast<const float4*>(a.mask[i] + (size_t)y * a.mp) + x4);
            const uint32_t* p = reinterpret_cast<const uint32_t*>(a.prev[i] + (size_t)y * a.ppitch) + x4 * 3;
            const uint32_t pw[3] = {__ldg(p), __ldg(p + 1), __ldg(p + 2)};
            const float wt[4] = {__fmul_rn(__fmul_rn(a.fw[i + 1], m.x), a.blend_strength), __fmul_rn(__fmul_rn(a.fw[i + 1], m.y), a.blend_strength),
                                 __fmul_rn(__fmul_rn(a.fw[i + 1], m.z), a.blend_strength), __fmul_rn(__fmul_rn(a.fw[i + 1], m.w), a.blend_strength)};
#pragma unroll
            for (int b = 0; b < 12; ++b) acc[b] = __fadd_rn(acc[b], __fmul_rn(wt[b / 3], (float)((pw[b >> 2] >> (8 * (b & 3))) & 0xffu)));
#pragma unroll
            for (int j = 0; j < 4; ++j) wsum[j] = __fadd_rn(wsum[j], wt[j]);
        }
    }
    float inv[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) inv[j] = __fdiv_rn(1.0f, wsum[j]);
    uint32_t ow[3] = {0, 0, 0};
#pragma unroll
    for (int b = 0; b < 12; ++b) {
        const uint32_t cb = (cw[b >> 2] >> (8 * (b & 3))) & 0xffu;
        ow[b >> 2] |= (uint32_t)f2u8(wsum[b / 3] > 0.f ? __fmul_rn(acc[b], inv[b / 3]) : (float)cb) << (8 * (b & 3));
    }
    o[0] = ow[0]; o[1] = ow[1]; o[2] = ow[2];
}

}  // namespace vp
