// vp_detail.cuh - SelectiveBilateral's selective and multi-scale branches:
//   createDetailMask            src/selective_bilateral.cpp:239-313
//   applyJointBilateral's blend src/selective_bilateral.cpp:412-428
//   applyMultiscaleBilateral    src/selective_bilateral.cpp:138-237 (cv::pyrDown :148-150, cv::pyrUp :188, blend :198-228)
// with the REPAIRED semantics of oracle/ref_chain.py: the literal code calls cv::sqrt on the CV_8U `local_var`
// (:276, undefined behaviour); here `local_var` is converted to float32 first, a zero maximum normalises to 0
// instead of NaN, and everything else is as written (including the saturating u8 subtract / multiply at :268-269).
//
// createDetailMask needs two whole-frame maxima (cv::minMaxLoc :281,:284) before any pixel can be normalised, so it
// is two kernels: k_detail_max (integer: max gx^2+gy^2 and max local_var - sqrt is monotonic; it also stores both integers
// per pixel in one 32-bit plane) and k_detail_mask, which normalises that plane, applies the sigmoid and the final blur.  Every stencil border
// is BORDER_REFLECT_101 of its own input; all the kernels involved are mirror symmetric (box, Gaussian, |Sobel|),
// so evaluating them on the reflect-padded gray tile reproduces OpenCV's per-stage padding.
#pragma once
#include "vp_common.cuh"
#include "vp_scene.cuh"   // gray_of

namespace vp {

constexpr int DT_TW = 64, DT_TH = 32, DT_NT = 256;

// gray tile of (DT_TW + 2*HALO) x (DT_TH + 2*HALO) pixels whose top-left corner is image pixel (x0-HALO, y0-HALO)
template <int HALO>
__device__ __forceinline__ void load_gray_tile(const uint8_t* __restrict__ src, size_t pitch, int w, int h, int x0, int y0,
                                               uint8_t* __restrict__ s_gray) {
    constexpr int GW = DT_TW + 2 * HALO, GH = DT_TH + 2 * HALO;
    for (int i = threadIdx.x; i < GW * GH; i += DT_NT) {
        const int ty = i / GW, tx = i - ty * GW;
        const int sx = reflect101(x0 - HALO + tx, w), sy = reflect101(y0 - HALO + ty, h);
        const uint8_t* p = src + (size_t)sy * pitch + (size_t)sx * 3;
        s_gray[ty * GW + tx] = (uint8_t)gray_of((uint32_t)__ldg(p) | ((uint32_t)__ldg(p + 1) << 8) | ((uint32_t)__ldg(p + 2) << 16));
    }
}

// K-tap row sums of a u8 plane: out[y][x] = sum in[y][x..x+K-1], x < iw-K+1
template <int K>
__device__ __forceinline__ void row_sum(const uint8_t* __restrict__ in, int iw, int ih, uint16_t* __restrict__ out) {
    const int ow = iw - (K - 1);
    for (int i = threadIdx.x; i < ow * ih; i += DT_NT) {
        const int y = i / ow, x = i - y * ow;
        const uint8_t* p = in + y * iw + x;
        unsigned s = 0;
#pragma unroll
        for (int k = 0; k < K; ++k) s += p[k];
        out[y * ow + x] = (uint16_t)s;
    }
}
// normalised KxK box filter on u8 (cv::blur / cv::filter2D with ones/K^2) at (x, y) of the row-sum plane:
// round(sum / K^2); sum / K^2 never lands on .5 for odd K
template <int K>
__device__ __forceinline__ int box_at(const uint16_t* __restrict__ rs, int rw, int x, int y) {
    const uint16_t* p = rs + y * rw + x;
    unsigned s = 0;
#pragma unroll
    for (int k = 0; k < K; ++k) s += p[k * rw];
    return (int)((2u * s + (unsigned)(K * K)) / (unsigned)(2 * K * K));
}

struct DetailMax { int g2; int var; };   // max gx^2+gy^2 (Sobel), max local_var; zeroed before k_detail_max

// gray tile (halo H) -> local mean (KxK box) -> saturating diff^2 plane of (DT_TW + 2*H - K + 1) x (DT_TH + 2*H - K + 1)
// -> its row sums: the "local variance" front end createDetailMask (K = 5, src/selective_bilateral.cpp:262-272) and
// calculateTextureMap (K = 7, src/adaptive_sharpening.cpp:368-383) share
template <int HALO, int K>
__device__ __forceinline__ void detail_variance_rows(const uint8_t* s_gray, uint16_t* s_rs1, uint8_t* s_dsq, uint16_t* s_rs2) {
    constexpr int GW = DT_TW + 2 * HALO, GH = DT_TH + 2 * HALO;
    constexpr int MW = GW - (K - 1), MH = GH - (K - 1);            // local_mean / diff_sq plane
    row_sum<K>(s_gray, GW, GH, s_rs1);
    __syncthreads();
    for (int i = threadIdx.x; i < MW * MH; i += DT_NT) {
        const int y = i / MW, x = i - y * MW;
        const int mean = box_at<K>(s_rs1, MW, x, y);
        const int d = max((int)s_gray[(y + K / 2) * GW + x + K / 2] - mean, 0);   // cv::subtract on u8 saturates
        s_dsq[i] = (uint8_t)min(d * d, 255);                                       // cv::multiply on u8 saturates
    }
    __syncthreads();
    row_sum<K>(s_dsq, MW, MH, s_rs2);
    __syncthreads();
}

__device__ __forceinline__ int sobel_g2(const uint8_t* __restrict__ g, int gw) {   // g points at the centre pixel
    const int a = g[-gw - 1], b = g[-gw], c = g[-gw + 1], d = g[-1], f = g[1], p = g[gw - 1], q = g[gw], r = g[gw + 1];
    const int gx = (c - a) + 2 * (f - d) + (r - p), gy = (p - a) + 2 * (q - b) + (r - c);
    return gx * gx + gy * gy;
}

// also stores (gx^2+gy^2) << 8 | local_var per pixel: k_detail_mask reads that plane instead of redoing the front end
__global__ void __launch_bounds__(DT_NT) k_detail_max(const uint8_t* __restrict__ src, size_t pitch, int w, int h, DetailMax* __restrict__ out,
                                                      uint32_t* __restrict__ dv, int dvp) {
    constexpr int HALO = 4, GW = DT_TW + 2 * HALO, GH = DT_TH + 2 * HALO, MW = GW - 4, MH = GH - 4, VW = MW - 4;
    __shared__ uint8_t s_gray[GW * GH];
    __shared__ uint16_t s_rs1[MW * GH];
    __shared__ uint8_t s_dsq[MW * MH];
    __shared__ uint16_t s_rs2[VW * MH];
    __shared__ int s_max[2];
    const int x0 = blockIdx.x * DT_TW, y0 = blockIdx.y * DT_TH;
    if (threadIdx.x < 2) s_max[threadIdx.x] = 0;
    load_gray_tile<HALO>(src, pitch, w, h, x0, y0, s_gray);
    __syncthreads();
    detail_variance_rows<HALO, 5>(s_gray, s_rs1, s_dsq, s_rs2);
    int g2 = 0, var = 0;
    for (int i = threadIdx.x; i < DT_TW * DT_TH; i += DT_NT) {
        const int y = i / DT_TW, x = i - y * DT_TW;
        if (x0 + x < w && y0 + y < h) {
            const int v = box_at<5>(s_rs2, VW, x, y);                           // :272 local_var
            const int q = sobel_g2(s_gray + (y + HALO) * GW + x + HALO, GW);    // :251-256, < 2^22
            var = max(var, v); g2 = max(g2, q);
            dv[(size_t)(y0 + y) * dvp + x0 + x] = ((uint32_t)q << 8) | (uint32_t)v;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        g2 = max(g2, __shfl_xor_sync(0xffffffffu, g2, o));
        var = max(var, __shfl_xor_sync(0xffffffffu, var, o));
    }
    if ((threadIdx.x & 31) == 0) { atomicMax(&s_max[0], g2); atomicMax(&s_max[1], var); }
    __syncthreads();
    if (threadIdx.x == 0) { atomicMax(&out->g2, s_max[0]); atomicMax(&out->var, s_max[1]); }
}

struct DetailMaskArgs {
    const uint32_t* dv; int dvp; int w, h;   // k_detail_max's plane: (gx^2+gy^2) << 8 | local_var
    const DetailMax* mx;
    double thr;                  // detail_threshold / 255.0 (:291)
    float gf[5];                 // (float)getGaussianKernel(5, 1.0) (:306)
    float* mask; size_t mpitch;  // float32 plane, pitch in bytes
};

__global__ void __launch_bounds__(DT_NT) k_detail_mask(const DetailMaskArgs a) {
    constexpr int VW = DT_TW + 4, VH = DT_TH + 4;
    __shared__ float s_sig[VW * VH];
    __shared__ float s_row[DT_TW * VH];
    const int x0 = blockIdx.x * DT_TW, y0 = blockIdx.y * DT_TH;
    // :281-285  Mat / max  ==  convertTo(alpha = 1/max): one float32 multiply by (float)(1.0 / max)
    const float magmax = __fsqrt_rn((float)a.mx->g2), texmax = __fsqrt_rn((float)a.mx->var);
    const float inv_mag = magmax > 0.f ? (float)(1.0 / (double)magmax) : 0.f;
    const float inv_tex = texmax > 0.f ? (float)(1.0 / (double)texmax) : 0.f;
    for (int i = threadIdx.x; i < VW * VH; i += DT_NT) {
        const int y = i / VW, x = i - y * VW;
        // BORDER_REFLECT_101 of the sigmoid plane (:306) == of the per-pixel inputs it is a function of
        const uint32_t q = a.dv[(size_t)reflect101(y0 - 2 + y, a.h) * a.dvp + reflect101(x0 - 2 + x, a.w)];
        const float tex = __fsqrt_rn((float)(q & 0xffu));                                       // :272,:276 (repaired)
        const float mag = __fsqrt_rn((float)(q >> 8));                                          // :251-256
        const float nm = __fmul_rn(mag, inv_mag), nt = __fmul_rn(tex, inv_tex);
        const float detail = (float)((double)nm * 0.7 + (double)nt * 0.3);                      // :288
        s_sig[i] = (float)(1.0 / (1.0 + exp(-((double)detail - a.thr) * 10.0)));                // :296-300 (double, float store)
    }
    __syncthreads();
    for (int i = threadIdx.x; i < DT_TW * VH; i += DT_NT) {                                     // :306 row pass
        const int y = i / DT_TW, x = i - y * DT_TW;
        const float* p = s_sig + y * VW + x;
        float acc = __fmul_rn(p[0], a.gf[0]);
#pragma unroll
        for (int k = 1; k < 5; ++k) acc = __fadd_rn(acc, __fmul_rn(p[k], a.gf[k]));
        s_row[i] = acc;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < DT_TW * DT_TH; i += DT_NT) {                                  // column pass
        const int y = i / DT_TW, x = i - y * DT_TW;
        if (x0 + x >= a.w || y0 + y >= a.h) continue;
        const float* p = s_row + y * DT_TW + x;
        float acc = __fmul_rn(p[0], a.gf[0]);
#pragma unroll
        for (int k = 1; k < 5; ++k) acc = __fadd_rn(acc, __fmul_rn(p[k * DT_TW], a.gf[k]));
        *reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(a.mask) + (size_t)(y0 + y) * a.mpitch + (size_t)(x0 + x) * 4) = acc;
    }
}

// cv::pyrDown on CV_8UC3 (:148-150): [1 4 6 4 1] x [1 4 6 4 1], BORDER_REFLECT_101, (v + 128) >> 8 at even (y, x)
// A thread owns a 2x2 block of outputs: their 5x5 windows overlap in a 7x7 source window (49 pixel loads for four outputs
// instead of 100) and the two horizontal sums of a source row feed both output rows.  Integer sums, so the result is the
// one of the plain per-output evaluation.
__global__ void __launch_bounds__(256) k_pyr_down(const uint8_t* __restrict__ src, size_t spitch, int sw, int sh,
                                                  uint8_t* __restrict__ dst, size_t dpitch, int dw, int dh) {
    const int ox = 2 * (blockIdx.x * 32 + (threadIdx.x & 31)), oy = 2 * (blockIdx.y * 8 + (threadIdx.x >> 5));
    if (ox >= dw || oy >= dh) return;
    int xs[7];
#pragma unroll
    for (int k = 0; k < 7; ++k) xs[k] = 3 * reflect101(2 * ox + k - 2, sw);
    int acc[2][2][3] = {};                                        // [output row][output column][channel]
#pragma unroll
    for (int j = 0; j < 7; ++j) {
        const uint8_t* row = src + (size_t)reflect101(2 * oy + j - 2, sh) * spitch;
        int r0[3], r1[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            int v[7];
#pragma unroll
            for (int k = 0; k < 7; ++k) v[k] = __ldg(row + xs[k] + c);
            r0[c] = v[0] + v[4] + 4 * (v[1] + v[3]) + 6 * v[2];
            r1[c] = v[2] + v[6] + 4 * (v[3] + v[5]) + 6 * v[4];
        }
        const int kw[5] = {1, 4, 6, 4, 1};
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            if (j < 5) { acc[0][0][c] += kw[j] * r0[c]; acc[0][1][c] += kw[j] * r1[c]; }
            if (j >= 2) { acc[1][0][c] += kw[j - 2] * r0[c]; acc[1][1][c] += kw[j - 2] * r1[c]; }
        }
    }
#pragma unroll
    for (int dy = 0; dy < 2; ++dy) {
        if (oy + dy >= dh) break;
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) {
            if (ox + dx >= dw) break;
            uint8_t* o = dst + (size_t)(oy + dy) * dpitch + (size_t)(ox + dx) * 3;
#pragma unroll
            for (int c = 0; c < 3; ++c) o[c] = (uint8_t)((acc[dy][dx][c] + 128) >> 8);
        }
    }
}

// one axis of cv::pyrUp: the (at most three) coarse indices and integer weights behind fine index x.
// Even outputs: s[i-1] + 6 s[i] + s[i+1], odd outputs: 4 (s[i] + s[i+1]); the low neighbour is reflect-101
// (s[-1] = s[1]), the high neighbour replicates (s[n] = s[n-1]).
__device__ __forceinline__ void pyr_up_taps(int x, int n, int idx[3], int wgt[3]) {
    const int i = x >> 1, hi = min(i + 1, n - 1);
    if (x & 1) { idx[0] = i; idx[1] = hi; idx[2] = i; wgt[0] = 4; wgt[1] = 4; wgt[2] = 0; }
    else { idx[0] = i > 0 ? i - 1 : (n > 1 ? 1 : 0); idx[1] = i; idx[2] = hi; wgt[0] = 1; wgt[1] = 6; wgt[2] = 1; }
}

__device__ __forceinline__ void pyr_up_pixel(const uint8_t* __restrict__ coarse, size_t cpitch, int cw, int ch, int x, int y, int up[3]) {
    int xi[3], xw[3], yi[3], yw[3];
    pyr_up_taps(x, cw, xi, xw);
    pyr_up_taps(y, ch, yi, yw);
    up[0] = up[1] = up[2] = 0;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const uint8_t* row = coarse + (size_t)yi[j] * cpitch;
        int r[3] = {0, 0, 0};
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const uint8_t* p = row + 3 * xi[k];
            r[0] += xw[k] * __ldg(p); r[1] += xw[k] * __ldg(p + 1); r[2] += xw[k] * __ldg(p + 2);
        }
        up[0] += yw[j] * r[0]; up[1] += yw[j] * r[1]; up[2] += yw[j] * r[2];
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) up[c] = (up[c] + 32) >> 6;
}

// cv::pyrUp on CV_8UC3 (:188), dst = 2*src or 2*src - 1 per axis
__global__ void __launch_bounds__(256) k_pyr_up(const uint8_t* __restrict__ src, size_t spitch, int sw, int sh,
                                                uint8_t* __restrict__ dst, size_t dpitch, int dw, int dh) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= dw || y >= dh) return;
    int up[3];
    pyr_up_pixel(src, spitch, sw, sh, x, y, up);
    uint8_t* o = dst + (size_t)y * dpitch + (size_t)x * 3;
    o[0] = (uint8_t)up[0]; o[1] = (uint8_t)up[1]; o[2] = (uint8_t)up[2];
}

// applyMultiscaleBilateral's coarse-to-fine step (:186-229): up = cv::pyrUp(coarse, fine.size()) evaluated on the fly,
// fine = saturate_cast<uchar>(fine * d + up * (1 - d)) in float32, in place
struct MsBlendArgs {
    uint8_t* fine; size_t fpitch; int fw, fh;
    const uint8_t* coarse; size_t cpitch; int cw, ch;
    const float* mask; size_t mpitch;
};
// A thread owns the 2x2 block of fine pixels behind one coarse cell: the four pyrUp values share one 3x3 coarse
// neighbourhood (27 byte loads for four pixels instead of 108, the row sums formed once), the integer sums are the very
// ones of pyr_up_pixel.  A2: rows of `fine` are 2-byte aligned and rows of `mask` 8-byte aligned (pixel pairs move as
// three 16-bit words and one float2); otherwise bytes and single floats.
template <bool A2>
__global__ void __launch_bounds__(256) k_ms_blend(const MsBlendArgs a) {
    const int i = blockIdx.x * 32 + (threadIdx.x & 31), j = blockIdx.y * 8 + (threadIdx.x >> 5);
    const int x = 2 * i, y = 2 * j;
    if (x >= a.fw || y >= a.fh) return;
    const int ci = min(i, a.cw - 1), cj = min(j, a.ch - 1);
    const int xs[3] = {ci > 0 ? ci - 1 : (a.cw > 1 ? 1 : 0), ci, min(ci + 1, a.cw - 1)};
    const int ys[3] = {cj > 0 ? cj - 1 : (a.ch > 1 ? 1 : 0), cj, min(cj + 1, a.ch - 1)};
    int he[3][3], ho[3][3];                                      // [coarse row][channel]: even / odd fine column
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const uint8_t* row = a.coarse + (size_t)ys[r] * a.cpitch;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const int lo = __ldg(row + 3 * xs[0] + c), mid = __ldg(row + 3 * xs[1] + c), hi = __ldg(row + 3 * xs[2] + c);
            he[r][c] = lo + 6 * mid + hi;
            ho[r][c] = 4 * (mid + hi);
        }
    }
    const bool x1 = x + 1 < a.fw;
#pragma unroll
    for (int dy = 0; dy < 2; ++dy) {
        if (y + dy >= a.fh) break;
        int up[2][3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            up[0][c] = ((dy ? 4 * (he[1][c] + he[2][c]) : he[0][c] + 6 * he[1][c] + he[2][c]) + 32) >> 6;
            up[1][c] = ((dy ? 4 * (ho[1][c] + ho[2][c]) : ho[0][c] + 6 * ho[1][c] + ho[2][c]) + 32) >> 6;
        }
        uint8_t* f = a.fine + (size_t)(y + dy) * a.fpitch + (size_t)x * 3;
        const float* mrow = reinterpret_cast<const float*>(reinterpret_cast<const uint8_t*>(a.mask) + (size_t)(y + dy) * a.mpitch) + x;
        if (A2 && x1) {
            const float2 d2 = *reinterpret_cast<const float2*>(mrow);
            uint16_t* f16 = reinterpret_cast<uint16_t*>(f);
            const uint32_t w0 = f16[0], w1 = f16[1], w2 = f16[2];
            const int fin[6] = {(int)(w0 & 0xff), (int)(w0 >> 8), (int)(w1 & 0xff), (int)(w1 >> 8), (int)(w2 & 0xff), (int)(w2 >> 8)};
            int o[6];
#pragma unroll
            for (int b = 0; b < 6; ++b) {
                const float d = b < 3 ? d2.x : d2.y;
                o[b] = rne_sat_u8(__fadd_rn(__fmul_rn((float)fin[b], d), __fmul_rn((float)up[b / 3][b % 3], __fsub_rn(1.0f, d))));
            }
            f16[0] = (uint16_t)(o[0] | (o[1] << 8)); f16[1] = (uint16_t)(o[2] | (o[3] << 8)); f16[2] = (uint16_t)(o[4] | (o[5] << 8));
        } else {
#pragma unroll
            for (int dx = 0; dx < 2; ++dx) {
                if (dx && !x1) break;
                const float d = mrow[dx];
                const float cwt = __fsub_rn(1.0f, d);
#pragma unroll
                for (int c = 0; c < 3; ++c)
                    f[3 * dx + c] = (uint8_t)rne_sat_u8(__fadd_rn(__fmul_rn((float)f[3 * dx + c], d), __fmul_rn((float)up[dx][c], cwt)));
            }
        }
    }
}

// applyJointBilateral's per-pixel loop (:412-428): out = sat(orig * d * p + filt * (1 - d * p)), float32, with
// `float p = 1.0f + d * (edge_preserve - 1.0f)` evaluated in double (edge_preserve is a double) and rounded once
struct JointBlendArgs {
    const uint8_t* orig; size_t opitch;
    const uint8_t* filt; size_t fpitch;
    const float* mask; size_t mpitch;
    uint8_t* dst; size_t dpitch;
    int w, h;
    double edge_preserve_m1;     // edge_preserve - 1.0
};
__global__ void __launch_bounds__(256) k_joint_blend(const JointBlendArgs a) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= a.w || y >= a.h) return;
    const float d = *reinterpret_cast<const float*>(reinterpret_cast<const uint8_t*>(a.mask) + (size_t)y * a.mpitch + (size_t)x * 4);
    const float p = (float)(1.0 + (double)d * a.edge_preserve_m1);
    const float dp = __fmul_rn(d, p);
    const float om = __fsub_rn(1.0f, dp);
    const uint8_t* o = a.orig + (size_t)y * a.opitch + (size_t)x * 3;
    const uint8_t* f = a.filt + (size_t)y * a.fpitch + (size_t)x * 3;
    uint8_t* q = a.dst + (size_t)y * a.dpitch + (size_t)x * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float t = __fmul_rn(__fmul_rn((float)__ldg(o + c), d), p);
        q[c] = (uint8_t)rne_sat_u8(__fadd_rn(t, __fmul_rn((float)__ldg(f + c), om)));
    }
}

}  // namespace vp
