// vp_varsharp.cuh - AdaptiveSharpening with adaptive_sigma = true:
//   calculateTextureMap            src/adaptive_sharpening.cpp:357-403
//   calculateAdaptiveSigma         src/adaptive_sharpening.cpp:405-440
//   applyVariableSigmaUnsharpMask  src/adaptive_sharpening.cpp:442-592
// with the REPAIRED semantics of oracle/ref_chain.py:sharpen_variable_sigma: the literal code calls cv::sqrt on the
// CV_8U `local_var` (:386, undefined behaviour) - here it is converted to float32 first - and degenerate ranges
// (max == min) give 0 instead of NaN.
//
// Two whole-frame reductions gate the data flow (cv::minMaxLoc of the texture map :390 and of the sigma map :457),
// so the stage is four kernels; the reductions are integer (sqrt and float ordering are monotonic) and the five
// sigma-dependent Q8 Gaussian kernels are built ON THE DEVICE from the second reduction (k_sigma_levels), so the
// host never waits:
//   k_texture_minmax (+ the u8 local-variance plane) -> k_sigma_map (+ min/max of the blurred sigma map) -> k_sigma_levels
//   -> k_varsharp
#pragma once
#include "vp_bilateral.cuh"   // f2u8
#include "vp_detail.cuh"
#include "vp_stats.cuh"

namespace vp {

constexpr int VS_LEVELS = 5;     // num_sigma_levels (:451)

struct SigmaStats {              // reset to {INT_MAX, 0, 0x7f800000, 0} before k_texture_minmax
    int var_min, var_max;        // extremes of the u8 local variance (sqrt is monotonic)
    int sig_min, sig_max;        // extremes of the blurred sigma map as float bit patterns (positive floats order as ints)
};
struct SigmaLevels {
    float mn, span;              // min_sigma, max_sigma - min_sigma (:459-460)
    int q[VS_LEVELS][7];         // Q0.8 taps of cv::GaussianBlur(u8, ksize, sigma_i) per level
};

__global__ void __launch_bounds__(DT_NT) k_texture_minmax(const uint8_t* __restrict__ src, size_t pitch, int w, int h, SigmaStats* __restrict__ st,
                                                          uint8_t* __restrict__ var_out, size_t vpitch) {
    constexpr int HALO = 6, K = 7, GW = DT_TW + 2 * HALO, GH = DT_TH + 2 * HALO, MW = GW - (K - 1), MH = GH - (K - 1), VW = MW - (K - 1);
    __shared__ uint8_t s_gray[GW * GH];
    __shared__ uint16_t s_rs1[MW * GH];
    __shared__ uint8_t s_dsq[MW * MH];
    __shared__ uint16_t s_rs2[VW * MH];
    __shared__ int s_mm[2];
    const int x0 = blockIdx.x * DT_TW, y0 = blockIdx.y * DT_TH;
    if (threadIdx.x == 0) { s_mm[0] = 0x7fffffff; s_mm[1] = 0; }
    load_gray_tile<HALO>(src, pitch, w, h, x0, y0, s_gray);
    __syncthreads();
    detail_variance_rows<HALO, K>(s_gray, s_rs1, s_dsq, s_rs2);
    int lo = 0x7fffffff, hi = 0;
    for (int i = threadIdx.x; i < DT_TW * DT_TH; i += DT_NT) {
        const int y = i / DT_TW, x = i - y * DT_TW;
        if (x0 + x < w && y0 + y < h) {
            const int v = box_at<K>(s_rs2, VW, x, y);
            lo = min(lo, v); hi = max(hi, v);
            var_out[(size_t)(y0 + y) * vpitch + x0 + x] = (uint8_t)v;      // kept for k_sigma_map: the front end runs once
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if ((threadIdx.x & 31) == 0) { atomicMin(&s_mm[0], lo); atomicMax(&s_mm[1], hi); }
    __syncthreads();
    if (threadIdx.x == 0) { atomicMin(&st->var_min, s_mm[0]); atomicMax(&st->var_max, s_mm[1]); }
}

struct SigmaMapArgs {
    const uint8_t* var; size_t vpitch; int w, h;   // k_texture_minmax's local-variance plane
    SigmaStats* st;
    float gf[5];                 // (float)getGaussianKernel(5, 1.0) (:431)
    float* sigma; size_t spitch; // float32 plane, pitch in bytes
};

// texture = sqrt(local_var) (repaired) -> normalise by the frame's extremes -> sigma = 2.5 - 1.7 t (:423) -> 5x5 sigma 1.0
// float Gaussian (:431, BORDER_REFLECT_101 of the sigma plane == of the variance plane) + min / max of the result
__global__ void __launch_bounds__(DT_NT) k_sigma_map(const SigmaMapArgs a) {
    constexpr int VW = DT_TW + 4, VH = DT_TH + 4;
    __shared__ float s_sig[VW * VH];
    __shared__ float s_row[DT_TW * VH];
    __shared__ int s_mm[2];
    const int x0 = blockIdx.x * DT_TW, y0 = blockIdx.y * DT_TH;
    if (threadIdx.x == 0) { s_mm[0] = 0x7f800000; s_mm[1] = 0; }
    // :390-391  (t - min) / (max - min) is one MatExpr: convertTo(alpha, beta) = t * alpha + beta in float32
    const float tmin = __fsqrt_rn((float)a.st->var_min), tmax = __fsqrt_rn((float)a.st->var_max);
    const bool flat = !(tmax > tmin);
    const double inv = flat ? 0.0 : 1.0 / ((double)tmax - (double)tmin);
    const float na = (float)inv, nb = (float)(-(double)tmin * inv);
    const float smax = 2.5f, sspan = __fsub_rn(2.5f, 0.8f);                                    // :412-413
    for (int i = threadIdx.x; i < VW * VH; i += DT_NT) {
        const int y = i / VW, x = i - y * VW;
        const int v = a.var[(size_t)reflect101(y0 - 2 + y, a.h) * a.vpitch + reflect101(x0 - 2 + x, a.w)];
        const float tex = __fsqrt_rn((float)v);                                                // :386 (repaired)
        const float t = flat ? 0.f : __fmaf_rn(tex, na, nb);
        s_sig[i] = __fsub_rn(smax, __fmul_rn(t, sspan));                                       // :423
    }
    __syncthreads();
    for (int i = threadIdx.x; i < DT_TW * VH; i += DT_NT) {                                    // :431 row pass
        const int y = i / DT_TW, x = i - y * DT_TW;
        const float* p = s_sig + y * VW + x;
        float acc = __fmul_rn(p[0], a.gf[0]);
#pragma unroll
        for (int k = 1; k < 5; ++k) acc = __fadd_rn(acc, __fmul_rn(p[k], a.gf[k]));
        s_row[i] = acc;
    }
    __syncthreads();
    int lo = 0x7f800000, hi = 0;
    for (int i = threadIdx.x; i < DT_TW * DT_TH; i += DT_NT) {                                 // column pass
        const int y = i / DT_TW, x = i - y * DT_TW;
        if (x0 + x >= a.w || y0 + y >= a.h) continue;
        const float* p = s_row + y * DT_TW + x;
        float acc = __fmul_rn(p[0], a.gf[0]);
#pragma unroll
        for (int k = 1; k < 5; ++k) acc = __fadd_rn(acc, __fmul_rn(p[k * DT_TW], a.gf[k]));
        *reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(a.sigma) + (size_t)(y0 + y) * a.spitch + (size_t)(x0 + x) * 4) = acc;
        const int bits = __float_as_int(acc);
        lo = min(lo, bits); hi = max(hi, bits);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if ((threadIdx.x & 31) == 0) { atomicMin(&s_mm[0], lo); atomicMax(&s_mm[1], hi); }
    __syncthreads();
    if (threadIdx.x == 0) { atomicMin(&a.st->sig_min, s_mm[0]); atomicMax(&a.st->sig_max, s_mm[1]); }
}

// one thread per level: sigma_i = min + (max - min) * i / (L - 1) in float32 (:465), then the Q0.8 kernel of
// cv::GaussianBlur(u8) for that sigma exactly as vp_host::gaussian_kernel_q8 builds it (double exp, outside-in
// error diffusion, centre = 256 - 2 * sides)
__global__ void k_sigma_levels(const SigmaStats* __restrict__ st, int ksize, SigmaLevels* __restrict__ out) {
    const int i = threadIdx.x;
    if (i >= VS_LEVELS) return;
    const float mn = __int_as_float(st->sig_min), mx = __int_as_float(st->sig_max);
    const float span = __fsub_rn(mx, mn);
    if (i == 0) { out->mn = mn; out->span = span; }
    const float sigma = __fadd_rn(mn, __fdiv_rn(__fmul_rn(span, (float)i), (float)(VS_LEVELS - 1)));
    const double sx = sigma > 0.f ? (double)sigma : ((ksize - 1) * 0.5 - 1) * 0.3 + 0.8;
    const double scale2x = -0.5 / (sx * sx);
    double g[7], sum = 0.0;
    for (int k = 0; k < ksize; ++k) {
        const double x = (double)k - (ksize - 1) * 0.5;
        g[k] = exp(scale2x * x * x);
        sum += g[k];
    }
    const int n = ksize / 2;
    double err = 0.0;
    int side = 0;
    for (int k = 0; k < 7; ++k) out->q[i][k] = 0;
    for (int k = 0; k < n; ++k) {
        const double v = 256.0 * (g[k] / sum) + err;
        const int qi = (int)rint(v);
        err = v - qi;
        out->q[i][k] = out->q[i][ksize - 1 - k] = qi;
        side += qi;
    }
    out->q[i][n] = 256 - 2 * side;
}

struct VarSharpArgs {
    const uint8_t* src; size_t spitch; int w, h;
    const float* emask; size_t epitch;       // createEdgeMask's output
    const float* sigma; size_t gpitch;       // k_sigma_map's output
    const SigmaLevels* lv;
    int gk;                                  // Gaussian kernel size 3 / 5 / 7
    float strength, edge_strength, smooth_strength;
    int preserve_tone;
    uint8_t* dst; size_t dpitch;
    unsigned long long* sums;                // optional: sum x / sum x^2 of the output (6 x u64)
    SelOut fin;
};

constexpr int VS_TW = 64, VS_TH = 32, VS_NT = 256;   // tile, threads

// The two Gaussian levels of a pixel are exact integer sums sum(p * q_a * q_b) < 2^24, so they are accumulated in float32
// (FFMA is full rate, IMAD is not) as packed pairs: (b, g) x w_low, (b, g) x w_high and r x (w_low, w_high) - three FFMA2 per
// tap, fed by 128-bit pixel loads and 64-bit weight-pair loads.  The kernel is bound by shared-memory wavefronts, so a thread
// owns four VERTICALLY adjacent pixels (lanes stay adjacent in x: conflict-free 128-bit loads; two or four horizontally
// adjacent pixels per thread put the lanes 32 / 64 bytes apart and cost 2x / 4x the wavefronts): a tile column costs GK+3
// pixel loads and - when the four pixels share their level pair, the usual case on the 5x5-blurred sigma map - GK weight
// loads for four outputs.  The sums are exact, so the tap order (column-major here) does not matter.
constexpr int VS_PX = 4;
template <int GK, bool SAME>
__device__ __forceinline__ void varsharp_cols(const float4* __restrict__ tile, int EW, const float2* const (&wp)[VS_PX],
                                              unsigned long long (&bg_lo)[VS_PX], unsigned long long (&bg_hi)[VS_PX],
                                              unsigned long long (&r_lh)[VS_PX]) {
#pragma unroll
    for (int tb = 0; tb < GK; ++tb) {
        unsigned long long bg[GK + VS_PX - 1], rr[GK + VS_PX - 1];
#pragma unroll
        for (int r = 0; r < GK + VS_PX - 1; ++r) {
            const float4 p = tile[r * EW + tb];
            bg[r] = pack2(__float_as_uint(p.x), __float_as_uint(p.y));
            rr[r] = pack2(__float_as_uint(p.z), __float_as_uint(p.z));
        }
#pragma unroll
        for (int ta = 0; ta < GK; ++ta) {
#pragma unroll
            for (int k = 0; k < VS_PX; ++k) {
                const float2 w = wp[SAME ? 0 : k][ta * GK + tb];          // SAME: one load, hoisted out of the k loop
                bg_lo[k] = ffma2(bg[ta + k], pack2(__float_as_uint(w.x), __float_as_uint(w.x)), bg_lo[k]);
                bg_hi[k] = ffma2(bg[ta + k], pack2(__float_as_uint(w.y), __float_as_uint(w.y)), bg_hi[k]);
                r_lh[k] = ffma2(pack2(__float_as_uint(w.x), __float_as_uint(w.y)), rr[ta + k], r_lh[k]);
            }
        }
    }
}

template <int GK>
__global__ void __launch_bounds__(VS_NT, 3) k_varsharp(const VarSharpArgs a) {
    constexpr int GR = GK / 2, EW = VS_TW + 2 * GR, EH = VS_TH + 2 * GR;
    static_assert(VS_TH % VS_PX == 0, "pixel groups tile the column");
    __shared__ float4 s_px[EW * EH];                     // (b, g, r, packed BGR0 bits), reflect-101 padded
    __shared__ float2 s_wp[VS_LEVELS - 1][GK * GK];      // (q[l][a] q[l][b], q[l+1][a] q[l+1][b]) for adjacent levels
    __shared__ unsigned long long s_part[(VS_NT / 32) * 6];
    __shared__ float s_lev[VS_LEVELS];                   // the sigma of each level, :470-473
    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * VS_TW, y0 = blockIdx.y * VS_TH;
    // the per-pixel sigma and edge-mask values of this thread's pixel groups: requested first, so their latency hides behind
    // the tile staging instead of stalling the first use after the barrier
    constexpr int NG = VS_TW * (VS_TH / VS_PX) / VS_NT;
    static_assert(VS_TW * (VS_TH / VS_PX) % VS_NT == 0, "whole pixel groups per thread");
    float sg[VS_PX], em[VS_PX];
    auto fetch = [&](int g, float (&s4)[VS_PX], float (&e4)[VS_PX]) {
        const int i4 = tid + g * VS_NT;
        const int gx = min(x0 + i4 % VS_TW, a.w - 1), yb = y0 + VS_PX * (i4 / VS_TW);
#pragma unroll
        for (int k = 0; k < VS_PX; ++k) {
            const int gyk = min(yb + k, a.h - 1);
            s4[k] = __ldg(reinterpret_cast<const float*>(reinterpret_cast<const uint8_t*>(a.sigma) + (size_t)gyk * a.gpitch + (size_t)gx * 4));
            e4[k] = __ldg(reinterpret_cast<const float*>(reinterpret_cast<const uint8_t*>(a.emask) + (size_t)gyk * a.epitch + (size_t)gx * 4));
        }
    };
    fetch(0, sg, em);
#pragma unroll 4
    for (int i = tid; i < EW * EH; i += VS_NT) {
        const int ty = i / EW, tx = i - ty * EW;
        const int sx = reflect101(x0 - GR + tx, a.w), sy = reflect101(y0 - GR + ty, a.h);
        const uint8_t* p = a.src + (size_t)sy * a.spitch + (size_t)sx * 3;
        const uint32_t b = __ldg(p), g = __ldg(p + 1), r = __ldg(p + 2);
        s_px[i] = make_float4((float)b, (float)g, (float)r, __uint_as_float(b | (g << 8) | (r << 16)));
    }
    for (int i = tid; i < (VS_LEVELS - 1) * GK * GK; i += VS_NT) {
        const int l = i / (GK * GK), t = i - l * (GK * GK), ta = t / GK, tb = t - ta * GK;
        s_wp[l][t] = make_float2((float)(a.lv->q[l][ta] * a.lv->q[l][tb]), (float)(a.lv->q[l + 1][ta] * a.lv->q[l + 1][tb]));
    }
    const float mn = a.lv->mn, span = a.lv->span;
    if (tid < VS_LEVELS) s_lev[tid] = __fadd_rn(mn, __fdiv_rn(__fmul_rn(span, (float)tid), (float)(VS_LEVELS - 1)));
    __syncthreads();
    unsigned acc6[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll 1
    for (int g = 0; g < NG; ++g) {              // not unrolled: the body is several thousand instructions (instruction cache)
        const int i4 = tid + g * VS_NT;
        const int yb = VS_PX * (i4 / VS_TW), x = i4 % VS_TW;
        const int gx = x0 + x;
        float sg_cur[VS_PX], em_cur[VS_PX];
#pragma unroll
        for (int k = 0; k < VS_PX; ++k) { sg_cur[k] = sg[k]; em_cur[k] = em[k]; }
        if (g + 1 < NG) fetch(g + 1, sg, em);   // the next group's values travel while this one is computed
        if (gx >= a.w || y0 + yb >= a.h) continue;
        int idx2[VS_PX]; float alpha2[VS_PX];
#pragma unroll
        for (int k = 0; k < VS_PX; ++k) {
            const float sigma = sg_cur[k];
            // :476-486  level pair and interpolation factor, float32 throughout
            int idx = 0;
            if (span > 0.f) idx = (int)__fmul_rn(__fdiv_rn(__fsub_rn(sigma, mn), span), (float)(VS_LEVELS - 1));
            idx = max(0, min(idx, VS_LEVELS - 2));
            const float slo = s_lev[idx], shi = s_lev[idx + 1];
            const float den = __fsub_rn(shi, slo);
            idx2[k] = idx;
            alpha2[k] = den > 0.f ? __fdiv_rn(__fsub_rn(sigma, slo), den) : 0.f;
        }
        // the two Gaussian levels at the four pixels: exact Q8.16 sums, (v + 2^15) >> 16
        unsigned long long bg_lo2[VS_PX], bg_hi2[VS_PX], r_lh2[VS_PX];
        const float2* wp[VS_PX];
        bool same = true;
#pragma unroll
        for (int k = 0; k < VS_PX; ++k) {
            bg_lo2[k] = bg_hi2[k] = r_lh2[k] = 0ull;
            wp[k] = s_wp[idx2[k]];
            same = same && idx2[k] == idx2[0];
        }
        if (same) varsharp_cols<GK, true>(s_px + yb * EW + x, EW, wp, bg_lo2, bg_hi2, r_lh2);
        else varsharp_cols<GK, false>(s_px + yb * EW + x, EW, wp, bg_lo2, bg_hi2, r_lh2);
#pragma unroll
      for (int k = 0; k < VS_PX; ++k) {
        const int y = yb + k, gy = y0 + y;
        if (gy >= a.h) continue;
        const float alpha = alpha2[k];
        float fl[3], fh[3];
        unpack2(bg_lo2[k], fl[0], fl[1]); unpack2(bg_hi2[k], fh[0], fh[1]); unpack2(r_lh2[k], fl[2], fh[2]);
        unsigned lo[3], hi[3];
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) { lo[ch] = __float2uint_rz(fl[ch]) + (1u << 15); hi[ch] = __float2uint_rz(fh[ch]) + (1u << 15); }
        const uint32_t cpx = __float_as_uint(s_px[(y + GR) * EW + x + GR].w);
        const int u[3] = {(int)(cpx & 0xff), (int)((cpx >> 8) & 0xff), (int)((cpx >> 16) & 0xff)};
        const float e = em_cur[k];
        const float t1 = __fmul_rn(e, a.edge_strength);
        const float t2 = __fmul_rn(__fsub_rn(1.0f, e), a.smooth_strength);
        const float st = __fmul_rn(a.strength, __fadd_rn(t1, t2));                              // :538-541
        const float om = __fsub_rn(1.0f, alpha);
        int o[3];
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
            // :496-498  static_cast<uchar>(low * (1 - alpha) + high * alpha): truncation
            const float bl = __fadd_rn(__fmul_rn((float)(lo[ch] >> 16), om), __fmul_rn((float)(hi[ch] >> 16), alpha));
            const int blur = min(max(__float2int_rz(bl), 0), 255);
            const int um = max(u[ch] - blur, 0);                                                // :505 cv::subtract saturates
            o[ch] = (int)f2u8(__fadd_rn((float)u[ch], __fmul_rn(st, (float)um)));              // :552-554
        }
        if (a.preserve_tone) {                                                                  // :563-584, as in k_upsharp
            const int yi = (4899 * u[2] + 9617 * u[1] + 1868 * u[0] + (1 << 13)) >> 14;
            const int cr = sat_u8(((u[2] - yi) * 11682 + (128 << 14) + (1 << 13)) >> 14) - 128;
            const int cb = sat_u8(((u[0] - yi) * 9241 + (128 << 14) + (1 << 13)) >> 14) - 128;
            const int yo = (4899 * o[2] + 9617 * o[1] + 1868 * o[0] + (1 << 13)) >> 14;
            o[0] = sat_u8(yo + ((cb * 29049 + (1 << 13)) >> 14));
            o[1] = sat_u8(yo + ((cb * -5636 + cr * -11698 + (1 << 13)) >> 14));
            o[2] = sat_u8(yo + ((cr * 22987 + (1 << 13)) >> 14));
        }
        uint8_t* q = a.dst + (size_t)gy * a.dpitch + (size_t)gx * 3;
        q[0] = (uint8_t)o[0]; q[1] = (uint8_t)o[1]; q[2] = (uint8_t)o[2];
        if (a.sums) {
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) { acc6[ch] += o[ch]; acc6[3 + ch] += o[ch] * o[ch]; }
        }
      }
    }
    if (a.sums) {
        warp_partials6(acc6, s_part);
        __syncthreads();
        if (tid < 32) warp_flush_and_finalize(s_part, VS_NT / 32, a.sums, a.fin, gridDim.x * gridDim.y);
    }
}

}  // namespace vp
