// vp_upsharp.cuh - fused  cv::resize(INTER_CUBIC)  ->  AdaptiveSharpening (createEdgeMask +
// applyUnsharpMask + tone preservation)  for CV_8UC3, one dst tile per CTA, HBM in -> HBM out.
//
// Reference call sites: cv::resize src/upscaler.cpp:81, src/test_enhancements.cpp:348;
// AdaptiveSharpening::createEdgeMask src/adaptive_sharpening.cpp:107-230; applyUnsharpMask :232-355.
// Arithmetic follows oracle/restate.py (pinned against OpenCV 4.13 through cv2):
//   resize   : int32 horizontal pass with 11-bit short taps, float32 vertical pass
//              t = fma(S0,b0,S1*b1) + fma(S2,b2,S3*b3), dst = sat(RNE(t))           -> bit-exact
//   gray     : (9798 R + 19235 G + 3735 B + 2^14) >> 15
//   edges    : e1 = RNE(fma(|Sx|,.5,|Sy|*.5)); comb = RNE(fma(e1,.6f,|Lap|*.4f)); sigmoid LUT[comb]
//   mask     : separable float32 5x5 sigma 1.5 Gaussian of the sigmoid image (reflect-101)
//   unsharp  : 8-bit Gaussian in Q8 fixed point (row Q8.8, column Q8.16, (v+2^15)>>16);
//              um = sat(up - blur); out = sat(RNE(up + st*(e*es + (1-e)*ss)*um))
//   tone     : Y from `out`, Cr/Cb from `up` (14-bit fixed-point YCrCb), back to BGR.
// All stencil borders are BORDER_REFLECT_101 of the *destination* image: the tile's halo-3 "ext"
// region is the reflect-padded upscaled image, and |Sobel|, Laplacian and the Gaussians are mirror
// symmetric, so evaluating them on the padded tile reproduces OpenCV's per-stage padding.
#pragma once
#include "vp_common.cuh"

namespace vp {

constexpr int US_TW = 64, US_TH = 32, US_NT = 256, US_HALO = 3;

struct ResizeTables {            // device memory, built on the host exactly as cv::resize does
    const int* xofs;             // [dst_w]  first tap column (s-1), unclamped
    const short4* xcoef;         // [dst_w]  11-bit taps
    const int* yofs;             // [dst_h]
    const float4* ycoef;         // [dst_h]  beta * 2^-22 as float32
};

struct UpsharpArgs {
    const uint8_t* src; size_t spitch; int sw, sh;
    uint8_t* dst; size_t dpitch; int dw, dh;        // dst may be null (mask-only)
    ResizeTables rt;
    int do_resize;               // 0: src is already dst-sized (standalone AdaptiveSharpening)
    int do_sharpen;              // 0: bicubic only
    int max_src_rows, max_src_cols;   // per-tile source window bounds (host computed)
    // sharpening parameters
    float strength, edge_strength, smooth_strength;
    int preserve_tone;
    int gk;                      // 8-bit Gaussian kernel size (3,5,7)
    int gq[7];                   // Q0.8 taps
    float gf[5];                 // float32 5x5 sigma 1.5 taps of the mask blur
    const float* sig_lut;        // 256-entry sigmoid table
    float* mask_out; size_t mask_out_pitch;          // optional: write the edge mask
    const float* mask_in; size_t mask_in_pitch;      // optional: use this mask instead of computing it
    unsigned long long* sums;    // optional: accumulate sum x / sum x^2 of the output (6 x u64)
};

struct UpsharpSmem {             // byte offsets into dynamic shared memory
    int raw, raw_pitch, hs, up, sig, rowf, tabs, total;
};

__host__ __device__ inline UpsharpSmem upsharp_layout(int halo, int max_src_rows, int max_src_cols, int do_sharpen) {
    UpsharpSmem L;
    const int EW = US_TW + 2 * halo, EH = US_TH + 2 * halo;
    int o = 0;
    L.raw_pitch = ((max_src_cols * 3 + 15 + 15) & ~15) + 16;
    L.raw = o; o += L.raw_pitch * max_src_rows;
    // hs: H-pass rows as float32 (max_src_rows x EW x 3); later reused for the Q8.8 row pass
    // (3 planes x (US_TH+6) x US_TW u16) - take the larger
    int hs = max_src_rows * EW * 3 * 4;
    int rowq = do_sharpen ? 3 * (US_TH + 6) * US_TW * 2 : 0;
    L.hs = o; o += ((hs > rowq ? hs : rowq) + 15) & ~15;
    L.up = o; o += EW * EH * 4;
    // sig (68 x 36 f32) later reused for the staged output tile (US_TH x US_TW*3 bytes)
    int sig = do_sharpen ? (US_TW + 4) * (US_TH + 4) * 4 : 0;
    int outb = US_TH * US_TW * 3;
    L.sig = o; o += ((sig > outb ? sig : outb) + 15) & ~15;
    L.rowf = o; o += do_sharpen ? US_TW * (US_TH + 4) * 4 : 0;
    // tables: xofs[EW] int, xcoef[EW] short4, yofs[EH] int, ycoef[EH] float4, lut[256] float, mbarrier
    L.tabs = o; o += EW * 4 + EW * 8 + EH * 4 + 4 /*pad*/ + EH * 16 + 256 * 4 + 16 + 64;
    L.total = o;
    return L;
}

__global__ void __launch_bounds__(US_NT) k_upsharp(const UpsharpArgs a) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ unsigned long long s_red[6];
    const int tid = threadIdx.x;
    const int halo = a.do_sharpen ? US_HALO : 0;
    const int EW = US_TW + 2 * halo, EH = US_TH + 2 * halo;
    const UpsharpSmem L = upsharp_layout(halo, a.max_src_rows, a.max_src_cols, a.do_sharpen);
    uint8_t* s_raw = smem + L.raw;
    float* s_hs = reinterpret_cast<float*>(smem + L.hs);
    uint16_t* s_rowq = reinterpret_cast<uint16_t*>(smem + L.hs);
    uint32_t* s_up = reinterpret_cast<uint32_t*>(smem + L.up);
    float* s_sig = reinterpret_cast<float*>(smem + L.sig);
    uint8_t* s_out = smem + L.sig;
    float* s_rowf = reinterpret_cast<float*>(smem + L.rowf);
    uint8_t* tp = smem + L.tabs;
    float4* s_yc = reinterpret_cast<float4*>(tp); tp += EH * 16;
    short4* s_xc = reinterpret_cast<short4*>(tp); tp += EW * 8;
    uint64_t* bar = reinterpret_cast<uint64_t*>(tp); tp += 16;
    int* s_xo = reinterpret_cast<int*>(tp); tp += EW * 4;
    int* s_yo = reinterpret_cast<int*>(tp); tp += EH * 4;
    float* s_lut = reinterpret_cast<float*>(tp);

    const int x0 = blockIdx.x * US_TW, y0 = blockIdx.y * US_TH;
    if (tid < 6) s_red[tid] = 0ull;

    // ---- tables for this tile + source window -------------------------------------------------
    const int dxa = max(0, x0 - halo), dxb = min(a.dw - 1, x0 + US_TW - 1 + halo);
    const int dya = max(0, y0 - halo), dyb = min(a.dh - 1, y0 + US_TH - 1 + halo);
    int sx0, sx1, sy0, sy1;   // inclusive source window
    if (a.do_resize) {
        sx0 = clampi(a.rt.xofs[dxa], 0, a.sw - 1); sx1 = clampi(a.rt.xofs[dxb] + 3, 0, a.sw - 1);
        sy0 = clampi(a.rt.yofs[dya], 0, a.sh - 1); sy1 = clampi(a.rt.yofs[dyb] + 3, 0, a.sh - 1);
        for (int i = tid; i < EW; i += US_NT) {
            const int dx = clampi(reflect101(x0 - halo + i, a.dw), dxa, dxb);   // clamp: unused columns of partial tiles
            s_xo[i] = a.rt.xofs[dx]; s_xc[i] = a.rt.xcoef[dx];
        }
        for (int i = tid; i < EH; i += US_NT) {
            const int dy = clampi(reflect101(y0 - halo + i, a.dh), dya, dyb);
            s_yo[i] = a.rt.yofs[dy]; s_yc[i] = a.rt.ycoef[dy];
        }
    } else {
        sx0 = dxa; sx1 = dxb; sy0 = dya; sy1 = dyb;
    }
    if (a.do_sharpen && tid < 256) s_lut[tid] = a.sig_lut[tid];
    const int bx0 = (sx0 * 3) & ~15, bx1 = (sx1 + 1) * 3;
    const int RP = L.raw_pitch;
    stage_rows(s_raw, RP, a.src, a.spitch, 3 * a.sw, sy0, sy1 + 1, bx0, bx1, aligned16(a.src, a.spitch), bar);

    if (a.do_resize) {
        // ---- P1: horizontal pass, int32 exact, stored as float32 (|H| < 2^24) ----------------
        const int nrows = sy1 - sy0 + 1;
        for (int i = tid; i < nrows * EW; i += US_NT) {
            const int r = i / EW, ex = i - r * EW;
            const int xo = s_xo[ex];
            const short4 c = s_xc[ex];
            const uint8_t* row = s_raw + r * RP - bx0;
            const uint8_t* p0 = row + 3 * clampi(xo, 0, a.sw - 1);
            const uint8_t* p1 = row + 3 * clampi(xo + 1, 0, a.sw - 1);
            const uint8_t* p2 = row + 3 * clampi(xo + 2, 0, a.sw - 1);
            const uint8_t* p3 = row + 3 * clampi(xo + 3, 0, a.sw - 1);
            float* h = s_hs + (size_t)i * 3;
#pragma unroll
            for (int ch = 0; ch < 3; ++ch)
                h[ch] = (float)((int)p0[ch] * c.x + (int)p1[ch] * c.y + (int)p2[ch] * c.z + (int)p3[ch] * c.w);
        }
        __syncthreads();
        // ---- P2: vertical pass, float32, RNE + saturate; gray packed into byte 3 -----------------
        for (int i = tid; i < EH * EW; i += US_NT) {
            const int ey = i / EW, ex = i - ey * EW;
            const int yo = s_yo[ey];
            const float4 b = s_yc[ey];
            const float* h0 = s_hs + ((clampi(yo, 0, a.sh - 1) - sy0) * EW + ex) * 3;
            const float* h1 = s_hs + ((clampi(yo + 1, 0, a.sh - 1) - sy0) * EW + ex) * 3;
            const float* h2 = s_hs + ((clampi(yo + 2, 0, a.sh - 1) - sy0) * EW + ex) * 3;
            const float* h3 = s_hs + ((clampi(yo + 3, 0, a.sh - 1) - sy0) * EW + ex) * 3;
            int v[3];
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
                const float t = __fadd_rn(__fmaf_rn(h0[ch], b.x, __fmul_rn(h1[ch], b.y)),
                                          __fmaf_rn(h2[ch], b.z, __fmul_rn(h3[ch], b.w)));
                v[ch] = rne_sat_u8(t);
            }
            const int gray = (9798 * v[2] + 19235 * v[1] + 3735 * v[0] + (1 << 14)) >> 15;
            s_up[i] = (uint32_t)v[0] | ((uint32_t)v[1] << 8) | ((uint32_t)v[2] << 16) | ((uint32_t)gray << 24);
        }
    } else {
        for (int i = tid; i < EH * EW; i += US_NT) {
            const int ey = i / EW, ex = i - ey * EW;
            const int sx = reflect101(x0 - halo + ex, a.dw), sy = reflect101(y0 - halo + ey, a.dh);
            uint32_t px = 0;
            if (sx >= sx0 && sx <= sx1 && sy >= sy0 && sy <= sy1) {
                const uint8_t* p = s_raw + (sy - sy0) * RP + (sx * 3 - bx0);
                const int v0 = p[0], v1 = p[1], v2 = p[2];
                const int gray = (9798 * v2 + 19235 * v1 + 3735 * v0 + (1 << 14)) >> 15;
                px = (uint32_t)v0 | ((uint32_t)v1 << 8) | ((uint32_t)v2 << 16) | ((uint32_t)gray << 24);
            }
            s_up[i] = px;
        }
    }
    __syncthreads();

    const int vw = min(US_TW, a.dw - x0), vh = min(US_TH, a.dh - y0);   // valid part of the tile

    if (!a.do_sharpen) {
        unsigned acc6[6] = {0, 0, 0, 0, 0, 0};
        for (int i = tid; i < US_TH * US_TW; i += US_NT) {
            const int y = i / US_TW, x = i - y * US_TW;
            const uint32_t px = s_up[y * EW + x];
            uint8_t* o = s_out + y * (US_TW * 3) + x * 3;
            const unsigned v0 = px & 0xff, v1 = (px >> 8) & 0xff, v2 = (px >> 16) & 0xff;
            o[0] = v0; o[1] = v1; o[2] = v2;
            if (x < vw && y < vh) {
                acc6[0] += v0; acc6[1] += v1; acc6[2] += v2;
                acc6[3] += v0 * v0; acc6[4] += v1 * v1; acc6[5] += v2 * v2;
            }
        }
        if (a.sums) reduce_sums6(acc6, a.sums, s_red);
    } else {
        const bool own_mask = (a.mask_in == nullptr);
        constexpr int CW = US_TW + 4, CH = US_TH + 4;
        if (own_mask) {
            // ---- P3: Sobel / Laplacian magnitude -> sigmoid, on the (TW+4)x(TH+4) region ----------
            const uint8_t* g8 = reinterpret_cast<const uint8_t*>(s_up) + 3;   // gray byte of each word
            for (int i = tid; i < CW * CH; i += US_NT) {
                const int cy = i / CW, cx = i - cy * CW;
                const uint8_t* g = g8 + ((cy * EW) + cx) * 4;    // top-left of the 3x3 window (ext cx, cy)
                const int g00 = g[0], g01 = g[4], g02 = g[8];
                const int g10 = g[EW * 4], g11 = g[EW * 4 + 4], g12 = g[EW * 4 + 8];
                const int g20 = g[EW * 8], g21 = g[EW * 8 + 4], g22 = g[EW * 8 + 8];
                const int sxv = (g02 + 2 * g12 + g22) - (g00 + 2 * g10 + g20);
                const int syv = (g20 + 2 * g21 + g22) - (g00 + 2 * g01 + g02);
                const int lap = 2 * (g00 + g02 + g20 + g22) - 8 * g11;
                const float ax = (float)min(abs(sxv), 255), ay = (float)min(abs(syv), 255), al = (float)min(abs(lap), 255);
                const float e1 = (float)rne_sat_u8(__fmaf_rn(ax, 0.5f, __fmul_rn(ay, 0.5f)));
                const int comb = rne_sat_u8(__fmaf_rn(e1, 0.6f, __fmul_rn(al, 0.4f)));
                s_sig[i] = s_lut[comb];
            }
            __syncthreads();
            // ---- P4a: float32 row pass of the mask blur ------------------------------------------
            for (int i = tid; i < US_TW * CH; i += US_NT) {
                const int cy = i / US_TW, x = i - cy * US_TW;
                const float* p = s_sig + cy * CW + x;
                float acc = __fmul_rn(p[0], a.gf[0]);
#pragma unroll
                for (int k = 1; k < 5; ++k) acc = __fadd_rn(acc, __fmul_rn(p[k], a.gf[k]));
                s_rowf[i] = acc;
            }
        }
        // ---- P4b: Q8.8 row pass of the 8-bit Gaussian (3 planes) ------------------------------
        {
            const int gr = a.gk >> 1;
            const int nrow = US_TH + 2 * gr;
            for (int i = tid; i < US_TW * nrow; i += US_NT) {
                const int ry = i / US_TW, x = i - ry * US_TW;
                const uint32_t* p = s_up + (ry + US_HALO - gr) * EW + (x + US_HALO - gr);
                int sb = 0, sg = 0, sr = 0;
                for (int k = 0; k < a.gk; ++k) {
                    const uint32_t px = p[k];
                    const int q = a.gq[k];
                    sb += (int)(px & 0xff) * q; sg += (int)((px >> 8) & 0xff) * q; sr += (int)((px >> 16) & 0xff) * q;
                }
                s_rowq[i] = (uint16_t)sb;
                s_rowq[(US_TH + 6) * US_TW + i] = (uint16_t)sg;
                s_rowq[2 * (US_TH + 6) * US_TW + i] = (uint16_t)sr;
            }
        }
        __syncthreads();
        // ---- P5: column passes, unsharp blend, tone preservation ------------------------------
        const int gr = a.gk >> 1;
        unsigned acc6[6] = {0, 0, 0, 0, 0, 0};
        for (int i = tid; i < US_TH * US_TW; i += US_NT) {
            const int y = i / US_TW, x = i - y * US_TW;
            float e;
            if (own_mask) {
                const float* p = s_rowf + y * US_TW + x;
                e = __fmul_rn(p[0], a.gf[0]);
#pragma unroll
                for (int k = 1; k < 5; ++k) e = __fadd_rn(e, __fmul_rn(p[k * US_TW], a.gf[k]));
            } else {
                const int gx = min(x0 + x, a.dw - 1), gy = min(y0 + y, a.dh - 1);
                e = *reinterpret_cast<const float*>(reinterpret_cast<const uint8_t*>(a.mask_in) + (size_t)gy * a.mask_in_pitch + (size_t)gx * 4);
            }
            if (a.mask_out && x < vw && y < vh)
                *reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(a.mask_out) + (size_t)(y0 + y) * a.mask_out_pitch + (size_t)(x0 + x) * 4) = e;
            int bl[3];
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
                const uint16_t* q = s_rowq + ch * (US_TH + 6) * US_TW + y * US_TW + x;
                int s = 0;
                for (int k = 0; k < a.gk; ++k) s += (int)q[k * US_TW] * a.gq[k];
                bl[ch] = (s + (1 << 15)) >> 16;
            }
            const uint32_t px = s_up[(y + US_HALO) * EW + x + US_HALO];
            const int u[3] = {(int)(px & 0xff), (int)((px >> 8) & 0xff), (int)((px >> 16) & 0xff)};
            const float t1 = __fmul_rn(e, a.edge_strength);
            const float t2 = __fmul_rn(__fsub_rn(1.0f, e), a.smooth_strength);
            const float st = __fmul_rn(a.strength, __fadd_rn(t1, t2));
            int o[3];
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
                const int um = max(u[ch] - bl[ch], 0);
                o[ch] = rne_sat_u8(__fadd_rn((float)u[ch], __fmul_rn(st, (float)um)));
            }
            if (a.preserve_tone) {
                const int yi = (4899 * u[2] + 9617 * u[1] + 1868 * u[0] + (1 << 13)) >> 14;
                const int cr = sat_u8(((u[2] - yi) * 11682 + (128 << 14) + (1 << 13)) >> 14) - 128;
                const int cb = sat_u8(((u[0] - yi) * 9241 + (128 << 14) + (1 << 13)) >> 14) - 128;
                const int yo = (4899 * o[2] + 9617 * o[1] + 1868 * o[0] + (1 << 13)) >> 14;
                o[0] = sat_u8(yo + ((cb * 29049 + (1 << 13)) >> 14));
                o[1] = sat_u8(yo + ((cb * -5636 + cr * -11698 + (1 << 13)) >> 14));
                o[2] = sat_u8(yo + ((cr * 22987 + (1 << 13)) >> 14));
            }
            uint8_t* ob = s_out + y * (US_TW * 3) + x * 3;
            ob[0] = (uint8_t)o[0]; ob[1] = (uint8_t)o[1]; ob[2] = (uint8_t)o[2];
            if (x < vw && y < vh) {
#pragma unroll
                for (int ch = 0; ch < 3; ++ch) { acc6[ch] += o[ch]; acc6[3 + ch] += o[ch] * o[ch]; }
            }
        }
        if (a.sums) reduce_sums6(acc6, a.sums, s_red);
        (void)gr;
    }
    __syncthreads();
    if (a.sums && tid < 6 && s_red[tid]) atomicAdd(&a.sums[tid], s_red[tid]);

    // ---- store the staged tile: 16-byte vectors when aligned, bytes otherwise ---------------------
    if (a.dst) {
        const int rb = vw * 3;                               // valid bytes per tile row
        const bool al = aligned16(a.dst, a.dpitch);          // x0*3 = 192*blockIdx.x is a multiple of 16
        const int nv = al ? rb >> 4 : 0;
        for (int i = tid; i < nv * vh; i += US_NT) {
            const int y = i / nv, q = i - y * nv;
            *reinterpret_cast<uint4*>(a.dst + (size_t)(y0 + y) * a.dpitch + (size_t)x0 * 3 + q * 16) =
                *reinterpret_cast<const uint4*>(s_out + y * (US_TW * 3) + q * 16);
        }
        const int tb = rb - nv * 16;
        for (int i = tid; i < tb * vh; i += US_NT) {
            const int y = i / tb, q = nv * 16 + (i - y * tb);
            a.dst[(size_t)(y0 + y) * a.dpitch + (size_t)x0 * 3 + q] = s_out[y * (US_TW * 3) + q];
        }
    }
}

}  // namespace vp
