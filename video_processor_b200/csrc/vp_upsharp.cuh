// vp_upsharp.cuh - fused  cv::resize(INTER_CUBIC)  ->  AdaptiveSharpening (createEdgeMask +
// applyUnsharpMask + tone preservation)  for CV_8UC3, one dst tile per CTA, HBM in -> HBM out.
//
// Reference call sites: cv::resize src/upscaler.cpp:81, src/test_enhancements.cpp:348;
// AdaptiveSharpening::createEdgeMask src/adaptive_sharpening.cpp:107-230; applyUnsharpMask :232-355.
// Arithmetic follows oracle/restate.py (pinned against OpenCV 4.13 through cv2):
//   resize   : int32 horizontal pass with 11-bit short taps, float32 vertical pass
//              t = fma(S0,b0,S1*b1) + fma(S2,b2,S3*b3), dst = sat(RNE(t))           -> bit-exact
//   gray     : (9798 R + 19235 G + 3735 B + 2^14) >> 15
//   edges    : e1 = RNE((|Sx|+|Sy|)/2); comb = RNE(fma(e1,.6f,|Lap|*.4f)) == (6 e1 + 4 |Lap| + 5)/10
//              (no .5 ties exist, tests/test_host_tables.py checks all 65,536 pairs); sigmoid LUT[comb]
//   mask     : separable float32 5x5 sigma 1.5 Gaussian of the sigmoid image (reflect-101)
//   unsharp  : 8-bit Gaussian in Q8 fixed point (row Q8.8, column Q8.16, (v+2^15)>>16);
//              um = sat(up - blur); out = sat(RNE(up + st*(e*es + (1-e)*ss)*um))
//   tone     : Y from `out`, Cr/Cb from `up` (14-bit fixed-point YCrCb), back to BGR.
// All stencil borders are BORDER_REFLECT_101 of the *destination* image: the tile's halo-3 "ext"
// region is the reflect-padded upscaled image, and |Sobel|, Laplacian and the Gaussians are mirror
// symmetric, so evaluating them on the padded tile reproduces OpenCV's per-stage padding.
//
// Data flow inside a CTA (64x32 dst tile, 256 threads), everything between HBM-in and HBM-out stays
// in shared memory / registers:
//   P0  TMA bulk row copies: source window -> raw bytes; raw -> BGR0 words with the replicate border
//       materialised (s_src is indexed in UNclamped source coordinates, so later passes do no clamping)
//   P1  horizontal cubic pass, two ext columns per thread (5 source pixels feed 8 taps)
//   P2  vertical cubic pass, runs of 4 ext rows per thread with a sliding 4-row register window
//   P3  Sobel/Laplacian -> sigmoid: a thread walks down a column with a 3-row register window
//   P4  row passes of the float mask blur and of the Q8 Gaussian (B,G packed in one 32-bit lane)
//   P5  column passes with sliding windows, unsharp blend, tone preservation, output statistics
//   P6  staged tile -> HBM by TMA bulk stores (16-byte aligned frames) or byte stores
#pragma once
#include "vp_bilateral.cuh"   // u8f / f2u8
#include "vp_common.cuh"
#include "vp_stats.cuh"

namespace vp {

constexpr int US_TW = 64, US_TH = 32, US_NT = 256, US_HALO = 3;
#ifndef VP_US_P5_FLOAT
#define VP_US_P5_FLOAT 1
#endif

struct ResizeTables {            // device memory, built on the host exactly as cv::resize does
    const int* xofs;             // [dst_w]  first tap column (s-1), unclamped
    const int4* xcoef;           // [dst_w]  11-bit taps
    const int* yofs;             // [dst_h]
    const float4* ycoef;         // [dst_h]  beta * 2^-22 as float32
};

struct UpsharpArgs {
    const uint8_t* src; size_t spitch; int sw, sh;
    uint8_t* dst; size_t dpitch; int dw, dh;        // dst may be null (mask-only)
    ResizeTables rt;
    int max_src_rows, max_src_cols;   // per-tile source window bounds in unclamped coordinates (host computed)
    int yperiod;                      // S when dst_h == S * src_h with S = 2 or 4 (host-verified row tables), else 0
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
    SelOut fin;                  // optional: last CTA turns the sums into the POST noise class
    // 16-byte aligned frames: the source window arrives (box raw_pitch - 16 bytes x max_src_rows) and the finished tile
    // leaves (box US_TW * 3 bytes x US_TH) as one tensor-map TMA box each
    int use_tm_src, use_tm_dst;
    // out_bgrx: the tile leaves as one 32-bit word per pixel (B, G, R, 1) - the lane buffer between this kernel and the POST
    // bilateral, whose halo tile then arrives by TMA already in the form its tap loop reads (no expansion pass there);
    // tm_dst is then a tensor of 32-bit units and use_tm_dst is set
    int out_bgrx;
    // src_bgrx: `src` holds 32-bit (B, G, R, 1) pixels (k_bilateral<., DSTX>): the source window arrives by TMA as the
    // BGR0-word tile the horizontal pass reads (box src_pitch_x x max_src_rows of 32-bit units, origin rounded down to 4
    // pixels); only tiles on the image border run a pass that replicates the zero-filled outside positions.  RESIZE only.
    int src_bgrx, src_pitch_x;
    alignas(64) CUtensorMap tm_src;
    alignas(64) CUtensorMap tm_dst;
};

struct UpsharpSmem {             // byte offsets into dynamic shared memory
    int raw, raw_pitch, src, src_pitch, hs, rowbg, rowr, rowf, up, gray, sig, tabs, total;
};

constexpr int US_GP = 80;        // gray plane pitch (bytes)

__host__ __device__ inline UpsharpSmem upsharp_layout(bool resize, bool sharpen, int max_src_rows, int max_src_cols) {
    UpsharpSmem L{};
    const int halo = sharpen ? US_HALO : 0;
    const int EW = US_TW + 2 * halo, EH = US_TH + 2 * halo;
    int o = 0;
    L.raw_pitch = ((max_src_cols * 3 + 15 + 15) & ~15) + 16;
    // region U: raw source bytes + BGR0 source words (dead after P1 / the expand), later the sigmoid
    // image (68 x 36 f32, P3-P4) and finally the staged output tile (US_TH x US_TW*3 bytes, P5-P6)
    L.src_pitch = (max_src_cols + 4) & ~3;      // whole 128-bit words per row (the word-wise expand stores uint4s)
    const int rawb = L.raw_pitch * max_src_rows;
    const int srcb = resize ? ((L.src_pitch * max_src_rows * 4 + 15) & ~15) : 0;
    const int sig = sharpen ? (US_TW + 4) * (US_TH + 4) * 4 : 0;
    const int outb = US_TH * US_TW * 4;         // staged output tile: 3 bytes per pixel, or one word (out_bgrx)
    const int sigout = ((sig > outb ? sig : outb) + 15) & ~15;
    L.raw = o; L.src = o + rawb; L.sig = o;
    o += (rawb + srcb > sigout ? rawb + srcb : sigout);
    // region A: H-pass rows (float x 3 per ext column) during P1/P2, then the row-pass planes of P4
    const int hs = resize ? max_src_rows * EW * 3 * 4 : 0;
    const int rowbg = (US_TH + 6) * US_TW * 4, rowr = (US_TH + 6) * US_TW * 2, rowf = (US_TH + 4) * US_TW * 4;
    const int p4 = sharpen ? rowbg + rowr + rowf : 0;
    L.hs = o; L.rowbg = o; L.rowr = o + rowbg; L.rowf = o + rowbg + rowr;
    o += ((hs > p4 ? hs : p4) + 15) & ~15;
    L.up = o; o += EW * EH * 4;
    L.gray = o; o += sharpen ? ((US_GP * EH + 15) & ~15) : 0;
    // tables: ycoef float4[EH] | xcoef int4[EW] | mbarrier(16) | xofs int[EW] | yofs int[EH] | lut float[256]
    L.tabs = o; o += EH * 16 + EW * 16 + 16 + EW * 4 + EH * 4 + 256 * 4 + 32;
    L.total = o;
    return L;
}

// ---- P2 for an integer vertical ratio S (2 or 4) on tiles whose ext rows need no mirroring -------------------------
// Tile rows start at multiples of US_TH, so ext row ey has phase (ey - HALO) mod S; the vertical tap origin advances
// exactly before a row of phase S/2.  For a run of four ext rows everything below is therefore known at compile time:
// which source rows each output row reads (no sliding-window register moves, no branches), which of the S coefficient
// sets it uses (registers, not shared memory), and which neighbouring rows share a window and can be filtered as one
// packed f32x2 pair.
__host__ __device__ constexpr int p2_phase(int S, int HALO, int j) { return (((j - HALO) % S) + S) % S; }
__host__ __device__ constexpr int p2_off(int S, int HALO, int j) {       // source-row offset of ext row j relative to row 0 of the run
    int o = 0;
    for (int k = 1; k <= j; ++k) o += (p2_phase(S, HALO, k) == S / 2) ? 1 : 0;
    return o;
}

template <bool SHARPEN>
__device__ __forceinline__ void p2_store(uint32_t* s_up, uint8_t* s_gray, int idx_up, int idx_gray, const uint32_t (&v)[3]) {
    s_up[idx_up] = v[0] | (v[1] << 8) | (v[2] << 16);
    if (SHARPEN) s_gray[idx_gray] = (uint8_t)((9798 * v[2] + 19235 * v[1] + 3735 * v[0] + (1 << 14)) >> 15);
}

template <int S, int HALO, bool SHARPEN, int J, int NR>
__device__ __forceinline__ void p2_rows(const float (&W)[NR][3], const float4 (&bj)[S], uint32_t* s_up, uint8_t* s_gray, int ey0, int ex, int eh) {
    constexpr int EW = US_TW + 2 * HALO;
    if constexpr (J < 4) {
        constexpr int o = p2_off(S, HALO, J);
        constexpr bool pair = (J + 1 < 4) && (p2_off(S, HALO, J + 1) == o);
        const int ey = ey0 + J;
        if constexpr (pair) {
            const float4 ba = bj[p2_phase(S, HALO, J)], bb = bj[p2_phase(S, HALO, J + 1)];
            const unsigned long long bx = pack2(__float_as_uint(ba.x), __float_as_uint(bb.x)), by = pack2(__float_as_uint(ba.y), __float_as_uint(bb.y));
            const unsigned long long bz = pack2(__float_as_uint(ba.z), __float_as_uint(bb.z)), bw = pack2(__float_as_uint(ba.w), __float_as_uint(bb.w));
            uint32_t va[3], vb[3];
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
                const unsigned long long p0 = pack2(__float_as_uint(W[o][ch]), __float_as_uint(W[o][ch])), p1 = pack2(__float_as_uint(W[o + 1][ch]), __float_as_uint(W[o + 1][ch]));
                const unsigned long long p2 = pack2(__float_as_uint(W[o + 2][ch]), __float_as_uint(W[o + 2][ch])), p3 = pack2(__float_as_uint(W[o + 3][ch]), __float_as_uint(W[o + 3][ch]));
                const unsigned long long t = fadd2(ffma2(p0, bx, fmul2(p1, by)), ffma2(p2, bz, fmul2(p3, bw)));
                float ta, tb;
                unpack2(t, ta, tb);
                va[ch] = f2u8(ta); vb[ch] = f2u8(tb);
            }
            if (ey < eh) p2_store<SHARPEN>(s_up, s_gray, ey * EW + ex, ey * US_GP + ex, va);
            if (ey + 1 < eh) p2_store<SHARPEN>(s_up, s_gray, (ey + 1) * EW + ex, (ey + 1) * US_GP + ex, vb);
            p2_rows<S, HALO, SHARPEN, J + 2, NR>(W, bj, s_up, s_gray, ey0, ex, eh);
        } else {
            const float4 b = bj[p2_phase(S, HALO, J)];
            uint32_t v[3];
#pragma unroll
            for (int ch = 0; ch < 3; ++ch)
                v[ch] = f2u8(__fadd_rn(__fmaf_rn(W[o][ch], b.x, __fmul_rn(W[o + 1][ch], b.y)), __fmaf_rn(W[o + 2][ch], b.z, __fmul_rn(W[o + 3][ch], b.w))));
            if (ey < eh) p2_store<SHARPEN>(s_up, s_gray, ey * EW + ex, ey * US_GP + ex, v);
            p2_rows<S, HALO, SHARPEN, J + 1, NR>(W, bj, s_up, s_gray, ey0, ex, eh);
        }
    }
}

template <int S, int HALO, bool SHARPEN>
__device__ __forceinline__ void p2_fast(const float* __restrict__ s_hs, const int* __restrict__ s_yo, const float4* __restrict__ s_yc,
                                        uint32_t* s_up, uint8_t* s_gray, int tid) {
    constexpr int EW = US_TW + 2 * HALO, EH = US_TH + 2 * HALO, NRUN = (EH + 3) / 4;
    constexpr int NR = 4 + p2_off(S, HALO, 3);
    float4 bj[S];                                   // coefficient set of phase p = that of the first ext row with that phase
#pragma unroll
    for (int p = 0; p < S; ++p) bj[p] = s_yc[(p + HALO) % S];
    for (int i = tid; i < NRUN * EW; i += US_NT) {
        const int run = i / EW, ex = i - run * EW;
        const int ey0 = run * 4;
        const float* hp = s_hs + (s_yo[ey0] * EW + ex) * 3;
        float W[NR][3];
#pragma unroll
        for (int k = 0; k < NR; ++k) { W[k][0] = hp[k * EW * 3]; W[k][1] = hp[k * EW * 3 + 1]; W[k][2] = hp[k * EW * 3 + 2]; }
        p2_rows<S, HALO, SHARPEN, 0, NR>(W, bj, s_up, s_gray, ey0, ex, EH);
    }
}

// OUTX: UpsharpArgs::out_bgrx as a compile-time flag (its own kernel: the three-byte form carries none of its code)
template <bool RESIZE, bool SHARPEN, int GK, bool OUTX = false>
__global__ void __launch_bounds__(US_NT) k_upsharp(const __grid_constant__ UpsharpArgs a) {
    constexpr int HALO = SHARPEN ? US_HALO : 0;
    constexpr int EW = US_TW + 2 * HALO, EH = US_TH + 2 * HALO;
    constexpr int GR = GK / 2;
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ unsigned long long s_part[(US_NT / 32) * 6];
    const int tid = threadIdx.x;
    const UpsharpSmem L = upsharp_layout(RESIZE, SHARPEN, a.max_src_rows, a.max_src_cols);
    uint8_t* s_raw = smem + L.raw;
    uint32_t* s_src = reinterpret_cast<uint32_t*>(smem + L.src);
    float* s_hs = reinterpret_cast<float*>(smem + L.hs);
    uint32_t* s_rowbg = reinterpret_cast<uint32_t*>(smem + L.rowbg);
    uint16_t* s_rowr = reinterpret_cast<uint16_t*>(smem + L.rowr);
    float* s_rowf = reinterpret_cast<float*>(smem + L.rowf);
    uint32_t* s_up = reinterpret_cast<uint32_t*>(smem + L.up);
    uint8_t* s_gray = smem + L.gray;
    float* s_sig = reinterpret_cast<float*>(smem + L.sig);
    uint8_t* s_out = smem + L.sig;
    uint8_t* tp = smem + L.tabs;
    float4* s_yc = reinterpret_cast<float4*>(tp); tp += EH * 16;
    int4* s_xc = reinterpret_cast<int4*>(tp); tp += EW * 16;
    uint64_t* bar = reinterpret_cast<uint64_t*>(tp); tp += 16;
    int* s_xo = reinterpret_cast<int*>(tp); tp += EW * 4;
    int* s_yo = reinterpret_cast<int*>(tp); tp += EH * 4;
    float* s_lut = reinterpret_cast<float*>(tp);

    const int x0 = blockIdx.x * US_TW, y0 = blockIdx.y * US_TH;

    // ---- P0: tables for this tile + source window -------------------------------------------------
    const int dxa = max(0, x0 - HALO), dxb = min(a.dw - 1, x0 + US_TW - 1 + HALO);
    const int dya = max(0, y0 - HALO), dyb = min(a.dh - 1, y0 + US_TH - 1 + HALO);
    int ux0, ux1, uy0, uy1;   // inclusive source window in unclamped coordinates
    if (RESIZE) {
        ux0 = a.rt.xofs[dxa]; ux1 = a.rt.xofs[dxb] + 3;
        uy0 = a.rt.yofs[dya]; uy1 = a.rt.yofs[dyb] + 3;
    } else {
        ux0 = dxa; ux1 = dxb; uy0 = dya; uy1 = dyb;
    }
    // the source window's bulk copies are in flight while the per-tile tables are filled
    const int sx0 = clampi(ux0, 0, a.sw - 1), sx1 = clampi(ux1, 0, a.sw - 1);
    const int sy0 = clampi(uy0, 0, a.sh - 1), sy1 = clampi(uy1, 0, a.sh - 1);
    const int bx0 = (sx0 * 3) & ~15, bx1 = (sx1 + 1) * 3;
    const bool tm_src = a.use_tm_src != 0;                    // implies fast_stage
    const int RP = L.raw_pitch - (tm_src ? 16 : 0);           // a TMA box is dense
    const bool fast_stage = aligned16(a.src, a.spitch);
    const bool srcx = RESIZE && a.src_bgrx != 0;
    const int ux0a = srcx ? ux0 - (((ux0 % 4) + 4) % 4) : ux0;   // 32-bit source: the box starts on a multiple of 4 pixels
    if (srcx) stage_box_issue(smem + L.raw, &a.tm_src, ux0a, uy0, (uint32_t)(a.src_pitch_x * a.max_src_rows * 4), bar);
    else if (tm_src) stage_box_issue(s_raw, &a.tm_src, bx0, sy0, (uint32_t)(RP * a.max_src_rows), bar);
    else stage_rows_issue(s_raw, RP, a.src, a.spitch, sy0, sy1 + 1, bx0, bx1, fast_stage, bar);
    if (RESIZE) {
        for (int i = tid; i < EW; i += US_NT) {
            const int dx = clampi(reflect101(x0 - HALO + i, a.dw), dxa, dxb);   // clamp: unused columns of edge tiles
            s_xo[i] = a.rt.xofs[dx] - ux0a; s_xc[i] = a.rt.xcoef[dx];
        }
        for (int i = tid; i < EH; i += US_NT) {
            const int dy = clampi(reflect101(y0 - HALO + i, a.dh), dya, dyb);
            s_yo[i] = a.rt.yofs[dy] - uy0; s_yc[i] = a.rt.ycoef[dy];
        }
    }
    if (SHARPEN) s_lut[tid] = a.sig_lut[tid];   // US_NT == 256
    stage_rows_wait(s_raw, RP, a.src, a.spitch, 3 * a.sw, sy0, sy1 + 1, bx0, bx1, fast_stage, bar);

    if (RESIZE) {
        const int SP = srcx ? a.src_pitch_x : L.src_pitch;
        if (srcx) s_src = reinterpret_cast<uint32_t*>(smem + L.raw);      // the TMA box IS the word tile
        const int ncols = ux1 - ux0 + 1, nrows = uy1 - uy0 + 1;
        // raw bytes -> BGR0 words, replicate border materialised; warp per row, lanes along x
        const int lane = tid & 31, wrp = tid >> 5;
        if (srcx) {
            // positions outside the image (zero-filled by the TMA unit) take their replicate-border source, which lies inside
            // the image and inside this window
            if (ux0a < 0 || uy0 < 0 || ux0a + SP > a.sw || uy0 + a.max_src_rows > a.sh) {
                for (int i = tid; i < a.max_src_rows * SP; i += US_NT) {
                    const int rr = i / SP, cc = i - rr * SP;
                    const int ux = ux0a + cc, uy = uy0 + rr;
                    if ((unsigned)ux < (unsigned)a.sw && (unsigned)uy < (unsigned)a.sh) continue;
                    const int qx = clampi(ux, 0, a.sw - 1) - ux0a, qy = clampi(uy, 0, a.sh - 1) - uy0;
                    if ((unsigned)qx < (unsigned)SP && (unsigned)qy < (unsigned)a.max_src_rows) s_src[i] = s_src[qy * SP + qx];
                }
            }
        } else if (ux0 >= 0 && ux1 < a.sw && aligned16(a.src, a.spitch)) {
            // no horizontal clamping in this tile: a lane turns 12 raw bytes (4 pixels, any byte phase) into four
            // BGR0 words with funnel shifts and byte permutes and stores them as one 128-bit word
            const int c0 = 3 * ux0 - bx0;               // byte offset of window column 0 inside a raw row, 0..15
            const int sh = 8 * (c0 & 3);
            // the (row, 4-pixel group) items are spread over all threads, not a warp per row: a row is ~10 groups
            const int ng = (ncols + 3) >> 2;
            const int drr = US_NT / ng, dg = US_NT - drr * ng;
            int rr = tid / ng, g = tid - rr * ng;
            for (; rr < nrows; rr += drr) {
                const uint32_t* q = reinterpret_cast<const uint32_t*>(s_raw + (clampi(uy0 + rr, 0, a.sh - 1) - sy0) * RP + (c0 & ~3)) + 3 * g;
                const uint32_t r0 = q[0], r1 = q[1], r2 = q[2], r3 = q[3];
                const uint32_t w0 = __funnelshift_r(r0, r1, sh), w1 = __funnelshift_r(r1, r2, sh), w2 = __funnelshift_r(r2, r3, sh);
                uint4 o;
                o.x = w0 & 0x00ffffffu;
                o.y = __byte_perm(w0, w1, 0x7543) & 0x00ffffffu;
                o.z = __byte_perm(w1, w2, 0x7432) & 0x00ffffffu;
                o.w = w2 >> 8;
                *reinterpret_cast<uint4*>(s_src + rr * SP + 4 * g) = o;
                g += dg;
                if (g >= ng) { g -= ng; ++rr; }
            }
        } else {
            for (int rr = wrp; rr < nrows; rr += US_NT / 32) {
                const uint8_t* rrow = s_raw + (clampi(uy0 + rr, 0, a.sh - 1) - sy0) * RP - bx0;
                for (int cc = lane; cc < ncols; cc += 32) {
                    const uint8_t* p = rrow + 3 * clampi(ux0 + cc, 0, a.sw - 1);
                    s_src[rr * SP + cc] = (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16);
                }
            }
        }
        __syncthreads();
        // ---- P1: horizontal pass, int32 exact, stored as float32 (|H| < 2^22); 2 ext columns/thread ----
        {
            constexpr int NP = EW / 2;
            for (int i = tid; i < nrows * NP; i += US_NT) {
                const int rr = i / NP, p = i - rr * NP;
                const int ca = s_xo[2 * p], cb = s_xo[2 * p + 1];
                const int4 ka = s_xc[2 * p], kb = s_xc[2 * p + 1];
                const uint32_t* row = s_src + rr * SP;
                const uint32_t a0 = row[ca], a1 = row[ca + 1], a2 = row[ca + 2], a3 = row[ca + 3];
                uint32_t b0, b1, b2, b3;
                const int d = cb - ca;
                if (d == 1) { b0 = a1; b1 = a2; b2 = a3; b3 = row[ca + 4]; }
                else if (d == 0) { b0 = a0; b1 = a1; b2 = a2; b3 = a3; }
                else { b0 = row[cb]; b1 = row[cb + 1]; b2 = row[cb + 2]; b3 = row[cb + 3]; }
                float h[6];
#pragma unroll
                for (int ch = 0; ch < 3; ++ch) {
                    const int sel = 0x4440 + ch;   // byte ch -> low byte, zeros above
                    const int ha = (int)__byte_perm(a0, 0, sel) * ka.x + (int)__byte_perm(a1, 0, sel) * ka.y +
                                   (int)__byte_perm(a2, 0, sel) * ka.z + (int)__byte_perm(a3, 0, sel) * ka.w;
                    const int hb = (int)__byte_perm(b0, 0, sel) * kb.x + (int)__byte_perm(b1, 0, sel) * kb.y +
                                   (int)__byte_perm(b2, 0, sel) * kb.z + (int)__byte_perm(b3, 0, sel) * kb.w;
                    // exact int -> float for |v| < 2^22: (v + 0x4B400000) reinterpreted is 1.5*2^23 + v
                    h[ch] = __int_as_float(ha + 0x4B400000) - 12582912.0f;
                    h[3 + ch] = __int_as_float(hb + 0x4B400000) - 12582912.0f;
                }
                float2* o = reinterpret_cast<float2*>(s_hs + (rr * EW + 2 * p) * 3);
                o[0] = make_float2(h[0], h[1]); o[1] = make_float2(h[2], h[3]); o[2] = make_float2(h[4], h[5]);
            }
        }
        __syncthreads();
        // ---- P2: vertical pass, float32, RNE + saturate; runs of 4 ext rows --------------------------------
        // integer ratio and no mirrored ext rows in this tile: the statically scheduled form (p2_fast)
        const bool rows_inside = (y0 - HALO >= 0) && (y0 + US_TH - 1 + HALO <= a.dh - 1);
        if (rows_inside && a.yperiod == 2) p2_fast<2, HALO, SHARPEN>(s_hs, s_yo, s_yc, s_up, s_gray, tid);
        else if (rows_inside && a.yperiod == 4) p2_fast<4, HALO, SHARPEN>(s_hs, s_yo, s_yc, s_up, s_gray, tid);
        else {   // any ratio / edge tiles: sliding 4-row window
            constexpr int RUN = 4, NRUN = (EH + RUN - 1) / RUN;
            for (int i = tid; i < NRUN * EW; i += US_NT) {
                const int run = i / EW, ex = i - run * EW;
                const int ey0 = run * RUN;
                float w[4][3];
                int yo = s_yo[ey0];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float* hp = s_hs + ((yo + k) * EW + ex) * 3;
                    w[k][0] = hp[0]; w[k][1] = hp[1]; w[k][2] = hp[2];
                }
#pragma unroll
                for (int j = 0; j < RUN; ++j) {
                    const int ey = ey0 + j;
                    if (ey < EH) {
                        if (j > 0) {
                            const int yn = s_yo[ey];
                            const int d = yn - yo;
                            if (d == 1) {
#pragma unroll
                                for (int ch = 0; ch < 3; ++ch) { w[0][ch] = w[1][ch]; w[1][ch] = w[2][ch]; w[2][ch] = w[3][ch]; }
                                const float* hp = s_hs + ((yn + 3) * EW + ex) * 3;
                                w[3][0] = hp[0]; w[3][1] = hp[1]; w[3][2] = hp[2];
                            } else if (d != 0) {
#pragma unroll
                                for (int k = 0; k < 4; ++k) {
                                    const float* hp = s_hs + ((yn + k) * EW + ex) * 3;
                                    w[k][0] = hp[0]; w[k][1] = hp[1]; w[k][2] = hp[2];
                                }
                            }
                            yo = yn;
                        }
                        const float4 b = s_yc[ey];
                        uint32_t v[3];
#pragma unroll
                        for (int ch = 0; ch < 3; ++ch)
                            v[ch] = f2u8(__fadd_rn(__fmaf_rn(w[0][ch], b.x, __fmul_rn(w[1][ch], b.y)),
                                                   __fmaf_rn(w[2][ch], b.z, __fmul_rn(w[3][ch], b.w))));
                        s_up[ey * EW + ex] = v[0] | (v[1] << 8) | (v[2] << 16);
                        if (SHARPEN)
                            s_gray[ey * US_GP + ex] = (uint8_t)((9798 * v[2] + 19235 * v[1] + 3735 * v[0] + (1 << 14)) >> 15);
                    }
                }
            }
        }
    } else {
        // standalone AdaptiveSharpening: ext tile straight from the (reflect-padded) source
        const int lane = tid & 31, wrp = tid >> 5;
        for (int ey = wrp; ey < EH; ey += US_NT / 32) {
            const int sy = clampi(reflect101(y0 - HALO + ey, a.dh), uy0, uy1);
            const uint8_t* rrow = s_raw + (sy - sy0) * RP - bx0;
            for (int ex = lane; ex < EW; ex += 32) {
                const int sx = clampi(reflect101(x0 - HALO + ex, a.dw), ux0, ux1);
                const uint8_t* p = rrow + sx * 3;
                const uint32_t v0 = p[0], v1 = p[1], v2 = p[2];
                s_up[ey * EW + ex] = v0 | (v1 << 8) | (v2 << 16);
                if (SHARPEN) s_gray[ey * US_GP + ex] = (uint8_t)((9798 * v2 + 19235 * v1 + 3735 * v0 + (1 << 14)) >> 15);
            }
        }
    }
    __syncthreads();

    const int vw = min(US_TW, a.dw - x0), vh = min(US_TH, a.dh - y0);   // valid part of the tile
    unsigned acc6[6] = {0, 0, 0, 0, 0, 0};

    if (!SHARPEN) {
        for (int i = tid; i < US_TH * US_TW; i += US_NT) {
            const int y = i >> 6, x = i & 63;
            const uint32_t px = s_up[y * EW + x];
            uint8_t* o = s_out + y * (US_TW * 3) + x * 3;
            const unsigned v0 = px & 0xff, v1 = (px >> 8) & 0xff, v2 = (px >> 16) & 0xff;
            if (OUTX) reinterpret_cast<uint32_t*>(s_out)[y * US_TW + x] = (px & 0xffffffu) | 0x01000000u;
            else { o[0] = v0; o[1] = v1; o[2] = v2; }
            if (a.sums && x < vw && y < vh) {
                acc6[0] += v0; acc6[1] += v1; acc6[2] += v2;
                acc6[3] += v0 * v0; acc6[4] += v1 * v1; acc6[5] += v2 * v2;
            }
        }
    } else {
        // caller-supplied / exported masks exist only on the stage-wise (no-resize) entry points
        constexpr bool MASKIO = !RESIZE;
        const bool own_mask = !MASKIO || (a.mask_in == nullptr);
        const bool mask_only = MASKIO && own_mask && a.dst == nullptr && a.sums == nullptr && a.mask_out != nullptr;
        constexpr int CW = US_TW + 4, CH = US_TH + 4;
        if (own_mask) {
            // ---- P3: Sobel / Laplacian magnitude -> sigmoid on (TW+4)x(TH+4); column walk, 3-row window ---
            constexpr int NG = 3, RG = CH / NG;   // 3 groups of 12 comb rows
            if (tid < CW * NG) {
                const int g = tid / CW, cx = tid - g * CW;
                const int cy0 = g * RG;
                const uint8_t* gp = s_gray + cy0 * US_GP + cx;
                int h0, s0, t0, h1, s1, t1, bm;
                { const int p = gp[0], q = gp[1], r = gp[2]; h0 = r - p; s0 = p + 2 * q + r; t0 = p + r; }
                { const int p = gp[US_GP], q = gp[US_GP + 1], r = gp[US_GP + 2]; h1 = r - p; s1 = p + 2 * q + r; t1 = p + r; bm = q; }
#pragma unroll 4
                for (int k = 0; k < RG; ++k) {
                    const uint8_t* g2 = gp + (k + 2) * US_GP;
                    const int p = g2[0], q = g2[1], r = g2[2];
                    const int h2 = r - p, s2 = p + 2 * q + r, t2 = p + r;
                    const int sxv = h0 + 2 * h1 + h2, syv = s2 - s0, lap = 2 * (t0 + t2) - 8 * bm;
                    const int ax = min(abs(sxv), 255), ay = min(abs(syv), 255), al = min(abs(lap), 255);
                    const int ss = ax + ay;
                    const int e1 = (ss + ((ss >> 1) & 1)) >> 1;                 // RNE((ax+ay)/2)
                    const int comb = ((6 * e1 + 4 * al + 5) * 6554) >> 16;     // RNE(fma(e1,.6f,al*.4f))
                    s_sig[(cy0 + k) * CW + cx] = s_lut[comb];
                    h0 = h1; s0 = s1; t0 = t1; h1 = h2; s1 = s2; t1 = t2; bm = q;
                }
            }
            __syncthreads();
        }
        {
            const int x = tid & 63, g = tid >> 6;
            // ---- P4a: float32 row pass of the mask blur ------------------------------------------
            if (own_mask) {
                for (int cy = g; cy < CH; cy += 4) {
                    const float* p = s_sig + cy * CW + x;
                    float acc = __fmul_rn(p[0], a.gf[0]);
#pragma unroll
                    for (int k = 1; k < 5; ++k) acc = __fadd_rn(acc, __fmul_rn(p[k], a.gf[k]));
                    s_rowf[cy * US_TW + x] = acc;
                }
            }
            // ---- P4b: Q8.8 row pass of the 8-bit Gaussian; B and G share one 32-bit accumulator ---
            if (!mask_only) {
                int q[GK];
#pragma unroll
                for (int k = 0; k < GK; ++k) q[k] = a.gq[k];
                constexpr int NROW = US_TH + 2 * GR;
                for (int ry = g; ry < NROW; ry += 4) {
                    const uint32_t* p = s_up + (ry + US_HALO - GR) * EW + (x + US_HALO - GR);
                    uint32_t sbg = 0, sr = 0;
#pragma unroll
                    for (int k = 0; k < GK; ++k) {
                        const uint32_t px = p[k];
                        sbg += __byte_perm(px, 0, 0x4140) * (uint32_t)q[k];    // [B,0,G,0]
                        sr += __byte_perm(px, 0, 0x4442) * (uint32_t)q[k];
                    }
                    s_rowbg[ry * US_TW + x] = sbg;
                    s_rowr[ry * US_TW + x] = (uint16_t)sr;
                }
            }
        }
        __syncthreads();
        if (mask_only) {
            // createEdgeMask alone (vp_edge_mask, the adaptive_sigma branch): column pass of the mask blur, nothing else
            const int x = tid & 63, yb = (tid >> 6) * 8;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int y = yb + i;
                float e = __fmul_rn(s_rowf[y * US_TW + x], a.gf[0]);
#pragma unroll
                for (int k = 1; k < 5; ++k) e = __fadd_rn(e, __fmul_rn(s_rowf[(y + k) * US_TW + x], a.gf[k]));
                if (x < vw && y < vh)
                    *reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(a.mask_out) + (size_t)(y0 + y) * a.mask_out_pitch + (size_t)(x0 + x) * 4) = e;
            }
        } else
        // ---- P5: column passes (sliding windows), unsharp blend, tone preservation ------------------
        {
            const int x = tid & 63, g = tid >> 6;
            const int yb = g * 8;
            float fw[5];
#if VP_US_P5_FLOAT
            unsigned long long bgf[GK], qf2[GK];
            float rf[GK], qf[GK];
#pragma unroll
            for (int k = 0; k < GK; ++k) { qf[k] = (float)a.gq[k]; qf2[k] = pack2(__float_as_uint(qf[k]), __float_as_uint(qf[k])); }
#else
            uint32_t bgw[GK], rw[GK];
            int q[GK];
#pragma unroll
            for (int k = 0; k < GK; ++k) q[k] = a.gq[k];
#endif
            const float gf0 = a.gf[0], gf1 = a.gf[1], gf2 = a.gf[2], gf3 = a.gf[3], gf4 = a.gf[4];
            if (own_mask) {
#pragma unroll
                for (int k = 0; k < 4; ++k) fw[k] = s_rowf[(yb + k) * US_TW + x];
            }
#pragma unroll
            for (int k = 0; k < GK - 1; ++k) {
#if VP_US_P5_FLOAT
                const uint32_t nbg = s_rowbg[(yb + k) * US_TW + x];
                const uint32_t nr = s_rowr[(yb + k) * US_TW + x];
                bgf[k] = fadd2(pack2(__byte_perm(nbg, 0x4B000000u, 0x7410), __byte_perm(nbg, 0x4B000000u, 0x7432)), pack2(0xCB000000u, 0xCB000000u));
                rf[k] = __uint_as_float(__byte_perm(nr, 0x4B000000u, 0x7410)) - 8388608.0f;
#else
                bgw[k] = s_rowbg[(yb + k) * US_TW + x]; rw[k] = s_rowr[(yb + k) * US_TW + x];
#endif
            }
            const float es = a.edge_strength, ssm = a.smooth_strength, stg = a.strength;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int y = yb + i;
                float e;
                if (own_mask) {
                    fw[(i + 4) % 5] = s_rowf[(y + 4) * US_TW + x];
                    e = __fmul_rn(fw[i % 5], gf0);
                    e = __fadd_rn(e, __fmul_rn(fw[(i + 1) % 5], gf1));
                    e = __fadd_rn(e, __fmul_rn(fw[(i + 2) % 5], gf2));
                    e = __fadd_rn(e, __fmul_rn(fw[(i + 3) % 5], gf3));
                    e = __fadd_rn(e, __fmul_rn(fw[(i + 4) % 5], gf4));
                } else {
                    const int gx = min(x0 + x, a.dw - 1), gy = min(y0 + y, a.dh - 1);
                    e = *reinterpret_cast<const float*>(reinterpret_cast<const uint8_t*>(a.mask_in) + (size_t)gy * a.mask_in_pitch + (size_t)gx * 4);
                }
                if (MASKIO && a.mask_out && x < vw && y < vh)
                    *reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(a.mask_out) + (size_t)(y0 + y) * a.mask_out_pitch + (size_t)(x0 + x) * 4) = e;
#if VP_US_P5_FLOAT
                // Q8.16 column pass, unsharp mask and blend in float32: the row sums (<= 65,280), their products with the
                // Q0.8 taps and the column sums (< 2^24) are integers a float holds exactly, so the packed FMA chains below
                // equal the integer arithmetic bit for bit; (v + 2^15) >> 16 is floor(v * 2^-16 + 0.5), taken by an add
                // rounded toward -inf onto 2^23, and `up` enters as 2^23 + u straight out of a PRMT, so u - blur needs no
                // unbiasing at all.
                {
                    const uint32_t nbg = s_rowbg[(y + GK - 1) * US_TW + x];
                    const uint32_t nr = s_rowr[(y + GK - 1) * US_TW + x];
                    bgf[(i + GK - 1) % GK] = fadd2(pack2(__byte_perm(nbg, 0x4B000000u, 0x7410), __byte_perm(nbg, 0x4B000000u, 0x7432)),
                                                   pack2(0xCB000000u, 0xCB000000u));
                    rf[(i + GK - 1) % GK] = __uint_as_float(__byte_perm(nr, 0x4B000000u, 0x7410)) - 8388608.0f;
                }
                unsigned long long cbg = fmul2(bgf[i % GK], qf2[0]);
                float crf = __fmul_rn(rf[i % GK], qf[0]);
#pragma unroll
                for (int k = 1; k < GK; ++k) {
                    cbg = ffma2(bgf[(i + k) % GK], qf2[k], cbg);
                    crf = __fmaf_rn(rf[(i + k) % GK], qf[k], crf);
                }
                const uint32_t px = s_up[(y + US_HALO) * EW + x + US_HALO];
                const int u[3] = {(int)(px & 0xff), (int)((px >> 8) & 0xff), (int)((px >> 16) & 0xff)};
                const float t1 = __fmul_rn(e, es);
                const float t2 = __fmul_rn(__fsub_rn(1.0f, e), ssm);
                const float st = __fmul_rn(stg, __fadd_rn(t1, t2));
                const unsigned long long c16 = pack2(0x37800000u, 0x37800000u), half2 = pack2(0x3F000000u, 0x3F000000u);   // 2^-16, 0.5
                const unsigned long long big = pack2(0x4B000000u, 0x4B000000u), nbig = pack2(0xCB000000u, 0xCB000000u);     // +-2^23
                const unsigned long long blb = fadd2_rm(ffma2(cbg, c16, half2), big);                     // 2^23 + blur (B, G)
                const float blr = __fadd_rd(__fmaf_rn(crf, 1.52587890625e-05f, 0.5f), 8388608.0f);         // 2^23 + blur R
                const unsigned long long ubg = pack2(__byte_perm(px, 0x4B000000u, 0x7440), __byte_perm(px, 0x4B000000u, 0x7441));   // 2^23 + u
                const float ubr = __uint_as_float(__byte_perm(px, 0x4B000000u, 0x7442));
                float umb, umg;
                unpack2(ffma2(blb, pack2(0xBF800000u, 0xBF800000u), ubg), umb, umg);                      // u - blur, exact
                const float umr = fmaxf(__fsub_rn(ubr, blr), 0.0f);
                const unsigned long long um2 = pack2(__float_as_uint(fmaxf(umb, 0.0f)), __float_as_uint(fmaxf(umg, 0.0f)));
                const uint32_t stb = __float_as_uint(st);
                float ob, og;
                unpack2(fadd2(fadd2(ubg, nbig), fmul2(pack2(stb, stb), um2)), ob, og);                    // u + st * um: separate mul, add
                int o[3];
                o[0] = (int)f2u8(ob); o[1] = (int)f2u8(og);
                o[2] = (int)f2u8(__fadd_rn(__fsub_rn(ubr, 8388608.0f), __fmul_rn(st, umr)));
#else
                bgw[(i + GK - 1) % GK] = s_rowbg[(y + GK - 1) * US_TW + x];
                rw[(i + GK - 1) % GK] = s_rowr[(y + GK - 1) * US_TW + x];
                uint32_t cb_ = 1u << 15, cg_ = 1u << 15, cr_ = 1u << 15;
#pragma unroll
                for (int k = 0; k < GK; ++k) {
                    const uint32_t bg = bgw[(i + k) % GK];
                    cb_ += (bg & 0xffffu) * (uint32_t)q[k];
                    cg_ += (bg >> 16) * (uint32_t)q[k];
                    cr_ += rw[(i + k) % GK] * (uint32_t)q[k];
                }
                const int bl[3] = {(int)(cb_ >> 16), (int)(cg_ >> 16), (int)(cr_ >> 16)};
                const uint32_t px = s_up[(y + US_HALO) * EW + x + US_HALO];
                const int u[3] = {(int)(px & 0xff), (int)((px >> 8) & 0xff), (int)((px >> 16) & 0xff)};
                const float t1 = __fmul_rn(e, es);
                const float t2 = __fmul_rn(__fsub_rn(1.0f, e), ssm);
                const float st = __fmul_rn(stg, __fadd_rn(t1, t2));
                int o[3];
#pragma unroll
                for (int ch = 0; ch < 3; ++ch) {
                    const int um = max(u[ch] - bl[ch], 0);
                    // I2F here on purpose: the conversion pipe is idle in this kernel, the issue slots are not
                    o[ch] = (int)f2u8(__fadd_rn((float)u[ch], __fmul_rn(st, (float)um)));
                }
#endif
                if (a.preserve_tone) {
                    const int yi = (4899 * u[2] + 9617 * u[1] + 1868 * u[0] + (1 << 13)) >> 14;
                    const int cr = sat_u8(((u[2] - yi) * 11682 + (128 << 14) + (1 << 13)) >> 14) - 128;
                    const int cb = sat_u8(((u[0] - yi) * 9241 + (128 << 14) + (1 << 13)) >> 14) - 128;
                    const int yo = (4899 * o[2] + 9617 * o[1] + 1868 * o[0] + (1 << 13)) >> 14;
                    o[0] = sat_u8(yo + ((cb * 29049 + (1 << 13)) >> 14));
                    o[1] = sat_u8(yo + ((cb * -5636 + cr * -11698 + (1 << 13)) >> 14));
                    o[2] = sat_u8(yo + ((cr * 22987 + (1 << 13)) >> 14));
                }
                if (OUTX) {
                    reinterpret_cast<uint32_t*>(s_out)[y * US_TW + x] = (uint32_t)o[0] | ((uint32_t)o[1] << 8) | ((uint32_t)o[2] << 16) | 0x01000000u;
                } else {
                    uint8_t* ob = s_out + y * (US_TW * 3) + x * 3;
                    ob[0] = (uint8_t)o[0]; ob[1] = (uint8_t)o[1]; ob[2] = (uint8_t)o[2];
                }
                if (a.sums && x < vw && y < vh) {
#pragma unroll
                    for (int ch = 0; ch < 3; ++ch) { acc6[ch] += o[ch]; acc6[3 + ch] += o[ch] * o[ch]; }
                }
            }
        }
    }
    if (a.sums) warp_partials6(acc6, s_part);
    fence_proxy_async();    // staged tile (generic-proxy writes) -> visible to the TMA store engine
    __syncthreads();
    // warp 1 publishes the statistics (global atomics, fence, ticket) while warp 0 drives the tile's TMA stores
    if (a.sums && (tid >> 5) == 1) warp_flush_and_finalize(s_part, US_NT / 32, a.sums, a.fin, gridDim.x * gridDim.y);

    // ---- P6: staged tile -> HBM: one TMA bulk store per row (16-byte aligned frames), else bytes ----
    if (a.dst && a.use_tm_dst) {
        // one box store; the part of an edge tile outside the frame is clipped by the TMA unit
        if (tid == 0) { tma_store_2d(&a.tm_dst, OUTX ? x0 : x0 * 3, y0, s_out); bulk_commit(); bulk_wait_read0(); }
    } else if (a.dst) {
        const int rb = vw * 3;                               // valid bytes per tile row
        const bool al = aligned16(a.dst, a.dpitch);          // x0*3 = 192*blockIdx.x is a multiple of 16
        const int nb16 = al ? (rb & ~15) : 0;
        if (nb16) {
            if (tid < vh) {
                bulk_s2g(a.dst + (size_t)(y0 + tid) * a.dpitch + (size_t)x0 * 3, s_out + tid * (US_TW * 3), (uint32_t)nb16);
                bulk_commit();
            }
        }
        const int tb = rb - nb16;
        for (int i = tid; i < tb * vh; i += US_NT) {
            const int y = i / tb, q = nb16 + (i - y * tb);
            a.dst[(size_t)(y0 + y) * a.dpitch + (size_t)x0 * 3 + q] = s_out[y * (US_TW * 3) + q];
        }
        if (nb16 && tid < vh) bulk_wait_read0();
    }
}

}  // namespace vp
