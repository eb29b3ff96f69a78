// vp_bicubic.cuh - lean cv::resize(INTER_CUBIC) for CV_8UC3 (reference call sites src/upscaler.cpp:81,
// src/test_enhancements.cpp:348, src/main.cpp:265): the bicubic-only path (BASELINE configs C1, C5 and
// Upscaler::BICUBIC).  Bit-exact to OpenCV (oracle/restate.py:resize_cubic): int32 horizontal pass with
// 11-bit taps, float32 vertical pass t = fma(S0,b0,S1*b1) + fma(S2,b2,S3*b3), RNE + saturate.
//
// One thread owns one destination column of a 128 x 32 tile and walks down its rows:
//   * the source window is staged by TMA bulk row copies as raw interleaved bytes and read in place;
//   * the column's four (border-clamped) tap offsets and taps are per-thread constants; a sliding window of four
//     horizontally filtered rows (3 channels) lives in registers and advances only when the vertical
//     tap origin changes (warp-uniform), so every H value is computed exactly once per tile column;
//   * four adjacent lanes repack their 4 x 3 output bytes into three aligned 32-bit words with one
//     shuffle and a funnel shift: no shared-memory staging of the output, 96 contiguous bytes per warp row.
// Any upscale ratio works (tables come from the host exactly as cv::resize builds them).
#pragma once
#include "vp_bilateral.cuh"   // f2u8
#include "vp_common.cuh"
#include "vp_upsharp.cuh"     // ResizeTables

namespace vp {

constexpr int BC_TW = 128, BC_TH = 32, BC_NT = 128;

struct BicubicArgs {
    const uint8_t* src; size_t spitch; int sw, sh;
    uint8_t* dst; size_t dpitch; int dw, dh;
    ResizeTables rt;
    int max_src_rows, max_src_cols;     // unclamped per-tile source window bounds
    int yperiod;                        // S when dst_h == S * src_h with S = 2 or 4 (host-verified row tables), else 0
};

struct BicubicSmem { int raw, raw_pitch, src, src_pitch, tabs, out, total; };

__host__ __device__ inline BicubicSmem bicubic_layout(int max_src_rows, int max_src_cols) {
    BicubicSmem L{};
    int o = 0;
    L.raw_pitch = ((max_src_cols * 3 + 15 + 15) & ~15) + 16;
    L.raw = o; o += L.raw_pitch * max_src_rows;
    L.src_pitch = 0; L.src = o;
    L.tabs = o; o += BC_TH * 16 + BC_TH * 4 + 16;      // ycoef float4[TH], yofs int[TH], mbarrier
    L.out = (o + 15) & ~15; o = L.out + BC_TH * BC_TW * 3;   // staged output tile (integer-ratio path, TMA bulk stores)
    L.total = o;
    return L;
}

// S = 0: any ratio (row tables read per row).  S = 2 / 4: the vertical tap origin advances exactly once every S output
// rows, at (y % S) == S/2, and the S coefficient sets repeat - so the step is unrolled statically, the coefficients live
// in registers, and the S rows that share a window are filtered two at a time with packed float32 FMAs.
template <int S>
__global__ void __launch_bounds__(BC_NT) k_bicubic(const BicubicArgs a) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int tid = threadIdx.x;
    const BicubicSmem L = bicubic_layout(a.max_src_rows, a.max_src_cols);
    uint8_t* s_raw = smem + L.raw;
    float4* s_yc = reinterpret_cast<float4*>(smem + L.tabs);
    int* s_yo = reinterpret_cast<int*>(smem + L.tabs + BC_TH * 16);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + L.tabs + BC_TH * 16 + BC_TH * 4);

    const int x0 = blockIdx.x * BC_TW, y0 = blockIdx.y * BC_TH;
    const int dxb = min(a.dw - 1, x0 + BC_TW - 1), dyb = min(a.dh - 1, y0 + BC_TH - 1);
    const int ux0 = a.rt.xofs[x0], ux1 = a.rt.xofs[dxb] + 3;      // unclamped source window (inclusive)
    const int uy0 = a.rt.yofs[y0], uy1 = a.rt.yofs[dyb] + 3;
    // the source window's bulk copies are in flight while the row tables and this column's taps are fetched
    const int sx0 = clampi(ux0, 0, a.sw - 1), sx1 = clampi(ux1, 0, a.sw - 1);
    const int sy0 = clampi(uy0, 0, a.sh - 1), sy1 = clampi(uy1, 0, a.sh - 1);
    const int bx0 = (sx0 * 3) & ~15, bx1 = (sx1 + 1) * 3;
    const int RP = L.raw_pitch;
    const bool fast_stage = aligned16(a.src, a.spitch);
    stage_rows_issue(s_raw, RP, a.src, a.spitch, sy0, sy1 + 1, bx0, bx1, fast_stage, bar);
    if (tid < BC_TH) {
        const int dy = min(y0 + tid, dyb);
        s_yo[tid] = a.rt.yofs[dy] - uy0;
        s_yc[tid] = a.rt.ycoef[dy];
    }
    // this thread's column: tap origin (relative to the window) and taps
    const int dx = min(x0 + tid, dxb);
    const int cx = a.rt.xofs[dx] - ux0;
    const int4 kx = a.rt.xcoef[dx];
    stage_rows_wait(s_raw, RP, a.src, a.spitch, 3 * a.sw, sy0, sy1 + 1, bx0, bx1, fast_stage, bar);
    __syncthreads();        // s_yo / s_yc

    // The four taps of this column as byte offsets into a staged row, replicate border applied once;
    // a source row index is clamped when it is used (only the image's first/last rows ever clamp).
    const int o0 = 3 * clampi(ux0 + cx, 0, a.sw - 1) - bx0, o1 = 3 * clampi(ux0 + cx + 1, 0, a.sw - 1) - bx0;
    const int o2 = 3 * clampi(ux0 + cx + 2, 0, a.sw - 1) - bx0, o3 = 3 * clampi(ux0 + cx + 3, 0, a.sw - 1) - bx0;
    const int ylo = sy0 - uy0, yhi = sy1 - uy0;       // window rows [ylo, yhi] exist in the staged tile

    // horizontal pass of one source row for this column: exact int32, returned as float32 (|H| < 2^22)
    auto hrow = [&](int rr, float (&h)[3]) {
        const uint8_t* row = s_raw + (min(max(rr, ylo), yhi) - ylo) * RP;
        const uint8_t *p0 = row + o0, *p1 = row + o1, *p2 = row + o2, *p3 = row + o3;
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
            const int v = (int)p0[ch] * kx.x + (int)p1[ch] * kx.y + (int)p2[ch] * kx.z + (int)p3[ch] * kx.w;
            h[ch] = __int_as_float(v + 0x4B400000) - 12582912.0f;
        }
    };

    const int vh = dyb - y0 + 1;
    const int gx = x0 + tid;
    const bool col_ok = gx < a.dw;
    const int lane = tid & 31, q = lane & 3;
    // word path: the whole quad of columns is inside the image and the row addresses are 4-byte aligned
    const bool quad_ok = ((gx | 3) < a.dw) && ((((uintptr_t)a.dst) | a.dpitch) & 3) == 0;
    float w0[3], w1[3], w2[3], w3[3];
    int yo = s_yo[0];
    hrow(yo, w0); hrow(yo + 1, w1); hrow(yo + 2, w2); hrow(yo + 3, w3);
    // one finished pixel of this column -> memory (quad repack through a shuffle: every lane of the warp must call it)
    auto emit = [&](int y, const uint32_t (&v)[3]) {
        const uint32_t px = v[0] | (v[1] << 8) | (v[2] << 16);
        const uint32_t nxt = __shfl_down_sync(0xffffffffu, px, 1);
        if (y >= vh) return;
        uint8_t* row = a.dst + (size_t)(y0 + y) * a.dpitch;
        if (quad_ok) {
            if (q != 3) {
                const uint32_t lo = px | (nxt << 24), hi = nxt >> 8;
                *reinterpret_cast<uint32_t*>(row + (size_t)(gx - q) * 3 + 4 * q) = __funnelshift_r(lo, hi, 8 * q);
            }
        } else if (col_ok) {
            uint8_t* o = row + (size_t)gx * 3;
            o[0] = (uint8_t)v[0]; o[1] = (uint8_t)v[1]; o[2] = (uint8_t)v[2];
        }
    };
    // 16-byte aligned frames: pixels are staged in shared memory and each tile row leaves as one TMA bulk store;
    // otherwise the quad-shuffle path
    uint8_t* s_out = smem + L.out;
    const bool staged = aligned16(a.dst, a.dpitch);
    auto put = [&](int y, const uint32_t (&v)[3]) {
        if (staged) {
            if (y < vh) { uint8_t* o = s_out + y * (BC_TW * 3) + tid * 3; o[0] = (uint8_t)v[0]; o[1] = (uint8_t)v[1]; o[2] = (uint8_t)v[2]; }
        } else emit(y, v);
    };
    auto flush_staged = [&]() {
        if (!staged) return;
        fence_proxy_async();
        __syncthreads();
        const int vw = dxb - x0 + 1, rb = vw * 3, nb16 = rb & ~15;
        if (tid < vh && nb16) {
            bulk_s2g(a.dst + (size_t)(y0 + tid) * a.dpitch + (size_t)x0 * 3, s_out + tid * (BC_TW * 3), (uint32_t)nb16);
            bulk_commit();
        }
        const int tb = rb - nb16;
        for (int i = tid; i < tb * vh; i += BC_NT) {
            const int y = i / tb, qb = nb16 + (i - y * tb);
            a.dst[(size_t)(y0 + y) * a.dpitch + (size_t)x0 * 3 + qb] = s_out[y * (BC_TW * 3) + qb];
        }
        if (tid < vh && nb16) bulk_wait_read0();
    };
    if constexpr (S > 0) {
        // rows y0 + y with y % S == j use coefficient set j (y0 is a multiple of BC_TH, hence of S)
        float4 bj[S];
#pragma unroll
        for (int j = 0; j < S; ++j) bj[j] = s_yc[j];
        // the four taps are 12 consecutive bytes unless the column touches the replicate border
        const bool contig = (o1 == o0 + 3) && (o2 == o0 + 6) && (o3 == o0 + 9);
        auto hrow_s = [&](int rr, float (&h)[3]) {
            const uint8_t* row = s_raw + (min(max(rr, ylo), yhi) - ylo) * RP;
            if (contig) {
                const uint8_t* p = row + o0;
#pragma unroll
                for (int ch = 0; ch < 3; ++ch) {
                    const int v = (int)p[ch] * kx.x + (int)p[3 + ch] * kx.y + (int)p[6 + ch] * kx.z + (int)p[9 + ch] * kx.w;
                    h[ch] = __int_as_float(v + 0x4B400000) - 12582912.0f;
                }
            } else {
                const uint8_t *p0 = row + o0, *p1 = row + o1, *p2 = row + o2, *p3 = row + o3;
#pragma unroll
                for (int ch = 0; ch < 3; ++ch) {
                    const int v = (int)p0[ch] * kx.x + (int)p1[ch] * kx.y + (int)p2[ch] * kx.z + (int)p3[ch] * kx.w;
                    h[ch] = __int_as_float(v + 0x4B400000) - 12582912.0f;
                }
            }
        };
        // window rows live in W[0..3] and rotate by index (the group loop is unrolled by four): no register moves
        float W[4][3];
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) { W[0][ch] = w0[ch]; W[1][ch] = w1[ch]; W[2][ch] = w2[ch]; W[3][ch] = w3[ch]; }
        auto row1 = [&](int y, const float (&r0)[3], const float (&r1)[3], const float (&r2)[3], const float (&r3)[3], const float4& b) {
            uint32_t v[3];
#pragma unroll
            for (int ch = 0; ch < 3; ++ch)
                v[ch] = f2u8(__fadd_rn(__fmaf_rn(r0[ch], b.x, __fmul_rn(r1[ch], b.y)), __fmaf_rn(r2[ch], b.z, __fmul_rn(r3[ch], b.w))));
            put(y, v);
        };
        // two rows that share the window: the same four products per channel as row1, as f32x2 lanes (row A, row B)
        auto row2 = [&](int y, const float (&r0)[3], const float (&r1)[3], const float (&r2)[3], const float (&r3)[3], const float4& ba, const float4& bb) {
            const unsigned long long bx = pack2(__float_as_uint(ba.x), __float_as_uint(bb.x)), by = pack2(__float_as_uint(ba.y), __float_as_uint(bb.y));
            const unsigned long long bz = pack2(__float_as_uint(ba.z), __float_as_uint(bb.z)), bw = pack2(__float_as_uint(ba.w), __float_as_uint(bb.w));
            uint32_t va[3], vb[3];
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
                const unsigned long long p0 = pack2(__float_as_uint(r0[ch]), __float_as_uint(r0[ch])), p1 = pack2(__float_as_uint(r1[ch]), __float_as_uint(r1[ch]));
                const unsigned long long p2 = pack2(__float_as_uint(r2[ch]), __float_as_uint(r2[ch])), p3 = pack2(__float_as_uint(r3[ch]), __float_as_uint(r3[ch]));
                const unsigned long long t = fadd2(ffma2(p0, bx, fmul2(p1, by)), ffma2(p2, bz, fmul2(p3, bw)));
                float ta, tb;
                unpack2(t, ta, tb);
                va[ch] = f2u8(ta); vb[ch] = f2u8(tb);
            }
            put(y, va);
            put(y + 1, vb);
        };
        // rows before the first step (phases 0 .. S/2-1) share the initial window
        if constexpr (S == 2) row1(0, W[0], W[1], W[2], W[3], bj[0]);
        else row2(0, W[0], W[1], W[2], W[3], bj[0], bj[1]);
        for (int yb = S / 2; yb < vh; yb += 4 * S) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {                 // step k overwrites the oldest row, W[k]; the window is W[k+1 .. k+4] mod 4
                const int y = yb + k * S;
                yo += 1;
                hrow_s(yo + 3, W[k]);
                if constexpr (S == 2) row2(y, W[(k + 1) & 3], W[(k + 2) & 3], W[(k + 3) & 3], W[k], bj[1], bj[0]);
                else {
                    row2(y, W[(k + 1) & 3], W[(k + 2) & 3], W[(k + 3) & 3], W[k], bj[2], bj[3]);
                    row2(y + 2, W[(k + 1) & 3], W[(k + 2) & 3], W[(k + 3) & 3], W[k], bj[0], bj[1]);
                }
            }
        }
        flush_staged();
        return;
    }
    for (int y = 0; y < vh; ++y) {
        const int yn = s_yo[y];           // uniform across the CTA: no divergence below
        const int d = yn - yo;
        if (d == 1) {                     // the only non-zero step an upscale produces
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) { w0[ch] = w1[ch]; w1[ch] = w2[ch]; w2[ch] = w3[ch]; }
            hrow(yn + 3, w3);
        } else if (d != 0) {
            hrow(yn, w0); hrow(yn + 1, w1); hrow(yn + 2, w2); hrow(yn + 3, w3);
        }
        yo = yn;
        const float4 b = s_yc[y];
        uint32_t v[3];
#pragma unroll
        for (int ch = 0; ch < 3; ++ch)
            v[ch] = f2u8(__fadd_rn(__fmaf_rn(w0[ch], b.x, __fmul_rn(w1[ch], b.y)),
                                   __fmaf_rn(w2[ch], b.z, __fmul_rn(w3[ch], b.w))));
        put(y, v);
    }
    flush_staged();
}

}  // namespace vp
