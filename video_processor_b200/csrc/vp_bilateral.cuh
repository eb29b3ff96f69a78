// vp_bilateral.cuh - cv::bilateralFilter on CV_8UC3 (reference call site
// src/selective_bilateral.cpp:110, SelectiveBilateral::applyBilateralFilter :84-119) with the
// device-side calculateAdaptiveParams selection (:315-371) and an optional fused temporal-blend
// epilogue (src/main.cpp:269-292 / TemporalConsistency::blendFrames, src/temporal_consistency.cpp:355-444).
//
// Arithmetic (OpenCV bilateral_filter, restated in oracle/restate.py:bilateral_u8c3):
//   taps (i,j) with i*i+j*j <= r*r in i-major order, BORDER_REFLECT_101;
//   w = space_w[tap] * color_w[|db|+|dg|+|dr|]  (float32 product), wsum += w, sum_c = fma(p_c, w, sum_c);
//   dst_c = RNE(sum_c * (1.f / wsum)).
// Layout: a CTA owns a TW x TH output tile.  The halo tile is staged from HBM by TMA bulk row copies
// into raw bytes, expanded once to packed BGR1 words with the border map resolved, and each thread
// then filters a strip of PX horizontally adjacent pixels so a staged neighbour is reused by up to
// PX outputs from registers (VABSDIFF4 + IDP.4A for the L1 colour distance).  The tap rows run in a
// rolled loop whose body is specialised on the row's half-width (a handful of variants per radius),
// which keeps the unrolled code inside the instruction cache; u8->f32 uses the exact PRMT "magic
// number" form (0x4B0000xx - 2^23) instead of I2F so the conversion pipe is not the limiter.
#pragma once
#include "vp_common.cuh"
#include "vp_scene.cuh"
#include "vp_stats.cuh"

namespace vp {

constexpr int BIL_TW = 64, BIL_TH = 16, BIL_PX = 4, BIL_NT = 256, BIL_RMAX = 7;

// Weight classes of a small window: taps at the same squared distance share their space weight, so for radius <= 3
// the host folds it into the colour LUT - class_w[c][n] = fl(space_w(d2) * color_w[n]), the very product the filter
// would form - and the tap loop drops its multiply.  d2 -> class index; the centre (d2 = 0) needs no table.
constexpr int BIL_CLS_RMAX = 3, BIL_NCLS = 6;          // radius 3: d2 in {1, 2, 4, 5, 8, 9}
__host__ __device__ constexpr int bil_class_of(int d2) {
    return d2 == 1 ? 0 : d2 == 2 ? 1 : d2 == 4 ? 2 : d2 == 5 ? 3 : d2 == 8 ? 4 : d2 == 9 ? 5 : -1;
}
__host__ __device__ constexpr int bil_nclasses(int r) { return r == 1 ? 1 : r == 2 ? 3 : r == 3 ? 6 : 0; }

struct __align__(16) BilHead {     // what every radius needs; a CTA fetches it into static shared memory with one bulk copy
    float color_w[768];
    float space_w[228];            // [(i+r)*(2r+1) + (j+r)], (2*7+1)^2 = 225 used
    int xmax[16];                  // half-width of tap row i (index i + r): floor(sqrt(r^2 - i^2))
    int radius;
    int d;
    int pad_[2];
};
struct __align__(16) BilCand {     // one (d, sigma) parameter set, device memory
    BilHead head;
    float class_w[BIL_NCLS][768];  // radius <= BIL_CLS_RMAX only; the first bil_nclasses(r) tables go to dynamic shared memory
};
static_assert(sizeof(BilHead) % 16 == 0, "bulk copies move whole 16-byte units");

struct BlendArgs {
    int mode;                      // VP_TEMPORAL_*
    int nprev;                     // previous frames available (0 = pass-through), at most 3
    const uint8_t* prev[3];        // prev[0] = t-1, prev[1] = t-2, prev[2] = t-3 (unblended states)
    size_t ppitch;
    uint8_t* state;                // where the unblended current frame is stored (may be null)
    size_t spitch;
    float w[4];                    // main blend: w[0]=alpha, w[1]=beta; blend_frames: per-frame weights
    float inv_wsum;                // blend_frames: 1.f / (w0+w1(+w2(+w3))) as accumulated in float32
};

struct BilArgs {
    const uint8_t* src; size_t spitch;
    uint8_t* dst; size_t dpitch;   // dst may be null when only `blend.state` is wanted
    int w, h;
    const BilCand* cand;           // candidates indexed by noise class (or 1 when sel == null)
    const int* sel;                // noise class chosen by the stats producer's last CTA (may be null)
    int radius[3];                 // radius of each candidate (host copy, saves a dependent global load)
    int ns;                        // strips per thread: the CTA tile is BIL_TW x (BIL_TH * ns)
    // CPUImpl::enhanceBicubicResult (src/upscaler.cpp:112-147): where the 2x2-dilated Canny edge map is set
    // the output is the 5-point sharpened source pixel (interior pixels) instead of the filtered one
    const uint8_t* emap; size_t epitch;
    // detectSceneChange accumulators (src/temporal_consistency.cpp:283-323): 64-bin histogram of the gray output and
    // sum |gray(out) - gray(previous unblended frame = blend.prev[0])|; null = off
    SceneAccum* scene;
    BlendArgs blend;
    // 16-byte aligned frames: the halo tile of candidate i arrives as one tensor-map box (bil_raw_pitch(r_i) - 16 bytes wide,
    // BIL_TH * ns + 2 r_i rows) instead of one bulk copy per row
    int use_tm;
    // src_bgrx: `src` holds one 32-bit word per pixel (B, G, R, 1) - k_upsharp's out_bgrx form.  The halo tile then arrives by
    // TMA directly as the tile words (box TP x TROWS of 32-bit units at (x0 - r, y0 - r), zero-filled outside the image);
    // only tiles on the image border run a pass that mirrors the outside positions.  Requires use_tm.
    int src_bgrx;
    int dst_bgrx;                  // `dst` takes 32-bit (B, G, R, 1) pixels; no blend / state / scene / edge map (k_bilateral<., DSTX>)
    alignas(64) CUtensorMap tm[3];
};

// exact u8 -> f32: bytes [b,0,0,0x4B] are the float 2^23 + b
__device__ __forceinline__ float u8f(uint32_t word, int byte) {
    return __uint_as_float(__byte_perm(word, 0x4B000000u, 0x7440 + byte)) - 8388608.0f;
}
// cv::saturate_cast<uchar>(float): one cvt (round-to-nearest-even, clamped to [0,255])
__device__ __forceinline__ uint32_t f2u8(float v) {
    uint32_t r;
    asm("cvt.rni.u8.f32 %0, %1;" : "=r"(r) : "f"(v));
    return r;
}

// packed f32x2 helpers (sm_100 FFMA2 / FADD2): two float lanes per 64-bit register pair, one issue slot
__device__ __forceinline__ unsigned long long pack2(uint32_t lo, uint32_t hi) {
    unsigned long long r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "r"(lo), "r"(hi)); return r;
}
__device__ __forceinline__ void unpack2(unsigned long long v, float& lo, float& hi) {
    asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d;
}
__device__ __forceinline__ unsigned long long fmul2(unsigned long long a, unsigned long long b) {
    unsigned long long d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d;
}
__device__ __forceinline__ unsigned long long fadd2_rm(unsigned long long a, unsigned long long b) {     // rounded toward -inf
    unsigned long long d; asm("add.rm.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d;
}
__device__ __forceinline__ unsigned long long fadd2(unsigned long long a, unsigned long long b) {
    unsigned long long d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d;
}

__device__ __forceinline__ uint32_t blend_byte(const BlendArgs& b, float cur, float p0, float p1, float p2) {
    if (b.mode == 1)               // cv::addWeighted AVX2 form: fma(cur, alpha, prev*beta)
        return f2u8(__fmaf_rn(cur, b.w[0], __fmul_rn(p0, b.w[1])));
    float acc = __fmul_rn(b.w[0], cur);
    acc = __fadd_rn(acc, __fmul_rn(b.w[1], p0));
    if (b.nprev > 1) acc = __fadd_rn(acc, __fmul_rn(b.w[2], p1));
    if (b.nprev > 2) acc = __fadd_rn(acc, __fmul_rn(b.w[3], p2));
    return f2u8(__fmul_rn(acc, b.inv_wsum));
}
// bytes k, k+1 of a word as an exact float32 pair (PRMT magic numbers, one packed subtract of 2^23)
__device__ __forceinline__ unsigned long long u8f_pair(uint32_t word, int k) {
    return fadd2(pack2(__byte_perm(word, 0x4B000000u, 0x7440 + k), __byte_perm(word, 0x4B000000u, 0x7441 + k)),
                 pack2(0xCB000000u, 0xCB000000u));
}
// blend_byte for two bytes at once: the same IEEE operations per lane (multiplies and adds stay separate)
__device__ __forceinline__ unsigned long long blend_pair(const BlendArgs& b, unsigned long long cur, unsigned long long p0,
                                                         unsigned long long p1, unsigned long long p2) {
    const uint32_t w0 = __float_as_uint(b.w[0]), w1 = __float_as_uint(b.w[1]);
    if (b.mode == 1) return ffma2(cur, pack2(w0, w0), fmul2(p0, pack2(w1, w1)));
    unsigned long long acc = fmul2(pack2(w0, w0), cur);
    acc = fadd2(acc, fmul2(pack2(w1, w1), p0));
    if (b.nprev > 1) { const uint32_t w2 = __float_as_uint(b.w[2]); acc = fadd2(acc, fmul2(pack2(w2, w2), p1)); }
    if (b.nprev > 2) { const uint32_t w3 = __float_as_uint(b.w[3]); acc = fadd2(acc, fmul2(pack2(w3, w3), p2)); }
    const uint32_t iw = __float_as_uint(b.inv_wsum);
    return fmul2(acc, pack2(iw, iw));
}
__device__ __forceinline__ uint32_t blend_word(const BlendArgs& b, uint32_t cur, uint32_t p0, uint32_t p1, uint32_t p2) {
    uint32_t r = 0;
#pragma unroll
    for (int k = 0; k < 4; k += 2) {
        float lo, hi;
        unpack2(blend_pair(b, u8f_pair(cur, k), u8f_pair(p0, k), u8f_pair(p1, k), u8f_pair(p2, k)), lo, hi);
        r |= (f2u8(lo) | (f2u8(hi) << 8)) << (8 * k);
    }
    return r;
}
// the same with the blend form (MODE 1 = addWeighted, 2 = blendFrames) and the frame count fixed at compile time:
// the epilogue of k_bilateral dispatches once per strip instead of testing nprev for every byte pair
template <int MODE, int NP>
__device__ __forceinline__ void blend12(const BlendArgs& b, const uint32_t (&cur)[3], const uint32_t (&p0)[3], const uint32_t (&p1)[3],
                                        const uint32_t (&p2)[3], uint32_t (&o)[3]) {
    const uint32_t w0 = __float_as_uint(b.w[0]), w1 = __float_as_uint(b.w[1]), w2 = __float_as_uint(b.w[2]), w3 = __float_as_uint(b.w[3]);
    const uint32_t iw = __float_as_uint(b.inv_wsum);
    const unsigned long long W0 = pack2(w0, w0), W1 = pack2(w1, w1), W2 = pack2(w2, w2), W3 = pack2(w3, w3), IW = pack2(iw, iw);
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        uint32_t r = 0;
#pragma unroll
        for (int k = 0; k < 4; k += 2) {
            unsigned long long acc;
            if (MODE == 1) acc = ffma2(u8f_pair(cur[q], k), W0, fmul2(u8f_pair(p0[q], k), W1));
            else {
                acc = fmul2(W0, u8f_pair(cur[q], k));
                acc = fadd2(acc, fmul2(W1, u8f_pair(p0[q], k)));
                if (NP > 1) acc = fadd2(acc, fmul2(W2, u8f_pair(p1[q], k)));
                if (NP > 2) acc = fadd2(acc, fmul2(W3, u8f_pair(p2[q], k)));
                acc = fmul2(acc, IW);
            }
            float lo, hi;
            unpack2(acc, lo, hi);
            r |= (f2u8(lo) | (f2u8(hi) << 8)) << (8 * k);
        }
        o[q] = r;
    }
}

// 12 packed bytes (4 BGR pixels) <-> memory; word accesses when the address is 4-byte aligned
__device__ __forceinline__ void store12(uint8_t* p, const uint32_t v[3], int npx) {
    if (npx == 4 && (((uintptr_t)p) & 3) == 0) {
        uint32_t* q = reinterpret_cast<uint32_t*>(p);
        q[0] = v[0]; q[1] = v[1]; q[2] = v[2];
    } else {
#pragma unroll
        for (int k = 0; k < 12; ++k)
            if (k < 3 * npx) p[k] = (uint8_t)(v[k >> 2] >> (8 * (k & 3)));
    }
}
__device__ __forceinline__ void load12(const uint8_t* p, uint32_t v[3], int npx) {
    if (npx == 4 && (((uintptr_t)p) & 3) == 0) {
        const uint32_t* q = reinterpret_cast<const uint32_t*>(p);
        v[0] = __ldg(q); v[1] = __ldg(q + 1); v[2] = __ldg(q + 2);
    } else {
        v[0] = v[1] = v[2] = 0;
#pragma unroll
        for (int k = 0; k < 12; ++k)
            if (k < 3 * npx) v[k >> 2] |= (uint32_t)p[k] << (8 * (k & 3));
    }
}

// XO: physical column of the tile's first logical column (x0 - R).  0 when the kernel builds the tile itself; in the
// 32-bit-source form the tile arrives by TMA, whose box must start on a 16-byte boundary of the source row: the box starts
// PAD = (R + 3) & ~3 pixels left of x0 and the logical columns sit XO = PAD - R words into each tile row.
__host__ __device__ constexpr int bil_pad(int r) { return (r + 3) & ~3; }
__host__ __device__ constexpr int bil_xo(int r) { return bil_pad(r) - r; }
template <int R, int XO = 0> struct BilGeom {
    static constexpr int TP = XO ? BIL_TW + 2 * (R + XO) : (BIL_TW + 2 * R + 3) & ~3;   // tile pitch in pixels (16-byte rows)
    static constexpr int NSEG = (BIL_PX + 2 * R + XO + 3) & ~3;                         // staged words per row, whole uint4s
    static constexpr int D = 2 * R + 1;
};

// one tap row of half-width XM: taps dx = -XM..XM against the PX outputs of the strip.
// Accumulators are (sum_b, sum_g) and (sum_r, wsum) pairs: the tile's 4th byte is 1, so the pixel's
// (r, 1.0) pair turns wsum += w into the second half of one FFMA2 (fma(1, w, wsum) == wsum + w).
// DY is the row's offset from the window centre when the caller knows it at compile time (small windows), else
// BIL_DY_ANY.  Two kinds of tap skip the per-pixel colour-distance lookup:
//   * the centre tap (colour distance 0 by construction): its weight space_w(0) * color_w[0] is formed once per row call;
//   * with class tables (R <= BIL_CLS_RMAX, DY known) the weight is class_w[class(DY^2 + dx^2)][n], the host-made product.
constexpr int BIL_DY_ANY = 99;
#ifndef VP_BIL_I2F
#define VP_BIL_I2F 1
#endif
#ifndef VP_BIL_PREFETCH
#define VP_BIL_PREFETCH 2      // previous frames of the temporal blend requested before the tap loop (0, 1 or 2 of them)
#endif
#ifndef VP_BIL_MINB
#define VP_BIL_MINB 4          // resident CTAs per SM the register allocation is held to (256 threads each: 64 registers)
#endif
template <int R, int XM, int DY, int XO>
__device__ __forceinline__ void bil_row(const uint32_t (&seg)[BilGeom<R, XO>::NSEG], const BilHead& T, const float* __restrict__ s_cls, int row,
                                        const uint32_t (&ctr)[BIL_PX],
                                        unsigned long long (&abg)[BIL_PX], unsigned long long (&arw)[BIL_PX]) {
    constexpr int N = BIL_PX + 2 * XM;                        // neighbours this row touches
    constexpr bool CENTER_ROW = (DY == 0) || (DY == BIL_DY_ANY && XM == R);      // only row 0 is R taps wide
    constexpr bool CLS = (R <= BIL_CLS_RMAX) && (DY != BIL_DY_ANY);
    const float* __restrict__ sw_row = T.space_w + row * BilGeom<R, XO>::D;
    unsigned long long pbg[N], pra[N];
    const unsigned long long bias = pack2(0xCB000000u, 0xCB000000u);   // (-2^23, -2^23)
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const uint32_t p = seg[i + R - XM + XO];
#if VP_BIL_I2F >= 1
        // (b, g) through the conversion unit (I2F.U8 with a byte selector, one instruction per value on the otherwise
        // idle XU pipe) instead of two PRMTs and half a FADD2 on the ALU / FMA pipes
        pbg[i] = pack2(__float_as_uint((float)(p & 0xffu)), __float_as_uint((float)((p >> 8) & 0xffu)));
#else
        pbg[i] = fadd2(pack2(__byte_perm(p, 0x4B000000u, 0x7440), __byte_perm(p, 0x4B000000u, 0x7441)), bias);
#endif
#if VP_BIL_I2F >= 2
        pra[i] = pack2(__float_as_uint((float)((p >> 16) & 0xffu)), 0x3f800000u);
#else
        pra[i] = fadd2(pack2(__byte_perm(p, 0x4B000000u, 0x7442), __byte_perm(p, 0x4B000000u, 0x7443)), bias);
#endif
    }
#pragma unroll
    for (int dx = -XM; dx <= XM; ++dx) {
        if (CENTER_ROW && dx == 0) {
            const uint32_t wgt = __float_as_uint(__fmul_rn(sw_row[R], T.color_w[0]));
            const unsigned long long w2 = pack2(wgt, wgt);
#pragma unroll
            for (int j = 0; j < BIL_PX; ++j) {
                abg[j] = ffma2(pbg[j + XM], w2, abg[j]);
                arw[j] = ffma2(pra[j + XM], w2, arw[j]);
            }
            continue;
        }
        const float sw = CLS ? 0.f : sw_row[dx + R];
        const float* __restrict__ lut = CLS ? s_cls + 768 * bil_class_of(CLS ? DY * DY + dx * dx : 1) : T.color_w;
#pragma unroll
        for (int j = 0; j < BIL_PX; ++j) {
            const int i = j + dx + XM;
            const unsigned n = __vsadu4(seg[i + R - XM + XO], ctr[j]);   // VABSDIFF4.ACC: |db|+|dg|+|dr| (+|1-1|)
            const uint32_t wgt = __float_as_uint(CLS ? lut[n] : __fmul_rn(sw, lut[n]));
            const unsigned long long w2 = pack2(wgt, wgt);
            abg[j] = ffma2(pbg[i], w2, abg[j]);
            arw[j] = ffma2(pra[i], w2, arw[j]);
        }
    }
}

template <int R, int XM, int XO>
struct BilRowSwitch {
    __device__ __forceinline__ static void run(int xm, const uint32_t (&seg)[BilGeom<R, XO>::NSEG], const BilHead& T, int row,
                                               const uint32_t (&ctr)[BIL_PX], unsigned long long (&abg)[BIL_PX],
                                               unsigned long long (&arw)[BIL_PX]) {
        if (xm == XM) bil_row<R, XM, BIL_DY_ANY, XO>(seg, T, nullptr, row, ctr, abg, arw);
        else BilRowSwitch<R, XM - 1, XO>::run(xm, seg, T, row, ctr, abg, arw);
    }
};
template <int R, int XO>
struct BilRowSwitch<R, -1, XO> {
    __device__ __forceinline__ static void run(int, const uint32_t (&)[BilGeom<R, XO>::NSEG], const BilHead&, int,
                                               const uint32_t (&)[BIL_PX], unsigned long long (&)[BIL_PX],
                                               unsigned long long (&)[BIL_PX]) {}
};

// half-width of tap row i of a radius-R window: taps (i,j) with i*i + j*j <= R*R  (the host's BilCand::xmax)
__host__ __device__ constexpr int bil_xmax(int R, int i) {
    int x = 0;
    while ((x + 1) * (x + 1) + i * i <= R * R) ++x;
    return x;
}

// small windows: every tap row unrolled with its compile-time half-width (no row switch, no xmax loads)
template <int R, int ROW, int XO>
__device__ __forceinline__ void bil_rows_unrolled(const uint32_t* __restrict__ s_tile, const BilHead& T, const float* __restrict__ s_cls, int lx, int ly,
                                                  const uint32_t (&ctr)[BIL_PX],
                                                  unsigned long long (&abg)[BIL_PX], unsigned long long (&arw)[BIL_PX]) {
    if constexpr (ROW <= 2 * R) {
        constexpr int TP = BilGeom<R, XO>::TP, NSEG = BilGeom<R, XO>::NSEG;
        const uint4* rowp = reinterpret_cast<const uint4*>(s_tile + (ly + ROW) * TP + lx);
        uint32_t seg[NSEG];
#pragma unroll
        for (int q = 0; q < NSEG / 4; ++q) {
            const uint4 v = rowp[q];
            seg[4 * q] = v.x; seg[4 * q + 1] = v.y; seg[4 * q + 2] = v.z; seg[4 * q + 3] = v.w;
        }
        bil_row<R, bil_xmax(R, ROW - R), ROW - R, XO>(seg, T, s_cls, ROW, ctr, abg, arw);
        bil_rows_unrolled<R, ROW + 1, XO>(s_tile, T, s_cls, lx, ly, ctr, abg, arw);
    }
}

constexpr int BIL_UNROLL_RMAX = 3;
static_assert(BIL_CLS_RMAX <= BIL_UNROLL_RMAX, "class tables need compile-time tap rows");

template <int R, int XO>
__device__ __forceinline__ void bil_strip(const uint32_t* __restrict__ s_tile, const BilHead& T, const float* __restrict__ s_cls, int lx, int ly,
                                          uint32_t out[BIL_PX]) {
    constexpr int TP = BilGeom<R, XO>::TP, NSEG = BilGeom<R, XO>::NSEG, D = BilGeom<R, XO>::D;
    uint32_t ctr[BIL_PX];
    unsigned long long abg[BIL_PX], arw[BIL_PX];
#pragma unroll
    for (int j = 0; j < BIL_PX; ++j) {
        ctr[j] = s_tile[(ly + R) * TP + lx + R + XO + j];
        abg[j] = arw[j] = 0ull;
    }
    if constexpr (R <= BIL_UNROLL_RMAX) {
        bil_rows_unrolled<R, 0, XO>(s_tile, T, s_cls, lx, ly, ctr, abg, arw);
    } else {
#pragma unroll 1
        for (int row = 0; row < D; ++row) {
            const uint4* rowp = reinterpret_cast<const uint4*>(s_tile + (ly + row) * TP + lx);
            uint32_t seg[NSEG];
#pragma unroll
            for (int q = 0; q < NSEG / 4; ++q) {
                const uint4 v = rowp[q];
                seg[4 * q] = v.x; seg[4 * q + 1] = v.y; seg[4 * q + 2] = v.z; seg[4 * q + 3] = v.w;
            }
            BilRowSwitch<R, R, XO>::run(T.xmax[row], seg, T, row, ctr, abg, arw);
        }
    }
#pragma unroll
    for (int j = 0; j < BIL_PX; ++j) {
        float sb, sg, sr, ws;
        unpack2(abg[j], sb, sg);
        unpack2(arw[j], sr, ws);
        const float inv = __frcp_rn(ws);
        out[j] = f2u8(__fmul_rn(sb, inv)) | (f2u8(__fmul_rn(sg, inv)) << 8) | (f2u8(__fmul_rn(sr, inv)) << 16);
    }
}

// dynamic shared memory: [ tile words | raw bytes | mbarrier | class tables (radius <= BIL_CLS_RMAX) ]; the head of the
// parameter set (colour LUT, space weights, row half-widths) is static shared memory so its base folds into the LDS immediates
__host__ __device__ inline int bil_tile_pitch(int r) { return (BIL_TW + 2 * r + 3) & ~3; }
__host__ __device__ inline int bil_tile_pitch_x(int r) { return BIL_TW + 2 * bil_pad(r); }       // 32-bit-source form (>= bil_tile_pitch)
__host__ __device__ inline int bil_raw_pitch(int r) { return (((BIL_TW + 2 * r) * 3 + 15 + 15) & ~15) + 16; }
__host__ __device__ inline int bil_class_bytes(int r) { return r <= BIL_CLS_RMAX ? bil_nclasses(r) * 768 * 4 : 0; }
__host__ __device__ inline size_t bil_tile_bytes(int r, int trows) { return ((size_t)bil_tile_pitch_x(r) * trows * 4 + 127) & ~(size_t)127; }
__host__ inline size_t bil_smem_bytes(int r, int ns) {
    size_t tile = bil_tile_bytes(r, BIL_TH * ns + 2 * r);
    size_t raw = (size_t)bil_raw_pitch(r) * (BIL_TH * ns + 2 * r);
    return tile + raw + 16 + bil_class_bytes(r);
}

// BGRX: the source holds 32-bit (B, G, R, 1) pixels (BilArgs::src_bgrx) - compiled as its own kernel so neither form carries
// the other's registers
// DSTX: the output goes out as one 32-bit word (B, G, R, 1) per pixel and nothing else happens in the epilogue
// (BilArgs::dst_bgrx: the PRE stage feeding k_upsharp's 32-bit source window)
template <bool BGRX, bool DSTX = false>
__global__ void __launch_bounds__(BIL_NT, VP_BIL_MINB) k_bilateral(const __grid_constant__ BilArgs a) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ BilHead s_cand;
    const int tid = threadIdx.x;
    const int sel = a.sel ? *a.sel : 0;
    const BilCand* __restrict__ cand = a.cand + sel;
    const int R = a.radius[sel];
    const int TP = BGRX ? bil_tile_pitch_x(R) : bil_tile_pitch(R);
    const int TH = BIL_TH * a.ns;
    const int TROWS = TH + 2 * R;
    const bool tm = a.use_tm != 0;                          // implies 16-byte aligned frames
    const int RP = bil_raw_pitch(R) - (tm ? 16 : 0);        // a TMA box is dense

    uint32_t* s_tile = reinterpret_cast<uint32_t*>(smem);
    uint8_t* s_raw = smem + bil_tile_bytes(R, TROWS);
    uint64_t* bar = reinterpret_cast<uint64_t*>(s_raw + (size_t)bil_raw_pitch(R) * TROWS);
    float* s_cls = reinterpret_cast<float*>(bar + 2);

    const int x0 = blockIdx.x * BIL_TW, y0 = blockIdx.y * TH;
    // valid source window (image coordinates) the halo tile touches
    const int vx0 = max(x0 - R, 0), vx1 = min(x0 + BIL_TW + R, a.w);
    const int vy0 = max(y0 - R, 0), vy1 = min(y0 + TH + R, a.h);
    const int bx0 = (vx0 * 3) & ~15, bx1 = vx1 * 3;
    const bool fast = aligned16(a.src, a.spitch);

    // the halo tile and the parameter set arrive by bulk copies on one mbarrier
    if (BGRX) stage_box_issue(reinterpret_cast<uint8_t*>(s_tile), &a.tm[sel], x0 - bil_pad(R), y0 - R, (uint32_t)(TP * TROWS * 4), bar,
                              &s_cand, &cand->head, (uint32_t)sizeof(BilHead), s_cls, cand->class_w, (uint32_t)bil_class_bytes(R));
    else if (tm) stage_box_issue(s_raw, &a.tm[sel], bx0, vy0, (uint32_t)(RP * TROWS), bar, &s_cand, &cand->head, (uint32_t)sizeof(BilHead),
                            s_cls, cand->class_w, (uint32_t)bil_class_bytes(R));
    else stage_rows_issue(s_raw, RP, a.src, a.spitch, vy0, vy1, bx0, bx1, fast, bar, &s_cand, &cand->head, (uint32_t)sizeof(BilHead),
                          s_cls, cand->class_w, (uint32_t)bil_class_bytes(R));
    stage_rows_wait(s_raw, RP, a.src, a.spitch, 3 * a.w, vy0, vy1, bx0, bx1, fast, bar);

    // expand raw bytes -> BGR0 words with BORDER_REFLECT_101 resolved; warp per row, lanes along x
    const bool simple = (a.w > R) && (a.h > R);          // one mirror step suffices
    if (BGRX) {
        // the tile words are already in place; positions outside the image (zero-filled by the TMA unit) take their
        // BORDER_REFLECT_101 mirror, which lies inside the image and inside this tile (host: w, h > 2 r)
        if (x0 < R || y0 < R || x0 + BIL_TW + R > a.w || y0 + TH + R > a.h) {
            const int ox = x0 - bil_pad(R), oy = y0 - R;       // image position of tile word (0, 0)
            for (int i = tid; i < TROWS * TP; i += BIL_NT) {
                const int ty = i / TP, tx = i - ty * TP;
                const int sx = ox + tx, sy = oy + ty;
                if ((unsigned)sx < (unsigned)a.w && (unsigned)sy < (unsigned)a.h) continue;
                int mx = sx < 0 ? -sx : sx; mx = mx >= a.w ? 2 * a.w - 2 - mx : mx;
                int my = sy < 0 ? -sy : sy; my = my >= a.h ? 2 * a.h - 2 - my : my;
                const int qx = mx - ox, qy = my - oy;
                if ((unsigned)qx < (unsigned)TP && (unsigned)qy < (unsigned)TROWS && mx >= 0 && my >= 0)
                    s_tile[i] = s_tile[qy * TP + qx];       // sources are in-image positions: never rewritten by this pass
            }
        }
    } else if (fast && simple && x0 >= R && x0 + BIL_TW + R <= a.w) {
        // no horizontal mirroring in this tile: a lane turns 12 raw bytes (4 pixels, any byte phase) into four
        // BGR1 words with funnel shifts and byte permutes and stores them as one 128-bit word
        const int c0 = 3 * vx0 - bx0;                    // byte offset of tile column 0 inside a raw row, 0..15
        const int sh = 8 * (c0 & 3);
        // the (row, 4-pixel group) items are spread over all threads, not a warp per row: TP/4 is 17-20 groups
        const int ng = TP >> 2;
        const int dty = BIL_NT / ng, dg = BIL_NT - dty * ng;
        int ty = tid / ng, g = tid - ty * ng;
        for (; ty < TROWS; ty += dty) {
            int sy = y0 - R + ty;
            sy = sy < 0 ? -sy : sy; sy = sy >= a.h ? 2 * a.h - 2 - sy : sy;
            uint4 o = make_uint4(0x01000000u, 0x01000000u, 0x01000000u, 0x01000000u);
            if (sy >= vy0 && sy < vy1) {                 // false only for unused rows of bottom-edge tiles
                const uint32_t* q = reinterpret_cast<const uint32_t*>(s_raw + (sy - vy0) * RP + (c0 & ~3)) + 3 * g;
                const uint32_t r0 = q[0], r1 = q[1], r2 = q[2], r3 = q[3];
                const uint32_t w0 = __funnelshift_r(r0, r1, sh), w1 = __funnelshift_r(r1, r2, sh), w2 = __funnelshift_r(r2, r3, sh);
                o.x = __byte_perm(w0, 0x01010101u, 0x4210);
                o.y = (__byte_perm(w0, w1, 0x3543) & 0x00ffffffu) | 0x01000000u;
                o.z = (__byte_perm(w1, w2, 0x3432) & 0x00ffffffu) | 0x01000000u;
                o.w = __byte_perm(w2, 0x01010101u, 0x4321);
            }
            *reinterpret_cast<uint4*>(s_tile + ty * TP + 4 * g) = o;
            g += dg;
            if (g >= ng) { g -= ng; ++ty; }
        }
    } else {
        const int TWH = BIL_TW + 2 * R;
        const int lane = tid & 31, wrp = tid >> 5;
        for (int ty = wrp; ty < TROWS; ty += BIL_NT / 32) {
            int sy = y0 - R + ty;
            if (simple) { sy = sy < 0 ? -sy : sy; sy = sy >= a.h ? 2 * a.h - 2 - sy : sy; }
            else sy = reflect101(sy, a.h);
            const bool rowok = sy >= vy0 && sy < vy1;
            const uint8_t* rrow = s_raw + (sy - vy0) * RP - bx0;
            for (int tx = lane; tx < TWH; tx += 32) {
                int sx = x0 - R + tx;
                if (simple) { sx = sx < 0 ? -sx : sx; sx = sx >= a.w ? 2 * a.w - 2 - sx : sx; }
                else sx = reflect101(sx, a.w);
                uint32_t v = 0x01000000u;                    // 4th byte 1: the (r, 1.0) pair accumulates wsum
                if (rowok && sx >= vx0 && sx < vx1) {        // false only for unused positions of edge tiles
                    const uint8_t* p = rrow + sx * 3;
                    v |= (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16);
                }
                s_tile[ty * TP + tx] = v;
            }
        }
    }
    __syncthreads();

    const BlendArgs& b = a.blend;
    const bool doblend = a.dst && b.mode != 0 && b.nprev > 0;
    const bool scene_prev = a.scene && b.nprev > 0;          // previous frame needed for the gray difference
    __shared__ unsigned s_hist[64];
    __shared__ unsigned s_mad;                               // <= 2048 px * 255 per tile: fits 32 bits (native atomic)
    if (a.scene) {                                           // uniform across the grid
        if (tid < 64) s_hist[tid] = 0;
        if (tid == 0) s_mad = 0u;
        __syncthreads();
    }
    unsigned mad = 0;
    const int lx = (tid & 15) * BIL_PX;
    const int gx = x0 + lx;
    const int npx = min(BIL_PX, a.w - gx);
    const size_t boff = (size_t)gx * 3;
    for (int s = 0; s < a.ns && gx < a.w; ++s) {
        const int ly = (tid >> 4) + s * BIL_TH;
        const int gy = y0 + ly;
        if (gy >= a.h) break;
        // previous frames of the temporal blend: issue the loads before the tap loop hides their latency
        uint32_t p0[3] = {0, 0, 0}, p1[3] = {0, 0, 0}, p2[3] = {0, 0, 0};
        // (two of them in the 32-bit-source kernel; the general kernel has no registers to spare for them - measured: it
        // spills with any prefetch and is 1-2 % faster without - and fetches all three behind the tap loop)
        constexpr int PREF = BGRX ? VP_BIL_PREFETCH : 0;
        if (PREF > 0 && (doblend || scene_prev)) {
            load12(b.prev[0] + (size_t)gy * b.ppitch + boff, p0, npx);
            if (PREF > 1 && doblend && b.nprev > 1) load12(b.prev[1] + (size_t)gy * b.ppitch + boff, p1, npx);
        }
        uint32_t out[BIL_PX];
        const bool late_p1 = PREF < 2 && doblend && b.nprev > 1;
        switch (R) {
            case 1: bil_strip<1, BGRX ? bil_xo(1) : 0>(s_tile, s_cand, s_cls, lx, ly, out); break;
            case 2: bil_strip<2, BGRX ? bil_xo(2) : 0>(s_tile, s_cand, s_cls, lx, ly, out); break;
            case 3: bil_strip<3, BGRX ? bil_xo(3) : 0>(s_tile, s_cand, s_cls, lx, ly, out); break;
            case 4: bil_strip<4, BGRX ? bil_xo(4) : 0>(s_tile, s_cand, s_cls, lx, ly, out); break;
            case 5: bil_strip<5, BGRX ? bil_xo(5) : 0>(s_tile, s_cand, s_cls, lx, ly, out); break;
            case 6: bil_strip<6, BGRX ? bil_xo(6) : 0>(s_tile, s_cand, s_cls, lx, ly, out); break;
            default: bil_strip<7, BGRX ? bil_xo(7) : 0>(s_tile, s_cand, s_cls, lx, ly, out); break;
        }
        if (DSTX) {      // 4 words per strip: one 128-bit store (gx is a multiple of 4, base and pitch 16-byte aligned)
            uint32_t* dp = reinterpret_cast<uint32_t*>(a.dst + (size_t)gy * a.dpitch) + gx;
            if (npx == 4) *reinterpret_cast<uint4*>(dp) = make_uint4(out[0] | 0x01000000u, out[1] | 0x01000000u, out[2] | 0x01000000u, out[3] | 0x01000000u);
            else
                for (int j = 0; j < npx; ++j) dp[j] = out[j] | 0x01000000u;
            continue;
        }
        if (!BGRX && a.emap) {
            // 2x2 dilate (anchor (1,1)): pixel x is an edge if map[y-1..y][x-1..x] holds a 2.  The strip's four
            // map bytes of a row arrive as one aligned word (gx % 4 == 0, pitch % 4 == 0) plus the byte left of it.
            uint32_t m = 0;
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int y = gy - r;
                if (y < 0) continue;
                const uint8_t* row = a.emap + (size_t)y * a.epitch + gx;
                const uint32_t e = __vcmpeq4(__ldg(reinterpret_cast<const uint32_t*>(row)), 0x02020202u);   // 0xff per edge byte
                const uint32_t left = (gx > 0 && __ldg(row - 1) == 2) ? 0xffu : 0u;
                m |= e | (e << 8) | left;
            }
            if (m) {
                const int TPr = bil_tile_pitch(R);
#pragma unroll
                for (int j = 0; j < BIL_PX; ++j) {
                    const int x = gx + j;
                    if (!((m >> (8 * j)) & 0xff) || x >= a.w) continue;
                    const uint32_t* t = s_tile + (ly + R) * TPr + lx + R + j;
                    uint32_t v = t[0] & 0xffffffu;
                    if (x >= 1 && x < a.w - 1 && gy >= 1 && gy < a.h - 1) {
                        const uint32_t up = t[-TPr], dn = t[TPr], lf = t[-1], rt = t[1];
                        uint32_t r = 0;
#pragma unroll
                        for (int ch = 0; ch < 3; ++ch) {
                            const int sh = 8 * ch;
                            const int val = 5 * (int)((v >> sh) & 0xff) - (int)((up >> sh) & 0xff) - (int)((dn >> sh) & 0xff) -
                                            (int)((lf >> sh) & 0xff) - (int)((rt >> sh) & 0xff);
                            r |= (uint32_t)sat_u8(val) << sh;
                        }
                        v = r;
                    }
                    out[j] = v;
                }
            }
        }
        if (PREF < 1 && (doblend || scene_prev)) load12(b.prev[0] + (size_t)gy * b.ppitch + boff, p0, npx);
        if (late_p1) load12(b.prev[1] + (size_t)gy * b.ppitch + boff, p1, npx);
        // epilogue: optional temporal blend, state store, output store (12 contiguous bytes per strip)
        uint32_t cur[3];   // 12 packed bytes B G R B | G R B G | R B G R
        cur[0] = (out[0] & 0xffffffu) | (out[1] << 24);
        cur[1] = ((out[1] >> 8) & 0xffffu) | (out[2] << 16);
        cur[2] = ((out[2] >> 16) & 0xffu) | (out[3] << 8);
        if (b.state) store12(b.state + (size_t)gy * b.spitch + boff, cur, npx);
        if (a.scene) {
            // previous pixels of the strip, unpacked from the 12 packed bytes
            const uint32_t q[BIL_PX] = {p0[0] & 0xffffffu, (p0[0] >> 24) | ((p0[1] & 0xffffu) << 8),
                                        (p0[1] >> 16) | ((p0[2] & 0xffu) << 16), p0[2] >> 8};
#pragma unroll
            for (int j = 0; j < BIL_PX; ++j) {
                // one shared-memory atomic per distinct bin per warp: neighbouring pixels share bins
                const int g = j < npx ? gray_of(out[j]) : -4;
                const int bin = g >> 2;                                  // -1 for lanes without a pixel
                const unsigned same = __match_any_sync(__activemask(), bin);
                if (bin >= 0 && (int)(__ffs(same) - 1) == (tid & 31)) atomicAdd(&s_hist[bin], (unsigned)__popc(same));
                if (scene_prev && j < npx) mad += abs(g - gray_of(q[j]));
            }
        }
        if (!a.dst) continue;
        uint8_t* dp = a.dst + (size_t)gy * a.dpitch + boff;
        if (doblend) {
            // the third previous frame is fetched late: holding 9 prefetched words across the tap loop spills
            if (b.nprev > 2) load12(b.prev[2] + (size_t)gy * b.ppitch + boff, p2, npx);
            uint32_t o[3];
            if (b.mode == 1) blend12<1, 1>(b, cur, p0, p1, p2, o);
            else if (b.nprev == 1) blend12<2, 1>(b, cur, p0, p1, p2, o);
            else if (b.nprev == 2) blend12<2, 2>(b, cur, p0, p1, p2, o);
            else blend12<2, 3>(b, cur, p0, p1, p2, o);
            store12(dp, o, npx);
        } else {
            store12(dp, cur, npx);
        }
    }
    if (a.scene) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mad += __shfl_xor_sync(0xffffffffu, mad, o);
        if ((tid & 31) == 0 && mad) atomicAdd(&s_mad, mad);
        __syncthreads();
        const int copy = (blockIdx.x + blockIdx.y * gridDim.x) % SC_COPIES;
        if (tid < 64 && s_hist[tid]) atomicAdd(&a.scene->hist[copy][tid], s_hist[tid]);
        if (tid == 0 && s_mad) atomicAdd(&a.scene->mad, (unsigned long long)s_mad);
    }
}

}  // namespace vp
