// vp_bicubic_int.cuh - cv::resize(INTER_CUBIC) for CV_8UC3 at an integer ratio S = 2 or 4 in both directions
// (BASELINE configs C1 / C5; reference call sites src/upscaler.cpp:81, src/main.cpp:265).  Bit-exact to OpenCV like
// k_bicubic (vp_bicubic.cuh), which remains the kernel for every other ratio; this form spends about half the
// instructions per output pixel.
//
// What an integer ratio buys (all of it verified on the host against cv::resize's own tables before this kernel is
// chosen, vpb200.cu:ensure_resize): the S output columns S*j .. S*j+S-1 read only the five source columns j-2 .. j+2
// (tap origin j-2 for the first S/2 of them, j-1 for the rest) with S fixed coefficient sets, and likewise for rows.
//   P0  the tile's source window is staged by TMA bulk row copies (raw interleaved bytes);
//   P1  every window pixel becomes ONE float4 (B, G, R, 0) - exact PRMT conversions - in window coordinates with the
//       replicate border materialised, so nothing below clamps;
//   P2  a thread owns S adjacent output columns of one row band and walks down: per source row five 128-bit shared
//       loads feed the horizontal pass of all S columns.  The horizontal pass runs in float32 - every product and
//       partial sum is an integer below 2^22, so the FMAs are exact and equal cv::resize's int32 pass - with B and G of a
//       pixel as the two lanes of packed f32x2 FMAs (the tap pairs (k, k) come from the kernel parameters);
//   P3  vertical pass t = fma(S0,b0,S1*b1) + fma(S2,b2,S3*b3) exactly as cv::resize's float row filter, again B/G
//       packed and the R values of two neighbouring columns packed; RNE + saturate by one cvt per value;
//   P4  finished pixels are staged in shared memory and every tile row leaves as one TMA bulk store.
#pragma once
#include "vp_bicubic.cuh"

namespace vp {

constexpr int BI_TW = 128, BI_TH = 32, BI_NT = 128;

struct BicubicIntArgs {
    const uint8_t* src; size_t spitch; int sw, sh;
    uint8_t* dst; size_t dpitch; int dw, dh;
    float2 xk[4][4];     // [column phase][tap] as (k, k): 11-bit integer taps, exact in float32
    float2 yk[4][4];     // [row phase][tap] as (b, b): beta * 2^-22 as float32 (ResizeTables::ycoef)
};

template <int S> struct BicubicIntGeom {
    static constexpr int NCG = BI_TW / S;          // column groups per tile
    static constexpr int NB = BI_NT / NCG;         // row bands per tile
    static constexpr int RB = BI_TH / NB;          // output rows per band
    static constexpr int WC = BI_TW / S + 4;       // source window of a tile: columns x0/S - 2 .. x0/S + TW/S + 1
    static constexpr int WR = BI_TH / S + 4;
    static constexpr int RP = ((WC * 3 + 15 + 15) & ~15) + 16;     // raw row pitch (16 bytes of slack for the aligned start)
    // the raw bytes are dead once P1 has converted them; the staged output tile takes their place
    static constexpr int RAW = 0, OUT = 0;
    static constexpr int RAWB = RP * WR, OUTB = BI_TH * BI_TW * 3;
    static constexpr int F4 = ((RAWB > OUTB ? RAWB : OUTB) + 15) & ~15;
    static constexpr int BAR = F4 + WC * WR * 16;
    static constexpr int TOTAL = BAR + 16;
};

// horizontal pass of one source row for the S columns of a thread: h[c] = (B, G) lanes, hr[c] = R
template <int S>
__device__ __forceinline__ void bi_hrow(const float4* __restrict__ p, const BicubicIntArgs& a, unsigned long long (&hbg)[S], float (&hr)[S]) {
    float4 P[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) P[k] = p[k];
#pragma unroll
    for (int c = 0; c < S; ++c) {
        constexpr int half = S / 2;
        const int d = c >= half ? 1 : 0;
        unsigned long long acc = fmul2(pack2(__float_as_uint(P[d].x), __float_as_uint(P[d].y)), pack2(__float_as_uint(a.xk[c][0].x), __float_as_uint(a.xk[c][0].y)));
        float r = __fmul_rn(P[d].z, a.xk[c][0].x);
#pragma unroll
        for (int t = 1; t < 4; ++t) {
            acc = ffma2(pack2(__float_as_uint(P[d + t].x), __float_as_uint(P[d + t].y)), pack2(__float_as_uint(a.xk[c][t].x), __float_as_uint(a.xk[c][t].y)), acc);
            r = __fmaf_rn(P[d + t].z, a.xk[c][t].x, r);
        }
        hbg[c] = acc; hr[c] = r;
    }
}

// vertical pass of one output row (coefficient set `ph`) for the thread's S columns, window rows w0..w3 (oldest first)
template <int S>
__device__ __forceinline__ void bi_vrow(const BicubicIntArgs& a, int ph,
                                        const unsigned long long (&b0)[S], const float (&r0)[S], const unsigned long long (&b1)[S], const float (&r1)[S],
                                        const unsigned long long (&b2)[S], const float (&r2)[S], const unsigned long long (&b3)[S], const float (&r3)[S],
                                        uint8_t* __restrict__ o) {
    const unsigned long long k0 = pack2(__float_as_uint(a.yk[ph][0].x), __float_as_uint(a.yk[ph][0].y));
    const unsigned long long k1 = pack2(__float_as_uint(a.yk[ph][1].x), __float_as_uint(a.yk[ph][1].y));
    const unsigned long long k2 = pack2(__float_as_uint(a.yk[ph][2].x), __float_as_uint(a.yk[ph][2].y));
    const unsigned long long k3 = pack2(__float_as_uint(a.yk[ph][3].x), __float_as_uint(a.yk[ph][3].y));
    uint32_t px[S];
#pragma unroll
    for (int c = 0; c < S; ++c) {
        float tb, tg;
        unpack2(fadd2(ffma2(b0[c], k0, fmul2(b1[c], k1)), ffma2(b2[c], k2, fmul2(b3[c], k3))), tb, tg);
        px[c] = f2u8(tb) | (f2u8(tg) << 8);
    }
#pragma unroll
    for (int c = 0; c < S; c += 2) {     // R of two neighbouring columns as one f32x2
        float ta, tb;
        unpack2(fadd2(ffma2(pack2(__float_as_uint(r0[c]), __float_as_uint(r0[c + 1])), k0, fmul2(pack2(__float_as_uint(r1[c]), __float_as_uint(r1[c + 1])), k1)),
                      ffma2(pack2(__float_as_uint(r2[c]), __float_as_uint(r2[c + 1])), k2, fmul2(pack2(__float_as_uint(r3[c]), __float_as_uint(r3[c + 1])), k3))), ta, tb);
        px[c] |= f2u8(ta) << 16; px[c + 1] |= f2u8(tb) << 16;
    }
    if constexpr (S == 2) {              // 6 bytes at a 2-byte aligned address
        uint16_t* q = reinterpret_cast<uint16_t*>(o);
        q[0] = (uint16_t)px[0];
        q[1] = (uint16_t)((px[0] >> 16) | (px[1] << 8));
        q[2] = (uint16_t)(px[1] >> 8);
    } else {                             // 12 bytes at a 4-byte aligned address
        uint32_t* q = reinterpret_cast<uint32_t*>(o);
        q[0] = px[0] | (px[1] << 24);
        q[1] = (px[1] >> 8) | (px[2] << 16);
        q[2] = (px[2] >> 16) | (px[3] << 8);
    }
}

template <int S>
__global__ void __launch_bounds__(BI_NT) k_bicubic_int(const __grid_constant__ BicubicIntArgs a) {
    using G = BicubicIntGeom<S>;
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* s_raw = smem + G::RAW;
    float4* s_f4 = reinterpret_cast<float4*>(smem + G::F4);
    uint8_t* s_out = smem + G::OUT;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + G::BAR);
    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * BI_TW, y0 = blockIdx.y * BI_TH;
    // window origin in unclamped source coordinates: cv::resize's xofs[S j] = j - 2 at an integer ratio
    const int ux0 = x0 / S - 2, uy0 = y0 / S - 2;
    const int sx0 = clampi(ux0, 0, a.sw - 1), sx1 = clampi(ux0 + G::WC - 1, 0, a.sw - 1);
    const int sy0 = clampi(uy0, 0, a.sh - 1), sy1 = clampi(uy0 + G::WR - 1, 0, a.sh - 1);
    const int bx0 = (sx0 * 3) & ~15, bx1 = (sx1 + 1) * 3;
    const bool fast_stage = aligned16(a.src, a.spitch);
    stage_rows_issue(s_raw, G::RP, a.src, a.spitch, sy0, sy1 + 1, bx0, bx1, fast_stage, bar);
    stage_rows_wait(s_raw, G::RP, a.src, a.spitch, 3 * a.sw, sy0, sy1 + 1, bx0, bx1, fast_stage, bar);

    // ---- P1: raw bytes -> float4 (B, G, R, 0) per window pixel, replicate border materialised ----------------------
    {
        const unsigned long long bias = pack2(0xCB000000u, 0xCB000000u);       // (-2^23, -2^23)
        for (int i = tid; i < G::WC * G::WR; i += BI_NT) {
            const int rr = i / G::WC, cc = i - rr * G::WC;
            const uint8_t* p = s_raw + (clampi(uy0 + rr, 0, a.sh - 1) - sy0) * G::RP + 3 * clampi(ux0 + cc, 0, a.sw - 1) - bx0;
            const uint32_t w = (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16);
            float b, g;
            unpack2(fadd2(pack2(__byte_perm(w, 0x4B000000u, 0x7440), __byte_perm(w, 0x4B000000u, 0x7441)), bias), b, g);
            const float r = __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7442)) - 8388608.0f;
            s_f4[i] = make_float4(b, g, r, 0.0f);
        }
    }
    __syncthreads();

    // ---- P2-P3: S columns x RB rows per thread ------------------------------------------------------------------------
    const int j = tid % G::NCG, band = tid / G::NCG;
    const int yb = band * G::RB;                                  // first output row of the band inside the tile
    const float4* col = s_f4 + (yb / S) * G::WC + j;              // window row of the band's first tap origin, column j - 2
    uint8_t* o = s_out + yb * (BI_TW * 3) + j * (3 * S);
    unsigned long long hb[4][S];
    float hr[4][S];
#pragma unroll
    for (int k = 0; k < 4; ++k) bi_hrow<S>(col + k * G::WC, a, hb[k], hr[k]);
    // rows of phase 0 .. S/2-1 share the initial window
#pragma unroll
    for (int p = 0; p < S / 2; ++p) bi_vrow<S>(a, p, hb[0], hr[0], hb[1], hr[1], hb[2], hr[2], hb[3], hr[3], o + p * (BI_TW * 3));
    // every further source row serves S output rows (phases S/2 .. S-1, then 0 .. S/2-1); the band's last step only its
    // first S/2 (the rest belong to the next band).  The window rotates by index: the loop is unrolled by four.
    constexpr int NSTEP = G::RB / S;
#pragma unroll
    for (int s = 0; s < NSTEP; ++s) {
        const int k = s & 3;                                      // overwrites the oldest row
        bi_hrow<S>(col + (s + 4) * G::WC, a, hb[k], hr[k]);
        uint8_t* orow = o + (S / 2 + s * S) * (BI_TW * 3);
#pragma unroll
        for (int p = 0; p < S; ++p) {
            if (s == NSTEP - 1 && p >= S / 2) break;
            bi_vrow<S>(a, (S / 2 + p) % S, hb[(k + 1) & 3], hr[(k + 1) & 3], hb[(k + 2) & 3], hr[(k + 2) & 3], hb[(k + 3) & 3], hr[(k + 3) & 3],
                       hb[k], hr[k], orow + p * (BI_TW * 3));
        }
    }

    // ---- P4: staged tile -> HBM ------------------------------------------------------------------------------------------
    fence_proxy_async();
    __syncthreads();
    const int vw = min(BI_TW, a.dw - x0), vh = min(BI_TH, a.dh - y0);
    const int rb = vw * 3;
    const int nb16 = aligned16(a.dst, a.dpitch) ? (rb & ~15) : 0;
    if (tid < vh && nb16) {
        bulk_s2g(a.dst + (size_t)(y0 + tid) * a.dpitch + (size_t)x0 * 3, s_out + tid * (BI_TW * 3), (uint32_t)nb16);
        bulk_commit();
    }
    const int tb = rb - nb16;
    for (int i = tid; i < tb * vh; i += BI_NT) {
        const int y = i / tb, qb = nb16 + (i - y * tb);
        a.dst[(size_t)(y0 + y) * a.dpitch + (size_t)x0 * 3 + qb] = s_out[y * (BI_TW * 3) + qb];
    }
    if (tid < vh && nb16) bulk_wait_read0();
}

}  // namespace vp
