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

constexpr int BI_TW = 128, BI_NT = 128;
#ifndef VP_BI_TH4
#define VP_BI_TH4 64
#endif
// tile height: at x4 a band of RB output rows costs 4 + RB/4 horizontally filtered source rows, so taller tiles halve the
// band overlap (RB 8 -> 16); at x2 the overlap is small already
#ifndef VP_BI_TH2
#define VP_BI_TH2 32
#endif
__host__ __device__ constexpr int bi_th(int S) { return S == 4 ? VP_BI_TH4 : VP_BI_TH2; }

struct BicubicIntArgs {
    const uint8_t* src; size_t spitch; int sw, sh;
    uint8_t* dst; size_t dpitch; int dw, dh;
    float2 xk[4][4];     // [column phase][tap] as (k, k): 11-bit integer taps, exact in float32
    float2 yk[4][4];     // [row phase][tap] as (b, b): beta * 2^-22 as float32 (ResizeTables::ycoef)
    // 16-byte aligned frames: the source window arrives and the finished tile leaves as ONE tensor-map TMA box each
    // (src: bytes x rows, box WB x WR; dst: 16-bit units x rows, box 3*TW/2 x TH - a box dimension holds at most 256 elements)
    int use_tm;
    alignas(64) CUtensorMap tm_src;
    alignas(64) CUtensorMap tm_dst;
};

#ifndef VP_BI_K
#define VP_BI_K 2
#endif
constexpr int BI_K = VP_BI_K;      // consecutive tiles (along x, then y) one CTA processes; tile i+1's TMA load flies under tile i's math

template <int S> struct BicubicIntGeom {
    static constexpr int NCG = BI_TW / S;          // column groups per tile
    static constexpr int NB = BI_NT / NCG;         // row bands per tile
    static constexpr int TH = bi_th(S);
    static constexpr int RB = TH / NB;             // output rows per band
    static constexpr int WC = BI_TW / S + 4;       // source window of a tile: columns x0/S - 2 .. x0/S + TW/S + 1
    static constexpr int WR = TH / S + 4;
    static constexpr int WB = (WC * 3 + 15 + 15) & ~15;            // bytes of a window row from its 16-byte aligned start: the TMA box width
    static constexpr int RP = WB + 16;                             // raw row pitch of the row-copy path (16 bytes of slack)
    // raw window | staged output tile | float4 window | mbarrier: the raw bytes are free again once P1 has converted
    // them (the next tile's load is issued right behind that barrier), the output tile drains by TMA during the next P1
    static constexpr int RAW = 0;
    static constexpr int OUT = (RP * WR + 127) & ~127;
    static constexpr int F4 = OUT + TH * BI_TW * 3;
    static_assert(WB <= 256 && BI_TW * 3 / 2 <= 256, "TMA box dimensions");
    static constexpr int BAR = F4 + WC * WR * 16;
    static constexpr int TOTAL = BAR + 16;
    static_assert(OUT % 16 == 0 && F4 % 16 == 0 && BAR % 16 == 0, "TMA and float4 alignment");
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

// source window of tile (x0, y0) in unclamped / clamped source coordinates
template <int S> struct BiWindow {
    int ux0, uy0, sy0, sy1, bx0, bx1;
    __device__ __forceinline__ BiWindow(const BicubicIntArgs& a, int x0, int y0) {
        using G = BicubicIntGeom<S>;
        ux0 = x0 / S - 2; uy0 = y0 / S - 2;       // cv::resize's xofs[S j] = j - 2 at an integer ratio
        const int sx0 = clampi(ux0, 0, a.sw - 1), sx1 = clampi(ux0 + G::WC - 1, 0, a.sw - 1);
        sy0 = clampi(uy0, 0, a.sh - 1); sy1 = clampi(uy0 + G::WR - 1, 0, a.sh - 1);
        bx0 = (sx0 * 3) & ~15; bx1 = (sx1 + 1) * 3;
    }
};

template <int S>
__global__ void __launch_bounds__(BI_NT, 5) k_bicubic_int(const __grid_constant__ BicubicIntArgs a, int ntx, int ntiles) {
    using G = BicubicIntGeom<S>;
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* s_raw = smem + G::RAW;
    float4* s_f4 = reinterpret_cast<float4*>(smem + G::F4);
    uint8_t* s_out = smem + G::OUT;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + G::BAR);
    const int tid = threadIdx.x;
    const bool fast_stage = aligned16(a.src, a.spitch);
    const bool tm = a.use_tm != 0;                 // implies fast_stage and a 16-byte aligned destination
    const int RPW = tm ? G::WB : G::RP;            // raw row pitch: the TMA box is dense
    const int t0 = blockIdx.x * BI_K, t1 = min(t0 + BI_K, ntiles);

    // warp 0 starts the TMA row copies of tile t's source window (one mbarrier phase per tile)
    auto issue = [&](int t) {
        if (!fast_stage || tid >= 32) return;
        const int ty = t / ntx, tx = t - ty * ntx;
        const BiWindow<S> w(a, tx * BI_TW, ty * G::TH);
        if (tm) {          // rows below the image are zero-filled (and counted): always the whole box
            if (tid == 0) {
                mbar_arrive_expect_tx(bar, (uint32_t)(G::WB * G::WR));
                tma_load_2d(s_raw, &a.tm_src, w.bx0, w.sy0, bar);
            }
            return;
        }
        const int nrows = w.sy1 + 1 - w.sy0, nb = ((w.bx1 + 15) & ~15) - w.bx0;
        if (tid == 0) mbar_arrive_expect_tx(bar, (uint32_t)(nb * nrows));
        __syncwarp();
        for (int r = tid; r < nrows; r += 32)
            bulk_g2s(s_raw + r * G::RP, a.src + (size_t)(w.sy0 + r) * a.spitch + w.bx0, (uint32_t)nb, bar);
    };
    if (fast_stage) {
        if (tid == 0) { mbar_init(bar, 1); mbar_fence_init(); }
        __syncthreads();
        issue(t0);
    }
    bool stores_pending = false;

    for (int t = t0; t < t1; ++t) {
        const int ty = t / ntx, tx = t - ty * ntx;
        const int x0 = tx * BI_TW, y0 = ty * G::TH;
        const BiWindow<S> w(a, x0, y0);
        if (fast_stage) mbar_wait(bar, (uint32_t)(t - t0) & 1u);    // long since complete for every tile but the first
        else stage_rows_wait(s_raw, G::RP, a.src, a.spitch, 3 * a.sw, w.sy0, w.sy1 + 1, w.bx0, w.bx1, false, bar);

        // ---- P1: raw bytes -> float4 (B, G, R, 0) per window pixel, replicate border materialised ------------------
        const bool interior = w.ux0 >= 0 && w.ux0 + G::WC <= a.sw && w.uy0 >= 0 && w.uy0 + G::WR <= a.sh && (RPW & 3) == 0;
        if (interior) {
            // no clamping in this tile: a pixel's three bytes come out of two aligned words by one funnel shift (consecutive
            // lanes write consecutive float4s: conflict-free, unlike four pixels per lane)
            const unsigned long long bias = pack2(0xCB000000u, 0xCB000000u);       // (-2^23, -2^23)
            const int c0 = 3 * w.ux0 - w.bx0;              // byte offset of window column 0 inside a raw row, 0..15
            const uint32_t K = 0x4B000000u;
            for (int i = tid; i < G::WC * G::WR; i += BI_NT) {
                const int rr = i / G::WC, cc = i - rr * G::WC;
                const int bo = c0 + 3 * cc;
                const uint32_t* q = reinterpret_cast<const uint32_t*>(s_raw + (w.uy0 + rr - w.sy0) * RPW + (bo & ~3));
                const uint32_t v = __funnelshift_r(q[0], q[1], 8 * (bo & 3));
                float b, g;
                unpack2(fadd2(pack2(__byte_perm(v, K, 0x7440), __byte_perm(v, K, 0x7441)), bias), b, g);
                const float r = __uint_as_float(__byte_perm(v, K, 0x7442)) - 8388608.0f;
                s_f4[i] = make_float4(b, g, r, 0.0f);
            }
        } else {
            const unsigned long long bias = pack2(0xCB000000u, 0xCB000000u);       // (-2^23, -2^23)
            for (int i = tid; i < G::WC * G::WR; i += BI_NT) {
                const int rr = i / G::WC, cc = i - rr * G::WC;
                const uint8_t* p = s_raw + (clampi(w.uy0 + rr, 0, a.sh - 1) - w.sy0) * RPW + 3 * clampi(w.ux0 + cc, 0, a.sw - 1) - w.bx0;
                const uint32_t v = (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16);
                float b, g;
                unpack2(fadd2(pack2(__byte_perm(v, 0x4B000000u, 0x7440), __byte_perm(v, 0x4B000000u, 0x7441)), bias), b, g);
                const float r = __uint_as_float(__byte_perm(v, 0x4B000000u, 0x7442)) - 8388608.0f;
                s_f4[i] = make_float4(b, g, r, 0.0f);
            }
        }
        if (stores_pending) bulk_wait_read0();       // the previous tile has left s_out (its stores ran under this P1)
        __syncthreads();                             // window converted; raw bytes and s_out free
        if (t + 1 < t1) issue(t + 1);                // next tile's window flies under this tile's math

        // ---- P2-P3: S columns x RB rows per thread ---------------------------------------------------------------------
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
        // every further source row serves S output rows (phases S/2 .. S-1, then 0 .. S/2-1); the band's last step only
        // its first S/2 (the rest belong to the next band).  The window rotates by index: the loop is unrolled by four.
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

        // ---- P4: staged tile -> HBM ------------------------------------------------------------------------------------
        fence_proxy_async();
        __syncthreads();
        if (tm) {          // one box store; the part of an edge tile outside the frame is clipped by the TMA unit
            if (tid == 0) { tma_store_2d(&a.tm_dst, x0 * 3 / 2, y0, s_out); bulk_commit(); }
            stores_pending = tid == 0;
            continue;
        }
        const int vw = min(BI_TW, a.dw - x0), vh = min(G::TH, a.dh - y0);
        const int rb = vw * 3;
        const int nb16 = aligned16(a.dst, a.dpitch) ? (rb & ~15) : 0;
        stores_pending = false;
        if (tid < vh && nb16) {
            bulk_s2g(a.dst + (size_t)(y0 + tid) * a.dpitch + (size_t)x0 * 3, s_out + tid * (BI_TW * 3), (uint32_t)nb16);
            bulk_commit();
            stores_pending = true;
        }
        const int tb = rb - nb16;
        for (int i = tid; i < tb * vh; i += BI_NT) {
            const int y = i / tb, qb = nb16 + (i - y * tb);
            a.dst[(size_t)(y0 + y) * a.dpitch + (size_t)x0 * 3 + qb] = s_out[y * (BI_TW * 3) + qb];
        }
        // (the byte-store tail reads s_out with ordinary loads: ordered before the next tile's writes by its barrier)
    }
    if (stores_pending) bulk_wait_read0();
}

}  // namespace vp
