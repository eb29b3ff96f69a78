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
// into raw bytes, expanded once to packed BGR0 words with the border map resolved, and each thread
// then filters a strip of PX horizontally adjacent pixels so a staged neighbour is reused by up to
// PX outputs from registers (VABSDIFF4 + IDP.4A for the L1 colour distance).
#pragma once
#include "vp_common.cuh"
#include "vp_stats.cuh"

namespace vp {

constexpr int BIL_TW = 64, BIL_TH = 16, BIL_PX = 4, BIL_NT = 256, BIL_RMAX = 7;

struct BilCand {                   // one (d, sigma) parameter set, device memory
    int radius;
    int d;
    float space_w[(2 * BIL_RMAX + 1) * (2 * BIL_RMAX + 1)];  // [(i+r)*(2r+1) + (j+r)]
    float color_w[768];
};

struct BlendArgs {
    int mode;                      // VP_TEMPORAL_*
    int nprev;                     // previous frames available (0 = pass-through)
    const uint8_t* prev[2];        // prev[0] = t-1, prev[1] = t-2 (unblended states)
    size_t ppitch;
    uint8_t* state;                // where the unblended current frame is stored (may be null)
    size_t spitch;
    float w[3];                    // main blend: w[0]=alpha, w[1]=beta; blend_frames: per-frame weights
    float inv_wsum;                // blend_frames: 1.f / (w0+w1(+w2)) as accumulated in float32
};

struct BilArgs {
    const uint8_t* src; size_t spitch;
    uint8_t* dst; size_t dpitch;   // dst may be null when only `blend.state` is wanted
    int w, h;
    const BilCand* cand;           // 3 candidates indexed by noise class (or 1 when sums == null)
    const unsigned long long* sums;
    double npx;
    int* sel_out;                  // class actually used (debug / vp_get_stage_params), may be null
    BlendArgs blend;
};

__device__ __forceinline__ int blend_byte(const BlendArgs& b, int cur, int p0, int p1) {
    if (b.mode == 1) {             // cv::addWeighted AVX2 form: fma(cur, alpha, prev*beta)
        return rne_sat_u8(__fmaf_rn((float)cur, b.w[0], __fmul_rn((float)p0, b.w[1])));
    }
    float acc = __fmul_rn(b.w[0], (float)cur);
    acc = __fadd_rn(acc, __fmul_rn(b.w[1], (float)p0));
    if (b.nprev > 1) acc = __fadd_rn(acc, __fmul_rn(b.w[2], (float)p1));
    return rne_sat_u8(__fmul_rn(acc, b.inv_wsum));
}


// 12 packed bytes (4 BGR pixels) <-> memory; word accesses when the address is 4-byte aligned
__device__ __forceinline__ void store12(uint8_t* p, const uint32_t v[3], int npx) {
    if (npx == 4 && (((uintptr_t)p) & 3) == 0) {
        uint32_t* q = reinterpret_cast<uint32_t*>(p);
        q[0] = v[0]; q[1] = v[1]; q[2] = v[2];
    } else {
        for (int k = 0; k < 3 * npx; ++k) p[k] = (uint8_t)(v[k >> 2] >> (8 * (k & 3)));
    }
}
__device__ __forceinline__ void load12(const uint8_t* p, uint32_t v[3], int npx) {
    if (npx == 4 && (((uintptr_t)p) & 3) == 0) {
        const uint32_t* q = reinterpret_cast<const uint32_t*>(p);
        v[0] = q[0]; v[1] = q[1]; v[2] = q[2];
    } else {
        v[0] = v[1] = v[2] = 0;
        for (int k = 0; k < 3 * npx; ++k) v[k >> 2] |= (uint32_t)p[k] << (8 * (k & 3));
    }
}
__device__ __forceinline__ void blend12(const BlendArgs& b, const uint32_t cur[3], const uint32_t p0[3],
                                        const uint32_t p1[3], uint32_t o[3]) {
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        uint32_t r = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k)
            r |= (uint32_t)blend_byte(b, (cur[q] >> (8 * k)) & 0xff, (p0[q] >> (8 * k)) & 0xff, (p1[q] >> (8 * k)) & 0xff) << (8 * k);
        o[q] = r;
    }
}

template <int R>
__device__ __forceinline__ void bil_strip(const uint32_t* __restrict__ s_tile, const float* __restrict__ s_sw,
                                          const float* __restrict__ s_cw, int lx, int ly, uint32_t out[BIL_PX]) {
    constexpr int TP = (BIL_TW + 2 * R + 3) & ~3;          // tile pitch in pixels (16-byte rows)
    constexpr int NSEG = (BIL_PX + 2 * R + 3) & ~3;        // staged neighbours per row, whole uint4s
    constexpr int D = 2 * R + 1;
    uint32_t ctr[BIL_PX];
    float sb[BIL_PX], sg[BIL_PX], sr[BIL_PX], ws[BIL_PX];
#pragma unroll
    for (int j = 0; j < BIL_PX; ++j) {
        ctr[j] = s_tile[(ly + R) * TP + lx + R + j];
        sb[j] = sg[j] = sr[j] = ws[j] = 0.f;
    }
#pragma unroll
    for (int dy = -R; dy <= R; ++dy) {
        const uint4* rowp = reinterpret_cast<const uint4*>(s_tile + (ly + R + dy) * TP + lx);
        uint32_t seg[NSEG];
#pragma unroll
        for (int q = 0; q < NSEG / 4; ++q) {
            const uint4 v = rowp[q];
            seg[4 * q] = v.x; seg[4 * q + 1] = v.y; seg[4 * q + 2] = v.z; seg[4 * q + 3] = v.w;
        }
        float fb[NSEG], fg[NSEG], fr[NSEG];
#pragma unroll
        for (int i = 0; i < BIL_PX + 2 * R; ++i) {
            fb[i] = (float)(seg[i] & 0xffu);
            fg[i] = (float)((seg[i] >> 8) & 0xffu);
            fr[i] = (float)((seg[i] >> 16) & 0xffu);
        }
#pragma unroll
        for (int dx = -R; dx <= R; ++dx) {
            if (dx * dx + dy * dy <= R * R) {
                const float sw = s_sw[(dy + R) * D + (dx + R)];
#pragma unroll
                for (int j = 0; j < BIL_PX; ++j) {
                    const int i = j + dx + R;
                    const unsigned n = __dp4a(__vabsdiffu4(seg[i], ctr[j]), 0x01010101u, 0u);
                    const float wgt = __fmul_rn(sw, s_cw[n]);
                    ws[j] = __fadd_rn(ws[j], wgt);
                    sb[j] = __fmaf_rn(fb[i], wgt, sb[j]);
                    sg[j] = __fmaf_rn(fg[i], wgt, sg[j]);
                    sr[j] = __fmaf_rn(fr[i], wgt, sr[j]);
                }
            }
        }
    }
#pragma unroll
    for (int j = 0; j < BIL_PX; ++j) {
        const float inv = __frcp_rn(ws[j]);
        const int b = rne_sat_u8(__fmul_rn(sb[j], inv));
        const int g = rne_sat_u8(__fmul_rn(sg[j], inv));
        const int r = rne_sat_u8(__fmul_rn(sr[j], inv));
        out[j] = (uint32_t)b | ((uint32_t)g << 8) | ((uint32_t)r << 16);
    }
}

// shared memory: [ tile words | space_w | color_w | raw bytes | mbarrier ]
__host__ __device__ inline int bil_tile_pitch(int r) { return (BIL_TW + 2 * r + 3) & ~3; }
__host__ __device__ inline int bil_raw_pitch(int r) { return (((BIL_TW + 2 * r) * 3 + 15 + 15) & ~15) + 16; }
__host__ inline size_t bil_smem_bytes(int r) {
    size_t tile = (size_t)bil_tile_pitch(r) * (BIL_TH + 2 * r) * 4;
    size_t sw = (size_t)(2 * r + 1) * (2 * r + 1) * 4;
    sw = (sw + 15) & ~(size_t)15;
    size_t cw = 768 * 4;
    size_t raw = (size_t)bil_raw_pitch(r) * (BIL_TH + 2 * r);
    return tile + sw + cw + raw + 16;
}

__global__ void __launch_bounds__(BIL_NT) k_bilateral(const BilArgs a) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ int s_sel;
    const int tid = threadIdx.x;
    if (tid == 0) {
        int sel = 0;
        if (a.sums) sel = noise_class_from_sums(a.sums, a.npx);
        s_sel = sel;
        if (a.sel_out && blockIdx.x == 0 && blockIdx.y == 0) *a.sel_out = sel;
    }
    __syncthreads();
    const BilCand* __restrict__ cand = a.cand + s_sel;
    const int R = cand->radius;
    const int D = 2 * R + 1;
    const int TP = bil_tile_pitch(R);
    const int TROWS = BIL_TH + 2 * R;
    const int RP = bil_raw_pitch(R);

    uint32_t* s_tile = reinterpret_cast<uint32_t*>(smem);
    float* s_sw = reinterpret_cast<float*>(smem + (size_t)TP * TROWS * 4);
    float* s_cw = s_sw + ((D * D + 3) & ~3);
    uint8_t* s_raw = reinterpret_cast<uint8_t*>(s_cw + 768);
    uint64_t* bar = reinterpret_cast<uint64_t*>(s_raw + (size_t)RP * TROWS);

    const int x0 = blockIdx.x * BIL_TW, y0 = blockIdx.y * BIL_TH;
    // valid source window (image coordinates) the halo tile touches
    const int vx0 = max(x0 - R, 0), vx1 = min(x0 + BIL_TW + R, a.w);
    const int vy0 = max(y0 - R, 0), vy1 = min(y0 + BIL_TH + R, a.h);
    const int bx0 = (vx0 * 3) & ~15, bx1 = vx1 * 3;
    const bool fast = aligned16(a.src, a.spitch);

    for (int i = tid; i < D * D; i += BIL_NT) s_sw[i] = cand->space_w[i];
    for (int i = tid; i < 768; i += BIL_NT) s_cw[i] = cand->color_w[i];
    stage_rows(s_raw, RP, a.src, a.spitch, 3 * a.w, vy0, vy1, bx0, bx1, fast, bar);

    // expand raw bytes -> BGR0 words with BORDER_REFLECT_101 resolved
    const int TWH = BIL_TW + 2 * R;
    for (int i = tid; i < TWH * TROWS; i += BIL_NT) {
        const int ty = i / TWH, tx = i - ty * TWH;
        const int sx = reflect101(x0 - R + tx, a.w), sy = reflect101(y0 - R + ty, a.h);
        uint32_t v = 0;
        if (sx >= vx0 && sx < vx1 && sy >= vy0 && sy < vy1) {   // always true for w,h > R
            const uint8_t* p = s_raw + (sy - vy0) * RP + (sx * 3 - bx0);
            v = (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16);
        }
        s_tile[ty * TP + tx] = v;
    }
    __syncthreads();

    const int lx = (tid & 15) * BIL_PX, ly = tid >> 4;
    const int gx = x0 + lx, gy = y0 + ly;
    if (gy >= a.h || gx >= a.w) return;
    uint32_t out[BIL_PX];
    switch (R) {
        case 1: bil_strip<1>(s_tile, s_sw, s_cw, lx, ly, out); break;
        case 2: bil_strip<2>(s_tile, s_sw, s_cw, lx, ly, out); break;
        case 3: bil_strip<3>(s_tile, s_sw, s_cw, lx, ly, out); break;
        case 4: bil_strip<4>(s_tile, s_sw, s_cw, lx, ly, out); break;
        case 5: bil_strip<5>(s_tile, s_sw, s_cw, lx, ly, out); break;
        case 6: bil_strip<6>(s_tile, s_sw, s_cw, lx, ly, out); break;
        default: bil_strip<7>(s_tile, s_sw, s_cw, lx, ly, out); break;
    }

    // epilogue: optional temporal blend, state store, output store (12 contiguous bytes per strip)
    const int npx = min(BIL_PX, a.w - gx);
    const BlendArgs& b = a.blend;
    uint32_t cur[3];   // 12 packed bytes B G R B | G R B G | R B G R
    cur[0] = (out[0] & 0xffffffu) | (out[1] << 24);
    cur[1] = ((out[1] >> 8) & 0xffffu) | (out[2] << 16);
    cur[2] = ((out[2] >> 16) & 0xffu) | (out[3] << 8);
    const size_t boff = (size_t)gx * 3;
    if (b.state) store12(b.state + (size_t)gy * b.spitch + boff, cur, npx);
    if (!a.dst) return;
    uint8_t* dp = a.dst + (size_t)gy * a.dpitch + boff;
    if (b.mode != 0 && b.nprev > 0) {
        uint32_t p0[3], p1[3], o[3];
        load12(b.prev[0] + (size_t)gy * b.ppitch + boff, p0, npx);
        if (b.nprev > 1) load12(b.prev[1] + (size_t)gy * b.ppitch + boff, p1, npx);
        else { p1[0] = p0[0]; p1[1] = p0[1]; p1[2] = p0[2]; }
        blend12(b, cur, p0, p1, o);
        store12(dp, o, npx);
    } else {
        store12(dp, cur, npx);
    }
}

}  // namespace vp
