// vp_canny.cuh - the pieces of CPUImpl's BICUBIC wrapper (reference src/upscaler.cpp:75-148) that are not
// already kernels of the enhancement chain:
//   k_gauss3      cv::GaussianBlur(input, Size(3,3), 0.5) on CV_8UC3         upscaler.cpp:78
//   k_canny_nms   cv::cvtColor(BGR2GRAY) + cv::Canny's Sobel / |dx|+|dy| / non-maximum suppression   :107-108
//   k_canny_hyst  cv::Canny's hysteresis (weak pixels 8-connected to a strong one)                      :108
// The dilate(2x2) + 5-point sharpen + per-pixel select (:112-147) live in k_bilateral's epilogue.
// Arithmetic is integer throughout and follows oracle/restate.py (gaussian_blur_u8, bgr2gray,
// canny_nms_map, canny_hysteresis), which tests/test_oracle_restate.py pins against cv2: bit-exact.
#pragma once
#include <cooperative_groups.h>

#include "vp_common.cuh"

namespace vp {

// ---------------------------------------------------------------------------------------------------
// 3x3 8-bit Gaussian, Q8 taps (q0,q1,q0), BORDER_REFLECT_101; exact: (sum q_i q_j p + 2^15) >> 16
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_gauss3(const uint8_t* __restrict__ src, size_t spitch, uint8_t* __restrict__ dst, size_t dpitch,
                                                int w, int h, int q0, int q1) {
    const int row_px = w;
    const long long total = (long long)row_px * h;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int y = (int)(i / row_px), x = (int)(i - (long long)y * row_px);
        const int xm = reflect101(x - 1, w), xp = reflect101(x + 1, w);
        const uint8_t* r0 = src + (size_t)reflect101(y - 1, h) * spitch;
        const uint8_t* r1 = src + (size_t)y * spitch;
        const uint8_t* r2 = src + (size_t)reflect101(y + 1, h) * spitch;
        uint8_t* o = dst + (size_t)y * dpitch + (size_t)x * 3;
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
            const int a0 = q0 * (r0[xm * 3 + ch] + r0[xp * 3 + ch]) + q1 * r0[x * 3 + ch];
            const int a1 = q0 * (r1[xm * 3 + ch] + r1[xp * 3 + ch]) + q1 * r1[x * 3 + ch];
            const int a2 = q0 * (r2[xm * 3 + ch] + r2[xp * 3 + ch]) + q1 * r2[x * 3 + ch];
            o[ch] = (uint8_t)((q0 * (a0 + a2) + q1 * a1 + (1 << 15)) >> 16);
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// gray + Sobel + magnitude + non-maximum suppression -> map: 0 none, 1 weak (low < m <= high), 2 strong
// ---------------------------------------------------------------------------------------------------
constexpr int CN_TW = 64, CN_TH = 32, CN_NT = 256;
constexpr int CN_GW = CN_TW + 4, CN_GH = CN_TH + 4;      // gray tile, halo 2
constexpr int CN_MW = CN_TW + 2, CN_MH = CN_TH + 2;      // magnitude tile, halo 1

struct CannyArgs {
    const uint8_t* img; size_t ipitch; int w, h;         // BGR8 image the edges are taken from
    uint8_t* map; size_t mpitch;                          // w x h bytes
    int low, high;
};

__global__ void __launch_bounds__(CN_NT) k_canny_nms(const CannyArgs a) {
    __shared__ uint8_t s_gray[CN_GH][CN_GW + 4];
    __shared__ int s_grad[CN_MH][CN_MW];                  // (dx & 0xffff) | (dy << 16)
    __shared__ uint16_t s_mag[CN_MH][CN_MW + 2];
    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * CN_TW, y0 = blockIdx.y * CN_TH;
    // 256 threads = 64 columns x 4 row phases: every loop below walks rows with a fixed column, no div/mod
    const int cx = tid & 63, cr = tid >> 6;
    // gray with BORDER_REPLICATE (what cv::Canny's Sobel uses): clamp the coordinates
    for (int gx = cx; gx < CN_GW; gx += 64) {
        const int x = clampi(x0 - 2 + gx, 0, a.w - 1);
        for (int gy = cr; gy < CN_GH; gy += 4) {
            const int y = clampi(y0 - 2 + gy, 0, a.h - 1);
            const uint8_t* p = a.img + (size_t)y * a.ipitch + (size_t)x * 3;
            s_gray[gy][gx] = (uint8_t)((9798 * p[2] + 19235 * p[1] + 3735 * p[0] + (1 << 14)) >> 15);
        }
    }
    __syncthreads();
    for (int mx = cx; mx < CN_MW; mx += 64) {
        const int x = x0 - 1 + mx;
        const bool xin = x >= 0 && x < a.w;
        for (int my = cr; my < CN_MH; my += 4) {
            const int y = y0 - 1 + my;
            int dx = 0, dy = 0, mag = 0;
            if (xin && y >= 0 && y < a.h) {                     // the magnitude is 0 outside the image
                const int g00 = s_gray[my][mx], g01 = s_gray[my][mx + 1], g02 = s_gray[my][mx + 2];
                const int g10 = s_gray[my + 1][mx], g12 = s_gray[my + 1][mx + 2];
                const int g20 = s_gray[my + 2][mx], g21 = s_gray[my + 2][mx + 1], g22 = s_gray[my + 2][mx + 2];
                dx = (g02 + 2 * g12 + g22) - (g00 + 2 * g10 + g20);
                dy = (g20 + 2 * g21 + g22) - (g00 + 2 * g01 + g02);
                mag = abs(dx) + abs(dy);
            }
            s_grad[my][mx] = (dx & 0xffff) | (dy << 16);
            s_mag[my][mx] = (uint16_t)mag;
        }
    }
    __syncthreads();
    {
        const int tx = cx, x = x0 + tx, mx = tx + 1;
        if (x < a.w) {
            for (int ty = cr; ty < CN_TH; ty += 4) {
                const int y = y0 + ty;
                if (y >= a.h) break;
                const int my = ty + 1;
                const int m = s_mag[my][mx];
                uint8_t out = 0;
                if (m > a.low) {
                    const int g = s_grad[my][mx];
                    const int dx = (int)(short)(g & 0xffff), dy = g >> 16;
                    const int ax = abs(dx), ay = abs(dy) << 15;       // |dy| <= 1020: fits
                    const int tg22x = ax * 13573;
                    bool lm;
                    if (ay < tg22x) {
                        lm = m > s_mag[my][mx - 1] && m >= s_mag[my][mx + 1];
                    } else {
                        const int tg67x = tg22x + (ax << 16);
                        if (ay > tg67x) {
                            lm = m > s_mag[my - 1][mx] && m >= s_mag[my + 1][mx];
                        } else {
                            const int s = ((dx ^ dy) < 0) ? -1 : 1;
                            lm = m > s_mag[my - 1][mx - s] && m > s_mag[my + 1][mx + s];
                        }
                    }
                    if (lm) out = m > a.high ? 2 : 1;
                }
                a.map[(size_t)y * a.mpitch + x] = out;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// hysteresis: one cooperative launch; per tile a breadth-first frontier in shared memory, tiles whose
// neighbourhood changed are revisited after a grid-wide barrier until nothing changes anywhere
// ---------------------------------------------------------------------------------------------------
constexpr int HY_STACK = 8;
constexpr int HY_TW = 128, HY_TH = 64, HY_NT = 256;
constexpr int HY_X0 = 4;                                  // interior starts at column 4: rows are whole aligned words
constexpr int HY_PW = HY_TW + 8, HY_PH = HY_TH + 2;      // columns [tx*128-4, tx*128+132), rows [ty*64-1, ty*64+65)

struct HystArgs {
    uint8_t* map; size_t mpitch; int w, h;                // mpitch % 4 == 0
    int tiles_x, tiles_y;
    unsigned char* dirty;          // 2 * ntiles flags (ping-pong); dirty[0..ntiles) preset to 1 by the host memset
    unsigned* again;               // 3 rotating counters, zeroed by the host memset
};

__global__ void __launch_bounds__(HY_NT) k_canny_hyst(const HystArgs a) {
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    __shared__ __align__(16) uint8_t s_map[HY_PW * HY_PH];     // 0 none, 1 weak, >= 2 edge (promotion ORs in bit 1)
    __shared__ unsigned short s_q[2][HY_PW * HY_PH];
    __shared__ unsigned short s_stack[HY_NT * HY_STACK];
    __shared__ int s_n[2];
    __shared__ unsigned s_border;                              // which of the 8 neighbours saw a promotion on their halo
    const int tid = threadIdx.x;
    const int ntiles = a.tiles_x * a.tiles_y;
    for (int iter = 0;; ++iter) {
        unsigned char* dcur = a.dirty + (iter & 1) * ntiles;
        unsigned char* dnext = a.dirty + ((iter + 1) & 1) * ntiles;
        for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
            if (!*(volatile unsigned char*)&dcur[t]) continue;
            const int ty = t / a.tiles_x, tx = t - ty * a.tiles_x;
            const int x0 = tx * HY_TW - HY_X0, y0 = ty * HY_TH - 1;   // image coordinates of s_map[0][0]
            if (tid == 0) { s_n[0] = 0; s_n[1] = 0; s_border = 0; }
            __syncthreads();
            // load whole words (4 map bytes); everything outside the image or outside the 1-pixel halo reads 0
            constexpr int WPR = HY_PW / 4;
            for (int i = tid; i < WPR * HY_PH; i += HY_NT) {
                const int py = i / WPR, wx = i - py * WPR;
                const int y = y0 + py, x = x0 + 4 * wx;
                uint32_t v = 0;
                if (y >= 0 && y < a.h && x + 3 >= 0 && x < a.w) {
                    v = __ldcg(reinterpret_cast<const uint32_t*>(a.map + (size_t)y * a.mpitch + x));   // L2: other CTAs write it
                    if (x + 3 >= a.w) v &= 0xffffffffu >> (8 * (x + 4 - a.w));      // bytes past the last column
                }
                if (wx == 0) v &= 0xff000000u;                        // only column 3 (= interior-1) is halo
                if (wx == WPR - 1) v &= 0x000000ffu;                  // only column 132
                reinterpret_cast<uint32_t*>(s_map)[i] = v;
                if (v & 0x02020202u) {
#pragma unroll
                    for (int b = 0; b < 4; ++b)
                        if ((v >> (8 * b)) & 2) s_q[0][atomicAdd(&s_n[0], 1)] = (unsigned short)(4 * i + b);
                }
            }
            __syncthreads();
            // Flood from the seeds without a barrier per step: a thread claims a seed and follows the chain of
            // weak pixels itself (depth-first, small private stack in shared memory; what does not fit goes to
            // the other queue and seeds the next pass).  atomicOr makes every pixel promoted exactly once.
            int cur = 0;
            while (true) {
                const int n = s_n[cur];
                if (n == 0) break;
                unsigned short* stk = s_stack + tid * HY_STACK;
                for (int k = tid; k < n; k += HY_NT) {
                    int sp = 0;
                    stk[sp++] = s_q[cur][k];
                    while (sp > 0) {
                        const int i = stk[--sp];
                        const int py = i / HY_PW, px = i - py * HY_PW;
#pragma unroll
                        for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
                            for (int dx = -1; dx <= 1; ++dx) {
                                const int qy = py + dy, qx = px + dx;
                                // only interior pixels are owned (and written back) by this tile
                                if ((dx | dy) != 0 && qy >= 1 && qy <= HY_TH && qx >= HY_X0 && qx < HY_X0 + HY_TW) {
                                    const int j = qy * HY_PW + qx;
                                    if (s_map[j] == 1) {
                                        const int sh = (j & 3) * 8;
                                        const unsigned old = atomicOr(reinterpret_cast<unsigned*>(s_map + (j & ~3)), 2u << sh);
                                        if (((old >> sh) & 0xff) == 1) {          // this thread promoted it
                                            const unsigned top = qy == 1, bot = qy == HY_TH, lef = qx == HY_X0, rig = qx == HY_X0 + HY_TW - 1;
                                            // bit k = neighbour (k/3-1, k%3-1) in (dy,dx) order; bit 4 = "something changed"
                                            atomicOr(&s_border, 16u | (top << 1) | (bot << 7) | (lef << 3) | (rig << 5) |
                                                                    ((top & lef) << 0) | ((top & rig) << 2) | ((bot & lef) << 6) | ((bot & rig) << 8));
                                            if (sp < HY_STACK) stk[sp++] = (unsigned short)j;
                                            else s_q[cur ^ 1][atomicAdd(&s_n[cur ^ 1], 1)] = (unsigned short)j;
                                        }
                                    }
                                }
                            }
                    }
                }
                __syncthreads();
                if (tid == 0) s_n[cur] = 0;
                cur ^= 1;
                __syncthreads();
            }
            const unsigned border = s_border;
            if (border) {
                for (int i = tid; i < (HY_TW / 4) * HY_TH; i += HY_NT) {       // interior words
                    const int py = 1 + i / (HY_TW / 4), wx = 1 + i % (HY_TW / 4);
                    const int y = y0 + py, x = x0 + 4 * wx;
                    if (y >= a.h || x >= a.w) continue;
                    const uint32_t v = reinterpret_cast<const uint32_t*>(s_map)[py * WPR + wx];
                    if (v & ((v >> 1) & 0x01010101u)) {                          // some byte == 3: promoted here
#pragma unroll
                        for (int b = 0; b < 4; ++b)
                            if (((v >> (8 * b)) & 0xff) == 3 && x + b < a.w) a.map[(size_t)y * a.mpitch + x + b] = 2;
                    }
                }
                if (tid < 9 && tid != 4 && ((border >> tid) & 1)) {              // neighbours whose halo changed look again
                    const int ny = ty + tid / 3 - 1, nx = tx + tid % 3 - 1;
                    if (ny >= 0 && ny < a.tiles_y && nx >= 0 && nx < a.tiles_x) {
                        dnext[ny * a.tiles_x + nx] = 1;
                        atomicAdd(&a.again[iter % 3], 1u);
                    }
                }
            }
            __syncthreads();
            if (tid == 0) dcur[t] = 0;
        }
        __threadfence();
        grid.sync();
        const unsigned more = *(volatile unsigned*)&a.again[iter % 3];
        // again[(iter+2)%3] was last read after the previous barrier, i.e. before every CTA reached this one
        if (blockIdx.x == 0 && tid == 0) a.again[(iter + 2) % 3] = 0;
        if (!more) break;
    }
}

}  // namespace vp
