// vp_stats.cuh - exact per-channel sum x / sum x^2 of a CV_8UC3 frame (cv::meanStdDev's integer
// sums, reference call site src/selective_bilateral.cpp:321-322) and the device-side
// calculateAdaptiveParams selector (src/selective_bilateral.cpp:315-371).
#pragma once
#include "vp_common.cuh"

namespace vp {

// One warp-row-chunk at a time: rows are read as 32-bit words in groups of 3 (12 bytes = 4 px) so
// the byte->channel map of each word is static; IDP.4A does the horizontal adds, warp shuffles and
// one 64-bit atomic per CTA per accumulator finish the reduction.
__global__ void __launch_bounds__(256) k_stats(const uint8_t* __restrict__ src, size_t pitch, int w, int h,
                                               unsigned long long* __restrict__ sums) {
    __shared__ unsigned long long s_red[6];
    if (threadIdx.x < 6) s_red[threadIdx.x] = 0ull;
    __syncthreads();
    const int row_bytes = 3 * w;
    unsigned acc[6] = {0, 0, 0, 0, 0, 0};
    const bool al4 = ((((uintptr_t)src) | pitch) & 3) == 0;
    const int ngroups = al4 ? row_bytes / 12 : 0;  // 12-byte groups per row
    // grid-stride over (row, chunk of 256 groups)
    const int chunks_per_row = (ngroups + 255) / 256;
    const long long nchunks = (long long)chunks_per_row * h;
    for (long long c = blockIdx.x; c < nchunks; c += gridDim.x) {
        const int y = (int)(c / chunks_per_row);
        const int g = (int)(c - (long long)y * chunks_per_row) * 256 + threadIdx.x;
        if (g < ngroups) {
            const uint32_t* p = reinterpret_cast<const uint32_t*>(src + (size_t)y * pitch) + 3 * g;
            const uint32_t w0 = __ldg(p), w1 = __ldg(p + 1), w2 = __ldg(p + 2);
            // byte->channel: w0 = c0 c1 c2 c0 ; w1 = c1 c2 c0 c1 ; w2 = c2 c0 c1 c2   (little endian)
            acc[0] += __dp4a(w0, 0x01000001u, 0u) + __dp4a(w1, 0x00010000u, 0u) + __dp4a(w2, 0x00000100u, 0u);
            acc[1] += __dp4a(w0, 0x00000100u, 0u) + __dp4a(w1, 0x01000001u, 0u) + __dp4a(w2, 0x00010000u, 0u);
            acc[2] += __dp4a(w0, 0x00010000u, 0u) + __dp4a(w1, 0x00000100u, 0u) + __dp4a(w2, 0x01000001u, 0u);
            acc[3] += __dp4a(w0 & 0xff0000ffu, w0, 0u) + __dp4a(w1 & 0x00ff0000u, w1, 0u) + __dp4a(w2 & 0x0000ff00u, w2, 0u);
            acc[4] += __dp4a(w0 & 0x0000ff00u, w0, 0u) + __dp4a(w1 & 0xff0000ffu, w1, 0u) + __dp4a(w2 & 0x00ff0000u, w2, 0u);
            acc[5] += __dp4a(w0 & 0x00ff0000u, w0, 0u) + __dp4a(w1 & 0x0000ff00u, w1, 0u) + __dp4a(w2 & 0xff0000ffu, w2, 0u);
        }
    }
    // Each thread saw at most ceil(nchunks/gridDim.x) groups of 4 px: < 2^32/(4*65025) = 16512 groups
    // is guaranteed by the host's grid sizing.
    // row tails (and whole rows when the frame is not 4-byte aligned): bytes [12*ngroups, row_bytes)
    const int tail0 = 12 * ngroups, tailn = row_bytes - tail0;
    if (tailn > 0) {
        const long long ntail = (long long)tailn * h;
        unsigned long long t[6] = {0, 0, 0, 0, 0, 0};
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < ntail; i += (long long)gridDim.x * blockDim.x) {
            const int y = (int)(i / tailn);
            const int b = tail0 + (int)(i - (long long)y * tailn);
            const unsigned v = src[(size_t)y * pitch + b];
            const int ch = b % 3;
            t[ch] += v;
            t[3 + ch] += v * v;
        }
#pragma unroll
        for (int k = 0; k < 6; ++k)
            if (t[k]) atomicAdd(&s_red[k], t[k]);
    }
    reduce_sums6(acc, sums, s_red);
    __syncthreads();
    if (threadIdx.x < 6 && s_red[threadIdx.x]) atomicAdd(&sums[threadIdx.x], s_red[threadIdx.x]);
}

// calculateAdaptiveParams: population std-dev per channel from the exact sums, mean of the three,
// -> noise factor class 0 (0.7: avg < 5), 1 (1.0), 2 (1.5: avg > 15).  Evaluated in double with
// explicitly rounded ops so it agrees with the host/numpy evaluation bit for bit.
__device__ __forceinline__ int noise_class_from_sums(const unsigned long long* __restrict__ sums, double npx) {
    double avg = 0.0;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const double s = (double)sums[c], sq = (double)sums[3 + c];
        const double mean = __ddiv_rn(s, npx);
        double var = __dsub_rn(__ddiv_rn(sq, npx), __dmul_rn(mean, mean));
        var = var > 0.0 ? var : 0.0;
        avg = __dadd_rn(avg, __dsqrt_rn(var));
    }
    avg = __ddiv_rn(avg, 3.0);
    return avg < 5.0 ? 0 : (avg > 15.0 ? 2 : 1);
}

}  // namespace vp
