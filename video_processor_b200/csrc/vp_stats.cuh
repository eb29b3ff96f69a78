// vp_stats.cuh - exact per-channel sum x / sum x^2 of a CV_8UC3 frame (cv::meanStdDev's integer
// sums, reference call site src/selective_bilateral.cpp:321-322) and the device-side
// calculateAdaptiveParams selector (src/selective_bilateral.cpp:315-371).
#pragma once
#include "vp_common.cuh"

namespace vp {

// calculateAdaptiveParams: population std-dev per channel from the exact sums, mean of the three,
// -> noise factor class 0 (0.7: avg < 5), 1 (1.0), 2 (1.5: avg > 15).  Evaluated in double with
// explicitly rounded ops so it agrees with the host/numpy evaluation bit for bit.
__device__ __forceinline__ int noise_class_from_sums(const unsigned long long* sums, double npx) {
    double avg = 0.0;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const double s = (double)__ldcg(sums + c), sq = (double)__ldcg(sums + 3 + c);
        const double mean = __ddiv_rn(s, npx);
        double var = __dsub_rn(__ddiv_rn(sq, npx), __dmul_rn(mean, mean));
        var = var > 0.0 ? var : 0.0;
        avg = __dadd_rn(avg, __dsqrt_rn(var));
    }
    avg = __ddiv_rn(avg, 3.0);
    return avg < 5.0 ? 0 : (avg > 15.0 ? 2 : 1);
}

// Where the producer of a frame's sums leaves the parameter class for the consumer kernel.
struct SelOut {
    unsigned* counter;             // CTA ticket counter (zero between launches); null = no finalize
    int* sel;                      // class written by the last CTA to finish
    double npx;
};

// Second half of the reduction, run by ONE whole warp after the barrier that published s_part: lanes 0..5 add the
// per-warp partials and flush them with one 64-bit global atomic each, then the last CTA of the grid evaluates
// the class once (threadfence-reduction pattern at warp scope) so the consumer does no double-precision work.
// Only this warp waits for the atomics and the fence; the CTA's other warps carry on (stores) or exit.
__device__ __forceinline__ void warp_flush_and_finalize(const unsigned long long* s_part, int nwarps,
                                                        unsigned long long* __restrict__ sums, const SelOut f, unsigned nblocks) {
    const int lane = threadIdx.x & 31;
    if (lane < 6) {
        unsigned long long t = 0;
        for (int w = 0; w < nwarps; ++w) t += s_part[w * 6 + lane];
        if (t) atomicAdd(&sums[lane], t);
        if (f.counter) __threadfence();
    }
    if (!f.counter) return;
    __syncwarp();
    if (lane == 0) {
        const unsigned ticket = atomicAdd(f.counter, 1u);
        if (ticket == nblocks - 1) {
            __threadfence();
            *f.sel = noise_class_from_sums(sums, f.npx);
            *f.counter = 0u;
        }
    }
}

// byte k of three consecutive words a|b|c (12 bytes = 4 BGR pixels) -> one word per channel
__device__ __forceinline__ void deinterleave12(uint32_t a, uint32_t b, uint32_t c, uint32_t& ch0, uint32_t& ch1, uint32_t& ch2) {
    // bytes: a = 0 1 2 0, b = 1 2 0 1, c = 2 0 1 2   (channel of each byte, little endian)
    ch0 = __byte_perm(__byte_perm(a, b, 0x0630), c, 0x5210 | 0);   // a0 a3 b2 c1
    ch1 = __byte_perm(__byte_perm(a, b, 0x0741), c, 0x6210 | 0);   // a1 b0 b3 c2
    ch2 = __byte_perm(__byte_perm(a, b, 0x0052), c, 0x7410 | 0);   // a2 b1 c0 c3
}

__device__ __forceinline__ void acc12(uint32_t a, uint32_t b, uint32_t c, unsigned acc[6]) {
    uint32_t c0, c1, c2;
    deinterleave12(a, b, c, c0, c1, c2);
    acc[0] = __dp4a(c0, 0x01010101u, acc[0]); acc[3] = __dp4a(c0, c0, acc[3]);
    acc[1] = __dp4a(c1, 0x01010101u, acc[1]); acc[4] = __dp4a(c1, c1, acc[4]);
    acc[2] = __dp4a(c2, 0x01010101u, acc[2]); acc[5] = __dp4a(c2, c2, acc[5]);
}

// Fast path: the frame is one contiguous, 16-byte aligned run of 3*w*h bytes (pitch == 3*w): it is
// streamed as 48-byte groups (three 128-bit loads = 16 pixels) with a grid-stride loop.  Otherwise rows
// are walked in 12-byte groups (4-byte aligned) or bytes.  Warp shuffles, one barrier and one 64-bit atomic
// per CTA per accumulator (issued by warp 0) finish the reduction.
__global__ void __launch_bounds__(256) k_stats(const uint8_t* __restrict__ src, size_t pitch, int w, int h,
                                               unsigned long long* __restrict__ sums, const SelOut fin) {
    __shared__ unsigned long long s_part[8 * 6];
    const int row_bytes = 3 * w;
    unsigned acc[6] = {0, 0, 0, 0, 0, 0};
    unsigned* tail = acc;      // byte tails: a few dozen bytes per thread at most, same 32-bit accumulators
    const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x, gsz = (long long)gridDim.x * blockDim.x;
    if (pitch == (size_t)row_bytes && ((uintptr_t)src & 15) == 0) {
        const long long total = (long long)row_bytes * h;
        const long long ngroups = total / 48;
        const uint4* p = reinterpret_cast<const uint4*>(src);
        for (long long g = gtid; g < ngroups; g += gsz) {      // host sizes the grid so < 4096 groups/thread
            const uint4 v0 = __ldg(p + 3 * g), v1 = __ldg(p + 3 * g + 1), v2 = __ldg(p + 3 * g + 2);
            acc12(v0.x, v0.y, v0.z, acc); acc12(v0.w, v1.x, v1.y, acc);
            acc12(v1.z, v1.w, v2.x, acc); acc12(v2.y, v2.z, v2.w, acc);
        }
        for (long long i = ngroups * 48 + gtid; i < total; i += gsz) {
            const unsigned v = src[i];
            const int ch = (int)(i % 3);
#pragma unroll
            for (int k = 0; k < 3; ++k)
                if (ch == k) { tail[k] += v; tail[3 + k] += v * v; }
        }
    } else {
        const bool al4 = ((((uintptr_t)src) | pitch) & 3) == 0;
        const int ngroups = al4 ? row_bytes / 12 : 0;            // 12-byte groups per row
        const int chunks_per_row = (ngroups + 255) / 256;
        const long long nchunks = (long long)chunks_per_row * h;
        for (long long c = blockIdx.x; c < nchunks; c += gridDim.x) {
            const int y = (int)(c / chunks_per_row);
            const int g = (int)(c - (long long)y * chunks_per_row) * 256 + threadIdx.x;
            if (g < ngroups) {
                const uint32_t* p = reinterpret_cast<const uint32_t*>(src + (size_t)y * pitch) + 3 * g;
                acc12(__ldg(p), __ldg(p + 1), __ldg(p + 2), acc);
            }
        }
        const int tail0 = 12 * ngroups, tailn = row_bytes - tail0;
        if (tailn > 0) {
            const long long ntail = (long long)tailn * h;
            for (long long i = gtid; i < ntail; i += gsz) {
                const int y = (int)(i / tailn);
                const int b = tail0 + (int)(i - (long long)y * tailn);
                const unsigned v = src[(size_t)y * pitch + b];
                const int ch = b % 3;
#pragma unroll
                for (int k = 0; k < 3; ++k)
                    if (ch == k) { tail[k] += v; tail[3 + k] += v * v; }
            }
        }
    }
    warp_partials6(acc, s_part);
    __syncthreads();
    if (threadIdx.x < 32) warp_flush_and_finalize(s_part, 8, sums, fin, gridDim.x);
}

}  // namespace vp
