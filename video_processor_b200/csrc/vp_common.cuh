// vp_common.cuh - device helpers shared by the sm_100a kernels of libvpb200.
// Border maps, TMA (bulk async copy) row staging with an mbarrier, saturating casts.
#pragma once
#include <cuda.h>            // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint, no libcuda link)
#include <cuda_runtime.h>
#include <stdint.h>

namespace vp {

// --------------------------------------------------------------------------------------------
// border maps
// --------------------------------------------------------------------------------------------
// cv::BORDER_REFLECT_101 (gfedcb|abcdefgh|gfedcba), valid for any offset.
__host__ __device__ __forceinline__ int reflect101(int i, int n) {
    if (n == 1) return 0;
    if ((unsigned)i < (unsigned)n) return i;
    int p = 2 * (n - 1);
    i %= p;
    if (i < 0) i += p;
    return i >= n ? p - i : i;
}
__host__ __device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

__device__ __forceinline__ int sat_u8(int v) { return min(max(v, 0), 255); }
// cv::saturate_cast<uchar>(float): cvRound (round-half-even) then clamp
__device__ __forceinline__ int rne_sat_u8(float v) { return sat_u8(__float2int_rn(v)); }

// --------------------------------------------------------------------------------------------
// mbarrier + bulk async copy (TMA, SASS UBLKCP) : global -> shared rows
// --------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    // try_wait suspends the thread in hardware for up to the hint (ns) before it returns false, so the loop body is
    // issued a handful of times, not once per scheduler slot
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(smem_u32(bar)), "r"(parity), "r"(0x989680u) : "memory");
    } while (!done);
}
// 16-byte aligned src, dst; bytes a multiple of 16
__device__ __forceinline__ void bulk_g2s(void* sdst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(sdst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* gdst, const void* ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// 2-D tensor-map TMA (SASS UTMALDG / UTMASTG): one instruction moves a whole box; rows outside the tensor are
// zero-filled on loads and clipped on stores.  Shared-memory side 128-byte aligned, dense rows of the box width.
__device__ __forceinline__ void tma_load_2d(void* sdst, const CUtensorMap* tm, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(sdst)), "l"(tm), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tm, int c0, int c1, const void* ssrc) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
                 ::"l"(tm), "r"(c0), "r"(c1), "r"(smem_u32(ssrc)) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tm) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(tm) : "memory");
}

// Stage image rows [y0,y1) x bytes [bx0,bx1) (bx0 a multiple of 16) into shared memory, row r at
// s_raw + (r-y0)*s_pitch, byte b at offset b-bx0.  `fast` (base and pitch 16-byte aligned) uses one TMA bulk copy
// per row issued by warp 0 and completing on `bar` (parity 0); otherwise all threads copy bytes.  The staging is split
// in two so a kernel can fill its tables while the bulk copies are in flight: stage_rows_issue() starts the copies
// (fast path: warp 0 only, no CTA barrier), stage_rows_wait() returns once the rows are visible to every thread.
// Only warp 0 polls the mbarrier - the other warps sleep in the CTA barrier behind it instead of spending issue slots
// on a spin loop (ncu: the all-warps spin was 3-8 % of the executed instructions of the tiled kernels).  Every thread
// must call both, with the same arguments.  `tab_*` / `tab2_*`: optional table blocks (16-byte aligned, a multiple of
// 16 bytes) fetched by one more bulk copy each on the same mbarrier; on the slow path the threads copy them word by word.
__device__ __forceinline__ void stage_rows_issue(uint8_t* s_raw, int s_pitch, const uint8_t* __restrict__ g, size_t gpitch,
                                                 int y0, int y1, int bx0, int bx1, bool fast, uint64_t* bar,
                                                 void* tab_dst = nullptr, const void* tab_src = nullptr, uint32_t tab_bytes = 0,
                                                 void* tab2_dst = nullptr, const void* tab2_src = nullptr, uint32_t tab2_bytes = 0) {
    const int tid = threadIdx.x;
    if (!fast) {
        for (uint32_t i = tid; i < tab_bytes / 4; i += blockDim.x)
            reinterpret_cast<uint32_t*>(tab_dst)[i] = reinterpret_cast<const uint32_t*>(tab_src)[i];
        for (uint32_t i = tid; i < tab2_bytes / 4; i += blockDim.x)
            reinterpret_cast<uint32_t*>(tab2_dst)[i] = reinterpret_cast<const uint32_t*>(tab2_src)[i];
        return;
    }
    if (tid < 32) {
        const int nrows = y1 - y0;
        const int nb = ((bx1 + 15) & ~15) - bx0;
        if (tid == 0) {
            mbar_init(bar, 1);
            mbar_fence_init();
            mbar_arrive_expect_tx(bar, (uint32_t)(nb * nrows) + tab_bytes + tab2_bytes);
            if (tab_bytes) bulk_g2s(tab_dst, tab_src, tab_bytes, bar);
            if (tab2_bytes) bulk_g2s(tab2_dst, tab2_src, tab2_bytes, bar);
        }
        __syncwarp();
        for (int r = tid; r < nrows; r += 32)
            bulk_g2s(s_raw + r * s_pitch, g + (size_t)(y0 + r) * gpitch + bx0, (uint32_t)nb, bar);
    }
}
__device__ __forceinline__ void stage_rows_wait(uint8_t* s_raw, int s_pitch, const uint8_t* __restrict__ g, size_t gpitch,
                                                int row_bytes, int y0, int y1, int bx0, int bx1, bool fast, uint64_t* bar) {
    if (fast) {
        if (threadIdx.x < 32) mbar_wait(bar, 0);
        __syncthreads();
        return;
    }
    const int tid = threadIdx.x, nt = blockDim.x, nrows = y1 - y0;
    const int e = min(bx1, row_bytes);
    const int nb = e - bx0;
    for (int i = tid; i < nb * nrows; i += nt) {
        int r = i / nb, b = i - r * nb;
        s_raw[r * s_pitch + b] = g[(size_t)(y0 + r) * gpitch + bx0 + b];
    }
    __syncthreads();
}

// The same staging as ONE tensor-map box (SASS UTMALDG): box rows are dense in shared memory (pitch = box width), the
// part of the box outside the image is zero-filled and still counted, so the expected byte count is the whole box.
// Thread 0 issues; pair with stage_rows_wait(..., fast = true, ...).  s_raw 128-byte aligned.
__device__ __forceinline__ void stage_box_issue(uint8_t* s_raw, const CUtensorMap* tm, int c0, int c1, uint32_t box_bytes, uint64_t* bar,
                                                void* tab_dst = nullptr, const void* tab_src = nullptr, uint32_t tab_bytes = 0,
                                                void* tab2_dst = nullptr, const void* tab2_src = nullptr, uint32_t tab2_bytes = 0) {
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        mbar_fence_init();
        mbar_arrive_expect_tx(bar, box_bytes + tab_bytes + tab2_bytes);
        tma_load_2d(s_raw, tm, c0, c1, bar);
        if (tab_bytes) bulk_g2s(tab_dst, tab_src, tab_bytes, bar);
        if (tab2_bytes) bulk_g2s(tab2_dst, tab2_src, tab2_bytes, bar);
    }
}

__device__ __forceinline__ bool aligned16(const void* p, size_t pitch) {
    return ((((uintptr_t)p) | pitch) & 15) == 0;
}

// First half of the CTA reduction of 6 uint32 per-thread partial sums (no shared-memory atomics: 64-bit shared
// atomics are CAS loops): warp shuffles, then lane 0 of every warp stores its 6 totals to s_part[warp*6 ..].
// The caller's next __syncthreads() publishes them; one warp then finishes with warp_flush_and_finalize().
__device__ __forceinline__ void warp_partials6(const unsigned v[6], unsigned long long* s_part) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        unsigned long long x = v[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if (lane == 0) s_part[warp * 6 + k] = x;
    }
}

}  // namespace vp
