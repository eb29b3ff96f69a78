// tma_probe.cu - what the tensor-map TMA unit accepts: element type, box width, negative box origins.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/tma_probe tools/tma_probe.cu   (no -lcuda: entry point by name)
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void k_load(const __grid_constant__ CUtensorMap tm, int c0, int c1, int box_bytes, uint8_t* out) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar)), "r"(box_bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     ::"r"(s32(smem)), "l"(&tm), "r"(c0), "r"(c1), "r"(s32(&bar)) : "memory");
        uint32_t done;
        do {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(s32(&bar)) : "memory");
        } while (!done);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < box_bytes; i += blockDim.x) out[i] = smem[i];
}
int main() {
    void* fp = nullptr; cudaDriverEntryPointQueryResult q;
    cudaFree(0);
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q) != cudaSuccess || !fp) { printf("no entry point\n"); return 1; }
    EncodeTiledFn enc = (EncodeTiledFn)fp;
    const int W = 320, H = 192;
    struct Case { const char* name; int eb; int bw, bh, c0, c1; } cases[] = {
        {"u8  box 224x20 at (16,2)", 1, 224, 20, 16, 2},  {"u8  box 224x20 at (-16,-2)", 1, 224, 20, -16, -2},
        {"u32 box 64x32 at (0,0)", 4, 64, 32, 0, 0},      {"u32 box 68x68 at (0,0)", 4, 68, 68, 0, 0},
        {"u32 box 68x68 at (-2,-2)", 4, 68, 68, -2, -2},  {"u32 box 68x68 at (254,126)", 4, 68, 68, 254, 126},
        {"u32 box 72x28 at (-4,-4)", 4, 72, 28, -4, -4},  {"u16 box 192x32 at (0,0)", 2, 192, 32, 0, 0}};
    for (auto& cs : cases) {
        const size_t pitch = ((size_t)W * 4 + 255) & ~(size_t)255;
        std::vector<uint8_t> h(pitch * H);
        for (size_t i = 0; i < h.size(); ++i) h[i] = (uint8_t)(i * 7 + (i / pitch) * 13 + 1);
        uint8_t *d, *o; cudaMalloc(&d, h.size()); cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
        const int box_bytes = cs.bw * cs.eb * cs.bh;
        cudaMalloc(&o, box_bytes); cudaMemset(o, 0xEE, box_bytes);
        CUtensorMap tm;
        const cuuint64_t gdim[2] = {(cuuint64_t)(W * 4 / cs.eb), (cuuint64_t)H}, gstr[1] = {pitch};
        const cuuint32_t box[2] = {(cuuint32_t)cs.bw, (cuuint32_t)cs.bh}, es[2] = {1, 1};
        CUresult r = enc(&tm, cs.eb == 4 ? CU_TENSOR_MAP_DATA_TYPE_UINT32 : cs.eb == 2 ? CU_TENSOR_MAP_DATA_TYPE_UINT16 : CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, gdim, gstr, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("%-32s encode failed %d\n", cs.name, (int)r); continue; }
        k_load<<<1, 128, box_bytes + 128>>>(tm, cs.c0, cs.c1, box_bytes, o);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%-32s KERNEL ERROR: %s\n", cs.name, cudaGetErrorString(e)); return 2; }
        std::vector<uint8_t> got(box_bytes); cudaMemcpy(got.data(), o, box_bytes, cudaMemcpyDeviceToHost);
        long bad = 0;
        for (int y = 0; y < cs.bh; ++y)
            for (int b = 0; b < cs.bw * cs.eb; ++b) {
                const long gx = (long)cs.c0 * cs.eb + b, gy = cs.c1 + y;
                const uint8_t want = (gx < 0 || gy < 0 || gx >= W * 4 || gy >= H) ? 0 : h[gy * pitch + gx];
                if (got[y * cs.bw * cs.eb + b] != want) ++bad;
            }
        printf("%-32s ok, mismatching bytes %ld\n", cs.name, bad);
        cudaFree(d); cudaFree(o);
    }
    return 0;
}
