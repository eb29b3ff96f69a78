// pipebench.cu - issue-rate microbenchmark for the sm_100a instruction mix of the enhance-chain kernels.
// Prints lane-ops per clock per SM for each op, measured with clock64() inside one CTA per SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipebench pipebench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITERS 2048
#define NCH 8
__device__ __forceinline__ unsigned long long pack2(float a, float b){ unsigned long long r; asm("mov.b64 %0, {%1,%2};":"=l"(r):"f"(a),"f"(b)); return r;}
template<int OP> __global__ void k(float* out, long long* cyc, const float* lut_g, int seed) {
  __shared__ float lut[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) lut[i] = lut_g[i];
  __syncthreads();
  float f[NCH]; unsigned u[NCH]; unsigned long long p[NCH];
  for (int i = 0; i < NCH; ++i) { f[i] = threadIdx.x * 0.001f + i + seed; u[i] = threadIdx.x * 2654435761u + i * 40503u + seed; p[i] = pack2(f[i], f[i] + 1.f); }
  float c1 = 1.0001f + seed, c2 = 0.5f + seed; unsigned uc = 0x01010101u * (seed + 1);
  long long t0 = clock64();
  #pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
    #pragma unroll
    for (int i = 0; i < NCH; ++i) {
      if (OP == 0) f[i] = __fmaf_rn(f[i], c1, c2);                       // FFMA reg,reg,reg
      if (OP == 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(p[(i+1)%NCH]), "l"(p[(i+2)%NCH]));
      if (OP == 2) f[i] = __fadd_rn(f[i], c1);
      if (OP == 3) f[i] = __fmul_rn(f[i], c1);
      if (OP == 4) u[i] = __byte_perm(u[i], uc, 0x7440);
      if (OP == 5) u[i] = __vabsdiffu4(u[i], uc);
      if (OP == 6) u[i] = __dp4a(u[i], uc, u[i]);
      if (OP == 7) u[i] = u[i] * 4 + uc;                                   // LEA/IMAD
      if (OP == 8) f[i] = lut[(__float_as_uint(f[i]) >> 3) & 1023] + f[i];   // LDS random + FADD
      if (OP == 9) f[i] = lut[(threadIdx.x + i + it) & 1023] + f[i];        // LDS conflict-free + FADD
      if (OP == 10) { unsigned r; asm volatile("cvt.rni.u8.f32 %0, %1;" : "=r"(r) : "f"(f[i])); u[i] += r; }
      if (OP == 11) f[i] = (float)(u[i] & 0xff) + f[i];                    // I2F + FADD
      if (OP == 12) u[i] = u[i] * uc + 12345u;                             // IMAD
      if (OP == 13) u[i] = (u[i] + uc) ^ (u[i] >> 3);                      // IADD3/LOP3/SHF mix
      if (OP == 14) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(p[(i+1)%NCH]));
      if (OP == 15) { f[i] = __fmaf_rn(f[i], c1, c2); u[i] = __byte_perm(u[i], uc, 0x7440); }  // FFMA + PRMT dual pipe
      if (OP == 16) { f[i] = __fmaf_rn(f[i], c1, c2); u[i] = __dp4a(u[i], uc, u[i]); }        // FFMA + IDP
      if (OP == 17) { f[i] = __fmaf_rn(f[i], c1, c2); u[i] = __vabsdiffu4(u[i], uc); }        // FFMA + VABSDIFF4
      if (OP == 18) { asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(p[(i+1)%NCH]), "l"(p[(i+2)%NCH])); u[i] = __byte_perm(u[i], uc, 0x7440); }
      if (OP == 19) f[i] = __fmaf_rn(f[i], 1.0001f, 0.5f);                  // FFMA imm
      if (OP == 20) { unsigned long long wb = pack2(c1, c1); asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(p[i]) : "l"(p[(i+1)%NCH]), "l"(wb)); }   // acc += data * w(bcast)
      if (OP == 21) { unsigned long long wb = pack2(f[i], f[i]); asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(p[i]) : "l"(p[(i+1)%NCH]), "l"(wb)); f[i] = __fmul_rn(f[i], c1); }  // FMUL + FFMA2 bcast
      if (OP == 22) { f[i] = __fmaf_rn(f[(i+1)%NCH], c1, f[i]); }            // FFMA acc form
      if (OP == 23) { unsigned long long wb = pack2(c1, c1); asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(wb)); }   // FADD2 bcast
      if (OP == 24) { unsigned long long wb = pack2(c1, c1); asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(wb)); }   // FMUL2 bcast
    }
  }
  long long t1 = clock64();
  float s = 0; unsigned su = 0; unsigned long long sp = 0;
  for (int i = 0; i < NCH; ++i) { s += f[i]; su += u[i]; sp += p[i]; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + su + (float)sp;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template<int OP> void run(const char* name, int opsPerIter, float* out, long long* cyc, float* lut) {
  int nsm = 148, nt = 1024;
  k<OP><<<nsm, nt>>>(out, cyc, lut, 0); cudaDeviceSynchronize();
  k<OP><<<nsm, nt>>>(out, cyc, lut, 1); cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < nsm; ++i) avg += h[i]; avg /= nsm;
  double warp_instr_per_clk_sm = (double)ITERS * NCH * opsPerIter * (nt / 32) / avg;
  printf("%-28s %8.3f warp-instr/clk/SM  (%6.1f lanes/clk/SM)  err=%s\n", name, warp_instr_per_clk_sm, warp_instr_per_clk_sm * 32, cudaGetErrorString(cudaGetLastError()));
}
int main() {
  float* out; long long* cyc; float* lut; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8); cudaMalloc(&lut, 4096); cudaMemset(lut, 0, 4096);
  run<0>("FFMA rrr", 1, out, cyc, lut); run<19>("FFMA imm", 1, out, cyc, lut); run<1>("FFMA2", 1, out, cyc, lut); run<2>("FADD", 1, out, cyc, lut); run<14>("FADD2", 1, out, cyc, lut);
  run<3>("FMUL", 1, out, cyc, lut); run<4>("PRMT", 1, out, cyc, lut); run<5>("VABSDIFF4", 1, out, cyc, lut); run<6>("IDP4A", 1, out, cyc, lut);
  run<7>("LEA/IMAD x4+c", 1, out, cyc, lut); run<12>("IMAD", 1, out, cyc, lut); run<13>("IADD+SHF+LOP", 3, out, cyc, lut);
  run<8>("LDS random + FADD", 2, out, cyc, lut); run<9>("LDS linear + FADD", 2, out, cyc, lut);
  run<10>("F2I.U8 + IADD", 2, out, cyc, lut); run<11>("I2F + FADD (+LOP)", 2, out, cyc, lut);
  run<20>("FFMA2 acc+=d*w.bcast", 1, out, cyc, lut); run<21>("FMUL + FFMA2 bcast", 2, out, cyc, lut); run<22>("FFMA acc+=d*w", 1, out, cyc, lut);
  run<23>("FADD2 bcast", 1, out, cyc, lut); run<24>("FMUL2 bcast", 1, out, cyc, lut);
  run<15>("FFMA + PRMT", 2, out, cyc, lut); run<16>("FFMA + IDP4A", 2, out, cyc, lut); run<17>("FFMA + VABSDIFF4", 2, out, cyc, lut); run<18>("FFMA2 + PRMT", 2, out, cyc, lut);
  return 0;
}
