// Integer pipe throughput microbenchmark for sm_100a: IMAD.WIDE.U32, IMAD (lo), IADD3(+X) and a mixed
// Shoup butterfly.  Prints warp-instructions per clock per SM.  nvcc -arch=sm_100a -O3 tools/pipe_bench.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
typedef uint64_t u64; typedef uint32_t u32;
#define ITER 4096
template <int OP> __global__ void k(u64 *out, u32 a0, u32 b0) {
  u64 acc[8]; u32 x[8];
  for (int i = 0; i < 8; i++) { acc[i] = threadIdx.x * 977 + i; x[i] = a0 + i + threadIdx.x; }
  u32 b = b0 | 1;
  for (int it = 0; it < ITER; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      if (OP == 0) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[i]) : "r"(x[i]), "r"(b));
      if (OP == 1) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(b), "r"(a0));
      if (OP == 2) asm volatile("add.cc.u32 %0, %0, %1; addc.u32 %2, %2, %3;" : "+r"(x[i]), "+r"(b) : "r"(a0), "r"(b0)); // 2 IADD3
      if (OP == 3) asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(b));
      if (OP == 4) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(b), "r"(a0));
      if (OP == 5) asm volatile("shf.r.wrap.b32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(b), "r"(a0));
    }
  }
  u64 s = 0;
  for (int i = 0; i < 8; i++) s += acc[i] + x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + b;
}
struct Tw { u64 w, wq; };
__global__ void kbf(u64 *out, u64 q, Tw t) {
  u64 x[8];
  for (int i = 0; i < 8; i++) x[i] = threadIdx.x * 977 + i;
  u64 q2 = 2 * q;
  for (int it = 0; it < ITER / 4; it++) {
#pragma unroll
    for (int i = 0; i < 8; i += 2) {
      u64 y = x[i + 1];
      u64 v = y * t.w + __umul64hi(y, t.wq) * (0 - q);
      u64 u = x[i];
      x[i] = u + v; x[i + 1] = u + q2 - v;
    }
  }
  u64 s = 0;
  for (int i = 0; i < 8; i++) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  u64 *d; cudaMalloc(&d, 148 * 16 * 256 * 8);
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  const char *names[] = {"IMAD.WIDE.U32", "IMAD(lo)", "IADD3+IADD3.X pair", "IMAD.HI.U32", "LOP3", "SHF"};
  for (int op = 0; op < 6; op++) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int rep = 0; rep < 2; rep++) {
      cudaEventRecord(a);
      int grid = 148 * 8;
      switch (op) { case 0: k<0><<<grid,256>>>(d,1,3); break; case 1: k<1><<<grid,256>>>(d,1,3); break; case 2: k<2><<<grid,256>>>(d,1,3); break;
                    case 3: k<3><<<grid,256>>>(d,1,3); break; case 4: k<4><<<grid,256>>>(d,1,3); break; case 5: k<5><<<grid,256>>>(d,1,3); break; }
      cudaEventRecord(b); cudaEventSynchronize(b);
    }
    float ms; cudaEventElapsedTime(&ms, a, b);
    double ninst = (double)148 * 8 * 8 /*warps*/ * ITER * 8 * (op == 2 ? 2 : 1);
    printf("%-22s %.2f warp-instr/clk/SM (assuming %.0f MHz)  %.3f ms\n", names[op], ninst / (ms * 1e-3) / (clk * 1e3) / 148, clk / 1e3, ms);
  }
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  Tw t{0x123456789abcdefULL, 0xfedcba9876543210ULL};
  for (int rep = 0; rep < 2; rep++) { cudaEventRecord(a); kbf<<<148 * 8, 256>>>(d, 0xFFFFFFFFFFC0001ULL, t); cudaEventRecord(b); cudaEventSynchronize(b); }
  float ms; cudaEventElapsedTime(&ms, a, b);
  double nb = (double)148 * 8 * 256 * (ITER / 4) * 4;
  printf("Shoup butterflies: %.1f G/s  => %.3f us per 2^15-point NTT (245760 butterflies) if perfectly parallel\n", nb / (ms * 1e-3) / 1e9, 245760.0 / (nb / (ms * 1e-3)) * 1e6);
  return 0;
}
