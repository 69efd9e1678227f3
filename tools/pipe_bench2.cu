// Mixed-pipe issue-rate microbenchmark for sm_100a: can IMAD.WIDE (fma pipe) and IADD3/LOP3 (alu pipe)
// co-issue, and how close do alternative 64-bit Shoup butterfly formulations get to the issue limit?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/pipe_bench2.cu -o /tmp/pb2 && /tmp/pb2
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
typedef uint64_t u64;
typedef uint32_t u32;
#define ITER 16384

// MIX: number of IADD3 issued per IMAD.WIDE (0 = wide only, 9 = iadd only)
template <int MIX> __global__ void kmix(u64 *out, u32 a0, u32 b0) {
  u64 acc[6];
  u32 x[6], y[6], z[6];
  for (int i = 0; i < 6; i++) {
    acc[i] = threadIdx.x * 977 + i;
    x[i] = a0 + i + threadIdx.x;
    y[i] = x[i] * 3;
    z[i] = x[i] * 5;
  }
  for (int it = 0; it < ITER; it++) {
#pragma unroll
    for (int i = 0; i < 6; i++) {
      if (MIX != 9) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[i]) : "r"(x[i]), "r"(y[i]));
      if (MIX == 1 || MIX == 9) asm volatile("add.u32 %0, %0, %1;" : "+r"(z[i]) : "r"(x[i]));
      if (MIX == 2) asm volatile("add.u32 %0, %0, %1; add.u32 %2, %2, %1;" : "+r"(z[i]), "+r"(y[i]) : "r"(x[i]));
      if (MIX == 3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(z[i]) : "r"(x[i]), "r"(y[i]));
    }
  }
  u64 s = 0;
  for (int i = 0; i < 6; i++) s += acc[i] + x[i] + y[i] + z[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

struct Tw {
  u64 w, wq;
};
__device__ __forceinline__ u64 mulhi_exact(u64 a, u64 b) { return __umul64hi(a, b); }
// approximate high product: drops the lo*lo partial product (result in [exact-2, exact])
__device__ __forceinline__ u64 mulhi_approx(u64 a, u64 b) {
  const u32 al = (u32)a, ah = (u32)(a >> 32), bl = (u32)b, bh = (u32)(b >> 32);
  const u64 t1 = (u64)ah * bl, t2 = (u64)al * bh;
  return (u64)ah * bh + (t1 >> 32) + (t2 >> 32);
}
// V: 0 = C reference form, 1 = approximate quotient, 2 = delta form of Q*q, 3 = both
template <int V> __global__ void kbf(u64 *out, u64 q, u32 delta, const Tw *tws) {
  u64 x[8];
  Tw t[4];
  for (int i = 0; i < 8; i++) x[i] = threadIdx.x * 977 + i;
  for (int i = 0; i < 4; i++) t[i] = tws[(threadIdx.x + i) & 63];
  const u64 q2 = 2 * q;
  for (int it = 0; it < ITER / 4; it++) {
#pragma unroll
    for (int i = 0; i < 8; i += 2) {
      const u64 y = x[i + 1];
      const u64 Q = (V & 1) ? mulhi_approx(y, t[i / 2].wq) : mulhi_exact(y, t[i / 2].wq);
      u64 v;
      if (V & 2) {
        // -Q*q = Q*delta - (Q << 60)
        v = y * t[i / 2].w + (u64)(u32)Q * delta + (((u64)((u32)(Q >> 32) * delta - ((u32)Q << 28))) << 32);
      } else {
        v = y * t[i / 2].w + Q * (0 - q);
      }
      const u64 u = x[i];
      x[i] = u + v;
      x[i + 1] = u + q2 - v;
    }
  }
  u64 s = 0;
  for (int i = 0; i < 8; i++) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F> float timeit(F f) {
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  float ms = 0;
  for (int rep = 0; rep < 3; rep++) {
    cudaEventRecord(a);
    f();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    cudaEventElapsedTime(&ms, a, b);
  }
  return ms;
}

int main() {
  u64 *d;
  cudaMalloc(&d, 148 * 16 * 256 * 8);
  Tw h[64];
  for (int i = 0; i < 64; i++) h[i] = Tw{0x123456789abcdefULL + i * 7919, 0xfedcba9876543210ULL - i * 104729};
  Tw *dt;
  cudaMalloc(&dt, sizeof(h));
  cudaMemcpy(dt, h, sizeof(h), cudaMemcpyHostToDevice);
  int clk;
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  const int grid = 148 * 8;
  const double warps = 148.0 * 8 * 8;
  const u64 q = 0xFFFFFFFFFFC0001ULL;
  {
    const char *nm[] = {"IMAD.WIDE only", "1 WIDE : 1 IADD", "1 WIDE : 2 IADD", "1 WIDE : 1 LOP3", "IADD only"};
    float ms[5];
    ms[0] = timeit([&] { kmix<0><<<grid, 256>>>(d, 1, 3); });
    ms[1] = timeit([&] { kmix<1><<<grid, 256>>>(d, 1, 3); });
    ms[2] = timeit([&] { kmix<2><<<grid, 256>>>(d, 1, 3); });
    ms[3] = timeit([&] { kmix<3><<<grid, 256>>>(d, 1, 3); });
    ms[4] = timeit([&] { kmix<9><<<grid, 256>>>(d, 1, 3); });
    const int per[] = {1, 2, 3, 2, 1};
    for (int i = 0; i < 5; i++) {
      double ninst = warps * ITER * 6 * per[i];
      printf("%-18s %.3f warp-instr/clk/SMSP  (%.3f ms)\n", nm[i], ninst / (ms[i] * 1e-3) / (clk * 1e3) / 148 / 4, ms[i]);
    }
  }
  {
    const char *nm[] = {"shoup exact", "approx quotient", "delta form", "approx + delta"};
    float ms[4];
    ms[0] = timeit([&] { kbf<0><<<grid, 256>>>(d, q, (u32)((1ULL << 60) - q), dt); });
    ms[1] = timeit([&] { kbf<1><<<grid, 256>>>(d, q, (u32)((1ULL << 60) - q), dt); });
    ms[2] = timeit([&] { kbf<2><<<grid, 256>>>(d, q, (u32)((1ULL << 60) - q), dt); });
    ms[3] = timeit([&] { kbf<3><<<grid, 256>>>(d, q, (u32)((1ULL << 60) - q), dt); });
    for (int i = 0; i < 4; i++) {
      double nb = warps * (ITER / 4) * 4;
      printf("butterfly %-16s %.1f clk per warp-butterfly per SMSP, %.2f G bfly/s (%.3f ms)\n", nm[i],
             (ms[i] * 1e-3) * (clk * 1e3) * 148 * 4 / nb, nb * 32 / (ms[i] * 1e-3) / 1e9, ms[i]);
    }
  }
  {
    // occupancy sweep: exact Shoup butterfly (4 independent chains per thread), w warps per SMSP
    for (int w = 1; w <= 16; w = (w < 4 ? w + 1 : w * 2)) {
      const int g = 148 * w; // CTAs of 128 threads = 1 warp per SMSP each
      float ms = timeit([&] { kbf<0><<<g, 128>>>(d, q, (u32)((1ULL << 60) - q), dt); });
      double nb = 148.0 * w * 4 * (ITER / 4) * 4;
      printf("occupancy %2d warps/SMSP: %.1f clk per warp-butterfly per SMSP (%.3f ms)\n", w, (ms * 1e-3) * (clk * 1e3) * 148 * 4 / nb, ms);
    }
  }
  return 0;
}
