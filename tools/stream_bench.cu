// Read-bandwidth probe for the key-switch key access pattern on sm_100a (one key = [13][2][14][N] u64, N = 2^15).
//   A: linear grid-stride read of one key (95 MB)
//   B: one CTA per (limb I, row r) reading its 26 rows of 2 KB, all loads in flight at once
//   C: as B, but in dependent batches of 4 digits (what a register-staged dot product does)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/stream_bench.cu -o tools/_sb
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
#define N 32768
#define L 14
#define DG 13
__global__ void kA(const ulonglong2 *p, size_t n, u64 *out) {
  u64 acc = 0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    ulonglong2 v = __ldg(p + i);
    acc ^= v.x ^ v.y;
  }
  if (acc == 0x1234567) out[0] = acc;
}
template <int BATCH> __global__ void __launch_bounds__(128) kB(const u64 *key, u64 *out) {
  const int job = blockIdx.x, I = L - 1 - (job >> 7), r = job & 127, tid = threadIdx.x;
  const u64 *kp = key + (size_t)I * N + r * 256 + 2 * tid;
  u64 acc = 0;
  for (int J0 = 0; J0 < DG; J0 += BATCH) {
    ulonglong2 v[BATCH][2];
#pragma unroll
    for (int b = 0; b < BATCH; b++)
      if (J0 + b < DG) {
        v[b][0] = __ldg(reinterpret_cast<const ulonglong2 *>(kp + ((size_t)(J0 + b) * 2 + 0) * L * N));
        v[b][1] = __ldg(reinterpret_cast<const ulonglong2 *>(kp + ((size_t)(J0 + b) * 2 + 1) * L * N));
      }
#pragma unroll
    for (int b = 0; b < BATCH; b++)
      if (J0 + b < DG) acc += v[b][0].x * v[b][1].y + v[b][0].y * v[b][1].x + acc * 3;
  }
  if (acc == 0x1234567) out[0] = acc;
}
int main() {
  const size_t kw = (size_t)DG * 2 * L * N;
  const int NK = 8;
  u64 *keys, *out;
  cudaMalloc(&keys, kw * 8 * NK);
  cudaMalloc(&out, 64);
  cudaMemset(keys, 1, kw * 8 * NK);
  cudaEvent_t a, b;
  cudaEventCreate(&a), cudaEventCreate(&b);
  auto run = [&](const char *nm, auto f) {
    float best = 1e9;
    for (int rep = 0; rep < 16; rep++) {
      cudaEventRecord(a);
      f(keys + (size_t)(rep % NK) * kw);
      cudaEventRecord(b);
      cudaEventSynchronize(b);
      float ms;
      cudaEventElapsedTime(&ms, a, b);
      if (rep >= 8 && ms < best) best = ms;
    }
    printf("%-40s %.1f us  %.2f TB/s\n", nm, best * 1e3, kw * 8 / (best * 1e-3) / 1e12);
  };
  run("A linear, 148x8 CTAs x 256", [&](u64 *k) { kA<<<148 * 8, 256>>>((const ulonglong2 *)k, kw / 2, out); });
  run("A linear, 148x16 CTAs x 256", [&](u64 *k) { kA<<<148 * 16, 256>>>((const ulonglong2 *)k, kw / 2, out); });
  run("B per-(I,row) CTA, all 26 loads at once", [&](u64 *k) { kB<13><<<L * 128, 128>>>(k, out); });
  run("C per-(I,row) CTA, batches of 4 digits", [&](u64 *k) { kB<4><<<L * 128, 128>>>(k, out); });
  run("C per-(I,row) CTA, batches of 2 digits", [&](u64 *k) { kB<2><<<L * 128, 128>>>(k, out); });
  run("C per-(I,row) CTA, batches of 1 digit", [&](u64 *k) { kB<1><<<L * 128, 128>>>(k, out); });
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
