// Butterfly-formulation microbenchmark for sm_100a (round 2): can a 60-bit lazy modular multiply be done in fewer
// half-rate integer instructions than the Shoup form (1 mul.hi.u64 + 2 mul.lo.u64 = 10-11 IMAD-class instructions)?
//
// Candidate ("split-32"): for a fixed twiddle w keep w2 = w * 2^32 mod q next to it.  For ANY 64-bit y = yh*2^32 + yl
//     y*w == yl*w + yh*w2 (mod q)                       4 IMAD.WIDE.U32 (32 x 60-bit products), < 2^93
// and with q = 2^60 - delta the 93-bit sum S is folded once at bit 61:  r = (S mod 2^61) + (S >> 61) * 2*delta,
// one more IMAD.WIDE.U32, r < 2^61 + 2^58 < 2.26 q.  5 IMAD.WIDE + 1 SHF + 1 LOP3 + a few full-rate IADD3.
//
// Prints clocks per warp-butterfly per SMSP for each formulation and checks every variant against a 128-bit reference.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/bfly_bench.cu -o tools/_bfly && tools/_bfly
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
typedef uint64_t u64;
typedef uint32_t u32;
#define ITER 8192

struct Tw {
  u64 w, wq; // wq: Shoup quotient (V0) or w * 2^32 mod q (V1, V2)
};

__device__ __forceinline__ u64 shoup_lazy(u64 x, Tw t, u64 q) { return x * t.w + __umul64hi(x, t.wq) * (0 - q); }

// r == y * w (mod q), r < 2^61 + 2^58, for any 64-bit y.  t.wq = w * 2^32 mod q.  d2 = 2 * delta.
__device__ __forceinline__ u64 mulw61(u64 y, Tw t, u32 d2) {
  const u32 yl = (u32)y, yh = (u32)(y >> 32);
  const u32 wl = (u32)t.w, wh = (u32)(t.w >> 32), vl = (u32)t.wq, vh = (u32)(t.wq >> 32);
  u64 p1, p3, h;
  asm("mul.wide.u32 %0, %1, %2;" : "=l"(p1) : "r"(yl), "r"(wl));
  asm("mul.wide.u32 %0, %1, %2;" : "=l"(p3) : "r"(yh), "r"(vl));
  asm("mul.wide.u32 %0, %1, %2;" : "=l"(h) : "r"(yl), "r"(wh));
  asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(h) : "r"(yh), "r"(vh)); // < 2^61 - 2^29
  // S = h * 2^32 + p1 + p3 ;  B = S >> 32 , lo = (u32) S
  u32 lo, blo, bhi;
  asm("{\n\t"
      ".reg .u32 t;\n\t"
      "add.cc.u32 %0, %3, %5;\n\t"      // lo = p1.lo + p3.lo
      "addc.cc.u32 t, %4, %6;\n\t"      // t = p1.hi + p3.hi + c      (33 bits)
      "addc.u32 %2, 0, 0;\n\t"          // carry of that
      "add.cc.u32 %1, t, %7;\n\t"       // B.lo = t + h.lo
      "addc.u32 %2, %2, %8;\n\t"        // B.hi = carry + h.hi + c
      "}"
      : "=&r"(lo), "=&r"(blo), "=&r"(bhi)
      : "r"((u32)p1), "r"((u32)(p1 >> 32)), "r"((u32)p3), "r"((u32)(p3 >> 32)), "r"((u32)h), "r"((u32)(h >> 32)));
  u32 sh;
  asm("shf.r.clamp.b32 %0, %1, %2, 29;" : "=r"(sh) : "r"(blo), "r"(bhi)); // S >> 61 (< 2^32)
  u64 r = ((u64)(blo & 0x1FFFFFFFu) << 32) | lo;
  asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(r) : "r"(sh), "r"(d2));
  return r;
}
// same, folded exactly at bit 60 (r < 2^60 + 2^59): one more IMAD for the 33rd bit of S >> 60
__device__ __forceinline__ u64 mulw60(u64 y, Tw t, u32 d) {
  const u32 yl = (u32)y, yh = (u32)(y >> 32);
  const u32 wl = (u32)t.w, wh = (u32)(t.w >> 32), vl = (u32)t.wq, vh = (u32)(t.wq >> 32);
  u64 p1, p3, h;
  asm("mul.wide.u32 %0, %1, %2;" : "=l"(p1) : "r"(yl), "r"(wl));
  asm("mul.wide.u32 %0, %1, %2;" : "=l"(p3) : "r"(yh), "r"(vl));
  asm("mul.wide.u32 %0, %1, %2;" : "=l"(h) : "r"(yl), "r"(wh));
  asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(h) : "r"(yh), "r"(vh));
  u32 lo, blo, bhi;
  asm("{\n\t"
      ".reg .u32 t;\n\t"
      "add.cc.u32 %0, %3, %5;\n\t"
      "addc.cc.u32 t, %4, %6;\n\t"
      "addc.u32 %2, 0, 0;\n\t"
      "add.cc.u32 %1, t, %7;\n\t"
      "addc.u32 %2, %2, %8;\n\t"
      "}"
      : "=&r"(lo), "=&r"(blo), "=&r"(bhi)
      : "r"((u32)p1), "r"((u32)(p1 >> 32)), "r"((u32)p3), "r"((u32)(p3 >> 32)), "r"((u32)h), "r"((u32)(h >> 32)));
  u32 sh;
  asm("shf.r.wrap.b32 %0, %1, %2, 28;" : "=r"(sh) : "r"(blo), "r"(bhi)); // low 32 bits of S >> 60
  const u32 top = bhi >> 28;                                               // bit 32 of S >> 60
  u64 r = ((u64)((blo & 0x0FFFFFFFu) + top * d) << 32) | lo;
  asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(r) : "r"(sh), "r"(d));
  return r;
}

template <int V> __global__ void kbf(u64 *out, u64 q, u32 delta, const Tw *tws) {
  u64 x[8];
  Tw t[4];
  for (int i = 0; i < 8; i++) x[i] = threadIdx.x * 977 + i;
  for (int i = 0; i < 4; i++) t[i] = tws[(threadIdx.x + i) & 63];
  const u64 q2 = 2 * q, q3 = 3 * q;
  for (int it = 0; it < ITER / 4; it++) {
#pragma unroll
    for (int i = 0; i < 8; i += 2) {
      const u64 y = x[i + 1], u = x[i];
      if (V == 0) {
        const u64 v = shoup_lazy(y, t[i / 2], q);
        x[i] = u + v, x[i + 1] = u + q2 - v;
      } else if (V == 1) {
        const u64 v = mulw61(y, t[i / 2], 2 * delta);
        x[i] = u + v, x[i + 1] = u + q3 - v;
      } else {
        const u64 v = mulw60(y, t[i / 2], delta);
        x[i] = u + v, x[i + 1] = u + q2 - v;
      }
    }
  }
  u64 s = 0;
  for (int i = 0; i < 8; i++) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// correctness: r == y*w mod q and r below the stated bound, random inputs
template <int V> __global__ void kcheck(u64 q, u32 delta, const u64 *ys, const Tw *tws, int n, unsigned long long *bad, u64 *maxr) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const u64 y = ys[i];
  const Tw t = tws[i];
  const u64 r = (V == 1) ? mulw61(y, t, 2 * delta) : mulw60(y, t, delta);
  const unsigned __int128 p = (unsigned __int128)(y % q) * t.w;
  const u64 ref = (u64)(p % q);
  if (r % q != ref) atomicAdd(bad, 1ull);
  atomicMax((unsigned long long *)maxr, (unsigned long long)r);
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
static u64 rnd() {
  static u64 s = 88172645463325252ull;
  s ^= s << 13, s ^= s >> 7, s ^= s << 17;
  return s;
}
int main() {
  const u64 q = 0xFFFFFFFFFFC0001ULL - 0; // q_13 ; also test the prime with the largest delta below
  const u64 qs[2] = {0xFFFFFFFFFFC0001ULL, (1ull << 60) - 25427967ull};
  u64 *d;
  cudaMalloc(&d, 148 * 16 * 256 * 8);
  int clk;
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  // correctness
  for (int pi = 0; pi < 2; pi++) {
    const u64 qq = qs[pi];
    const int n = 1 << 20;
    u64 *hy = (u64 *)malloc(n * 8);
    Tw *ht = (Tw *)malloc(n * sizeof(Tw));
    for (int i = 0; i < n; i++) {
      hy[i] = (i & 7) == 0 ? ~0ull - (rnd() & 0xFFFF) : rnd();
      const u64 w = (i & 15) == 1 ? qq - 1 - (rnd() & 0xFF) : rnd() % qq;
      ht[i].w = w;
      ht[i].wq = (u64)((((unsigned __int128)w) << 32) % qq);
    }
    u64 *dy, *dmax;
    Tw *dt;
    unsigned long long *dbad;
    cudaMalloc(&dy, n * 8), cudaMalloc(&dt, n * sizeof(Tw)), cudaMalloc(&dbad, 8), cudaMalloc(&dmax, 8);
    cudaMemcpy(dy, hy, n * 8, cudaMemcpyHostToDevice), cudaMemcpy(dt, ht, n * sizeof(Tw), cudaMemcpyHostToDevice);
    for (int v = 1; v <= 2; v++) {
      cudaMemset(dbad, 0, 8), cudaMemset(dmax, 0, 8);
      if (v == 1)
        kcheck<1><<<n / 256, 256>>>(qq, (u32)((1ull << 60) - qq), dy, dt, n, dbad, dmax);
      else
        kcheck<2><<<n / 256, 256>>>(qq, (u32)((1ull << 60) - qq), dy, dt, n, dbad, dmax);
      unsigned long long bad = 0;
      u64 mx = 0;
      cudaMemcpy(&bad, dbad, 8, cudaMemcpyDeviceToHost), cudaMemcpy(&mx, dmax, 8, cudaMemcpyDeviceToHost);
      printf("check q=%llx variant %d: %llu mismatches of %d, max r = %.4f * 2^60\n", (unsigned long long)qq, v, bad, n, (double)mx / (double)(1ull << 60));
    }
    cudaFree(dy), cudaFree(dt), cudaFree(dbad), cudaFree(dmax);
    free(hy), free(ht);
  }
  Tw h[64];
  for (int i = 0; i < 64; i++) h[i] = Tw{0x123456789abcdefULL + i * 7919, 0x0edcba9876543210ULL - i * 104729};
  Tw *dt;
  cudaMalloc(&dt, sizeof(h));
  cudaMemcpy(dt, h, sizeof(h), cudaMemcpyHostToDevice);
  const int grid = 148 * 8;
  const double warps = 148.0 * 8 * 8;
  const char *nm[] = {"shoup (mul.hi + 2 mul.lo)", "split-32, fold at 2^61", "split-32, fold at 2^60"};
  for (int v = 0; v < 3; v++) {
    float ms = timeit([&] {
      if (v == 0) kbf<0><<<grid, 256>>>(d, q, (u32)((1ULL << 60) - q), dt);
      if (v == 1) kbf<1><<<grid, 256>>>(d, q, (u32)((1ULL << 60) - q), dt);
      if (v == 2) kbf<2><<<grid, 256>>>(d, q, (u32)((1ULL << 60) - q), dt);
    });
    const double nb = warps * (ITER / 4) * 4;
    printf("butterfly %-28s %.1f clk per warp-butterfly per SMSP, %.2f G bfly/s (%.3f ms) => %.3f us per 2^15 limb-NTT\n", nm[v],
           (ms * 1e-3) * (clk * 1e3) * 148 * 4 / nb, nb * 32 / (ms * 1e-3) / 1e9, ms, 245760.0 / (nb * 32 / (ms * 1e-3)) * 1e6);
  }
  return 0;
}
