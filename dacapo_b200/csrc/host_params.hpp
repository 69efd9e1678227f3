// Host-side parameter generation for libB200_HEVM.so: RNS primes, 2N-th roots, twiddle
// tables and RNS constants.  One-off setup work (the reference does the same inside
// seal::SEALContext, reference: lib/Runtime/SEAL_HEVM.cpp:46-59,93-99); nothing here runs
// on the hot path.  Independent of oracle/ by construction (different code, same maths).
//
// SEAL 4.0 rules restated: primes = the `count` largest primes
// p = 1 (mod 2N) below 2^bits, in ascending order (q_0 smallest, q_{L-1} = special prime);
// psi = the smallest primitive 2N-th root of unity mod p; tw[bitrev(k)] = psi^k.
#pragma once
#include "ntt_core.cuh"
#include <cstdio>
#include <cstdlib>
#include <vector>

namespace hp {
typedef unsigned __int128 u128;

inline u64 mul(u64 a, u64 b, u64 q) { return (u64)((u128)a * b % q); }
inline u64 pw(u64 b, u64 e, u64 q) {
  u64 r = 1;
  for (b %= q; e; e >>= 1, b = mul(b, b, q))
    if (e & 1) r = mul(r, b, q);
  return r;
}
inline u64 inv(u64 a, u64 q) { return pw(a, q - 2, q); }
inline bool miller_rabin(u64 n) {
  if (n < 4) return n == 2 || n == 3;
  if (!(n & 1)) return false;
  u64 d = n - 1;
  int s = 0;
  while (!(d & 1)) d >>= 1, ++s;
  static const u64 bases[] = {2, 325, 9375, 28178, 450775, 9780504, 1795265022}; // deterministic for 64-bit
  for (u64 a : bases) {
    u64 x = pw(a % n, d, n);
    if (x == 0 || x == 1 || x == n - 1) continue;
    bool witness = true;
    for (int i = 1; i < s && witness; i++) {
      x = mul(x, x, n);
      if (x == n - 1) witness = false;
    }
    if (witness) return false;
  }
  return true;
}
inline std::vector<u64> prime_chain(int logN, int bits, int count) {
  const u64 step = (u64)2 << logN;
  std::vector<u64> out(count);
  u64 cand = ((((u64)1 << bits) - 1) / step) * step + 1;
  for (int k = count - 1; k >= 0; cand -= step) {
    if (cand < ((u64)1 << (bits - 1))) {
      std::fprintf(stderr, "[b200-hevm] fatal: prime chain exhausted\n");
      std::abort();
    }
    if (miller_rabin(cand)) out[k--] = cand;
  }
  return out;
}
inline u64 smallest_primitive_root(int logN, u64 q) {
  const u64 n2 = (u64)2 << logN;
  u64 g = 0;
  for (u64 c = 2;; c++) {
    g = pw(c, (q - 1) / n2, q);
    if (pw(g, n2 >> 1, q) == q - 1) break;
  }
  // all primitive roots are the odd powers of g
  u64 best = g, cur = g, gg = mul(g, g, q);
  for (u64 k = 1; k < (n2 >> 1); k++) {
    cur = mul(cur, gg, q);
    if (cur < best) best = cur;
  }
  return best;
}
inline u32 revbits(u32 v, int bits) {
  u32 r = 0;
  for (int i = 0; i < bits; i++, v >>= 1) r = (r << 1) | (v & 1);
  return r;
}
inline Tw shoup(u64 v, u64 q) {
  Tw t;
  t.w = v;
  t.wq = (u64)(((u128)v << 64) / q);
  return t;
}

struct HostParams {
  int logN = 0, L = 0;
  size_t N = 0;
  std::vector<u64> q, psi;
  std::vector<Tw> tw, itw; // [L][N]
  // pass-B copies: twB[prime][row][TWB_ROW] holds the 255 twiddles row `row` needs, contiguous and in the (bank-conflict
  // free, padded) order the warp kernels stage them (entry twB_pos(k, g)  <-  tw[(rows << k) + (row << k) + g], k < 8, g < 2^k)
  std::vector<Tw> twB, itwB;
  void build_rowwise() {
    const size_t rows = N >> 8;
    twB.assign((size_t)L * rows * TWB_ROW, Tw{0, 0});
    itwB.assign((size_t)L * rows * TWB_ROW, Tw{0, 0});
    for (int i = 0; i < L; i++)
      for (size_t r = 0; r < rows; r++)
        for (int k = 0; k < 8; k++)
          for (size_t g = 0; g < ((size_t)1 << k); g++) {
            const size_t dst = ((size_t)i * rows + r) * TWB_ROW + (size_t)twB_pos(k, (int)g), src = (size_t)i * N + (rows << k) + (r << k) + g;
            twB[dst] = tw[src];
            itwB[dst] = itw[src];
          }
  }
  NttTables tab{};         // tw/itw pointers are filled by the owner (device or host copies)

  void build(int logn, int nprimes, int bits) {
    if (nprimes > HEVM_MAXL) {
      std::fprintf(stderr, "[b200-hevm] fatal: at most %d primes supported\n", HEVM_MAXL);
      std::abort();
    }
    if (bits != 60) {
      std::fprintf(stderr, "[b200-hevm] fatal: the kernels assume 60-bit primes (SEAL_HEVM.cpp:48-53)\n");
      std::abort();
    }
    logN = logn, L = nprimes, N = (size_t)1 << logn;
    q = prime_chain(logn, bits, nprimes);
    for (u64 p : q)
      if ((((u64)1 << 60) - p) >> 32) {
        std::fprintf(stderr, "[b200-hevm] fatal: prime too far below 2^60 for lazy folding\n");
        std::abort();
      }
    psi.resize(L);
    tw.resize((size_t)L * N);
    itw.resize((size_t)L * N);
    tab.logN = logN, tab.L = L;
    for (int i = 0; i < L; i++) {
      const u64 p = q[i];
      psi[i] = smallest_primitive_root(logn, p);
      const u64 ipsi = inv(psi[i], p);
      u64 f = 1, b = 1;
      for (size_t k = 0; k < N; k++) {
        size_t pos = revbits((u32)k, logn);
        tw[(size_t)i * N + pos] = shoup(f, p);
        itw[(size_t)i * N + pos] = shoup(b, p);
        f = mul(f, psi[i], p);
        b = mul(b, ipsi, p);
      }
      u128 full = ~(u128)0;
      u128 ratio = full / p; // floor(2^128/p) since p does not divide 2^128
      tab.mod[i].q = p;
      tab.mod[i].ratio_lo = (u64)ratio;
      tab.mod[i].ratio_hi = (u64)(ratio >> 64);
      tab.mod[i].delta = ((u64)1 << 60) - p; // lazy-range folding needs p = 2^60 - delta, delta < 2^32
      u64 ninv = inv((u64)N % p, p);
      tab.invn[i] = shoup(ninv, p);
      tab.invn_w[i] = shoup(mul(ninv, itw[(size_t)i * N + 1].w, p), p);
    }
    for (int a = 0; a < L; a++)
      for (int b = 0; b < L; b++)
        tab.qinv[a][b] = (a == b) ? Tw{0, 0} : shoup(inv(q[a] % q[b], q[b]), q[b]);
  }
};
} // namespace hp
