// =============================================================================
// oracle/ckks_oracle.hpp  --  TEST INFRASTRUCTURE ONLY (CPU oracle)
//
// A plain C++ CPU restatement of the arithmetic that the reference HEVM runtime
// (reference: lib/Runtime/SEAL_HEVM.cpp) delegates to Microsoft SEAL 4.0.0
// (third-party, tag 4.0.0, NOT vendored in the reference: CMakeLists.txt:60,
// README.md:65-74, versions.txt:4).  Each routine names the reference call site
// that reaches it and the SEAL routine whose published algorithm it restates.
//
// PARITY STATUS: "parity unpinned" at the SEAL boundary.  The reference holds no
// golden ciphertexts / KATs and SEAL cannot be built in
// this environment, so this oracle is pinned only against (i) independent
// Python big-integer restatements (tests/pyref.py), (ii) externally checkable
// constants (q_13 = 0xFFFFFFFFFFC0001 for N=2^15, all primes = 1 mod 2N) and
// (iii) algebraic identities (decrypt(encrypt(x)) ~ x, homomorphic identities).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
// reference legs may use this code.  The product (dacapo_b200/) never links,
// imports or calls it.
// =============================================================================
#pragma once
#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace orc {

using u64 = uint64_t;
using u128 = unsigned __int128;

[[noreturn]] inline void die(const char *msg) {
  std::fprintf(stderr, "[oracle] fatal: %s\n", msg);
  std::abort();
}

// ----------------------------------------------------------------------------
// Modular arithmetic (SEAL util/uintarithsmallmod.h semantics: all public
// results canonical in [0,q); Barrett with floor(2^128/q), Shoup operands for
// twiddles/scalars).  SURVEY A.2.3.
// ----------------------------------------------------------------------------
struct Modulus {
  u64 q = 0;
  u64 ratio_lo = 0, ratio_hi = 0; // floor(2^128 / q)
  explicit Modulus(u64 v = 0) : q(v) {
    if (v) {
      // floor(2^128/q) via 2^128 = (2^128-1) + 1 ; q never divides 2^128 (q odd > 1)
      u128 r = (~(u128)0) / v;
      ratio_lo = (u64)r;
      ratio_hi = (u64)(r >> 64);
    }
  }
};

inline u64 mulhi(u64 a, u64 b) { return (u64)(((u128)a * b) >> 64); }

// barrett_reduce_128: input < 2^128 -> [0,q)
inline u64 reduce128(u64 lo, u64 hi, const Modulus &m) {
  // quotient estimate = floor( (hi:lo) * ratio / 2^128 ), computed like SEAL's
  // barrett_reduce_128 (only the words that matter).
  u64 tmp1, tmp3, carry;
  u128 t = (u128)lo * m.ratio_lo;
  carry = (u64)(t >> 64);
  t = (u128)lo * m.ratio_hi;
  u64 tmp2_lo = (u64)t, tmp2_hi = (u64)(t >> 64);
  u128 s = (u128)tmp2_lo + carry;
  tmp1 = (u64)s;
  tmp3 = tmp2_hi + (u64)(s >> 64);
  t = (u128)hi * m.ratio_lo;
  tmp2_lo = (u64)t;
  tmp2_hi = (u64)(t >> 64);
  s = (u128)tmp1 + tmp2_lo;
  carry = tmp2_hi + (u64)(s >> 64);
  u64 qhat = hi * m.ratio_hi + tmp3 + carry;
  u64 r = lo - qhat * m.q;
  return r >= m.q ? r - m.q : r;
}
inline u64 reduce64(u64 x, const Modulus &m) {
  u64 qhat = mulhi(x, m.ratio_hi);
  u64 r = x - qhat * m.q;
  return r >= m.q ? r - m.q : r;
}
inline u64 mulmod(u64 a, u64 b, const Modulus &m) {
  u128 p = (u128)a * b;
  return reduce128((u64)p, (u64)(p >> 64), m);
}
inline u64 addmod(u64 a, u64 b, u64 q) {
  u64 s = a + b;
  return s >= q ? s - q : s;
}
inline u64 submod(u64 a, u64 b, u64 q) { return a >= b ? a - b : a + q - b; }
inline u64 negmod(u64 a, u64 q) { return a ? q - a : 0; }
inline u64 powmod(u64 b, u64 e, const Modulus &m) {
  u64 r = 1;
  while (e) {
    if (e & 1) r = mulmod(r, b, m);
    b = mulmod(b, b, m);
    e >>= 1;
  }
  return r;
}
inline u64 invmod(u64 a, const Modulus &m) { return powmod(a, m.q - 2, m); } // q prime

struct Shoup { // SEAL MultiplyUIntModOperand
  u64 w = 0, wq = 0;
  void set(u64 v, u64 q) {
    w = v;
    wq = (u64)((((u128)v) << 64) / q);
  }
};
// x*w mod q, lazy: result in [0,2q) for ANY 64-bit x
inline u64 shoup_lazy(u64 x, const Shoup &s, u64 q) { return x * s.w - mulhi(x, s.wq) * q; }
inline u64 shoup_mul(u64 x, const Shoup &s, u64 q) {
  u64 r = shoup_lazy(x, s, q);
  return r >= q ? r - q : r;
}

// ----------------------------------------------------------------------------
// Number theory: deterministic Miller-Rabin for 64-bit, SEAL get_primes order,
// minimal primitive 2N-th root.  SURVEY A.2.1 / A.2.4
// (SEAL util/numth.cpp: get_primes, try_minimal_primitive_root).
// ----------------------------------------------------------------------------
inline bool is_prime(u64 n) {
  if (n < 2) return false;
  for (u64 p : {2ull, 3ull, 5ull, 7ull, 11ull, 13ull, 17ull, 19ull, 23ull, 29ull, 31ull, 37ull}) {
    if (n % p == 0) return n == p;
  }
  u64 d = n - 1;
  int r = 0;
  while (!(d & 1)) d >>= 1, r++;
  Modulus m(n);
  for (u64 a : {2ull, 3ull, 5ull, 7ull, 11ull, 13ull, 17ull, 19ull, 23ull, 29ull, 31ull, 37ull}) {
    u64 x = powmod(a % n, d, m);
    if (x == 1 || x == n - 1) continue;
    bool comp = true;
    for (int i = 1; i < r; i++) {
      x = mulmod(x, x, m);
      if (x == n - 1) {
        comp = false;
        break;
      }
    }
    if (comp) return false;
  }
  return true;
}

// CoeffModulus::Create(N, {bits x count}) : scan downward from the largest
// value = 1 mod 2N below 2^bits; first found prime becomes the LAST modulus.
inline std::vector<u64> seal_primes(size_t N, int bits, int count) {
  u64 factor = 2 * (u64)N;
  u64 value = ((((u64)1) << bits) - 1) / factor * factor + 1;
  u64 lower = ((u64)1) << (bits - 1);
  std::vector<u64> found;
  while ((int)found.size() < count && value > lower) {
    if (is_prime(value)) found.push_back(value);
    value -= factor;
  }
  if ((int)found.size() < count) die("not enough primes");
  std::reverse(found.begin(), found.end()); // q_0 smallest ... q_{L-1} largest
  return found;
}

inline u64 minimal_primitive_root(u64 degree /*2N*/, const Modulus &m) {
  // any primitive 'degree'-th root: g = x^((q-1)/degree) with g^(degree/2) = -1
  u64 e = (m.q - 1) / degree;
  u64 g = 0;
  for (u64 x = 2;; x++) {
    g = powmod(x, e, m);
    if (powmod(g, degree / 2, m) == m.q - 1) break;
  }
  // minimum over all odd powers (all primitive roots)
  u64 g2 = mulmod(g, g, m), cur = g, best = g;
  for (u64 i = 0; i < degree / 2; i++) {
    if (cur < best) best = cur;
    cur = mulmod(cur, g2, m);
  }
  return best;
}

inline uint32_t bitrev(uint32_t x, int bits) {
  uint32_t r = 0;
  for (int i = 0; i < bits; i++) r |= ((x >> i) & 1u) << (bits - 1 - i);
  return r;
}

// ----------------------------------------------------------------------------
// Negacyclic NTT (SEAL util/ntt.cpp + dwthandler.h).  Forward: Cooley-Tukey,
// twiddle table w[bitrev(i)] = psi^i, natural in -> bit-reversed out:
// out[i] = f(psi^(2*bitrev(i)+1)).  Inverse: Gentleman-Sande + N^-1.
// ----------------------------------------------------------------------------
struct NttTable {
  size_t N = 0;
  int logN = 0;
  Modulus mod;
  u64 psi = 0;
  std::vector<Shoup> w, iw; // w[m+i]: forward twiddle of stage m group i; iw[m+i] = its inverse
  Shoup invN;
  void init(int logn, u64 q) {
    logN = logn;
    N = (size_t)1 << logn;
    mod = Modulus(q);
    psi = minimal_primitive_root(2 * N, mod);
    w.resize(N);
    iw.resize(N);
    u64 ipsi = invmod(psi, mod);
    u64 p = 1, ip = 1;
    for (size_t i = 0; i < N; i++) {
      size_t r = bitrev((uint32_t)i, logn);
      w[r].set(p, q);
      iw[r].set(ip, q);
      p = mulmod(p, psi, mod);
      ip = mulmod(ip, ipsi, mod);
    }
    invN.set(invmod((u64)N % q, mod), q);
  }
  // lazy forward: input [0,4q) -> output [0,4q)
  void forward_lazy(u64 *a) const {
    const u64 q = mod.q, two_q = 2 * q;
    size_t gap = N >> 1;
    for (size_t m = 1; m < N; m <<= 1, gap >>= 1) {
      for (size_t i = 0; i < m; i++) {
        const Shoup s = w[m + i];
        u64 *x = a + 2 * i * gap, *y = x + gap;
        for (size_t j = 0; j < gap; j++) {
          u64 u = x[j];
          u -= (u >= two_q) ? two_q : 0;
          u64 v = shoup_lazy(y[j], s, q);
          x[j] = u + v;
          y[j] = u + two_q - v;
        }
      }
    }
  }
  void forward(u64 *a) const {
    forward_lazy(a);
    const u64 q = mod.q, two_q = 2 * q;
    for (size_t i = 0; i < N; i++) {
      u64 v = a[i];
      v -= (v >= two_q) ? two_q : 0;
      v -= (v >= q) ? q : 0;
      a[i] = v;
    }
  }
  // inverse: input [0,2q) (canonical is fine) -> output [0,q)
  void inverse(u64 *a) const {
    const u64 q = mod.q, two_q = 2 * q;
    size_t gap = 1;
    for (size_t m = N >> 1; m >= 1; m >>= 1, gap <<= 1) {
      for (size_t i = 0; i < m; i++) {
        const Shoup s = iw[m + i];
        u64 *x = a + 2 * i * gap, *y = x + gap;
        for (size_t j = 0; j < gap; j++) {
          u64 u = x[j], v = y[j];
          u64 t = u + v;
          t -= (t >= two_q) ? two_q : 0;
          x[j] = t;
          y[j] = shoup_lazy(u + two_q - v, s, q);
        }
      }
    }
    for (size_t i = 0; i < N; i++) a[i] = shoup_mul(a[i], invN, q);
  }
};

// ----------------------------------------------------------------------------
// Deterministic counter-based sampler (our own spec, shared BY SPECIFICATION with
// the product so that keys / encryptions are reproducible on both sides; SEAL's
// Blake2xb PRNG is seeded from random_device and cannot be matched).
//   rnd(key, stream, idx) = 64-bit word (idx & 7) of the ChaCha20 block with the 256-bit `key`, nonce = stream and
//                           block counter = idx >> 3 (original 64-bit counter / 64-bit nonce layout; word j of a
//                           block = state[2j] | state[2j+1] << 32)  -- a keyed cryptographic PRF, so the public `a`
//                           polynomials reveal nothing about the streams that produced the secret key and the errors
//   uniform mod q : (rnd(2k) * 2^64 + rnd(2k+1)) mod q
//   ternary       : rnd(k) % 3 - 1                  (SEAL sample_poly_ternary)
//   cbd           : popc(w & 0x1FFFFF) - popc((w>>21) & 0x1FFFFF)  (SEAL sample_poly_cbd, 21+21 bits)
// Key material is derived from the 256-bit seed of hevm_params.bin; the encryption randomness (u, e0, e1) uses a
// SEPARATE 256-bit key that is fresh per VM unless a test pins it (hevmx_set_enc_counter), see oracle_hevm.cpp.
// ----------------------------------------------------------------------------
struct Seed256 {
  uint32_t k[8];
};
inline uint32_t rotl32(uint32_t v, int c) { return (v << c) | (v >> (32 - c)); }
inline void chacha20_block(const Seed256 &key, u64 nonce, u64 counter, u64 out[8]) {
  uint32_t x[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u, key.k[0], key.k[1], key.k[2], key.k[3], key.k[4], key.k[5], key.k[6],
                    key.k[7], (uint32_t)counter, (uint32_t)(counter >> 32), (uint32_t)nonce, (uint32_t)(nonce >> 32)};
  uint32_t in[16];
  for (int i = 0; i < 16; i++) in[i] = x[i];
#define ORC_QR(a, b, c, d)                                                                                             \
  x[a] += x[b], x[d] = rotl32(x[d] ^ x[a], 16), x[c] += x[d], x[b] = rotl32(x[b] ^ x[c], 12), x[a] += x[b],            \
      x[d] = rotl32(x[d] ^ x[a], 8), x[c] += x[d], x[b] = rotl32(x[b] ^ x[c], 7)
  for (int r = 0; r < 10; r++) {
    ORC_QR(0, 4, 8, 12), ORC_QR(1, 5, 9, 13), ORC_QR(2, 6, 10, 14), ORC_QR(3, 7, 11, 15);
    ORC_QR(0, 5, 10, 15), ORC_QR(1, 6, 11, 12), ORC_QR(2, 7, 8, 13), ORC_QR(3, 4, 9, 14);
  }
#undef ORC_QR
  for (int j = 0; j < 8; j++) out[j] = (u64)(x[2 * j] + in[2 * j]) | ((u64)(x[2 * j + 1] + in[2 * j + 1]) << 32);
}
struct Prf { // one stream of the sampler; consecutive indices reuse the cached block
  const Seed256 &key;
  u64 stream, cur = ~0ull, w[8];
  Prf(const Seed256 &k, u64 st) : key(k), stream(st) {}
  u64 operator()(u64 idx) {
    if ((idx >> 3) != cur) cur = idx >> 3, chacha20_block(key, stream, cur, w);
    return w[idx & 7];
  }
};
inline u64 uniform_mod(Prf &r, u64 k, const Modulus &m) {
  u64 hi = r(2 * k), lo = r(2 * k + 1);
  return (u64)((((u128)hi << 64) | lo) % m.q);
}
inline int ternary(Prf &r, u64 k) { return (int)(r(k) % 3) - 1; }
inline int cbd(Prf &r, u64 k) {
  u64 w = r(k);
  return __builtin_popcountll(w & 0x1FFFFF) - __builtin_popcountll((w >> 21) & 0x1FFFFF);
}
// stream ids
enum : u64 {
  ST_SK = 1,        // secret key (ternary)
  ST_PK_A = 2,      // public key a (uniform; +limb)
  ST_PK_E = 3,      // public key e (cbd)
  ST_KSK_BASE = 16, // key-switch keys: stream(key_id,digit,kind,limb)
  ST_ENC_BASE = 1ull << 40
};
inline u64 ksk_stream(u64 key_id, u64 digit, u64 kind /*0=a,1=e*/, u64 limb) {
  return (((key_id * 64 + digit) * 2 + kind) * 64 + limb) + (ST_KSK_BASE << 20);
}
inline u64 pk_a_stream(u64 limb) { return (ST_PK_A << 20) + limb; }
inline u64 enc_stream(u64 counter, u64 which /*0=u,1=e0,2=e1*/) { return ST_ENC_BASE + counter * 4 + which; }

// ----------------------------------------------------------------------------
// Context: primes, NTT tables, RNS constants, Galois tool, encoder tables, keys
// ----------------------------------------------------------------------------
struct Ct {
  int level = 0; // number of data limbs
  int size = 2;
  double scale = 1.0;
  std::vector<u64> d; // [size][level][N]
};
struct Pt {
  int level = 0;
  double scale = 1.0;
  std::vector<u64> d; // [level][N], NTT form
};
struct KswKey {
  std::vector<u64> d; // [digit J < L-1][K < 2][limb I < L][N]
};

struct Context {
  int logN = 0, L = 0; // L = number of primes incl. the special (last) one
  size_t N = 0;
  Seed256 seed{};     // key material
  Seed256 enc_seed{}; // encryption randomness (fresh per VM unless pinned by a test)
  std::vector<Modulus> q;
  std::vector<NttTable> ntt;
  // encoder tables (SEAL ckks.cpp CKKSEncoder ctor)
  std::vector<uint32_t> slot_index;             // matrix_reps_index_map, size N
  std::vector<std::complex<double>> fft_roots;  // root_powers_ (forward, decode)
  std::vector<std::complex<double>> ifft_roots; // inv_root_powers_ (inverse, encode)
  // keys
  std::vector<u64> sk;        // [L][N] NTT form
  std::vector<u64> pk;        // [2][L][N] NTT form
  KswKey relin;               // key_id 0
  std::map<u64, KswKey> gal;  // by galois element; key_id = 1 + (elt-1)/2 ... see keygen
  u64 enc_counter = 0;

  size_t max_level() const { return (size_t)L - 1; }

  void init_params(int logn, int nprimes, int bits, const Seed256 &sd) {
    logN = logn;
    L = nprimes;
    N = (size_t)1 << logn;
    seed = sd;
    enc_seed = sd;
    auto ps = seal_primes(N, bits, nprimes);
    q.clear();
    ntt.resize(nprimes);
    for (int i = 0; i < nprimes; i++) {
      q.emplace_back(ps[i]);
      ntt[i].init(logn, ps[i]);
    }
    init_encoder();
  }

  // -------- Galois tool (SEAL util/galois.cpp) ------------------------------
  u64 galois_elt_from_step(int step) const {
    u64 m = 2 * N;
    if (step == 0) return m - 1;
    bool neg = step < 0;
    u64 pos = (u64)std::abs(step);
    if (pos >= (N >> 1)) die("rotation step count too large");
    u64 s = neg ? (N >> 1) - pos : pos;
    u64 elt = 1;
    for (u64 i = 0; i < s; i++) elt = (elt * 3) & (m - 1);
    return elt;
  }
  std::vector<u64> galois_elts_all() const {
    u64 m = 2 * N;
    std::vector<u64> r;
    r.push_back(m - 1);
    u64 pos = 3, neg = 0;
    for (u64 x = 1; x < m; x += 2)
      if (((x * 3) & (m - 1)) == 1) {
        neg = x;
        break;
      }
    for (int i = 0; i < logN - 1; i++) {
      r.push_back(pos);
      pos = (pos * pos) & (m - 1);
      r.push_back(neg);
      neg = (neg * neg) & (m - 1);
    }
    return r;
  }
  // apply_galois_ntt permutation table: out[i] = in[table[i]]
  mutable std::map<u64, std::vector<uint32_t>> galois_table_cache; // GaloisTool keeps its tables too
  const std::vector<uint32_t> &galois_table(u64 elt) const {
    auto it = galois_table_cache.find(elt);
    if (it != galois_table_cache.end()) return it->second;
    std::vector<uint32_t> &t = galois_table_cache[elt];
    t.resize(N);
    for (size_t i = 0; i < N; i++) {
      u64 rev = 2 * (u64)bitrev((uint32_t)i, logN) + 1;
      u64 raw = ((elt * rev) >> 1) & (N - 1);
      t[i] = bitrev((uint32_t)raw, logN);
    }
    return t;
  }

  // -------- encoder tables ----------------------------------------------------
  void init_encoder() {
    size_t slots = N >> 1;
    u64 m = 2 * N;
    slot_index.resize(N);
    u64 pos = 1;
    for (size_t i = 0; i < slots; i++) {
      u64 i1 = (pos - 1) >> 1, i2 = (m - pos - 1) >> 1;
      slot_index[i] = bitrev((uint32_t)i1, logN);
      slot_index[slots | i] = bitrev((uint32_t)i2, logN);
      pos = (pos * 3) & (m - 1);
    }
    // ComplexRoots(2N): roots_[i] = polar(1, 2*pi*i/m) for i <= m/8, 8-fold symmetry otherwise
    const double PI = 3.1415926535897932384626433832795028842;
    std::vector<std::complex<double>> base(m / 8 + 1);
    for (size_t i = 0; i <= m / 8; i++) base[i] = std::polar<double>(1.0, 2 * PI * (double)i / (double)m);
    auto mirror = [](std::complex<double> a) { return std::complex<double>(a.imag(), a.real()); };
    std::function<std::complex<double>(size_t)> get_root = [&](size_t idx) -> std::complex<double> {
      idx &= m - 1;
      if (idx <= m / 8) return base[idx];
      if (idx <= m / 4) return mirror(base[m / 4 - idx]);
      if (idx <= m / 2) return -std::conj(get_root(m / 2 - idx));
      if (idx <= 3 * m / 4) return -get_root(idx - m / 2);
      return std::conj(get_root(m - idx));
    };
    fft_roots.assign(N, {0, 0});
    ifft_roots.assign(N, {0, 0});
    for (size_t i = 1; i < N; i++) {
      fft_roots[i] = get_root(bitrev((uint32_t)i, logN));
      ifft_roots[i] = std::conj(get_root((size_t)bitrev((uint32_t)(i - 1), logN) + 1));
    }
  }

  // -------- sampling helpers ----------------------------------------------------
  void sample_ternary_rns(const Seed256 &key, u64 stream, int limbs, u64 *out) const { // coefficient form
    Prf r(key, stream);
    for (size_t k = 0; k < N; k++) {
      int t = ternary(r, k);
      for (int i = 0; i < limbs; i++) out[(size_t)i * N + k] = t < 0 ? q[i].q - 1 : (u64)t;
    }
  }
  void sample_cbd_rns(const Seed256 &key, u64 stream, int limbs, u64 *out) const {
    Prf r(key, stream);
    for (size_t k = 0; k < N; k++) {
      int t = cbd(r, k);
      for (int i = 0; i < limbs; i++) out[(size_t)i * N + k] = t < 0 ? q[i].q - (u64)(-t) : (u64)t;
    }
  }

  // -------- key generation (SEAL keygenerator.cpp / util/rlwe.cpp) ------------
  // symmetric zero encryption at key level in NTT form: (c0, c1) = (-(a s + e), a)
  void encrypt_zero_symmetric(u64 a_stream_base, u64 e_stream, u64 *c0, u64 *c1) const {
    std::vector<u64> e((size_t)L * N);
    sample_cbd_rns(seed, e_stream, L, e.data());
    for (int i = 0; i < L; i++) {
      u64 *a = c1 + (size_t)i * N;
      Prf ra(seed, a_stream_base + i);
      for (size_t k = 0; k < N; k++) a[k] = uniform_mod(ra, k, q[i]); // sampled in NTT form
      u64 *ei = e.data() + (size_t)i * N;
      ntt[i].forward(ei);
      const u64 *s = sk.data() + (size_t)i * N;
      u64 *o = c0 + (size_t)i * N;
      for (size_t k = 0; k < N; k++) o[k] = negmod(addmod(mulmod(a[k], s[k], q[i]), ei[k], q[i].q), q[i].q);
    }
  }
  // generate_one_kswitch_key for every digit J: key[J] = enc_zero_sym ; c0[limb J] += (p mod q_J) * newkey[J]
  void make_ksk(u64 key_id, const u64 *newkey /*[L][N] NTT*/, KswKey &out) const {
    int D = L - 1;
    out.d.assign((size_t)D * 2 * L * N, 0);
    u64 p = q[L - 1].q;
    for (int J = 0; J < D; J++) {
      u64 *c0 = out.d.data() + ((size_t)J * 2 + 0) * L * N;
      u64 *c1 = out.d.data() + ((size_t)J * 2 + 1) * L * N;
      encrypt_zero_symmetric(ksk_stream(key_id, J, 0, 0), ksk_stream(key_id, J, 1, 0), c0, c1);
      u64 factor = reduce64(p, q[J]);
      u64 *t = c0 + (size_t)J * N;
      const u64 *nk = newkey + (size_t)J * N;
      for (size_t k = 0; k < N; k++) t[k] = addmod(t[k], mulmod(nk[k], factor, q[J]), q[J].q);
    }
  }
  static u64 galois_key_id(u64 elt) { return 1 + ((elt - 1) >> 1); } // GaloisKeys::get_index(elt)+1
  void keygen() {
    sk.assign((size_t)L * N, 0);
    sample_ternary_rns(seed, ST_SK << 20, L, sk.data());
    for (int i = 0; i < L; i++) ntt[i].forward(sk.data() + (size_t)i * N);
    pk.assign((size_t)2 * L * N, 0);
    encrypt_zero_symmetric(pk_a_stream(0), ST_PK_E << 20, pk.data(), pk.data() + (size_t)L * N);
    // relin: new key = s^2
    std::vector<u64> s2((size_t)L * N);
    for (int i = 0; i < L; i++)
      for (size_t k = 0; k < N; k++) s2[(size_t)i * N + k] = mulmod(sk[(size_t)i * N + k], sk[(size_t)i * N + k], q[i]);
    make_ksk(0, s2.data(), relin);
    // galois keys (create_galois_keys default set): new key = apply_galois_ntt(sk, elt)
    // HEVM_GALOIS_STEPS="1,-2,64": only these rotation steps get a key (SEAL KeyGenerator::create_galois_keys(steps));
    // unset = the default set of create_galois_keys() used by the reference (SEAL_HEVM.cpp:82-83)
    if (const char *e = std::getenv("HEVM_GALOIS_STEPS")) {
      for (const char *p = e; *p;) {
        char *end = nullptr;
        long st = std::strtol(p, &end, 10);
        if (end == p) break;
        make_galois_key(galois_elt_from_step((int)st));
        p = (*end == ',') ? end + 1 : end;
      }
    } else {
      for (u64 elt : galois_elts_all()) make_galois_key(elt);
    }
  }
  void make_galois_key(u64 elt) {
    if (gal.count(elt)) return;
    const auto &tab = galois_table(elt);
    std::vector<u64> rs((size_t)L * N);
    for (int i = 0; i < L; i++)
      for (size_t k = 0; k < N; k++) rs[(size_t)i * N + k] = sk[(size_t)i * N + tab[k]];
    make_ksk(galois_key_id(elt), rs.data(), gal[elt]);
  }

  // -------- CKKS encode (SEAL CKKSEncoder::encode_internal), `level` limbs ---
  // reference call site: SEAL_HEVM.cpp:256-267 (encode at top level then drop limbs:
  // limb-wise independent, so encoding `level` limbs directly is bit-identical).
  void fft_from_rev(std::complex<double> *v, double fix) const { // DWTHandler::transform_from_rev + scalar
    size_t n = N, gap = 1, m = n >> 1, ri = 0;
    for (; m > 1; m >>= 1, gap <<= 1) {
      size_t off = 0;
      for (size_t i = 0; i < m; i++, off += gap << 1) {
        std::complex<double> r = ifft_roots[++ri];
        for (size_t j = 0; j < gap; j++) {
          auto u = v[off + j], w = v[off + gap + j];
          v[off + j] = cadd(u, w);
          v[off + gap + j] = cmul(csub(u, w), r);
        }
      }
    }
    std::complex<double> r = ifft_roots[++ri];
    std::complex<double> sr(r.real() * fix, r.imag() * fix);
    for (size_t j = 0; j < gap; j++) {
      auto u = v[j], w = v[gap + j];
      auto s = cadd(u, w);
      v[j] = std::complex<double>(s.real() * fix, s.imag() * fix);
      v[gap + j] = cmul(csub(u, w), sr);
    }
  }
  void fft_to_rev(std::complex<double> *v) const { // DWTHandler::transform_to_rev
    size_t n = N, gap = n >> 1, ri = 0;
    for (size_t m = 1; m < n; m <<= 1, gap >>= 1) {
      size_t off = 0;
      for (size_t i = 0; i < m; i++, off += gap << 1) {
        std::complex<double> r = fft_roots[++ri];
        for (size_t j = 0; j < gap; j++) {
          auto u = v[off + j];
          auto w = cmul(v[off + gap + j], r);
          v[off + j] = cadd(u, w);
          v[off + gap + j] = csub(u, w);
        }
      }
    }
  }
  static std::complex<double> cadd(std::complex<double> a, std::complex<double> b) {
    return {a.real() + b.real(), a.imag() + b.imag()};
  }
  static std::complex<double> csub(std::complex<double> a, std::complex<double> b) {
    return {a.real() - b.real(), a.imag() - b.imag()};
  }
  static std::complex<double> cmul(std::complex<double> a, std::complex<double> b) {
    // naive product, separately rounded (compile with -ffp-contract=off)
    return {a.real() * b.real() - a.imag() * b.imag(), a.real() * b.imag() + a.imag() * b.real()};
  }

  // values: N/2 real slots
  void encode(const double *values, int level, double scale, Pt &out) const {
    size_t slots = N >> 1;
    std::vector<std::complex<double>> cv(N);
    for (size_t i = 0; i < slots; i++) {
      cv[slot_index[i]] = std::complex<double>(values[i], 0.0);
      cv[slot_index[slots + i]] = std::complex<double>(values[i], -0.0);
    }
    double fix = scale / (double)N;
    fft_from_rev(cv.data(), fix);
    double maxc = 0;
    for (size_t i = 0; i < N; i++) maxc = std::max(maxc, std::fabs(cv[i].real()));
    int bits = (int)std::ceil(std::log2(std::max(maxc, 1.0)));
    out.level = level;
    out.scale = scale;
    out.d.assign((size_t)level * N, 0);
    const double two64 = std::pow(2.0, 64);
    if (bits <= 64) {
      for (size_t i = 0; i < N; i++) {
        double c = std::round(cv[i].real());
        bool negv = std::signbit(c);
        u64 cu = (u64)std::fabs(c);
        for (int j = 0; j < level; j++) {
          u64 r = reduce64(cu, q[j]);
          out.d[(size_t)j * N + i] = negv ? negmod(r, q[j].q) : r;
        }
      }
    } else if (bits <= 128) {
      for (size_t i = 0; i < N; i++) {
        double c = std::round(cv[i].real());
        bool negv = std::signbit(c);
        c = std::fabs(c);
        u64 lo = (u64)std::fmod(c, two64), hi = (u64)(c / two64);
        for (int j = 0; j < level; j++) {
          u64 r = reduce128(lo, hi, q[j]);
          out.d[(size_t)j * N + i] = negv ? negmod(r, q[j].q) : r;
        }
      }
    } else {
      die("encode: coefficients wider than 128 bits are not supported by the oracle");
    }
    for (int j = 0; j < level; j++) ntt[j].forward(out.d.data() + (size_t)j * N);
  }

  // -------- CKKS decode (SEAL CKKSEncoder::decode_internal) ------------------
  void decode(const Pt &p, double *out /*N/2*/) const {
    int l = p.level;
    std::vector<u64> c(p.d);
    for (int j = 0; j < l; j++) ntt[j].inverse(c.data() + (size_t)j * N);
    // CRT compose into l-word integers (unique representative in [0,Q))
    std::vector<std::vector<u64>> punct(l, std::vector<u64>(l, 0)); // Q/q_j (l words)
    std::vector<u64> Q(l, 0), invp(l);
    {
      // Q = prod q_j
      std::vector<u64> acc(l, 0);
      acc[0] = 1;
      for (int j = 0; j < l; j++) mp_mul_word(acc, q[j].q);
      Q = acc;
      for (int j = 0; j < l; j++) {
        std::vector<u64> a(l, 0);
        a[0] = 1;
        for (int k = 0; k < l; k++)
          if (k != j) mp_mul_word(a, q[k].q);
        punct[j] = a;
        u64 r = 1;
        for (int k = 0; k < l; k++)
          if (k != j) r = mulmod(r, reduce64(q[k].q, q[j]), q[j]);
        invp[j] = invmod(r, q[j]);
      }
    }
    std::vector<u64> half(Q); // upper_half_threshold = (Q+1)/2
    {
      // (Q + 1) >> 1
      u64 carry = 1;
      for (int k = 0; k < l; k++) {
        u64 s = half[k] + carry;
        carry = s < carry ? 1 : 0;
        half[k] = s;
      }
      for (int k = 0; k < l; k++) half[k] = (half[k] >> 1) | (k + 1 < l ? (half[k + 1] << 63) : (carry << 63));
    }
    std::vector<std::complex<double>> res(N);
    double inv_scale = 1.0 / p.scale;
    const double two64 = std::pow(2.0, 64);
    std::vector<u64> val(l + 1), tmp(l + 1);
    for (size_t i = 0; i < N; i++) {
      std::fill(val.begin(), val.end(), 0);
      for (int j = 0; j < l; j++) {
        u64 t = mulmod(c[(size_t)j * N + i], invp[j], q[j]);
        // tmp = punct[j] * t  (l+1 words) ; val = (val + tmp) mod Q
        u64 carry = 0;
        for (int k = 0; k < l; k++) {
          u128 pr = (u128)punct[j][k] * t + carry;
          tmp[k] = (u64)pr;
          carry = (u64)(pr >> 64);
        }
        tmp[l] = carry;
        mp_mod_small_multiple(tmp, Q, l); // tmp < q_j * Q/q_j*... reduce by repeated compare (tmp < Q*? see fn)
        // val += tmp (both < Q), subtract Q if >= Q
        u64 cy = 0;
        for (int k = 0; k < l; k++) {
          u128 s = (u128)val[k] + tmp[k] + cy;
          val[k] = (u64)s;
          cy = (u64)(s >> 64);
        }
        val[l] = cy;
        if (mp_geq(val, Q, l)) mp_sub(val, Q, l);
      }
      double r = 0.0;
      bool upper = mp_geq_n(val, half, l);
      double s64 = inv_scale;
      if (upper) {
        for (int j = 0; j < l; j++, s64 *= two64) {
          if (val[j] > Q[j]) {
            u64 diff = val[j] - Q[j];
            r += diff ? (double)diff * s64 : 0.0;
          } else {
            u64 diff = Q[j] - val[j];
            r -= diff ? (double)diff * s64 : 0.0;
          }
        }
      } else {
        for (int j = 0; j < l; j++, s64 *= two64) {
          u64 cc = val[j];
          r += cc ? (double)cc * s64 : 0.0;
        }
      }
      res[i] = std::complex<double>(r, 0.0);
    }
    fft_to_rev(res.data());
    size_t slots = N >> 1;
    for (size_t i = 0; i < slots; i++) out[i] = res[slot_index[i]].real();
  }
  // --- tiny multiprecision helpers (little-endian words) ---
  static void mp_mul_word(std::vector<u64> &a, u64 w) {
    u64 carry = 0;
    for (size_t k = 0; k < a.size(); k++) {
      u128 p = (u128)a[k] * w + carry;
      a[k] = (u64)p;
      carry = (u64)(p >> 64);
    }
  }
  static bool mp_geq(const std::vector<u64> &a /*l+1*/, const std::vector<u64> &b /*l*/, int l) {
    if (a[l]) return true;
    for (int k = l - 1; k >= 0; k--) {
      if (a[k] != b[k]) return a[k] > b[k];
    }
    return true;
  }
  static bool mp_geq_n(const std::vector<u64> &a, const std::vector<u64> &b, int l) {
    for (int k = l - 1; k >= 0; k--) {
      if (a[k] != b[k]) return a[k] > b[k];
    }
    return true;
  }
  static void mp_sub(std::vector<u64> &a /*l+1*/, const std::vector<u64> &b /*l*/, int l) {
    u64 borrow = 0;
    for (int k = 0; k < l; k++) {
      u128 d = (u128)a[k] - b[k] - borrow;
      a[k] = (u64)d;
      borrow = (u64)(d >> 64) ? 1 : 0;
    }
    a[l] -= borrow;
  }
  // tmp = punct_j * t with t < q_j  =>  tmp < Q ; nothing to do (kept for clarity)
  static void mp_mod_small_multiple(std::vector<u64> &, const std::vector<u64> &, int) {}

  // -------- encrypt / decrypt (SEAL encryptor.cpp / decryptor.cpp / rlwe.cpp) --
  // divide_and_round_q_last_ntt_inplace on one poly with `k` limbs (result in first k-1)
  void divide_round_last_ntt(u64 *poly, size_t stride, int k) const {
    u64 *last = poly + (size_t)(k - 1) * stride;
    const Modulus &ql = q[k - 1];
    std::vector<u64> r(last, last + N);
    ntt[k - 1].inverse(r.data());
    u64 half = ql.q >> 1;
    for (size_t i = 0; i < N; i++) r[i] = addmod(r[i], half, ql.q);
    std::vector<u64> t(N);
    for (int i = 0; i < k - 1; i++) {
      const Modulus &qi = q[i];
      u64 neg_half = qi.q - reduce64(half, qi);
      for (size_t c = 0; c < N; c++) t[c] = reduce64(r[c], qi) + neg_half; // < 2 q_i
      ntt[i].forward(t.data());                                            // canonical
      u64 inv = invmod(reduce64(ql.q, qi), qi);
      u64 *x = poly + (size_t)i * stride;
      for (size_t c = 0; c < N; c++) x[c] = mulmod(submod(x[c], t[c], qi.q), inv, qi);
    }
  }
  void encrypt_zero_asym(int limbs /*use q_0..q_{limbs-1}*/, u64 counter, std::vector<u64> &c /*[2][limbs][N]*/) const {
    c.assign((size_t)2 * limbs * N, 0);
    std::vector<u64> u((size_t)limbs * N);
    sample_ternary_rns(enc_seed, enc_stream(counter, 0), limbs, u.data());
    for (int i = 0; i < limbs; i++) {
      ntt[i].forward(u.data() + (size_t)i * N);
      for (int j = 0; j < 2; j++) {
        const u64 *pkj = pk.data() + ((size_t)j * L + i) * N;
        u64 *o = c.data() + ((size_t)j * limbs + i) * N;
        for (size_t k = 0; k < N; k++) o[k] = mulmod(u[(size_t)i * N + k], pkj[k], q[i]);
      }
    }
    for (int j = 0; j < 2; j++) {
      sample_cbd_rns(enc_seed, enc_stream(counter, 1 + j), limbs, u.data());
      for (int i = 0; i < limbs; i++) {
        ntt[i].forward(u.data() + (size_t)i * N);
        u64 *o = c.data() + ((size_t)j * limbs + i) * N;
        for (size_t k = 0; k < N; k++) o[k] = addmod(o[k], u[(size_t)i * N + k], q[i].q);
      }
    }
  }
  // Encryptor::encrypt(plain): zero-encrypt at level+1 limbs, mod-switch (divide & round by
  // q_level) to `level`, then c0 += plain.   Reference call sites: SEAL_HEVM.cpp:333,444
  void encrypt(const Pt &p, Ct &out) {
    int l = p.level;
    std::vector<u64> z;
    u64 counter = enc_counter++;
    encrypt_zero_asym(l + 1, counter, z);
    out.level = l;
    out.size = 2;
    out.scale = p.scale;
    out.d.assign((size_t)2 * l * N, 0);
    for (int j = 0; j < 2; j++) {
      u64 *poly = z.data() + (size_t)j * (l + 1) * N;
      divide_round_last_ntt(poly, N, l + 1);
      std::memcpy(out.d.data() + (size_t)j * l * N, poly, (size_t)l * N * sizeof(u64));
    }
    for (int i = 0; i < l; i++)
      for (size_t k = 0; k < N; k++) out.d[(size_t)i * N + k] = addmod(out.d[(size_t)i * N + k], p.d[(size_t)i * N + k], q[i].q);
  }
  // Decryptor::decrypt (ckks_decrypt): c0 + c1*s  (+ c2*s^2 for size 3)
  void decrypt(const Ct &c, Pt &out) const {
    int l = c.level;
    out.level = l;
    out.scale = c.scale;
    out.d.assign((size_t)l * N, 0);
    for (int i = 0; i < l; i++) {
      const u64 *s = sk.data() + (size_t)i * N;
      for (size_t k = 0; k < N; k++) {
        u64 acc = c.d[((size_t)0 * l + i) * N + k];
        u64 sp = s[k];
        for (int j = 1; j < c.size; j++) {
          acc = addmod(acc, mulmod(c.d[((size_t)j * l + i) * N + k], sp, q[i]), q[i].q);
          sp = mulmod(sp, s[k], q[i]);
        }
        out.d[(size_t)i * N + k] = acc;
      }
    }
  }

  // -------- Evaluator (SEAL evaluator.cpp) -------------------------------------
  // Evaluator::add / add_plain / negate  (call sites SEAL_HEVM.cpp:302,309,278)
  void add(const Ct &a, const Ct &b, Ct &o) const {
    if (a.level != b.level) die("add: level mismatch");
    Ct r;
    r.level = a.level;
    r.size = std::max(a.size, b.size);
    r.scale = a.scale;
    r.d.assign((size_t)r.size * r.level * N, 0);
    for (int j = 0; j < r.size; j++)
      for (int i = 0; i < r.level; i++)
        for (size_t k = 0; k < N; k++) {
          size_t ix = ((size_t)j * r.level + i) * N + k;
          u64 x = j < a.size ? a.d[ix] : 0, y = j < b.size ? b.d[ix] : 0;
          r.d[ix] = addmod(x, y, q[i].q);
        }
    o = std::move(r);
  }
  void add_plain(const Ct &a, const Pt &p, Ct &o) const {
    if (a.level != p.level) die("add_plain: level mismatch");
    Ct r = a;
    for (int i = 0; i < r.level; i++)
      for (size_t k = 0; k < N; k++) r.d[(size_t)i * N + k] = addmod(r.d[(size_t)i * N + k], p.d[(size_t)i * N + k], q[i].q);
    o = std::move(r);
  }
  void negate(const Ct &a, Ct &o) const {
    Ct r = a;
    for (int j = 0; j < r.size; j++)
      for (int i = 0; i < r.level; i++)
        for (size_t k = 0; k < N; k++) {
          size_t ix = ((size_t)j * r.level + i) * N + k;
          r.d[ix] = negmod(r.d[ix], q[i].q);
        }
    o = std::move(r);
  }
  // Evaluator::multiply_plain (NTT form): SEAL_HEVM.cpp:322
  void multiply_plain(const Ct &a, const Pt &p, Ct &o) const {
    if (a.level != p.level) die("multiply_plain: level mismatch");
    Ct r = a;
    for (int j = 0; j < r.size; j++)
      for (int i = 0; i < r.level; i++)
        for (size_t k = 0; k < N; k++) {
          size_t ix = ((size_t)j * r.level + i) * N + k;
          r.d[ix] = mulmod(r.d[ix], p.d[(size_t)i * N + k], q[i]);
        }
    r.scale = a.scale * p.scale;
    o = std::move(r);
  }
  // Evaluator::multiply (ckks_multiply, 2x2 -> 3): SEAL_HEVM.cpp:315
  void multiply(const Ct &a, const Ct &b, Ct &o) const {
    if (a.level != b.level || a.size != 2 || b.size != 2) die("multiply: bad operands");
    Ct r;
    int l = a.level;
    r.level = l;
    r.size = 3;
    r.scale = a.scale * b.scale;
    r.d.assign((size_t)3 * l * N, 0);
    for (int i = 0; i < l; i++)
      for (size_t k = 0; k < N; k++) {
        u64 a0 = a.d[((size_t)0 * l + i) * N + k], a1 = a.d[((size_t)1 * l + i) * N + k];
        u64 b0 = b.d[((size_t)0 * l + i) * N + k], b1 = b.d[((size_t)1 * l + i) * N + k];
        r.d[((size_t)0 * l + i) * N + k] = mulmod(a0, b0, q[i]);
        r.d[((size_t)1 * l + i) * N + k] = addmod(mulmod(a0, b1, q[i]), mulmod(a1, b0, q[i]), q[i].q);
        r.d[((size_t)2 * l + i) * N + k] = mulmod(a1, b1, q[i]);
      }
    o = std::move(r);
  }
  // Evaluator::switch_key_inplace: ct(size>=2, first two polys) += KS(target)
  void switch_key(Ct &ct, const u64 *target /*[l][N] NTT*/, const KswKey &key) const {
    int l = ct.level;
    int sp = L - 1; // special prime index
    std::vector<u64> t((size_t)l * N);
    std::memcpy(t.data(), target, t.size() * sizeof(u64));
    for (int j = 0; j < l; j++) ntt[j].inverse(t.data() + (size_t)j * N);
    // acc[K][I], I in {0..l-1, sp}
    std::vector<u64> prod((size_t)2 * (l + 1) * N);
    std::vector<u64> tn(N);
    std::vector<u128> acc0(N), acc1(N);
    for (int Ii = 0; Ii <= l; Ii++) {
      int I = (Ii == l) ? sp : Ii;
      std::fill(acc0.begin(), acc0.end(), (u128)0);
      std::fill(acc1.begin(), acc1.end(), (u128)0);
      for (int J = 0; J < l; J++) {
        const u64 *opnd;
        if (I == J) {
          opnd = target + (size_t)J * N;
        } else {
          const u64 *src = t.data() + (size_t)J * N;
          if (q[J].q <= q[I].q)
            std::memcpy(tn.data(), src, N * sizeof(u64));
          else
            for (size_t k = 0; k < N; k++) tn[k] = reduce64(src[k], q[I]);
          ntt[I].forward_lazy(tn.data()); // [0,4q) like SEAL
          opnd = tn.data();
        }
        const u64 *k0 = key.d.data() + (((size_t)J * 2 + 0) * L + I) * N;
        const u64 *k1 = key.d.data() + (((size_t)J * 2 + 1) * L + I) * N;
        for (size_t k = 0; k < N; k++) {
          acc0[k] += (u128)opnd[k] * k0[k];
          acc1[k] += (u128)opnd[k] * k1[k];
        }
      }
      u64 *p0 = prod.data() + ((size_t)0 * (l + 1) + Ii) * N;
      u64 *p1 = prod.data() + ((size_t)1 * (l + 1) + Ii) * N;
      for (size_t k = 0; k < N; k++) {
        p0[k] = reduce128((u64)acc0[k], (u64)(acc0[k] >> 64), q[I]);
        p1[k] = reduce128((u64)acc1[k], (u64)(acc1[k] >> 64), q[I]);
      }
    }
    // mod-down by the special prime with rounding, accumulate into ct
    const Modulus &qp = q[sp];
    u64 half = qp.q >> 1;
    for (int K = 0; K < 2; K++) {
      u64 *last = prod.data() + ((size_t)K * (l + 1) + l) * N;
      ntt[sp].inverse(last);
      for (size_t k = 0; k < N; k++) last[k] = reduce64(last[k] + half, qp);
      for (int i = 0; i < l; i++) {
        const Modulus &qi = q[i];
        u64 fix = qi.q - reduce64(half, qi);
        for (size_t k = 0; k < N; k++) tn[k] = (qp.q > qi.q ? reduce64(last[k], qi) : last[k]) + fix;
        ntt[i].forward(tn.data());
        u64 inv = invmod(reduce64(qp.q, qi), qi);
        const u64 *pi = prod.data() + ((size_t)K * (l + 1) + i) * N;
        u64 *dst = ct.d.data() + ((size_t)K * l + i) * N;
        for (size_t k = 0; k < N; k++) dst[k] = addmod(dst[k], mulmod(submod(pi[k], tn[k], qi.q), inv, qi), qi.q);
      }
    }
  }
  // Evaluator::relinearize_inplace (size 3 -> 2): SEAL_HEVM.cpp:316
  void relinearize(Ct &c) const {
    if (c.size != 3) return;
    int l = c.level;
    std::vector<u64> c2(c.d.begin() + (size_t)2 * l * N, c.d.end());
    c.d.resize((size_t)2 * l * N);
    c.size = 2;
    switch_key(c, c2.data(), relin);
  }
  // Evaluator::apply_galois_inplace (CKKS, NTT form)
  void apply_galois(Ct &c, u64 elt) {
    int l = c.level;
    if (!gal.count(elt)) die("Galois key not present");
    const auto &tab = galois_table(elt);
    std::vector<u64> t0((size_t)l * N), t1((size_t)l * N);
    for (int i = 0; i < l; i++)
      for (size_t k = 0; k < N; k++) {
        t0[(size_t)i * N + k] = c.d[((size_t)0 * l + i) * N + tab[k]];
        t1[(size_t)i * N + k] = c.d[((size_t)1 * l + i) * N + tab[k]];
      }
    std::memcpy(c.d.data(), t0.data(), t0.size() * sizeof(u64));
    std::memset(c.d.data() + (size_t)l * N, 0, (size_t)l * N * sizeof(u64));
    switch_key(c, t1.data(), gal.at(elt));
  }
  static std::vector<int> naf(int value) { // SEAL util/numth.h naf == reference Support.h:10-27
    std::vector<int> res;
    bool sign = value < 0;
    value = std::abs(value);
    for (int i = 0; value; i++) {
      int zi = (value & 1) ? 2 - (value & 3) : 0;
      value = (value - zi) >> 1;
      if (zi) res.push_back((sign ? -zi : zi) * (1 << i));
    }
    return res;
  }
  // Evaluator::rotate_vector -> rotate_internal   (SEAL_HEVM.cpp:273)
  void rotate_inplace(Ct &c, int steps) {
    if (steps == 0) return;
    u64 elt = galois_elt_from_step(steps);
    if (gal.count(elt)) {
      apply_galois(c, elt);
      return;
    }
    auto terms = naf(steps);
    if (terms.size() == 1) die("Galois key not present");
    for (int s : terms)
      if ((size_t)std::abs(s) != (N >> 1)) rotate_inplace(c, s);
  }
  // Evaluator::rescale_to_next: SEAL_HEVM.cpp:283
  void rescale(const Ct &a, Ct &o) const {
    int l = a.level;
    if (l < 2) die("rescale: already at the last level");
    Ct tmp = a;
    Ct r;
    r.level = l - 1;
    r.size = a.size;
    r.scale = a.scale / (double)q[l - 1].q;
    r.d.assign((size_t)r.size * (l - 1) * N, 0);
    for (int j = 0; j < a.size; j++) {
      u64 *poly = tmp.d.data() + (size_t)j * l * N;
      divide_round_last_ntt(poly, N, l);
      std::memcpy(r.d.data() + (size_t)j * (l - 1) * N, poly, (size_t)(l - 1) * N * sizeof(u64));
    }
    o = std::move(r);
  }
  // Evaluator::mod_switch_to_next (CKKS: drop last limb): SEAL_HEVM.cpp:289-291
  void mod_switch(const Ct &a, Ct &o) const {
    int l = a.level;
    if (l < 2) die("mod_switch: already at the last level");
    Ct r;
    r.level = l - 1;
    r.size = a.size;
    r.scale = a.scale;
    r.d.assign((size_t)r.size * (l - 1) * N, 0);
    for (int j = 0; j < a.size; j++)
      std::memcpy(r.d.data() + (size_t)j * (l - 1) * N, a.d.data() + (size_t)j * l * N, (size_t)(l - 1) * N * sizeof(u64));
    o = std::move(r);
  }
};

} // namespace orc
