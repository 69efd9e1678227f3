// Warp-job bodies of the NTT-based kernels (N = 2^LOGA rows x 256 cols; LOGA = 6..9 <=> N = 2^14..2^17).
// One warp executes one job (a (limb, tile) or (limb, row) pair); the key-switch inner-product
// kernel uses one 4-warp CTA per (output prime, row) and splits the digit loop over its warps.
// See ntt_core.cuh for the schedule, kernels.cu for the __global__ wrappers.
//
// Fusions (reference call sites SEAL_HEVM.cpp:273,283,315-316):
//   body_intt_B  : [Galois gather | ct x ct product d2 = a1*b1 | plain load] + inverse pass B
//   body_intt_A  : inverse pass A + N^-1 [+ "add q_last/2" rounding of rescale / mod-down]
//   body_fwd_A   : [mod-up reduction "t_J mod q_I" | rounding fix-up] + forward pass A
//   body_fwd_B   : forward pass B + one of
//        CANON    canonical store (plain NTT)
//        MODDOWN  (acc - u) * p^-1 + addend   with addend = permuted c0 | tensor product d0/d1
//        RESCALE  (c - u) * q_last^-1
//   body_mac_*   : forward pass B of every digit into shared memory (radix-2^30 split), then the key-switch
//                  inner product per coefficient over all digits (carry-free 64-bit column accumulators,
//                  one Barrett at the end): the l x (l+1) matrix of NTT'd digits never exists in HBM.
// Every body takes an optional target range (t0/nt/i_end) for the limb-sharded multi-GPU key switch.
#pragma once
#include "ntt_core.cuh"

enum { LD_PLAIN = 0, LD_GALOIS = 1, LD_PRODUCT = 2, LD_DECRYPT = 3 /* c0 + c1 * s: src = c1, src2 = s, c0 = c0 */ };
enum { PRE_NONE = 0, PRE_MODUP = 1, PRE_ROUND = 2 };
enum { EPI_CANON = 0, EPI_MAC = 1, EPI_MODDOWN_GALOIS = 2, EPI_MODDOWN_RELIN = 3, EPI_RESCALE = 4 };

#define MAC_WARPS 4
// shared memory of one MAC CTA (words): staged twiddles | MAC_WARPS NTT tiles | MAC_WARPS staged rows | l transformed digit rows
#define TILE_B_WORDS 272 // 256 values + 256/16 padding
#define MAC_TW_WORDS (2 * TWB_ROW) // the staged (padded) twiddle row
#define MAC_ROW_WORDS 256 // per warp: staged pass-A row of the next digit
#define MAC_FIXED_WORDS (MAC_TW_WORDS + MAC_WARPS * TILE_B_WORDS + MAC_WARPS * MAC_ROW_WORDS)
// transformed digit rows: 256 values + two unused words after every 16 (the 64-byte lane stride of the layout-C
// stores would otherwise put four lanes of a quarter-warp on the same banks; 16-byte alignment is kept)
#define MAC_XROW 288
HD int xpad(int i) { return i + ((i >> 4) << 1); }
HD size_t mac_smem_words(int l) { return MAC_FIXED_WORDS + (size_t)MAC_XROW * l; }
// special-prime CTAs, phase 3 (after the last CTA barrier everything but the two accumulator rows in tiles[0, 512) is
// dead): warp K transforms row K with a twiddle table and an exchange tile of its own --
//   K = 1: the CTA's table, tile rowbufs[512, 784);  K = 0: table tiles[512, 512 + 2 TWB_ROW) (runs 64 words into
//   rowbufs), tile rowbufs[64, 336)
HD u64 *mac_tail_tile(u64 *rowbufs, int K) { return rowbufs + (K ? 512 : 64); }
HD Tw *mac_tail_tw(u64 *tiles, Tw *tw_s, int K) { return K ? tw_s : reinterpret_cast<Tw *>(tiles + 512); }
static_assert(512 + 2 * TWB_ROW <= MAC_WARPS * TILE_B_WORDS + 64, "tail table of warp 0 overlaps its tile");
static_assert(64 + TILE_B_WORDS <= 512 && 512 + TILE_B_WORDS <= MAC_WARPS * MAC_ROW_WORDS, "tail tiles do not fit the row buffers");
// radix-2^30 split of a value < 2^60 + 2^36: low 30 bits in the low word, the rest in the high word.  Key-switch keys
// are stored this way, and the transformed digits are converted to it, so that the key inner product is four
// 32x32->64 multiply-adds per term into four 64-bit column accumulators with no carries at all
// (every partial product is < 2^60 + 2^31, 15 of them fit in 64 bits).
#define RADIX30_MASK 0x3FFFFFFFull
HD u64 split30(u64 v) { return (v & RADIX30_MASK) | ((v >> 30) << 32); }
HD u64 join30(u64 w) { return (w & 0xFFFFFFFFull) + ((w >> 32) << 30); }
#define MAC_FLUSH_DIGITS 15

HD Tw *warp_tw(u64 *sm) { return reinterpret_cast<Tw *>(sm + WARP_TILE_WORDS); }
template <int LOGA> HD constexpr bool pass_a_needs_fold(int pre) { return pre == 2 /*PRE_ROUND*/ || LOGA >= 8; }

HD void load8_stream(const u64 *p, u64 (&v)[8]) {
  ldg_stream4(p, v[0], v[1], v[2], v[3]);
  ldg_stream4(p + 4, v[4], v[5], v[6], v[7]);
}
HD void load8_ro(const u64 *p, u64 (&v)[8]) { // key material
  ldg_ro4(p, v[0], v[1], v[2], v[3]);
  ldg_ro4(p + 4, v[4], v[5], v[6], v[7]);
}
HD void store8(u64 *p, const u64 (&v)[8]) {
  stg4(p, v[0], v[1], v[2], v[3]);
  stg4(p + 4, v[4], v[5], v[6], v[7]);
}

// ---------------------------------------------------------------------------------------------
// inverse pass B.  limbs: `nl` limbs, limb k uses prime (prime0 + k*pstep); src limb pitch = N.
// ---------------------------------------------------------------------------------------------
struct ArgsInttB {
  const NttTables *T;
  const u64 *src;  // PLAIN/GALOIS: [nl][N] ; PRODUCT: a1
  const u64 *src2; // PRODUCT: b1
  u64 *dst;        // [nl][N]
  // GALOIS only: also emit the permuted c0 (so the mod-down epilogue adds it with coalesced loads)
  const u64 *c0;
  u64 *pc0;
  int nl, prime0, pstep;
  u32 elt;
  size_t sstride; // words between consecutive source limbs (0 = N)
};

template <int LOGA, int LD> HD void body_intt_B(const ArgsInttB &a, int job, LaneB8 *st, u64 *sm) {
  const NttTables &T = *a.T;
  const int N = 1 << T.logN;
  const int limb = job >> LOGA, r = job & (Geo<LOGA>::ROWS - 1);
  const int p = a.prime0 + limb * a.pstep;
  const ModQ m = T.mod[p];
  const u64 *src = a.src + (size_t)limb * (a.sstride ? a.sstride : (size_t)N);
  Tw *tw = warp_tw(sm);
  LANE_DECL;
  FOR_LANES(S, st, {
    grid_dep_launch();
    stage_tw_B<LOGA>(tw, T.itwB + (size_t)p * twB_prime_stride(N), r, lane);
    grid_dep_wait();
    const int base = r * 256 + lane * 8;
    if (LD == LD_PLAIN) {
      load8_stream(src + base, S.x);
    } else if (LD == LD_GALOIS) {
      const u64 *c0 = a.c0 + (size_t)limb * N;
      u64 v[8];
      _Pragma("unroll")
      for (int e = 0; e < 8; e++) {
        u32 si = galois_src_index((u32)(base + e), a.elt, T.logN);
        S.x[e] = ldg_stream(src + si);
        v[e] = ldg_stream(c0 + si);
      }
      store8(a.pc0 + (size_t)limb * N + base, v);
    } else {
      u64 u[8], v[8];
      load8_stream(src + base, u);
      load8_stream(a.src2 + (size_t)limb * N + base, v);
      _Pragma("unroll")
      for (int e = 0; e < 8; e++) S.x[e] = mulmod(u[e], v[e], m);
      if (LD == LD_DECRYPT) { // Decryptor: c0 + c1 * s (SEAL_HEVM.cpp:330, 449)
        load8_stream(a.c0 + (size_t)limb * N + base, u);
        _Pragma("unroll")
        for (int e = 0; e < 8; e++) S.x[e] = csub(S.x[e] + u[e], m.q);
      }
    }
    cp_async_wait();
  });
  warp_invB8_regs(st, sm, tw, m);
  u64 *dst = a.dst + (size_t)limb * N + r * 256;
  FOR_LANES(S, st, {
    _Pragma("unroll")
    for (int e = 0; e < 8; e++) dst[idxH(lane, e)] = S.x[e];
  });
}

// ---------------------------------------------------------------------------------------------
// inverse pass A (+ N^-1, canonical).  ROUND: also add floor(q/2) mod q (rescale / mod-down rounding).
// ---------------------------------------------------------------------------------------------
struct ArgsInttA {
  const NttTables *T;
  const u64 *src; // [nl][N] (output of inverse pass B)
  u64 *dst;       // [nl][N] coefficient form, canonical
  int nl, prime0, pstep;
  int round; // add half
};
template <int LOGA> HD void body_intt_A(const ArgsInttA &a, int job, LaneA *st, u64 *sm) {
  const NttTables &T = *a.T;
  const int N = 1 << T.logN;
  const int limb = job >> Geo<LOGA>::LOGT, tile = job & (Geo<LOGA>::TILES - 1);
  const int p = a.prime0 + limb * a.pstep;
  const ModQ m = T.mod[p];
  const u64 q = m.q;
  Tw *tw = warp_tw(sm);
  LANE_DECL;
  FOR_LANES(S, st, {
    (void)S;
    grid_dep_launch();
    stage_tw_A<LOGA>(tw, T.itw + (size_t)p * N, lane);
    grid_dep_wait();
    cp_async_wait();
  });
  warp_invA_to_regs<LOGA>(st, sm, a.src + (size_t)limb * N, tile * Geo<LOGA>::C, tw, m, T.invn[p], T.invn_w[p]);
  u64 *dst = a.dst + (size_t)limb * N + tile * Geo<LOGA>::C;
  const u64 half = q >> 1;
  FOR_LANES(S, st, {
    u64 *dstR = dst + ((size_t)rowR<LOGA>(lane, 0) << 8) + colA<LOGA>(lane);
    _Pragma("unroll")
    for (int e = 0; e < 16; e++) {
      u64 v = S.x[e];
      if (a.round) v = csub(v + half, q);
      dstR[(size_t)e * (Geo<LOGA>::ROWS / 16) << 8] = v; // row rowR(lane, e)
    }
  });
}

// ---------------------------------------------------------------------------------------------
// forward pass A with optional pre-processing (output values are lazy, < 8q).
//   PRE_NONE : dst limb d <- src limb d, prime prime0 + d*pstep                       (nd limbs)
//   PRE_MODUP: d = Iidx*l + J : dst s2[Iidx][J] <- (t[J] mod q_I), I = Iidx<l ? Iidx : sp; skip I==J
//   PRE_ROUND: d = K*nlim + i : dst s4[K][i]  <- (r[K] mod q_i) + q_i - (half mod q_i), half = q_plast/2
// ---------------------------------------------------------------------------------------------
struct ArgsFwdA {
  const NttTables *T;
  const u64 *src;
  u64 *dst;
  int nd, prime0, pstep;
  int l;     // MODUP: level ; ROUND: nlim (number of target limbs)
  int sp;    // MODUP: special prime index
  int plast; // ROUND: prime index of the divided-out modulus
  // limb-sharded key switching: only the targets [t0, t0 + nt) are produced (nt = 0: all of them)
  int t0, nt;
  int j0, nj; // MODUP: only the digits [j0, j0 + nj) (nj = 0: all l) -- the sharded key switch starts on a peer's digits as they arrive
  int pmod; // PRE_NONE: limb d uses prime prime0 + (d % pmod) * pstep (several polynomials in one launch); 0 = no wrap
};
template <int LOGA, int PRE> HD void body_fwd_A(const ArgsFwdA &a, int job, LaneA *st, u64 *sm) {
  const NttTables &T = *a.T;
  const int N = 1 << T.logN;
  const int d = job >> Geo<LOGA>::LOGT, tile = job & (Geo<LOGA>::TILES - 1);
  int ps, pd, sl; // source prime, destination prime, source limb
  if (PRE == PRE_NONE) {
    sl = d;
    ps = pd = a.prime0 + (a.pmod ? d % a.pmod : d) * a.pstep;
  } else if (PRE == PRE_MODUP) {
    const int nj = a.nj ? a.nj : a.l;
    const int Iidx = d / nj + a.t0; // job rows enumerate the owned targets only
    sl = a.j0 + d - (Iidx - a.t0) * nj;
    ps = sl;
    pd = (Iidx == a.l) ? a.sp : Iidx;
    if (pd == ps) return; // diagonal: the NTT-form input is used directly by the MAC kernel
  } else {
    const int nt = a.nt ? a.nt : a.l;
    const int K = d / nt;
    sl = K;
    ps = a.plast;
    pd = a.t0 + d - K * nt;
  }
  // row of the destination matrix: s2[Iidx][J] (MODUP), s4[K][i] (ROUND), limb d (NONE)
  const int drow = (PRE == PRE_MODUP) ? ((pd == a.sp ? a.l : pd) * a.l + sl) : (PRE == PRE_ROUND) ? (sl * a.l + pd) : d;
  const ModQ m = T.mod[pd];
  const u64 *src = a.src + (size_t)sl * N + tile * Geo<LOGA>::C;
  u64 fix = 0;
  if (PRE == PRE_ROUND) fix = m.q - reduce64(T.mod[ps].q >> 1, m);
  Tw *tw = warp_tw(sm);
  LANE_DECL;
  FOR_LANES(S, st, {
    grid_dep_launch();
    stage_tw_A<LOGA>(tw, T.tw + (size_t)pd * N, lane);
    grid_dep_wait();
    _Pragma("unroll")
    for (int e = 0; e < 16; e++) S.y[e] = ldg_stream(src + ((size_t)rowR<LOGA>(lane, 0) << 8) + colA<LOGA>(lane) + ((size_t)e * (Geo<LOGA>::ROWS / 16) << 8));
    // No modular reduction: every prime is 2^60 - delta with delta < 2^32, so a canonical residue of
    // ANY prime is < 2^60 < 2*q_dst, i.e. already a valid lazy representative mod q_dst.
    _Pragma("unroll")
    for (int e = 0; e < 16; e++) S.y[e] += fix;
    cp_async_wait();
  });
  (void)ps;
  warp_fwdA_from_regs<LOGA, pass_a_needs_fold<LOGA>(PRE)>(st, sm, a.dst + (size_t)drow * N, tile * Geo<LOGA>::C, tw, m);
}


// ---------------------------------------------------------------------------------------------
// fused: inverse pass A (+N^-1, canonical) of one source limb tile, then for a group of targets
// [mod-up reduction | rounding + fix-up] + forward pass A, all on the same 128x4 tile held in
// registers.  job = ((source limb * ngroups) + group) * 64 + tile.
//   PRE_MODUP: source limb J (prime J) of s1[l][N]; targets Iidx in [0,l] except the diagonal;
//              dst s2[Iidx][J]
//   PRE_ROUND: source limb K (prime plast) of s1[2][N]; x <- (x + q_last/2) mod q_last;
//              targets i in [0,nlim): dst s4[K][i] <- (x mod q_i) + q_i - (q_last/2 mod q_i)
// Forward twiddle tables are double-buffered in the warp's staging area (cp.async groups).
// ---------------------------------------------------------------------------------------------
struct ArgsInvFwdA {
  const NttTables *T;
  const u64 *src;
  u64 *dst;
  int nsrc, l, sp, plast, ngroups;
};
template <int LOGA, int PRE> HD void body_invA_fwdA(const ArgsInvFwdA &a, int job, LaneA *st, u64 *sm) {
  const NttTables &T = *a.T;
  const int N = 1 << T.logN;
  const int tile = job & (Geo<LOGA>::TILES - 1), rest = job >> Geo<LOGA>::LOGT;
  const int sl = rest / a.ngroups, grp = rest - sl * a.ngroups;
  const int ntargets = (PRE == PRE_MODUP) ? a.l + 1 : a.l;
  const int tpj = (ntargets + a.ngroups - 1) / a.ngroups;
  const int tend = (grp + 1) * tpj < ntargets ? (grp + 1) * tpj : ntargets;
  const int ps = (PRE == PRE_MODUP) ? sl : a.plast;
  const ModQ ms = T.mod[ps];
  // target slot -> prime (MODUP: slot l is the special prime); -1 when past the end
  auto prime_of = [&](int t) { return (PRE == PRE_MODUP && t == a.l) ? a.sp : t; };
  auto next_valid = [&](int t) {
    while (t < tend && prime_of(t) == ps) t++; // the diagonal is taken from the NTT-form input by the MAC kernel
    return t;
  };
  int t = next_valid(grp * tpj);
  if (t >= tend) return;
  Tw *tw_inv = warp_tw(sm), *tw_fwd = tw_inv + Geo<LOGA>::ROWS;
  int buf = 0;
  LANE_DECL;
  FOR_LANES(S, st, {
    (void)S;
    grid_dep_launch();
    stage_tw_A<LOGA>(tw_inv, T.itw + (size_t)ps * N, lane);
    stage_tw_A<LOGA>(tw_fwd, T.tw + (size_t)prime_of(t) * N, lane);
    grid_dep_wait();
    cp_async_wait();
  });
  warp_invA_to_regs<LOGA>(st, sm, a.src + (size_t)sl * N, tile * Geo<LOGA>::C, tw_inv, ms, T.invn[ps], T.invn_w[ps]);
  if (PRE == PRE_ROUND) {
    const u64 half = ms.q >> 1;
    FOR_LANES(S, st, {
      _Pragma("unroll")
      for (int e = 0; e < 16; e++) S.x[e] = csub(S.x[e] + half, ms.q);
    });
  }
  while (t < tend) {
    const int pd = prime_of(t);
    const int tn = next_valid(t + 1);
    const ModQ m = T.mod[pd];
    u64 fix = 0;
    if (PRE == PRE_ROUND) fix = m.q - reduce64(ms.q >> 1, m);
    FOR_LANES(S, st, {
      if (Geo<LOGA>::TW_DOUBLE && tn < tend)
        stage_tw_A<LOGA>(tw_fwd + Geo<LOGA>::ROWS * (buf ^ 1), T.tw + (size_t)prime_of(tn) * N, lane); // prefetch next table
      cp_async_commit();
      // no modular reduction needed (see body_fwd_A): x < 2^60 < 2*q_dst is a valid lazy representative
      _Pragma("unroll")
      for (int e = 0; e < 16; e++) S.y[e] = S.x[e] + fix;
      cp_async_wait_keep1(); // the current table (older group) has landed
    });
    u64 *dst = (PRE == PRE_MODUP) ? a.dst + ((size_t)t * a.l + sl) * N : a.dst + ((size_t)sl * a.l + t) * N;
    warp_fwdA_from_regs<LOGA, pass_a_needs_fold<LOGA>(PRE)>(st, sm, dst, tile * Geo<LOGA>::C, tw_fwd + Geo<LOGA>::ROWS * buf, m);
    if (Geo<LOGA>::TW_DOUBLE) {
      buf ^= 1;
    } else if (tn < tend) { // single buffer: the next table can only be staged once this transform is done with it
      FOR_LANES(S, st, {
        (void)S;
        stage_tw_A<LOGA>(tw_fwd, T.tw + (size_t)prime_of(tn) * N, lane);
        cp_async_commit();
      });
    }
    t = tn;
  }
  FOR_LANES(S, st, {
    (void)S;
    cp_async_wait();
  });
}

// ---------------------------------------------------------------------------------------------
// forward pass B + epilogue
// ---------------------------------------------------------------------------------------------
struct ArgsFwdB {
  const NttTables *T;
  const u64 *src; // CANON: [nd][N]; MAC: s2 [l+1][l][N]; MODDOWN/RESCALE: s4 [2][l][N]
  u64 *dst;       // CANON: [nd][N]; MAC: acc [2][l+1][N]; MODDOWN/RESCALE: ct out (poly pitch = pitch)
  int nd, prime0, pstep;
  int l, sp;
  // MAC
  const u64 *key; // [L-1][2][Ltot][N]; limb-sharded storage: only the limbs [key_t0, key_t0 + Ltot) of every key polynomial
  int Ltot;       // limbs per stored key polynomial (L for a whole key); the special limb is always the LAST stored limb
  int key_t0;     // prime index of the first stored limb (0 for a whole key)
  int ld;            // how the NTT-form target (diagonal term) is obtained: LD_*
  const u64 *tgt;    // PLAIN: target [l][N]; GALOIS: c1 (gather with elt); PRODUCT: a1
  const u64 *tgt2;   // PRODUCT: b1
  u32 elt;
  // MODDOWN / RESCALE
  const u64 *acc;    // MODDOWN: acc [2][l+1][N]
  const u64 *add0;   // GALOIS: pc0 [l][N];  RELIN: a (ct, poly pitch) ; RESCALE: input ct
  const u64 *add1;   // RELIN: b (ct)
  size_t pitch;      // poly pitch (words) of ct operands / output
  size_t spitch;     // RESCALE: poly pitch of the INPUT ciphertext when it differs from `pitch` (0 = same)
  int plast;         // prime divided out (sp for MODDOWN, l_in-1 for RESCALE)
  u64 *sp_rows;      // MAC: [2][N] inverse pass-B output of the special-prime accumulators (mod-down input)
  // limb-sharded key switching: MAC jobs cover the targets Iidx = i_end-1, i_end-2, ... (i_end = 0: l+1, i.e. all);
  // MODDOWN_GALOIS jobs cover the limbs [t0, t0 + nt) (nt = 0: all l)
  int i_end, t0, nt;
  int pmod; // CANON: limb d uses prime prime0 + (d % pmod) * pstep; 0 = no wrap
};

// ---- key-switch inner product: one CTA of MAC_WARPS warps per (Iidx, row) ----------------------
// phase 0 (all threads of the CTA): stage the row's twiddles of prime I
// job -> (Iidx, row): the special prime (Iidx = l) comes first because its CTAs carry the extra tail
template <int LOGA> HD int mac_Iidx(const ArgsFwdB &a, int job) { return (a.i_end ? a.i_end - 1 : a.l) - (job >> LOGA); }
template <int LOGA> HD void body_mac_stage(const ArgsFwdB &a, int job, int tid, Tw *tw_s) {
  const NttTables &T = *a.T;
  const int N = 1 << T.logN;
  const int r = job & (Geo<LOGA>::ROWS - 1), Iidx = mac_Iidx<LOGA>(a, job), I = (Iidx == a.l) ? a.sp : Iidx;
  grid_dep_launch();
  stage_tw_B<LOGA>(tw_s, T.twB + (size_t)I * twB_prime_stride(N), r, tid, MAC_WARPS * 32);
  grid_dep_wait();
  cp_async_wait();
}
// phase 1 (per warp): digits p = w, w + MAC_WARPS, ... of a rotated digit order (data primes: J = (I + p) mod l, so the
// cheap diagonal term always falls to warp 0, the warp with one digit more).  Forward pass B of digit J, then the
// result (folded below 2^60 + 2^36, radix-2^30 split) goes to xbuf[J][coefficient] in shared memory.
// The pass-A row of the warp's NEXT digit is staged into `rowbuf` with cp.async while the current one is transformed.
template <int LOGA> HD void body_mac_warp(const ArgsFwdB &a, int job, int w, LaneB8 *st, u64 *tile, const Tw *tw_s, u64 *rowbuf, u64 *xbuf) {
  const NttTables &T = *a.T;
  const int N = 1 << T.logN;
  const int r = job & (Geo<LOGA>::ROWS - 1), Iidx = mac_Iidx<LOGA>(a, job), I = (Iidx == a.l) ? a.sp : Iidx;
  const ModQ m = T.mod[I];
  LANE_DECL;
  auto digit_of = [&](int p) { // p-th digit of the rotated order
    if (Iidx == a.l) return p;
    const int J = p + I;
    return J >= a.l ? J - a.l : J;
  };
  auto stage = [&](int J) {
    if (J == I) return;
    const u64 *src = a.src + ((size_t)Iidx * a.l + J) * N + r * 256;
    FOR_LANES(S, st, {
      (void)S;
      _Pragma("unroll")
      for (int i = 0; i < 4; i++) cp_async16_cg(rowbuf + 2 * (i * 32 + lane), src + 2 * (i * 32 + lane));
    });
  };
  int p = w;
  if (p < a.l) stage(digit_of(p));
  for (; p < a.l; p += MAC_WARPS) {
    const int J = digit_of(p);
    if (J == I) {
      // diagonal: NTT-form target limb J, layout C, canonical
      FOR_LANES(S, st, {
        const int base = r * 256 + lane * 8;
        if (a.ld == LD_PLAIN) {
          load8_stream(a.tgt + (size_t)J * N + base, S.x);
        } else if (a.ld == LD_GALOIS) {
          const u64 *t = a.tgt + (size_t)J * N;
          _Pragma("unroll")
          for (int e = 0; e < 8; e++) S.x[e] = ldg_stream(t + galois_src_index((u32)(base + e), a.elt, T.logN));
        } else {
          u64 u[8], v[8];
          load8_stream(a.tgt + (size_t)J * N + base, u);
          load8_stream(a.tgt2 + (size_t)J * N + base, v);
          _Pragma("unroll")
          for (int e = 0; e < 8; e++) S.x[e] = mulmod(u[e], v[e], m);
        }
      });
    } else {
      FOR_LANES(S, st, {
        cp_async_wait();
        (void)S;
      });
      FOR_LANES(S, st, {
        _Pragma("unroll")
        for (int e = 0; e < 8; e++) S.x[e] = rowbuf[idxH(lane, e)];
      });
    }
    if (p + MAC_WARPS < a.l) stage(digit_of(p + MAC_WARPS)); // every lane has drained rowbuf
    if (J != I) warp_fwdB8_regs(st, tile, tw_s, m);
    u64 *xo = xbuf + (size_t)J * MAC_XROW;
    FOR_LANES(S, st, {
      u64 *xl = xo + xpad(lane * 8); // 8 consecutive words (four 16-byte stores): the pad never splits them
      _Pragma("unroll")
      for (int e = 0; e < 8; e++) xl[e] = split30(fold60(S.x[e], m.delta));
    });
  }
}
// phase 2 (all threads, after a CTA barrier): thread t owns coefficients 2t, 2t+1 of the row for both key polys and
// runs over all digits: x from shared memory, the two key rows straight from HBM (one 16-byte load each, issued
// MAC_KEY_BATCH digits ahead), four 32x32->64 multiply-adds per term.  Then Barrett and a 16-byte store.
// Data primes: store acc[K][Iidx][row].  Special prime: keep the two rows in shared memory (`rows`,
// [2][256]) for phase 3 instead -- nobody else reads them.
#ifndef MAC_KEY_BATCH
#define MAC_KEY_BATCH 4
#endif
#ifndef MAC_KEY_EVICT_FIRST
#define MAC_KEY_EVICT_FIRST 1
#endif
struct U2 {
  u64 a, b;
};
HD U2 ldg_key2(const u64 *p) { // 16 bytes of key material (read-only for the life of the VM)
#if defined(__CUDA_ARCH__)
#if MAC_KEY_EVICT_FIRST
  // every key byte is used exactly once per key switch: mark its L2 lines evict-first so that streaming 95 MB of key
  // does not push the mod-up matrix s2 (written by the previous kernel, read by this one) out of the 126 MB L2
  u64 pol, a, b;
  asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.u64 {%0,%1}, [%2], %3;" : "=l"(a), "=l"(b) : "l"(p), "l"(pol));
  return U2{a, b};
#else
  const ulonglong2 v = __ldg(reinterpret_cast<const ulonglong2 *>(p));
  return U2{v.x, v.y};
#endif
#else
  return U2{p[0], p[1]};
#endif
}
HD void mac30(u64 (&c)[4], u64 xw, u64 kw) {
  const u32 x0 = (u32)xw, x1 = (u32)(xw >> 32), k0 = (u32)kw, k1 = (u32)(kw >> 32);
#if defined(__CUDA_ARCH__)
  // one IMAD.WIDE.U32 each (the compiler's own lowering of the C form below splits the 64-bit accumulation)
  asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(c[0]) : "r"(x0), "r"(k0));
  asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(c[1]) : "r"(x0), "r"(k1));
  asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(c[2]) : "r"(x1), "r"(k0));
  asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(c[3]) : "r"(x1), "r"(k1));
#else
  c[0] += (u64)x0 * k0;
  c[1] += (u64)x0 * k1;
  c[2] += (u64)x1 * k0;
  c[3] += (u64)x1 * k1;
#endif
}
// (lo,hi) += c0 + (c1 + c2) * 2^30 + c3 * 2^60 ; columns cleared
HD void flush30(u64 &lo, u64 &hi, u64 (&c)[4]) {
  const u64 mid_lo = c[1] + c[2];
  const u64 mid_hi = mid_lo < c[1] ? 1 : 0; // 65-bit sum
  u64 add_lo = c[0], add_hi = 0;
  u64 t = mid_lo << 30;
  add_lo += t, add_hi += (add_lo < t ? 1 : 0) + (mid_lo >> 34) + (mid_hi << 30);
  t = c[3] << 60;
  add_lo += t, add_hi += (add_lo < t ? 1 : 0) + (c[3] >> 4);
  lo += add_lo;
  hi += add_hi + (lo < add_lo ? 1 : 0);
  c[0] = c[1] = c[2] = c[3] = 0;
}
template <int LOGA> HD void body_mac_dot(const ArgsFwdB &a, int job, int tid, const u64 *xbuf, u64 *rows) {
  const NttTables &T = *a.T;
  const int N = 1 << T.logN;
  const int r = job & (Geo<LOGA>::ROWS - 1), Iidx = mac_Iidx<LOGA>(a, job), I = (Iidx == a.l) ? a.sp : Iidx;
  const ModQ m = T.mod[I];
  const size_t kstride = (size_t)a.Ltot * N; // words between key[J][0] and key[J][1]
  const int Iloc = (Iidx == a.l) ? a.Ltot - 1 : Iidx - a.key_t0; // stored limb index of target I
  const u64 *kp = a.key + (size_t)Iloc * N + r * 256 + 2 * tid;
  const size_t dstride = 2 * kstride; // words between consecutive digits of the key
  const u64 pol = 0;
  (void)pol;
  // digits [Jb, Je), Je - Jb <= MAC_FLUSH_DIGITS: a straight multiply-add loop over batches of MAC_KEY_BATCH digits
  // (key rows of a batch are loaded together; digits past Je contribute a zero key), then one recombination
  auto chunk = [&](int Jb, int Je, u64 (&lo)[2][2], u64 (&hi)[2][2]) {
    u64 c[2][2][4]; // [key poly][coefficient][column]
    _Pragma("unroll")
    for (int K = 0; K < 2; K++)
      _Pragma("unroll")
      for (int j = 0; j < 2; j++) c[K][j][0] = c[K][j][1] = c[K][j][2] = c[K][j][3] = 0;
    // running pointers: key rows of digit J (kq: poly 0, kq + kstride: poly 1) and its x row
    const u64 *kq = kp + (size_t)Jb * dstride;
    const u64 *xq = xbuf + (size_t)Jb * MAC_XROW + xpad(2 * tid);
    int J = Jb;
    for (; J + MAC_KEY_BATCH <= Je; J += MAC_KEY_BATCH) { // full batches: the key rows of MAC_KEY_BATCH digits in flight
      U2 k[MAC_KEY_BATCH][2];
      _Pragma("unroll")
      for (int b = 0; b < MAC_KEY_BATCH; b++) {
        k[b][0] = ldg_key2(kq);
        k[b][1] = ldg_key2(kq + kstride);
        kq += dstride;
      }
      _Pragma("unroll")
      for (int b = 0; b < MAC_KEY_BATCH; b++) {
        const u64 xa = xq[b * MAC_XROW], xb = xq[b * MAC_XROW + 1];
        mac30(c[0][0], xa, k[b][0].a);
        mac30(c[0][1], xb, k[b][0].b);
        mac30(c[1][0], xa, k[b][1].a);
        mac30(c[1][1], xb, k[b][1].b);
      }
      xq += MAC_KEY_BATCH * MAC_XROW;
    }
    for (; J < Je; J++) { // remaining digits one by one
      const U2 k0 = ldg_key2(kq), k1 = ldg_key2(kq + kstride);
      const u64 xa = xq[0], xb = xq[1];
      mac30(c[0][0], xa, k0.a);
      mac30(c[0][1], xb, k0.b);
      mac30(c[1][0], xa, k1.a);
      mac30(c[1][1], xb, k1.b);
      kq += dstride;
      xq += MAC_XROW;
    }
    _Pragma("unroll")
    for (int K = 0; K < 2; K++)
      _Pragma("unroll")
      for (int j = 0; j < 2; j++) flush30(lo[K][j], hi[K][j], c[K][j]);
  };
  u64 lo[2][2] = {{0, 0}, {0, 0}}, hi[2][2] = {{0, 0}, {0, 0}};
  for (int Jc = 0; Jc < a.l; Jc += MAC_FLUSH_DIGITS) chunk(Jc, a.l < Jc + MAC_FLUSH_DIGITS ? a.l : Jc + MAC_FLUSH_DIGITS, lo, hi);
  _Pragma("unroll")
  for (int K = 0; K < 2; K++) {
    u64 v[2];
    _Pragma("unroll")
    for (int j = 0; j < 2; j++) v[j] = reduce128(lo[K][j], hi[K][j], m);
    u64 *o = (Iidx == a.l) ? rows + K * 256 + 2 * tid : a.dst + ((size_t)K * (a.l + 1) + Iidx) * N + r * 256 + 2 * tid;
    o[0] = v[0], o[1] = v[1];
  }
}
// phase 3 (special-prime CTAs only, warps K = 0,1, after a CTA barrier): inverse pass B of row r of the
// accumulator acc[K][special] (first step of the mod-down), straight from shared memory.
template <int LOGA> HD void body_mac_tail(const ArgsFwdB &a, int job, int K, LaneB8 *st, u64 *tile, Tw *tw_s, const u64 *rows) {
  const NttTables &T = *a.T;
  const int N = 1 << T.logN;
  const int r = job & (Geo<LOGA>::ROWS - 1);
  const ModQ m = T.mod[a.sp];
  LANE_DECL;
  FOR_LANES(S, st, {
    stage_tw_B<LOGA>(tw_s, T.itwB + (size_t)a.sp * twB_prime_stride(N), r, lane); // `tw_s` is private to this warp (see the callers)
    _Pragma("unroll")
    for (int e = 0; e < 8; e++) S.x[e] = rows[K * 256 + lane * 8 + e];
    cp_async_wait();
  });
  warp_invB8_regs(st, tile, tw_s, m);
  u64 *dst = a.sp_rows + (size_t)K * N + r * 256;
  FOR_LANES(S, st, {
    _Pragma("unroll")
    for (int e = 0; e < 8; e++) dst[idxH(lane, e)] = S.x[e];
  });
}

template <int LOGA, int EPI> HD void body_fwd_B(const ArgsFwdB &a, int job, LaneB8 *st, u64 *sm) {
  const NttTables &T = *a.T;
  const int N = 1 << T.logN;
  const int r = job & (Geo<LOGA>::ROWS - 1), d = job >> LOGA;
  Tw *tw = warp_tw(sm);
  LANE_DECL;
  if (EPI == EPI_CANON) {
    const int p = a.prime0 + (a.pmod ? d % a.pmod : d) * a.pstep;
    const ModQ m = T.mod[p];
    const u64 *src = a.src + (size_t)d * N + r * 256;
    FOR_LANES(S, st, {
      grid_dep_launch();
      stage_tw_B<LOGA>(tw, T.twB + (size_t)p * twB_prime_stride(N), r, lane);
      grid_dep_wait();
      _Pragma("unroll")
      for (int e = 0; e < 8; e++) S.x[e] = ldg_stream(src + idxH(lane, e));
      cp_async_wait();
    });
    warp_fwdB8_regs(st, sm, tw, m);
    u64 *o = a.dst + (size_t)d * N + r * 256;
    FOR_LANES(S, st, {
      u64 v[8];
      _Pragma("unroll")
      for (int e = 0; e < 8; e++) v[e] = canon60(S.x[e], m.q, m.delta);
      store8(o + lane * 8, v);
    });
    return;
  }
  if (EPI == EPI_MODDOWN_RELIN) {
    // d = limb i; both output polys are produced by the same warp so that every input
    // (a0,a1,b0,b1 at this position) is read before either output is written (dst may alias a or b).
    const int i = d + a.t0; // limb-sharded: jobs enumerate the owned limbs [t0, t0 + nt)
    const ModQ m = T.mod[i];
    const u64 q = m.q;
    const Tw inv = T.qinv[a.plast][i];
    const size_t off = (size_t)i * N + r * 256;
    for (int K = 0; K < 2; K++) {
      const u64 *src = a.src + ((size_t)K * a.l + i) * N + r * 256;
      FOR_LANES(S, st, {
        if (K == 0) {
          grid_dep_launch();
          stage_tw_B<LOGA>(tw, T.twB + (size_t)i * twB_prime_stride(N), r, lane);
          grid_dep_wait();
        }
        if (K == 1) {
          _Pragma("unroll")
          for (int e = 0; e < 8; e++) S.z[e] = S.x[e];
        }
        _Pragma("unroll")
        for (int e = 0; e < 8; e++) S.x[e] = ldg_stream(src + idxH(lane, e));
        cp_async_wait();
      });
      warp_fwdB8_regs(st, sm, tw, m);
    }
    FOR_LANES(S, st, {
      const int b = lane * 8;
      u64 c0[8], c1[8], a0[8], a1[8], b0[8], b1[8], v0[8], v1[8];
      load8_stream(a.acc + ((size_t)0 * (a.l + 1) + i) * N + r * 256 + b, c0);
      load8_stream(a.acc + ((size_t)1 * (a.l + 1) + i) * N + r * 256 + b, c1);
      load8_stream(a.add0 + off + b, a0);
      load8_stream(a.add0 + a.pitch + off + b, a1);
      load8_stream(a.add1 + off + b, b0);
      load8_stream(a.add1 + a.pitch + off + b, b1);
      _Pragma("unroll")
      for (int e = 0; e < 8; e++) {
        u64 u0 = canon60(S.z[e], q, m.delta), u1 = canon60(S.x[e], q, m.delta);
        u64 t0 = shoup_mul(c0[e] + q - u0, inv, q);
        u64 t1 = shoup_mul(c1[e] + q - u1, inv, q);
        v0[e] = csub(t0 + mulmod(a0[e], b0[e], m), q);
        // d1 = a0*b1 + a1*b0 : one 128-bit sum, one Barrett (same canonical value as two mulmods + add)
        u64 lo = 0, hi = 0;
        mac128(lo, hi, a0[e], b1[e]);
        mac128(lo, hi, a1[e], b0[e]);
        v1[e] = csub(t1 + reduce128(lo, hi, m), q);
      }
      store8(a.dst + off + b, v0);
      store8(a.dst + a.pitch + off + b, v1);
    });
    return;
  }
  // MODDOWN_GALOIS / RESCALE: d = K*nt + (i - t0)
  {
    const int nt = a.nt ? a.nt : a.l;
    const int K = d / nt, i = a.t0 + d - K * nt;
    const ModQ m = T.mod[i];
    const u64 q = m.q;
    const u64 *src = a.src + ((size_t)K * a.l + i) * N + r * 256;
    FOR_LANES(S, st, {
      grid_dep_launch();
      stage_tw_B<LOGA>(tw, T.twB + (size_t)i * twB_prime_stride(N), r, lane);
      grid_dep_wait();
      _Pragma("unroll")
      for (int e = 0; e < 8; e++) S.x[e] = ldg_stream(src + idxH(lane, e));
      cp_async_wait();
    });
    warp_fwdB8_regs(st, sm, tw, m);
    const Tw inv = T.qinv[a.plast][i];
    const size_t off = (size_t)i * N + r * 256;
    u64 *o = a.dst + (size_t)K * a.pitch + off;
    FOR_LANES(S, st, {
      const int b = lane * 8;
      u64 c[8], v[8];
      const u64 *cin = (EPI == EPI_RESCALE) ? a.add0 + (size_t)K * (a.spitch ? a.spitch : a.pitch) + off + b
                                            : a.acc + ((size_t)K * (a.l + 1) + i) * N + r * 256 + b;
      load8_stream(cin, c);
      _Pragma("unroll")
      for (int e = 0; e < 8; e++) {
        u64 u = canon60(S.x[e], q, m.delta);    // canonical NTT of the rounding term
        v[e] = shoup_mul(c[e] + q - u, inv, q); // (c - u) * plast^-1 mod q
      }
      if (EPI == EPI_MODDOWN_GALOIS && K == 0) {
        u64 w[8];
        load8_stream(a.add0 + off + b, w);
        _Pragma("unroll")
        for (int e = 0; e < 8; e++) v[e] = csub(v[e] + w[e], q);
      }
      if (EPI == EPI_RESCALE && K == 0 && a.add1) { // public-key encryption: c0 += plaintext (Encryptor::encrypt, last step)
        u64 w[8];
        load8_stream(a.add1 + off + b, w);
        _Pragma("unroll")
        for (int e = 0; e < 8; e++) v[e] = csub(v[e] + w[e], q);
      }
      store8(o + b, v);
    });
  }
}
