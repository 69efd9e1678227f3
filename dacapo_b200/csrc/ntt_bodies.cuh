// Warp-job bodies of the NTT-based kernels (N = 2^15: 128 rows x 256 cols, LOGB = 8).
// One warp executes one job; a job is a (limb, tile) or (limb, row) pair.  See ntt_core.cuh
// for the schedule, and kernels.cu for the __global__ wrappers and the launch geometry.
//
// Fusions (reference call sites SEAL_HEVM.cpp:273,283,315-316):
//   body_intt_B  : [Galois gather | ct x ct product d2 = a1*b1 | plain load] + inverse pass B
//   body_intt_A  : inverse pass A + N^-1 [+ "add q_last/2" rounding of rescale / mod-down]
//   body_fwd_A   : [mod-up reduction "t_J mod q_I" | rounding fix-up] + forward pass A
//   body_fwd_B   : forward pass B + one of
//        CANON    canonical store (plain NTT)
//        MAC      key-switch inner product over all digits J (128-bit lazy accumulators in
//                 registers, Barrett at the end) -- the l x (l+1) digit matrix never exists in HBM
//        MODDOWN  (acc - u) * p^-1 + addend   with addend = permuted c0 | tensor product d0/d1
//        RESCALE  (c - u) * q_last^-1
#pragma once
#include "ntt_core.cuh"

enum { LD_PLAIN = 0, LD_GALOIS = 1, LD_PRODUCT = 2 };
enum { PRE_NONE = 0, PRE_MODUP = 1, PRE_ROUND = 2 };
enum { EPI_CANON = 0, EPI_MAC = 1, EPI_MODDOWN_GALOIS = 2, EPI_MODDOWN_RELIN = 3, EPI_RESCALE = 4 };

#define LOGB8 8
#define ROWS 128
#define TILES_A 64 // 256 cols / 4

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
};

template <int LD> HD void body_intt_B(const ArgsInttB &a, int job, LaneB8 *st, u64 *sm) {
  const NttTables &T = *a.T;
  const int N = 1 << T.logN;
  const int limb = job >> 7, r = job & 127;
  const int p = a.prime0 + limb * a.pstep;
  const ModQ m = T.mod[p];
  const u64 *src = a.src + (size_t)limb * N;
  LANE_DECL;
  FOR_LANES(S, st, {
    const int base = r * 256 + lane * 8;
    if (LD == LD_PLAIN) {
      ldg_stream4(src + base, S.x[0], S.x[1], S.x[2], S.x[3]);
      ldg_stream4(src + base + 4, S.x[4], S.x[5], S.x[6], S.x[7]);
    } else if (LD == LD_GALOIS) {
      const u64 *c0 = a.c0 + (size_t)limb * N;
      u64 *pc0 = a.pc0 + (size_t)limb * N;
      u64 v[8];
_Pragma("unroll")
      for (int e = 0; e < 8; e++) {
        u32 si = galois_src_index((u32)(base + e), a.elt, T.logN);
        S.x[e] = ldg_stream(src + si);
        v[e] = ldg_stream(c0 + si);
      }
      stg4(pc0 + base, v[0], v[1], v[2], v[3]);
      stg4(pc0 + base + 4, v[4], v[5], v[6], v[7]);
    } else {
      const u64 *b = a.src2 + (size_t)limb * N;
      u64 u[8], v[8];
      ldg_stream4(src + base, u[0], u[1], u[2], u[3]);
      ldg_stream4(src + base + 4, u[4], u[5], u[6], u[7]);
      ldg_stream4(b + base, v[0], v[1], v[2], v[3]);
      ldg_stream4(b + base + 4, v[4], v[5], v[6], v[7]);
_Pragma("unroll")
      for (int e = 0; e < 8; e++) S.x[e] = mulmod(u[e], v[e], m);
    }
  });
  warp_invB8_regs(st, sm, r, T.itw + (size_t)p * N, m.q);
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
HD void body_intt_A(const ArgsInttA &a, int job, LaneA *st, u64 *sm) {
  const NttTables &T = *a.T;
  const int N = 1 << T.logN;
  const int limb = job >> 6, tile = job & 63;
  const int p = a.prime0 + limb * a.pstep;
  const u64 q = T.mod[p].q;
  warp_invA_to_regs<LOGB8>(st, sm, a.src + (size_t)limb * N, tile * 4, T.itw + (size_t)p * N, q, T.invn[p], T.invn_w[p]);
  u64 *dst = a.dst + (size_t)limb * N + tile * 4;
  const u64 half = q >> 1;
  LANE_DECL;
  FOR_LANES(S, st, {
_Pragma("unroll")
    for (int e = 0; e < 16; e++) {
      u64 v = S.x[e];
      if (a.round) v = csub(v + half, q);
      dst[((size_t)rowR(lane, e) << LOGB8) + (lane & 3)] = v;
    }
  });
}

// ---------------------------------------------------------------------------------------------
// forward pass A with optional pre-processing.
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
};
template <int PRE> HD void body_fwd_A(const ArgsFwdA &a, int job, LaneA *st, u64 *sm) {
  const NttTables &T = *a.T;
  const int N = 1 << T.logN;
  const int d = job >> 6, tile = job & 63;
  int ps, pd, sl; // source prime, destination prime, source limb
  if (PRE == PRE_NONE) {
    sl = d;
    ps = pd = a.prime0 + d * a.pstep;
  } else if (PRE == PRE_MODUP) {
    const int Iidx = d / a.l;
    sl = d - Iidx * a.l;
    ps = sl;
    pd = (Iidx == a.l) ? a.sp : Iidx;
    if (pd == ps) return; // diagonal: the NTT-form input is used directly by the MAC kernel
  } else {
    const int K = d / a.l;
    sl = K;
    ps = a.plast;
    pd = d - K * a.l;
  }
  const ModQ m = T.mod[pd];
  const u64 *src = a.src + (size_t)sl * N + tile * 4;
  u64 fix = 0;
  if (PRE == PRE_ROUND) fix = m.q - reduce64(T.mod[ps].q >> 1, m);
  LANE_DECL;
  FOR_LANES(S, st, {
_Pragma("unroll")
    for (int e = 0; e < 16; e++) {
      u64 v = ldg_stream(src + ((size_t)rowR(lane, e) << LOGB8) + (lane & 3));
      if (PRE != PRE_NONE && ps > pd) v = reduce64(v, m);
      S.y[e] = v + fix;
    }
  });
  warp_fwdA_from_regs<LOGB8>(st, sm, a.dst + (size_t)d * N, tile * 4, T.tw + (size_t)pd * N, m.q);
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
  const u64 *key; // [L-1][2][L][N]
  int Ltot;
  int ld;            // how the NTT-form target (diagonal term) is obtained: LD_*
  const u64 *tgt;    // PLAIN: target [l][N]; GALOIS: c1 (gather with elt); PRODUCT: a1
  const u64 *tgt2;   // PRODUCT: b1
  u32 elt;
  // MODDOWN / RESCALE
  const u64 *acc;    // MODDOWN: acc [2][l+1][N]
  const u64 *add0;   // GALOIS: pc0 [l][N];  RELIN: a (ct, poly pitch) ; RESCALE: input ct
  const u64 *add1;   // RELIN: b (ct)
  size_t pitch;      // poly pitch (words) of ct operands / output
  int plast;         // prime divided out (sp for MODDOWN, l_in-1 for RESCALE)
};

template <int EPI> HD void body_fwd_B(const ArgsFwdB &a, int job, LaneB8 *st, u64 *sm) {
  const NttTables &T = *a.T;
  const int N = 1 << T.logN;
  const int r = job & 127, d = job >> 7;
  LANE_DECL;
  if (EPI == EPI_MAC) {
    // d = Iidx in [0, l]; loop over digits J
    const int Iidx = d, I = (Iidx == a.l) ? a.sp : Iidx;
    const ModQ m = T.mod[I];
    const Tw *tw = T.tw + (size_t)I * N;
    u64 lo0[NLANE_STATE][8], hi0[NLANE_STATE][8], lo1[NLANE_STATE][8], hi1[NLANE_STATE][8];
    FOR_LANES(S, st, {
      (void)S;
      const int li = (NLANE_STATE == 1) ? 0 : lane;
_Pragma("unroll")
      for (int e = 0; e < 8; e++) lo0[li][e] = hi0[li][e] = lo1[li][e] = hi1[li][e] = 0;
    });
    for (int J = 0; J < a.l; J++) {
      if (J == I) {
        // diagonal: NTT-form target limb J, layout C
        FOR_LANES(S, st, {
          const int base = r * 256 + lane * 8;
          if (a.ld == LD_PLAIN) {
            const u64 *t = a.tgt + (size_t)J * N + base;
            ldg_stream4(t, S.x[0], S.x[1], S.x[2], S.x[3]);
            ldg_stream4(t + 4, S.x[4], S.x[5], S.x[6], S.x[7]);
          } else if (a.ld == LD_GALOIS) {
            const u64 *t = a.tgt + (size_t)J * N;
_Pragma("unroll")
            for (int e = 0; e < 8; e++) S.x[e] = ldg_stream(t + galois_src_index((u32)(base + e), a.elt, T.logN));
          } else {
            const u64 *t = a.tgt + (size_t)J * N + base;
            const u64 *t2 = a.tgt2 + (size_t)J * N + base;
            u64 u[8], v[8];
            ldg_stream4(t, u[0], u[1], u[2], u[3]);
            ldg_stream4(t + 4, u[4], u[5], u[6], u[7]);
            ldg_stream4(t2, v[0], v[1], v[2], v[3]);
            ldg_stream4(t2 + 4, v[4], v[5], v[6], v[7]);
_Pragma("unroll")
            for (int e = 0; e < 8; e++) S.x[e] = mulmod(u[e], v[e], m);
          }
        });
      } else {
        const u64 *src = a.src + ((size_t)Iidx * a.l + J) * N + r * 256;
        FOR_LANES(S, st, {
_Pragma("unroll")
          for (int e = 0; e < 8; e++) S.x[e] = ldg_stream(src + idxH(lane, e));
        });
        warp_fwdB8_regs(st, sm, r, tw, m.q);
      }
      const u64 *k0 = a.key + (((size_t)J * 2 + 0) * a.Ltot + I) * N + r * 256;
      const u64 *k1 = a.key + (((size_t)J * 2 + 1) * a.Ltot + I) * N + r * 256;
      FOR_LANES(S, st, {
        const int li = (NLANE_STATE == 1) ? 0 : lane;
        u64 ka[8], kb[8];
        ldg_stream4(k0 + lane * 8, ka[0], ka[1], ka[2], ka[3]);
        ldg_stream4(k0 + lane * 8 + 4, ka[4], ka[5], ka[6], ka[7]);
        ldg_stream4(k1 + lane * 8, kb[0], kb[1], kb[2], kb[3]);
        ldg_stream4(k1 + lane * 8 + 4, kb[4], kb[5], kb[6], kb[7]);
_Pragma("unroll")
        for (int e = 0; e < 8; e++) {
          mac128(lo0[li][e], hi0[li][e], S.x[e], ka[e]);
          mac128(lo1[li][e], hi1[li][e], S.x[e], kb[e]);
        }
      });
    }
    u64 *o0 = a.dst + ((size_t)0 * (a.l + 1) + Iidx) * N + r * 256;
    u64 *o1 = a.dst + ((size_t)1 * (a.l + 1) + Iidx) * N + r * 256;
    FOR_LANES(S, st, {
      (void)S;
      const int li = (NLANE_STATE == 1) ? 0 : lane;
      u64 v0[8], v1[8];
_Pragma("unroll")
      for (int e = 0; e < 8; e++) {
        v0[e] = reduce128(lo0[li][e], hi0[li][e], m);
        v1[e] = reduce128(lo1[li][e], hi1[li][e], m);
      }
      stg4(o0 + lane * 8, v0[0], v0[1], v0[2], v0[3]);
      stg4(o0 + lane * 8 + 4, v0[4], v0[5], v0[6], v0[7]);
      stg4(o1 + lane * 8, v1[0], v1[1], v1[2], v1[3]);
      stg4(o1 + lane * 8 + 4, v1[4], v1[5], v1[6], v1[7]);
    });
    return;
  }

  // --- single-limb variants: d indexes the limb ---
  if (EPI == EPI_CANON) {
    const int p = a.prime0 + d * a.pstep;
    const ModQ m = T.mod[p];
    const u64 q = m.q, q2 = 2 * m.q;
    const u64 *src = a.src + (size_t)d * N + r * 256;
    FOR_LANES(S, st, {
_Pragma("unroll")
      for (int e = 0; e < 8; e++) S.x[e] = ldg_stream(src + idxH(lane, e));
    });
    warp_fwdB8_regs(st, sm, r, T.tw + (size_t)p * N, q);
    u64 *o = a.dst + (size_t)d * N + r * 256;
    FOR_LANES(S, st, {
      u64 v[8];
_Pragma("unroll")
      for (int e = 0; e < 8; e++) v[e] = csub(csub(S.x[e], q2), q);
      stg4(o + lane * 8, v[0], v[1], v[2], v[3]);
      stg4(o + lane * 8 + 4, v[4], v[5], v[6], v[7]);
    });
    return;
  }
  if (EPI == EPI_MODDOWN_RELIN) {
    // d = limb i; both output polys are produced by the same warp so that every input
    // (a0,a1,b0,b1 at this position) is read before either output is written (dst may alias a or b).
    const int i = d;
    const ModQ m = T.mod[i];
    const u64 q = m.q, q2 = 2 * m.q;
    const Tw inv = T.qinv[a.plast][i];
    const size_t off = (size_t)i * N + r * 256;
    for (int K = 0; K < 2; K++) {
      const u64 *src = a.src + ((size_t)K * a.l + i) * N + r * 256;
      FOR_LANES(S, st, {
        if (K == 1) {
_Pragma("unroll")
          for (int e = 0; e < 8; e++) S.z[e] = S.x[e];
        }
_Pragma("unroll")
        for (int e = 0; e < 8; e++) S.x[e] = ldg_stream(src + idxH(lane, e));
      });
      warp_fwdB8_regs(st, sm, r, T.tw + (size_t)i * N, q);
    }
    FOR_LANES(S, st, {
      const int b = lane * 8;
      const u64 *c0p = a.acc + ((size_t)0 * (a.l + 1) + i) * N + r * 256 + b;
      const u64 *c1p = a.acc + ((size_t)1 * (a.l + 1) + i) * N + r * 256 + b;
      const u64 *pa0 = a.add0 + off + b, *pa1 = a.add0 + a.pitch + off + b;
      const u64 *pb0 = a.add1 + off + b, *pb1 = a.add1 + a.pitch + off + b;
      u64 c0[8], c1[8], a0[8], a1[8], b0[8], b1[8], v0[8], v1[8];
      ldg_stream4(c0p, c0[0], c0[1], c0[2], c0[3]);
      ldg_stream4(c0p + 4, c0[4], c0[5], c0[6], c0[7]);
      ldg_stream4(c1p, c1[0], c1[1], c1[2], c1[3]);
      ldg_stream4(c1p + 4, c1[4], c1[5], c1[6], c1[7]);
      ldg_stream4(pa0, a0[0], a0[1], a0[2], a0[3]);
      ldg_stream4(pa0 + 4, a0[4], a0[5], a0[6], a0[7]);
      ldg_stream4(pa1, a1[0], a1[1], a1[2], a1[3]);
      ldg_stream4(pa1 + 4, a1[4], a1[5], a1[6], a1[7]);
      ldg_stream4(pb0, b0[0], b0[1], b0[2], b0[3]);
      ldg_stream4(pb0 + 4, b0[4], b0[5], b0[6], b0[7]);
      ldg_stream4(pb1, b1[0], b1[1], b1[2], b1[3]);
      ldg_stream4(pb1 + 4, b1[4], b1[5], b1[6], b1[7]);
_Pragma("unroll")
      for (int e = 0; e < 8; e++) {
        u64 u0 = csub(csub(S.z[e], q2), q), u1 = csub(csub(S.x[e], q2), q);
        u64 t0 = shoup_mul(c0[e] + q - u0, inv, q);
        u64 t1 = shoup_mul(c1[e] + q - u1, inv, q);
        v0[e] = csub(t0 + mulmod(a0[e], b0[e], m), q);
        // d1 = a0*b1 + a1*b0 : one 128-bit sum, one Barrett (same canonical value as two mulmods + add)
        u64 lo = 0, hi = 0;
        mac128(lo, hi, a0[e], b1[e]);
        mac128(lo, hi, a1[e], b0[e]);
        v1[e] = csub(t1 + reduce128(lo, hi, m), q);
      }
      u64 *o0 = a.dst + off + b, *o1 = a.dst + a.pitch + off + b;
      stg4(o0, v0[0], v0[1], v0[2], v0[3]);
      stg4(o0 + 4, v0[4], v0[5], v0[6], v0[7]);
      stg4(o1, v1[0], v1[1], v1[2], v1[3]);
      stg4(o1 + 4, v1[4], v1[5], v1[6], v1[7]);
    });
    return;
  }
  // MODDOWN_GALOIS / RESCALE: d = K*l + i
  {
    const int K = d / a.l, i = d - K * a.l;
    const ModQ m = T.mod[i];
    const u64 q = m.q, q2 = 2 * m.q;
    const u64 *src = a.src + (size_t)d * N + r * 256;
    FOR_LANES(S, st, {
_Pragma("unroll")
      for (int e = 0; e < 8; e++) S.x[e] = ldg_stream(src + idxH(lane, e));
    });
    warp_fwdB8_regs(st, sm, r, T.tw + (size_t)i * N, q);
    const Tw inv = T.qinv[a.plast][i];
    const size_t off = (size_t)i * N + r * 256;
    u64 *o = a.dst + (size_t)K * a.pitch + off;
    FOR_LANES(S, st, {
      const int b = lane * 8;
      u64 c[8], v[8];
      const u64 *cin = (EPI == EPI_RESCALE) ? a.add0 + (size_t)K * a.pitch + off + b
                                            : a.acc + ((size_t)K * (a.l + 1) + i) * N + r * 256 + b;
      ldg_stream4(cin, c[0], c[1], c[2], c[3]);
      ldg_stream4(cin + 4, c[4], c[5], c[6], c[7]);
_Pragma("unroll")
      for (int e = 0; e < 8; e++) {
        u64 u = csub(csub(S.x[e], q2), q);      // canonical NTT of the rounding term
        v[e] = shoup_mul(c[e] + q - u, inv, q); // (c - u) * plast^-1 mod q
      }
      if (EPI == EPI_MODDOWN_GALOIS && K == 0) {
        const u64 *p0 = a.add0 + off + b;
        u64 w[8];
        ldg_stream4(p0, w[0], w[1], w[2], w[3]);
        ldg_stream4(p0 + 4, w[4], w[5], w[6], w[7]);
_Pragma("unroll")
        for (int e = 0; e < 8; e++) v[e] = csub(v[e] + w[e], q);
      }
      stg4(o + b, v[0], v[1], v[2], v[3]);
      stg4(o + b + 4, v[4], v[5], v[6], v[7]);
    });
  }
}
