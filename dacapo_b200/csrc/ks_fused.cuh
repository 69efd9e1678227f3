// Single-launch key switch / rescale: the five (three) dependent kernel launches of ops.hpp::keyswitch (rescale)
// become ONE persistent kernel whose CTAs pull work units from a global ticket counter, in an order in which every
// unit only depends on units with smaller tickets.  A unit waits for its producers on per-limb completion counters
// in global memory (release: stores -> __threadfence -> atomicAdd; acquire: ld.acquire poll -> data loads through L2),
// so consecutive phases overlap limb by limb instead of being separated by launch boundaries, and nothing requires
// co-residency of the whole grid (a ticket holder is always resident, lower tickets never wait on higher ones).
//
// Reference semantics unchanged (SEAL 4.0 Evaluator::switch_key_inplace / divide_and_round_q_last_ntt_inplace as used
// by lib/Runtime/SEAL_HEVM.cpp:273,283,315-316): the units ARE the warp-job bodies of
// ntt_bodies.cuh, only their sequencing changes.
//
//   key switch (MODE = LD_GALOIS | LD_PRODUCT), per ciphertext of the batch:
//     P1  inverse pass B of the l target limbs (+ Galois gather / tensor product d2)       -> s1      signals c1[J]
//     P2  inverse pass A + mod-up + forward pass A of limb J for a group of targets         -> s2      waits c1[J], signals c2[I]
//     P3  forward pass B of all digits + key inner product for (target I, row r)            -> acc     waits c2[I], signals c3[I]
//         (special prime: + inverse pass B of the two accumulator rows                     -> t[0..1])
//     P4  inverse pass A + rounding + forward pass A of the special limb, group of targets -> s4      waits c3[l], signals c4[K][i]
//     P5  forward pass B + (acc - u) * p^-1 + addend                                        -> dst     waits c4[K][i], c3[i]
//   rescale (MODE = FUSED_RESCALE): P1 (last limb of both polys) -> P4 (waits c1[K]) -> P5 with the rescale epilogue.
//
// The ticket decoding, the argument structs of every phase and the wait / signal lists are plain HD functions shared
// with the test-only CPU emulator, which replays the tickets in order and ASSERTS that every wait is already satisfied
// -- i.e. it checks the dependency bookkeeping without a GPU.
#pragma once
#include "ntt_bodies.cuh"

#define FUSED_RESCALE 4
#define KS_MAX_BATCH 32
#define KS_CTL_WORDS 128 // u64 words reserved per scratch area for the counters (5 L + 4 u32 <= 164 for L = 32)

struct Scratch {      // device (or emulator-host) buffers, sized for the top level
  u64 *s1 = nullptr;  // [L][N]      inverse pass-B output
  u64 *t = nullptr;   // [L][N]      coefficient-form digits (sharded path) / special-limb rows of the fused key switch
  u64 *s2 = nullptr;  // [L][L-1][N] forward pass-A output of every (I,J) digit  (also NTT staging)
  u64 *acc = nullptr; // [2][L][N]   key-switch accumulators
  u64 *s4 = nullptr;  // [2][L][N]   forward pass-A output of the rounding terms
  u64 *pc0 = nullptr; // [L][N]      permuted c0 (rotate)
  u64 *rnd = nullptr; // [2][N]      limb-sharded key switch: rounded special-limb coefficients (broadcast by their owner)
  unsigned *ctl = nullptr; // completion counters + ticket of the fused kernels (all zero between launches)
  static HD size_t words(int L, size_t N) { return ((size_t)L + L + (size_t)L * (L - 1) + 2 * L + 2 * L + L + 2) * N + KS_CTL_WORDS; }
  HD void carve(u64 *base, int L, size_t N) {
    s1 = base;
    t = s1 + (size_t)L * N;
    s2 = t + (size_t)L * N;
    acc = s2 + (size_t)L * (L - 1) * N;
    s4 = acc + (size_t)2 * L * N;
    pc0 = s4 + (size_t)2 * L * N;
    rnd = pc0 + (size_t)L * N;
    ctl = reinterpret_cast<unsigned *>(rnd + 2 * N);
  }
};

struct KsCt { // one ciphertext of a batched launch
  const u64 *a, *b; // key switch: operands (b only for LD_PRODUCT); rescale: a = source, b = optional plaintext added to poly 0
  u64 *dst;
  const u64 *key;   // key-switch key of this ciphertext (rotations of one batch may use different Galois keys)
  u64 *scratch;     // base of a Scratch area
  u32 elt, pad;
};
struct KsFusedArgs {
  const NttTables *T;
  int l;            // key switch: level; rescale: level of the INPUT (l-1 limbs come out)
  int sp, Ltot, nct, ng2, ng4;
  size_t pitch;     // poly pitch of ciphertext operands / key-switch output
  size_t spitch;    // rescale: poly pitch of the input (0 = pitch)
  KsCt ct[KS_MAX_BATCH];
};

// counter indices inside Scratch::ctl (u32 words); L = Ltot
HD int ctl_c1(int /*L*/, int J) { return J; }
HD int ctl_c2(int L, int t) { return L + t; }
HD int ctl_c3(int L, int t) { return 2 * L + 1 + t; }
HD int ctl_c4(int L, int K, int i) { return 3 * L + 2 + K * L + i; }
HD int ctl_ticket(int L) { return 5 * L + 2; }
HD int ctl_nexit(int L) { return 5 * L + 3; }
HD int ctl_count(int L) { return 5 * L + 4; }

struct KsUnits {
  int u[5];
  int per_ct, total;
};
template <int LOGA, int MODE> HD KsUnits ks_units(const KsFusedArgs &A) {
  constexpr int ROWS = Geo<LOGA>::ROWS, TILES = Geo<LOGA>::TILES;
  KsUnits k;
  if (MODE == FUSED_RESCALE) {
    k.u[0] = 2 * ROWS / 4, k.u[1] = 0, k.u[2] = 0, k.u[3] = 2 * A.ng4 * TILES / 4, k.u[4] = 2 * (A.l - 1) * ROWS / 4;
  } else {
    k.u[0] = A.l * ROWS / 4, k.u[1] = A.l * A.ng2 * TILES / 4, k.u[2] = (A.l + 1) * ROWS, k.u[3] = 2 * A.ng4 * TILES / 4;
    k.u[4] = (MODE == LD_GALOIS ? 2 * A.l : A.l) * ROWS / 4;
  }
  k.per_ct = k.u[0] + k.u[1] + k.u[2] + k.u[3] + k.u[4];
  k.total = k.per_ct * A.nct;
  return k;
}
struct KsUnit {
  int phase, c, u; // phase 0..4, ciphertext, unit inside (phase, ciphertext)
};
// tickets are phase-major: all P1 units of the batch, then all P2 units, ...
HD KsUnit ks_decode(const KsUnits &k, int nct, int ticket) {
  KsUnit r{0, 0, ticket};
  for (int p = 0; p < 5; p++) {
    const int n = k.u[p] * nct;
    if (r.u < n) {
      r.phase = p;
      r.c = r.u / k.u[p];
      r.u -= r.c * k.u[p];
      return r;
    }
    r.u -= n;
  }
  r.phase = 5;
  return r;
}

// ---- argument structs of the phases (same fields ops.hpp fills for the separate launches) ----
template <int MODE> HD ArgsInttB ks_args_p1(const KsFusedArgs &A, const KsCt &c, const Scratch &sc) {
  ArgsInttB x{};
  x.T = A.T, x.dst = sc.s1;
  if (MODE == FUSED_RESCALE) {
    const size_t N = (size_t)1 << A.T->logN;
    x.src = c.a + (size_t)(A.l - 1) * N, x.sstride = A.spitch ? A.spitch : A.pitch, x.nl = 2, x.prime0 = A.l - 1, x.pstep = 0;
  } else {
    x.nl = A.l, x.prime0 = 0, x.pstep = 1, x.elt = c.elt;
    x.src = c.a + A.pitch;
    if (MODE == LD_GALOIS)
      x.c0 = c.a, x.pc0 = sc.pc0;
    else
      x.src2 = c.b + A.pitch;
  }
  return x;
}
HD ArgsInvFwdA ks_args_p2(const KsFusedArgs &A, const Scratch &sc) {
  ArgsInvFwdA x{};
  x.T = A.T, x.src = sc.s1, x.dst = sc.s2, x.nsrc = A.l, x.l = A.l, x.sp = A.sp, x.ngroups = A.ng2;
  return x;
}
template <int MODE> HD ArgsFwdB ks_args_p3(const KsFusedArgs &A, const KsCt &c, const Scratch &sc) {
  ArgsFwdB x{};
  x.T = A.T, x.src = sc.s2, x.dst = sc.acc, x.l = A.l, x.sp = A.sp, x.key = c.key, x.Ltot = A.Ltot, x.ld = MODE, x.elt = c.elt;
  x.tgt = c.a + A.pitch, x.tgt2 = (MODE == LD_PRODUCT) ? c.b + A.pitch : nullptr;
  x.sp_rows = sc.t; // NOT s1: P2 units of other target groups may still be reading s1[0..1]
  return x;
}
template <int MODE> HD ArgsInvFwdA ks_args_p4(const KsFusedArgs &A, const Scratch &sc) {
  ArgsInvFwdA x{};
  x.T = A.T, x.dst = sc.s4, x.nsrc = 2, x.ngroups = A.ng4;
  if (MODE == FUSED_RESCALE)
    x.src = sc.s1, x.l = A.l - 1, x.plast = A.l - 1;
  else
    x.src = sc.t, x.l = A.l, x.plast = A.sp;
  return x;
}
template <int MODE> HD ArgsFwdB ks_args_p5(const KsFusedArgs &A, const KsCt &c, const Scratch &sc) {
  ArgsFwdB w{};
  w.T = A.T, w.src = sc.s4, w.dst = c.dst, w.pitch = A.pitch;
  if (MODE == FUSED_RESCALE) {
    w.l = A.l - 1, w.add0 = c.a, w.spitch = A.spitch, w.plast = A.l - 1, w.add1 = c.b;
  } else {
    w.l = A.l, w.sp = A.sp, w.acc = sc.acc, w.plast = A.sp;
    if (MODE == LD_GALOIS)
      w.add0 = sc.pc0;
    else
      w.add0 = c.a, w.add1 = c.b;
  }
  return w;
}

// ---- dependencies: what a unit waits for / signals (indices into ctl, u32 targets / increments) ----
struct KsDeps {
  int nwait, widx[3];
  unsigned wtarget[3];
  int sig_first, sig_count, sig_skip; // signals: ctl[sig_first + k] += sig_inc for k < sig_count, k != sig_skip
  unsigned sig_inc;
};
template <int LOGA, int MODE> HD KsDeps ks_deps(const KsFusedArgs &A, const KsUnit &un) {
  constexpr int ROWS = Geo<LOGA>::ROWS, TILES = Geo<LOGA>::TILES, LOGT = Geo<LOGA>::LOGT;
  const int L = A.Ltot;
  KsDeps d{};
  d.sig_skip = -1;
  const int j0 = un.u * 4; // first warp job of a 4-warp unit
  switch (un.phase) {
  case 0: { // P1: limb = job / ROWS
    const int limb = j0 >> LOGA;
    d.sig_first = ctl_c1(L, limb), d.sig_count = 1, d.sig_inc = 4;
    break;
  }
  case 1: { // P2: job = ((sl * ng2) + grp) * TILES + tile ; targets [grp*tpj, ...) of l+1, the diagonal (target == sl) is skipped
    const int rest = j0 >> LOGT, sl = rest / A.ng2, grp = rest - sl * A.ng2;
    const int ntargets = A.l + 1, tpj = (ntargets + A.ng2 - 1) / A.ng2;
    const int t0 = grp * tpj, t1 = (grp + 1) * tpj < ntargets ? (grp + 1) * tpj : ntargets;
    d.nwait = 1, d.widx[0] = ctl_c1(L, sl), d.wtarget[0] = ROWS;
    d.sig_first = ctl_c2(L, t0), d.sig_count = t1 > t0 ? t1 - t0 : 0, d.sig_inc = 4;
    if (sl >= t0 && sl < t1) d.sig_skip = sl - t0; // prime_of(t) == sl  <=>  t == sl (t < l)
    break;
  }
  case 2: { // P3: one CTA job = (Iidx, row), Iidx = l - job / ROWS
    const int Iidx = A.l - (un.u >> LOGA);
    const unsigned need = (unsigned)((A.l - (Iidx < A.l ? 1 : 0)) * TILES);
    if (need) d.nwait = 1, d.widx[0] = ctl_c2(L, Iidx), d.wtarget[0] = need;
    d.sig_first = ctl_c3(L, Iidx), d.sig_count = 1, d.sig_inc = 1;
    break;
  }
  case 3: { // P4: job = ((K * ng4) + grp) * TILES + tile ; targets [grp*tpj, ...) of nlim
    const int rest = j0 >> LOGT, K = rest / A.ng4, grp = rest - K * A.ng4;
    const int ntargets = (MODE == FUSED_RESCALE) ? A.l - 1 : A.l, tpj = (ntargets + A.ng4 - 1) / A.ng4;
    const int t0 = grp * tpj, t1 = (grp + 1) * tpj < ntargets ? (grp + 1) * tpj : ntargets;
    d.nwait = 1, d.wtarget[0] = ROWS;
    d.widx[0] = (MODE == FUSED_RESCALE) ? ctl_c1(L, K) : ctl_c3(L, A.l);
    d.sig_first = ctl_c4(L, K, t0), d.sig_count = t1 > t0 ? t1 - t0 : 0, d.sig_inc = 4;
    break;
  }
  default: { // P5
    const int dd = j0 >> LOGA;
    if (MODE == LD_PRODUCT) { // job = i * ROWS + row, both polys in one job
      d.nwait = 3;
      d.widx[0] = ctl_c4(L, 0, dd), d.wtarget[0] = TILES;
      d.widx[1] = ctl_c4(L, 1, dd), d.wtarget[1] = TILES;
      d.widx[2] = ctl_c3(L, dd), d.wtarget[2] = ROWS;
    } else {
      const int nlim = (MODE == FUSED_RESCALE) ? A.l - 1 : A.l;
      const int K = dd / nlim, i = dd - K * nlim;
      d.nwait = 1, d.widx[0] = ctl_c4(L, K, i), d.wtarget[0] = TILES;
      if (MODE == LD_GALOIS) d.nwait = 2, d.widx[1] = ctl_c3(L, i), d.wtarget[1] = ROWS;
    }
    break;
  }
  }
  return d;
}
