// Host-side orchestration of the NTT-based operations: which warp-job kernels run, in
// which order, over which scratch buffers.  Templated on a Launcher so that the very same
// sequencing is exercised by the GPU launcher (kernels.cu) and by the test-only CPU warp
// emulator (tests/emul/warp_emul.cpp).  No arithmetic happens here.
//
// Reference semantics (SEAL 4.0 via lib/Runtime/SEAL_HEVM.cpp):
//   keyswitch : Evaluator::switch_key_inplace   (rotate: SEAL_HEVM.cpp:273, mulcc+relin: 315-316)
//   rescale   : RNSTool::divide_and_round_q_last_ntt_inplace (SEAL_HEVM.cpp:283, also inside encrypt)
#pragma once
#include "ks_fused.cuh"

// ring-size independent interface (the VM holds one of these per lane)
struct OpsIface {
  Scratch sc;
  int shard_j0 = 0, shard_nj = 0; // ks_shard_stage(20, ...): the digit range of this call
  int key_L = 0, key_t0 = 0; // limb-sharded KEY STORAGE (this rank stores limbs [key_t0, key_t0 + key_L) of every key); 0 = whole keys
  // warp jobs a fused inverse+forward pass-A launch should reach (see pick_groups): more groups = lower latency of a lone
  // op but the inverse pass is recomputed per group; the VM lowers it while many independent ops are in flight
#ifndef GROUP_TARGET_WARPS
#define GROUP_TARGET_WARPS 1184
#endif
  int group_warps = GROUP_TARGET_WARPS;
  virtual ~OpsIface() {}
  // limb k is transformed under prime prime0 + (pmod ? k % pmod : k) * pstep (pmod: several polynomials of pmod limbs each)
  virtual void ntt_fwd(const u64 *src, u64 *dst, int nl, int prime0, int pstep, int pmod = 0) = 0;
  virtual void ntt_inv(const u64 *src, u64 *dst, int nl, int prime0, int pstep, int round = 0) = 0;
  // coefficient form of the decrypted plaintext c0 + c1 * s of a level-l ciphertext (both polys at `pitch`), canonical
  virtual void decrypt_inv(const u64 *ct, size_t pitch, const u64 *sk, u64 *dst, int l) = 0;
  virtual void keyswitch(int mode, const u64 *a, const u64 *b, u64 *dst, size_t pitch, int l, const u64 *key, u32 elt) = 0;
  // add_pt (optional, [l-1][N]): added to polynomial 0 of the result (the plaintext of a public-key encryption)
  virtual void rescale(const u64 *src, size_t src_pitch, u64 *dst, size_t dst_pitch, int l, const u64 *add_pt = nullptr) = 0;
  // limb-sharded key switch: this rank owns the key-switch targets [tlo, thi) of [0, l]
  // (target l = the special prime); between the stages the caller exchanges sc.t (all-gather) and sc.rnd (broadcast)
  // mode LD_GALOIS: rotation of ciphertext `a` (b unused); mode LD_PRODUCT: multiply a*b + relinearise
  virtual void ks_shard_stage(int stage, int mode, const u64 *a, const u64 *b, u64 *dst, size_t pitch, int l, const u64 *key, u32 elt, int tlo, int thi) = 0;
  // batched single-launch forms (independent ciphertexts of one level in ONE kernel; every item brings its own scratch
  // area, see KsCt): key switch of `n` ciphertexts / rescale of `n` ciphertexts (item.a = source, item.b = optional plaintext)
  virtual void keyswitch_batch(int mode, int n, const KsCt *items, size_t pitch, int l) = 0;
  virtual void rescale_batch(int n, const KsCt *items, size_t src_pitch, size_t dst_pitch, int l) = 0;
};

template <class LA, int LOGA> struct HeOps : OpsIface {
  static constexpr int ROWS = Geo<LOGA>::ROWS, TILES_A = Geo<LOGA>::TILES;
  LA &la;
  const NttTables *T; // device-visible tables
  int logN, L;
  size_t N;
  HeOps(LA &l, const NttTables *tab, int nprimes) : la(l), T(tab), logN(LOGA + 8), L(nprimes), N((size_t)1 << (LOGA + 8)) {}
  int sp() const { return L - 1; }

  // forward NTT of nl limbs (limb k under prime prime0 + k*pstep); src may equal dst; canonical output
  void ntt_fwd(const u64 *src, u64 *dst, int nl, int prime0, int pstep, int pmod = 0) override {
    for (int done = 0; done < nl;) { // staged through s2 in chunks
      int chunk = nl - done;
      int cap = L * (L - 1);
      if (chunk > cap) chunk = cap;
      ArgsFwdA a{};
      a.T = T, a.src = src + (size_t)done * N, a.dst = sc.s2, a.nd = chunk, a.prime0 = prime0 + (pmod ? 0 : done * pstep), a.pstep = pstep;
      a.pmod = pmod; // (with pmod the whole batch must fit one chunk: checked below)
      la.template fwd_A<LOGA, PRE_NONE>(a, chunk * TILES_A);
      ArgsFwdB b{};
      b.T = T, b.src = sc.s2, b.dst = dst + (size_t)done * N, b.nd = chunk, b.prime0 = a.prime0, b.pstep = pstep, b.pmod = pmod;
      la.template fwd_B<LOGA, EPI_CANON>(b, chunk * ROWS);
      done += chunk;
    }
  }
  // inverse NTT (canonical coefficients).  round=1 adds floor(q/2) mod q.
  void ntt_inv(const u64 *src, u64 *dst, int nl, int prime0, int pstep, int round = 0) override {
    for (int done = 0; done < nl;) {
      int chunk = nl - done;
      int cap = L * (L - 1);
      if (chunk > cap) chunk = cap;
      ArgsInttB a{};
      a.T = T, a.src = src + (size_t)done * N, a.dst = sc.s2, a.nl = chunk, a.prime0 = prime0 + done * pstep, a.pstep = pstep;
      la.template intt_B<LOGA, LD_PLAIN>(a, chunk * ROWS);
      ArgsInttA b{};
      b.T = T, b.src = sc.s2, b.dst = dst + (size_t)done * N, b.nl = chunk, b.prime0 = a.prime0, b.pstep = pstep, b.round = round;
      la.template intt_A<LOGA>(b, chunk * TILES_A);
      done += chunk;
    }
  }

  void decrypt_inv(const u64 *ct, size_t pitch, const u64 *sk, u64 *dst, int l) override {
    ArgsInttB a{};
    a.T = T, a.src = ct + pitch, a.src2 = sk, a.c0 = ct, a.dst = sc.s2, a.nl = l, a.prime0 = 0, a.pstep = 1;
    la.template intt_B<LOGA, LD_DECRYPT>(a, l * ROWS);
    ArgsInttA b{};
    b.T = T, b.src = sc.s2, b.dst = dst, b.nl = l, b.prime0 = 0, b.pstep = 1, b.round = 0;
    la.template intt_A<LOGA>(b, l * TILES_A);
  }
  // how many target groups the fused inverse+forward pass-A kernel is split into: enough warp jobs to
  // cover the machine (148 SMs x 8 warps) without recomputing the inverse pass more than needed
  int pick_groups(int nsrc, int ntargets, int nct = 1) const { // nct: ciphertexts sharing the launch (batched forms)
    const int base = nsrc * TILES_A * nct;
    int g = (group_warps + base - 1) / base;
    if (g < 1) g = 1;
    if (g > ntargets) g = ntargets;
    return g;
  }

  // ---- key switching --------------------------------------------------------------------
  // mode LD_GALOIS : dst = (perm(c0), 0) + KS(perm(c1))   with src ciphertext `a`, Galois element elt
  // mode LD_PRODUCT: dst = (a0 b0, a0 b1 + a1 b0) + KS(a1 b1)           (multiply + relinearize)
  // `pitch` = words between the two polys of every ciphertext operand.  dst may alias a or b.
  // Five launches: B' | A'+mod-up+A | B+MAC(+B' of the special limb) | A'+round+A | B+mod-down epilogue
  void keyswitch(int mode, const u64 *a, const u64 *b, u64 *dst, size_t pitch, int l, const u64 *key, u32 elt) override {
    if (la.fused()) { // one persistent launch (ks_fused.cuh)
      KsCt it{a, b, dst, key, sc.s1, elt, 0};
      keyswitch_batch(mode, 1, &it, pitch, l);
      return;
    }
    // 1. inverse pass B of the target, fused with the Galois gather / the tensor product d2
    {
      ArgsInttB x{};
      x.T = T, x.dst = sc.s1, x.nl = l, x.prime0 = 0, x.pstep = 1, x.elt = elt;
      if (mode == LD_GALOIS) {
        x.src = a + pitch, x.c0 = a, x.pc0 = sc.pc0;
        la.template intt_B<LOGA, LD_GALOIS>(x, l * ROWS);
      } else {
        x.src = a + pitch, x.src2 = b + pitch;
        la.template intt_B<LOGA, LD_PRODUCT>(x, l * ROWS);
      }
    }
    // 2. inverse pass A -> coefficient digits t_J (registers only) -> (t_J mod q_I) -> forward pass A
    {
      ArgsInvFwdA x{};
      x.T = T, x.src = sc.s1, x.dst = sc.s2, x.nsrc = l, x.l = l, x.sp = sp(), x.ngroups = pick_groups(l, l + 1);
      la.template invA_fwdA<LOGA, PRE_MODUP>(x, l * x.ngroups * TILES_A);
    }
    // 3. forward pass B + inner product with the key over all digits; the special-prime CTAs also run the
    //    inverse pass B of their accumulator rows (first step of the mod-down) into s1[0..1]
    {
      ArgsFwdB x{};
      x.T = T, x.src = sc.s2, x.dst = sc.acc, x.l = l, x.sp = sp(), x.key = key, x.Ltot = L, x.ld = mode, x.elt = elt;
      x.tgt = a + pitch, x.tgt2 = (mode == LD_PRODUCT) ? b + pitch : nullptr;
      x.sp_rows = sc.s1;
      la.template mac<LOGA>(x, (l + 1) * ROWS);
    }
    // 4. mod-down: inverse pass A of the special limb + rounding + per-target fix-up + forward pass A
    {
      ArgsInvFwdA x{};
      x.T = T, x.src = sc.s1, x.dst = sc.s4, x.nsrc = 2, x.l = l, x.plast = sp(), x.ngroups = pick_groups(2, l);
      la.template invA_fwdA<LOGA, PRE_ROUND>(x, 2 * x.ngroups * TILES_A);
    }
    // 5. forward pass B + (acc - u) * p^-1 + addend
    {
      ArgsFwdB w{};
      w.T = T, w.src = sc.s4, w.dst = dst, w.l = l, w.sp = sp(), w.acc = sc.acc, w.pitch = pitch, w.plast = sp();
      if (mode == LD_GALOIS) {
        w.add0 = sc.pc0;
        la.template fwd_B<LOGA, EPI_MODDOWN_GALOIS>(w, 2 * l * ROWS);
      } else {
        w.add0 = a, w.add1 = b;
        la.template fwd_B<LOGA, EPI_MODDOWN_RELIN>(w, l * ROWS);
      }
    }
  }

  void keyswitch_batch(int mode, int n, const KsCt *items, size_t pitch, int l) override {
    for (int done = 0; done < n; done += KS_MAX_BATCH) {
      KsFusedArgs x{};
      x.T = T, x.l = l, x.sp = sp(), x.Ltot = L, x.pitch = pitch;
      x.nct = n - done < KS_MAX_BATCH ? n - done : KS_MAX_BATCH;
      x.ng2 = pick_groups(l, l + 1, x.nct), x.ng4 = pick_groups(2, l, x.nct);
      for (int k = 0; k < x.nct; k++) x.ct[k] = items[done + k];
      if (mode == LD_GALOIS)
        la.template ks_fused<LOGA, LD_GALOIS>(x);
      else
        la.template ks_fused<LOGA, LD_PRODUCT>(x);
    }
  }
  void rescale_batch(int n, const KsCt *items, size_t src_pitch, size_t dst_pitch, int l) override {
    for (int done = 0; done < n; done += KS_MAX_BATCH) {
      KsFusedArgs x{};
      x.T = T, x.l = l, x.sp = sp(), x.Ltot = L, x.pitch = dst_pitch, x.spitch = src_pitch;
      x.nct = n - done < KS_MAX_BATCH ? n - done : KS_MAX_BATCH;
      x.ng2 = 1, x.ng4 = pick_groups(2, l - 1, x.nct);
      for (int k = 0; k < x.nct; k++) x.ct[k] = items[done + k];
      la.template ks_fused<LOGA, FUSED_RESCALE>(x);
    }
  }

  // ---- limb-sharded rotation key switch ------------------------------------------------------------
  // The output limbs (targets) of SEAL's switch_key are partitioned over ranks; a rank holds / reads only its limbs
  // of the key and of the ciphertext.  Same arithmetic as keyswitch(LD_GALOIS), cut where data must cross ranks:
  //   stage 1  own data limbs of perm(c1): inverse NTT -> coefficient digits sc.t[J], perm(c0) -> sc.pc0[J]
  //            <all-gather of sc.t rows over the ranks>
  //   stage 2  own targets I: (t_J mod q_I) forward NTT for every digit J, inner product with key[J][.][I];
  //            owner of the special limb: inverse NTT + rounding of its two accumulator rows -> sc.rnd
  //            <broadcast of sc.rnd from the special limb's owner>
  //   stage 3  own data limbs: NTT of the rounding term, (acc - u) * p^-1 + perm(c0) -> dst limbs
  void ks_shard_stage(int stage, int mode, const u64 *a, const u64 *b, u64 *dst, size_t pitch, int l, const u64 *key, u32 elt, int tlo, int thi) override {
    const int dhi = thi < l ? thi : l, nd = dhi - tlo; // owned data limbs [tlo, dhi)
    const bool own_sp = thi == l + 1;
    if (stage == 1) {
      if (nd <= 0) return;
      ArgsInttB x{};
      x.T = T, x.dst = sc.s1 + (size_t)tlo * N, x.nl = nd, x.prime0 = tlo, x.pstep = 1, x.elt = elt;
      if (mode == LD_GALOIS) {
        x.src = a + pitch + (size_t)tlo * N, x.c0 = a + (size_t)tlo * N, x.pc0 = sc.pc0 + (size_t)tlo * N;
        la.template intt_B<LOGA, LD_GALOIS>(x, nd * ROWS);
      } else { // d2 = a1 * b1 on the own limbs
        x.src = a + pitch + (size_t)tlo * N, x.src2 = b + pitch + (size_t)tlo * N;
        la.template intt_B<LOGA, LD_PRODUCT>(x, nd * ROWS);
      }
      ArgsInttA y{};
      y.T = T, y.src = sc.s1 + (size_t)tlo * N, y.dst = sc.t + (size_t)tlo * N, y.nl = nd, y.prime0 = tlo, y.pstep = 1, y.round = 0;
      la.template intt_A<LOGA>(y, nd * TILES_A);
    } else if (stage == 2 || stage >= 20) {
      // stage 2 = 20 (mod-up pass A of the digits [j0, j0 + nj), all of them by default) + 21 (inner product, rounding)
      const int nt = thi - tlo;
      if (nt <= 0) return;
      if (stage != 21) {
        const int j0 = stage == 20 ? shard_j0 : 0, nj = stage == 20 ? shard_nj : l;
        if (nj > 0) {
          ArgsFwdA x{};
          x.T = T, x.src = sc.t, x.dst = sc.s2, x.l = l, x.sp = sp(), x.t0 = tlo, x.nt = nt, x.j0 = j0, x.nj = nj;
          la.template fwd_A<LOGA, PRE_MODUP>(x, nt * nj * TILES_A);
        }
        if (stage == 20) return;
      }
      ArgsFwdB m{};
      m.T = T, m.src = sc.s2, m.dst = sc.acc, m.l = l, m.sp = sp(), m.key = key, m.Ltot = key_L ? key_L : L, m.key_t0 = key_t0, m.ld = mode, m.elt = elt;
      m.tgt = a + pitch, m.tgt2 = (mode == LD_PRODUCT) ? b + pitch : nullptr, m.sp_rows = sc.s1, m.i_end = thi;
      la.template mac<LOGA>(m, nt * ROWS);
      if (own_sp) {
        ArgsInttA y{};
        y.T = T, y.src = sc.s1, y.dst = sc.rnd, y.nl = 2, y.prime0 = sp(), y.pstep = 0, y.round = 1;
        la.template intt_A<LOGA>(y, 2 * TILES_A);
      }
    } else {
      if (nd <= 0) return;
      ArgsFwdA x{};
      x.T = T, x.src = sc.rnd, x.dst = sc.s4, x.l = l, x.plast = sp(), x.t0 = tlo, x.nt = nd;
      la.template fwd_A<LOGA, PRE_ROUND>(x, 2 * nd * TILES_A);
      ArgsFwdB w{};
      w.T = T, w.src = sc.s4, w.dst = dst, w.l = l, w.sp = sp(), w.acc = sc.acc, w.pitch = pitch, w.plast = sp();
      w.t0 = tlo, w.nt = nd;
      if (mode == LD_GALOIS) {
        w.add0 = sc.pc0;
        la.template fwd_B<LOGA, EPI_MODDOWN_GALOIS>(w, 2 * nd * ROWS);
      } else { // (a0 b0, a0 b1 + a1 b0) + mod-down, own limbs only; dst may alias a or b
        w.add0 = a, w.add1 = b;
        la.template fwd_B<LOGA, EPI_MODDOWN_RELIN>(w, nd * ROWS);
      }
    }
  }

  // ---- rescale: 2 polys with l limbs -> l-1 limbs (divide by q_{l-1} and round); three launches -------
  // src/dst poly pitches may differ (encrypt uses a compact (l)-limb temporary).
  void rescale(const u64 *src, size_t src_pitch, u64 *dst, size_t dst_pitch, int l, const u64 *add_pt = nullptr) override {
    if (la.fused()) {
      KsCt it{src, add_pt, dst, nullptr, sc.s1, 0, 0};
      rescale_batch(1, &it, src_pitch, dst_pitch, l);
      return;
    }
    {
      ArgsInttB x{};
      x.T = T, x.src = src + (size_t)(l - 1) * N, x.sstride = src_pitch, x.dst = sc.s1, x.nl = 2, x.prime0 = l - 1, x.pstep = 0;
      la.template intt_B<LOGA, LD_PLAIN>(x, 2 * ROWS);
    }
    {
      ArgsInvFwdA x{};
      x.T = T, x.src = sc.s1, x.dst = sc.s4, x.nsrc = 2, x.l = l - 1, x.plast = l - 1, x.ngroups = pick_groups(2, l - 1);
      la.template invA_fwdA<LOGA, PRE_ROUND>(x, 2 * x.ngroups * TILES_A);
    }
    {
      ArgsFwdB w{};
      w.T = T, w.src = sc.s4, w.dst = dst, w.l = l - 1, w.add0 = src, w.pitch = dst_pitch, w.spitch = src_pitch, w.plast = l - 1, w.add1 = add_pt;
      la.template fwd_B<LOGA, EPI_RESCALE>(w, 2 * (l - 1) * ROWS);
    }
  }
};

// ring-size dispatch: N = 2^(LOGA+8)
template <class LA> OpsIface *make_ops(LA &la, const NttTables *tab, int logN, int nprimes) {
  switch (logN) {
  case 14: return new HeOps<LA, 6>(la, tab, nprimes);
  case 15: return new HeOps<LA, 7>(la, tab, nprimes);
  case 16: return new HeOps<LA, 8>(la, tab, nprimes);
  case 17: return new HeOps<LA, 9>(la, tab, nprimes);
  }
  return nullptr;
}
