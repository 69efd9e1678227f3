// TEST INFRASTRUCTURE ONLY -- CPU warp emulator for the CUDA warp-job bodies.
//
// Compiles dacapo_b200/csrc/{ntt_core,ntt_bodies}.cuh and the launcher-templated
// orchestration (ops.hpp) for the host: every warp job is replayed as a 32-iteration loop
// over lane states with a plain array standing in for the warp's shared-memory tile.  This
// lets the CPU-only test tier (-m "not gpu") check the index maths, twiddle indexing and
// fusion logic of the kernels bit-for-bit against the oracle without a GPU.  It is never
// linked into libB200_HEVM.so and is not a fallback: the product has no host execution path.
#include "../../dacapo_b200/csrc/host_params.hpp"
#include "../../dacapo_b200/csrc/ops.hpp"
#include <cstdio>
#include <cstring>

struct EmuLauncher {
  bool use_fused = false;
  long waits_checked = 0;
  bool fused() const { return use_fused; }
  // single-launch key switch / rescale: the tickets are replayed IN ORDER by one "CTA"; every wait of a unit must
  // already be satisfied by the signals of lower tickets (that is the deadlock-freedom argument of ks_fused.cuh), and
  // at the end every counter must sit exactly at the value its consumers wait for.
  template <int LOGA, int MODE> void ks_fused(const KsFusedArgs &A) {
    const KsUnits K = ks_units<LOGA, MODE>(A);
    const size_t N = (size_t)1 << (LOGA + 8);
    const int L = A.Ltot;
    for (int ticket = 0; ticket < K.total; ticket++) {
      const KsUnit un = ks_decode(K, A.nct, ticket);
      const KsCt &c = A.ct[un.c];
      Scratch sc;
      sc.carve(c.scratch, L, N);
      const KsDeps d = ks_deps<LOGA, MODE>(A, un);
      for (int k = 0; k < d.nwait; k++, waits_checked++)
        if (sc.ctl[d.widx[k]] < d.wtarget[k]) {
          std::fprintf(stderr, "emul: ticket %d (phase %d unit %d) waits on counter %d = %u < %u: dependency order broken\n", ticket, un.phase, un.u,
                       d.widx[k], sc.ctl[d.widx[k]], d.wtarget[k]);
          std::abort();
        }
      if (un.phase == 2) {
        mac<LOGA>(ks_args_p3<MODE>(A, c, sc), 1, un.u);
      } else {
        for (int w = 0; w < 4; w++) {
          const int job = un.u * 4 + w;
          alignas(16) u64 sm[Geo<LOGA>::WARP_WORDS];
          if (un.phase == 0) {
            LaneB8 st[32];
            body_intt_B<LOGA, MODE == FUSED_RESCALE ? LD_PLAIN : MODE>(ks_args_p1<MODE>(A, c, sc), job, st, sm);
          } else if (un.phase == 1) {
            LaneA st[32];
            body_invA_fwdA<LOGA, PRE_MODUP>(ks_args_p2(A, sc), job, st, sm);
          } else if (un.phase == 3) {
            LaneA st[32];
            body_invA_fwdA<LOGA, PRE_ROUND>(ks_args_p4<MODE>(A, sc), job, st, sm);
          } else {
            LaneB8 st[32];
            body_fwd_B<LOGA, MODE == LD_GALOIS ? EPI_MODDOWN_GALOIS : MODE == LD_PRODUCT ? EPI_MODDOWN_RELIN : EPI_RESCALE>(ks_args_p5<MODE>(A, c, sc), job, st, sm);
          }
        }
      }
      for (int k = 0; k < d.sig_count; k++)
        if (k != d.sig_skip) sc.ctl[d.sig_first + k] += d.sig_inc;
    }
    for (int cc = 0; cc < A.nct; cc++) { // the last CTA's reset
      Scratch sc;
      sc.carve(A.ct[cc].scratch, L, N);
      for (int i = 0; i < ctl_count(L); i++) sc.ctl[i] = 0;
    }
  }
  template <int LOGA, int LD> void intt_B(const ArgsInttB &a, int njobs) {
    for (int j = 0; j < njobs; j++) {
      LaneB8 st[32];
      alignas(16) u64 sm[Geo<LOGA>::WARP_WORDS];
      body_intt_B<LOGA, LD>(a, j, st, sm);
    }
  }
  template <int LOGA> void intt_A(const ArgsInttA &a, int njobs) {
    for (int j = 0; j < njobs; j++) {
      LaneA st[32];
      alignas(16) u64 sm[Geo<LOGA>::WARP_WORDS];
      body_intt_A<LOGA>(a, j, st, sm);
    }
  }
  template <int LOGA, int PRE> void fwd_A(const ArgsFwdA &a, int njobs) {
    for (int j = 0; j < njobs; j++) {
      LaneA st[32];
      alignas(16) u64 sm[Geo<LOGA>::WARP_WORDS];
      body_fwd_A<LOGA, PRE>(a, j, st, sm);
    }
  }
  template <int LOGA, int EPI> void fwd_B(const ArgsFwdB &a, int njobs) {
    for (int j = 0; j < njobs; j++) {
      LaneB8 st[32];
      alignas(16) u64 sm[Geo<LOGA>::WARP_WORDS];
      body_fwd_B<LOGA, EPI>(a, j, st, sm);
    }
  }
  template <int LOGA, int PRE> void invA_fwdA(const ArgsInvFwdA &a, int njobs) {
    for (int j = 0; j < njobs; j++) {
      LaneA st[32];
      alignas(16) u64 sm[Geo<LOGA>::WARP_WORDS];
      body_invA_fwdA<LOGA, PRE>(a, j, st, sm);
    }
  }
  template <int LOGA> void mac(const ArgsFwdB &a, int njobs, int job0 = 0) { // one CTA of MAC_WARPS warps per job; CTA barriers = phase boundaries
    std::vector<u64> smv(mac_smem_words(a.l) + 2);
    u64 *sm = smv.data() + ((reinterpret_cast<uintptr_t>(smv.data()) & 8) ? 1 : 0); // 16-byte aligned
    for (int j = job0; j < job0 + njobs; j++) {
      Tw *tw_s = reinterpret_cast<Tw *>(sm);
      u64 *tiles = sm + MAC_TW_WORDS, *rowbufs = tiles + MAC_WARPS * TILE_B_WORDS, *xbuf = rowbufs + MAC_WARPS * MAC_ROW_WORDS;
      for (int tid = 0; tid < MAC_WARPS * 32; tid++) body_mac_stage<LOGA>(a, j, tid, tw_s);
      for (int w = 0; w < MAC_WARPS; w++) {
        LaneB8 st[32];
        body_mac_warp<LOGA>(a, j, w, st, tiles + w * TILE_B_WORDS, tw_s, rowbufs + w * MAC_ROW_WORDS, xbuf);
      }
      for (int tid = 0; tid < MAC_WARPS * 32; tid++) body_mac_dot<LOGA>(a, j, tid, xbuf, tiles);
      if (mac_Iidx<LOGA>(a, j) == a.l)
        for (int K = 0; K < 2; K++) {
          LaneB8 st[32];
          body_mac_tail<LOGA>(a, j, K, st, mac_tail_tile(rowbufs, K), mac_tail_tw(tiles, tw_s, K), tiles);
        }
    }
  }
};

struct Emu {
  hp::HostParams P;
  EmuLauncher la;
  OpsIface *ops;
  std::vector<u64> scratch;
};

extern "C" {
void *emul_create(int logN, int L, int bits) {
  auto e = new Emu();
  e->P.build(logN, L, bits);
  e->P.tab.tw = e->P.tw.data();
  e->P.tab.itw = e->P.itw.data();
  e->P.build_rowwise();
  e->P.tab.twB = e->P.twB.data();
  e->P.tab.itwB = e->P.itwB.data();
  e->ops = make_ops(e->la, &e->P.tab, logN, L);
  if (!e->ops) return nullptr;
  e->scratch.resize(Scratch::words(L, e->P.N));
  e->ops->sc.carve(e->scratch.data(), L, e->P.N);
  return e;
}
void emul_primes(void *h, u64 *q, u64 *psi) {
  auto e = (Emu *)h;
  for (int i = 0; i < e->P.L; i++) q[i] = e->P.q[i], psi[i] = e->P.psi[i];
}
void emul_ntt(void *h, u64 *data, int prime, int count, int inverse) {
  auto e = (Emu *)h;
  if (inverse)
    e->ops->ntt_inv(data, data, count, prime, 0);
  else
    e->ops->ntt_fwd(data, data, count, prime, 0);
}
// ciphertext operands: compact [2][l][N]; dst may alias a or b
void emul_keyswitch(void *h, int mode, const u64 *a, const u64 *b, u64 *dst, int l, const u64 *key, u32 elt) {
  auto e = (Emu *)h;
  // the product stores key-switch keys in radix-2^30 split form (ntt_bodies.cuh split30); `key` is the oracle's canonical key
  const size_t kw = (size_t)(e->P.L - 1) * 2 * e->P.L * e->P.N;
  std::vector<u64> ks(kw);
  for (size_t i = 0; i < kw; i++) ks[i] = split30(key[i]);
  e->ops->keyswitch(mode, a, b, dst, (size_t)l * e->P.N, l, ks.data(), elt);
}
// limb-sharded rotation key switch with `ranks` simulated ranks sharing one scratch area: every stage is run for all
// ranks before the next one starts, which is exactly what the all-gather / broadcast between the stages provide
void emul_keyswitch_sharded(void *h, int mode, const u64 *a, const u64 *b, u64 *dst, int l, const u64 *key, u32 elt, int ranks) {
  auto e = (Emu *)h;
  const size_t kw = (size_t)(e->P.L - 1) * 2 * e->P.L * e->P.N;
  std::vector<u64> ks(kw);
  for (size_t i = 0; i < kw; i++) ks[i] = split30(key[i]);
  for (int stage = 1; stage <= 3; stage++)
    for (int g = 0; g < ranks; g++) {
      const int tlo = (int)((long)(l + 1) * g / ranks), thi = (int)((long)(l + 1) * (g + 1) / ranks);
      e->ops->ks_shard_stage(stage, mode, a, b, dst, (size_t)l * e->P.N, l, ks.data(), elt, tlo, thi);
    }
}
// the same, with stage 2 cut like the peer-to-peer path does: mod-up pass A per SOURCE rank (digit ranges, own rank first),
// then the inner product
void emul_keyswitch_sharded_split(void *h, int mode, const u64 *a, const u64 *b, u64 *dst, int l, const u64 *key, u32 elt, int ranks) {
  auto e = (Emu *)h;
  const size_t kw = (size_t)(e->P.L - 1) * 2 * e->P.L * e->P.N;
  std::vector<u64> ks(kw);
  for (size_t i = 0; i < kw; i++) ks[i] = split30(key[i]);
  auto range = [&](int g, int &lo, int &hi) { lo = (int)((long)(l + 1) * g / ranks), hi = (int)((long)(l + 1) * (g + 1) / ranks); };
  for (int stage = 1; stage <= 3; stage++)
    for (int g = 0; g < ranks; g++) {
      int tlo, thi;
      range(g, tlo, thi);
      if (stage != 2) {
        e->ops->ks_shard_stage(stage, mode, a, b, dst, (size_t)l * e->P.N, l, ks.data(), elt, tlo, thi);
        continue;
      }
      for (int k = 0; k < ranks; k++) {
        int rlo, rhi;
        range((g + k) % ranks, rlo, rhi);
        if (rhi > l) rhi = l;
        if (rhi <= rlo) continue;
        e->ops->shard_j0 = rlo, e->ops->shard_nj = rhi - rlo;
        e->ops->ks_shard_stage(20, mode, a, b, dst, (size_t)l * e->P.N, l, ks.data(), elt, tlo, thi);
      }
      e->ops->ks_shard_stage(21, mode, a, b, dst, (size_t)l * e->P.N, l, ks.data(), elt, tlo, thi);
    }
}
void emul_set_fused(void *h, int on) { ((Emu *)h)->la.use_fused = on != 0; }
// shared-memory layout functions of the kernels (bank-conflict checks of tests/test_kernel_emul.py)
int emul_xpad(int i) { return xpad(i); }
int emul_twB_pos(int k, int g) { return twB_pos(k, g); }
int emul_layout_const(int which) { return which == 0 ? TWB_ROW : which == 1 ? MAC_XROW : which == 2 ? MAC_TW_WORDS : TILE_B_WORDS; }
// warp-job target of the fused pass-A launches (OpsIface::group_warps): 1 184 = lone op, 148 = scheduled program
void emul_set_group_warps(void *h, int w) { ((Emu *)h)->ops->group_warps = w; }
long emul_waits_checked(void *h) { return ((Emu *)h)->la.waits_checked; }
// batched single-launch key switch: n ciphertexts (compact [2][l][N] each, contiguous), one scratch area per ciphertext
void emul_keyswitch_batch(void *h, int mode, int n, const u64 *a, const u64 *b, u64 *dst, int l, const u64 *key, const u32 *elts) {
  auto e = (Emu *)h;
  const size_t kw = (size_t)(e->P.L - 1) * 2 * e->P.L * e->P.N, ctw = (size_t)2 * l * e->P.N, sw = Scratch::words(e->P.L, e->P.N);
  std::vector<u64> ks(kw), scr(sw * n, 0);
  for (size_t i = 0; i < kw; i++) ks[i] = split30(key[i]);
  std::vector<KsCt> items(n);
  for (int k = 0; k < n; k++) items[k] = KsCt{a + k * ctw, b ? b + k * ctw : nullptr, dst + k * ctw, ks.data(), scr.data() + k * sw, elts ? elts[k] : 0, 0};
  const bool was = e->la.use_fused;
  e->la.use_fused = true;
  e->ops->keyswitch_batch(mode, n, items.data(), (size_t)l * e->P.N, l);
  e->la.use_fused = was;
}
void emul_rescale(void *h, const u64 *src, u64 *dst, int l) {
  auto e = (Emu *)h;
  e->ops->rescale(src, (size_t)l * e->P.N, dst, (size_t)(l - 1) * e->P.N, l);
}
}
