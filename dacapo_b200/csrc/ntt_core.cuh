// Warp-synchronous negacyclic NTT building blocks (N = 2^(LOGA+8)).
//
// Replaces SEAL 4.0 ntt_negacyclic_harvey(_lazy) / inverse_ntt_negacyclic_harvey that the
// reference reaches through seal::Evaluator (reference: lib/Runtime/SEAL_HEVM.cpp:273,283,
// 315-316).  Same transform: forward = Cooley-Tukey with twiddle table tw[bitrev(k)] = psi^k,
// natural in -> bit-reversed out; inverse = Gentleman-Sande + N^-1.  Any schedule of the
// same butterfly DAG gives the same canonical residues, so the schedule is B200-first:
//
//   N = 2^LOGA (rows) x 256 (cols), LOGA = 6,7,8.  Two passes per transform, one HBM/L2 round trip between:
//   pass A  "strided":    a warp owns a tile of 2^LOGA rows x C cols (512 values, 16 per lane; C = 4 and
//                         32 B sectors at N = 2^15) and does the LOGA stages whose span is >= one row.
//   pass B  "contiguous": a warp owns one row (256 contiguous values, 8 per lane) and does the
//                         remaining 8 stages.
//   Inside a pass, values live in registers; a warp-private padded shared-memory tile is used
//   only to re-distribute values between rounds of 2-4 register-resident stages (__syncwarp,
//   never __syncthreads).  Shared tile index idx -> idx + idx/16 makes every 64-bit access
//   pattern used here bank-conflict free.
//   The twiddles a job needs (127 per pass-A prime, 255 per pass-B row) are staged into shared
//   memory once per job with cp.async, so no butterfly stage waits on a global-memory round trip.
//   Range control is lazy: butterflies never compare; values are folded into [0,2q) with
//   fold60() (q = 2^60 - delta) only where a round of stages could otherwise exceed 16q < 2^64.
//
// Everything is __host__ __device__: `FOR_LANES` runs the per-lane code on the 32 lanes of a
// real warp on the GPU, and as a 32-iteration loop over an array of lane states in the
// test-only CPU warp emulator (tests/emul/warp_emul.cpp).  The product never runs it on the host.
#pragma once
#include "modarith.cuh"

#define HEVM_MAXL 32

// Row-wise pass-B twiddle table: the 255 twiddles of one 256-point row, local stage k (2^k groups), group g.
// Stages 0..5 sit at 2^k - 1 + g.  In the last two stages every lane reads its own 2 / 4 consecutive entries (layout C:
// g = 2 lane + j / 4 lane + j), i.e. 32- / 64-byte lane strides that put 2 / 4 lanes of a quarter-warp on the same banks
// (LDS.128); one unused entry after every 2 / 4 entries makes the lane stride 48 / 80 bytes = conflict-free.
#define TWB_ROW 320 // entries per row (319 used)
HD constexpr int twB_pos(int k, int g) { return k < 6 ? (1 << k) - 1 + g : k == 6 ? 63 + g + (g >> 1) : 159 + g + (g >> 2); }
HD size_t twB_prime_stride(size_t N) { return (N >> 8) * TWB_ROW; }

struct NttTables {
  int logN, L;
  const Tw *tw;  // [L][N]  forward twiddles, tw[m+i] for stage with m groups, group i
  const Tw *itw; // [L][N]  inverses of the above
  const Tw *twB, *itwB; // [L][rows][TWB_ROW] the same values regrouped per pass-B row (host_params.hpp build_rowwise)
  ModQ mod[HEVM_MAXL];
  Tw invn[HEVM_MAXL];           // N^-1
  Tw invn_w[HEVM_MAXL];         // N^-1 * itw[1]
  Tw qinv[HEVM_MAXL][HEVM_MAXL]; // qinv[a][b] = q_a^-1 mod q_b  (a != b)
};

#if defined(__CUDA_ARCH__)
#define LANE_DECL const int lane = threadIdx.x & 31
#define FOR_LANES(S, st, ...)                                                                                         \
  {                                                                                                                    \
    auto &S = st[0];                                                                                                   \
    __VA_ARGS__;                                                                                                       \
    __syncwarp();                                                                                                      \
  }
#define NLANE_STATE 1
HD Tw ldtw(const Tw *p) { // twiddle from the warp's shared-memory copy (LDS.128)
  const ulonglong2 v = *reinterpret_cast<const ulonglong2 *>(p);
  Tw t;
  t.w = v.x;
  t.wq = v.y;
  return t;
}
HD void cp_async16(void *smem_dst, const void *gsrc) { // LDGSTS: global -> shared without registers
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gsrc) : "memory");
}
HD void cp_async16_cg(void *smem_dst, const void *gsrc) { // L2-only (coherent with a still-draining predecessor kernel)
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gsrc) : "memory");
}
HD void cp_async_wait() { asm volatile("cp.async.wait_all;" ::: "memory"); }
HD void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
HD void cp_async_wait_keep1() { asm volatile("cp.async.wait_group 1;" ::: "memory"); } // all but the newest group
// Programmatic dependent launch: a kernel may start (and stage its twiddles) while its predecessor
// drains; it must not touch the predecessor's output before grid_dep_wait().
HD void grid_dep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
HD void grid_dep_launch() {
#ifndef HEVM_NO_PDL_TRIGGER
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}
// Streaming 64-bit / 256-bit loads of data produced by EARLIER KERNELS (ciphertext limbs, scratch):
// ld.global.cg = cached in L2 only.  They must not use the non-coherent (.nc) path: with programmatic
// dependent launch a kernel becomes resident -- and its SM's L1 is invalidated -- before its
// predecessor has finished writing, so a .nc load could hit a stale L1 line of a reused scratch buffer.
HD u64 ldg_stream(const u64 *p) {
  u64 v;
  asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
HD void ldg_stream4(const u64 *p, u64 &a, u64 &b, u64 &c, u64 &d) {
  asm volatile("ld.global.cg.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p) : "memory");
}
// read-only for the whole life of the VM (key-switch keys): non-coherent path, no L1 allocation
HD void ldg_ro4(const u64 *p, u64 &a, u64 &b, u64 &c, u64 &d) {
  asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
}
HD void stg4(u64 *p, u64 a, u64 b, u64 c, u64 d) {
  asm volatile("st.global.v4.u64 [%0], {%1,%2,%3,%4};" ::"l"(p), "l"(a), "l"(b), "l"(c), "l"(d) : "memory");
}
#else
#define LANE_DECL
#define FOR_LANES(S, st, ...)                                                                                         \
  for (int lane = 0; lane < 32; lane++) {                                                                              \
    auto &S = st[lane];                                                                                                \
    __VA_ARGS__;                                                                                                       \
  }
#define NLANE_STATE 32
HD Tw ldtw(const Tw *p) { return *p; }
HD void cp_async16(void *smem_dst, const void *gsrc) { *(Tw *)smem_dst = *(const Tw *)gsrc; }
HD void cp_async16_cg(void *smem_dst, const void *gsrc) { *(Tw *)smem_dst = *(const Tw *)gsrc; }
HD void cp_async_wait() {}
HD void cp_async_commit() {}
HD void cp_async_wait_keep1() {}
HD void grid_dep_wait() {}
HD void grid_dep_launch() {}
HD u64 ldg_stream(const u64 *p) { return *p; }
HD void ldg_stream4(const u64 *p, u64 &a, u64 &b, u64 &c, u64 &d) { a = p[0], b = p[1], c = p[2], d = p[3]; }
HD void ldg_ro4(const u64 *p, u64 &a, u64 &b, u64 &c, u64 &d) { a = p[0], b = p[1], c = p[2], d = p[3]; }
HD void stg4(u64 *p, u64 a, u64 b, u64 c, u64 d) { p[0] = a, p[1] = b, p[2] = c, p[3] = d; }
#endif

HD int padx(int idx) { return idx + (idx >> 4); }
#define WARP_TILE_WORDS 544  // 512 values + 512/16 padding

// Ring geometry: N = 2^LOGA rows x 256 columns (LOGA = logN - 8; 6, 7, 8, 9 <=> N = 2^14 .. 2^17).
// Pass A tile = 2^LOGA rows x C columns with 2^LOGA * C = 512 values (16 per lane).
template <int LOGA> struct Geo {
  static constexpr int ROWS = 1 << LOGA;
  static constexpr int LOGC = 9 - LOGA;      // log2(columns per pass-A tile)
  static constexpr int C = 1 << LOGC;
  static constexpr int TILES = 256 / C;      // pass-A tiles per limb
  static constexpr int LOGT = 8 - LOGC;
  // staged twiddles per warp: pass B 255; fused inverse+forward pass A: ROWS + 2 x ROWS (the forward table is
  // double-buffered up to 256 rows; at 512 rows it is single-buffered to keep two CTAs per SM)
  static constexpr bool TW_DOUBLE = LOGA <= 8;
  static constexpr int TW = (TW_DOUBLE ? 3 : 2) * ROWS > TWB_ROW ? (TW_DOUBLE ? 3 : 2) * ROWS : TWB_ROW;
  static constexpr int WARP_WORDS = WARP_TILE_WORDS + 2 * TW;
};

// ---- twiddle staging (one cp.async batch per job) ------------------------------------------------
// pass A: entries [0,ROWS) of the prime's table (tw[m+g], m+g < ROWS)
template <int LOGA> HD void stage_tw_A(Tw *dst, const Tw *table, int lane) {
  _Pragma("unroll")
  for (int i = 0; i < Geo<LOGA>::ROWS / 32; i++) cp_async16(dst + lane + 32 * i, table + lane + 32 * i);
}
// pass B, row r: local entry twB_pos(k, g) = table[(ROWS << k) + (r << k) + g],  k < 8, g < 2^k.
// `tid`/`nthr`: the threads sharing the copy.
// `tableB` = the prime's row-wise copy (NttTables::twB / itwB): one contiguous 5 KB block per row, already in the padded
// order, ten 16-byte cp.async per lane with immediate offsets
template <int LOGA> HD void stage_tw_B(Tw *dst, const Tw *tableB, int r, int tid, int nthr = 32) {
  const Tw *src = tableB + (size_t)r * TWB_ROW;
  _Pragma("unroll")
  for (int g = tid; g < TWB_ROW; g += nthr) cp_async16(dst + g, src + g);
}

// =====================================================================================
// Pass A (strided): tile = ROWS rows x C cols (512 values), tile index = row*C + col.
//   layout R  ("row-major lanes"): x[e] <-> row = e*(ROWS/16) + (lane>>LOGC), col = lane&(C-1), idx = e*32 + lane
//   layout S  ("stage lanes"):     x[e] <-> row = (lane>>LOGC)*16 + e,        col = lane&(C-1)
// forward: stages 0-3 in layout R (top four row bits), stages 4..LOGA-1 in layout S (low row bits)
// inverse: row gaps 1..2^(LOGA-5) in layout S, the four largest gaps in layout R (last one carries N^-1)
// `tw` points at the staged table (entries 0..ROWS-1 of the prime's table).
// =====================================================================================
template <int LOGA> HD int rowR(int lane, int e) { return e * (Geo<LOGA>::ROWS / 16) + (lane >> Geo<LOGA>::LOGC); }
template <int LOGA> HD int rowS(int lane, int e) { return (lane >> Geo<LOGA>::LOGC) * 16 + e; }
template <int LOGA> HD int colA(int lane) { return lane & (Geo<LOGA>::C - 1); }
HD int idxR(int lane, int e) { return e * 32 + lane; }
template <int LOGA> HD int idxS(int lane, int e) { return rowS<LOGA>(lane, e) * Geo<LOGA>::C + colA<LOGA>(lane); }

// layout X ("middle bits", only for 512 rows x 1 column): x[e] <-> row = (lane>>1)*32 + e*2 + (lane&1), i.e. the
// register index carries row bits 4..1.  With 9 row bits the 16 registers of layouts R (bits 8..5) and S (bits 3..0)
// leave bit 4 uncovered; a third round in layout X does the stages with row gaps 16, 8, 4, 2.
HD int rowX(int lane, int e) { return ((lane >> 1) << 5) | (e << 1) | (lane & 1); }
// forward stages 4..7 of a 9-stage pass A (row gaps 16, 8, 4, 2); every stage adds 2q to the bound
template <int LOGA> HD void fwdA_stages_X(u64 (&x)[16], int lane, const Tw *tw, u64 q, u64 q2) {
  _Pragma("unroll")
  for (int s = 4; s < 8; s++) {
    const int half = 8 >> (s - 4);
    _Pragma("unroll")
    for (int e = 0; e < 16; e++)
      if (!(e & half)) {
        Tw t = ldtw(tw + (1 << s) + (rowX(lane, e) >> (LOGA - s)));
        ct_bfly_lazy(x[e], x[e + half], t, q, q2);
      }
  }
}
// inverse: the row gap 16 (j = 4) of a 9-stage pass A; in/out < 2q
template <int LOGA> HD void invA_stages_X(u64 (&x)[16], int lane, const Tw *itw, u64 q, u64 q2, u64 dl) {
  const int j = 4, half = 8;
  _Pragma("unroll")
  for (int e = 0; e < 16; e++)
    if (!(e & half)) {
      Tw t = ldtw(itw + ((Geo<LOGA>::ROWS / 2) >> j) + (rowX(lane, e) >> (j + 1)));
      gs_bfly_fold(x[e], x[e + half], t, q, q2, dl);
    }
}

// in: < 2q   out: < 10q
HD void fwdA_stages_R(u64 (&x)[16], const Tw *tw, u64 q, u64 q2) {
  _Pragma("unroll")
  for (int s = 0; s < 4; s++) {
    const int half = 8 >> s;
    Tw t[8];
    _Pragma("unroll")
    for (int g = 0; g < (1 << s); g++) t[g] = ldtw(tw + (1 << s) + g);
    _Pragma("unroll")
    for (int e = 0; e < 16; e++)
      if (!(e & half)) ct_bfly_lazy(x[e], x[e + half], t[e >> (4 - s)], q, q2);
  }
}
// stages 4..LOGA-1; every stage adds 2q to the bound
template <int LOGA> HD void fwdA_stages_S(u64 (&x)[16], int lane, const Tw *tw, u64 q, u64 q2) {
  const int rbase = (lane >> Geo<LOGA>::LOGC) * 16;
  _Pragma("unroll")
  for (int s = (LOGA > 8 ? LOGA - 1 : 4); s < LOGA; s++) { // 512 rows: stages 4..7 were done in layout X
    const int half = 1 << (LOGA - 1 - s);
    _Pragma("unroll")
    for (int e = 0; e < 16; e++)
      if (!(e & half)) {
        Tw t = ldtw(tw + (1 << s) + ((rbase + e) >> (LOGA - s)));
        ct_bfly_lazy(x[e], x[e + half], t, q, q2);
      }
  }
}
// inverse, layout S: row gaps 2^j, j < LOGA-4  (m = ROWS/2 >> j groups); in/out < 2q
template <int LOGA> HD void invA_stages_S(u64 (&x)[16], int lane, const Tw *itw, u64 q, u64 q2, u64 dl) {
  const int rbase = (lane >> Geo<LOGA>::LOGC) * 16;
  _Pragma("unroll")
  for (int j = 0; j < (LOGA - 4 < 4 ? LOGA - 4 : 4); j++) { // at most the four gaps 1, 2, 4, 8 fit the 16 registers
    const int half = 1 << j;
    _Pragma("unroll")
    for (int e = 0; e < 16; e++)
      if (!(e & half)) {
        Tw t = ldtw(itw + ((Geo<LOGA>::ROWS / 2) >> j) + ((rbase + e) >> (j + 1)));
        gs_bfly_fold(x[e], x[e + half], t, q, q2, dl);
      }
  }
}
// inverse, layout R: the four largest row gaps (m = 8,4,2,1 groups); the last stage applies N^-1; output canonical
HD void invA_stages_R(u64 (&x)[16], const Tw *itw, u64 q, u64 q2, u64 dl, Tw invn, Tw invn_w) {
  _Pragma("unroll")
  for (int jj = 0; jj < 3; jj++) {
    const int half = 1 << jj;
    _Pragma("unroll")
    for (int e = 0; e < 16; e++)
      if (!(e & half)) {
        const Tw t = ldtw(itw + (8 >> jj) + (e >> (jj + 1))); // warp-uniform address: broadcast LDS.128
        gs_bfly_fold(x[e], x[e + half], t, q, q2, dl);
      }
  }
  _Pragma("unroll")
  for (int e = 0; e < 8; e++) {
    u64 u = x[e], v = x[e + 8];
    x[e] = csub(shoup_lazy(u + v, invn, q), q);
    x[e + 8] = csub(shoup_lazy(u + q2 - v, invn_w, q), q);
  }
}

// =====================================================================================
// Pass B (contiguous): a row of 256 values, 8 per lane, row index r in [0,128).
//   layout H ("high bits in registers"): x[e] <-> idx = e*32 + lane
//   layout M ("middle bits in registers"): idx = hi*32 + e*4 + (lane&3), hi = ((lane>>2)&3)*2 + (lane>>4)
//   layout C ("consecutive"):            x[e] <-> idx = lane*8 + e
// forward: stages k=0..2 in H (distance 128,64,32), k=3..5 in M (16,8,4), k=6,7 in C (2,1)
// inverse: gaps 1,2 in C; 4,8,16 in M; 32,64,128 in H
// `tw` points at the staged row table: entry twB_pos(k, g) = twiddle of local stage k, local group g.
// =====================================================================================
HD int idxH(int lane, int e) { return e * 32 + lane; }
HD int idxM8(int lane, int e) { return ((((lane >> 2) & 3) * 2 + (lane >> 4)) << 5) + e * 4 + (lane & 3); }
HD int idxC8(int lane, int e) { return lane * 8 + e; }

HD void fwdB8_stages_H(u64 (&x)[8], const Tw *tw, u64 q, u64 q2) { // in < 2q, out < 8q
  _Pragma("unroll")
  for (int k = 0; k < 3; k++) {
    const int half = 4 >> k;
    Tw t[4];
    _Pragma("unroll")
    for (int g = 0; g < (1 << k); g++) t[g] = ldtw(tw + twB_pos(k, g));
    _Pragma("unroll")
    for (int e = 0; e < 8; e++)
      if (!(e & half)) ct_bfly_lazy(x[e], x[e + half], t[e >> (3 - k)], q, q2);
  }
}
HD void fwdB8_stages_M(u64 (&x)[8], int lane, const Tw *tw, u64 q, u64 q2) { // in < 2q, out < 8q
  const int base = idxM8(lane, 0);
  _Pragma("unroll")
  for (int k = 3; k < 6; k++) {
    const int half = 4 >> (k - 3);
    _Pragma("unroll")
    for (int e = 0; e < 8; e++)
      if (!(e & half)) {
        Tw t = ldtw(tw + twB_pos(k, (base + e * 4) >> (8 - k)));
        ct_bfly_lazy(x[e], x[e + half], t, q, q2);
      }
  }
}
HD void fwdB8_stages_C(u64 (&x)[8], int lane, const Tw *tw, u64 q, u64 q2) { // in < 2q, out < 6q
  const int base = lane * 8;
  _Pragma("unroll")
  for (int k = 6; k < 8; k++) {
    const int half = 2 >> (k - 6);
    _Pragma("unroll")
    for (int e = 0; e < 8; e++)
      if (!(e & half)) {
        Tw t = ldtw(tw + twB_pos(k, (base + e) >> (8 - k)));
        ct_bfly_lazy(x[e], x[e + half], t, q, q2);
      }
  }
}
// inverse: gap 2^j <-> 128>>j groups = local stage k = 7 - j, staged entry twB_pos(k, group); in/out < 2q
HD void invB8_stages_C(u64 (&x)[8], int lane, const Tw *itw, u64 q, u64 q2, u64 dl) {
  const int base = lane * 8;
  _Pragma("unroll")
  for (int j = 0; j < 2; j++) {
    const int half = 1 << j;
    _Pragma("unroll")
    for (int e = 0; e < 8; e++)
      if (!(e & half)) {
        Tw t = ldtw(itw + twB_pos(7 - j, (base + e) >> (j + 1)));
        gs_bfly_fold(x[e], x[e + half], t, q, q2, dl);
      }
  }
}
HD void invB8_stages_M(u64 (&x)[8], int lane, const Tw *itw, u64 q, u64 q2, u64 dl) {
  const int base = idxM8(lane, 0);
  _Pragma("unroll")
  for (int j = 2; j < 5; j++) {
    const int half = 1 << (j - 2);
    _Pragma("unroll")
    for (int e = 0; e < 8; e++)
      if (!(e & half)) {
        Tw t = ldtw(itw + twB_pos(7 - j, (base + e * 4) >> (j + 1)));
        gs_bfly_fold(x[e], x[e + half], t, q, q2, dl);
      }
  }
}
HD void invB8_stages_H(u64 (&x)[8], const Tw *itw, u64 q, u64 q2, u64 dl) {
  _Pragma("unroll")
  for (int j = 5; j < 8; j++) {
    const int half = 1 << (j - 5), ml = 128 >> j;
    Tw t[4];
    _Pragma("unroll")
    for (int g = 0; g < ml; g++) t[g] = ldtw(itw + twB_pos(7 - j, g));
    _Pragma("unroll")
    for (int e = 0; e < 8; e++)
      if (!(e & half)) gs_bfly_fold(x[e], x[e + half], t[e >> (j - 4)], q, q2, dl);
  }
}

// ---- lane state carried across FOR_LANES phases ---------------------------------------
struct LaneA {
  u64 x[16];
  u64 y[16];
};
struct LaneB8 {
  u64 x[8];
  u64 z[8]; // second result (ct x ct mod-down handles both output polys in one job)
};

// =====================================================================================
// Warp-level pass bodies.  `sm` = warp-private tile (WARP_TILE_WORDS words); `tw` = staged twiddles.
// =====================================================================================

// forward pass A on values already held in layout R by st[].y; result (lazy, < 16q) -> dst tile
// (ROWS x C tile at column c0 of the limb `dst`, row pitch = 256)
// Range: every stage adds at most 2q to the bound, so inputs < 2q leave the 7 stages < 16q < 2^64
// without any correction (FOLD = false); inputs up to 4q need one fold between the rounds (FOLD = true).
template <int LOGA, bool FOLD>
HD void warp_fwdA_from_regs(LaneA *st, u64 *sm, u64 *dst, int c0, const Tw *tw, const ModQ &m) {
  const u64 q = m.q, q2 = 2 * m.q, dl = m.delta;
  LANE_DECL;
  FOR_LANES(S, st, {
    fwdA_stages_R(S.y, tw, q, q2);
    _Pragma("unroll")
    for (int e = 0; e < 16; e++) sm[padx(idxR(lane, e))] = FOLD ? fold60(S.y[e], dl) : S.y[e];
  });
  if (LOGA > 8) { // 512 rows: middle round in layout X (tile index = row, one column)
    FOR_LANES(S, st, {
      _Pragma("unroll")
      for (int e = 0; e < 16; e++) S.y[e] = sm[padx(rowX(lane, e))];
      fwdA_stages_X<LOGA>(S.y, lane, tw, q, q2);
      _Pragma("unroll")
      for (int e = 0; e < 16; e++) sm[padx(rowX(lane, e))] = S.y[e];
    });
  }
  FOR_LANES(S, st, {
    u64 *dstS = dst + ((size_t)rowS<LOGA>(lane, 0) << 8) + c0 + colA<LOGA>(lane);
    _Pragma("unroll")
    for (int e = 0; e < 16; e++) S.y[e] = sm[padx(idxS<LOGA>(lane, e))];
    fwdA_stages_S<LOGA>(S.y, lane, tw, q, q2);
    _Pragma("unroll")
    for (int e = 0; e < 16; e++) dstS[e << 8] = S.y[e]; // row rowS(lane, e): one base + compile-time offsets
  });
}

// inverse pass A: src tile (layout S load, values < 2q) -> canonical coefficients in st[].x, layout R
template <int LOGA>
HD void warp_invA_to_regs(LaneA *st, u64 *sm, const u64 *src, int c0, const Tw *itw, const ModQ &m, Tw invn, Tw invn_w) {
  const u64 q = m.q, q2 = 2 * m.q, dl = m.delta;
  LANE_DECL;
  FOR_LANES(S, st, {
    const u64 *srcS = src + ((size_t)rowS<LOGA>(lane, 0) << 8) + c0 + colA<LOGA>(lane);
    _Pragma("unroll")
    for (int e = 0; e < 16; e++) S.x[e] = ldg_stream(srcS + (e << 8)); // row rowS(lane, e)
    invA_stages_S<LOGA>(S.x, lane, itw, q, q2, dl);
    _Pragma("unroll")
    for (int e = 0; e < 16; e++) sm[padx(idxS<LOGA>(lane, e))] = S.x[e];
  });
  if (LOGA > 8) {
    FOR_LANES(S, st, {
      _Pragma("unroll")
      for (int e = 0; e < 16; e++) S.x[e] = sm[padx(rowX(lane, e))];
      invA_stages_X<LOGA>(S.x, lane, itw, q, q2, dl);
      _Pragma("unroll")
      for (int e = 0; e < 16; e++) sm[padx(rowX(lane, e))] = S.x[e];
    });
  }
  FOR_LANES(S, st, {
    _Pragma("unroll")
    for (int e = 0; e < 16; e++) S.x[e] = sm[padx(idxR(lane, e))];
    invA_stages_R(S.x, itw, q, q2, dl, invn, invn_w);
  });
}

// forward pass B: values in st[].x layout H (any lazy bound) -> st[].x layout C (< 12q):
// fold on load (< 2q), H stages (< 8q), fold (< 2q), M stages (< 8q), C stages (< 12q)
HD void warp_fwdB8_regs(LaneB8 *st, u64 *sm, const Tw *tw, const ModQ &m) {
  const u64 q = m.q, q2 = 2 * m.q, dl = m.delta;
  LANE_DECL;
  FOR_LANES(S, st, {
    _Pragma("unroll")
    for (int e = 0; e < 8; e++) S.x[e] = fold60(S.x[e], dl);
    fwdB8_stages_H(S.x, tw, q, q2);
    _Pragma("unroll")
    for (int e = 0; e < 8; e++) sm[padx(idxH(lane, e))] = fold60(S.x[e], dl);
  });
  FOR_LANES(S, st, {
    _Pragma("unroll")
    for (int e = 0; e < 8; e++) S.x[e] = sm[padx(idxM8(lane, e))];
    fwdB8_stages_M(S.x, lane, tw, q, q2);
    _Pragma("unroll")
    for (int e = 0; e < 8; e++) sm[padx(idxM8(lane, e))] = S.x[e];
  });
  FOR_LANES(S, st, {
    _Pragma("unroll")
    for (int e = 0; e < 8; e++) S.x[e] = sm[padx(idxC8(lane, e))];
    fwdB8_stages_C(S.x, lane, tw, q, q2);
  });
}
// inverse pass B: st[].x layout C (values < 2q) -> st[].x layout H (< 2q)
HD void warp_invB8_regs(LaneB8 *st, u64 *sm, const Tw *itw, const ModQ &m) {
  const u64 q = m.q, q2 = 2 * m.q, dl = m.delta;
  LANE_DECL;
  FOR_LANES(S, st, {
    invB8_stages_C(S.x, lane, itw, q, q2, dl);
    _Pragma("unroll")
    for (int e = 0; e < 8; e++) sm[padx(idxC8(lane, e))] = S.x[e];
  });
  FOR_LANES(S, st, {
    _Pragma("unroll")
    for (int e = 0; e < 8; e++) S.x[e] = sm[padx(idxM8(lane, e))];
    invB8_stages_M(S.x, lane, itw, q, q2, dl);
    _Pragma("unroll")
    for (int e = 0; e < 8; e++) sm[padx(idxM8(lane, e))] = S.x[e];
  });
  FOR_LANES(S, st, {
    _Pragma("unroll")
    for (int e = 0; e < 8; e++) S.x[e] = sm[padx(idxH(lane, e))];
    invB8_stages_H(S.x, itw, q, q2, dl);
  });
}

// Galois automorphism in the bit-reversed NTT domain (SEAL GaloisTool::apply_galois_ntt):
// out[i] = in[ bitrev( ((elt * (2*bitrev(i)+1)) >> 1) & (N-1) ) ]
HD u32 brev32(u32 v) {
#if defined(__CUDA_ARCH__)
  return __brev(v);
#else
  v = ((v >> 1) & 0x55555555u) | ((v & 0x55555555u) << 1);
  v = ((v >> 2) & 0x33333333u) | ((v & 0x33333333u) << 2);
  v = ((v >> 4) & 0x0F0F0F0Fu) | ((v & 0x0F0F0F0Fu) << 4);
  v = ((v >> 8) & 0x00FF00FFu) | ((v & 0x00FF00FFu) << 8);
  return (v >> 16) | (v << 16);
#endif
}
HD u32 galois_src_index(u32 i, u32 elt, int logN) {
  u32 rev = brev32(i) >> (32 - logN);
  u32 raw = (u32)((((u64)elt * (2 * (u64)rev + 1)) >> 1) & ((1u << logN) - 1));
  return brev32(raw) >> (32 - logN);
}
