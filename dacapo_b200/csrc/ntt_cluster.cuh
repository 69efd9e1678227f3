// Single-pass forward NTT of one 2^15-point limb inside an 8-CTA thread-block cluster (round-2 experiment, VERDICT r1
// item 4): the limb is read from HBM ONCE (TMA 2-D tile loads of the strided 128 x 4 pass-A tiles, cp.async.bulk.tensor
// + mbarrier), pass A runs in registers, its output is scattered to the CTA that owns each row through DISTRIBUTED SHARED
// MEMORY (st.shared::cluster via cluster.map_shared_rank), one cluster barrier, pass B runs from local shared memory and the
// canonical result is written ONCE.  No L2 / HBM round trip between the passes (the two-launch path writes and re-reads the
// limb).  Same butterflies as k_fwd_A<PRE_NONE> + k_fwd_B<EPI_CANON> (SEAL ntt_negacyclic_harvey), so the output is
// bit-identical; selected with hevmx_ntt(..., inverse = 2) / hevmx_ntt_bench(..., inverse = 2).  Device-only code.
#pragma once
#include "ntt_bodies.cuh"
#include <cooperative_groups.h>
#include <cuda.h>

#define NC_CLUSTER 8                      // CTAs per limb
#define NC_WARPS 4                        // warps per CTA: 2 pass-A tiles and 4 pass-B rows per warp
#define NC_ROWS_PER_CTA 16                // 128 rows / 8 CTAs
#define NC_ROWPITCH 264                   // words per staged row (256 + 8: rows of one pass-A tile land in different banks)
#define NC_TILE_WORDS 512                 // one dense TMA tile: 128 rows x 4 columns
// shared memory (words): rows[16][264] | per warp: TMA box 512 = exchange tile 544 (aliased: the box is in registers before
// the exchange tile is written, the next box is requested after it was read) | twiddles 2 x 256 | mbarrier 2
#define NC_TW_WORDS (2 * TWB_ROW)          // staged (padded) pass-B twiddle row; the pass-A table (128 entries) uses the same words
#define NC_WARP_WORDS (WARP_TILE_WORDS + NC_TW_WORDS + 16) // multiple of 16 words: every TMA box stays 128-byte aligned
#define NC_SMEM_WORDS (NC_ROWS_PER_CTA * NC_ROWPITCH + NC_WARPS * NC_WARP_WORDS)

__device__ __forceinline__ void nc_mbar_init(u64 *bar) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void nc_mbar_expect_tx(u64 *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void nc_mbar_wait(u64 *bar, unsigned parity) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
               "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
               "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(a), "r"(parity) : "memory");
}
// 3-D tensor (columns, rows, limbs) -> dense shared-memory box of 4 columns x 128 rows
__device__ __forceinline__ void nc_tma_load_tile(u64 *dst, const CUtensorMap *map, int c0, int limb, u64 *bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                   (unsigned)__cvta_generic_to_shared(dst)),
               "l"(map), "r"(c0), "r"(0), "r"(limb), "r"((unsigned)__cvta_generic_to_shared(bar))
               : "memory");
}

template <int LOGA>
__global__ void __cluster_dims__(NC_CLUSTER, 1, 1) __launch_bounds__(NC_WARPS * 32)
    k_ntt_fwd_cluster(const __grid_constant__ CUtensorMap tmap, const NttTables *T, u64 *dst, int prime0, int pstep) {
  static_assert(LOGA == 7, "the cluster NTT is instantiated for N = 2^15 (128 rows x 256 columns)");
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ __align__(128) u64 nc_smem[];
  const int crank = (int)cluster.block_rank(), limb = blockIdx.x / NC_CLUSTER;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int N = 1 << (LOGA + 8);
  const int p = prime0 + limb * pstep;
  const ModQ m = T->mod[p];
  const u64 q = m.q, q2 = 2 * m.q, dl = m.delta;
  u64 *rows = nc_smem;
  u64 *wbase = nc_smem + NC_ROWS_PER_CTA * NC_ROWPITCH + warp * NC_WARP_WORDS;
  u64 *tile = wbase, *xch = wbase;
  Tw *tw = reinterpret_cast<Tw *>(xch + WARP_TILE_WORDS);
  u64 *bar = wbase + WARP_TILE_WORDS + NC_TW_WORDS;
  // ---- pass A: tiles crank*8 + warp*2 + k ------------------------------------------------------------------------
  if (lane == 0) {
    nc_mbar_init(bar);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  stage_tw_A<LOGA>(tw, T->tw + (size_t)p * N, lane); // 128 pass-A twiddles of this prime (one cp.async batch)
  cluster.sync(); // distributed shared memory may only be written once every CTA of the cluster is known to be running
  for (int k = 0; k < 2; k++) {
    const int c0 = (crank * 8 + warp * 2 + k) * Geo<LOGA>::C;
    if (lane == 0) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // generic-proxy accesses of the aliased tile before the TMA write
      nc_mbar_expect_tx(bar, NC_TILE_WORDS * 8);
      nc_tma_load_tile(tile, &tmap, c0, limb, bar);
    }
    nc_mbar_wait(bar, k & 1);
    u64 y[16];
#pragma unroll
    for (int e = 0; e < 16; e++) y[e] = tile[idxR(lane, e)]; // dense box [row][4 columns]: row*4 + col == idxR for layout R
    __syncwarp();                                            // the box may be overwritten by the next TMA from here on
    if (k == 0) cp_async_wait();
    __syncwarp();
    fwdA_stages_R(y, tw, q, q2);
#pragma unroll
    for (int e = 0; e < 16; e++) xch[padx(idxR(lane, e))] = fold60(y[e], dl);
    __syncwarp();
#pragma unroll
    for (int e = 0; e < 16; e++) y[e] = xch[padx(idxS<LOGA>(lane, e))];
    __syncwarp(); // every lane has drained the exchange tile: the next TMA box may land in it
    fwdA_stages_S<LOGA>(y, lane, tw, q, q2);
    // scatter: row rowS(lane, e) lives in CTA row / 16 of the cluster, at [row % 16][c0 + col]
    const int col = c0 + colA<LOGA>(lane);
#pragma unroll
    for (int e = 0; e < 16; e++) {
      const int row = rowS<LOGA>(lane, e); // (lane >> 2) * 16 + e  ->  owner CTA = lane >> 2
      u64 *remote = cluster.map_shared_rank(rows, row >> 4);
      remote[(row & 15) * NC_ROWPITCH + col] = y[e];
    }
  }
  cluster.sync(); // every row of this CTA has arrived (release / acquire at cluster scope)
  // ---- pass B: rows crank*16 + warp*4 + k ------------------------------------------------------------------------
  LaneB8 st[1];
  for (int k = 0; k < 4; k++) {
    const int lrow = warp * 4 + k, r = crank * NC_ROWS_PER_CTA + lrow;
    stage_tw_B<LOGA>(tw, T->twB + (size_t)p * twB_prime_stride(N), r, lane);
#pragma unroll
    for (int e = 0; e < 8; e++) st[0].x[e] = rows[lrow * NC_ROWPITCH + idxH(lane, e)];
    cp_async_wait();
    __syncwarp();
    warp_fwdB8_regs(st, xch, tw, m);
    u64 v[8];
#pragma unroll
    for (int e = 0; e < 8; e++) v[e] = canon60(st[0].x[e], q, dl);
    store8(dst + (size_t)limb * N + (size_t)r * 256 + lane * 8, v);
    __syncwarp(); // the row's twiddles are restaged next
  }
}
