// sm_100a kernels of libB200_HEVM.so.
//
//  * NTT-based warp-job kernels: thin __global__ wrappers around ntt_bodies.cuh.  A CTA is
//    4 independent warps (128 threads); each warp owns one job and a private padded
//    shared-memory tile; only the key inner-product kernel has CTA barriers (its phase boundaries).
//  * element-wise RNS kernels (256-bit vector loads/stores, one HBM pass)
//  * samplers, key-generation / encryption helpers
//  * fp64 CKKS encoder / decoder kernels (radix-2 special FFT, four stages per launch, explicit *_rn
//    intrinsics so that no FMA contraction changes bits w.r.t. the host restatement)
#include "kernels.h"
#include "ntt_cluster.cuh"
#include <algorithm>

unsigned long long g_launch_count = 0;
bool g_pdl_suspended = false;
KProfiler g_prof;
const char *const g_kernel_class_names[KC_COUNT] = {
    "intt_B_plain", "intt_B_galois", "intt_B_product", "intt_A", "fwd_A_plain", "fwd_A_modup", "fwd_A_round",
    "fwd_B_canon", "fwd_B_mac", "fwd_B_moddown_galois", "fwd_B_moddown_relin", "fwd_B_rescale", "invA_fwdA_modup",
    "invA_fwdA_round", "elementwise", "ks_fused_galois", "ks_fused_relin", "rescale_fused", "other"};
cudaEvent_t KProfiler::ev() {
  if (used == pool.size()) {
    cudaEvent_t e;
    CUDA_CHECK(cudaEventCreate(&e));
    pool.push_back(e);
  }
  return pool[used++];
}
void KProfiler::begin(cudaStream_t s, int cls) {
  if (!on) return;
  Rec r{cls, ev(), ev()};
  CUDA_CHECK(cudaEventRecord(r.a, s));
  recs.push_back(r);
  pending = true;
}
void KProfiler::end(cudaStream_t s) {
  if (!on || !pending) return;
  pending = false;
  CUDA_CHECK(cudaEventRecord(recs.back().b, s));
}
void KProfiler::collect() {
  for (auto &r : recs) {
    CUDA_CHECK(cudaEventSynchronize(r.b));
    float t = 0;
    CUDA_CHECK(cudaEventElapsedTime(&t, r.a, r.b));
    ms[r.cls] += t;
    cnt[r.cls]++;
  }
  recs.clear();
  used = 0;
}
void KProfiler::reset() {
  collect();
  for (int i = 0; i < KC_COUNT; i++) ms[i] = 0, cnt[i] = 0;
}

#define WARPS_PER_CTA 4
#define CTA_THREADS (WARPS_PER_CTA * 32)

// Every NTT kernel: a CTA is WARPS_PER_CTA independent warps; warp = one job; per-warp shared region
// (dynamic shared memory) = padded value tile + staged twiddles.  No __syncthreads (except the MAC
// kernel's CTA barriers).
extern __shared__ __align__(16) u64 dyn_smem[];
#define WARP_KERNEL_PROLOGUE(LaneT)                                                                                    \
  const int warp = threadIdx.x >> 5, job = blockIdx.x * WARPS_PER_CTA + warp;                                          \
  if (job >= njobs) return;                                                                                            \
  u64 *sm = dyn_smem + (size_t)warp * Geo<LOGA>::WARP_WORDS;                                                           \
  LaneT st[1];

template <int LOGA, int LD> __global__ void __launch_bounds__(CTA_THREADS) k_intt_B(ArgsInttB a, int njobs) {
  WARP_KERNEL_PROLOGUE(LaneB8)
  body_intt_B<LOGA, LD>(a, job, st, sm);
}
template <int LOGA> __global__ void __launch_bounds__(CTA_THREADS) k_intt_A(ArgsInttA a, int njobs) {
  WARP_KERNEL_PROLOGUE(LaneA)
  body_intt_A<LOGA>(a, job, st, sm);
}
template <int LOGA, int PRE> __global__ void __launch_bounds__(CTA_THREADS) k_fwd_A(ArgsFwdA a, int njobs) {
  WARP_KERNEL_PROLOGUE(LaneA)
  body_fwd_A<LOGA, PRE>(a, job, st, sm);
}
template <int LOGA, int EPI> __global__ void __launch_bounds__(CTA_THREADS) k_fwd_B(ArgsFwdB a, int njobs) {
  WARP_KERNEL_PROLOGUE(LaneB8)
  body_fwd_B<LOGA, EPI>(a, job, st, sm);
}
#ifndef FUSEA_MIN_CTAS
#define FUSEA_MIN_CTAS 4
#endif
template <int LOGA, int PRE> __global__ void __launch_bounds__(CTA_THREADS, (LOGA >= 8 ? 3 : FUSEA_MIN_CTAS)) k_invA_fwdA(ArgsInvFwdA a, int njobs) {
  WARP_KERNEL_PROLOGUE(LaneA)
  body_invA_fwdA<LOGA, PRE>(a, job, st, sm);
}
// key-switch inner product: CTA = one (output prime, row) job, its MAC_WARPS warps split the digits
#ifndef MAC_MIN_CTAS
#define MAC_MIN_CTAS 4
#endif
template <int LOGA> __global__ void __launch_bounds__(MAC_WARPS * 32, MAC_MIN_CTAS) k_mac(ArgsFwdB a) {
  u64 *sm = dyn_smem;
  Tw *tw_s = reinterpret_cast<Tw *>(sm);
  u64 *tiles = sm + MAC_TW_WORDS;
  u64 *rowbufs = tiles + MAC_WARPS * TILE_B_WORDS;
  u64 *xbuf = rowbufs + MAC_WARPS * MAC_ROW_WORDS;
  const int job = blockIdx.x, warp = threadIdx.x >> 5;
  body_mac_stage<LOGA>(a, job, threadIdx.x, tw_s);
  __syncthreads();
  LaneB8 st[1];
  body_mac_warp<LOGA>(a, job, warp, st, tiles + warp * TILE_B_WORDS, tw_s, rowbufs + warp * MAC_ROW_WORDS, xbuf);
  __syncthreads();
  body_mac_dot<LOGA>(a, job, threadIdx.x, xbuf, tiles);
  if (mac_Iidx<LOGA>(a, job) == a.l) { // special prime: continue with the inverse pass B of the two accumulator rows
    __syncthreads();
    // the two tail warps each stage the row's inverse twiddles into a table of their own (warp 1: the CTA's table, warp 0:
    // the free words behind the two accumulator rows kept in `tiles`), so nothing is written or read across warps
    if (warp < 2) body_mac_tail<LOGA>(a, job, warp, st, mac_tail_tile(rowbufs, warp), mac_tail_tw(tiles, tw_s, warp), tiles);
  }
}


// ---- single-launch key switch / rescale (ks_fused.cuh): persistent CTAs, ticket-ordered work units, per-limb
// completion counters instead of kernel boundaries -----------------------------------------------------------------
__device__ __forceinline__ unsigned ld_relaxed_u32(const unsigned *p) {
  unsigned v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned *p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
#ifndef FUSED_POLL_ALL
#define FUSED_POLL_ALL 1
#endif
#ifndef FUSED_POLL_NS
#define FUSED_POLL_NS 32
#endif
template <int LOGA, int MODE> __global__ void __launch_bounds__(CTA_THREADS, 4) k_ks_fused(KsFusedArgs A) {
  __shared__ int s_ticket, s_last;
  const int L = A.Ltot;
  const size_t N = (size_t)1 << (LOGA + 8);
  grid_dep_wait(); // the previous kernel on this stream has zeroed the counters / produced the operands
  grid_dep_launch();
  const KsUnits K = ks_units<LOGA, MODE>(A);
  unsigned *const ctl0 = reinterpret_cast<unsigned *>(A.ct[0].scratch + Scratch::words(L, N) - KS_CTL_WORDS); // launch-wide ticket / exit counters
  if (threadIdx.x == 0) s_ticket = (int)atomicAdd(ctl0 + ctl_ticket(L), 1u);
  for (;;) {
    __syncthreads();
    const int ticket = s_ticket;
    if (ticket >= K.total) break;
    const int warp = threadIdx.x >> 5;
    {
      const KsUnit un = ks_decode(K, A.nct, ticket);
      const KsDeps d = ks_deps<LOGA, MODE>(A, un);
#if FUSED_POLL_ALL
      if (d.nwait) { // every thread polls with an acquire load
        const unsigned *ctl = reinterpret_cast<const unsigned *>(A.ct[un.c].scratch + Scratch::words(L, N) - KS_CTL_WORDS);
#pragma unroll
        for (int k = 0; k < 3; k++)
          if (k < d.nwait)
            while (ld_acquire_u32(ctl + d.widx[k]) < d.wtarget[k]) __nanosleep(FUSED_POLL_NS);
      }
#else
      if ((threadIdx.x & 31) == 0 && d.nwait) { // lane 0 of every warp polls (relaxed), then one acquire fence
        const unsigned *ctl = reinterpret_cast<const unsigned *>(A.ct[un.c].scratch + Scratch::words(L, N) - KS_CTL_WORDS);
#pragma unroll
        for (int k = 0; k < 3; k++)
          if (k < d.nwait)
            while (ld_relaxed_u32(ctl + d.widx[k]) < d.wtarget[k]) __nanosleep(FUSED_POLL_NS);
        asm volatile("fence.acq_rel.gpu;" ::: "memory");
      }
#endif
      __syncwarp();
    }
    {
      const KsUnit un = ks_decode(K, A.nct, ticket);
      const KsCt &c = A.ct[un.c];
      Scratch sc;
      sc.carve(c.scratch, L, N);
      u64 *sm = dyn_smem + (size_t)warp * Geo<LOGA>::WARP_WORDS;
      const int job = un.u * WARPS_PER_CTA + warp;
      switch (un.phase) {
      case 0: {
        LaneB8 st[1];
        body_intt_B<LOGA, MODE == FUSED_RESCALE ? LD_PLAIN : MODE>(ks_args_p1<MODE>(A, c, sc), job, st, sm);
        break;
      }
      case 1: {
        LaneA st[1];
        body_invA_fwdA<LOGA, PRE_MODUP>(ks_args_p2(A, sc), job, st, sm);
        break;
      }
      case 2: {
        const ArgsFwdB a = ks_args_p3<MODE>(A, c, sc);
        Tw *tw_s = reinterpret_cast<Tw *>(dyn_smem);
        u64 *tiles = dyn_smem + MAC_TW_WORDS, *rowbufs = tiles + MAC_WARPS * TILE_B_WORDS, *xbuf = rowbufs + MAC_WARPS * MAC_ROW_WORDS;
        body_mac_stage<LOGA>(a, un.u, threadIdx.x, tw_s);
        __syncthreads();
        LaneB8 st[1];
        body_mac_warp<LOGA>(a, un.u, warp, st, tiles + warp * TILE_B_WORDS, tw_s, rowbufs + warp * MAC_ROW_WORDS, xbuf);
        __syncthreads();
        body_mac_dot<LOGA>(a, un.u, threadIdx.x, xbuf, tiles);
        if (mac_Iidx<LOGA>(a, un.u) == a.l) {
          __syncthreads();
          if (warp < 2) body_mac_tail<LOGA>(a, un.u, warp, st, mac_tail_tile(rowbufs, warp), mac_tail_tw(tiles, tw_s, warp), tiles);
        }
        break;
      }
      case 3: {
        LaneA st[1];
        body_invA_fwdA<LOGA, PRE_ROUND>(ks_args_p4<MODE>(A, sc), job, st, sm);
        break;
      }
      default: {
        LaneB8 st[1];
        body_fwd_B<LOGA, MODE == LD_GALOIS ? EPI_MODDOWN_GALOIS : MODE == LD_PRODUCT ? EPI_MODDOWN_RELIN : EPI_RESCALE>(ks_args_p5<MODE>(A, c, sc), job, st, sm);
        break;
      }
      }
    }
    __syncthreads(); // every warp of the unit has issued its stores (and read s_ticket)
    // (the next ticket is NOT requested ahead of time: a held ticket would delay a unit that idle CTAs could start now)
    if (threadIdx.x == 0) s_ticket = (int)atomicAdd(ctl0 + ctl_ticket(L), 1u);
    if (threadIdx.x < 32) { // signals (recomputed rather than kept live across the unit body)
      const KsUnit un = ks_decode(K, A.nct, ticket);
      const KsDeps d = ks_deps<LOGA, MODE>(A, un);
      if ((int)threadIdx.x < d.sig_count && (int)threadIdx.x != d.sig_skip) {
        unsigned *ctl = reinterpret_cast<unsigned *>(A.ct[un.c].scratch + Scratch::words(L, N) - KS_CTL_WORDS);
        __threadfence();
        atomicAdd(ctl + d.sig_first + threadIdx.x, d.sig_inc);
      }
    }
  }
  // the last CTA to leave zeroes every counter block for the next launch on these scratch areas
  if (threadIdx.x == 0) s_last = atomicAdd(ctl0 + ctl_nexit(L), 1u) == gridDim.x - 1;
  __syncthreads();
  if (s_last)
    for (int cc = 0; cc < A.nct; cc++) {
      unsigned *ctl = reinterpret_cast<unsigned *>(A.ct[cc].scratch + Scratch::words(L, N) - KS_CTL_WORDS);
      for (int i = threadIdx.x; i < ctl_count(L); i += CTA_THREADS) ctl[i] = 0;
    }
}

static inline int ctas_for(int njobs) { return (njobs + WARPS_PER_CTA - 1) / WARPS_PER_CTA; }
// launch with programmatic stream serialization: the kernel may begin (twiddle staging) while its
// predecessor drains; every such kernel executes griddepcontrol.wait before touching global data
template <class... KArgs, class... Args>
static void launch_pdl(void (*kern)(KArgs...), int grid, int block, size_t smem, cudaStream_t s, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid), cfg.blockDim = dim3(block), cfg.dynamicSmemBytes = smem, cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  static const bool no_pdl = std::getenv("HEVM_PDL") && std::atoi(std::getenv("HEVM_PDL")) == 0;
  cfg.attrs = at, cfg.numAttrs = (no_pdl || g_pdl_suspended) ? 0 : 1;
  CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, args...));
}
// dynamic shared memory of the warp kernels (may exceed the 48 KB default: opt in once per kernel)
template <int LOGA, class K> static size_t warp_smem(K kern) {
  const size_t bytes = (size_t)WARPS_PER_CTA * Geo<LOGA>::WARP_WORDS * sizeof(u64);
  static bool done = false; // one instance per (LOGA, K) instantiation... K is a function-pointer TYPE: guard per pointer below
  static const void *seen[16];
  static int nseen = 0;
  (void)done;
  for (int i = 0; i < nseen; i++)
    if (seen[i] == (const void *)kern) return bytes;
  CUDA_CHECK(cudaFuncSetAttribute((const void *)kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  if (nseen < 16) seen[nseen++] = (const void *)kern;
  return bytes;
}
#define PRE_LAUNCH(s, cls) g_prof.begin(s, cls)
#define POST_LAUNCH_S(s)                                                                                               \
  do {                                                                                                                 \
    g_prof.end(s);                                                                                                     \
    g_launch_count++;                                                                                                  \
    CUDA_CHECK(cudaGetLastError());                                                                                    \
  } while (0)

template <int LOGA, int LD> void GpuLauncher::intt_B(const ArgsInttB &a, int njobs) {
  if (njobs <= 0) return;
  PRE_LAUNCH(stream, LD == LD_DECRYPT ? KC_INTT_B_PLAIN : KC_INTT_B_PLAIN + LD);
  launch_pdl(k_intt_B<LOGA, LD>, ctas_for(njobs), CTA_THREADS, warp_smem<LOGA>(k_intt_B<LOGA, LD>), stream, a, njobs);
  POST_LAUNCH_S(stream);
}
template <int LOGA> void GpuLauncher::intt_A(const ArgsInttA &a, int njobs) {
  if (njobs <= 0) return;
  PRE_LAUNCH(stream, KC_INTT_A);
  launch_pdl(k_intt_A<LOGA>, ctas_for(njobs), CTA_THREADS, warp_smem<LOGA>(k_intt_A<LOGA>), stream, a, njobs);
  POST_LAUNCH_S(stream);
}
template <int LOGA, int PRE> void GpuLauncher::fwd_A(const ArgsFwdA &a, int njobs) {
  if (njobs <= 0) return;
  PRE_LAUNCH(stream, KC_FWD_A_NONE + PRE);
  launch_pdl(k_fwd_A<LOGA, PRE>, ctas_for(njobs), CTA_THREADS, warp_smem<LOGA>(k_fwd_A<LOGA, PRE>), stream, a, njobs);
  POST_LAUNCH_S(stream);
}
template <int LOGA, int EPI> void GpuLauncher::fwd_B(const ArgsFwdB &a, int njobs) {
  if (njobs <= 0) return;
  PRE_LAUNCH(stream, KC_FWD_B_CANON + EPI);
  launch_pdl(k_fwd_B<LOGA, EPI>, ctas_for(njobs), CTA_THREADS, warp_smem<LOGA>(k_fwd_B<LOGA, EPI>), stream, a, njobs);
  POST_LAUNCH_S(stream);
}
template <int LOGA> void GpuLauncher::mac(const ArgsFwdB &a, int njobs) {
  if (njobs <= 0) return;
  PRE_LAUNCH(stream, KC_FWD_B_MAC);
  static bool optin = false; // per LOGA instantiation
  if (!optin) {
    CUDA_CHECK(cudaFuncSetAttribute((const void *)k_mac<LOGA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(mac_smem_words(HEVM_MAXL) * sizeof(u64))));
    optin = true;
  }
  launch_pdl(k_mac<LOGA>, njobs, MAC_WARPS * 32, mac_smem_words(a.l) * sizeof(u64), stream, a);
  POST_LAUNCH_S(stream);
}
template <int LOGA, int PRE> void GpuLauncher::invA_fwdA(const ArgsInvFwdA &a, int njobs) {
  if (njobs <= 0) return;
  PRE_LAUNCH(stream, PRE == PRE_MODUP ? KC_INVA_FWDA_MODUP : KC_INVA_FWDA_ROUND);
  launch_pdl(k_invA_fwdA<LOGA, PRE>, ctas_for(njobs), CTA_THREADS, warp_smem<LOGA>(k_invA_fwdA<LOGA, PRE>), stream, a, njobs);
  POST_LAUNCH_S(stream);
}

// ---- single-pass cluster NTT (ntt_cluster.cuh): TMA tensor map over the source limbs, 8 CTAs per limb ----
void launch_ntt_fwd_cluster(cudaStream_t s, const NttTables *T, int logN, const u64 *src, u64 *dst, int nl, int prime0, int pstep) {
  if (logN != 15) {
    std::fprintf(stderr, "[b200-hevm] fatal: the cluster NTT is built for N = 2^15\n");
    std::abort();
  }
  typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                               const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn encode = nullptr;
  if (!encode) {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qr;
    CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr));
    if (!fn || qr != cudaDriverEntryPointSuccess) {
      std::fprintf(stderr, "[b200-hevm] fatal: cuTensorMapEncodeTiled not available\n");
      std::abort();
    }
    encode = (EncodeFn)fn;
    CUDA_CHECK(cudaFuncSetAttribute((const void *)k_ntt_fwd_cluster<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(NC_SMEM_WORDS * sizeof(u64))));
  }
  const size_t N = (size_t)1 << logN;
  CUtensorMap map;
  const cuuint64_t dims[3] = {256, N / 256, (cuuint64_t)nl};
  const cuuint64_t strides[2] = {256 * sizeof(u64), N * sizeof(u64)};
  const cuuint32_t box[3] = {4, (cuuint32_t)(N / 256), 1}, estr[3] = {1, 1, 1};
  const CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, (void *)src, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    std::fprintf(stderr, "[b200-hevm] fatal: cuTensorMapEncodeTiled failed (%d)\n", (int)r);
    std::abort();
  }
  PRE_LAUNCH(s, KC_OTHER);
  k_ntt_fwd_cluster<7><<<nl * NC_CLUSTER, NC_WARPS * 32, NC_SMEM_WORDS * sizeof(u64), s>>>(map, T, dst, prime0, pstep);
  POST_LAUNCH_S(s);
}

static int fused_grid_cap() {
  static const int cap = std::getenv("HEVM_FUSED_GRID") ? std::max(1, std::atoi(std::getenv("HEVM_FUSED_GRID"))) : 148 * 4;
  return cap;
}
// Single ops keep the five (three) PDL-chained launches: measured on B200 (profiles/r02_fused_vs_launches.md) the
// persistent single-launch form is SLOWER for one ciphertext (rotate l = 13: 177 us vs 147 us; l = 1: 44 vs 44 us) --
// a phase costs 8-10 us because of the serial instruction stream of its warp jobs, not because of the launch --
// and only pays off when one launch carries a batch of independent ciphertexts.  HEVM_FUSED=1 forces it everywhere.
bool GpuLauncher::fused() const {
  static const bool on = std::getenv("HEVM_FUSED") && std::atoi(std::getenv("HEVM_FUSED")) != 0;
  return on;
}
template <int LOGA, int MODE> void GpuLauncher::ks_fused(const KsFusedArgs &a) {
  const KsUnits K = ks_units<LOGA, MODE>(a);
  if (K.total <= 0) return;
  const size_t warp_bytes = (size_t)WARPS_PER_CTA * Geo<LOGA>::WARP_WORDS * sizeof(u64);
  const size_t mac_bytes = MODE == FUSED_RESCALE ? 0 : mac_smem_words(a.l) * sizeof(u64);
  static bool optin = false; // per (LOGA, MODE) instantiation
  if (!optin) {
    const size_t mx = std::max(warp_bytes, mac_smem_words(HEVM_MAXL) * sizeof(u64));
    CUDA_CHECK(cudaFuncSetAttribute((const void *)k_ks_fused<LOGA, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mx));
    optin = true;
  }
  PRE_LAUNCH(stream, MODE == LD_GALOIS ? KC_KS_FUSED_GALOIS : MODE == LD_PRODUCT ? KC_KS_FUSED_RELIN : KC_RESCALE_FUSED);
  launch_pdl(k_ks_fused<LOGA, MODE>, std::min(K.total, fused_grid_cap()), CTA_THREADS, std::max(warp_bytes, mac_bytes), stream, a);
  POST_LAUNCH_S(stream);
}
#define INSTANTIATE(LOGA)                                                                                              \
  template void GpuLauncher::intt_B<LOGA, LD_PLAIN>(const ArgsInttB &, int);                                           \
  template void GpuLauncher::intt_B<LOGA, LD_GALOIS>(const ArgsInttB &, int);                                          \
  template void GpuLauncher::intt_B<LOGA, LD_PRODUCT>(const ArgsInttB &, int);                                         \
  template void GpuLauncher::intt_B<LOGA, LD_DECRYPT>(const ArgsInttB &, int);                                         \
  template void GpuLauncher::intt_A<LOGA>(const ArgsInttA &, int);                                                     \
  template void GpuLauncher::fwd_A<LOGA, PRE_NONE>(const ArgsFwdA &, int);                                             \
  template void GpuLauncher::fwd_A<LOGA, PRE_MODUP>(const ArgsFwdA &, int);                                            \
  template void GpuLauncher::fwd_A<LOGA, PRE_ROUND>(const ArgsFwdA &, int);                                            \
  template void GpuLauncher::fwd_B<LOGA, EPI_CANON>(const ArgsFwdB &, int);                                            \
  template void GpuLauncher::fwd_B<LOGA, EPI_MODDOWN_GALOIS>(const ArgsFwdB &, int);                                   \
  template void GpuLauncher::fwd_B<LOGA, EPI_MODDOWN_RELIN>(const ArgsFwdB &, int);                                    \
  template void GpuLauncher::fwd_B<LOGA, EPI_RESCALE>(const ArgsFwdB &, int);                                          \
  template void GpuLauncher::mac<LOGA>(const ArgsFwdB &, int);                                                         \
  template void GpuLauncher::invA_fwdA<LOGA, PRE_MODUP>(const ArgsInvFwdA &, int);                                     \
  template void GpuLauncher::invA_fwdA<LOGA, PRE_ROUND>(const ArgsInvFwdA &, int);                                     \
  template void GpuLauncher::ks_fused<LOGA, LD_GALOIS>(const KsFusedArgs &);                                           \
  template void GpuLauncher::ks_fused<LOGA, LD_PRODUCT>(const KsFusedArgs &);                                          \
  template void GpuLauncher::ks_fused<LOGA, FUSED_RESCALE>(const KsFusedArgs &);
INSTANTIATE(6)
INSTANTIATE(7)
INSTANTIATE(8)
INSTANTIATE(9)

// =====================================================================================
// Peer-to-peer exchange kernels of the limb-sharded key switch: plain stores into peer HBM over NVLink (the peer blocks
// are CUDA-IPC mappings), completion by an epoch flag written with system-scope release after a system-scope fence.
// =====================================================================================
__global__ void __launch_bounds__(256) k_p2p_push(const u64 *src, size_t words, PeerPtrs peer, size_t dst_off, int self, int world,
                                                  unsigned *done_ctr, size_t flag_off, unsigned long long epoch) {
  // blockIdx.y = destination peer (one load + one 16-byte store per thread: many independent stores in flight per link)
  const size_t nvec = words / 2; // 16-byte pieces (limb rows are multiples of 16 bytes)
  const int g = (self + 1 + blockIdx.y) % world;
  for (size_t v = blockIdx.x * (size_t)blockDim.x + threadIdx.x; v < nvec; v += (size_t)gridDim.x * blockDim.x) {
    u64 a, b;
    asm volatile("ld.global.cg.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(src + 2 * v) : "memory");
    asm volatile("st.global.v2.u64 [%0], {%1,%2};" ::"l"(peer.p[g] + dst_off + 2 * v), "l"(a), "l"(b) : "memory");
  }
  __threadfence_system();
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0) last = atomicAdd(done_ctr, 1u) == gridDim.x * gridDim.y - 1;
  __syncthreads();
  if (last) { // every CTA's stores are ordered before its increment: publish
    if (threadIdx.x == 0) *done_ctr = 0;
    if ((int)threadIdx.x < world && (int)threadIdx.x != self) {
      __threadfence_system();
      asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(peer.p[threadIdx.x] + flag_off + self), "l"(epoch) : "memory");
    }
  }
}
// The same push with the TMA engine: one elected thread per CTA bulk-loads a 16 KB chunk into shared memory
// (cp.async.bulk global -> shared, mbarrier completion) and bulk-stores it into the peer's block (cp.async.bulk shared ->
// global on the CUDA-IPC mapped address): large NVLink transactions instead of 16-byte stores.  grid = (chunks, peers).
#define P2P_CHUNK_WORDS 2048
__global__ void __launch_bounds__(32) k_p2p_push_tma(const u64 *src, size_t words, PeerPtrs peer, size_t dst_off, int self, int world,
                                                     unsigned *done_ctr, size_t flag_off, unsigned long long epoch) {
  __shared__ __align__(128) u64 buf[P2P_CHUNK_WORDS];
  __shared__ __align__(8) u64 bar;
  const int g = (self + 1 + blockIdx.y) % world;
  const size_t off = (size_t)blockIdx.x * P2P_CHUNK_WORDS;
  if (threadIdx.x == 0 && off < words) {
    const unsigned bytes = (unsigned)((words - off < P2P_CHUNK_WORDS ? words - off : P2P_CHUNK_WORDS) * 8);
    const unsigned sb = (unsigned)__cvta_generic_to_shared(buf), sbar = (unsigned)__cvta_generic_to_shared(&bar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sbar) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sbar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sb), "l"(src + off), "r"(bytes), "r"(sbar)
                 : "memory");
    asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(sbar)
                 : "memory");
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(peer.p[g] + dst_off + off), "r"(sb), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); // the bulk store has completed (not merely been read out of shared memory)
    asm volatile("fence.proxy.async;" ::: "memory");
  }
  __threadfence_system();
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0) last = atomicAdd(done_ctr, 1u) == gridDim.x * gridDim.y - 1;
  __syncthreads();
  if (last) {
    if (threadIdx.x == 0) *done_ctr = 0;
    if ((int)threadIdx.x < world && (int)threadIdx.x != self) {
      __threadfence_system();
      asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(peer.p[threadIdx.x] + flag_off + self), "l"(epoch) : "memory");
    }
  }
}
__global__ void k_p2p_wait(const u64 *block, size_t flag_off, int self, int world, int only, unsigned long long epoch) {
  const int g = threadIdx.x;
  if (g < world && g != self && (only < 0 || g == only)) {
    unsigned long long v;
    const long long t0 = clock64();
    do {
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(block + flag_off + g) : "memory");
      if (v < epoch) {
        __nanosleep(64);
        if (clock64() - t0 > 20000000000ll) { // ~10 s: a peer never arrived -- fail loudly instead of hanging the GPU
          printf("[b200-hevm] fatal: p2p wait timed out (rank %d waiting for rank %d, epoch %llu)\n", self, g, epoch);
          __trap();
        }
      }
    } while (v < epoch);
  }
}
void launch_p2p_push(cudaStream_t s, const u64 *src, size_t words, const PeerPtrs &peer, size_t dst_off, int self, int world,
                     unsigned *done_ctr, size_t flag_off, unsigned long long epoch) {
  static const int mode = std::getenv("HEVM_P2P_PUSH") ? std::atoi(std::getenv("HEVM_P2P_PUSH")) : 1; // 1 = TMA bulk copies, 0 = 16-byte stores
  if (mode == 1 && words % 2 == 0) {
    const size_t chunks = words ? (words + P2P_CHUNK_WORDS - 1) / P2P_CHUNK_WORDS : 1;
    k_p2p_push_tma<<<dim3((unsigned)chunks, (unsigned)(world > 1 ? world - 1 : 1)), 32, 0, s>>>(src, words, peer, dst_off, self, world, done_ctr, flag_off, epoch);
    POST_LAUNCH_S(s);
    return;
  }
  size_t g = (words / 2 + 255) / 256;
  if (g < 1) g = 1;
  if (g > 148) g = 148;
  k_p2p_push<<<dim3((unsigned)g, (unsigned)(world > 1 ? world - 1 : 1)), 256, 0, s>>>(src, words, peer, dst_off, self, world, done_ctr, flag_off, epoch);
  POST_LAUNCH_S(s);
}
__global__ void k_p2p_flag(u64 *flag, unsigned long long epoch) {
  __threadfence_system();
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(flag), "l"(epoch) : "memory");
}
void launch_p2p_flag(cudaStream_t s, u64 *flag, unsigned long long epoch) {
  k_p2p_flag<<<1, 1, 0, s>>>(flag, epoch);
  POST_LAUNCH_S(s);
}
void launch_p2p_wait(cudaStream_t s, const u64 *block, size_t flag_off, int self, int world, int only, unsigned long long epoch) {
  k_p2p_wait<<<1, 32, 0, s>>>(block, flag_off, self, world, only, epoch);
  POST_LAUNCH_S(s);
}

// =====================================================================================
// element-wise kernels.  One thread = 4 consecutive coefficients (256-bit accesses).
// =====================================================================================
template <int OP>
__global__ void __launch_bounds__(256) k_elementwise(const NttTables *T, int logN, u64 *out, const u64 *a, const u64 *b,
                                                     const u64 *p, size_t pitch, int l, size_t nvec) {
  const size_t polyw = (size_t)l << logN;
  for (size_t v = blockIdx.x * (size_t)blockDim.x + threadIdx.x; v < nvec; v += (size_t)gridDim.x * blockDim.x) {
    const size_t w = v * 4;
    const int K = w >= polyw;
    const size_t rem = w - (K ? polyw : 0);
    const int i = (int)(rem >> logN);
    const size_t off = (size_t)K * pitch + rem;
    const u64 q = T->mod[i].q;
    u64 x[4], y[4], r[4];
    ldg_stream4(a + off, x[0], x[1], x[2], x[3]);
    if (OP == EW_ADD) {
      ldg_stream4(b + off, y[0], y[1], y[2], y[3]);
#pragma unroll
      for (int e = 0; e < 4; e++) r[e] = csub(x[e] + y[e], q);
    } else if (OP == EW_NEG) {
#pragma unroll
      for (int e = 0; e < 4; e++) r[e] = negmod(x[e], q);
    } else if (OP == EW_ADDP) {
      if (K == 0) {
        ldg_stream4(p + rem, y[0], y[1], y[2], y[3]);
#pragma unroll
        for (int e = 0; e < 4; e++) r[e] = csub(x[e] + y[e], q);
      } else {
#pragma unroll
        for (int e = 0; e < 4; e++) r[e] = x[e];
      }
    } else if (OP == EW_MULP) {
      const ModQ m = T->mod[i];
      ldg_stream4(p + rem, y[0], y[1], y[2], y[3]);
#pragma unroll
      for (int e = 0; e < 4; e++) r[e] = mulmod(x[e], y[e], m);
    } else if (OP == EW_MULP_ADD) { // multiply_plain followed by add of another ciphertext: same canonical values, one HBM pass
      const ModQ m = T->mod[i];
      u64 z[4];
      ldg_stream4(p + rem, y[0], y[1], y[2], y[3]);
      ldg_stream4(b + off, z[0], z[1], z[2], z[3]);
#pragma unroll
      for (int e = 0; e < 4; e++) r[e] = csub(mulmod(x[e], y[e], m) + z[e], q);
    } else {
#pragma unroll
      for (int e = 0; e < 4; e++) r[e] = x[e];
    }
    stg4(out + off, r[0], r[1], r[2], r[3]);
  }
}
// acc + x0*p0 + x1*p1 + ...: the canonical value after every step equals the one the separate mulcp / addcc kernels
// produce, so the final residues are bit-identical
__global__ void __launch_bounds__(256) k_mulp_add_n(const NttTables *T, int logN, u64 *out, const u64 *acc, MulpTerms t, size_t pitch,
                                                    int l, size_t nvec) {
  const size_t polyw = (size_t)l << logN;
  for (size_t v = blockIdx.x * (size_t)blockDim.x + threadIdx.x; v < nvec; v += (size_t)gridDim.x * blockDim.x) {
    const size_t w = v * 4;
    const int K = w >= polyw;
    const size_t rem = w - (K ? polyw : 0);
    const int i = (int)(rem >> logN);
    const size_t off = (size_t)K * pitch + rem;
    const ModQ m = T->mod[i];
    u64 r[4];
    ldg_stream4(acc + off, r[0], r[1], r[2], r[3]);
    for (int k = 0; k < t.n; k++) {
      u64 x[4], y[4];
      ldg_stream4(t.x[k] + off, x[0], x[1], x[2], x[3]);
      ldg_stream4(t.p[k] + rem, y[0], y[1], y[2], y[3]);
#pragma unroll
      for (int e = 0; e < 4; e++) r[e] = csub(mulmod(x[e], y[e], m) + r[e], m.q);
    }
    stg4(out + off, r[0], r[1], r[2], r[3]);
  }
}
static int ew_grid(size_t nthreads);
void launch_mulp_add_n(cudaStream_t s, const NttTables *T, int logN, u64 *out, const u64 *acc, const MulpTerms &t, size_t pitch, int l) {
  const size_t nvec = ((size_t)2 * l << logN) / 4;
  PRE_LAUNCH(s, KC_ELEMENTWISE);
  k_mulp_add_n<<<ew_grid(nvec), 256, 0, s>>>(T, logN, out, acc, t, pitch, l, nvec);
  POST_LAUNCH_S(s);
}
static int ew_grid(size_t nthreads) {
  size_t g = (nthreads + 255) / 256;
  const size_t cap = 148 * 8; // 8 resident CTAs of 256 threads per SM, 148 SMs
  return (int)(g < cap ? g : cap);
}
void launch_elementwise(cudaStream_t s, int op, const NttTables *T, int logN, u64 *out, const u64 *a, const u64 *b,
                        const u64 *p, size_t pitch, int l) {
  const size_t nvec = ((size_t)2 * l << logN) / 4;
  const int g = ew_grid(nvec);
  PRE_LAUNCH(s, KC_ELEMENTWISE);
  switch (op) {
  case EW_ADD: k_elementwise<EW_ADD><<<g, 256, 0, s>>>(T, logN, out, a, b, p, pitch, l, nvec); break;
  case EW_NEG: k_elementwise<EW_NEG><<<g, 256, 0, s>>>(T, logN, out, a, b, p, pitch, l, nvec); break;
  case EW_ADDP: k_elementwise<EW_ADDP><<<g, 256, 0, s>>>(T, logN, out, a, b, p, pitch, l, nvec); break;
  case EW_MULP: k_elementwise<EW_MULP><<<g, 256, 0, s>>>(T, logN, out, a, b, p, pitch, l, nvec); break;
  case EW_MULP_ADD: k_elementwise<EW_MULP_ADD><<<g, 256, 0, s>>>(T, logN, out, a, b, p, pitch, l, nvec); break;
  default: k_elementwise<EW_COPY><<<g, 256, 0, s>>>(T, logN, out, a, b, p, pitch, l, nvec); break;
  }
  POST_LAUNCH_S(s);
}

// =====================================================================================
// samplers (specification shared with the oracle, oracle/ckks_oracle.hpp "sampler"):
//   rnd(key, stream, idx) = 64-bit word (idx & 7) of the ChaCha20 block with the 256-bit key, nonce = stream,
//   block counter = idx >> 3.  One thread computes one block = the values of 8 (small) or 4 (uniform) coefficients.
// =====================================================================================
__device__ __forceinline__ u32 rotl32(u32 v, int c) { return __funnelshift_l(v, v, c); }
__device__ __forceinline__ void chacha20_block(const Seed256 &key, u64 nonce, u64 counter, u64 (&out)[8]) {
  u32 x[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u, key.k[0], key.k[1], key.k[2], key.k[3], key.k[4], key.k[5], key.k[6],
               key.k[7], (u32)counter, (u32)(counter >> 32), (u32)nonce, (u32)(nonce >> 32)};
  u32 in[16];
#pragma unroll
  for (int i = 0; i < 16; i++) in[i] = x[i];
#define QR(a, b, c, d)                                                                                                 \
  x[a] += x[b], x[d] = rotl32(x[d] ^ x[a], 16), x[c] += x[d], x[b] = rotl32(x[b] ^ x[c], 12), x[a] += x[b],            \
      x[d] = rotl32(x[d] ^ x[a], 8), x[c] += x[d], x[b] = rotl32(x[b] ^ x[c], 7)
#pragma unroll 1
  for (int r = 0; r < 10; r++) {
    QR(0, 4, 8, 12), QR(1, 5, 9, 13), QR(2, 6, 10, 14), QR(3, 7, 11, 15);
    QR(0, 5, 10, 15), QR(1, 6, 11, 12), QR(2, 7, 8, 13), QR(3, 4, 9, 14);
  }
#undef QR
#pragma unroll
  for (int j = 0; j < 8; j++) out[j] = (u64)(x[2 * j] + in[2 * j]) | ((u64)(x[2 * j + 1] + in[2 * j + 1]) << 32);
}
__device__ __forceinline__ int small_from_word(u64 w, int cbd) {
  return cbd ? __popcll(w & 0x1FFFFF) - __popcll((w >> 21) & 0x1FFFFF) : (int)(w % 3) - 1;
}
template <int CBD>
__global__ void k_sample_small(const NttTables *T, int logN, u64 *out, int limbs, Seed256 key, u64 stream, const u64 *ctr) {
  const size_t N = (size_t)1 << logN;
  if (ctr) stream += *ctr * 4; // encryption counter base kept on the device (graph replays advance it)
  for (size_t b = blockIdx.x * (size_t)blockDim.x + threadIdx.x; b < N / 8; b += (size_t)gridDim.x * blockDim.x) {
    u64 w[8];
    chacha20_block(key, stream, b, w);
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const int t = small_from_word(w[j], CBD);
      for (int i = 0; i < limbs; i++) out[(size_t)i * N + 8 * b + j] = t < 0 ? T->mod[i].q - (u64)(-t) : (u64)t;
    }
  }
}
// the three small polynomials of one public-key encryption in one launch: blockIdx.y = 0 ternary u (stream0),
// 1 and 2 the centred-binomial errors e0, e1 (stream0 + 1, + 2); out = [3][limbs][N]; same values as three
// k_sample_small launches
__global__ void k_sample_enc(const NttTables *T, int logN, u64 *out, int limbs, Seed256 key, u64 stream0, const u64 *ctr) {
  const size_t N = (size_t)1 << logN;
  const int which = blockIdx.y;
  u64 stream = stream0 + which;
  if (ctr) stream += *ctr * 4;
  u64 *o = out + (size_t)which * limbs * N;
  for (size_t b = blockIdx.x * (size_t)blockDim.x + threadIdx.x; b < N / 8; b += (size_t)gridDim.x * blockDim.x) {
    u64 w[8];
    chacha20_block(key, stream, b, w);
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const int t = small_from_word(w[j], which != 0);
      for (int i = 0; i < limbs; i++) o[(size_t)i * N + 8 * b + j] = t < 0 ? T->mod[i].q - (u64)(-t) : (u64)t;
    }
  }
}
static int ew_grid(size_t nthreads);
void launch_sample_enc(cudaStream_t s, const NttTables *T, int logN, u64 *out, int limbs, const Seed256 &key, u64 stream0, const u64 *ctr) {
  k_sample_enc<<<dim3(ew_grid(((size_t)1 << logN) / 8), 3), 256, 0, s>>>(T, logN, out, limbs, key, stream0, ctr);
  POST_LAUNCH_S(s);
}
__global__ void k_sample_uniform(const NttTables *T, int logN, u64 *out, int limbs, Seed256 key, u64 stream_base) {
  const size_t N = (size_t)1 << logN, total = (N / 4) * limbs; // one ChaCha block = the (hi, lo) pairs of 4 coefficients
  for (size_t v = blockIdx.x * (size_t)blockDim.x + threadIdx.x; v < total; v += (size_t)gridDim.x * blockDim.x) {
    const int i = (int)(v / (N / 4));
    const size_t b = v - (size_t)i * (N / 4);
    u64 w[8];
    chacha20_block(key, stream_base + i, b, w);
#pragma unroll
    for (int j = 0; j < 4; j++) out[(size_t)i * N + 4 * b + j] = reduce128(w[2 * j + 1], w[2 * j], T->mod[i]);
  }
}
void launch_sample_ternary(cudaStream_t s, const NttTables *T, int logN, u64 *out, int limbs, const Seed256 &key, u64 stream, const u64 *ctr) {
  k_sample_small<0><<<ew_grid(((size_t)1 << logN) / 8), 256, 0, s>>>(T, logN, out, limbs, key, stream, ctr);
  POST_LAUNCH_S(s);
}
void launch_sample_cbd(cudaStream_t s, const NttTables *T, int logN, u64 *out, int limbs, const Seed256 &key, u64 stream, const u64 *ctr) {
  k_sample_small<1><<<ew_grid(((size_t)1 << logN) / 8), 256, 0, s>>>(T, logN, out, limbs, key, stream, ctr);
  POST_LAUNCH_S(s);
}
void launch_sample_uniform(cudaStream_t s, const NttTables *T, int logN, u64 *out, int limbs, const Seed256 &key, u64 stream_base) {
  k_sample_uniform<<<ew_grid(((size_t)limbs << logN) / 4), 256, 0, s>>>(T, logN, out, limbs, key, stream_base);
  POST_LAUNCH_S(s);
}

// =====================================================================================
// key generation / encryption helpers
// =====================================================================================
__global__ void k_ksk_finish(const NttTables *T, int logN, int L, u64 *c0, const u64 *c1, const u64 *sk, const u64 *e,
                             const u64 *newkey, int digit) {
  const size_t N = (size_t)1 << logN, total = N * L;
  const u64 pspecial = T->mod[L - 1].q;
  for (size_t v = blockIdx.x * (size_t)blockDim.x + threadIdx.x; v < total; v += (size_t)gridDim.x * blockDim.x) {
    const int i = (int)(v >> logN);
    const ModQ m = T->mod[i];
    u64 r = negmod(csub(mulmod(c1[v], sk[v], m) + e[v], m.q), m.q);
    if (i == digit) r = csub(r + mulmod(newkey[v], reduce64(pspecial, m), m), m.q);
    c0[v] = r;
  }
}
// canonical key-switch key -> radix-2^30 split storage (the only consumer is the key inner product, body_mac_dot)
__global__ void k_key_split(u64 *key, size_t words) {
  for (size_t v = blockIdx.x * (size_t)blockDim.x + threadIdx.x; v < words; v += (size_t)gridDim.x * blockDim.x) key[v] = split30(key[v]);
}
void launch_key_split(cudaStream_t s, u64 *key, size_t words) { k_key_split<<<ew_grid(words), 256, 0, s>>>(key, words); }
void launch_ksk_finish(cudaStream_t s, const NttTables *T, int logN, int L, u64 *c0, const u64 *c1, const u64 *sk,
                       const u64 *e, const u64 *newkey, int digit) {
  k_ksk_finish<<<ew_grid((size_t)L << logN), 256, 0, s>>>(T, logN, L, c0, c1, sk, e, newkey, digit);
  POST_LAUNCH_S(s);
}
__global__ void k_square(const NttTables *T, int logN, int L, u64 *out, const u64 *in) {
  const size_t total = (size_t)L << logN;
  for (size_t v = blockIdx.x * (size_t)blockDim.x + threadIdx.x; v < total; v += (size_t)gridDim.x * blockDim.x)
    out[v] = mulmod(in[v], in[v], T->mod[v >> logN]);
}
void launch_square(cudaStream_t s, const NttTables *T, int logN, int L, u64 *out, const u64 *in) {
  k_square<<<ew_grid((size_t)L << logN), 256, 0, s>>>(T, logN, L, out, in);
  POST_LAUNCH_S(s);
}
__global__ void k_galois_gather(int logN, int L, u64 *out, const u64 *in, u32 elt) {
  const size_t N = (size_t)1 << logN, total = N * L;
  for (size_t v = blockIdx.x * (size_t)blockDim.x + threadIdx.x; v < total; v += (size_t)gridDim.x * blockDim.x) {
    const size_t i = v >> logN;
    out[v] = in[i * N + galois_src_index((u32)(v & (N - 1)), elt, logN)];
  }
}
void launch_galois_gather(cudaStream_t s, int logN, int L, u64 *out, const u64 *in, u32 elt) {
  k_galois_gather<<<ew_grid((size_t)L << logN), 256, 0, s>>>(logN, L, out, in, elt);
  POST_LAUNCH_S(s);
}
__global__ void k_enc_combine(const NttTables *T, int logN, int nl, u64 *c, const u64 *u, const u64 *pk, size_t pk_pitch,
                              const u64 *e) {
  const size_t polyw = (size_t)nl << logN, total = 2 * polyw;
  for (size_t v = blockIdx.x * (size_t)blockDim.x + threadIdx.x; v < total; v += (size_t)gridDim.x * blockDim.x) {
    const int j = v >= polyw;
    const size_t rem = v - (j ? polyw : 0);
    const ModQ m = T->mod[rem >> logN];
    c[v] = csub(mulmod(u[rem], pk[(size_t)j * pk_pitch + rem], m) + e[v], m.q);
  }
}
void launch_enc_combine(cudaStream_t s, const NttTables *T, int logN, int nl, u64 *c, const u64 *u, const u64 *pk,
                        size_t pk_pitch, const u64 *e) {
  k_enc_combine<<<ew_grid((size_t)2 * nl << logN), 256, 0, s>>>(T, logN, nl, c, u, pk, pk_pitch, e);
  POST_LAUNCH_S(s);
}
__global__ void k_decrypt(const NttTables *T, int logN, int l, u64 *pt, const u64 *ct, size_t pitch, const u64 *sk) {
  const size_t total = (size_t)l << logN;
  for (size_t v = blockIdx.x * (size_t)blockDim.x + threadIdx.x; v < total; v += (size_t)gridDim.x * blockDim.x) {
    const ModQ m = T->mod[v >> logN];
    pt[v] = csub(ct[v] + mulmod(ct[pitch + v], sk[v], m), m.q);
  }
}
void launch_decrypt(cudaStream_t s, const NttTables *T, int logN, int l, u64 *pt, const u64 *ct, size_t pitch, const u64 *sk) {
  k_decrypt<<<ew_grid((size_t)l << logN), 256, 0, s>>>(T, logN, l, pt, ct, pitch, sk);
  POST_LAUNCH_S(s);
}

// =====================================================================================
// CKKS encoder / decoder (fp64).  Butterfly DAG and operation order of SEAL's
// DWTHandler::transform_from_rev / transform_to_rev with complex<double> arithmetic
// (SURVEY.md A.2.9); *_rn intrinsics are never contracted into FMAs.
// =====================================================================================
__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(__dadd_rn(a.x, b.x), __dadd_rn(a.y, b.y)); }
__device__ __forceinline__ double2 csubc(double2 a, double2 b) { return make_double2(__dsub_rn(a.x, b.x), __dsub_rn(a.y, b.y)); }
__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
  return make_double2(__dsub_rn(__dmul_rn(a.x, b.x), __dmul_rn(a.y, b.y)), __dadd_rn(__dmul_rn(a.x, b.y), __dmul_rn(a.y, b.x)));
}
__global__ void k_enc_scatter(int logN, const double *vals, int len, const u32 *slot_index, double2 *work,
                              unsigned long long *maxbits) {
  const size_t slots = (size_t)1 << (logN - 1);
  if (blockIdx.x == 0 && threadIdx.x == 0) *maxbits = 0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < slots; i += (size_t)gridDim.x * blockDim.x) {
    const double v = vals[i % len];
    work[slot_index[i]] = make_double2(v, 0.0);
    work[slot_index[slots + i]] = make_double2(v, -0.0);
  }
}
// decode's gather followed by encode's scatter of the same N/2 values, in place (the wrapper bootstrap re-encodes what it
// just decoded): position slot_index[i] is read and rewritten by thread i alone, slot_index[slots + i] is only written
__global__ void k_recode(int logN, const u32 *slot_index, double2 *work, unsigned long long *maxbits) {
  const size_t slots = (size_t)1 << (logN - 1);
  if (blockIdx.x == 0 && threadIdx.x == 0) *maxbits = 0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < slots; i += (size_t)gridDim.x * blockDim.x) {
    const double v = work[slot_index[i]].x;
    work[slot_index[i]] = make_double2(v, 0.0);
    work[slot_index[slots + i]] = make_double2(v, -0.0);
  }
}
// Up to four consecutive radix-2 stages per launch: a thread owns the 2^K points  base + t * gap_lo  (t < 2^K) that
// are closed under the stages with gaps gap_lo .. gap_lo * 2^(K-1), keeps them in registers and performs exactly the
// butterflies of SEAL's stage-by-stage loops on them (Gentleman-Sande with m groups of gap n/(2m), roots consumed at
// index (n - 2m) + 1 + group; Cooley-Tukey with root index m + group), in the same order per point -- so every
// intermediate value is the same fp64 number and the results stay bit-identical, with a quarter of the launches and
// memory passes of one kernel per stage.
template <int K, bool GS> __global__ void k_fft_multi(int logN, double2 *v, const double2 *roots, int gap_lo, double fix, int has_last) {
  constexpr int P = 1 << K;
  const size_t n = (size_t)1 << logN, sets = n >> K;
  for (size_t sidx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; sidx < sets; sidx += (size_t)gridDim.x * blockDim.x) {
    const size_t lo = sidx % (size_t)gap_lo, hi = sidx / (size_t)gap_lo;
    const size_t base = hi * ((size_t)gap_lo << K) + lo;
    double2 x[P];
#pragma unroll
    for (int t = 0; t < P; t++) x[t] = v[base + (size_t)t * gap_lo];
#pragma unroll
    for (int st = 0; st < K; st++) {
      const int half = GS ? (1 << st) : (P >> 1 >> st);   // register distance of the stage
      const size_t gap = (size_t)gap_lo * half;            // its distance in the array
      const size_t m = n / (2 * gap);
#pragma unroll
      for (int t = 0; t < P; t++)
        if (!(t & half)) {
          const size_t off = base + (size_t)t * gap_lo, g = off / (2 * gap);
          if (GS) {
            const double2 r = roots[(n - 2 * m) + 1 + g];
            const double2 u = x[t], w = x[t + half];
            if (has_last && m == 1) {
              const double2 sr = make_double2(__dmul_rn(r.x, fix), __dmul_rn(r.y, fix));
              const double2 sm_ = cadd(u, w);
              x[t] = make_double2(__dmul_rn(sm_.x, fix), __dmul_rn(sm_.y, fix));
              x[t + half] = cmul(csubc(u, w), sr);
            } else {
              x[t] = cadd(u, w);
              x[t + half] = cmul(csubc(u, w), r);
            }
          } else {
            const double2 r = roots[m + g];
            const double2 u = x[t], w = cmul(x[t + half], r);
            x[t] = cadd(u, w);
            x[t + half] = csubc(u, w);
          }
        }
    }
#pragma unroll
    for (int t = 0; t < P; t++) v[base + (size_t)t * gap_lo] = x[t];
  }
}
// The FFT_LC stages whose butterflies stay inside 2^FFT_LC-point contiguous chunks (gaps 1 .. 2^(FFT_LC-1)) in ONE
// launch: a 1024-thread CTA owns a chunk in shared memory, one butterfly per thread and stage, __syncthreads between
// stages.  Same butterflies on the same operands as the per-stage loops => bit-identical.
#define FFT_LC 11
template <bool GS> __global__ void __launch_bounds__(1024) k_fft_chunk(int logN, double2 *v, const double2 *roots, double fix, int has_last) {
  __shared__ double2 sh[1 << FFT_LC];
  const size_t n = (size_t)1 << logN, base = (size_t)blockIdx.x << FFT_LC;
  const int tid = threadIdx.x;
  sh[tid] = v[base + tid];
  sh[tid + 1024] = v[base + tid + 1024];
  __syncthreads();
  for (int st = 0; st < FFT_LC; st++) {
    const int gap = GS ? (1 << st) : (1 << (FFT_LC - 1 - st));
    const size_t m = n / (2 * (size_t)gap);
    const int gl = tid / gap, j = tid - gl * gap, off = 2 * gl * gap + j;
    const size_t g = (base + off) / (2 * (size_t)gap);
    const double2 u = sh[off], w = sh[off + gap];
    if (GS) {
      const double2 r = roots[(n - 2 * m) + 1 + g];
      if (has_last && m == 1) {
        const double2 sr = make_double2(__dmul_rn(r.x, fix), __dmul_rn(r.y, fix));
        const double2 sm_ = cadd(u, w);
        sh[off] = make_double2(__dmul_rn(sm_.x, fix), __dmul_rn(sm_.y, fix));
        sh[off + gap] = cmul(csubc(u, w), sr);
      } else {
        sh[off] = cadd(u, w);
        sh[off + gap] = cmul(csubc(u, w), r);
      }
    } else {
      const double2 r = roots[m + g];
      const double2 wr = cmul(w, r);
      sh[off] = cadd(u, wr);
      sh[off + gap] = csubc(u, wr);
    }
    __syncthreads();
  }
  v[base + tid] = sh[tid];
  v[base + tid + 1024] = sh[tid + 1024];
}
// The K = logN - FFT_LC strided stages (gaps 2^FFT_LC * 2^s) in ONE launch: a CTA owns 2^(FFT_LC-K) neighbouring columns
// (index mod 2^FFT_LC) with their 2^K points each, 2^FFT_LC points in shared memory, one butterfly per thread and stage.
// Loads / stores are contiguous runs of 2^(FFT_LC-K) points.  Same butterflies, same operands => bit-identical.
template <bool GS> __global__ void __launch_bounds__(1024) k_fft_cols(int logN, double2 *v, const double2 *roots, double fix) {
  __shared__ double2 sh[1 << FFT_LC];
  const int K = logN - FFT_LC, logC = FFT_LC - K, C = 1 << logC;
  const size_t n = (size_t)1 << logN;
  const int tid = threadIdx.x;
  const size_t col0 = (size_t)blockIdx.x << logC;
  for (int i = tid; i < (1 << FFT_LC); i += 1024) sh[i] = v[col0 + (i & (C - 1)) + ((size_t)(i >> logC) << FFT_LC)];
  __syncthreads();
  const int c = tid & (C - 1), bidx = tid >> logC;
  for (int st = 0; st < K; st++) {
    const int sft = GS ? st : K - 1 - st;
    const int t = ((bidx >> sft) << (sft + 1)) | (bidx & ((1 << sft) - 1));
    const size_t gap = (size_t)1 << (FFT_LC + sft), m = n / (2 * gap);
    const size_t off = col0 + c + ((size_t)t << FFT_LC), g = off / (2 * gap);
    const int i0 = (t << logC) + c, i1 = ((t + (1 << sft)) << logC) + c;
    const double2 u = sh[i0], w = sh[i1];
    if (GS) {
      const double2 r = roots[(n - 2 * m) + 1 + g];
      if (m == 1) {
        const double2 sr = make_double2(__dmul_rn(r.x, fix), __dmul_rn(r.y, fix));
        const double2 sm_ = cadd(u, w);
        sh[i0] = make_double2(__dmul_rn(sm_.x, fix), __dmul_rn(sm_.y, fix));
        sh[i1] = cmul(csubc(u, w), sr);
      } else {
        sh[i0] = cadd(u, w);
        sh[i1] = cmul(csubc(u, w), r);
      }
    } else {
      const double2 r = roots[m + g];
      const double2 wr = cmul(w, r);
      sh[i0] = cadd(u, wr);
      sh[i1] = csubc(u, wr);
    }
    __syncthreads();
  }
  for (int i = tid; i < (1 << FFT_LC); i += 1024) v[col0 + (i & (C - 1)) + ((size_t)(i >> logC) << FFT_LC)] = sh[i];
}
// all logN stages: the FFT_LC contiguous stages in one chunk kernel, the strided ones in register groups of four
// (GS walks the gaps upward: chunk kernel first; CT walks them downward: chunk kernel last)
template <bool GS> static void launch_fft_all(cudaStream_t s, int logN, double2 *work, const double2 *roots, double fix) {
  const size_t n = (size_t)1 << logN;
  const bool chunked = logN >= FFT_LC;
  int done = 0;
  if (chunked && GS) {
    k_fft_chunk<true><<<(unsigned)(n >> FFT_LC), 1024, 0, s>>>(logN, work, roots, fix, logN == FFT_LC);
    POST_LAUNCH_S(s);
    done = FFT_LC;
  }
  const int strided_end = (chunked && !GS) ? logN - FFT_LC : logN;
  const int Kc = logN - FFT_LC;
  if (chunked && Kc >= 1 && Kc <= 6) { // all strided stages in one shared-memory launch
    k_fft_cols<GS><<<1u << Kc, 1024, 0, s>>>(logN, work, roots, fix);
    POST_LAUNCH_S(s);
    done = strided_end;
  }
  while (done < strided_end) {
    const int k = (strided_end - done >= 4) ? 4 : strided_end - done;
    // GS walks gaps 1, 2, 4, ... upward; CT walks gaps n/2, n/4, ... downward
    const int gap_lo = GS ? (1 << done) : (int)(n >> (done + k));
    const size_t sets = n >> k;
    const int has_last = GS && done + k == logN;
    switch (k) {
    case 4: k_fft_multi<4, GS><<<ew_grid(sets), 256, 0, s>>>(logN, work, roots, gap_lo, fix, has_last); break;
    case 3: k_fft_multi<3, GS><<<ew_grid(sets), 256, 0, s>>>(logN, work, roots, gap_lo, fix, has_last); break;
    case 2: k_fft_multi<2, GS><<<ew_grid(sets), 256, 0, s>>>(logN, work, roots, gap_lo, fix, has_last); break;
    default: k_fft_multi<1, GS><<<ew_grid(sets), 256, 0, s>>>(logN, work, roots, gap_lo, fix, has_last); break;
    }
    POST_LAUNCH_S(s);
    done += k;
  }
  if (chunked && !GS) {
    k_fft_chunk<false><<<(unsigned)(n >> FFT_LC), 1024, 0, s>>>(logN, work, roots, fix, 0);
    POST_LAUNCH_S(s);
  }
}
__global__ void k_enc_max(int logN, const double2 *work, unsigned long long *maxbits) {
  const size_t n = (size_t)1 << logN;
  double mx = 0.0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    mx = fmax(mx, fabs(work[i].x));
  for (int o = 16; o; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) atomicMax(maxbits, (unsigned long long)__double_as_longlong(mx)); // mx >= 0: bit order == value order
}
__global__ void k_enc_round(const NttTables *T, int logN, int level, const double2 *work, const unsigned long long *maxbits,
                            u64 *out) {
  const size_t n = (size_t)1 << logN;
  // ceil(log2(max(maxc,1))) computed exactly from the exponent / mantissa
  const double maxc = fmax(__longlong_as_double((long long)*maxbits), 1.0);
  int ex;
  const double mant = frexp(maxc, &ex); // maxc = mant * 2^ex, mant in [0.5,1)
  const int bits = (mant == 0.5) ? ex - 1 : ex;
  if (bits > 128) {
    if (blockIdx.x == 0 && threadIdx.x == 0) printf("[b200-hevm] fatal: encode: coefficient wider than 128 bits\n");
    __trap();
  }
  const double two64 = 18446744073709551616.0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    double c = round(work[i].x);
    const bool neg = signbit(c);
    c = fabs(c);
    if (bits <= 64) {
      const u64 cu = (u64)c;
      for (int j = 0; j < level; j++) {
        const ModQ m = T->mod[j];
        const u64 r = reduce64(cu, m);
        out[(size_t)j * n + i] = neg ? negmod(r, m.q) : r;
      }
    } else {
      const u64 lo = (u64)fmod(c, two64), hi = (u64)(c / two64);
      for (int j = 0; j < level; j++) {
        const ModQ m = T->mod[j];
        const u64 r = reduce128(lo, hi, m);
        out[(size_t)j * n + i] = neg ? negmod(r, m.q) : r;
      }
    }
  }
}
void launch_encode(cudaStream_t s, const NttTables *T, const EncoderTables &E, int logN, const double *vals, int len,
                   int level, double scale, double2 *work, unsigned long long *maxbits, u64 *out) {
  const size_t n = (size_t)1 << logN;
  if (vals) { // nullptr: the slots are already in `work` (launch_decode with out == nullptr re-scattered them)
    k_enc_scatter<<<ew_grid(n / 2), 256, 0, s>>>(logN, vals, len, E.slot_index, work, maxbits);
    POST_LAUNCH_S(s);
  }
  const double fix = scale / (double)n;
  launch_fft_all<true>(s, logN, work, E.inv_root, fix);
  k_enc_max<<<ew_grid(n), 256, 0, s>>>(logN, work, maxbits);
  POST_LAUNCH_S(s);
  k_enc_round<<<ew_grid(n), 256, 0, s>>>(T, logN, level, work, maxbits, out);
  POST_LAUNCH_S(s);
}

#define DEC_MAXW 33
__global__ void k_dec_compose(const NttTables *T, int logN, int l, const u64 *coeff, DecodeTables D, double inv_scale,
                              double2 *work) {
  const size_t n = (size_t)1 << logN;
  const double two64 = 18446744073709551616.0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    u64 val[DEC_MAXW];
    for (int k = 0; k <= l; k++) val[k] = 0;
    for (int j = 0; j < l; j++) {
      const ModQ m = T->mod[j];
      const u64 t = mulmod(coeff[(size_t)j * n + i], D.invp[j], m);
      // val += punct[j] * t   (< Q), then one conditional subtraction of Q
      u64 carry = 0, cy = 0;
      for (int k = 0; k < l; k++) {
        const u64 pw = D.punct[(size_t)j * l + k];
        u64 lo = pw * t, hi = __umul64hi(pw, t);
        lo += carry;
        hi += (lo < carry);
        carry = hi;
        u64 s1 = val[k] + lo;
        u64 c1 = s1 < lo;
        u64 s2 = s1 + cy;
        u64 c2 = s2 < cy;
        val[k] = s2;
        cy = c1 + c2;
      }
      val[l] = carry + cy; // carry is 0 because punct[j]*t < Q
      bool ge = val[l] != 0;
      if (!ge) {
        ge = true;
        for (int k = l - 1; k >= 0; k--) {
          const u64 qk = D.Q[k];
          if (val[k] != qk) {
            ge = val[k] > qk;
            break;
          }
        }
      }
      if (ge) {
        u64 borrow = 0;
        for (int k = 0; k < l; k++) {
          const u64 qk = D.Q[k];
          const u64 d1 = val[k] - qk;
          const u64 b1 = val[k] < qk;
          const u64 d2 = d1 - borrow;
          const u64 b2 = d1 < borrow;
          val[k] = d2;
          borrow = b1 + b2;
        }
        val[l] -= borrow;
      }
    }
    bool upper = true; // val >= half ?
    for (int k = l - 1; k >= 0; k--) {
      const u64 hk = D.half[k];
      if (val[k] != hk) {
        upper = val[k] > hk;
        break;
      }
    }
    double r = 0.0, s64 = inv_scale;
    if (upper) {
      for (int j = 0; j < l; j++, s64 = __dmul_rn(s64, two64)) {
        const u64 qj = D.Q[j];
        if (val[j] > qj) {
          const u64 diff = val[j] - qj;
          r = __dadd_rn(r, diff ? __dmul_rn((double)diff, s64) : 0.0);
        } else {
          const u64 diff = qj - val[j];
          r = __dsub_rn(r, diff ? __dmul_rn((double)diff, s64) : 0.0);
        }
      }
    } else {
      for (int j = 0; j < l; j++, s64 = __dmul_rn(s64, two64)) {
        const u64 cc = val[j];
        r = __dadd_rn(r, cc ? __dmul_rn((double)cc, s64) : 0.0);
      }
    }
    work[i] = make_double2(r, 0.0);
  }
}
__global__ void k_dec_gather(int logN, const double2 *work, const u32 *slot_index, double *out) {
  const size_t slots = (size_t)1 << (logN - 1);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < slots; i += (size_t)gridDim.x * blockDim.x)
    out[i] = work[slot_index[i]].x;
}
void launch_decode(cudaStream_t s, const NttTables *T, const EncoderTables &E, const DecodeTables &D, int logN, int l,
                   const u64 *coeff, double scale, double2 *work, double *out, unsigned long long *maxbits) {
  const size_t n = (size_t)1 << logN;
  if (l + 1 > DEC_MAXW) {
    std::fprintf(stderr, "[b200-hevm] fatal: decode supports at most %d limbs\n", DEC_MAXW - 1);
    std::abort();
  }
  k_dec_compose<<<ew_grid(n), 256, 0, s>>>(T, logN, l, coeff, D, 1.0 / scale, work);
  POST_LAUNCH_S(s);
  launch_fft_all<false>(s, logN, work, E.fwd_root, 0.0);
  if (out) {
    k_dec_gather<<<ew_grid(n / 2), 256, 0, s>>>(logN, work, E.slot_index, out);
  } else { // decode -> encode round trip: leave the values in `work`, scattered for the encoder
    k_recode<<<ew_grid(n / 2), 256, 0, s>>>(logN, E.slot_index, work, maxbits);
  }
  POST_LAUNCH_S(s);
}
