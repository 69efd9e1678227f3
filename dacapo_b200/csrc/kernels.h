// Launch interface of the sm_100a kernels (kernels.cu) used by the HEVM host code (vm.cu).
#pragma once
#include "ops.hpp"
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>

#define CUDA_CHECK(x)                                                                                                  \
  do {                                                                                                                 \
    cudaError_t e_ = (x);                                                                                              \
    if (e_ != cudaSuccess) {                                                                                           \
      std::fprintf(stderr, "[b200-hevm] fatal: CUDA error %s at %s:%d: %s\n", cudaGetErrorName(e_), __FILE__, __LINE__, \
                   cudaGetErrorString(e_));                                                                            \
      std::abort();                                                                                                    \
    }                                                                                                                  \
  } while (0)

// counts every kernel this library launches (bench.py reports it as gpu_launches)
extern unsigned long long g_launch_count;
extern bool g_pdl_suspended; // set while run() is being captured into a CUDA graph (plain edges there)

// Optional per-kernel-class timing with CUDA events on the launching stream (bench.py roofline).
enum {
  KC_INTT_B_PLAIN = 0, KC_INTT_B_GALOIS, KC_INTT_B_PRODUCT, KC_INTT_A, KC_FWD_A_NONE, KC_FWD_A_MODUP, KC_FWD_A_ROUND,
  KC_FWD_B_CANON, KC_FWD_B_MAC, KC_FWD_B_MODDOWN_GALOIS, KC_FWD_B_MODDOWN_RELIN, KC_FWD_B_RESCALE, KC_INVA_FWDA_MODUP, KC_INVA_FWDA_ROUND, KC_ELEMENTWISE,
  KC_KS_FUSED_GALOIS, KC_KS_FUSED_RELIN, KC_RESCALE_FUSED, KC_OTHER, KC_COUNT
};
extern const char *const g_kernel_class_names[KC_COUNT];
struct KProfiler {
  bool on = false, pending = false;
  struct Rec { int cls; cudaEvent_t a, b; };
  std::vector<Rec> recs;
  std::vector<cudaEvent_t> pool;
  size_t used = 0;
  double ms[KC_COUNT] = {0};
  long long cnt[KC_COUNT] = {0};
  cudaEvent_t ev();
  void begin(cudaStream_t s, int cls);
  void end(cudaStream_t s);
  void collect(); // synchronises the recorded events and accumulates ms / cnt
  void reset();
};
extern KProfiler g_prof;

struct GpuLauncher {
  cudaStream_t stream = nullptr;
  template <int LOGA, int LD> void intt_B(const ArgsInttB &a, int njobs);
  template <int LOGA> void intt_A(const ArgsInttA &a, int njobs);
  template <int LOGA, int PRE> void fwd_A(const ArgsFwdA &a, int njobs);
  template <int LOGA, int EPI> void fwd_B(const ArgsFwdB &a, int njobs);
  template <int LOGA, int PRE> void invA_fwdA(const ArgsInvFwdA &a, int njobs);
  template <int LOGA> void mac(const ArgsFwdB &a, int njobs); // one 4-warp CTA per job
  // single-launch key switch / rescale (ks_fused.cuh); HEVM_FUSED=0 falls back to the separate launches
  bool fused() const;
  template <int LOGA, int MODE> void ks_fused(const KsFusedArgs &a);
};

// ---- single-pass forward NTT of `nl` limbs (N = 2^15) in 8-CTA clusters: TMA tile loads + DSMEM transpose (ntt_cluster.cuh) ----
void launch_ntt_fwd_cluster(cudaStream_t s, const NttTables *T, int logN, const u64 *src, u64 *dst, int nl, int prime0, int pstep);

// ---- peer-to-peer exchange of the limb-sharded key switch (NVLink, CUDA IPC mapped peer memory) ----
// Every rank owns one exchange block with the same layout; peer.p[g] = that block of rank g mapped into this process.
struct PeerPtrs {
  u64 *p[8];
};
// copy `words` u64 from `src` (local) to offset `dst_off` of every peer's block, then -- once every CTA's stores are
// system-visible -- write `epoch` into flag word `flag_off + self` of every peer's block
void launch_p2p_push(cudaStream_t s, const u64 *src, size_t words, const PeerPtrs &peer, size_t dst_off, int self, int world,
                     unsigned *done_ctr, size_t flag_off, unsigned long long epoch);
void launch_p2p_flag(cudaStream_t s, u64 *flag, unsigned long long epoch); // one system-scope release store (after a copy-engine push)
// spin until flag word `flag_off + g` of the LOCAL block is >= epoch for every peer g (only >= 0: just that peer)
void launch_p2p_wait(cudaStream_t s, const u64 *block, size_t flag_off, int self, int world, int only, unsigned long long epoch);

// ---- element-wise ciphertext kernels (SEAL add/negate/add_plain/multiply_plain, limb drop) ----
// out = acc + sum_k x[k] * p[k] (k < n <= 8): a chain of multiply_plain + add, one HBM pass
struct MulpTerms {
  const u64 *x[8];
  const u64 *p[8];
  int n;
};
void launch_mulp_add_n(cudaStream_t s, const NttTables *T, int logN, u64 *out, const u64 *acc, const MulpTerms &t, size_t pitch, int l);
enum { EW_ADD = 0, EW_NEG = 1, EW_ADDP = 2, EW_MULP = 3, EW_COPY = 4, EW_MULP_ADD = 5 /* out = a * p + b */ };
// out[K][i][n] = op(a[K][i][n], b...) for K<2, i<l ; ct poly pitch = `pitch` words; plaintext `p` is [l][N]
void launch_elementwise(cudaStream_t s, int op, const NttTables *T, int logN, u64 *out, const u64 *a, const u64 *b,
                        const u64 *p, size_t pitch, int l);

// ---- samplers (specification shared with the oracle: oracle/ckks_oracle.hpp "sampler"): ChaCha20 keyed by 256 bits ----
struct Seed256 {
  u32 k[8];
};
void launch_sample_ternary(cudaStream_t s, const NttTables *T, int logN, u64 *out, int limbs, const Seed256 &key, u64 stream, const u64 *ctr = nullptr);
void launch_sample_cbd(cudaStream_t s, const NttTables *T, int logN, u64 *out, int limbs, const Seed256 &key, u64 stream, const u64 *ctr = nullptr);
void launch_sample_uniform(cudaStream_t s, const NttTables *T, int logN, u64 *out, int limbs, const Seed256 &key, u64 stream_base);

// ---- key generation / encryption helpers ----
// c0[i] = -(c1[i]*sk[i] + e[i]); if (digit >= 0 && i == digit) c0[i] += (p mod q_i) * newkey[i]
void launch_sample_enc(cudaStream_t s, const NttTables *T, int logN, u64 *out, int limbs, const Seed256 &key, u64 stream0, const u64 *ctr);
void launch_key_split(cudaStream_t s, u64 *key, size_t words);
void launch_ksk_finish(cudaStream_t s, const NttTables *T, int logN, int L, u64 *c0, const u64 *c1, const u64 *sk,
                       const u64 *e, const u64 *newkey, int digit);
void launch_square(cudaStream_t s, const NttTables *T, int logN, int L, u64 *out, const u64 *in);
void launch_galois_gather(cudaStream_t s, int logN, int L, u64 *out, const u64 *in, u32 elt);
// c[j][i] = u[i]*pk[j][i] + e[j][i]   (limbs i < nl; pk poly pitch pk_pitch; c/e poly pitch nl*N)
void launch_enc_combine(cudaStream_t s, const NttTables *T, int logN, int nl, u64 *c, const u64 *u, const u64 *pk,
                        size_t pk_pitch, const u64 *e);
// pt[i] = c0[i] + c1[i]*sk[i]
void launch_decrypt(cudaStream_t s, const NttTables *T, int logN, int l, u64 *pt, const u64 *ct, size_t pitch, const u64 *sk);

// ---- CKKS encoder (fp64; SEAL CKKSEncoder::encode_internal / decode_internal) ----
struct EncoderTables {
  const u32 *slot_index;   // [N]  matrix_reps_index_map
  const double2 *fwd_root; // [N]  root_powers_
  const double2 *inv_root; // [N]  inv_root_powers_
};
// values (device, `len` doubles, tiled to N/2 slots) -> coefficient-form RNS limbs [level][N] (not yet NTT'd)
void launch_encode(cudaStream_t s, const NttTables *T, const EncoderTables &E, int logN, const double *vals, int len,
                   int level, double scale, double2 *work, unsigned long long *maxbits, u64 *out);
struct DecodeTables { // per level, device
  const u64 *punct;   // [l][l]  Q/q_j
  const u64 *invp;    // [l]     (Q/q_j)^-1 mod q_j
  const u64 *Q;       // [l]
  const u64 *half;    // [l]     (Q+1)/2
};
// coeff: [l][N] coefficient-form canonical limbs -> out: N/2 doubles (device)
void launch_decode(cudaStream_t s, const NttTables *T, const EncoderTables &E, const DecodeTables &D, int logN, int l,
                   const u64 *coeff, double scale, double2 *work, double *out, unsigned long long *maxbits = nullptr);
