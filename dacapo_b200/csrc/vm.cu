// libB200_HEVM.so -- the HEVM interpreter on B200 and its C ABI.
//
// Drop-in for the reference runtime (reference: lib/Runtime/SEAL_HEVM.cpp:15-504): same 18
// exported symbols (include/hevm_abi.h), same .hevm / .cst files, same opcode semantics
// including the wrapper quirks of SURVEY.md A.3.  Ciphertext / plaintext registers, the
// secret / public / relinearisation / Galois keys and every table live in HBM for the life of
// the VM; host<->device traffic happens only in load/preprocess (constants), encrypt (input
// slots) and decrypt (output slots) -- nothing crosses during run().
//
// There is no CPU execution path: without a CUDA device every entry point aborts.
#include "host_params.hpp"
#include "kernels.h"
#include <cmath>
#include <cuda_profiler_api.h>
#include <nvtx3/nvToolsExt.h> // header-only NVTX v3: one range per HEVM opcode class
#include <complex>
#include <cstring>
#include <fstream>
#include <functional>
#include <algorithm>
#include <map>
#include <random>
#include <set>
#include <string>
#include <vector>

namespace {

[[noreturn]] void die(const char *msg) {
  std::fprintf(stderr, "[b200-hevm] fatal: %s\n", msg);
  std::abort();
}

// ---- on-disk formats ---------------------------------------------------------------------
// HEVM container (reference: include/hecate/Support/HEVMHeader.h:9-34; writer EmitHEVM.cpp:31-119)
#pragma pack(push, 1)
struct HevmHead {
  uint32_t magic, head_size;
  uint64_t n_args, n_res;
};
struct HevmConfig {
  uint64_t body_len, n_ops, n_ct, n_pt, init_level;
};
struct HevmOp {
  uint16_t opcode, dst, lhs, rhs;
};
#pragma pack(pop)
struct ParamFile { // <dir>/hevm_params.bin : ring geometry + 256-bit key seed (own format, version 2; SURVEY 8b "Files consumed")
  uint64_t magic, logN, L, bits;
  uint32_t seed[8];
};
const uint64_t PARAM_MAGIC = 0x3230324D56454842ull;

// sampler stream ids (specification shared with the oracle)
inline u64 ksk_stream(u64 key_id, u64 digit, u64 kind, u64 limb) { return (((key_id * 64 + digit) * 2 + kind) * 64 + limb) + (16ull << 20); }
inline u64 enc_stream(u64 counter, u64 which) { return (1ull << 40) + counter * 4 + which; }

struct CtReg {
  u64 *d = nullptr; // [2][L-1][N] (poly pitch fixed), NTT form
  int level = 0;
  double scale = 1.0;
};
struct PtReg {
  u64 *d = nullptr; // [level][N]
  int level = 0, cap = 0;
  double scale = 1.0;
};

template <class T> T *dalloc(size_t n) {
  T *p = nullptr;
  CUDA_CHECK(cudaMalloc(&p, n * sizeof(T)));
  return p;
}
template <class T> T *upload(const std::vector<T> &v) {
  T *p = dalloc<T>(v.size());
  CUDA_CHECK(cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  return p;
}

struct DecTabs {
  u64 *punct, *invp, *Q, *half;
};

// Everything one in-flight operation needs privately: a stream, the NTT scratch, the encoder /
// encryptor temporaries.  run() schedules independent HEVM ops onto several lanes.
struct Lane {
  cudaStream_t stream = nullptr;
  GpuLauncher la;
  OpsIface *ops = nullptr;
  double2 *d_work = nullptr;
  unsigned long long *d_maxbits = nullptr;
  double *d_vals = nullptr, *d_vals_in = nullptr;
  u64 *d_u = nullptr, *d_e = nullptr, *d_ue = nullptr, *d_tmpct = nullptr, *d_coef = nullptr;
  PtReg boot_pt;
  double load = 0; // scheduling estimate
  // asynchronous encrypt() calls issued on this lane (C ABI): private counter word, pinned staging copy of the caller's
  // values, completion event; `async_pending` = lane 0 has not been ordered behind that work yet (VM::join_async)
  u64 *d_ctr = nullptr;
  double *h_stage = nullptr;
  cudaEvent_t ev_async = nullptr;
  bool async_pending = false, stage_busy = false;
};

struct VM {
  hp::HostParams P;
  int logN = 0, L = 0;
  size_t N = 0, pitch = 0;
  // Key material comes from the 256-bit seed of the parameter file through a keyed ChaCha20 PRF; the encryption
  // randomness (u, e0, e1 of every encrypt / bootstrap) uses a separate key that is FRESH PER VM (std::random_device)
  // -- two processes sharing a key directory never encrypt with the same randomness -- unless a parity test pins it
  // to the key seed with hevmx_set_enc_counter.
  Seed256 seed{}, enc_seed{};
  u64 enc_counter = 0;
  // Limb-sharded KEY STORAGE across the GPUs of one node ("each GPU holds only its limbs of every key"):
  // with HEVM_SHARD_RANK / HEVM_SHARD_WORLD set, rank g keeps only the limbs [own_lo, own_hi) of the L key-level limbs
  // of every key-switch key (static, contiguous, balanced; the special limb L-1 belongs to the last rank) and the VM
  // only accepts the limb-sharded ops (hevmx_ks_shard_p2p).  At level l its key-switch targets are that range cut to l.
  int shard_rank = 0, shard_world = 1, own_lo = 0, own_hi = 0;
  // balanced contiguous ranges with the remainder on the FIRST ranks: the last rank, which also carries the special
  // limb's extra work (inverse NTT + rounding of two rows, and every other rank waits for them), never has more limbs
  // than anybody else
  static void own_range(int L, int g, int G, int &lo, int &hi) {
    const int base = L / G, rem = L % G;
    lo = g * base + std::min(g, rem);
    hi = lo + base + (g < rem ? 1 : 0);
  }
  bool keys_sharded() const { return shard_world > 1; }
  int key_limbs() const { return keys_sharded() ? own_hi - own_lo : L; }
  void my_targets(int l, int &tlo, int &thi) const {
    tlo = std::min(own_lo, l);
    thi = (own_hi == L) ? l + 1 : std::min(own_hi, l);
  }
  // peer-to-peer exchange block of the sharded key switch (one cudaMalloc, exported over CUDA IPC):
  //   digits t[2][L][N] | rounded rows rnd[2][2][N] | flags[2][8] u64 | done counter      ([2] = double buffer by epoch parity)
  struct P2P {
    bool on = false;
    int rank = 0, world = 1;
    u64 *block = nullptr;
    PeerPtrs peer{};
    unsigned long long epoch = 0;
    cudaEvent_t ev[6] = {};
    bool timed = false;
    // copy-engine push (HEVM_P2P_PUSH=2): the per-peer copies run on side streams, each followed by that peer's flag
    cudaStream_t side[4] = {};
    cudaEvent_t fork = nullptr, copied[2][4] = {};
    bool copied_valid[2] = {false, false};
  } p2p;
  // push `words` from `src` (in this rank's block, offset dst_off) into every peer's block + raise flag slot `flag_off`
  void p2p_push(const u64 *src, size_t words, size_t dst_off, size_t flag_off, unsigned long long e, int par) {
    Lane &L0 = lanes[0];
    unsigned *done = reinterpret_cast<unsigned *>(p2p.block + p2p_words() - 2);
    static const int mode = std::getenv("HEVM_P2P_PUSH") ? std::atoi(std::getenv("HEVM_P2P_PUSH")) : 1;
    if (mode != 2 || p2p.world == 1) {
      launch_p2p_push(L0.stream, src, words, p2p.peer, dst_off, p2p.rank, p2p.world, done, flag_off, e);
      return;
    }
    if (!p2p.fork) {
      CUDA_CHECK(cudaEventCreateWithFlags(&p2p.fork, cudaEventDisableTiming));
      for (auto &st : p2p.side) CUDA_CHECK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
      for (auto &row : p2p.copied)
        for (auto &ev : row) CUDA_CHECK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    }
    CUDA_CHECK(cudaEventRecord(p2p.fork, L0.stream));
    for (int i = 0; i < 4; i++) CUDA_CHECK(cudaStreamWaitEvent(p2p.side[i], p2p.fork, 0));
    for (int k = 0; k + 1 < p2p.world; k++) {
      const int g = (p2p.rank + 1 + k) % p2p.world;
      cudaStream_t st = p2p.side[k % 4];
      if (words) CUDA_CHECK(cudaMemcpyAsync(p2p.peer.p[g] + dst_off, src, words * 8, cudaMemcpyDeviceToDevice, st));
      launch_p2p_flag(st, p2p.peer.p[g] + flag_off + p2p.rank, e);
    }
    for (int i = 0; i < 4; i++) CUDA_CHECK(cudaEventRecord(p2p.copied[par][i], p2p.side[i]));
    p2p.copied_valid[par] = true;
  }
  size_t p2p_t_off(int parity) const { return (size_t)parity * L * N; }
  size_t p2p_rnd_off(int parity) const { return (size_t)2 * L * N + (size_t)parity * 2 * N; }
  size_t p2p_flag_off(int slot) const { return (size_t)2 * L * N + 4 * N + (size_t)slot * 8; }
  size_t p2p_words() const { return (size_t)2 * L * N + 4 * N + 16 + 2; }
  u64 *d_ctr_base = nullptr; // device copy of the encryption counter base (read by the samplers)
  std::vector<Lane> lanes;
  Lane *ln = nullptr; // lane the host code is currently issuing on
  NttTables *dT = nullptr;
  // encoder
  EncoderTables E{};
  std::map<int, DecTabs> dec;
  // keys
  u64 *d_sk = nullptr, *d_pk = nullptr, *d_relin = nullptr;
  std::map<u64, u64 *> d_gal;
  // run() schedule / CUDA graph
  cudaGraphExec_t graph_exec = nullptr;
  int graph_nboot = 0;
  unsigned long long graph_launches = 0;
  // register renaming: logical ciphertext register -> physical buffer.  `home[r]` is the identity buffer of
  // register r (where encrypt / ct_write put data and where every run() starts); `spare` are extra buffers.
  std::vector<u64 *> home, spare, final_map;
  // register metadata (level, scale) of every ciphertext register when the graph was captured: `entry_meta` is what the
  // captured kernels were specialised for (a replay is only valid from the same entry state of the registers the
  // program reads before writing), `final_meta` is what a run leaves behind (restored after every replay)
  std::vector<std::pair<int, double>> entry_meta, final_meta;
  std::vector<char> live_in;
  bool use_graph = true;
  // program
  std::vector<std::vector<double>> consts;
  HevmHead head{};
  HevmConfig cfg{};
  std::vector<HevmOp> prog;
  std::vector<uint64_t> arg_scale, arg_level, res_scale, res_level, res_dst;
  std::vector<CtReg> ct;
  std::vector<PtReg> pt;
  bool debug = false;
  cudaEvent_t ev_start = nullptr, ev_stop = nullptr;

  // ---------------------------------------------------------------- setup
  u64 *new_scratch() { // one NTT scratch area; the fused kernels' completion counters at its end start out (and stay) zero
    u64 *p = dalloc<u64>(Scratch::words(L, N));
    CUDA_CHECK(cudaMemset(p + Scratch::words(L, N) - KS_CTL_WORDS, 0, KS_CTL_WORDS * sizeof(u64)));
    return p;
  }
  void init(const ParamFile &pf) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) die("no CUDA device visible: libB200_HEVM.so has no CPU path");
    if (const char *e = std::getenv("HEVM_DEVICE")) CUDA_CHECK(cudaSetDevice(std::atoi(e)));
    if (pf.logN < 14 || pf.logN > 17) die("supported ring sizes: N = 2^14 .. 2^17 (HEVM_LOGN = 14..17)");
    if (pf.L < 2 || pf.L > HEVM_MAXL) die("number of primes out of range");
    logN = (int)pf.logN, L = (int)pf.L, N = (size_t)1 << logN;
    std::memcpy(seed.k, pf.seed, sizeof seed.k);
    {
      std::random_device rd;
      for (auto &w : enc_seed.k) w = rd();
    }
    pitch = (size_t)(L - 1) * N;
    if (const char *e = std::getenv("HEVM_SHARD_WORLD")) {
      shard_world = std::max(1, std::atoi(e));
      shard_rank = std::getenv("HEVM_SHARD_RANK") ? std::atoi(std::getenv("HEVM_SHARD_RANK")) : 0;
      if (shard_world > 8 || shard_rank < 0 || shard_rank >= shard_world) die("HEVM_SHARD_RANK / HEVM_SHARD_WORLD out of range (at most 8 ranks)");
      if (2 * shard_world > (int)pf.L) die("limb-sharded key storage needs at least two limbs per rank");
    }
    own_range(L, shard_rank, shard_world, own_lo, own_hi);
    P.build(logN, L, (int)pf.bits);
    Tw *d_tw = upload(P.tw), *d_itw = upload(P.itw);
    P.tab.tw = d_tw, P.tab.itw = d_itw;
    P.build_rowwise();
    P.tab.twB = upload(P.twB), P.tab.itwB = upload(P.itwB);
    P.twB.clear(), P.twB.shrink_to_fit(), P.itwB.clear(), P.itwB.shrink_to_fit();
    dT = dalloc<NttTables>(1);
    CUDA_CHECK(cudaMemcpy(dT, &P.tab, sizeof(NttTables), cudaMemcpyHostToDevice));
    d_ctr_base = dalloc<u64>(1);
    CUDA_CHECK(cudaMemset(d_ctr_base, 0, 8));
    int nl = 16;
    if (const char *e = std::getenv("HEVM_STREAMS")) nl = std::max(1, std::min(32, std::atoi(e)));
    if (const char *e = std::getenv("HEVM_GRAPH")) use_graph = std::atoi(e) != 0;
    lanes.resize(nl);
    const size_t slots = N / 2;
    for (Lane &l : lanes) {
      CUDA_CHECK(cudaStreamCreateWithFlags(&l.stream, cudaStreamNonBlocking));
      l.la.stream = l.stream;
      l.ops = make_ops(l.la, dT, logN, L);
      l.ops->sc.carve(new_scratch(), L, N);
      if (keys_sharded()) l.ops->key_L = key_limbs(), l.ops->key_t0 = own_lo;
      if (const char *e = std::getenv("HEVM_GROUP_WARPS")) l.ops->group_warps = std::max(1, std::atoi(e));
      l.d_work = dalloc<double2>(N);
      l.d_maxbits = dalloc<unsigned long long>(1);
      l.d_vals = dalloc<double>(slots);
      l.d_vals_in = dalloc<double>(slots);
      l.d_u = dalloc<u64>((size_t)L * N);
      l.d_e = dalloc<u64>((size_t)2 * L * N);
      l.d_ue = dalloc<u64>((size_t)3 * L * N); // encryption randomness u | e0 | e1, contiguous
      l.d_tmpct = dalloc<u64>((size_t)2 * L * N);
      l.d_coef = dalloc<u64>((size_t)L * N);
      l.boot_pt.d = dalloc<u64>((size_t)(L - 1) * N); // never reallocated (graph capture forbids cudaMalloc)
      l.boot_pt.cap = L - 1;
      l.d_ctr = dalloc<u64>(1);
      CUDA_CHECK(cudaMallocHost(&l.h_stage, slots * sizeof(double)));
      CUDA_CHECK(cudaEventCreateWithFlags(&l.ev_async, cudaEventDisableTiming));
    }
    CUDA_CHECK(cudaEventCreateWithFlags(&ev_front, cudaEventDisableTiming));
    ln = &lanes[0];
    init_encoder();
    for (int l = 1; l <= L - 1; l++) dec_tabs(l);
    CUDA_CHECK(cudaDeviceSynchronize()); // the table uploads / memsets above ran on the legacy stream; the lanes do not wait for it
    keygen();
    CUDA_CHECK(cudaStreamSynchronize(ln->stream));
  }

  // CKKSEncoder tables (SEAL ckks.cpp constructor): slot index map through powers of 3,
  // complex 2N-th roots from the first octant (std::polar) + 8-fold symmetry.
  void init_encoder() {
    const size_t m = 2 * N, slots = N / 2;
    std::vector<u32> idx(N);
    u64 pos = 1;
    for (size_t i = 0; i < slots; i++, pos = (pos * 3) & (m - 1)) {
      idx[i] = hp::revbits((u32)((pos - 1) >> 1), logN);
      idx[slots + i] = hp::revbits((u32)((m - pos - 1) >> 1), logN);
    }
    const double PI = 3.1415926535897932384626433832795028842;
    std::vector<std::complex<double>> oct(m / 8 + 1);
    for (size_t i = 0; i <= m / 8; i++) oct[i] = std::polar<double>(1.0, 2 * PI * (double)i / (double)m);
    // ComplexRoots::get_root: first octant table + 8-fold symmetry (same case order as SEAL's util/croots.cpp)
    std::function<std::complex<double>(size_t)> root = [&](size_t k) -> std::complex<double> {
      k &= m - 1;
      if (k <= m / 8) return oct[k];
      if (k <= m / 4) return std::complex<double>(oct[m / 4 - k].imag(), oct[m / 4 - k].real());
      if (k <= m / 2) return -std::conj(root(m / 2 - k));
      if (k <= 3 * m / 4) return -root(k - m / 2);
      return std::conj(root(m - k));
    };
    std::vector<double2> fr(N), ir(N);
    fr[0] = ir[0] = make_double2(0, 0);
    for (size_t i = 1; i < N; i++) {
      auto f = root(hp::revbits((u32)i, logN));
      auto b = std::conj(root((size_t)hp::revbits((u32)(i - 1), logN) + 1));
      fr[i] = make_double2(f.real(), f.imag());
      ir[i] = make_double2(b.real(), b.imag());
    }
    E.slot_index = upload(idx);
    E.fwd_root = upload(fr);
    E.inv_root = upload(ir);
  }

  // per-level CRT tables for decode (SEAL RNSBase punctured products; multi-precision, little endian)
  const DecTabs &dec_tabs(int l) {
    auto it = dec.find(l);
    if (it != dec.end()) return it->second;
    auto mulw = [](std::vector<u64> &a, u64 w) {
      u64 c = 0;
      for (auto &x : a) {
        unsigned __int128 p = (unsigned __int128)x * w + c;
        x = (u64)p, c = (u64)(p >> 64);
      }
    };
    std::vector<u64> Q(l, 0), punct((size_t)l * l, 0), invp(l), half(l);
    Q[0] = 1;
    for (int j = 0; j < l; j++) mulw(Q, P.q[j]);
    for (int j = 0; j < l; j++) {
      std::vector<u64> a(l, 0);
      a[0] = 1;
      u64 r = 1;
      for (int k = 0; k < l; k++)
        if (k != j) mulw(a, P.q[k]), r = hp::mul(r, P.q[k] % P.q[j], P.q[j]);
      std::copy(a.begin(), a.end(), punct.begin() + (size_t)j * l);
      invp[j] = hp::inv(r, P.q[j]);
    }
    { // (Q+1) >> 1
      std::vector<u64> t(Q);
      u64 c = 1;
      for (int k = 0; k < l && c; k++) c = (++t[k] == 0);
      for (int k = 0; k < l; k++) half[k] = (t[k] >> 1) | ((k + 1 < l ? t[k + 1] : c) << 63);
    }
    DecTabs d{upload(punct), upload(invp), upload(Q), upload(half)};
    return dec[l] = d;
  }

  // ---------------------------------------------------------------- keys (on device, from the seed)
  void enc_zero_sym(u64 a_stream_base, u64 e_stream, u64 *c0, u64 *c1, const u64 *newkey, int digit) {
    launch_sample_uniform(ln->stream, dT, logN, c1, L, seed, a_stream_base); // a: sampled directly in NTT form
    launch_sample_cbd(ln->stream, dT, logN, ln->d_e, L, seed, e_stream);
    ln->ops->ntt_fwd(ln->d_e, ln->d_e, L, 0, 1);
    launch_ksk_finish(ln->stream, dT, logN, L, c0, c1, d_sk, ln->d_e, newkey, digit);
  }
  u64 *ksk_tmp = nullptr; // [2][L][N]: one digit of a key, before its owned limbs are kept (sharded storage only)
  u64 *make_ksk(u64 key_id, const u64 *newkey) {
    const size_t KL = (size_t)key_limbs();
    u64 *key = dalloc<u64>((size_t)(L - 1) * 2 * KL * N);
    if (keys_sharded() && !ksk_tmp) ksk_tmp = dalloc<u64>((size_t)2 * L * N);
    for (int J = 0; J < L - 1; J++) {
      u64 *c0 = key + ((size_t)J * 2 + 0) * KL * N, *c1 = key + ((size_t)J * 2 + 1) * KL * N;
      if (!keys_sharded()) {
        enc_zero_sym(ksk_stream(key_id, J, 0, 0), ksk_stream(key_id, J, 1, 0), c0, c1, newkey, J);
      } else { // the digit is generated whole (its randomness is defined over all limbs), then only the owned limbs are stored
        u64 *t0 = ksk_tmp, *t1 = ksk_tmp + (size_t)L * N;
        enc_zero_sym(ksk_stream(key_id, J, 0, 0), ksk_stream(key_id, J, 1, 0), t0, t1, newkey, J);
        CUDA_CHECK(cudaMemcpyAsync(c0, t0 + (size_t)own_lo * N, KL * N * 8, cudaMemcpyDeviceToDevice, ln->stream));
        CUDA_CHECK(cudaMemcpyAsync(c1, t1 + (size_t)own_lo * N, KL * N * 8, cudaMemcpyDeviceToDevice, ln->stream));
      }
    }
    launch_key_split(ln->stream, key, (size_t)(L - 1) * 2 * KL * N); // storage format of the key inner product
    return key;
  }
  u64 galois_elt_from_step(int step) const {
    const u64 m = 2 * N;
    if (step == 0) return m - 1;
    const u64 pos = (u64)std::abs(step);
    if (pos >= N / 2) die("rotation step count too large");
    u64 s = step < 0 ? N / 2 - pos : pos, elt = 1;
    for (u64 i = 0; i < s; i++) elt = (elt * 3) & (m - 1);
    return elt;
  }
  void keygen() {
    d_sk = dalloc<u64>((size_t)L * N);
    launch_sample_ternary(ln->stream, dT, logN, d_sk, L, seed, 1ull << 20);
    ln->ops->ntt_fwd(d_sk, d_sk, L, 0, 1);
    d_pk = dalloc<u64>((size_t)2 * L * N);
    enc_zero_sym(2ull << 20, 3ull << 20, d_pk, d_pk + (size_t)L * N, nullptr, -1);
    u64 *nk = dalloc<u64>((size_t)L * N);
    launch_square(ln->stream, dT, logN, L, nk, d_sk);
    d_relin = make_ksk(0, nk);
    // default Galois key set (SEAL GaloisTool::get_elts_all): conjugation + 3^(+-2^i)
    const u64 m = 2 * N;
    std::vector<u64> elts{m - 1};
    u64 posp = 3, negp = 0;
    for (u64 x = 1; x < m; x += 2)
      if (((x * 3) & (m - 1)) == 1) {
        negp = x;
        break;
      }
    for (int i = 0; i < logN - 1; i++) {
      elts.push_back(posp), posp = (posp * posp) & (m - 1);
      elts.push_back(negp), negp = (negp * negp) & (m - 1);
    }
    // HEVM_GALOIS_STEPS="1,-2,64": only these rotation steps get a key (SEAL KeyGenerator::create_galois_keys(steps));
    // unset = the default set of create_galois_keys() that the reference uses (SEAL_HEVM.cpp:82-83)
    if (const char *e = std::getenv("HEVM_GALOIS_STEPS")) {
      elts.clear();
      for (const char *p = e; *p;) {
        char *end = nullptr;
        const long st = std::strtol(p, &end, 10);
        if (end == p) break;
        elts.push_back(galois_elt_from_step((int)st));
        p = (*end == ',') ? end + 1 : end;
      }
    }
    for (u64 elt : elts) {
      if (d_gal.count(elt)) continue;
      launch_galois_gather(ln->stream, logN, L, nk, d_sk, (u32)elt);
      d_gal[elt] = make_ksk(1 + ((elt - 1) >> 1), nk);
    }
    CUDA_CHECK(cudaStreamSynchronize(ln->stream));
    CUDA_CHECK(cudaFree(nk));
  }

  // ---------------------------------------------------------------- registers
  CtReg &ctr(size_t r) {
    if (r >= ct.size()) die("ciphertext register index out of range");
    CtReg &c = ct[r];
    if (!c.d) c.d = dalloc<u64>(2 * pitch);
    return c;
  }
  PtReg &ptr(size_t r) {
    if (r >= pt.size()) die("plaintext register index out of range");
    return pt[r];
  }
  void pt_reserve(PtReg &p, int level) {
    if (p.cap < level) {
      if (p.d) CUDA_CHECK(cudaFree(p.d));
      p.d = dalloc<u64>((size_t)level * N);
      p.cap = level;
    }
  }
  void resize_regs(size_t nct, size_t npt) {
    invalidate_graph();
    eager_run_done = false;
    for (size_t r = 0; r < home.size(); r++) ct[r].d = home[r]; // so that each home buffer is freed exactly once
    home.clear();
    final_map.clear();
    for (auto &c : ct)
      if (c.d) CUDA_CHECK(cudaFree(c.d));
    for (auto &p : pt)
      if (p.d) CUDA_CHECK(cudaFree(p.d));
    ct.clear();
    ct.reserve(nct + 1);
    ct.resize(nct);
    pt.assign(npt, PtReg());
  }

  // ---------------------------------------------------------------- encode / encrypt / decrypt
  // SEAL_HEVM.cpp:256-267 : tile to N/2 slots, encode at 2^scale_bits, keep `level` limbs
  void encode_internal(PtReg &dst, const double *d_src, int len, int64_t level, int64_t scale_bits) {
    if (level < 1 || level > L - 1) die("encode: level out of range");
    pt_reserve(dst, (int)level);
    dst.level = (int)level;
    dst.scale = std::pow(2.0, (double)scale_bits);
    launch_encode(ln->stream, dT, E, logN, d_src, len, (int)level, dst.scale, ln->d_work, ln->d_maxbits, dst.d);
    ln->ops->ntt_fwd(dst.d, dst.d, (int)level, 0, 1);
  }
  void stage_host_values(const double *h, size_t len) {
    if (len == 0) die("empty value vector");
    CUDA_CHECK(cudaMemcpyAsync(ln->d_vals_in, h, std::min(len, N / 2) * sizeof(double), cudaMemcpyHostToDevice, ln->stream));
  }
  // Encryptor::encrypt (public key): zero-encrypt on level+1 limbs, divide-and-round by the extra
  // prime, add the plaintext.  Reference call sites SEAL_HEVM.cpp:333,444.
  // The sampler stream id is enc_stream(*d_ctr_base + k, which): `k` is static (position of the
  // encryption inside run()), the base lives on the device so that graph replays advance it.
  void encrypt_pt(const PtReg &p, CtReg &out, u64 k, const u64 *ctr_base = nullptr) {
    if (!ctr_base) ctr_base = d_ctr_base;
    const int l = p.level, nl = l + 1;
    const u64 counter = k;
    // u, e0, e1 live back to back in d_ue ([3][nl][N]) so that one launch pair transforms all three
    u64 *u = ln->d_ue, *e01 = ln->d_ue + (size_t)nl * N;
    launch_sample_enc(ln->stream, dT, logN, u, nl, enc_seed, enc_stream(counter, 0), ctr_base); // streams +0 (u), +1, +2 (e0, e1)
    if (3 * nl <= L * (L - 1)) {
      ln->ops->ntt_fwd(u, u, 3 * nl, 0, 1, nl);
    } else {
      for (int j = 0; j < 3; j++) ln->ops->ntt_fwd(u + (size_t)j * nl * N, u + (size_t)j * nl * N, nl, 0, 1);
    }
    launch_enc_combine(ln->stream, dT, logN, nl, ln->d_tmpct, u, d_pk, (size_t)L * N, e01);
    ln->ops->rescale(ln->d_tmpct, (size_t)nl * N, out.d, pitch, nl, p.d); // ... + plaintext, in the same epilogue
    out.level = l;
    out.scale = p.scale;
  }
  // eager single encryption (encrypt() ABI call, hevmx hook): publish the counter, use offset 0
  void encrypt_pt_now(const PtReg &p, CtReg &out) {
    // the counter word is the issuing lane's own: encrypt() calls on different lanes run concurrently
    CUDA_CHECK(cudaMemcpyAsync(ln->d_ctr, &enc_counter, 8, cudaMemcpyHostToDevice, ln->stream));
    encrypt_pt(p, out, 0, ln->d_ctr);
    enc_counter++;
  }
  // ---- asynchronous API work (encrypt() of independent arguments, decrypt_result() of all results) -----------------
  // encrypt(i) is issued on lane i mod #lanes and returns without waiting; every other entry point first orders lane 0
  // behind that work (join_async, called by V()).  decrypt_result() decrypts ALL result registers at its first call
  // after a run -- one per lane, concurrently -- into pinned host memory and serves the later calls from there.
  cudaEvent_t ev_front = nullptr; // "everything issued on lane 0 so far"
  void join_async() {
    for (Lane &l : lanes)
      if (l.async_pending) {
        CUDA_CHECK(cudaStreamWaitEvent(lanes[0].stream, l.ev_async, 0));
        l.async_pending = false;
      }
    ln = &lanes[0];
  }
  void encrypt_async(size_t i, const double *dat, size_t len) {
    Lane &l = lanes[i % lanes.size()];
    const size_t n = std::min(len, N / 2);
    if (len == 0) die("empty value vector");
    if (l.stage_busy) CUDA_CHECK(cudaEventSynchronize(l.ev_async)); // the lane's previous encrypt still owns the staging buffer
    std::memcpy(l.h_stage, dat, n * sizeof(double));                // the caller may reuse `dat` as soon as we return
    if (&l != &lanes[0]) { // ordered behind whatever lane 0 was given so far (it may still read or write the register)
      CUDA_CHECK(cudaEventRecord(ev_front, lanes[0].stream));
      CUDA_CHECK(cudaStreamWaitEvent(l.stream, ev_front, 0));
    }
    ln = &l;
    CUDA_CHECK(cudaMemcpyAsync(l.d_vals_in, l.h_stage, n * sizeof(double), cudaMemcpyHostToDevice, l.stream));
    encode_internal(l.boot_pt, l.d_vals_in, (int)n, (int64_t)arg_level[i], (int64_t)arg_scale[i]);
    CtReg &c = ctr(i);
    rehome(c);
    encrypt_pt_now(l.boot_pt, c);
    CUDA_CHECK(cudaEventRecord(l.ev_async, l.stream));
    l.async_pending = &l != &lanes[0];
    l.stage_busy = true;
    ln = &lanes[0];
  }
  struct ResCache {
    bool valid = false;
    double *d = nullptr, *h = nullptr; // [n_res][N/2] device / pinned host
    size_t cap = 0;
    std::vector<cudaEvent_t> ev;
    std::vector<char> ok;
  } res_cache;
  void decrypt_all_results() {
    const size_t nres = res_dst.size(), slots = N / 2;
    ResCache &rc = res_cache;
    if (rc.cap < nres) {
      if (rc.d) {
        CUDA_CHECK(cudaFree(rc.d));
        CUDA_CHECK(cudaFreeHost(rc.h));
      }
      rc.d = dalloc<double>(nres * slots);
      CUDA_CHECK(cudaMallocHost(&rc.h, nres * slots * sizeof(double)));
      rc.cap = nres;
      while (rc.ev.size() < nres) {
        cudaEvent_t e;
        CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        rc.ev.push_back(e);
      }
    }
    rc.ok.assign(nres, 0);
    CUDA_CHECK(cudaEventRecord(ev_front, lanes[0].stream));
    for (size_t r = 0; r < nres; r++) {
      if (res_dst[r] >= ct.size() || ct[res_dst[r]].level < 1 || !ct[res_dst[r]].d) continue; // decrypt() reports it
      Lane &l = lanes[r % lanes.size()];
      if (&l != &lanes[0]) CUDA_CHECK(cudaStreamWaitEvent(l.stream, ev_front, 0));
      ln = &l;
      decrypt_to_pt(ct[res_dst[r]], l.boot_pt);
      decode_pt(l.boot_pt, rc.d + r * slots);
      CUDA_CHECK(cudaMemcpyAsync(rc.h + r * slots, rc.d + r * slots, slots * sizeof(double), cudaMemcpyDeviceToHost, l.stream));
      CUDA_CHECK(cudaEventRecord(rc.ev[r], l.stream));
      rc.ok[r] = 1;
    }
    ln = &lanes[0];
    // lane 0 must not run ahead of the readers (a later run() or encrypt() overwrites the registers): wait for the last
    // event of every other lane that was given work
    for (size_t k = 1; k < lanes.size(); k++) {
      long last = -1;
      for (size_t r = k; r < nres; r += lanes.size())
        if (rc.ok[r]) last = (long)r;
      if (last >= 0) CUDA_CHECK(cudaStreamWaitEvent(lanes[0].stream, rc.ev[(size_t)last], 0));
    }
    rc.valid = true;
  }
  void decrypt_to_pt(const CtReg &c, PtReg &p) {
    pt_reserve(p, c.level);
    p.level = c.level;
    p.scale = c.scale;
    launch_decrypt(ln->stream, dT, logN, c.level, p.d, c.d, pitch, d_sk);
  }
  void decode_pt(const PtReg &p, double *d_out) {
    ln->ops->ntt_inv(p.d, ln->d_coef, p.level, 0, 1);
    const DecTabs &t = dec_tabs(p.level);
    DecodeTables D{t.punct, t.invp, t.Q, t.half};
    launch_decode(ln->stream, dT, E, D, logN, p.level, ln->d_coef, p.scale, ln->d_work, d_out, ln->d_maxbits);
  }

  // ---------------------------------------------------------------- the opcodes (SEAL_HEVM.cpp:269-334)
  static std::vector<int> naf(int value) {
    std::vector<int> res;
    const bool sign = value < 0;
    value = std::abs(value);
    for (int i = 0; value; i++) {
      int zi = (value & 1) ? 2 - (value & 3) : 0;
      value = (value - zi) >> 1;
      if (zi) res.push_back((sign ? -zi : zi) * (1 << i));
    }
    return res;
  }
  void rotate_steps(int steps, std::vector<int> &out) const { // Evaluator::rotate_internal
    if (steps == 0) return;
    if (d_gal.count(galois_elt_from_step(steps))) {
      out.push_back(steps);
      return;
    }
    auto terms = naf(steps);
    if (terms.size() == 1) die("Galois key not present");
    for (int t : terms)
      if ((size_t)std::abs(t) != N / 2) rotate_steps(t, out);
  }
  void copy_ct(const CtReg &s, CtReg &d) {
    if (s.d != d.d) launch_elementwise(ln->stream, EW_COPY, dT, logN, d.d, s.d, nullptr, nullptr, pitch, s.level);
    d.level = s.level, d.scale = s.scale;
  }
  u64 boot_index = 0; // encryptions issued so far inside the current run()
  double shard_scale = 0; // hevmx_mulcc_shard_stage: result scale computed at stage 1
  bool meta_only = false; // exec() of addcc / mulcp updates register metadata only (their kernels were fused by the scheduler)
  struct NvtxRange { // host-side range around the issue of one op (visible in Nsight Systems; free when no tool is attached)
    explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
  };
  static const char *op_class_name(unsigned opcode) {
    static const char *const names[11] = {"hevm.encode", "hevm.rotate", "hevm.negate", "hevm.rescale", "hevm.modswitch", "hevm.upscale",
                                          "hevm.addcc", "hevm.addcp", "hevm.mulcc", "hevm.mulcp", "hevm.bootstrap"};
    return opcode < 11 ? names[opcode] : "hevm.placeholder";
  }
  void exec(const HevmOp &op) {
    NvtxRange nvtx(op_class_name(op.opcode));
    switch (op.opcode) {
    case 1: { // rotate
      CtReg &s = ctr(op.lhs), &d = ctr(op.dst);
      if (s.level < 1) die("rotate: empty source register");
      if (keys_sharded() && (int16_t)op.rhs != 0) die("this VM stores limb-sharded keys (HEVM_SHARD_WORLD): use the sharded ops");
      std::vector<int> steps;
      rotate_steps((int16_t)op.rhs, steps);
      if (steps.empty()) {
        copy_ct(s, d);
        break;
      }
      const u64 *cur = s.d;
      for (int st : steps) {
        const u64 elt = galois_elt_from_step(st);
        ln->ops->keyswitch(LD_GALOIS, cur, nullptr, d.d, pitch, s.level, d_gal.at(elt), (u32)elt);
        cur = d.d;
      }
      d.level = s.level, d.scale = s.scale;
      break;
    }
    case 2: { // negate
      CtReg &s = ctr(op.lhs), &d = ctr(op.dst);
      launch_elementwise(ln->stream, EW_NEG, dT, logN, d.d, s.d, nullptr, nullptr, pitch, s.level);
      d.level = s.level, d.scale = s.scale;
      break;
    }
    case 3: { // rescale
      CtReg &s = ctr(op.lhs), &d = ctr(op.dst);
      if (s.level < 2) die("rescale: already at the last level");
      ln->ops->rescale(s.d, pitch, d.d, pitch, s.level);
      const int l = s.level;
      const double sc = s.scale / (double)P.q[l - 1];
      d.level = l - 1, d.scale = sc;
      break;
    }
    case 4: { // modswitch by downFactor limbs
      const int down = (int16_t)op.rhs;
      if (down <= 0) break; // reference: nothing happens (SEAL_HEVM.cpp:288-292)
      CtReg &s = ctr(op.lhs), &d = ctr(op.dst);
      if (s.level - down < 1) die("modswitch: already at the last level");
      const int nl = s.level - down;
      const double sc = s.scale;
      if (s.d != d.d) launch_elementwise(ln->stream, EW_COPY, dT, logN, d.d, s.d, nullptr, nullptr, pitch, nl);
      d.level = nl, d.scale = sc;
      break;
    }
    case 5: die("This VM does not support native upscale op");
    case 6: { // addcc
      CtReg &a = ctr(op.lhs), &b = ctr(op.rhs), &d = ctr(op.dst);
      a.scale = b.scale; // SEAL_HEVM.cpp:301
      if (a.level != b.level) die("addcc: level mismatch");
      const int l = a.level;
      const double sc = a.scale;
      if (!meta_only) launch_elementwise(ln->stream, EW_ADD, dT, logN, d.d, a.d, b.d, nullptr, pitch, l);
      d.level = l, d.scale = sc;
      break;
    }
    case 7: { // addcp
      CtReg &a = ctr(op.lhs), &d = ctr(op.dst);
      PtReg &p = ptr(op.rhs);
      a.scale = p.scale; // SEAL_HEVM.cpp:308
      if (a.level != p.level) die("addcp: level mismatch");
      const int l = a.level;
      const double sc = a.scale;
      launch_elementwise(ln->stream, EW_ADDP, dT, logN, d.d, a.d, nullptr, p.d, pitch, l);
      d.level = l, d.scale = sc;
      break;
    }
    case 8: { // mulcc = multiply + relinearize, one fused key-switch pipeline
      CtReg &a = ctr(op.lhs), &b = ctr(op.rhs), &d = ctr(op.dst);
      if (a.level != b.level || a.level < 1) die("mulcc: level mismatch");
      if (keys_sharded()) die("this VM stores limb-sharded keys (HEVM_SHARD_WORLD): use the sharded ops");
      const int l = a.level;
      const double sc = a.scale * b.scale;
      ln->ops->keyswitch(LD_PRODUCT, a.d, b.d, d.d, pitch, l, d_relin, 0);
      d.level = l, d.scale = sc;
      break;
    }
    case 9: { // mulcp
      CtReg &a = ctr(op.lhs), &d = ctr(op.dst);
      PtReg &p = ptr(op.rhs);
      if (a.level != p.level) die("mulcp: level mismatch");
      const int l = a.level;
      const double sc = a.scale * p.scale;
      if (!meta_only) launch_elementwise(ln->stream, EW_MULP, dT, logN, d.d, a.d, nullptr, p.d, pitch, l);
      d.level = l, d.scale = sc;
      break;
    }
    case 10: { // "bootstrap" = decrypt + re-encrypt at the target level, entirely on device
      CtReg &s = ctr(op.lhs), &d = ctr(op.dst);
      // decrypt (c0 + c1 s) fused into the inverse NTT's load; the N/2 decoded values stay in the lane's FFT buffer,
      // already scattered for the encoder
      ln->ops->decrypt_inv(s.d, pitch, d_sk, ln->d_coef, s.level);
      {
        const DecTabs &t = dec_tabs(s.level);
        DecodeTables D{t.punct, t.invp, t.Q, t.half};
        launch_decode(ln->stream, dT, E, D, logN, s.level, ln->d_coef, s.scale, ln->d_work, nullptr, ln->d_maxbits);
      }
      const int64_t sb = (int64_t)std::log2(s.scale); // SEAL_HEVM.cpp:332 truncation
      encode_internal(ln->boot_pt, nullptr, (int)(N / 2), op.rhs, sb);
      encrypt_pt(ln->boot_pt, d, boot_index++);
      break;
    }
    default: break; // 0 = encode (done in preprocess), 0xFFFF = tensor.empty placeholder
    }
  }

  // ---------------------------------------------------------------- limb-sharded key switch over peer memory
  // One call = the whole sharded rotation / multiply+relinearise of this rank, no host synchronisation:
  //   stage 1 (own data limbs -> coefficient digits, written into this rank's exchange block)
  //   push the own digit rows into every peer's block over NVLink + epoch flag ; wait for every peer's flag
  //   stage 2 (mod-up + key inner product for the own targets; the special limb's owner also rounds it)
  //   owner of the special limb pushes the two rounded rows + flag ; the others wait for it
  //   stage 3 (mod-down of the own data limbs)
  // The exchange buffers are double-buffered by the parity of the key-switch sequence number: a rank that runs ahead
  // writes buffer (n+1) & 1 while a slower peer may still read buffer n & 1; it cannot get two key switches ahead
  // because each one needs every peer's digits.
  // `phase` 0 = the whole op; 1 / 2 / 3 = only up to the first push / from the first wait to the second push or wait /
  // from there to the end -- a test that emulates several ranks with several VMs on ONE GPU issues the phases in
  // lockstep, so that no wait kernel is ever queued in front of the kernels it waits for.
  void ks_shard_p2p(int mode, CtReg &d, const CtReg &a, const CtReg *b, const u64 *key, u32 elt, int phase = 0) {
    if (!p2p.on) die("hevmx_p2p_setup has not been called");
    Lane &L0 = lanes[0];
    ln = &L0;
    const int l = a.level;
    int tlo, thi;
    my_targets(l, tlo, thi);
    const int dhi = std::min(thi, l), nd = std::max(0, dhi - tlo);
    const bool own_sp = thi == l + 1;
    int sp_owner = p2p.world - 1;
    if (phase <= 1) ++p2p.epoch;
    const unsigned long long e = p2p.epoch;
    const int par = (int)(e & 1);
    if (phase <= 1 && p2p.copied_valid[par])  // copy-engine pushes of the key switch before last read this parity's buffers
      for (int i = 0; i < 4; i++) CUDA_CHECK(cudaStreamWaitEvent(L0.stream, p2p.copied[par][i], 0));
    Scratch &sc = L0.ops->sc;
    u64 *save_t = sc.t, *save_rnd = sc.rnd;
    sc.t = p2p.block + p2p_t_off(par), sc.rnd = p2p.block + p2p_rnd_off(par);
    auto mark = [&](int i) {
      if (p2p.timed) CUDA_CHECK(cudaEventRecord(p2p.ev[i], L0.stream));
    };
    const u64 *bd = b ? b->d : nullptr;
    if (phase <= 1) {
      mark(0);
      L0.ops->ks_shard_stage(1, mode, a.d, bd, d.d, pitch, l, key, elt, tlo, thi);
      mark(1);
      p2p_push(sc.t + (size_t)tlo * N, (size_t)nd * N, p2p_t_off(par) + (size_t)tlo * N, p2p_flag_off(0), e, par);
    }
    if (phase == 0 || phase == 2) {
      mark(2);
      // HEVM_P2P_OVERLAP=1: mod-up pass A source rank by source rank, own digits first -- a peer's digits are waited for
      // only when their turn comes, so the exchange overlaps the pass-A work on what has already arrived.  Measured on
      // 8 B200s (N = 2^16, level 29) it LOSES: eight quarter-wave launches cost 224 us against 176 us for one launch
      // behind one wait, so the default is to wait for every peer and transform all digits in one launch.
      static const int overlap_mode = std::getenv("HEVM_P2P_OVERLAP") ? std::atoi(std::getenv("HEVM_P2P_OVERLAP")) : 0;
      const bool overlap = overlap_mode == 1;
      if (overlap_mode == 2 && nd > 0) {
        // HEVM_P2P_OVERLAP=2: only the OWN digits (already local) are transformed while the push is in flight, then one
        // wait for every peer, then the digits below and above the own range
        L0.ops->shard_j0 = tlo, L0.ops->shard_nj = nd;
        L0.ops->ks_shard_stage(20, mode, a.d, bd, d.d, pitch, l, key, elt, tlo, thi);
        launch_p2p_wait(L0.stream, p2p.block, p2p_flag_off(0), p2p.rank, p2p.world, -1, e);
        if (tlo > 0) {
          L0.ops->shard_j0 = 0, L0.ops->shard_nj = tlo;
          L0.ops->ks_shard_stage(20, mode, a.d, bd, d.d, pitch, l, key, elt, tlo, thi);
        }
        if (dhi < l) {
          L0.ops->shard_j0 = dhi, L0.ops->shard_nj = l - dhi;
          L0.ops->ks_shard_stage(20, mode, a.d, bd, d.d, pitch, l, key, elt, tlo, thi);
        }
      } else if (!overlap) {
        launch_p2p_wait(L0.stream, p2p.block, p2p_flag_off(0), p2p.rank, p2p.world, -1, e);
        L0.ops->shard_j0 = 0, L0.ops->shard_nj = l;
        L0.ops->ks_shard_stage(20, mode, a.d, bd, d.d, pitch, l, key, elt, tlo, thi);
      }
      for (int k = 0; overlap && k < p2p.world; k++) {
        const int r = (p2p.rank + k) % p2p.world;
        int rlo, rhi;
        own_range(L, r, p2p.world, rlo, rhi);
        const int j0 = std::min(rlo, l), j1 = std::min(rhi, l);
        if (j1 <= j0) continue;
        if (r != p2p.rank) launch_p2p_wait(L0.stream, p2p.block, p2p_flag_off(0), p2p.rank, p2p.world, r, e);
        L0.ops->shard_j0 = j0, L0.ops->shard_nj = j1 - j0;
        L0.ops->ks_shard_stage(20, mode, a.d, bd, d.d, pitch, l, key, elt, tlo, thi);
      }
      L0.ops->ks_shard_stage(21, mode, a.d, bd, d.d, pitch, l, key, elt, tlo, thi);
      mark(3);
      if (own_sp) p2p_push(sc.rnd, (size_t)2 * N, p2p_rnd_off(par), p2p_flag_off(1), e, par);
    }
    if (phase == 0 || phase == 3) {
      if (!own_sp) launch_p2p_wait(L0.stream, p2p.block, p2p_flag_off(1), p2p.rank, p2p.world, sp_owner, e);
      mark(4);
      L0.ops->ks_shard_stage(3, mode, a.d, bd, d.d, pitch, l, key, elt, tlo, thi);
      mark(5);
    }
    sc.t = save_t, sc.rnd = save_rnd;
  }

  // ---------------------------------------------------------------- batched ops (independent ciphertexts, one launch)
  // n independent ops of one opcode whose sources sit at the same level: rotate (single key-switch steps), mulcc and
  // rescale go out as ONE persistent kernel over all n ciphertexts (ks_fused.cuh; BASELINE configs[1] "batched
  // ciphertexts"); everything else, and any batch that does not qualify, is issued op by op.
  std::vector<u64 *> batch_scratch;
  u64 *scratch_for(size_t k) {
    while (batch_scratch.size() <= k) batch_scratch.push_back(new_scratch());
    return batch_scratch[k];
  }
  void exec_batch(int opcode, int n, const int64_t *dst, const int64_t *lhs, const int64_t *rhs) {
    bool ok = (opcode == 1 || opcode == 3 || opcode == 8) && n > 1;
    int lvl = 0;
    std::vector<KsCt> items;
    for (int k = 0; ok && k < n; k++) {
      CtReg &a = ctr((size_t)lhs[k]), &d = ctr((size_t)dst[k]);
      if (k == 0) lvl = a.level;
      if (a.level != lvl || lvl < (opcode == 3 ? 2 : 1)) ok = false;
      for (int j = 0; ok && j < k; j++) // not independent: an item writes what another item reads or writes
        if (dst[j] == dst[k] || dst[j] == lhs[k] || dst[k] == lhs[j] || (opcode == 8 && (dst[j] == rhs[k] || dst[k] == rhs[j]))) ok = false;
      if (!ok) break;
      KsCt it{a.d, nullptr, d.d, nullptr, scratch_for((size_t)k), 0, 0};
      if (opcode == 1) {
        std::vector<int> steps;
        rotate_steps((int16_t)rhs[k], steps);
        if (steps.size() != 1) {
          ok = false;
          break;
        }
        const u64 elt = galois_elt_from_step(steps[0]);
        it.key = d_gal.at(elt), it.elt = (u32)elt;
      } else if (opcode == 8) {
        CtReg &b = ctr((size_t)rhs[k]);
        if (b.level != lvl) ok = false;
        it.b = b.d, it.key = d_relin;
      }
      items.push_back(it);
    }
    if (!ok) {
      for (int k = 0; k < n; k++) exec(HevmOp{(uint16_t)opcode, (uint16_t)dst[k], (uint16_t)lhs[k], (uint16_t)rhs[k]});
      return;
    }
    if (opcode == 3)
      ln->ops->rescale_batch(n, items.data(), pitch, pitch, lvl);
    else
      ln->ops->keyswitch_batch(opcode == 1 ? LD_GALOIS : LD_PRODUCT, n, items.data(), pitch, lvl);
    for (int k = 0; k < n; k++) { // register metadata, as exec() would have left it
      CtReg &a = ctr((size_t)lhs[k]), &d = ctr((size_t)dst[k]);
      if (opcode == 1) d.level = lvl, d.scale = a.scale;
      if (opcode == 3) d.scale = a.scale / (double)P.q[lvl - 1], d.level = lvl - 1;
      if (opcode == 8) d.scale = a.scale * ctr((size_t)rhs[k]).scale, d.level = lvl;
    }
  }

  // ---------------------------------------------------------------- run(): multi-lane schedule + CUDA graph
  // The program is static, so run() is issued once as a dependency-respecting schedule over the lanes
  // (independent ciphertext ops overlap on the GPU) while being captured into a CUDA graph; later
  // run() calls replay the graph.  Register metadata (level, scale) is host state and is advanced in
  // program order while issuing; a replay leaves it in the same final state.
  std::vector<cudaEvent_t> ev_pool;
  size_t ev_used = 0;
  void reserve_events(size_t n) {
    while (ev_pool.size() < n) {
      cudaEvent_t e;
      CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      ev_pool.push_back(e);
    }
  }
  cudaEvent_t new_event() {
    if (ev_used == ev_pool.size()) {
      cudaEvent_t e;
      CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      ev_pool.push_back(e);
    }
    return ev_pool[ev_used++];
  }
  static double op_cost(const HevmOp &op, int level) {
    switch (op.opcode) {
    // microseconds measured on one B200 at N = 2^15 (profiled_B200_GPU.json): key switch 51 us at level 3 .. 148 us at 13
    case 1: case 8: return 30.0 + 7.0 * level + 0.16 * level * level;
    case 3: return 14.0 + 1.1 * level;
    case 10: return 178.0 + 4.0 * level;
    default: return 3.5 + 0.25 * level;
    }
  }
  // registers the program reads before (or without) writing them: their level / scale at entry shapes the schedule
  void compute_live_in() {
    live_in.assign(ct.size(), 0);
    std::vector<char> written(ct.size(), 0);
    auto rd = [&](unsigned r) {
      if (r < ct.size() && !written[r]) live_in[r] = 1;
    };
    for (const HevmOp &o : prog) {
      const bool ct_op = ((o.opcode >= 1 && o.opcode <= 4) || (o.opcode >= 6 && o.opcode <= 10)) && !(o.opcode == 4 && (int16_t)o.rhs <= 0);
      if (!ct_op) continue;
      rd(o.lhs);
      if (o.opcode == 6 || o.opcode == 8) rd(o.rhs);
      if (o.dst < ct.size()) written[o.dst] = 1;
    }
  }
  bool eager_run_done = false; // this program has been run once without a graph (see run_program)
  static bool lazy_graph() {
    static const bool on = !(std::getenv("HEVM_GRAPH_LAZY") && std::atoi(std::getenv("HEVM_GRAPH_LAZY")) == 0);
    return on;
  }
  void invalidate_graph() {
    if (graph_exec) CUDA_CHECK(cudaGraphExecDestroy(graph_exec));
    graph_exec = nullptr;
  }
  void preallocate_registers() {
    for (size_t r = 0; r < ct.size(); r++) ctr(r);
    if (home.size() != ct.size()) {
      home.resize(ct.size());
      for (size_t r = 0; r < ct.size(); r++) home[r] = ct[r].d;
      int nspare = 128; // 0.9 GB at N = 2^15: enough for every false dependency of the ResNet-20 program to disappear
      if (const char *e = std::getenv("HEVM_RENAME_POOL")) nspare = std::max(0, std::atoi(e));
      if (lanes.size() == 1) nspare = 0; // nothing to overlap
      while ((int)spare.size() < nspare) spare.push_back(dalloc<u64>(2 * pitch));
    }
  }
  void reset_mapping() { // every run() starts from the identity mapping (so that a captured graph can be replayed)
    for (size_t r = 0; r < home.size(); r++) ct[r].d = home[r];
  }
  void rehome(CtReg &c) { // data written through the API always lands in the register's home buffer
    const size_t r = &c - ct.data();
    if (r < home.size()) c.d = home[r];
  }
  // issue every op of the program over the lanes with event dependencies
  // While a whole program is scheduled over the lanes the machine is shared by many independent ops, so the fused pass-A
  // launches are cut into as few target groups as possible (every extra group recomputes the inverse pass: measured
  // +10 % ops/s on the op microbench, -4 % ResNet-20 latency against the single-op setting); a lone op issued through
  // hevmx_exec keeps the latency-oriented split.  HEVM_GROUP_WARPS_RUN overrides.
  struct GroupWarpsScope {
    VM &vm;
    std::vector<int> saved;
    explicit GroupWarpsScope(VM &v) : vm(v) {
      static const int run_warps = std::getenv("HEVM_GROUP_WARPS_RUN") ? std::max(1, std::atoi(std::getenv("HEVM_GROUP_WARPS_RUN"))) : 148;
      for (Lane &l : vm.lanes) saved.push_back(l.ops->group_warps), l.ops->group_warps = std::min(l.ops->group_warps, run_warps);
    }
    ~GroupWarpsScope() {
      for (size_t i = 0; i < vm.lanes.size(); i++) vm.lanes[i].ops->group_warps = saved[i];
    }
  };
  void issue_scheduled() {
    GroupWarpsScope group_scope(*this);
    const int nl = (int)lanes.size();
    struct BufState { // dependency state of one PHYSICAL buffer
      cudaEvent_t wr = nullptr;
      int wr_lane = -1;
      double ready = 0; // modelled time at which the last write lands
      std::vector<std::pair<int, cudaEvent_t>> readers;
    };
    std::map<u64 *, BufState> bs;
    // FIFO of free physical buffers (oldest first).  Only SPARE buffers circulate: a register's home buffer is where
    // encrypt() / hevmx_ct_write put data between runs, so it must never end up backing another register.
    std::vector<u64 *> pool(spare.begin(), spare.end());
    const std::set<u64 *> home_set(home.begin(), home.end());
    auto is_home = [&](u64 *b) { return home_set.count(b) != 0; };
    size_t pool_head = 0;
    reset_mapping();
    for (Lane &l : lanes) l.load = 0;
    ev_used = 0;
    boot_index = 0;
    // fork: every lane starts after lane 0's current position
    cudaEvent_t fork = new_event();
    CUDA_CHECK(cudaEventRecord(fork, lanes[0].stream));
    for (int i = 1; i < nl; i++) CUDA_CHECK(cudaStreamWaitEvent(lanes[i].stream, fork, 0));
    // peephole: mulcp t <- x * p immediately followed by addcc d <- t + y (either operand order) where t is dead
    // afterwards becomes ONE kernel d <- x * p + y (same canonical residues; register metadata of both ops is
    // still applied in order).  Liveness is conservative: every register is live at the end of the program.
    std::vector<char> fuse(prog.size(), 0), acc_dead(prog.size(), 0);
    {
      static const bool on = !(std::getenv("HEVM_FUSE") && std::atoi(std::getenv("HEVM_FUSE")) == 0);
      std::vector<char> live(ct.size(), 1);
      for (size_t i = prog.size(); on && i-- > 0;) {
        const HevmOp &o = prog[i];
        // (modswitch with downFactor <= 0 leaves dst untouched, SEAL_HEVM.cpp:288-292: not a write, not a read)
        const bool writes = ((o.opcode >= 1 && o.opcode <= 4) || (o.opcode >= 6 && o.opcode <= 10)) && !(o.opcode == 4 && (int16_t)o.rhs <= 0);
        if (!writes || o.dst >= ct.size()) continue;
        if (i > 0 && o.opcode == 6 && prog[i - 1].opcode == 9 && o.lhs < ct.size() && o.rhs < ct.size()) {
          const HevmOp &m = prog[i - 1];
          const int t = m.dst;
          const bool t_once = (o.lhs == t) != (o.rhs == t);
          const bool t_dead = (o.dst == t) || !live[t]; // liveness after the addcc
          if (t_once && t_dead && m.lhs < ct.size()) {
            fuse[i - 1] = 1;
            const int y = (o.lhs == t) ? o.rhs : o.lhs; // the other addend: dead after this pair?
            acc_dead[i - 1] = (o.dst == y) || !live[y];
          }
        }
        live[o.dst] = 0;
        live[o.lhs < ct.size() ? o.lhs : o.dst] = 1;
        if (o.opcode == 6 || o.opcode == 8) live[o.rhs < ct.size() ? o.rhs : o.dst] = 1;
      }
    }
    // chain_cont[pc] (pc = a fused pair with result register D): the next op that touches D is another fused pair that
    // takes D as its addend, does not multiply D itself, and after which D is dead -- the running sum then never has
    // to exist in D
    std::vector<char> chain_cont(prog.size(), 0);
    for (size_t pc = 0; pc + 1 < prog.size(); pc++) {
      if (!fuse[pc]) continue;
      const int D = prog[pc + 1].dst;
      for (size_t j = pc + 2; j < prog.size() && j < pc + 2 + 96; j++) {
        const HevmOp &o = prog[j];
        const bool ct_op = ((o.opcode >= 1 && o.opcode <= 4) || (o.opcode >= 6 && o.opcode <= 10)) && !(o.opcode == 4 && (int16_t)o.rhs <= 0);
        if (!ct_op) continue;
        const bool two = o.opcode == 6 || o.opcode == 8;
        if (fuse[j]) {
          const HevmOp &a2 = prog[j + 1];
          const int t2 = o.dst, y2 = (a2.lhs == t2) ? a2.rhs : a2.lhs;
          if (o.lhs == D || t2 == D) break;                       // reads / clobbers D
          if (y2 == D) {
            chain_cont[pc] = acc_dead[j];
            break;
          }
          if (a2.dst == D) break;
          j++; // skip the addcc of that pair
          continue;
        }
        if (o.dst == D || o.lhs == D || (two && o.rhs == D)) break;
      }
    }
    const bool chains_on = !spare.empty() && !(std::getenv("HEVM_CHAIN") && std::atoi(std::getenv("HEVM_CHAIN")) == 0);
    struct Pending {
      bool active = false;
      int accreg = -1, lvl = 0;
      u64 *acc_buf = nullptr;
      std::vector<u64 *> xs;
      std::vector<const u64 *> ps;
    } pend;
    auto pinned = [&](const u64 *b) {
      if (!pend.active) return false;
      if (b == pend.acc_buf) return true;
      for (u64 *x : pend.xs)
        if (x == b) return true;
      return false;
    };
    auto flush_pending = [&]() {
      if (!pend.active) return;
      const int wr = pend.accreg, lvl = pend.lvl;
      std::vector<u64 *> srcs(pend.xs);
      srcs.push_back(pend.acc_buf);
      double dep_ready = 0, min_start = 1e300;
      for (u64 *b : srcs) dep_ready = std::max(dep_ready, bs[b].ready);
      for (int i = 0; i < nl; i++) min_start = std::min(min_start, std::max(lanes[i].load, dep_ready));
      int best = -1;
      for (size_t k = 0; k < srcs.size() && best < 0; k++) {
        const int pl = bs[srcs[k]].wr_lane;
        if (pl >= 0 && std::max(lanes[pl].load, dep_ready) <= min_start + 2.0) best = pl;
      }
      if (best < 0)
        for (int i = 0; i < nl; i++) {
          if (std::max(lanes[i].load, dep_ready) > min_start + 1e-9) continue;
          if (best < 0 || lanes[i].load > lanes[best].load) best = i;
        }
      const double best_start = std::max(lanes[best].load, dep_ready);
      Lane &L0 = lanes[best];
      auto wait_on = [&](int lane, cudaEvent_t e) {
        if (e && lane != best) CUDA_CHECK(cudaStreamWaitEvent(L0.stream, e, 0));
      };
      u64 *old_buf = ct[wr].d, *dst_buf = old_buf;
      if (pool_head < pool.size()) {
        dst_buf = pool[pool_head++];
        if (!is_home(old_buf)) pool.push_back(old_buf);
      }
      for (u64 *b : srcs) wait_on(bs[b].wr_lane, bs[b].wr);
      BufState &D = bs[dst_buf];
      wait_on(D.wr_lane, D.wr);
      for (auto &r : D.readers) wait_on(r.first, r.second);
      ln = &L0;
      if (pend.xs.size() == 1) {
        launch_elementwise(L0.stream, EW_MULP_ADD, dT, logN, dst_buf, pend.xs[0], pend.acc_buf, pend.ps[0], pitch, lvl);
      } else {
        MulpTerms mt{};
        mt.n = (int)pend.xs.size();
        for (int k = 0; k < mt.n; k++) mt.x[k] = pend.xs[k], mt.p[k] = pend.ps[k];
        launch_mulp_add_n(L0.stream, dT, logN, dst_buf, pend.acc_buf, mt, pitch, lvl);
      }
      ct[wr].d = dst_buf;
      cudaEvent_t done = new_event();
      CUDA_CHECK(cudaEventRecord(done, L0.stream));
      for (u64 *b : srcs)
        if (b != dst_buf) bs[b].readers.emplace_back(best, done);
      BufState &D2 = bs[dst_buf];
      D2.wr = done, D2.wr_lane = best, D2.readers.clear();
      L0.load = best_start + 4.0 + 1.5 * (double)pend.xs.size();
      D2.ready = L0.load;
      pend.active = false, pend.xs.clear(), pend.ps.clear();
    };
    for (size_t pc = 0; pc < prog.size(); pc++) {
      const HevmOp &op = prog[pc];
      if (fuse[pc]) {
        const HevmOp &m = op, &ad = prog[pc + 1];
        const int t = m.dst, y = (ad.lhs == t) ? ad.rhs : ad.lhs, wr = ad.dst;
        const int lvl = ct[m.lhs].level;
        if (ct[y].level != lvl) die("addcc: level mismatch");
        // accumulation chains s1 = s0 + x1*p1; s2 = s1 + x2*p2; ... (convolution taps): the terms are collected and issued
        // as ONE kernel when the chain ends -- nothing else reads the intermediate sums (chain_cont), the captured source
        // buffers stay pinned until then
        const bool extend = pend.active && pend.accreg == y && pend.lvl == lvl && pend.xs.size() < 8;
        if (!extend) flush_pending();
        if (!pend.active) pend.active = true, pend.acc_buf = ct[y].d, pend.lvl = lvl;
        pend.accreg = wr; // the register that holds the running sum from here on
        pend.xs.push_back(ct[m.lhs].d);
        pend.ps.push_back(ptr(m.rhs).d);
        meta_only = true; // both ops' level / scale bookkeeping, in program order (SEAL_HEVM.cpp:297-323)
        exec(m);
        exec(ad);
        meta_only = false;
        if (!(chain_cont[pc] && chains_on) || pend.xs.size() >= 8) flush_pending();
        pc++;
        continue;
      }
      int rd[2], nrd = 0, wr = -1;
      switch (op.opcode) {
      case 1: case 2: case 3: case 10: rd[nrd++] = op.lhs, wr = op.dst; break;
      case 4:
        if ((int16_t)op.rhs > 0) rd[nrd++] = op.lhs, wr = op.dst;
        break;
      case 6: case 8: rd[nrd++] = op.lhs, rd[nrd++] = op.rhs, wr = op.dst; break;
      case 7: case 9: rd[nrd++] = op.lhs, wr = op.dst; break;
      case 5: die("This VM does not support native upscale op");
      default: break;
      }
      if (wr < 0) continue;
      for (int k = 0; k < nrd; k++)
        if ((size_t)rd[k] >= ct.size()) die("ciphertext register index out of range");
      if ((size_t)wr >= ct.size()) die("ciphertext register index out of range");
      if (pend.active) {
        bool touch = wr == pend.accreg || (pool_head < pool.size() && pinned(pool[pool_head]));
        for (int k = 0; k < nrd; k++) touch |= rd[k] == pend.accreg;
        if (touch) flush_pending();
      }
      if (op.opcode == 4 && op.lhs == op.dst) { // in-place limb drop: metadata only, no kernel, no dependency
        exec(op);
        continue;
      }
      // lane choice = list scheduling on a modelled clock: an op can start when its sources are ready and the lane
      // is free (lane.load = time at which the lane drains); take the lane with the earliest start, preferring
      // on ties the producer of a source (no cross-stream wait) and otherwise the lane that has been idle longest
      // -- an op never queues behind unrelated work that is itself still waiting for its inputs
      u64 *src_buf[2] = {ct[rd[0]].d, nrd > 1 ? ct[rd[1]].d : nullptr};
      double dep_ready = 0;
      for (int k = 0; k < nrd; k++) dep_ready = std::max(dep_ready, bs[src_buf[k]].ready);
      double min_start = 1e300;
      for (int i = 0; i < nl; i++) min_start = std::min(min_start, std::max(lanes[i].load, dep_ready));
      int best = -1;
      for (int k = 0; k < nrd && best < 0; k++) { // a producer lane that is (almost) as early as any
        const int pl = bs[src_buf[k]].wr_lane;
        if (pl >= 0 && std::max(lanes[pl].load, dep_ready) <= min_start + 2.0) best = pl;
      }
      if (best < 0)
        for (int i = 0; i < nl; i++) { // best fit: the busiest lane among those that can start at min_start
          if (std::max(lanes[i].load, dep_ready) > min_start + 1e-9) continue;
          if (best < 0 || lanes[i].load > lanes[best].load) best = i;
        }
      const double best_start = std::max(lanes[best].load, dep_ready);
      Lane &L0 = lanes[best];
      auto wait_on = [&](int lane, cudaEvent_t e) {
        if (e && lane != best) CUDA_CHECK(cudaStreamWaitEvent(L0.stream, e, 0));
      };
      // renaming: the destination gets a fresh physical buffer; the old one returns to the FIFO, so the
      // write-after-read / write-after-write hazards of recycled registers (ReuseBuffer.cpp:27-55) disappear
      u64 *old_buf = ct[wr].d, *dst_buf = old_buf;
      if (pool_head < pool.size()) {
        dst_buf = pool[pool_head++];
        if (!is_home(old_buf)) pool.push_back(old_buf);
      }
      for (int k = 0; k < nrd; k++) wait_on(bs[src_buf[k]].wr_lane, bs[src_buf[k]].wr); // RAW
      BufState &D = bs[dst_buf];
      wait_on(D.wr_lane, D.wr);                                                        // WAW
      for (auto &r : D.readers) wait_on(r.first, r.second);                            // WAR
      ln = &L0;
      const int lvl = ct[rd[0]].level;
      // run the op out of place into dst_buf: temporarily point the destination register at it
      if (dst_buf != old_buf) {
        if (wr == rd[0] || (nrd > 1 && wr == rd[1])) {
          // source and destination are the same logical register: read the old buffer through a shadow register
          CtReg shadow = ct[wr];
          ct.push_back(shadow); // index ct.size()-1 (capacity reserved below, references stay valid)
          HevmOp o2 = op;
          if (o2.lhs == wr) o2.lhs = (uint16_t)(ct.size() - 1);
          if (nrd > 1 && o2.rhs == wr) o2.rhs = (uint16_t)(ct.size() - 1);
          ct[wr].d = dst_buf;
          exec(o2);
          ct.pop_back();
        } else {
          ct[wr].d = dst_buf;
          exec(op);
        }
      } else {
        exec(op);
      }
      cudaEvent_t done = new_event();
      CUDA_CHECK(cudaEventRecord(done, L0.stream));
      for (int k = 0; k < nrd; k++)
        if (src_buf[k] != dst_buf) bs[src_buf[k]].readers.emplace_back(best, done);
      BufState &D2 = bs[dst_buf];
      D2.wr = done, D2.wr_lane = best, D2.readers.clear();
      L0.load = best_start + op_cost(op, lvl);
      D2.ready = L0.load;
    }
    flush_pending();
    final_map.resize(home.size());
    for (size_t r = 0; r < home.size(); r++) final_map[r] = ct[r].d;
    if (std::getenv("HEVM_SCHED_DEBUG")) {
      std::fprintf(stderr, "[b200-hevm] schedule: %zu events, lane loads:", ev_used);
      for (Lane &l : lanes) std::fprintf(stderr, " %.0f", l.load);
      std::fprintf(stderr, "\n");
    }
    // join: lane 0 waits for every other lane
    for (int i = 1; i < nl; i++) {
      cudaEvent_t e = new_event();
      CUDA_CHECK(cudaEventRecord(e, lanes[i].stream));
      CUDA_CHECK(cudaStreamWaitEvent(lanes[0].stream, e, 0));
    }
    ln = &lanes[0];
  }
  void run_program() {
    ln = &lanes[0];
    preallocate_registers();
    CUDA_CHECK(cudaMemcpyAsync(d_ctr_base, &enc_counter, 8, cudaMemcpyHostToDevice, lanes[0].stream));
    if (debug) { // setDebug(true): the reference's per-op trace (SEAL_HEVM.cpp:340-350, 270-271), ops in program order
      boot_index = 0;
      reset_mapping();
      long i = (long)((head.head_size + cfg.body_len) / 8), j = 0;
      for (auto &op : prog) {
        std::printf("\n%lo %ld\nopcode [%u], dst [%u], lhs [%u], rhs [%u]\n", i++, j++, (unsigned)op.opcode, (unsigned)op.dst, (unsigned)op.lhs,
                    (unsigned)op.rhs);
        if (op.opcode == 1 && op.lhs < ct.size()) std::printf("%g\n", std::log2(ct[op.lhs].scale));
        exec(op);
      }
      std::fflush(stdout);
      enc_counter += boot_index;
    } else if (g_prof.on || lanes.size() == 1 && !use_graph) { // kernel-class profiling / plain mode: in order on lane 0
      boot_index = 0;
      reset_mapping();
      for (auto &op : prog) exec(op);
      enc_counter += boot_index;
    } else if (!use_graph) {
      issue_scheduled();
      enc_counter += boot_index;
    } else if (!graph_exec && !eager_run_done && lazy_graph()) {
      // FIRST run() of a program: issue the schedule on the lanes directly.  Capturing + instantiating a graph of tens of
      // thousands of nodes costs 2-3x the run itself, and the reference's hc-test flow is ONE cold run per process
      // (examples/tests/ResNet.py:109-111); the graph is built by the second run() and replayed from the third on.
      issue_scheduled();
      enc_counter += boot_index;
      eager_run_done = true;
    } else {
      if (graph_exec) { // the captured kernels are specialised for the entry levels of the live-in registers
        for (size_t r = 0; r < ct.size() && graph_exec; r++)
          if (live_in[r] && (ct[r].level != entry_meta[r].first || ct[r].scale != entry_meta[r].second)) invalidate_graph();
      }
      if (!graph_exec) {
        compute_live_in();
        entry_meta.resize(ct.size());
        for (size_t r = 0; r < ct.size(); r++) entry_meta[r] = {ct[r].level, ct[r].scale};
        cudaGraph_t g = nullptr;
        // HEVM_GRAPH_PDL=1: keep programmatic dependent launch while capturing (programmatic edges inside the graph)
        static const bool graph_pdl = std::getenv("HEVM_GRAPH_PDL") && std::atoi(std::getenv("HEVM_GRAPH_PDL")) != 0;
        g_pdl_suspended = !graph_pdl;
        const unsigned long long c0 = g_launch_count;
        CUDA_CHECK(cudaStreamBeginCapture(lanes[0].stream, cudaStreamCaptureModeRelaxed));
        issue_scheduled();
        CUDA_CHECK(cudaStreamEndCapture(lanes[0].stream, &g));
        g_pdl_suspended = false;
        graph_launches = g_launch_count - c0; // kernels recorded, not yet executed
        g_launch_count = c0;
        CUDA_CHECK(cudaGraphInstantiate(&graph_exec, g, 0));
        CUDA_CHECK(cudaGraphDestroy(g));
        graph_nboot = (int)boot_index;
        final_meta.resize(ct.size());
        for (size_t r = 0; r < ct.size(); r++) final_meta[r] = {ct[r].level, ct[r].scale};
      }
      CUDA_CHECK(cudaGraphLaunch(graph_exec, lanes[0].stream));
      for (size_t r = 0; r < final_map.size(); r++) { // where (and in which state) the replay leaves each register
        ct[r].d = final_map[r];
        ct[r].level = final_meta[r].first, ct[r].scale = final_meta[r].second;
      }
      g_launch_count += graph_launches; // every replay executes all recorded kernels
      enc_counter += graph_nboot;
    }
    CUDA_CHECK(cudaStreamSynchronize(lanes[0].stream));
  }
};

// Every entry point that touches VM state goes through V(): lane 0 is ordered behind the asynchronous encrypt() calls
// and the cached results of decrypt_result() are dropped.  Vq() = pure queries / the decrypt_result path itself.
VM *Vq(void *h) { return static_cast<VM *>(h); }
VM *V(void *h) {
  VM *vm = static_cast<VM *>(h);
  vm->join_async();
  vm->res_cache.valid = false;
  return vm;
}

void read_params(const char *dir, ParamFile &pf) {
  std::ifstream f(std::string(dir) + "/hevm_params.bin", std::ios::binary);
  if (!f) die("cannot open <dir>/hevm_params.bin (call create_context first)");
  f.read((char *)&pf, sizeof pf);
  if (pf.magic != PARAM_MAGIC) die("bad hevm_params.bin");
}

} // namespace

extern "C" {

// ================= the reference's 18 symbols (include/hevm_abi.h) =========================
void create_context(char *dir) {
  ParamFile pf{};
  pf.magic = PARAM_MAGIC;
  const char *e;
  pf.logN = (e = std::getenv("HEVM_LOGN")) ? std::strtoull(e, nullptr, 10) : 15;
  pf.L = (e = std::getenv("HEVM_NUM_PRIMES")) ? std::strtoull(e, nullptr, 10) : 14;
  pf.bits = (e = std::getenv("HEVM_PRIME_BITS")) ? std::strtoull(e, nullptr, 10) : 60;
  if ((e = std::getenv("HEVM_SEED"))) { // tests: a fixed, public seed
    const uint64_t v = std::strtoull(e, nullptr, 0);
    pf.seed[0] = (uint32_t)v, pf.seed[1] = (uint32_t)(v >> 32);
  } else {
    std::random_device rd;
    for (auto &w : pf.seed) w = rd();
  }
  std::ofstream f(std::string(dir) + "/hevm_params.bin", std::ios::binary);
  if (!f) die("create_context: cannot write hevm_params.bin");
  f.write((const char *)&pf, sizeof pf);
}
void *initFullVM(char *dir, bool /*device: always the GPU*/) {
  ParamFile pf;
  read_params(dir, pf);
  auto vm = new VM();
  vm->init(pf);
  return vm;
}
void *initClientVM(char *dir) { return initFullVM(dir, true); }
void *initServerVM(char *dir) { return initFullVM(dir, true); }

void load(void *h, char *cst, char *hevm) {
  VM *vm = V(h);
  {
    std::ifstream f(cst, std::ios::binary);
    if (!f) die("cannot open constants file");
    int64_t n = 0;
    f.read((char *)&n, 8);
    vm->consts.assign((size_t)n, {});
    for (auto &c : vm->consts) {
      int64_t len = 0;
      f.read((char *)&len, 8);
      c.resize((size_t)len);
      f.read((char *)c.data(), len * 8);
    }
  }
  std::ifstream f(hevm, std::ios::binary);
  if (!f) die("cannot open hevm file");
  f.read((char *)&vm->head, sizeof(HevmHead));
  if (vm->head.magic != 0x4845564D) die("bad HEVM magic");
  f.read((char *)&vm->cfg, sizeof(HevmConfig));
  auto rd = [&](std::vector<uint64_t> &v, size_t n) {
    v.resize(n);
    f.read((char *)v.data(), n * 8);
  };
  rd(vm->arg_scale, vm->head.n_args), rd(vm->arg_level, vm->head.n_args);
  rd(vm->res_scale, vm->head.n_res), rd(vm->res_level, vm->head.n_res), rd(vm->res_dst, vm->head.n_res);
  vm->prog.resize(vm->cfg.n_ops);
  f.read((char *)vm->prog.data(), vm->prog.size() * sizeof(HevmOp));
  vm->resize_regs(vm->cfg.n_ct, vm->cfg.n_pt);
}
void loadClient(void *, void *) { die("client/server split is vestigial in the reference (SURVEY A.3); use initFullVM"); }

void preprocess(void *h) { // SEAL_HEVM.cpp:242-254: encode every opcode-0 constant into its plaintext register
  VM *vm = V(h);
  static const double one = 1.0;
  for (auto &op : vm->prog)
    if (op.opcode == 0) {
      const double *src = &one;
      size_t len = 1;
      if (op.lhs != 0xFFFF) {
        if (op.lhs >= vm->consts.size()) die("encode: constant index out of range");
        src = vm->consts[op.lhs].data(), len = vm->consts[op.lhs].size();
      }
      vm->stage_host_values(src, len);
      vm->encode_internal(vm->ptr(op.dst), vm->ln->d_vals_in, (int)std::min(len, vm->N / 2), op.rhs >> 10, op.rhs & 0x3FF);
    }
  CUDA_CHECK(cudaStreamSynchronize(vm->ln->stream));
  // set-up that the first run() would otherwise pay for (the reference's hc-test times run() only): ciphertext registers,
  // the renaming pool and the scheduler's event pool (about two events per op)
  vm->preallocate_registers();
  vm->reserve_events(2 * vm->prog.size() + 64);
}
void encrypt(void *h, int64_t i, double *dat, int len) {
  // SEAL_HEVM.cpp:436-446.  Asynchronous: the values are copied to a pinned staging buffer, encode + encrypt are issued on
  // lane i mod #lanes (independent arguments overlap) and the call returns; run() / decrypt() / every other entry point
  // is ordered behind it.
  VM *vm = Vq(h);
  vm->res_cache.valid = false;
  if ((size_t)i >= vm->arg_level.size()) die("encrypt: argument index out of range");
  vm->encrypt_async((size_t)i, dat, (size_t)len);
}
void decrypt(void *h, int64_t i, double *dat) {
  VM *vm = V(h);
  CtReg &c = vm->ctr((size_t)i);
  if (c.level < 1) die("decrypt: empty register");
  vm->decrypt_to_pt(c, vm->ln->boot_pt);
  vm->decode_pt(vm->ln->boot_pt, vm->ln->d_vals);
  CUDA_CHECK(cudaMemcpyAsync(dat, vm->ln->d_vals, (vm->N / 2) * sizeof(double), cudaMemcpyDeviceToHost, vm->ln->stream));
  CUDA_CHECK(cudaStreamSynchronize(vm->ln->stream));
}
void decrypt_result(void *h, int64_t i, double *dat) { // SEAL_HEVM.cpp:459-461
  VM *vm = Vq(h);
  vm->join_async();
  if (!vm->res_cache.valid) vm->decrypt_all_results(); // all results at once, one per lane; later calls only copy
  if ((size_t)i >= vm->res_dst.size() || !vm->res_cache.ok[(size_t)i]) {
    decrypt(h, (int64_t)vm->res_dst.at((size_t)i), dat); // reports the error
    return;
  }
  CUDA_CHECK(cudaEventSynchronize(vm->res_cache.ev[(size_t)i]));
  std::memcpy(dat, vm->res_cache.h + (size_t)i * (vm->N / 2), (vm->N / 2) * sizeof(double));
}
int64_t getResIdx(void *h, int64_t i) { return (int64_t)Vq(h)->res_dst.at((size_t)i); }
void *getCtxt(void *h, int64_t id) { return &V(h)->ctr((size_t)id); }
void run(void *h) { // SEAL_HEVM.cpp:336-401; returns only when the results are complete
  VM *vm = V(h);
  vm->run_program();
}
int64_t getArgLen(void *h) { return (int64_t)Vq(h)->head.n_args; }
int64_t getResLen(void *h) { return (int64_t)Vq(h)->head.n_res; }
void setDebug(void *h, bool e) { V(h)->debug = e; }
void setToGPU(void *, bool) {} // always resident on the GPU
void printMem(void *h) {
  if (!V(h)->debug) return;
  size_t fr = 0, tot = 0;
  CUDA_CHECK(cudaMemGetInfo(&fr, &tot));
  std::printf("MemUsage: %.2fGB (%.1f%%)\n", (tot - fr) / 1e9, 100.0 * (tot - fr) / tot);
}

// ================= hevmx_* hooks (include/hevm_ext.h) =========================================
int64_t hevmx_param(void *h, int what) {
  VM *vm = V(h);
  switch (what) {
  case 0: return vm->logN;
  case 1: return vm->L;
  case 2: return (int64_t)(vm->seed.k[0] | ((uint64_t)vm->seed.k[1] << 32));
  case 3: return (int64_t)vm->ct.size();
  case 4: return (int64_t)vm->pt.size();
  case 5: return (int64_t)vm->d_gal.size();
  case 6: return (int64_t)g_launch_count;
  }
  return -1;
}
void hevmx_primes(void *h, uint64_t *out) {
  for (int i = 0; i < V(h)->L; i++) out[i] = V(h)->P.q[i];
}
void hevmx_roots(void *h, uint64_t *out) {
  for (int i = 0; i < V(h)->L; i++) out[i] = V(h)->P.psi[i];
}
void hevmx_resize(void *h, int64_t nct, int64_t npt) { V(h)->resize_regs((size_t)nct, (size_t)npt); }
void hevmx_ct_info(void *h, int64_t r, int64_t *level, double *scale) {
  CtReg &c = V(h)->ctr((size_t)r);
  *level = c.level, *scale = c.scale;
}
void hevmx_ct_read(void *h, int64_t r, uint64_t *out) {
  VM *vm = V(h);
  CtReg &c = vm->ctr((size_t)r);
  const size_t w = (size_t)c.level * vm->N;
  for (int K = 0; K < 2; K++) CUDA_CHECK(cudaMemcpyAsync(out + K * w, c.d + K * vm->pitch, w * 8, cudaMemcpyDeviceToHost, vm->ln->stream));
  CUDA_CHECK(cudaStreamSynchronize(vm->ln->stream));
}
void hevmx_ct_write(void *h, int64_t r, const uint64_t *in, int64_t level, double scale) {
  VM *vm = V(h);
  CtReg &c = vm->ctr((size_t)r);
  vm->rehome(c);
  if (level < 1 || level > vm->L - 1) die("ct_write: level out of range");
  const size_t w = (size_t)level * vm->N;
  for (int K = 0; K < 2; K++) CUDA_CHECK(cudaMemcpyAsync(c.d + K * vm->pitch, in + K * w, w * 8, cudaMemcpyHostToDevice, vm->ln->stream));
  CUDA_CHECK(cudaStreamSynchronize(vm->ln->stream));
  c.level = (int)level, c.scale = scale;
}
void hevmx_pt_info(void *h, int64_t r, int64_t *level, double *scale) {
  PtReg &p = V(h)->ptr((size_t)r);
  *level = p.level, *scale = p.scale;
}
void hevmx_pt_read(void *h, int64_t r, uint64_t *out) {
  VM *vm = V(h);
  PtReg &p = vm->ptr((size_t)r);
  CUDA_CHECK(cudaMemcpyAsync(out, p.d, (size_t)p.level * vm->N * 8, cudaMemcpyDeviceToHost, vm->ln->stream));
  CUDA_CHECK(cudaStreamSynchronize(vm->ln->stream));
}
void hevmx_pt_write(void *h, int64_t r, const uint64_t *in, int64_t level, double scale) {
  VM *vm = V(h);
  PtReg &p = vm->ptr((size_t)r);
  CUDA_CHECK(cudaStreamSynchronize(vm->ln->stream));
  vm->pt_reserve(p, (int)level);
  CUDA_CHECK(cudaMemcpyAsync(p.d, in, (size_t)level * vm->N * 8, cudaMemcpyHostToDevice, vm->ln->stream));
  CUDA_CHECK(cudaStreamSynchronize(vm->ln->stream));
  p.level = (int)level, p.scale = scale;
}
void hevmx_exec(void *h, int64_t opcode, int64_t dst, int64_t lhs, int64_t rhs) {
  HevmOp op{(uint16_t)opcode, (uint16_t)dst, (uint16_t)lhs, (uint16_t)rhs};
  VM *vm = V(h);
  vm->ln = &vm->lanes[0];
  if (opcode == 10) {
    CUDA_CHECK(cudaMemcpyAsync(vm->d_ctr_base, &vm->enc_counter, 8, cudaMemcpyHostToDevice, vm->ln->stream));
    vm->boot_index = 0;
  }
  vm->exec(op);
  if (opcode == 10) vm->enc_counter++;
}
void hevmx_exec_batch(void *h, int64_t opcode, int64_t n, const int64_t *dst, const int64_t *lhs, const int64_t *rhs) {
  VM *vm = V(h);
  vm->ln = &vm->lanes[0];
  vm->exec_batch((int)opcode, (int)n, dst, lhs, rhs);
}
void hevmx_sync(void *h) { CUDA_CHECK(cudaStreamSynchronize(V(h)->ln->stream)); }
void hevmx_ntt(void *h, uint64_t *data, int64_t prime_idx, int64_t count, int inverse) {
  VM *vm = V(h);
  const size_t w = (size_t)count * vm->N;
  u64 *d = dalloc<u64>(w);
  CUDA_CHECK(cudaMemcpyAsync(d, data, w * 8, cudaMemcpyHostToDevice, vm->ln->stream));
  if (inverse == 2) // the single-pass cluster NTT (forward), for parity tests of the experiment
    launch_ntt_fwd_cluster(vm->ln->stream, vm->dT, vm->logN, d, d, (int)count, (int)prime_idx, 0);
  else if (inverse)
    vm->ln->ops->ntt_inv(d, d, (int)count, (int)prime_idx, 0);
  else
    vm->ln->ops->ntt_fwd(d, d, (int)count, (int)prime_idx, 0);
  CUDA_CHECK(cudaMemcpyAsync(data, d, w * 8, cudaMemcpyDeviceToHost, vm->ln->stream));
  CUDA_CHECK(cudaStreamSynchronize(vm->ln->stream));
  CUDA_CHECK(cudaFree(d));
}
// device-resident NTT throughput: `batch` limbs under each of the first `nprimes` primes, transformed in place `reps`
// times (forward or inverse); returns the average milliseconds of one pass over all batch * nprimes limbs
double hevmx_ntt_bench(void *h, int64_t batch, int64_t nprimes, int inverse, int64_t reps) {
  VM *vm = V(h);
  vm->ln = &vm->lanes[0];
  if (nprimes < 1 || nprimes > vm->L || batch < 1) die("ntt_bench: bad arguments");
  const size_t w = (size_t)batch * nprimes * vm->N;
  u64 *d = dalloc<u64>(w);
  CUDA_CHECK(cudaMemsetAsync(d, 0x11, w * 8, vm->ln->stream)); // 0x1111... < q: valid residues
  auto pass = [&] {
    for (int64_t i = 0; i < nprimes; i++) {
      u64 *p = d + (size_t)i * batch * vm->N;
      if (inverse == 2)
        launch_ntt_fwd_cluster(vm->ln->stream, vm->dT, vm->logN, p, p, (int)batch, (int)i, 0);
      else if (inverse)
        vm->ln->ops->ntt_inv(p, p, (int)batch, (int)i, 0);
      else
        vm->ln->ops->ntt_fwd(p, p, (int)batch, (int)i, 0);
    }
  };
  pass();
  cudaEvent_t a, b;
  CUDA_CHECK(cudaEventCreate(&a));
  CUDA_CHECK(cudaEventCreate(&b));
  CUDA_CHECK(cudaEventRecord(a, vm->ln->stream));
  for (int64_t r = 0; r < reps; r++) pass();
  CUDA_CHECK(cudaEventRecord(b, vm->ln->stream));
  CUDA_CHECK(cudaEventSynchronize(b));
  float ms = 0;
  CUDA_CHECK(cudaEventElapsedTime(&ms, a, b));
  CUDA_CHECK(cudaEventDestroy(a));
  CUDA_CHECK(cudaEventDestroy(b));
  CUDA_CHECK(cudaFree(d));
  return (double)ms / (double)reps;
}
void hevmx_encode(void *h, int64_t ptreg, const double *vals, int64_t len, int64_t level, int64_t scale_bits) {
  VM *vm = V(h);
  vm->stage_host_values(vals, (size_t)len);
  vm->encode_internal(vm->ptr((size_t)ptreg), vm->ln->d_vals_in, (int)std::min((size_t)len, vm->N / 2), level, scale_bits);
  CUDA_CHECK(cudaStreamSynchronize(vm->ln->stream));
}
void hevmx_decode(void *h, int64_t ptreg, double *out) {
  VM *vm = V(h);
  vm->decode_pt(vm->ptr((size_t)ptreg), vm->ln->d_vals);
  CUDA_CHECK(cudaMemcpyAsync(out, vm->ln->d_vals, (vm->N / 2) * sizeof(double), cudaMemcpyDeviceToHost, vm->ln->stream));
  CUDA_CHECK(cudaStreamSynchronize(vm->ln->stream));
}
void hevmx_decrypt_to_pt(void *h, int64_t ctreg, int64_t ptreg) { V(h)->decrypt_to_pt(V(h)->ctr((size_t)ctreg), V(h)->ptr((size_t)ptreg)); }
void hevmx_encrypt_pt(void *h, int64_t ptreg, int64_t ctreg) { V(h)->encrypt_pt_now(V(h)->ptr((size_t)ptreg), V(h)->ctr((size_t)ctreg)); }
// TEST SWITCH: pins the encryption randomness to (key seed, counter) so that this library and the oracle, started from
// the same hevm_params.bin, produce identical ciphertexts.  A captured run() graph carries the key it was built with.
void hevmx_set_enc_counter(void *h, uint64_t c) {
  VM *vm = V(h);
  if (std::memcmp(vm->enc_seed.k, vm->seed.k, sizeof vm->seed.k) != 0) vm->invalidate_graph();
  vm->enc_seed = vm->seed;
  vm->enc_counter = c;
}
int64_t hevmx_key_read(void *h, int which, uint64_t elt, uint64_t *out) {
  VM *vm = V(h);
  const u64 *src = nullptr;
  size_t w = 0;
  const size_t kw = (size_t)(vm->L - 1) * 2 * vm->key_limbs() * vm->N; // limb-sharded storage: [L-1][2][own limbs][N]
  if (which == 0) src = vm->d_sk, w = (size_t)vm->L * vm->N;
  if (which == 1) src = vm->d_pk, w = (size_t)2 * vm->L * vm->N;
  if (which == 2) src = vm->d_relin, w = kw;
  if (which == 3) {
    auto it = vm->d_gal.find(elt);
    if (it == vm->d_gal.end()) return -1;
    src = it->second, w = kw;
  }
  if (!src) return -1;
  if (out) {
    CUDA_CHECK(cudaMemcpyAsync(out, src, w * 8, cudaMemcpyDeviceToHost, vm->ln->stream));
    CUDA_CHECK(cudaStreamSynchronize(vm->ln->stream));
    if (which >= 2) // key-switch keys live in radix-2^30 split form on the device: hand back canonical residues
      for (size_t i = 0; i < w; i++) out[i] = join30(out[i]);
  }
  return (int64_t)w;
}
int64_t hevmx_galois_elt(void *h, int64_t step) { return (int64_t)V(h)->galois_elt_from_step((int)step); }
const char *hevmx_backend(void) { return "b200-cuda-sm_100a"; }
// CUDA-event stopwatch on the VM's own stream: which=0 start, which=1 stop -> elapsed milliseconds
double hevmx_timer(void *h, int which) {
  VM *vm = V(h);
  if (!vm->ev_start) {
    CUDA_CHECK(cudaEventCreate(&vm->ev_start));
    CUDA_CHECK(cudaEventCreate(&vm->ev_stop));
  }
  if (which == 0) {
    CUDA_CHECK(cudaEventRecord(vm->ev_start, vm->ln->stream));
    return 0.0;
  }
  CUDA_CHECK(cudaEventRecord(vm->ev_stop, vm->ln->stream));
  CUDA_CHECK(cudaEventSynchronize(vm->ev_stop));
  float ms = 0;
  CUDA_CHECK(cudaEventElapsedTime(&ms, vm->ev_start, vm->ev_stop));
  return (double)ms;
}
// cudaProfilerStart/Stop so that `ncu --profile-from-start off` skips key generation
void hevmx_ks_shard_stage(void *h, int stage, int64_t dst, int64_t src, int64_t step, int64_t tlo, int64_t thi) {
  VM *vm = V(h);
  vm->ln = &vm->lanes[0];
  CtReg &s = vm->ctr((size_t)src), &d = vm->ctr((size_t)dst);
  if (s.level < 1) die("ks_shard: empty source register");
  if (tlo < 0 || thi > s.level + 1 || tlo > thi) die("ks_shard: bad target range");
  const u64 elt = vm->galois_elt_from_step((int)step);
  auto it = vm->d_gal.find(elt);
  if (it == vm->d_gal.end()) die("ks_shard: no Galois key for this step (use a power of two)");
  vm->ln->ops->ks_shard_stage(stage, LD_GALOIS, s.d, nullptr, d.d, vm->pitch, s.level, it->second, (u32)elt, (int)tlo, (int)thi);
  if (stage == 3) d.level = s.level, d.scale = s.scale;
}
void hevmx_mulcc_shard_stage(void *h, int stage, int64_t dst, int64_t lhs, int64_t rhs, int64_t tlo, int64_t thi) {
  VM *vm = V(h);
  vm->ln = &vm->lanes[0];
  CtReg &a = vm->ctr((size_t)lhs), &b = vm->ctr((size_t)rhs), &d = vm->ctr((size_t)dst);
  if (a.level != b.level || a.level < 1) die("mulcc_shard: level mismatch");
  if (tlo < 0 || thi > a.level + 1 || tlo > thi) die("mulcc_shard: bad target range");
  const int l = a.level;
  // product scale, fixed at stage 1: dst may alias an operand, and stage 3 may run once per owned range
  if (stage == 1) vm->shard_scale = a.scale * b.scale;
  vm->ln->ops->ks_shard_stage(stage, LD_PRODUCT, a.d, b.d, d.d, vm->pitch, l, vm->d_relin, 0, (int)tlo, (int)thi);
  if (stage == 3) d.level = l, d.scale = vm->shard_scale;
}
// ---- peer-to-peer limb-sharded key switch (include/hevm_ext.h) ----
void hevmx_p2p_setup(void *h, int64_t rank, int64_t world, uint8_t *handle_out /*64 bytes*/) {
  VM *vm = V(h);
  if (world < 1 || world > 8 || rank < 0 || rank >= world) die("p2p_setup: at most 8 ranks");
  if (world > 1 && 2 * world > vm->L) die("p2p_setup: at least two limbs per rank");
  if (vm->keys_sharded() && (vm->shard_rank != rank || vm->shard_world != world)) die("p2p_setup: rank / world differ from HEVM_SHARD_RANK / HEVM_SHARD_WORLD");
  if (!vm->keys_sharded()) { // whole keys, sharded work only: same static ownership
    vm->shard_rank = (int)rank, vm->shard_world = (int)world;
    VM::own_range(vm->L, (int)rank, (int)world, vm->own_lo, vm->own_hi);
    vm->shard_world = 1; // keys stay whole (keys_sharded() false); ownership is kept in own_lo / own_hi
  }
  auto &p = vm->p2p;
  if (!p.block) {
    p.block = dalloc<u64>(vm->p2p_words());
    CUDA_CHECK(cudaMemset(p.block, 0, vm->p2p_words() * 8));
    CUDA_CHECK(cudaDeviceSynchronize()); // legacy-stream memset: the (non-blocking) lane streams do not wait for it
    for (auto &e : p.ev) CUDA_CHECK(cudaEventCreate(&e));
  }
  p.rank = (int)rank, p.world = (int)world, p.on = true;
  p.peer.p[rank] = p.block;
  if (handle_out) {
    cudaIpcMemHandle_t hd;
    CUDA_CHECK(cudaIpcGetMemHandle(&hd, p.block));
    static_assert(sizeof(hd) == 64, "cudaIpcMemHandle_t is 64 bytes");
    std::memcpy(handle_out, &hd, 64);
  }
}
// map a peer's exchange block: `handle` from its hevmx_p2p_setup (another process), or -- same process, for tests with
// several VMs on one GPU -- the peer VM itself
void hevmx_p2p_connect(void *h, int64_t peer, const uint8_t *handle, void *peer_vm_same_process) {
  VM *vm = V(h);
  if (peer < 0 || peer >= vm->p2p.world || peer == vm->p2p.rank) die("p2p_connect: bad peer");
  if (peer_vm_same_process) {
    vm->p2p.peer.p[peer] = V(peer_vm_same_process)->p2p.block;
    return;
  }
  cudaIpcMemHandle_t hd;
  std::memcpy(&hd, handle, 64);
  void *ptr = nullptr;
  CUDA_CHECK(cudaIpcOpenMemHandle(&ptr, hd, cudaIpcMemLazyEnablePeerAccess));
  vm->p2p.peer.p[peer] = (u64 *)ptr;
}
// asynchronous: opcode 1 = rotate by `rhs` (a step with its own Galois key), 8 = multiply + relinearise with register `rhs`
void hevmx_ks_shard_p2p_phase(void *h, int64_t phase, int64_t opcode, int64_t dst, int64_t lhs, int64_t rhs);
void hevmx_ks_shard_p2p(void *h, int64_t opcode, int64_t dst, int64_t lhs, int64_t rhs) { hevmx_ks_shard_p2p_phase(h, 0, opcode, dst, lhs, rhs); }
void hevmx_ks_shard_p2p_phase(void *h, int64_t phase, int64_t opcode, int64_t dst, int64_t lhs, int64_t rhs) {
  VM *vm = V(h);
  if (phase < 0 || phase > 3) die("ks_shard_p2p_phase: phase must be 0..3");
  CtReg &a = vm->ctr((size_t)lhs), &d = vm->ctr((size_t)dst);
  if (a.level < 1) die("ks_shard_p2p: empty source register");
  if (opcode == 1) {
    const u64 elt = vm->galois_elt_from_step((int)(int16_t)rhs);
    auto it = vm->d_gal.find(elt);
    if (it == vm->d_gal.end()) die("ks_shard_p2p: no Galois key for this step");
    const double sc = a.scale;
    const int l = a.level;
    vm->ks_shard_p2p(LD_GALOIS, d, a, nullptr, it->second, (u32)elt, (int)phase);
    if (phase == 0 || phase == 3) d.level = l, d.scale = sc;
  } else if (opcode == 8) {
    CtReg &b = vm->ctr((size_t)rhs);
    if (a.level != b.level) die("ks_shard_p2p: level mismatch");
    if (phase <= 1) vm->shard_scale = a.scale * b.scale; // dst may alias an operand: fix the product scale before stage 3 overwrites it
    const int l = a.level;
    vm->ks_shard_p2p(LD_PRODUCT, d, a, &b, vm->d_relin, 0, (int)phase);
    if (phase == 0 || phase == 3) d.level = l, d.scale = vm->shard_scale;
  } else {
    die("ks_shard_p2p: opcode must be 1 (rotate) or 8 (mulcc)");
  }
}
// bandwidth probe of the exchange path: push `words` u64 of this rank's block into every peer `reps` times (no flags are
// consumed: epoch 0); returns the average milliseconds of one push on this rank.  All ranks must call it together.
double hevmx_p2p_bench(void *h, int64_t words, int64_t reps) {
  VM *vm = V(h);
  auto &p = vm->p2p;
  if (!p.on || words < 0 || (size_t)words > (size_t)vm->L * vm->N) die("p2p_bench: bad arguments");
  Lane &L0 = vm->lanes[0];
  unsigned *done = reinterpret_cast<unsigned *>(p.block + vm->p2p_words() - 2);
  cudaEvent_t a, b;
  CUDA_CHECK(cudaEventCreate(&a));
  CUDA_CHECK(cudaEventCreate(&b));
  launch_p2p_push(L0.stream, p.block, (size_t)words, p.peer, 0, p.rank, p.world, done, vm->p2p_flag_off(1), 0);
  CUDA_CHECK(cudaEventRecord(a, L0.stream));
  for (int64_t r = 0; r < reps; r++) launch_p2p_push(L0.stream, p.block, (size_t)words, p.peer, 0, p.rank, p.world, done, vm->p2p_flag_off(1), 0);
  CUDA_CHECK(cudaEventRecord(b, L0.stream));
  CUDA_CHECK(cudaEventSynchronize(b));
  float ms = 0;
  CUDA_CHECK(cudaEventElapsedTime(&ms, a, b));
  CUDA_CHECK(cudaEventDestroy(a));
  CUDA_CHECK(cudaEventDestroy(b));
  return (double)ms / (double)reps;
}
// the owned key-switch targets [tlo, thi) of this rank at `level` (target `level` = the special limb)
void hevmx_p2p_targets(void *h, int64_t level, int64_t *tlo, int64_t *thi) {
  int a, b;
  V(h)->my_targets((int)level, a, b);
  *tlo = a, *thi = b;
}
// CUDA-event breakdown of the NEXT sharded key switch: on = 1 arms it; after hevmx_sync, on = 0 reads five durations (ms):
// stage 1 | digit push | stage 2 (mod-up pass A per source rank behind that rank's flag, then the inner product) |
// rounded-row push / wait | stage 3
void hevmx_p2p_timing(void *h, int on, double *out5) {
  VM *vm = V(h);
  if (on) {
    vm->p2p.timed = true;
    return;
  }
  vm->p2p.timed = false;
  CUDA_CHECK(cudaEventSynchronize(vm->p2p.ev[5]));
  for (int i = 0; i < 5; i++) {
    float ms = 0;
    CUDA_CHECK(cudaEventElapsedTime(&ms, vm->p2p.ev[i], vm->p2p.ev[i + 1]));
    out5[i] = ms;
  }
}
// key material from the host (canonical residues, SEAL layouts): which = 0 sk[L][N], 1 pk[2][L][N], 2 relin[L-1][2][L][N],
// 3 Galois key of `elt` (created if absent).  Used by the .seal loader (dacapo_b200/seal_format.py).
void hevmx_key_write(void *h, int which, uint64_t elt, const uint64_t *in) {
  VM *vm = V(h);
  if (vm->keys_sharded()) die("key_write: not available with limb-sharded key storage");
  vm->invalidate_graph();
  const size_t LN = (size_t)vm->L * vm->N, kw = (size_t)(vm->L - 1) * 2 * LN;
  cudaStream_t st = vm->ln->stream;
  CUDA_CHECK(cudaStreamSynchronize(st));
  // every copy is ordered on the lane's own (non-blocking) stream: a legacy-stream cudaMemcpy from pageable memory may
  // return before its DMA has landed, and nothing orders it against kernels of a non-blocking stream
  if (which == 0) CUDA_CHECK(cudaMemcpyAsync(vm->d_sk, in, LN * 8, cudaMemcpyHostToDevice, st));
  if (which == 1) CUDA_CHECK(cudaMemcpyAsync(vm->d_pk, in, 2 * LN * 8, cudaMemcpyHostToDevice, st));
  if (which == 2 || which == 3) {
    u64 *&dst = (which == 2) ? vm->d_relin : vm->d_gal[elt];
    if (!dst) dst = dalloc<u64>(kw);
    CUDA_CHECK(cudaMemcpyAsync(dst, in, kw * 8, cudaMemcpyHostToDevice, st));
    launch_key_split(st, dst, kw);
  }
  CUDA_CHECK(cudaStreamSynchronize(st));
}
// drop every Galois key (before loading a foreign key set whose rotation steps differ from the default set)
void hevmx_galois_clear(void *h) {
  VM *vm = V(h);
  vm->invalidate_graph();
  for (auto &kv : vm->d_gal)
    if (kv.second) CUDA_CHECK(cudaFree(kv.second));
  vm->d_gal.clear();
}
void *hevmx_dev_ptr(void *h, int64_t which) {
  VM *vm = V(h);
  if (which == 0) return vm->lanes[0].ops->sc.t;
  if (which == 1) return vm->lanes[0].ops->sc.rnd;
  if (which >= 16) return vm->ctr((size_t)(which - 16)).d;
  return nullptr;
}
void *hevmx_stream(void *h) { return (void *)V(h)->lanes[0].stream; }
void hevmx_profiler_range(void *h, int on) {
  CUDA_CHECK(cudaStreamSynchronize(V(h)->ln->stream));
  if (on)
    CUDA_CHECK(cudaProfilerStart());
  else
    CUDA_CHECK(cudaProfilerStop());
}
// per-kernel-class CUDA-event timing: on=1 reset+enable, on=0 collect+disable
void hevmx_profile(void *h, int on) {
  CUDA_CHECK(cudaStreamSynchronize(V(h)->ln->stream));
  if (on) {
    g_prof.reset();
    g_prof.on = true;
  } else {
    g_prof.on = false;
    g_prof.collect();
  }
}
const char *hevmx_profile_read(void *, int cls, double *ms, int64_t *count) {
  if (cls < 0 || cls >= KC_COUNT) return nullptr;
  *ms = g_prof.ms[cls], *count = g_prof.cnt[cls];
  return g_kernel_class_names[cls];
}
}
