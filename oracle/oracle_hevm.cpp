// =============================================================================
// oracle/oracle_hevm.cpp  --  TEST INFRASTRUCTURE ONLY (CPU oracle, libORACLE_HEVM.so)
//
// CPU restatement of the reference HEVM interpreter (reference:
// lib/Runtime/SEAL_HEVM.cpp:15-504) on top of ckks_oracle.hpp.  Exports the same
// 18 C symbols as the reference plus the hevmx_* inspection hooks declared in
// include/hevm_ext.h, so the parity tests drive both libraries identically.
// "parity unpinned" at the SEAL boundary -- see ckks_oracle.hpp header.
// =============================================================================
#include "ckks_oracle.hpp"
#include <fstream>
#include <functional>
#include <random>

using namespace orc;

namespace {

// on-disk HEVM container (reference: include/hecate/Support/HEVMHeader.h:9-34,
// writer lib/Dialect/CKKS/Transforms/EmitHEVM.cpp:31-119). Own declaration.
#pragma pack(push, 1)
struct FileHead {
  uint32_t magic;       // 0x4845564D
  uint32_t head_size;   // 24
  uint64_t n_args, n_res;
};
struct FileConfig {
  uint64_t body_len, n_ops, n_ct, n_pt, init_level;
};
struct FileOp {
  uint16_t opcode, dst, lhs, rhs;
};
#pragma pack(pop)

struct ParamFile {
  uint64_t magic;
  uint64_t logN, L, bits;
  uint32_t seed[8]; // 256-bit key seed (format version 2)
};
const uint64_t PARAM_MAGIC = 0x3230324D56454842ull; // "BHEVM202" little-endian tag: version 2 = 256-bit seed

struct VM {
  Context ctx;
  std::vector<std::vector<double>> consts;
  FileHead head{};
  FileConfig cfg{};
  std::vector<FileOp> ops;
  std::vector<uint64_t> arg_scale, arg_level, res_scale, res_level, res_dst;
  std::vector<Ct> ct;
  std::vector<Pt> pt;
  bool debug = false;

  void encode_internal(Pt &dst, const double *src, size_t len, int64_t level, int64_t scale_bits) {
    // SEAL_HEVM.cpp:256-267
    size_t slots = ctx.N >> 1;
    std::vector<double> d(slots);
    for (size_t i = 0; i < slots; i++) d[i] = src[i % len];
    ctx.encode(d.data(), (int)level, std::pow(2.0, (double)scale_bits), dst);
  }
  void exec(const FileOp &op) {
    switch (op.opcode) {
    case 1: { // rotate  SEAL_HEVM.cpp:269-274
      Ct r = ct[op.lhs];
      ctx.rotate_inplace(r, (int16_t)op.rhs);
      ct[op.dst] = std::move(r);
      break;
    }
    case 2: ctx.negate(ct[op.lhs], ct[op.dst]); break;           // 275-279
    case 3: ctx.rescale(ct[op.lhs], ct[op.dst]); break;          // 280-284
    case 4: {                                                    // 285-293
      int16_t down = (int16_t)op.rhs;
      if (down > 0) ctx.mod_switch(ct[op.lhs], ct[op.dst]);
      for (int i = 1; i < down; i++) ctx.mod_switch(ct[op.dst], ct[op.dst]);
      break;
    }
    case 5: die("This VM does not support native upscale op"); // 294-296
    case 6:                                                    // 297-303
      ct[op.lhs].scale = ct[op.rhs].scale;
      ctx.add(ct[op.lhs], ct[op.rhs], ct[op.dst]);
      break;
    case 7: // 304-310
      ct[op.lhs].scale = pt[op.rhs].scale;
      ctx.add_plain(ct[op.lhs], pt[op.rhs], ct[op.dst]);
      break;
    case 8: { // 311-317
      Ct r;
      ctx.multiply(ct[op.lhs], ct[op.rhs], r);
      ctx.relinearize(r);
      ct[op.dst] = std::move(r);
      break;
    }
    case 9: ctx.multiply_plain(ct[op.lhs], pt[op.rhs], ct[op.dst]); break; // 318-323
    case 10: {                                                             // 324-334 (Release build)
      Pt p;
      ctx.decrypt(ct[op.lhs], p);
      std::vector<double> v(ctx.N >> 1);
      ctx.decode(p, v.data());
      int64_t sb = (int64_t)std::log2(ct[op.lhs].scale);
      Pt e;
      encode_internal(e, v.data(), v.size(), op.rhs, sb);
      ctx.encrypt(e, ct[op.dst]);
      break;
    }
    default: break; // 0 (encode: preprocess only), 0xFFFF placeholders
    }
  }
};

void read_params(const std::string &dir, ParamFile &pf) {
  std::ifstream f(dir + "/hevm_params.bin", std::ios::binary);
  if (!f) die("cannot open hevm_params.bin (run create_context first)");
  f.read((char *)&pf, sizeof pf);
  if (pf.magic != PARAM_MAGIC) die("bad hevm_params.bin");
}

} // namespace

extern "C" {

// ---- the reference's 18 symbols (SEAL_HEVM.cpp:404-504) -------------------------
void create_context(char *dir) {
  ParamFile pf{};
  pf.magic = PARAM_MAGIC;
  const char *e;
  pf.logN = (e = std::getenv("HEVM_LOGN")) ? std::strtoull(e, nullptr, 10) : 15;   // SEAL_HEVM.cpp:39
  pf.L = (e = std::getenv("HEVM_NUM_PRIMES")) ? std::strtoull(e, nullptr, 10) : 14; // SEAL_HEVM.cpp:40
  pf.bits = (e = std::getenv("HEVM_PRIME_BITS")) ? std::strtoull(e, nullptr, 10) : 60;
  if ((e = std::getenv("HEVM_SEED"))) { // tests: a fixed, public seed
    const uint64_t v = std::strtoull(e, nullptr, 0);
    pf.seed[0] = (uint32_t)v, pf.seed[1] = (uint32_t)(v >> 32);
  } else {
    std::random_device rd;
    for (auto &w : pf.seed) w = rd();
  }
  std::ofstream f(std::string(dir) + "/hevm_params.bin", std::ios::binary);
  f.write((const char *)&pf, sizeof pf);
}
void *initFullVM(char *dir, bool /*device*/) {
  auto vm = new VM();
  ParamFile pf;
  read_params(dir, pf);
  Seed256 sd;
  std::memcpy(sd.k, pf.seed, sizeof sd.k);
  vm->ctx.init_params((int)pf.logN, (int)pf.L, (int)pf.bits, sd);
  { // encryption randomness: fresh per VM (hevmx_set_enc_counter pins it to the key seed for the parity tests)
    std::random_device rd;
    for (auto &w : vm->ctx.enc_seed.k) w = rd();
  }
  vm->ctx.keygen();
  return vm;
}
void *initClientVM(char *dir) { return initFullVM(dir, false); }
void *initServerVM(char *dir) { return initFullVM(dir, false); }

void load(void *h, char *cst, char *hevm) {
  auto vm = (VM *)h;
  {
    std::ifstream f(cst, std::ios::binary);
    if (!f) die("cannot open constants file");
    int64_t n = 0;
    f.read((char *)&n, 8);
    vm->consts.assign(n, {});
    for (int64_t i = 0; i < n; i++) {
      int64_t len = 0;
      f.read((char *)&len, 8);
      vm->consts[i].resize(len);
      f.read((char *)vm->consts[i].data(), len * 8);
    }
  }
  std::ifstream f(hevm, std::ios::binary);
  if (!f) die("cannot open hevm file");
  f.read((char *)&vm->head, sizeof(FileHead));
  if (vm->head.magic != 0x4845564D) die("bad HEVM magic");
  f.read((char *)&vm->cfg, sizeof(FileConfig));
  auto rd = [&](std::vector<uint64_t> &v, size_t n) {
    v.resize(n);
    f.read((char *)v.data(), n * 8);
  };
  rd(vm->arg_scale, vm->head.n_args);
  rd(vm->arg_level, vm->head.n_args);
  rd(vm->res_scale, vm->head.n_res);
  rd(vm->res_level, vm->head.n_res);
  rd(vm->res_dst, vm->head.n_res);
  vm->ops.resize(vm->cfg.n_ops);
  f.read((char *)vm->ops.data(), vm->ops.size() * sizeof(FileOp));
  vm->ct.assign(vm->cfg.n_ct, Ct());
  vm->pt.assign(vm->cfg.n_pt, Pt());
}
void loadClient(void *, void *) { die("client/server split is vestigial in the reference (SURVEY A.3); use initFullVM"); }

void preprocess(void *h) { // SEAL_HEVM.cpp:242-254
  auto vm = (VM *)h;
  std::vector<double> ones(1, 1.0);
  for (auto &op : vm->ops)
    if (op.opcode == 0) {
      const std::vector<double> &src = (op.lhs == 0xFFFF) ? ones : vm->consts.at(op.lhs);
      vm->encode_internal(vm->pt.at(op.dst), src.data(), src.size(), op.rhs >> 10, op.rhs & 0x3FF);
    }
}
void encrypt(void *h, int64_t i, double *dat, int len) { // 439-445
  auto vm = (VM *)h;
  Pt p;
  vm->encode_internal(p, dat, (size_t)len, (int64_t)vm->arg_level.at(i), (int64_t)vm->arg_scale.at(i));
  vm->ctx.encrypt(p, vm->ct.at(i));
}
void decrypt(void *h, int64_t i, double *dat) { // 446-455 (single copy)
  auto vm = (VM *)h;
  Pt p;
  vm->ctx.decrypt(vm->ct.at(i), p);
  vm->ctx.decode(p, dat);
}
void decrypt_result(void *h, int64_t i, double *dat) { decrypt(h, (int64_t)((VM *)h)->res_dst.at(i), dat); }
int64_t getResIdx(void *h, int64_t i) { return (int64_t)((VM *)h)->res_dst.at(i); }
void *getCtxt(void *h, int64_t id) { return &((VM *)h)->ct.at(id); }
void run(void *h) { // 336-401
  auto vm = (VM *)h;
  for (auto &op : vm->ops) vm->exec(op);
}
int64_t getArgLen(void *h) { return (int64_t)((VM *)h)->head.n_args; }
int64_t getResLen(void *h) { return (int64_t)((VM *)h)->head.n_res; }
void setDebug(void *h, bool e) { ((VM *)h)->debug = e; }
void setToGPU(void *, bool) { die("This Library does not support GPU Backend"); }
void printMem(void *) {}

// ---- hevmx_* inspection hooks (include/hevm_ext.h) --------------------------------
int64_t hevmx_param(void *h, int what) {
  auto vm = (VM *)h;
  switch (what) {
  case 0: return vm->ctx.logN;
  case 1: return vm->ctx.L;
  case 2: return (int64_t)(vm->ctx.seed.k[0] | ((uint64_t)vm->ctx.seed.k[1] << 32));
  case 3: return (int64_t)vm->ct.size();
  case 4: return (int64_t)vm->pt.size();
  case 5: return (int64_t)vm->ctx.gal.size();
  }
  return -1;
}
void hevmx_primes(void *h, uint64_t *out) {
  auto vm = (VM *)h;
  for (int i = 0; i < vm->ctx.L; i++) out[i] = vm->ctx.q[i].q;
}
void hevmx_roots(void *h, uint64_t *out) {
  auto vm = (VM *)h;
  for (int i = 0; i < vm->ctx.L; i++) out[i] = vm->ctx.ntt[i].psi;
}
void hevmx_resize(void *h, int64_t nct, int64_t npt) {
  auto vm = (VM *)h;
  vm->ct.assign(nct, Ct());
  vm->pt.assign(npt, Pt());
}
void hevmx_ct_info(void *h, int64_t r, int64_t *level, double *scale) {
  auto vm = (VM *)h;
  *level = vm->ct.at(r).level;
  *scale = vm->ct.at(r).scale;
}
void hevmx_ct_read(void *h, int64_t r, uint64_t *out) {
  auto &c = ((VM *)h)->ct.at(r);
  std::memcpy(out, c.d.data(), c.d.size() * 8);
}
void hevmx_ct_write(void *h, int64_t r, const uint64_t *in, int64_t level, double scale) {
  auto vm = (VM *)h;
  auto &c = vm->ct.at(r);
  c.level = (int)level;
  c.size = 2;
  c.scale = scale;
  c.d.assign(in, in + (size_t)2 * level * vm->ctx.N);
}
void hevmx_pt_info(void *h, int64_t r, int64_t *level, double *scale) {
  auto vm = (VM *)h;
  *level = vm->pt.at(r).level;
  *scale = vm->pt.at(r).scale;
}
void hevmx_pt_read(void *h, int64_t r, uint64_t *out) {
  auto &p = ((VM *)h)->pt.at(r);
  std::memcpy(out, p.d.data(), p.d.size() * 8);
}
void hevmx_pt_write(void *h, int64_t r, const uint64_t *in, int64_t level, double scale) {
  auto vm = (VM *)h;
  auto &p = vm->pt.at(r);
  p.level = (int)level;
  p.scale = scale;
  p.d.assign(in, in + (size_t)level * vm->ctx.N);
}
void hevmx_exec(void *h, int64_t opcode, int64_t dst, int64_t lhs, int64_t rhs) {
  FileOp op{(uint16_t)opcode, (uint16_t)dst, (uint16_t)lhs, (uint16_t)rhs};
  ((VM *)h)->exec(op);
}
void hevmx_sync(void *) {}
// NTT of `count` consecutive N-word limbs, all under prime `prime_idx`
void hevmx_ntt(void *h, uint64_t *data, int64_t prime_idx, int64_t count, int inverse) {
  auto vm = (VM *)h;
  for (int64_t c = 0; c < count; c++) {
    if (inverse)
      vm->ctx.ntt.at(prime_idx).inverse(data + c * vm->ctx.N);
    else
      vm->ctx.ntt.at(prime_idx).forward(data + c * vm->ctx.N);
  }
}
void hevmx_encode(void *h, int64_t ptreg, const double *vals, int64_t len, int64_t level, int64_t scale_bits) {
  auto vm = (VM *)h;
  vm->encode_internal(vm->pt.at(ptreg), vals, (size_t)len, level, scale_bits);
}
void hevmx_decode(void *h, int64_t ptreg, double *out) {
  auto vm = (VM *)h;
  vm->ctx.decode(vm->pt.at(ptreg), out);
}
// plaintext register <- decrypt(ciphertext register)
void hevmx_decrypt_to_pt(void *h, int64_t ctreg, int64_t ptreg) {
  auto vm = (VM *)h;
  vm->ctx.decrypt(vm->ct.at(ctreg), vm->pt.at(ptreg));
}
// ciphertext register <- encrypt(plaintext register)
void hevmx_encrypt_pt(void *h, int64_t ptreg, int64_t ctreg) {
  auto vm = (VM *)h;
  vm->ctx.encrypt(vm->pt.at(ptreg), vm->ct.at(ctreg));
}
// TEST SWITCH: pins the encryption randomness to (key seed, counter) so that two libraries started from the same
// hevm_params.bin produce identical ciphertexts; without it every VM encrypts with fresh entropy
void hevmx_set_enc_counter(void *h, uint64_t c) {
  auto vm = (VM *)h;
  vm->ctx.enc_seed = vm->ctx.seed;
  vm->ctx.enc_counter = c;
}
// key material read-out: which = 0 sk [L][N], 1 pk [2][L][N], 2 relin, 3 galois(elt)
int64_t hevmx_key_read(void *h, int which, uint64_t elt, uint64_t *out) {
  auto vm = (VM *)h;
  const std::vector<u64> *src = nullptr;
  if (which == 0) src = &vm->ctx.sk;
  if (which == 1) src = &vm->ctx.pk;
  if (which == 2) src = &vm->ctx.relin.d;
  if (which == 3) {
    auto it = vm->ctx.gal.find(elt);
    if (it == vm->ctx.gal.end()) return -1;
    src = &it->second.d;
  }
  if (!src) return -1;
  if (out) std::memcpy(out, src->data(), src->size() * 8);
  return (int64_t)src->size();
}
int64_t hevmx_galois_elt(void *h, int64_t step) { return (int64_t)((VM *)h)->ctx.galois_elt_from_step((int)step); }
const char *hevmx_backend(void) { return "oracle-cpu"; }
}
