/* =============================================================================
 * include/hevm_ext.h -- inspection / micro-benchmark hooks (own extension)
 *
 * The reference exposes ciphertexts only as an opaque seal::Ciphertext*
 * (getCtxt, SEAL_HEVM.cpp:470-473).  Bit-exact parity tests and the op
 * microbenchmarks need raw residues, so both libB200_HEVM.so and the CPU oracle
 * export these hooks with identical meaning.  All buffers are HOST memory,
 * ciphertext layout is SEAL's: u64[2][level][N], NTT form.
 * ========================================================================== */
#ifndef HEVM_EXT_H
#define HEVM_EXT_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* what: 0 logN, 1 L (#primes incl. special), 2 seed, 3 #ct regs, 4 #pt regs, 5 #galois keys */
int64_t hevmx_param(void *vm, int what);
void hevmx_primes(void *vm, uint64_t *out /*[L]*/);
void hevmx_roots(void *vm, uint64_t *out /*[L] minimal primitive 2N-th roots*/);
void hevmx_resize(void *vm, int64_t n_ct, int64_t n_pt); /* size register files without a program */
void hevmx_ct_info(void *vm, int64_t reg, int64_t *level, double *scale);
void hevmx_ct_read(void *vm, int64_t reg, uint64_t *out);
void hevmx_ct_write(void *vm, int64_t reg, const uint64_t *in, int64_t level, double scale);
void hevmx_pt_info(void *vm, int64_t reg, int64_t *level, double *scale);
void hevmx_pt_read(void *vm, int64_t reg, uint64_t *out);
void hevmx_pt_write(void *vm, int64_t reg, const uint64_t *in, int64_t level, double scale);
/* execute ONE HEVM operation {opcode,dst,lhs,rhs} (HEVMHeader.h:27-32) asynchronously */
void hevmx_exec(void *vm, int64_t opcode, int64_t dst, int64_t lhs, int64_t rhs);
void hevmx_sync(void *vm);
/* NTT / INTT of `count` consecutive N-word limbs under prime `prime_idx` (host in/out) */
void hevmx_ntt(void *vm, uint64_t *data, int64_t prime_idx, int64_t count, int inverse);
void hevmx_encode(void *vm, int64_t ptreg, const double *vals, int64_t len, int64_t level, int64_t scale_bits);
void hevmx_decode(void *vm, int64_t ptreg, double *out /*N/2*/);
void hevmx_decrypt_to_pt(void *vm, int64_t ctreg, int64_t ptreg);
void hevmx_encrypt_pt(void *vm, int64_t ptreg, int64_t ctreg);
void hevmx_set_enc_counter(void *vm, uint64_t counter);
/* which: 0 sk[L][N], 1 pk[2][L][N], 2 relin[L-1][2][L][N], 3 galois key of `elt`; returns #words (or -1) */
int64_t hevmx_key_read(void *vm, int which, uint64_t elt, uint64_t *out);
int64_t hevmx_galois_elt(void *vm, int64_t step);
const char *hevmx_backend(void);
/* --- libB200_HEVM.so only (measurement; not part of the oracle) --- */
/* device-resident NTT throughput: batch limbs under each of the first nprimes primes, in place; returns ms per pass */
double hevmx_ntt_bench(void *vm, int64_t batch, int64_t nprimes, int inverse, int64_t reps);
/* n independent ops of ONE opcode (rotate with a single-key step / mulcc / rescale, sources at one level) as one
 * batched kernel launch (BASELINE configs[1] "batched ciphertexts"); anything else is issued op by op */
void hevmx_exec_batch(void *vm, int64_t opcode, int64_t n, const int64_t *dst, const int64_t *lhs, const int64_t *rhs);
double hevmx_timer(void *vm, int which);        /* CUDA events on the VM stream: 0 start, 1 stop -> ms */
void hevmx_profiler_range(void *vm, int on);  /* cudaProfilerStart/Stop (ncu --profile-from-start off) */
void hevmx_profile(void *vm, int on);           /* per-kernel-class CUDA-event timing */
const char *hevmx_profile_read(void *vm, int cls, double *ms, int64_t *count); /* NULL past the last class */
/* --- limb-sharded rotation key switch across GPUs (host side: dacapo_b200/sharded.py) ---
 * The l+1 key-switch targets (data limbs 0..l-1, special limb = target l) are partitioned over ranks; this rank owns
 * [tlo, thi).  stage 1: own limbs of perm(c1) -> coefficient digits; <all-gather digits>; stage 2: mod-up + key inner
 * product for the own targets, owner of the special limb also rounds it; <broadcast rounding rows>; stage 3: mod-down
 * of the own data limbs into `dst`.  `step` must have a Galois key of its own (one key switch). */
void hevmx_ks_shard_stage(void *vm, int stage, int64_t dst, int64_t src, int64_t step, int64_t tlo, int64_t thi);
/* the same three stages for mulcc (multiply + relinearise): stage 1 forms d2 = a1*b1 on the own limbs, stage 3 adds the
 * tensor-product terms d0, d1 of the own limbs; dst may alias lhs or rhs */
void hevmx_mulcc_shard_stage(void *vm, int stage, int64_t dst, int64_t lhs, int64_t rhs, int64_t tlo, int64_t thi);
/* device addresses for the exchanges: which 0 = coefficient digits [L][N] u64, 1 = rounding rows [2][N] u64,
 * 16 + r = limb data of ciphertext register r ([2][L-1][N] u64) */
void *hevmx_dev_ptr(void *vm, int64_t which);
void *hevmx_stream(void *vm);                   /* the cudaStream_t the VM issues hevmx_* work on */
/* --- the same key switch with the exchanges done by the library itself over peer memory (NVLink; CUDA IPC) ---
 * hevmx_p2p_setup allocates this rank's exchange block (digits + rounded rows, double-buffered, + epoch flags) and
 * returns its 64-byte cudaIpcMemHandle_t; hevmx_p2p_connect maps a peer's block (handle from another process, or the
 * peer VM itself when several VMs share a process in tests).  hevmx_ks_shard_p2p = the whole sharded op of this rank,
 * asynchronous: stage 1, push of the own digit rows into every peer + flag, wait, stage 2, push / wait of the rounded
 * special-limb rows, stage 3.  Ownership of targets is static: rank g owns the limbs [L*g/G, L*(g+1)/G) of the key level
 * (hevmx_p2p_targets gives the range at a level).  With HEVM_SHARD_RANK / HEVM_SHARD_WORLD set at initFullVM the VM
 * also STORES only those limbs of every key-switch key (key memory / G) and accepts only the sharded ops. */
void hevmx_p2p_setup(void *vm, int64_t rank, int64_t world, uint8_t *handle_out /*64 bytes or NULL*/);
void hevmx_p2p_connect(void *vm, int64_t peer, const uint8_t *handle /*64 bytes*/, void *peer_vm_same_process /*or NULL*/);
void hevmx_ks_shard_p2p(void *vm, int64_t opcode /*1 rotate, 8 mulcc*/, int64_t dst, int64_t lhs, int64_t rhs /*step | register*/);
/* the same op cut at its two exchanges (phase 1: stage 1 + digit push; 2: wait + stage 2 + row push; 3: wait + stage 3;
 * 0 = all): lets a single-GPU test that emulates the ranks with several VMs issue the phases in lockstep */
void hevmx_ks_shard_p2p_phase(void *vm, int64_t phase, int64_t opcode, int64_t dst, int64_t lhs, int64_t rhs);
double hevmx_p2p_bench(void *vm, int64_t words, int64_t reps); /* ms per push of `words` u64 into every peer (all ranks together) */
void hevmx_p2p_targets(void *vm, int64_t level, int64_t *tlo, int64_t *thi);
void hevmx_p2p_timing(void *vm, int on, double *out5 /*ms: stage1, digit exchange, stage2, row exchange, stage3*/);
/* --- key import (the .seal loader, dacapo_b200/seal_format.py): canonical residues in SEAL's layouts --- */
void hevmx_key_write(void *vm, int which /*0 sk, 1 pk, 2 relin, 3 galois*/, uint64_t elt, const uint64_t *in);
void hevmx_galois_clear(void *vm);

#ifdef __cplusplus
}
#endif
#endif
