/* =============================================================================
 * include/hevm_abi.h -- the drop-in C ABI of libB200_HEVM.so
 *
 * Exactly the 18 `extern "C"` symbols the reference HEVM runtime exports
 * (reference: lib/Runtime/SEAL_HEVM.cpp:404-504) and that the reference's Python
 * driver binds through ctypes (reference: python/hecate/hecate/runner.py:34-71).
 * Plain pointers and sizes only; no torch / CUDA types cross this boundary.
 *
 * Ownership / errors: init*VM returns a heap object that is never
 * freed by the caller (the reference has no destroy symbol); `dat` buffers are
 * caller-owned; there are no error returns -- unsupported operations and CUDA
 * failures abort() with a message (the reference asserts / lets SEAL throw).
 * Every call is synchronous: results are complete (device-synchronised) on return.
 * ========================================================================== */
#ifndef HEVM_ABI_H
#define HEVM_ABI_H
#include <stdbool.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* SEAL_HEVM.cpp:421 / 44-89 -- write the parameter + key-seed file into `dir`.
 * Ring geometry defaults to the reference's compile-time constants N=2^15, L=14
 * 60-bit primes (SEAL_HEVM.cpp:39-40,48-53); HEVM_LOGN / HEVM_NUM_PRIMES /
 * HEVM_SEED environment variables override them (own extension). */
void create_context(char *dir);

/* SEAL_HEVM.cpp:405-409 (+ HEAAN_HEVM.cpp:470 for the `device` flag that
 * runner.py:196-198 always passes).  Generates sk/pk/relin/Galois keys on the GPU
 * from the seed and keeps them resident in HBM. */
void *initFullVM(char *dir, bool device);
void *initClientVM(char *dir); /* SEAL_HEVM.cpp:410-414 (vestigial split: == full) */
void *initServerVM(char *dir); /* SEAL_HEVM.cpp:415-419 */

void load(void *vm, char *constant, char *vmfile); /* SEAL_HEVM.cpp:424-428 (.cst + .hevm) */
void loadClient(void *vm, void *is);               /* SEAL_HEVM.cpp:431-436 (aborts: vestigial) */

void encrypt(void *vm, int64_t i, double *dat, int len); /* SEAL_HEVM.cpp:439-445 */
void decrypt(void *vm, int64_t i, double *dat);          /* SEAL_HEVM.cpp:446-455 (writes N/2 doubles) */
void decrypt_result(void *vm, int64_t i, double *dat);   /* SEAL_HEVM.cpp:458-461 */
int64_t getResIdx(void *vm, int64_t i);                  /* SEAL_HEVM.cpp:464-467 */
void *getCtxt(void *vm, int64_t id);                     /* SEAL_HEVM.cpp:470-473 (borrowed) */

void preprocess(void *vm); /* SEAL_HEVM.cpp:475-478 -> 242-254: encode every opcode-0 constant */
void run(void *vm);        /* SEAL_HEVM.cpp:479-482 -> 336-401: the hot loop */
int64_t getArgLen(void *vm); /* SEAL_HEVM.cpp:483-486 */
int64_t getResLen(void *vm); /* SEAL_HEVM.cpp:487-490 */
void setDebug(void *vm, bool enable); /* SEAL_HEVM.cpp:491-494 */
void setToGPU(void *vm, bool ongpu);  /* SEAL_HEVM.cpp:495-497 (no-op here: always on GPU) */
void printMem(void *vm);              /* SEAL_HEVM.cpp:498-503 / HEAAN_HEVM.cpp:100-107 */

#ifdef __cplusplus
}
#endif
#endif
