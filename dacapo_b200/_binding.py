"""ctypes prototypes of lib*_HEVM.so (the same block binds any library exporting the reference ABI).

Mirrors the prototype block of the reference driver
(reference: python/hecate/hecate/runner.py:34-71) and adds the hevmx_* hooks of
include/hevm_ext.h.  Nothing here computes: it only declares signatures.
"""
import ctypes as C
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
B200_LIB = Path(__file__).resolve().parent / "libB200_HEVM.so"

ABI_SYMBOLS = [
    "create_context", "initFullVM", "initClientVM", "initServerVM", "load", "loadClient",
    "encrypt", "decrypt", "decrypt_result", "getResIdx", "getCtxt", "preprocess", "run",
    "getArgLen", "getResLen", "setDebug", "setToGPU", "printMem",
]
EXT_SYMBOLS = [
    "hevmx_param", "hevmx_primes", "hevmx_roots", "hevmx_resize", "hevmx_ct_info", "hevmx_ct_read",
    "hevmx_ct_write", "hevmx_pt_info", "hevmx_pt_read", "hevmx_pt_write", "hevmx_exec", "hevmx_sync",
    "hevmx_ntt", "hevmx_encode", "hevmx_decode", "hevmx_decrypt_to_pt", "hevmx_encrypt_pt",
    "hevmx_set_enc_counter", "hevmx_key_read", "hevmx_galois_elt", "hevmx_backend",
]

B200_ONLY_SYMBOLS = ["hevmx_ntt_bench", "hevmx_timer", "hevmx_profile", "hevmx_profile_read", "hevmx_profiler_range",
                     "hevmx_ks_shard_stage", "hevmx_mulcc_shard_stage", "hevmx_dev_ptr", "hevmx_stream", "hevmx_exec_batch",
                     "hevmx_p2p_setup", "hevmx_p2p_connect", "hevmx_ks_shard_p2p", "hevmx_ks_shard_p2p_phase", "hevmx_p2p_targets", "hevmx_p2p_timing", "hevmx_p2p_bench",
                     "hevmx_key_write", "hevmx_galois_clear"]

_u64p = C.POINTER(C.c_uint64)
_f64p = C.POINTER(C.c_double)
_i64p = C.POINTER(C.c_int64)


def bind(path):
    """dlopen `path` and attach argtypes/restype to every exported entry point."""
    path = Path(path)
    if not path.is_file():
        raise FileNotFoundError(f"{path} is missing - run `python -c 'import __graft_entry__ as g; g.build()'`")
    lw = C.CDLL(str(path))
    # --- the reference's 18 symbols (runner.py:34-71) ---
    lw.initFullVM.argtypes = [C.c_char_p, C.c_bool]
    lw.initFullVM.restype = C.c_void_p
    lw.initClientVM.argtypes = [C.c_char_p]
    lw.initClientVM.restype = C.c_void_p
    lw.initServerVM.argtypes = [C.c_char_p]
    lw.initServerVM.restype = C.c_void_p
    lw.create_context.argtypes = [C.c_char_p]
    lw.load.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p]
    lw.loadClient.argtypes = [C.c_void_p, C.c_void_p]
    lw.getArgLen.argtypes = [C.c_void_p]
    lw.getArgLen.restype = C.c_int64
    lw.getResLen.argtypes = [C.c_void_p]
    lw.getResLen.restype = C.c_int64
    lw.encrypt.argtypes = [C.c_void_p, C.c_int64, _f64p, C.c_int]
    lw.decrypt.argtypes = [C.c_void_p, C.c_int64, _f64p]
    lw.decrypt_result.argtypes = [C.c_void_p, C.c_int64, _f64p]
    lw.getResIdx.argtypes = [C.c_void_p, C.c_int64]
    lw.getResIdx.restype = C.c_int64
    lw.getCtxt.argtypes = [C.c_void_p, C.c_int64]
    lw.getCtxt.restype = C.c_void_p
    lw.preprocess.argtypes = [C.c_void_p]
    lw.run.argtypes = [C.c_void_p]
    lw.setDebug.argtypes = [C.c_void_p, C.c_bool]
    lw.setToGPU.argtypes = [C.c_void_p, C.c_bool]
    lw.printMem.argtypes = [C.c_void_p]
    # --- hevmx_* hooks (include/hevm_ext.h) ---
    lw.hevmx_param.argtypes = [C.c_void_p, C.c_int]
    lw.hevmx_param.restype = C.c_int64
    lw.hevmx_primes.argtypes = [C.c_void_p, _u64p]
    lw.hevmx_roots.argtypes = [C.c_void_p, _u64p]
    lw.hevmx_resize.argtypes = [C.c_void_p, C.c_int64, C.c_int64]
    lw.hevmx_ct_info.argtypes = [C.c_void_p, C.c_int64, _i64p, _f64p]
    lw.hevmx_ct_read.argtypes = [C.c_void_p, C.c_int64, _u64p]
    lw.hevmx_ct_write.argtypes = [C.c_void_p, C.c_int64, _u64p, C.c_int64, C.c_double]
    lw.hevmx_pt_info.argtypes = [C.c_void_p, C.c_int64, _i64p, _f64p]
    lw.hevmx_pt_read.argtypes = [C.c_void_p, C.c_int64, _u64p]
    lw.hevmx_pt_write.argtypes = [C.c_void_p, C.c_int64, _u64p, C.c_int64, C.c_double]
    lw.hevmx_exec.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int64]
    lw.hevmx_sync.argtypes = [C.c_void_p]
    lw.hevmx_ntt.argtypes = [C.c_void_p, _u64p, C.c_int64, C.c_int64, C.c_int]
    lw.hevmx_encode.argtypes = [C.c_void_p, C.c_int64, _f64p, C.c_int64, C.c_int64, C.c_int64]
    lw.hevmx_decode.argtypes = [C.c_void_p, C.c_int64, _f64p]
    lw.hevmx_decrypt_to_pt.argtypes = [C.c_void_p, C.c_int64, C.c_int64]
    lw.hevmx_encrypt_pt.argtypes = [C.c_void_p, C.c_int64, C.c_int64]
    lw.hevmx_set_enc_counter.argtypes = [C.c_void_p, C.c_uint64]
    lw.hevmx_key_read.argtypes = [C.c_void_p, C.c_int, C.c_uint64, _u64p]
    lw.hevmx_key_read.restype = C.c_int64
    lw.hevmx_galois_elt.argtypes = [C.c_void_p, C.c_int64]
    lw.hevmx_galois_elt.restype = C.c_int64
    lw.hevmx_backend.argtypes = []
    lw.hevmx_backend.restype = C.c_char_p
    if hasattr(lw, "hevmx_timer"):  # libB200_HEVM.so only
        lw.hevmx_ntt_bench.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_int64]
        lw.hevmx_ntt_bench.restype = C.c_double
        lw.hevmx_timer.argtypes = [C.c_void_p, C.c_int]
        lw.hevmx_timer.restype = C.c_double
        lw.hevmx_profile.argtypes = [C.c_void_p, C.c_int]
        lw.hevmx_profiler_range.argtypes = [C.c_void_p, C.c_int]
        lw.hevmx_profile_read.argtypes = [C.c_void_p, C.c_int, _f64p, _i64p]
        lw.hevmx_profile_read.restype = C.c_char_p
        lw.hevmx_ks_shard_stage.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int64]
        lw.hevmx_mulcc_shard_stage.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int64]
        lw.hevmx_dev_ptr.argtypes = [C.c_void_p, C.c_int64]
        lw.hevmx_dev_ptr.restype = C.c_void_p
        lw.hevmx_exec_batch.argtypes = [C.c_void_p, C.c_int64, C.c_int64, _i64p, _i64p, _i64p]
        lw.hevmx_p2p_setup.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_char_p]
        lw.hevmx_p2p_connect.argtypes = [C.c_void_p, C.c_int64, C.c_char_p, C.c_void_p]
        lw.hevmx_ks_shard_p2p.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int64]
        lw.hevmx_ks_shard_p2p_phase.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int64]
        lw.hevmx_p2p_bench.argtypes = [C.c_void_p, C.c_int64, C.c_int64]
        lw.hevmx_p2p_bench.restype = C.c_double
        lw.hevmx_p2p_targets.argtypes = [C.c_void_p, C.c_int64, _i64p, _i64p]
        lw.hevmx_p2p_timing.argtypes = [C.c_void_p, C.c_int, _f64p]
        lw.hevmx_key_write.argtypes = [C.c_void_p, C.c_int, C.c_uint64, _u64p]
        lw.hevmx_galois_clear.argtypes = [C.c_void_p]
        lw.hevmx_stream.argtypes = [C.c_void_p]
        lw.hevmx_stream.restype = C.c_void_p
    return lw
