"""Reader / writer for Microsoft SEAL 4.0 binary serialisation (`*.seal`): SURVEY.md 8(f) row 4.

The reference writes its key directory with SEAL's `save()` (reference: lib/Runtime/SEAL_HEVM.cpp:56-88 --
`parm.seal` EncryptionParameters, `pub.seal` PublicKey, `sec.seal` SecretKey, `relin.seal` RelinKeys, `gal.seal`
GaloisKeys) and reads it back with `load()` (91-129).  This module restates that container so that a key directory (or
a ciphertext) produced by a REAL SEAL 4.0 elsewhere can be dropped next to this backend and loaded into the GPU VM
(`runner.HEVM(seal_dir=...)` / `load_seal_keys`), which is what turns "bit-exact against our restatement" into
"bit-exact against SEAL"; and so that keys / ciphertexts of this backend can be exported for a SEAL installation to open.

SEAL is not vendored in /root/reference and cannot be installed offline, so the layout below is restated from SEAL 4.0's
published sources (native/src/seal/serialization.h, encryptionparams.cpp, plaintext.cpp, ciphertext.cpp, secretkey.h,
publickey.h, kswitchkeys.cpp, dynarray.h, modulus.cpp) and is UNPINNED until a SEAL-produced file is available: the
round-trip tests pin reader and writer to each other and to the documented header constants only.

Every serialised object = a 16-byte header followed by the payload:

    u16 magic = 0xA15E | u8 header_size = 16 | u8 version_major = 4 | u8 version_minor = 0 | u8 compr_mode | u16 reserved = 0
    u64 size            (total bytes including the header)
    compr_mode: 0 none, 1 zlib (deflate stream), 2 zstd.  The payload of a compressed object is the compressed image of
    exactly the bytes an uncompressed save would have produced.

Payloads (little-endian, no padding):
    Modulus               u64 value
    DynArray<u64>         u64 count | count x u64                    (always wrapped in its own header)
    EncryptionParameters  u8 scheme (2 = CKKS) | u64 poly_modulus_degree | u64 coeff_modulus_size |
                          coeff_modulus_size x [header + Modulus] | [header + Modulus] plain_modulus
    Plaintext             4 x u64 parms_id | u64 coeff_count | f64 scale | [header + DynArray]
    Ciphertext            4 x u64 parms_id | u8 is_ntt_form | u64 size | u64 poly_modulus_degree | u64 coeff_modulus_size |
                          u64 correction_factor | f64 scale | [header + DynArray]   (size*coeff_modulus_size*N words,
                          poly-major then limb-major: exactly this backend's [2][l][N] layout; seeded ciphertexts --
                          a marker in the second polynomial -- are rejected: they need SEAL's PRNG to expand)
    SecretKey             [header + Plaintext]  (key level, NTT form, coeff_count = L*N)
    PublicKey             [header + Ciphertext] (key level)
    KSwitchKeys           4 x u64 parms_id | u64 dim1 | dim1 x ( u64 dim2 | dim2 x [header + PublicKey payload] )
                          RelinKeys: dim1 = 1 (index 0 = s^2); GaloisKeys: dim1 = index(elt)+1 slots, index = (elt-1)/2,
                          empty slots have dim2 = 0.  keys[i][J] = digit J, a size-2 ciphertext at key level.
"""
import struct
import zlib
from typing import Dict, List, Tuple

import numpy as np

MAGIC, HEADER_SIZE, VERSION = 0xA15E, 16, (4, 0)
COMPR_NONE, COMPR_ZLIB, COMPR_ZSTD = 0, 1, 2
SCHEME_CKKS = 2
SEED_MARKER = 0xFFFFFFFFFFFFFFFF  # first word of a seeded second polynomial (util/rlwe.h)


class SealFormatError(ValueError):
    pass


# ------------------------------------------------------------------------------------------------ container
def wrap(payload: bytes, compr_mode: int = COMPR_NONE) -> bytes:
    body = payload
    if compr_mode == COMPR_ZLIB:
        body = zlib.compress(payload)
    elif compr_mode == COMPR_ZSTD:
        import pyarrow as pa
        body = pa.Codec("zstd").compress(payload, asbytes=True)
    elif compr_mode != COMPR_NONE:
        raise SealFormatError(f"unknown compr_mode {compr_mode}")
    return struct.pack("<HBBBBHQ", MAGIC, HEADER_SIZE, VERSION[0], VERSION[1], compr_mode, 0, HEADER_SIZE + len(body)) + body


def unwrap(buf: bytes, off: int = 0) -> Tuple[bytes, int]:
    """-> (uncompressed payload, offset just past this object)."""
    if len(buf) - off < HEADER_SIZE:
        raise SealFormatError("truncated SEAL header")
    magic, hsize, vmaj, vmin, compr, reserved, size = struct.unpack_from("<HBBBBHQ", buf, off)
    if magic != MAGIC or hsize != HEADER_SIZE:
        raise SealFormatError(f"not a SEAL object at offset {off} (magic {magic:#x}, header size {hsize})")
    if vmaj not in (3, 4):
        raise SealFormatError(f"unsupported SEAL version {vmaj}.{vmin}")
    if size < HEADER_SIZE or off + size > len(buf):
        raise SealFormatError("SEAL header size field out of range")
    body = bytes(buf[off + HEADER_SIZE:off + size])
    if compr == COMPR_ZLIB:
        body = zlib.decompress(body)
    elif compr == COMPR_ZSTD:
        body = _zstd_decompress(body)
    elif compr != COMPR_NONE:
        raise SealFormatError(f"unknown compr_mode {compr}")
    return body, off + size


def _zstd_decompress(body: bytes) -> bytes:
    try:
        import zstandard
        return zstandard.ZstdDecompressor().decompressobj().decompress(body)
    except ImportError:
        pass
    import pyarrow as pa  # pyarrow ships a zstd codec; it wants the decompressed size, which the zstd frame header carries
    size = _zstd_content_size(body)
    return pa.Codec("zstd").decompress(body, decompressed_size=size, asbytes=True)


def _zstd_content_size(frame: bytes) -> int:
    if frame[:4] != b"\x28\xb5\x2f\xfd":
        raise SealFormatError("bad zstd frame magic")
    fhd = frame[4]
    fcs_flag, single, dict_flag = fhd >> 6, (fhd >> 5) & 1, fhd & 3
    pos = 5 + (0 if single else 1) + (0, 1, 2, 4)[dict_flag]
    nbytes = (1 if single else 0, 2, 4, 8)[fcs_flag]
    if nbytes == 0:
        raise SealFormatError("zstd frame without content size (streamed): install `zstandard` to read it")
    v = int.from_bytes(frame[pos:pos + nbytes], "little")
    return v + 256 if nbytes == 2 else v


# ------------------------------------------------------------------------------------------------ leaf objects
def _dynarray(words: np.ndarray) -> bytes:
    w = np.ascontiguousarray(words, dtype="<u8").ravel()
    return wrap(struct.pack("<Q", w.size) + w.tobytes())


def _read_dynarray(buf: bytes, off: int) -> Tuple[np.ndarray, int]:
    body, nxt = unwrap(buf, off)
    (n,) = struct.unpack_from("<Q", body, 0)
    if len(body) != 8 + 8 * n:
        raise SealFormatError("DynArray length mismatch")
    return np.frombuffer(body, dtype="<u8", count=n, offset=8).copy(), nxt


def parms_id_placeholder(primes: List[int], n: int) -> Tuple[int, int, int, int]:
    """SEAL's parms_id is a BLAKE2b-256 hash of (scheme, N, moduli, plain modulus); hashlib has blake2b, and the input
    block is [scheme, N, q_0 .. q_{k-1}, plain_modulus(=0)] as u64 words (encryptionparams.cpp compute_parms_id)."""
    import hashlib
    words = np.array([SCHEME_CKKS, n] + [int(q) for q in primes] + [0], dtype="<u8").tobytes()
    h = hashlib.blake2b(words, digest_size=32).digest()
    return struct.unpack("<4Q", h)


# ------------------------------------------------------------------------------------------------ typed objects
def write_parms(n: int, primes: List[int], compr_mode: int = COMPR_NONE) -> bytes:
    p = struct.pack("<BQQ", SCHEME_CKKS, n, len(primes))
    for q in primes:
        p += wrap(struct.pack("<Q", int(q)))
    p += wrap(struct.pack("<Q", 0))  # plain_modulus (unused by CKKS)
    return wrap(p, compr_mode)


def read_parms(buf: bytes) -> Dict:
    body, _ = unwrap(buf)
    scheme, n, k = struct.unpack_from("<BQQ", body, 0)
    off, primes = 17, []
    for _ in range(k):
        m, off = unwrap(body, off)
        primes.append(struct.unpack("<Q", m)[0])
    pm, off = unwrap(body, off)
    if scheme != SCHEME_CKKS:
        raise SealFormatError(f"scheme {scheme} is not CKKS")
    return {"poly_modulus_degree": n, "coeff_modulus": primes, "plain_modulus": struct.unpack("<Q", pm)[0]}


def _ciphertext_payload(data: np.ndarray, n: int, nlimbs: int, parms_id, scale: float = 1.0, is_ntt: bool = True) -> bytes:
    size = data.size // (n * nlimbs)
    return (struct.pack("<4QBQQQQd", *parms_id, 1 if is_ntt else 0, size, n, nlimbs, 1, scale) + _dynarray(data))


def _read_ciphertext_payload(body: bytes, off: int = 0) -> Tuple[Dict, int]:
    pid = struct.unpack_from("<4Q", body, off)
    is_ntt, size, n, nlimbs, corr, scale = struct.unpack_from("<BQQQQd", body, off + 32)
    data, nxt = _read_dynarray(body, off + 32 + 41)
    if data.size != size * n * nlimbs:
        if data.size and size == 2 and data.size < size * n * nlimbs and data[n * nlimbs] == SEED_MARKER:
            raise SealFormatError("seeded (compressed-randomness) ciphertext / public key: expand it with SEAL before exporting")
        raise SealFormatError("ciphertext data length mismatch")
    return {"parms_id": pid, "is_ntt_form": bool(is_ntt), "size": size, "poly_modulus_degree": n, "coeff_modulus_size": nlimbs,
            "correction_factor": corr, "scale": scale, "data": data.reshape(size, nlimbs, n)}, nxt


def write_ciphertext(ct: np.ndarray, scale: float, parms_id, compr_mode: int = COMPR_NONE) -> bytes:
    """ct: u64[size][limbs][N], NTT form."""
    size, nlimbs, n = ct.shape
    return wrap(_ciphertext_payload(ct, n, nlimbs, parms_id, scale), compr_mode)


def read_ciphertext(buf: bytes) -> Dict:
    body, _ = unwrap(buf)
    return _read_ciphertext_payload(body)[0]


def write_public_key(pk: np.ndarray, parms_id, compr_mode: int = COMPR_NONE) -> bytes:
    """pk: u64[2][L][N] at key level, NTT form."""
    _, nlimbs, n = pk.shape
    return wrap(wrap(_ciphertext_payload(pk, n, nlimbs, parms_id)), compr_mode)


def read_public_key(buf: bytes) -> Dict:
    body, _ = unwrap(buf)
    inner, _ = unwrap(body)
    return _read_ciphertext_payload(inner)[0]


def write_secret_key(sk: np.ndarray, parms_id, compr_mode: int = COMPR_NONE) -> bytes:
    """sk: u64[L][N] at key level, NTT form."""
    pt = struct.pack("<4QQd", *parms_id, sk.size, 1.0) + _dynarray(sk)
    return wrap(wrap(pt), compr_mode)


def read_secret_key(buf: bytes) -> Dict:
    body, _ = unwrap(buf)
    inner, _ = unwrap(body)
    pid = struct.unpack_from("<4Q", inner, 0)
    count, scale = struct.unpack_from("<Qd", inner, 32)
    data, _ = _read_dynarray(inner, 48)
    if data.size != count:
        raise SealFormatError("secret key length mismatch")
    return {"parms_id": pid, "coeff_count": count, "data": data}


def write_kswitch_keys(keys: Dict[int, np.ndarray], parms_id, compr_mode: int = COMPR_NONE) -> bytes:
    """keys: {slot index: u64[digits][2][L][N]}.  RelinKeys: {0: relin}; GaloisKeys: {(elt-1)//2: key}."""
    dim1 = max(keys) + 1 if keys else 0
    p = struct.pack("<4QQ", *parms_id, dim1)
    for i in range(dim1):
        k = keys.get(i)
        if k is None:
            p += struct.pack("<Q", 0)
            continue
        p += struct.pack("<Q", k.shape[0])
        for J in range(k.shape[0]):
            _, nlimbs, n = k[J].shape
            p += wrap(wrap(_ciphertext_payload(k[J], n, nlimbs, parms_id)))
    return wrap(p, compr_mode)


def read_kswitch_keys(buf: bytes) -> Dict:
    body, _ = unwrap(buf)
    pid = struct.unpack_from("<4Q", body, 0)
    (dim1,) = struct.unpack_from("<Q", body, 32)
    off, keys = 40, {}
    for i in range(dim1):
        (dim2,) = struct.unpack_from("<Q", body, off)
        off += 8
        digits = []
        for _ in range(dim2):
            outer, off = unwrap(body, off)
            inner, _ = unwrap(outer)
            digits.append(_read_ciphertext_payload(inner)[0]["data"])
        if digits:
            keys[i] = np.stack(digits)  # [digits][2][L][N]
    return {"parms_id": pid, "keys": keys}


def galois_index(elt: int) -> int:
    return (elt - 1) >> 1  # GaloisKeys::get_index


# ------------------------------------------------------------------------------------------------ key directory
FILES = {"parm": "parm.seal", "pub": "pub.seal", "sec": "sec.seal", "relin": "relin.seal", "gal": "gal.seal"}


def write_key_dir(path, n: int, primes: List[int], sk: np.ndarray, pk: np.ndarray, relin: np.ndarray, galois: Dict[int, np.ndarray],
                  compr_mode: int = COMPR_NONE):
    """The five files of SEAL_HEVM.cpp:56-88 from this backend's key material (canonical residues, NTT form)."""
    from pathlib import Path
    d = Path(path)
    d.mkdir(parents=True, exist_ok=True)
    pid = parms_id_placeholder(primes, n)
    (d / FILES["parm"]).write_bytes(write_parms(n, primes, compr_mode))
    (d / FILES["sec"]).write_bytes(write_secret_key(sk, pid, compr_mode))
    (d / FILES["pub"]).write_bytes(write_public_key(pk, pid, compr_mode))
    (d / FILES["relin"]).write_bytes(write_kswitch_keys({0: relin}, pid, compr_mode))
    (d / FILES["gal"]).write_bytes(write_kswitch_keys({galois_index(e): k for e, k in galois.items()}, pid, compr_mode))


def read_key_dir(path) -> Dict:
    """-> {"n", "primes", "sk" [L][N], "pk" [2][L][N], "relin" [L-1][2][L][N], "galois" {elt: [L-1][2][L][N]}}"""
    from pathlib import Path
    d = Path(path)
    parms = read_parms((d / FILES["parm"]).read_bytes())
    n, primes = parms["poly_modulus_degree"], parms["coeff_modulus"]
    L = len(primes)
    sk = read_secret_key((d / FILES["sec"]).read_bytes())["data"]
    if sk.size != L * n:
        raise SealFormatError("secret key is not at key level")
    pk = read_public_key((d / FILES["pub"]).read_bytes())["data"]
    relin = read_kswitch_keys((d / FILES["relin"]).read_bytes())["keys"]
    gal = read_kswitch_keys((d / FILES["gal"]).read_bytes())["keys"]
    if 0 not in relin:
        raise SealFormatError("relin.seal holds no key for s^2")
    return {"n": n, "primes": primes, "sk": sk.reshape(L, n), "pk": pk, "relin": relin[0],
            "galois": {2 * i + 1: k for i, k in gal.items()}}


# ------------------------------------------------------------------------------------------------ VM <-> key directory
def export_vm_keys(lib, vm, path, compr_mode: int = COMPR_NONE):
    """Write the key material of a libB200_HEVM.so VM (hevmx_key_read: canonical residues) as a SEAL key directory."""
    import ctypes as C
    u64p = C.POINTER(C.c_uint64)
    logn, L = lib.hevmx_param(vm, 0), lib.hevmx_param(vm, 1)
    n = 1 << logn
    primes = np.zeros(L, dtype=np.uint64)
    lib.hevmx_primes(vm, primes.ctypes.data_as(u64p))

    def read(which, elt=0):
        w = lib.hevmx_key_read(vm, which, elt, None)
        if w < 0:
            return None
        out = np.zeros(w, dtype=np.uint64)
        lib.hevmx_key_read(vm, which, elt, out.ctypes.data_as(u64p))
        return out

    gal = {}
    m = 2 * n
    elts, pos, neg = [m - 1], 3, pow(3, -1, m)
    for _ in range(logn - 1):  # GaloisTool::get_elts_all: 3^(+-2^i) and the conjugation
        elts += [pos, neg]
        pos, neg = pos * pos % m, neg * neg % m
    for e in elts:
        k = read(3, e)
        if k is not None:
            gal[e] = k.reshape(L - 1, 2, L, n)
    write_key_dir(path, n, [int(q) for q in primes], read(0).reshape(L, n), read(1).reshape(2, L, n), read(2).reshape(L - 1, 2, L, n), gal, compr_mode)


def load_seal_keys(lib, vm, path):
    """Replace the VM's keys with those of a SEAL key directory (SEAL_HEVM.cpp:91-129 `loadSEAL`).  The directory's
    parameters must be this VM's ring: same degree and the same prime chain (SEAL's CoeffModulus::Create order)."""
    import ctypes as C
    u64p = C.POINTER(C.c_uint64)
    kd = read_key_dir(path)
    logn, L = lib.hevmx_param(vm, 0), lib.hevmx_param(vm, 1)
    primes = np.zeros(L, dtype=np.uint64)
    lib.hevmx_primes(vm, primes.ctypes.data_as(u64p))
    if kd["n"] != 1 << logn or [int(q) for q in primes] != [int(q) for q in kd["primes"]]:
        raise SealFormatError("the key directory was made for other encryption parameters than this VM's")

    def put(which, arr, elt=0):
        a = np.ascontiguousarray(arr, dtype=np.uint64)
        lib.hevmx_key_write(vm, which, elt, a.ctypes.data_as(u64p))

    if kd["pk"].shape != (2, L, kd["n"]) or kd["relin"].shape != (L - 1, 2, L, kd["n"]):
        raise SealFormatError("unexpected key shapes")
    put(0, kd["sk"])
    put(1, kd["pk"])
    put(2, kd["relin"])
    lib.hevmx_galois_clear(vm)
    for elt, k in kd["galois"].items():
        put(3, k, elt)
    return kd
