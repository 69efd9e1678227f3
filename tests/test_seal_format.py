"""CPU tier: the SEAL 4.0 `.seal` container (dacapo_b200/seal_format.py; reference writes / reads these files at
lib/Runtime/SEAL_HEVM.cpp:56-88 / 91-129).  No SEAL-produced file exists offline, so the reader is pinned to the documented
header constants with a hand-assembled file and to the writer by round trips (plain / zlib / zstd)."""
import struct
import zlib

import numpy as np
import pytest

from dacapo_b200 import seal_format as sf

N, PRIMES = 16, [97, 193, 257]


def test_hand_assembled_header_and_parameters():
    """Byte-for-byte: magic 0xA15E, header size 16, version 4.0, compr_mode, reserved, u64 total size; then scheme = CKKS (2),
    degree, modulus count and one wrapped Modulus per prime."""
    def obj(payload, compr=0):
        return bytes([0x5E, 0xA1, 0x10, 0x04, 0x00, compr, 0x00, 0x00]) + struct.pack("<Q", 16 + len(payload)) + payload
    body = bytes([2]) + struct.pack("<QQ", N, len(PRIMES)) + b"".join(obj(struct.pack("<Q", q)) for q in PRIMES) + obj(struct.pack("<Q", 0))
    blob = obj(body)
    assert blob == sf.write_parms(N, PRIMES)
    p = sf.read_parms(blob)
    assert p == {"poly_modulus_degree": N, "coeff_modulus": PRIMES, "plain_modulus": 0}
    # the same object, deflate-compressed as SEAL's compr_mode zlib would store it
    z = zlib.compress(body)
    assert sf.read_parms(obj(z, compr=1)) == p
    with pytest.raises(sf.SealFormatError):
        sf.read_parms(b"\x00" * 40)
    with pytest.raises(sf.SealFormatError):
        sf.read_parms(blob[:-3])  # size field larger than the buffer


@pytest.mark.parametrize("compr", [sf.COMPR_NONE, sf.COMPR_ZLIB, sf.COMPR_ZSTD])
def test_key_directory_round_trip(tmp_path, compr):
    rng = np.random.default_rng(3)
    L = len(PRIMES)
    def res(*shape):
        a = np.zeros(shape + (L, N), dtype=np.uint64)
        for i, q in enumerate(PRIMES):
            a[..., i, :] = rng.integers(0, q, size=shape + (N,), dtype=np.uint64)
        return a
    sk, pk, relin = res(), res(2), res(L - 1, 2)
    gal = {3: res(L - 1, 2), 2 * N - 1: res(L - 1, 2), 11: res(L - 1, 2)}
    sf.write_key_dir(tmp_path, N, PRIMES, sk, pk, relin, gal, compr)
    assert sorted(p.name for p in tmp_path.iterdir()) == ["gal.seal", "parm.seal", "pub.seal", "relin.seal", "sec.seal"]
    kd = sf.read_key_dir(tmp_path)
    assert kd["n"] == N and kd["primes"] == PRIMES
    assert np.array_equal(kd["sk"], sk) and np.array_equal(kd["pk"], pk) and np.array_equal(kd["relin"], relin)
    assert sorted(kd["galois"]) == sorted(gal) and all(np.array_equal(kd["galois"][e], gal[e]) for e in gal)
    # GaloisKeys index = (elt - 1) / 2: elt 3 -> slot 1, elt 11 -> slot 5, elt 2N-1 -> slot N-1; empty slots carry dim2 = 0
    body, _ = sf.unwrap((tmp_path / "gal.seal").read_bytes())
    assert struct.unpack_from("<Q", body, 32)[0] == N


def test_ciphertext_round_trip_and_seeded_rejection():
    rng = np.random.default_rng(4)
    ct = rng.integers(0, 97, size=(2, 2, N), dtype=np.uint64)
    pid = sf.parms_id_placeholder(PRIMES[:2], N)
    back = sf.read_ciphertext(sf.write_ciphertext(ct, 2.0 ** 40, pid))
    assert np.array_equal(back["data"], ct) and back["scale"] == 2.0 ** 40 and back["is_ntt_form"] and back["parms_id"] == pid
    # a seeded ciphertext (second polynomial replaced by a marker + seed) must be refused, not mis-read
    short = np.concatenate([ct[0].ravel(), np.array([sf.SEED_MARKER, 1, 2, 3], dtype=np.uint64)])
    payload = struct.pack("<4QBQQQQd", *pid, 1, 2, N, 2, 1, 1.0) + sf._dynarray(short)
    with pytest.raises(sf.SealFormatError):
        sf.read_ciphertext(sf.wrap(payload))
