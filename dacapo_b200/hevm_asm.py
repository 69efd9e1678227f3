"""Assembler / disassembler for the HEVM program container and the constants file.

File layout restated from the reference (include/hecate/Support/HEVMHeader.h:9-34,
writer lib/Dialect/CKKS/Transforms/EmitHEVM.cpp:31-119, reader
lib/Runtime/SEAL_HEVM.cpp:182-234; constants writer
lib/Dialect/Earth/Transforms/ElideConstant.cpp:42-53).  Little-endian:

    u32 magic=0x4845564D  u32 header_size=24  u64 arg_len  u64 res_len
    u64 config_body_len=40+8*(2A+3R)  u64 num_ops  u64 num_ct  u64 num_pt  u64 init_level
    u64 arg_scale[A] arg_level[A] res_scale[R] res_level[R] res_dst[R]
    {u16 opcode,dst,lhs,rhs}[num_ops]

Opcode numbers follow lib/Dialect/CKKS/IR/CKKSOps.td:69-222.
"""
import struct
from dataclasses import dataclass, field
from typing import List, Sequence

import numpy as np

MAGIC = 0x4845564D
ENCODE, ROTATE, NEGATE, RESCALE, MODSWITCH, UPSCALE, ADDCC, ADDCP, MULCC, MULCP, BOOTSTRAP = range(11)
PLACEHOLDER = 0xFFFF
OP_NAMES = {ENCODE: "encode", ROTATE: "rotate", NEGATE: "negate", RESCALE: "rescale", MODSWITCH: "modswitch",
            UPSCALE: "upscale", ADDCC: "addcc", ADDCP: "addcp", MULCC: "mulcc", MULCP: "mulcp",
            BOOTSTRAP: "bootstrap", PLACEHOLDER: "empty"}


@dataclass
class Program:
    arg_scale: List[int] = field(default_factory=list)
    arg_level: List[int] = field(default_factory=list)
    res_scale: List[int] = field(default_factory=list)
    res_level: List[int] = field(default_factory=list)
    res_dst: List[int] = field(default_factory=list)
    ops: List[tuple] = field(default_factory=list)  # (opcode, dst, lhs, rhs) as u16
    num_ct: int = 0
    num_pt: int = 0
    init_level: int = 13
    constants: List[np.ndarray] = field(default_factory=list)

    # ---- builder helpers ------------------------------------------------------
    def arg(self, scale_bits, level):
        self.arg_scale.append(scale_bits)
        self.arg_level.append(level)
        self.num_ct = max(self.num_ct, len(self.arg_scale))
        return len(self.arg_scale) - 1

    def result(self, reg, scale_bits=0, level=0):
        self.res_scale.append(scale_bits)
        self.res_level.append(level)
        self.res_dst.append(reg)

    def new_ct(self):
        self.num_ct += 1
        return self.num_ct - 1

    def new_pt(self):
        self.num_pt += 1
        return self.num_pt - 1

    def const(self, values):
        self.constants.append(np.asarray(values, dtype=np.float64).ravel())
        return len(self.constants) - 1

    def emit(self, opcode, dst, lhs=0, rhs=0):
        """Append one 8-byte op {u16 opcode, dst, lhs, rhs} (HEVMHeader.h:27-32).  Operands that do not fit 16 bits are
        an error, not wrapped: index 0xFFFF is the placeholder / all-ones marker, so a 65 536th register or constant
        would silently alias another one."""
        if opcode != PLACEHOLDER:
            for name, v in (("dst", dst), ("lhs", lhs), ("rhs", rhs)):
                if not 0 <= int(v) <= 0xFFFF:
                    raise ValueError(f"HEVM operand {name}={v} does not fit the 16-bit field of opcode {opcode}")
            if opcode != ENCODE and (int(dst) == 0xFFFF or int(lhs) == 0xFFFF):
                raise ValueError("register index 0xFFFF is reserved (tensor.empty placeholder, EmitHEVM.cpp:54-58)")
        self.ops.append((opcode & 0xFFFF, dst & 0xFFFF, lhs & 0xFFFF, rhs & 0xFFFF))
        return dst

    def encode(self, pt, const_idx, level, scale_bits):
        """opcode 0: rhs packs (level<<10)+scale (CKKSOps.td:75); lhs=-1 encodes all-ones."""
        if not 0 <= int(level) < 64 or not 0 <= int(scale_bits) < 1024:
            raise ValueError(f"encode: level {level} / scale {scale_bits} do not fit the 6 + 10 bit packing (CKKSOps.td:75)")
        if int(const_idx) >= 0xFFFF or int(pt) >= 0xFFFF:
            raise ValueError("encode: more than 65 534 constants / plaintext registers cannot be addressed by a 16-bit operand")
        return self.emit(ENCODE, pt, PLACEHOLDER if const_idx < 0 else const_idx, (level << 10) + scale_bits)

    def rotate(self, dst, src, offset):
        if not -32768 <= int(offset) <= 0xFFFF:
            raise ValueError(f"rotate offset {offset} does not fit 16 bits (SEAL_HEVM.cpp:269 reads an int16)")
        return self.emit(ROTATE, dst, src, offset & 0xFFFF)  # low 16 bits, runtime reads int16

    # ---- serialisation --------------------------------------------------------------
    def hevm_bytes(self) -> bytes:
        A, R = len(self.arg_scale), len(self.res_dst)
        out = struct.pack("<IIQQ", MAGIC, 24, A, R)
        out += struct.pack("<QQQQQ", 40 + 8 * (2 * A + 3 * R), len(self.ops), self.num_ct, self.num_pt, self.init_level)
        for arr in (self.arg_scale, self.arg_level, self.res_scale, self.res_level, self.res_dst):
            out += struct.pack(f"<{len(arr)}Q", *arr)
        out += np.asarray(self.ops, dtype="<u2").reshape(-1, 4).tobytes() if self.ops else b""
        return out

    def cst_bytes(self) -> bytes:
        out = [struct.pack("<q", len(self.constants))]
        for c in self.constants:
            out.append(struct.pack("<q", c.size))
            out.append(c.astype("<f8").tobytes())
        return b"".join(out)

    def save(self, cst_path, hevm_path):
        with open(cst_path, "wb") as f:
            f.write(self.cst_bytes())
        with open(hevm_path, "wb") as f:
            f.write(self.hevm_bytes())


def parse_hevm(data: bytes) -> Program:
    magic, hsize, A, R = struct.unpack_from("<IIQQ", data, 0)
    if magic != MAGIC:
        raise ValueError("bad HEVM magic")
    body_len, n_ops, n_ct, n_pt, init_level = struct.unpack_from("<QQQQQ", data, 24)
    off = 64
    arrs = []
    for n in (A, A, R, R, R):
        arrs.append(list(struct.unpack_from(f"<{n}Q", data, off)))
        off += 8 * n
    if off != hsize + body_len:
        raise ValueError("inconsistent config_body_length")
    ops = np.frombuffer(data, dtype="<u2", count=4 * n_ops, offset=off).reshape(-1, 4)
    p = Program(*arrs, ops=[tuple(int(x) for x in o) for o in ops], num_ct=n_ct, num_pt=n_pt, init_level=init_level)
    return p


def parse_cst(data: bytes) -> List[np.ndarray]:
    (n,) = struct.unpack_from("<q", data, 0)
    off, out = 8, []
    for _ in range(n):
        (ln,) = struct.unpack_from("<q", data, off)
        off += 8
        out.append(np.frombuffer(data, dtype="<f8", count=ln, offset=off).copy())
        off += 8 * ln
    return out


def disassemble(p: Program) -> str:
    lines = []
    for i, (oc, d, l, r) in enumerate(p.ops):
        name = OP_NAMES.get(oc, f"op{oc}")
        if oc == ENCODE:
            lines.append(f"{i:6d}  encode   p{d} <- const[{l if l != PLACEHOLDER else 'ones'}] level={r >> 10} scale=2^{r & 0x3FF}")
        elif oc == ROTATE:
            lines.append(f"{i:6d}  rotate   c{d} <- c{l} by {np.int16(np.uint16(r))}")
        else:
            lines.append(f"{i:6d}  {name:8s} {d} <- {l}, {r}")
    return "\n".join(lines)
