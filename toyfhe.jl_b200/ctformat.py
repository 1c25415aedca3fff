"""On-disk container for RNS ciphertext batches.  The reference has no ciphertext format
(only BSON for the plaintext Flux model, examples/encrypted_mnist/infer.jl:38); this is the engine's own: the exact
buffer the C-ABI consumes, prefixed by what is needed to rebuild the ring it lives in.

    offset  size   field
    0       8      magic  b"TFB2CT\\0\\0"
    8       4      version (1)                      u32 LE
    12      4      flags: bit 0 = dual (NTT) domain, bit 1 = CKKS scale present
    16      4      N   (ring degree)                u32
    20      4      L   (RNS primes)                 u32
    24      4      components per ciphertext        u32
    28      4      reserved (0)
    32      8      batch                            u64
    40      8      scale (float64; 0.0 when flag bit 1 is clear)
    48      8 L    q[L]   moduli                    u64 LE each   (crt.jl:282-295 order)
    ..      8 L    psi[L] primitive 2N-th roots     u64 LE each   (pow2_cyc_rings.jl:27-47)
    ..      8 batch*components*L*N   residues, [batch][components][L][N], canonical in [0, q_i), u64 LE
    ..      4      CRC-32 (zlib) of everything before it

Each prime row [N] is what one StructArray field array of the reference holds (crt.jl:150-156), so a Julia reader fills
`fieldarrays(sa)[i]` with `reinterpret(PrimeField{Int64,q_i}, row)`."""
from __future__ import annotations

import struct
import zlib
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np

MAGIC = b"TFB2CT\0\0"
VERSION = 1
_HDR = struct.Struct("<8sIIIIIIQd")


class FormatError(ValueError):
    pass


@dataclass
class CtFile:
    N: int
    qs: list
    psis: list
    residues: np.ndarray            # uint64 [batch][components][L][N]
    dual: bool = False
    scale: Optional[float] = None


def dumps(residues: np.ndarray, qs: Sequence[int], psis: Sequence[int], dual: bool = False, scale: Optional[float] = None) -> bytes:
    a = np.ascontiguousarray(residues, dtype=np.uint64)
    if a.ndim != 4 or a.shape[2] != len(qs) or len(qs) != len(psis):
        raise FormatError("residues must be [batch][components][L][N] with L = len(qs) = len(psis)")
    B, comps, L, N = a.shape
    if N < 2 or N & (N - 1):
        raise FormatError("N must be a power of two")
    for i, q in enumerate(qs):
        if int(a[:, :, i].max(initial=0)) >= q:
            raise FormatError(f"residues under prime {i} are not canonical")
    flags = (1 if dual else 0) | (2 if scale is not None else 0)
    head = _HDR.pack(MAGIC, VERSION, flags, N, L, comps, 0, B, float(scale) if scale is not None else 0.0)
    body = head + np.asarray(qs, dtype="<u8").tobytes() + np.asarray(psis, dtype="<u8").tobytes() + a.astype("<u8", copy=False).tobytes()
    return body + struct.pack("<I", zlib.crc32(body) & 0xFFFFFFFF)


def loads(buf: bytes) -> CtFile:
    if len(buf) < _HDR.size + 4:
        raise FormatError("truncated file")
    magic, version, flags, N, L, comps, _res, B, scale = _HDR.unpack_from(buf, 0)
    if magic != MAGIC:
        raise FormatError("not a toyfhe_b200 ciphertext file")
    if version != VERSION:
        raise FormatError(f"unsupported version {version}")
    need = _HDR.size + 16 * L + 8 * B * comps * L * N + 4
    if len(buf) != need:
        raise FormatError(f"size mismatch: {len(buf)} bytes, header says {need}")
    (crc,) = struct.unpack_from("<I", buf, need - 4)
    if zlib.crc32(buf[:need - 4]) & 0xFFFFFFFF != crc:
        raise FormatError("checksum mismatch")
    off = _HDR.size
    qs = np.frombuffer(buf, dtype="<u8", count=L, offset=off).tolist()
    psis = np.frombuffer(buf, dtype="<u8", count=L, offset=off + 8 * L).tolist()
    res = np.frombuffer(buf, dtype="<u8", count=B * comps * L * N, offset=off + 16 * L).reshape(B, comps, L, N).astype(np.uint64)
    return CtFile(N, qs, psis, res, bool(flags & 1), scale if flags & 2 else None)


def save(path: str, residues: np.ndarray, qs, psis, dual: bool = False, scale: Optional[float] = None) -> None:
    with open(path, "wb") as f:
        f.write(dumps(residues, qs, psis, dual, scale))


def load(path: str) -> CtFile:
    with open(path, "rb") as f:
        return loads(f.read())
