"""Reader/writer for the "GTBA" named-array container used by fixtures and golden vectors.

Layout: 8-byte magic ``GTBA0001``, u64 n_arrays, then per array: char name[48], u32 elem_size,
u32 kind (0 unsigned, 1 signed, 2 float), u64 count, raw little-endian data padded to 8 bytes.
"""
from __future__ import annotations

import gzip
import os
import struct
from typing import Dict

import numpy as np

_DT = {(1, 0): np.uint8, (2, 0): np.uint16, (4, 0): np.uint32, (8, 0): np.uint64,
       (1, 1): np.int8, (2, 1): np.int16, (4, 1): np.int32, (8, 1): np.int64,
       (4, 2): np.float32, (8, 2): np.float64}


def load(path: str) -> Dict[str, np.ndarray]:
    """Reads PATH, or PATH.gz when only the gzip-compressed file exists (the larger committed fixtures)."""
    if not os.path.exists(path) and os.path.exists(path + ".gz"):
        path += ".gz"
    if path.endswith(".gz"):
        with gzip.open(path, "rb") as f:
            buf = f.read()
    else:
        with open(path, "rb") as f:
            buf = f.read()
    if buf[:8] != b"GTBA0001":
        raise ValueError(f"{path}: not a GTBA file")
    (n,) = struct.unpack_from("<Q", buf, 8)
    off = 16
    out: Dict[str, np.ndarray] = {}
    for _ in range(n):
        name = buf[off:off + 48].split(b"\0", 1)[0].decode()
        esize, kind, count = struct.unpack_from("<IIQ", buf, off + 48)
        off += 64
        nbytes = esize * count
        out[name] = np.frombuffer(buf, dtype=_DT[(esize, kind)], count=count, offset=off).copy()
        off += nbytes + (8 - nbytes % 8) % 8
    return out


def save(path: str, arrays: Dict[str, np.ndarray]) -> None:
    with open(path, "wb") as f:
        f.write(b"GTBA0001")
        f.write(struct.pack("<Q", len(arrays)))
        for name, a in arrays.items():
            a = np.ascontiguousarray(a)
            kind = 2 if a.dtype.kind == "f" else (1 if a.dtype.kind == "i" else 0)
            f.write(name.encode()[:47].ljust(48, b"\0"))
            f.write(struct.pack("<IIQ", a.dtype.itemsize, kind, a.size))
            b = a.tobytes()
            f.write(b)
            f.write(b"\0" * ((8 - len(b) % 8) % 8))
