"""TEST INFRASTRUCTURE ONLY -- reader of the state dumps written by the dump-wrapper problem
generators (oracle/pgens/dump_common.hpp) from the reference's own entity.xc."""
from __future__ import annotations

import struct

import numpy as np

_DT = {0: np.float32, 1: np.int32, 2: np.int16, 3: np.float64, 4: np.uint32}


def read(path: str) -> dict:
    out = {}
    with open(path, "rb") as f:
        buf = f.read()
    pos = 0
    while pos < len(buf):
        (nl,) = struct.unpack_from("<I", buf, pos)
        pos += 4
        name = buf[pos:pos + nl].decode()
        pos += nl
        dt, nd = struct.unpack_from("<II", buf, pos)
        pos += 8
        shape = struct.unpack_from(f"<{nd}Q", buf, pos)
        pos += 8 * nd
        dtype = np.dtype(_DT[dt])
        n = int(np.prod(shape)) if nd else 1
        out[name] = np.frombuffer(buf, dtype, n, pos).reshape(shape).copy()
        pos += n * dtype.itemsize
    return out
