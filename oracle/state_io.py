"""TEST INFRASTRUCTURE — reader for the state dumps written by oracle/probe_hook.hpp.

Record = char name[32]; char type ('d'|'i'); int64 rows, cols; data column-major.
"""
import struct
import numpy as np


def read_state(path):
    out = {}
    with open(path, "rb") as f:
        buf = f.read()
    pos = 0
    while pos < len(buf):
        name = buf[pos:pos + 32].split(b"\0", 1)[0].decode()
        typ = chr(buf[pos + 32])
        rows, cols = struct.unpack_from("<qq", buf, pos + 33)
        pos += 49
        dt = np.float64 if typ == "d" else np.int32
        n = rows * cols
        arr = np.frombuffer(buf, dtype=dt, count=n, offset=pos).copy()
        pos += n * (8 if typ == "d" else 4)
        out[name] = arr.reshape((cols, rows)).T.copy() if cols > 1 else arr  # col-major -> (rows, cols)
    return out
