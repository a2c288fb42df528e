"""Reader for the tagged-array dumps written by oracle/ref_harness_*.cpp (oracle/dumpfmt.h)."""
import struct
import numpy as np

_DT = {b"f": np.float32, b"d": np.float64, b"i": np.int32, b"B": np.uint8, b"H": np.uint16}


def read_dump(path):
    out = {}
    with open(path, "rb") as f:
        data = f.read()
    p = 0
    while p < len(data):
        (nl,) = struct.unpack_from("<I", data, p); p += 4
        name = data[p:p + nl].decode(); p += nl
        dt = _DT[data[p:p + 1]]; p += 1
        (nd,) = struct.unpack_from("<I", data, p); p += 4
        dims = struct.unpack_from("<%dQ" % nd, data, p); p += 8 * nd
        n = int(np.prod(dims)) if nd else 1
        arr = np.frombuffer(data, dtype=dt, count=n, offset=p).reshape(dims).copy()
        p += n * np.dtype(dt).itemsize
        out[name] = arr
    return out
