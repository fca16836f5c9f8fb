"""TEST INFRASTRUCTURE: reader for the record container written by oracle/ref_harness.cpp."""
import struct
import numpy as np


def read_dump(path):
    """Return {name: float64 ndarray} for every record in the file."""
    out = {}
    with open(path, "rb") as fh:
        data = fh.read()
    off = 0
    n = len(data)
    while off < n:
        (nl,) = struct.unpack_from("<I", data, off); off += 4
        name = data[off:off + nl].decode(); off += nl
        (nd,) = struct.unpack_from("<I", data, off); off += 4
        dims = struct.unpack_from("<%dq" % nd, data, off); off += 8 * nd
        cnt = int(np.prod(dims))
        out[name] = np.frombuffer(data, dtype="<f8", count=cnt, offset=off).reshape(dims).copy()
        off += 8 * cnt
    return out


def write_dump(path, records):
    with open(path, "wb") as fh:
        for name, arr in records.items():
            arr = np.ascontiguousarray(arr, dtype="<f8")
            nb = name.encode()
            fh.write(struct.pack("<I", len(nb))); fh.write(nb)
            fh.write(struct.pack("<I", arr.ndim)); fh.write(struct.pack("<%dq" % arr.ndim, *arr.shape))
            fh.write(arr.tobytes())
