"""Reader/writer for the TRBDUMP1 container used to move flattened basins and state
snapshots between the reference harness, the C oracle, the host shim and the tests.

Layout: 8-byte magic ``TRBDUMP1`` then records
``name[48] | int32 dtype (0=f64, 1=i32) | int32 pad | int64 count | payload``.
"""
from __future__ import annotations

import numpy as np

MAGIC = b"TRBDUMP1"
_DT = {0: np.dtype("<f8"), 1: np.dtype("<i4")}


def read_trb(path: str, prefix: str | None = None) -> dict:
    out = {}
    with open(path, "rb") as f:
        buf = f.read()
    if buf[:8] != MAGIC:
        raise ValueError(f"{path}: not a TRBDUMP1 file")
    o = 8
    n = len(buf)
    while o < n:
        name = buf[o:o + 48].split(b"\0", 1)[0].decode()
        dt, _pad = np.frombuffer(buf, "<i4", 2, o + 48)
        cnt = int(np.frombuffer(buf, "<i8", 1, o + 56)[0])
        o += 64
        d = _DT[int(dt)]
        nb = cnt * d.itemsize
        if prefix is None or name.startswith(prefix):
            out[name] = np.frombuffer(buf, d, cnt, o).copy()
        o += nb
    return out


def write_trb(path: str, arrays: dict) -> None:
    with open(path, "wb") as f:
        f.write(MAGIC)
        for name, a in arrays.items():
            a = np.ascontiguousarray(a)
            if a.dtype.kind == "f":
                a = a.astype("<f8"); dt = 0
            else:
                a = a.astype("<i4"); dt = 1
            nm = name.encode()[:47]
            f.write(nm + b"\0" * (48 - len(nm)))
            f.write(np.array([dt, 0], "<i4").tobytes())
            f.write(np.array([a.size], "<i8").tobytes())
            f.write(a.tobytes())
