"""Build the C part of the oracle (TEST INFRASTRUCTURE): ``gcc -O2 -ffp-contract=off`` of
``oracle/nl_oracle.c`` into ``oracle/_build/libnl_oracle.so``.  Called by ``__graft_entry__.build()``
and lazily by the tests.  ``-ffp-contract=off`` keeps the float64 arithmetic un-fused so that the
distance test is the literal restated expression."""
from __future__ import annotations

import ctypes
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libnl_oracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "nl_oracle.c")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        os.makedirs(OUT_DIR, exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", LIB, src, "-lm"])
    return LIB


def load():
    lib = ctypes.CDLL(build())
    i64p = ctypes.POINTER(ctypes.c_int64)
    lib.nl_pbc_rows.restype = ctypes.c_int64
    lib.nl_pbc_rows.argtypes = [ctypes.POINTER(ctypes.c_float), ctypes.c_int64, ctypes.POINTER(ctypes.c_float),
                                ctypes.c_double, i64p, ctypes.c_int64, ctypes.c_int64, i64p, i64p, i64p]
    return lib


def nl_pbc_rows(pos, cell, rc, centres=None):
    """numpy front-end: returns (i, j, S) for the given centres (all atoms if None)."""
    import numpy as np
    lib = load()
    pos = np.ascontiguousarray(pos, dtype=np.float32)
    cell = np.ascontiguousarray(cell, dtype=np.float32).reshape(9)
    n = pos.shape[0]
    centres = np.arange(n, dtype=np.int64) if centres is None else np.ascontiguousarray(centres, dtype=np.int64)
    fp = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))
    ip = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_int64))
    z = np.zeros(1, dtype=np.int64)
    cnt = lib.nl_pbc_rows(fp(pos), n, fp(cell), float(rc), ip(centres), len(centres), 0, ip(z), ip(z), ip(z))
    oi = np.zeros(max(cnt, 1), dtype=np.int64); oj = np.zeros(max(cnt, 1), dtype=np.int64)
    os_ = np.zeros((max(cnt, 1), 3), dtype=np.int64)
    lib.nl_pbc_rows(fp(pos), n, fp(cell), float(rc), ip(centres), len(centres), cnt, ip(oi), ip(oj), ip(os_))
    return oi[:cnt], oj[:cnt], os_[:cnt]


if __name__ == "__main__":
    print(build(force=True))
