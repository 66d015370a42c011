"""TEST INFRASTRUCTURE — ctypes wrapper of oracle/lsap.c (CPU restatement of scipy.optimize.linear_sum_assignment)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def _load():
    global _lib
    if _lib is None:
        so = os.path.join(_HERE, "_build", "liblsap_oracle.so")
        if not os.path.exists(so):
            subprocess.run(["make", "-C", _HERE, "-s"], check=True)
        _lib = ctypes.CDLL(so)
    return _lib


def linear_sum_assignment(cost, maximize=False):
    cost = np.ascontiguousarray(cost, dtype=np.float64)
    nr, nc = cost.shape
    k = min(nr, nc)
    rows = np.zeros(k, dtype=np.int64)
    cols = np.zeros(k, dtype=np.int64)
    rc = _load().lsap_oracle(cost.ctypes.data_as(ctypes.c_void_p), nr, nc, 1 if maximize else 0,
                             rows.ctypes.data_as(ctypes.c_void_p), cols.ctypes.data_as(ctypes.c_void_p))
    if rc != 0:
        raise ValueError("cost matrix is infeasible" if rc == -1 else "matrix contains invalid numeric entries")
    return rows, cols
