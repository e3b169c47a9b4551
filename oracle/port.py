"""ctypes wrapper of the CPU restatement ``oracle/fill_port.c`` -- TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libfill_port.so")
_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", HERE, "libfill_port.so"])


def load():
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB) or os.path.getmtime(LIB) < os.path.getmtime(os.path.join(HERE, "fill_port.c")):
            build()
        _lib = C.CDLL(LIB)
    return _lib


def port_fill(problem, ija, st, delta_t=0.0, theta=0.0, time=0.0, h_elem_avg=0.0, U_norm=0.0,
              assemble_residual=True, assemble_jacobian=True, num_owned_nodes=None):
    """Returns (rc, a, resid) in the MSR layout of ``ija``."""
    from goma_b200 import capi  # only the struct definition of the C ABI header

    lib = load()
    pst, keep = capi.make_problem_struct(problem, num_owned_nodes=num_owned_nodes)
    n = pst.num_unknowns
    ija = np.ascontiguousarray(ija, np.int32)
    nnz_plus = int(ija[n])
    a = np.zeros(nnz_plus + 1)
    r = np.zeros(n)
    dp = C.POINTER(C.c_double)
    z = np.zeros(n)
    arr = lambda k: np.ascontiguousarray(st.get(k, z), np.float64)
    x, xo, xd = arr("x"), arr("x_old"), arr("xdot")
    lib.goma_port_fill.argtypes = [C.c_void_p, C.POINTER(C.c_int), dp, dp, dp, C.c_double, C.c_double, C.c_double,
                                   C.c_double, C.c_double, C.c_int, C.c_int, dp, dp]
    rc = lib.goma_port_fill(C.byref(pst), ija.ctypes.data_as(C.POINTER(C.c_int)), x.ctypes.data_as(dp),
                            xo.ctypes.data_as(dp), xd.ctypes.data_as(dp), delta_t, theta, time, h_elem_avg, U_norm,
                            int(assemble_residual), int(assemble_jacobian), a.ctypes.data_as(dp), r.ctypes.data_as(dp))
    return rc, a, r
