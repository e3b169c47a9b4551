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


def port_flops(problem, st, delta_t=0.0, theta=0.0, h_elem_avg=0.0, U_norm=0.0):
    """Floating-point operations ONE fill of ``problem`` executes in the restatement, counted by running
    ``fill_port.c`` with an operation-counting number type (``flop_count.cpp``).
    Returns (flops per element, dict of totals)."""
    from goma_b200 import capi

    lib_path = os.path.join(HERE, "libflop_count.so")
    srcs = [os.path.join(HERE, f) for f in ("flop_count.cpp", "fill_port.c")]
    if not os.path.isfile(lib_path) or os.path.getmtime(lib_path) < max(os.path.getmtime(f) for f in srcs):
        subprocess.check_call(["make", "-s", "-C", HERE, "libflop_count.so"])
    lib = C.CDLL(lib_path)
    pst, keep = capi.make_problem_struct(problem)
    n = pst.num_unknowns
    ija = np.ascontiguousarray(capi.pattern_msr(problem), np.int32)
    a = np.zeros(int(ija[n]) + 1)
    r = np.zeros(n)
    dp = C.POINTER(C.c_double)
    z = np.zeros(n)
    arr = lambda k: np.ascontiguousarray(st.get(k, z), np.float64)
    x, xo, xd = arr("x"), arr("x_old"), arr("xdot")
    cnt = (C.c_longlong * 4)()
    lib.goma_port_fill_counted.argtypes = [C.c_void_p, C.POINTER(C.c_int), dp, dp, dp, C.c_double, C.c_double, C.c_double,
                                           C.c_double, C.c_double, C.c_int, C.c_int, dp, dp, C.POINTER(C.c_longlong)]
    rc = lib.goma_port_fill_counted(C.byref(pst), ija.ctypes.data_as(C.POINTER(C.c_int)), x.ctypes.data_as(dp),
                                    xo.ctypes.data_as(dp), xd.ctypes.data_as(dp), delta_t, theta, 0.0, h_elem_avg, U_norm,
                                    1, 1, a.ctypes.data_as(dp), r.ctypes.data_as(dp), cnt)
    assert rc == 0
    tot = {"add": int(cnt[0]), "mul": int(cnt[1]), "div": int(cnt[2]), "fn": int(cnt[3])}
    return sum(tot.values()) / problem.mesh.num_elems, tot
