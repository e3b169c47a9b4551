"""ctypes binding of the C ABI in ``include/goma_gpu_fill.h`` (``libgoma_gpu_fill.so``).

This is the same binding a Goma maintainer would write in C (INTEGRATION.md shows the
shim for ``matrix_fill_full``); Python is only the test/bench harness.  There is no CPU
fallback: if the library is missing or no CUDA device is present the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GOMA_GPU_LIB") or os.path.join(_HERE, "libgoma_gpu_fill.so")

NSLOT = 12
MAX_KINDS = 4
SLOTS = {"U": 0, "V": 1, "W": 2, "T": 3, "Y0": 4, "Y1": 5, "Y2": 6, "Y3": 7, "DX": 8, "DY": 9, "DZ": 10, "P": 11}
PRESSURE_P1, PRESSURE_EQ = 1, 2

EXPORTED = [
    "goma_gpu_fill_init", "goma_gpu_fill_destroy", "goma_gpu_fill_get_msr", "goma_gpu_fill_export_msr",
    "goma_gpu_fill", "goma_gpu_fill_device_buffers", "goma_gpu_fill_device", "goma_gpu_global_h_U",
    "goma_gpu_pack_dofs", "goma_gpu_unpack_dofs", "goma_gpu_fill_last_stats", "goma_gpu_fill_set_option",
    "goma_gpu_last_error", "goma_gpu_pattern_msr", "goma_gpu_exchange_export", "goma_gpu_exchange_setup",
    "goma_gpu_exchange_dof", "goma_gpu_matvec", "goma_gpu_row_sum_scale", "goma_gpu_scale_buffer", "goma_gpu_vector_norms", "goma_gpu_csr_structure",
    "goma_gpu_csr_values", "goma_gpu_exchange_status", "goma_gpu_fill_device_async", "goma_gpu_fill_wait", "goma_gpu_fill_setup_stats",
    "goma_gpu_fill_value_count", "goma_gpu_csr_rows", "goma_gpu_node_graph", "goma_gpu_exchange_fence",
]

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_bp = C.POINTER(C.c_ubyte)



class Material(C.Structure):
    """``struct goma_gpu_material`` (include/goma_gpu_fill.h)."""

    _fields_ = [
        ("rho", C.c_double), ("mu", C.c_double), ("conductivity", C.c_double), ("heat_capacity", C.c_double),
        ("volume_expansion", C.c_double), ("reference_temperature", C.c_double),
        ("diffusivity", C.c_double * 4), ("momentum_source", C.c_double * 3), ("momentum_source_model", C.c_int),
        ("heat_source", C.c_double), ("lame_mu", C.c_double), ("lame_lambda", C.c_double),
    ]


class GomaGpuProblem(C.Structure):
    _fields_ = [
        ("dim", C.c_int), ("elem_type", C.c_int), ("num_nodes", C.c_int), ("num_owned_nodes", C.c_int),
        ("num_elems", C.c_int), ("elem_connect", _ip), ("coord", _dp * 3),
        ("num_unknowns", C.c_int), ("first_unknown", _ip), ("num_kinds", C.c_int), ("node_kind", _bp),
        ("kind_slot", (C.c_int * NSLOT) * MAX_KINDS), ("kind_num_unknowns", C.c_int * MAX_KINDS),
        ("ija", _ip),
        ("pressure_interp", C.c_int), ("energy", C.c_int), ("num_species", C.c_int), ("ale", C.c_int),
        ("transient", C.c_int), ("pspg", C.c_int), ("ps_scaling", C.c_double),
        ("etm_momentum", C.c_double * 6), ("etm_continuity", C.c_double * 2), ("etm_energy", C.c_double * 5),
        ("etm_species", C.c_double * 5), ("etm_mesh", C.c_double * 5),
        ("rho", C.c_double), ("mu", C.c_double), ("conductivity", C.c_double), ("heat_capacity", C.c_double),
        ("volume_expansion", C.c_double), ("reference_temperature", C.c_double),
        ("diffusivity", C.c_double * 4), ("momentum_source", C.c_double * 3), ("momentum_source_model", C.c_int),
        ("heat_source", C.c_double), ("lame_mu", C.c_double), ("lame_lambda", C.c_double),
        ("dbc_flag", _bp), ("dbc_value", _dp),
        ("num_elem_blocks", C.c_int), ("num_materials", C.c_int), ("elem_material", _ip),
        ("materials", C.POINTER(Material)), ("matrix_layout", C.c_int),
        ("host_stream_chunks", C.c_int),
    ]


IPC_HANDLE_BYTES = 64
MAX_NEIGHBORS = 32


class ExchangeHandles(C.Structure):
    """``struct goma_gpu_exchange_handles``: CUDA IPC handles of x, xdot, x_old and of the flag block."""
    _fields_ = [("vec", (C.c_ubyte * IPC_HANDLE_BYTES) * 3), ("flags", C.c_ubyte * IPC_HANDLE_BYTES),
                ("device", C.c_int)]


class Csr(C.Structure):
    """``struct goma_gpu_csr``."""
    _fields_ = [("num_rows", C.c_int), ("nnz", C.c_longlong), ("d_rowptr", C.c_void_p), ("d_colind", C.c_void_p),
                ("d_values", C.c_void_p)]


class DeviceBuffers(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in
                ("d_x", "d_x_old", "d_x_older", "d_xdot", "d_xdot_old", "d_a", "d_resid", "stream")]


_lib = None


def load_library():
    """Load ``libgoma_gpu_fill.so`` (built in-tree by ``__graft_entry__.build()``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback for the GPU fill)")
    lib = C.CDLL(LIB_PATH)
    lib.goma_gpu_last_error.restype = C.c_char_p
    lib.goma_gpu_fill_init.argtypes = [C.POINTER(GomaGpuProblem), C.c_int, C.POINTER(C.c_void_p)]
    lib.goma_gpu_fill_destroy.argtypes = [C.c_void_p]
    lib.goma_gpu_fill_destroy.restype = None
    lib.goma_gpu_fill_get_msr.argtypes = [C.c_void_p, C.POINTER(C.c_longlong)]
    lib.goma_gpu_fill_export_msr.argtypes = [C.c_void_p, C.POINTER(GomaGpuProblem), _ip]
    lib.goma_gpu_pattern_msr.argtypes = [C.POINTER(GomaGpuProblem), C.POINTER(C.c_longlong), _ip]
    lib.goma_gpu_fill.argtypes = [C.c_void_p, _dp, _dp, _dp, _dp, _dp, C.c_double, C.c_double, C.c_double,
                                  C.c_double, C.c_double, C.c_int, C.c_int, _dp, _dp, _ip]
    lib.goma_gpu_fill_device_buffers.argtypes = [C.c_void_p, C.POINTER(DeviceBuffers)]
    lib.goma_gpu_fill_device.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double,
                                         C.c_int, C.c_int, _ip]
    lib.goma_gpu_fill_device_async.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double,
                                               C.c_int, C.c_int, C.POINTER(C.c_void_p)]
    lib.goma_gpu_fill_wait.argtypes = [C.c_void_p, _ip]
    lib.goma_gpu_global_h_U.argtypes = [C.c_void_p, _bp, _dp]
    lib.goma_gpu_exchange_export.argtypes = [C.c_void_p, C.POINTER(ExchangeHandles)]
    lib.goma_gpu_exchange_setup.argtypes = [C.c_void_p, C.c_int, C.POINTER(ExchangeHandles), _ip, _ip, _ip, C.c_int]
    lib.goma_gpu_exchange_dof.argtypes = [C.c_void_p, C.c_int]
    lib.goma_gpu_exchange_status.argtypes = [C.c_void_p]
    lib.goma_gpu_exchange_fence.argtypes = [C.c_void_p, C.c_int]
    lib.goma_gpu_row_sum_scale.argtypes = [C.c_void_p, _dp, _ip]
    lib.goma_gpu_matvec.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.goma_gpu_scale_buffer.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), _ip]
    lib.goma_gpu_vector_norms.argtypes = [C.c_void_p, C.c_int, _dp]
    lib.goma_gpu_csr_structure.argtypes = [C.c_void_p, C.POINTER(GomaGpuProblem), C.POINTER(Csr)]
    lib.goma_gpu_csr_values.argtypes = [C.c_void_p]
    lib.goma_gpu_pack_dofs.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    lib.goma_gpu_unpack_dofs.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    lib.goma_gpu_fill_last_stats.argtypes = [C.c_void_p, _dp, _ip]
    lib.goma_gpu_fill_setup_stats.argtypes = [C.c_void_p, _dp]
    lib.goma_gpu_fill_value_count.argtypes = [C.c_void_p, C.POINTER(C.c_longlong)]
    lib.goma_gpu_csr_rows.argtypes = [C.c_void_p, C.POINTER(Csr)]
    lib.goma_gpu_node_graph.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
    lib.goma_gpu_fill_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
    _lib = lib
    return lib


class GomaGpuError(RuntimeError):
    pass


def check(rc, what):
    """0 ok, -1 = matrix_fill_full's own 'domain failure' return; anything else raises."""
    if rc not in (0, -1):
        msg = load_library().goma_gpu_last_error().decode(errors="replace")
        raise GomaGpuError(f"{what} failed ({rc}): {msg}")
    return rc


def _ptr(a, typ):
    return a.ctypes.data_as(typ) if a is not None else typ()


LAYOUT_MSR, LAYOUT_CSR = 0, 1


def make_problem_struct(problem, ija=None, num_owned_nodes=None, layout="msr", host_stream_chunks=0):
    """Fill ``struct goma_gpu_problem`` from a :class:`goma_b200.problem.Problem`.

    Returns (struct, keepalive) -- keepalive holds the numpy arrays the struct points into.
    """
    m = problem.mesh
    first, node_kind, kinds = problem.unknown_map()
    is_dbc, value, hard = problem.dirichlet_table()
    keep = {
        "conn": np.ascontiguousarray(m.conn, np.int32),
        "coords": [np.ascontiguousarray(m.coords[d], np.float64) for d in range(m.dim)],
        "first": np.ascontiguousarray(first[:-1], np.int32),
        "kind": np.ascontiguousarray(node_kind, np.uint8),
        "dbc_flag": np.ascontiguousarray(np.where(is_dbc == 1, np.where(hard == 1, 2, 1), 0), np.uint8),
        "dbc_value": np.ascontiguousarray(value, np.float64),
        "ija": None if ija is None else np.ascontiguousarray(ija, np.int32),
    }
    p = GomaGpuProblem()
    p.dim = m.dim
    p.elem_type = m.npe
    p.num_nodes = m.num_nodes
    p.num_owned_nodes = m.num_nodes if num_owned_nodes is None else int(num_owned_nodes)
    p.num_elems = m.num_elems
    p.elem_connect = _ptr(keep["conn"], _ip)
    for d in range(m.dim):
        p.coord[d] = _ptr(keep["coords"][d], _dp)
    p.num_unknowns = int(first[-1])
    p.first_unknown = _ptr(keep["first"], _ip)
    p.num_kinds = len(kinds)
    p.node_kind = _ptr(keep["kind"], _bp)
    for k in range(MAX_KINDS):
        for s in range(NSLOT):
            p.kind_slot[k][s] = -1
        p.kind_num_unknowns[k] = 0
    for k, slots in enumerate(kinds):
        p.kind_num_unknowns[k] = len(slots)
        for off, name in enumerate(slots):
            if p.kind_slot[k][SLOTS[name]] < 0:  # first dof of a multi-dof variable
                p.kind_slot[k][SLOTS[name]] = off
    p.ija = _ptr(keep["ija"], _ip)
    p.pressure_interp = PRESSURE_P1 if problem.interp == "Q2P1" else PRESSURE_EQ
    p.energy = int(problem.energy)
    p.num_species = problem.n_species
    p.ale = int(problem.ale)
    p.transient = int(problem.transient)
    p.pspg = {None: 0, "global": 1, "local": 2}[problem.pspg]
    p.ps_scaling = problem.ps_scaling
    for name, src in (("etm_momentum", problem.etm_momentum), ("etm_continuity", problem.etm_continuity),
                      ("etm_energy", problem.etm_energy), ("etm_species", problem.etm_species),
                      ("etm_mesh", problem.etm_mesh)):
        arr = getattr(p, name)
        for i, v in enumerate(src):
            arr[i] = v
    p.rho, p.mu, p.conductivity, p.heat_capacity = problem.rho, problem.mu, problem.k, problem.Cp
    p.volume_expansion, p.reference_temperature = problem.beta, problem.Tref
    for w in range(4):
        p.diffusivity[w] = problem.diffusivity[w]
    for d in range(3):
        p.momentum_source[d] = problem.gravity[d]
    p.momentum_source_model = {"CONSTANT": 0, "BOUSS": 1, "BOUSSINESQ": 2}[problem.ns_source]
    p.heat_source = problem.heat_source
    p.lame_mu, p.lame_lambda = problem.lame_mu, problem.lame_lambda
    p.dbc_flag = _ptr(keep["dbc_flag"], _bp)
    p.dbc_value = _ptr(keep["dbc_value"], _dp)
    if getattr(problem, "extra_bc_cards", None):
        raise GomaGpuError("integrated boundary conditions (BC cards on side sets) are not assembled by the GPU fill (SURVEY.md §8f-3)")
    p.num_elem_blocks = int(getattr(m, "num_elem_blocks", 1))
    p.num_materials = int(getattr(problem, "num_materials", 1))
    if p.num_materials > 1:  # material of each element = its element block (Matilda[ebn] == ebn in the decks written here)
        keep["elem_material"] = np.ascontiguousarray(m.elem_block, np.int32)
        mats = (Material * p.num_materials)()
        for k in range(p.num_materials):
            d = problem.material(k)
            mats[k].rho, mats[k].mu, mats[k].conductivity, mats[k].heat_capacity = d["rho"], d["mu"], d["k"], d["Cp"]
            mats[k].volume_expansion, mats[k].reference_temperature = d["beta"], d["Tref"]
            for w in range(4):
                mats[k].diffusivity[w] = d["diffusivity"][w]
            for a in range(3):
                mats[k].momentum_source[a] = d["gravity"][a]
            mats[k].momentum_source_model = {"CONSTANT": 0, "BOUSS": 1, "BOUSSINESQ": 2}[d["ns_source"]]
            mats[k].heat_source = d["heat_source"]
            mats[k].lame_mu, mats[k].lame_lambda = d["lame_mu"], d["lame_lambda"]
        keep["materials"] = mats
        p.elem_material = _ptr(keep["elem_material"], _ip)
        p.materials = C.cast(mats, C.POINTER(Material))
    p.matrix_layout = {"msr": LAYOUT_MSR, "csr": LAYOUT_CSR}[layout]
    p.host_stream_chunks = int(host_stream_chunks)
    return p, keep


def pattern_msr(problem):
    """MSR graph (``ija``) derived on the host by the library -- no GPU needed."""
    lib = load_library()
    st, keep = make_problem_struct(problem)
    nnz = C.c_longlong()
    check(lib.goma_gpu_pattern_msr(C.byref(st), C.byref(nnz), None), "goma_gpu_pattern_msr")
    ija = np.zeros(int(nnz.value) + 1, np.int32)
    check(lib.goma_gpu_pattern_msr(C.byref(st), C.byref(nnz), ija.ctypes.data_as(_ip)), "goma_gpu_pattern_msr")
    return ija[: int(nnz.value)]
