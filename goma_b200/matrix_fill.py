"""Host-side mirror of the reference's assembly entry point.

``MatrixFill.matrix_fill_full`` keeps the argument meaning and return convention of
``int matrix_fill_full(ams, x, resid_vector, x_old, x_older, xdot, xdot_old, x_update,
ptr_delta_t, ptr_theta, first_elem_side_BC_array, ptr_time_value, exo, dpi,
ptr_num_total_nodes, ptr_h_elem_avg, ptr_U_norm, estifm)`` (reference
``include/mm_fill.h:41-58``, ``src/mm_fill.c:158-312``): the state vectors go in, the MSR
values ``a`` (``ams->val``) and ``resid_vector`` come out, the return value is 0 or -1 with
the three domain-failure flags.  All compute happens in ``libgoma_gpu_fill.so``.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi


class MatrixFill:
    def __init__(self, problem, device: int = 0, ija=None, num_owned_nodes=None, layout: str = "msr",
                 host_stream_chunks: int = 0):
        """``layout``: "msr" = the reference's ``ams->val`` (default); "csr" = CSR of the owned rows with the diagonal in
        place, assembled directly (``goma_gpu_problem.matrix_layout``).  ``host_stream_chunks`` = K > 1: host-buffer
        fills sweep the elements in K chunks and copy finished rows to the host under the assembly of the later chunks
        (``goma_gpu_problem.host_stream_chunks``)."""
        self.lib = capi.load_library()
        self.problem = problem
        self.device = int(device)
        self.layout = layout
        self._struct, self._keep = capi.make_problem_struct(problem, ija=ija, num_owned_nodes=num_owned_nodes, layout=layout,
                                                               host_stream_chunks=host_stream_chunks)
        self._ctx = C.c_void_p()
        capi.check(self.lib.goma_gpu_fill_init(C.byref(self._struct), device, C.byref(self._ctx)),
                   "goma_gpu_fill_init")
        nnz = C.c_longlong()
        capi.check(self.lib.goma_gpu_fill_get_msr(self._ctx, C.byref(nnz)), "goma_gpu_fill_get_msr")
        self.nnz_plus = int(nnz.value)
        capi.check(self.lib.goma_gpu_fill_value_count(self._ctx, C.byref(nnz)), "goma_gpu_fill_value_count")
        self.value_count = int(nnz.value)  # doubles in `a`: nnz_plus + 1 (MSR) or the CSR nnz
        self.num_unknowns = int(self._struct.num_unknowns)
        self.flags = np.zeros(3, np.int32)

    # -- MSR graph (ams->bindx) as derived by the library
    def export_msr(self) -> np.ndarray:
        ija = np.zeros(self.nnz_plus + 1, np.int32)
        capi.check(self.lib.goma_gpu_fill_export_msr(self._ctx, C.byref(self._struct),
                                                     ija.ctypes.data_as(capi._ip)), "goma_gpu_fill_export_msr")
        return ija[: self.nnz_plus]

    def set_option(self, name: str, value: int):
        capi.check(self.lib.goma_gpu_fill_set_option(self._ctx, name.encode(), int(value)), "set_option")

    def matrix_fill_full(self, x, x_old=None, x_older=None, xdot=None, xdot_old=None, delta_t=0.0, theta=0.0,
                         time_value=0.0, h_elem_avg=0.0, U_norm=0.0, assemble_residual=True,
                         assemble_jacobian=True, a=None, resid_vector=None):
        """Returns (err, a, resid_vector).  ``a`` has the MSR layout of ``ams->val`` (length nnz_plus+1)."""
        n = self.num_unknowns

        def prep(v):
            if v is None:
                return None
            v = np.ascontiguousarray(v, np.float64)
            if v.shape != (n,):
                raise ValueError(f"state vector has shape {v.shape}, expected ({n},)")
            return v

        x, x_old, x_older, xdot, xdot_old = (prep(v) for v in (x, x_old, x_older, xdot, xdot_old))
        if a is None and assemble_jacobian:
            a = np.empty(self.value_count, np.float64)
        if resid_vector is None and assemble_residual:
            resid_vector = np.empty(n, np.float64)
        p = lambda v: capi._ptr(v, capi._dp)
        err = capi.check(self.lib.goma_gpu_fill(self._ctx, p(x), p(x_old), p(x_older), p(xdot), p(xdot_old),
                                                float(delta_t), float(theta), float(time_value),
                                                float(h_elem_avg), float(U_norm), int(assemble_residual),
                                                int(assemble_jacobian), p(a), p(resid_vector),
                                                self.flags.ctypes.data_as(capi._ip)), "goma_gpu_fill")
        return err, a, resid_vector

    def fill_raw(self, ptrs, delta_t=0.0, theta=0.0, time_value=0.0, h_elem_avg=0.0, U_norm=0.0,
                 assemble_residual=True, assemble_jacobian=True):
        """Host-buffer call on raw addresses (pinned buffers owned by the caller): ptrs =
        (x, x_old, x_older, xdot, xdot_old, a, resid) as ints (0 = NULL)."""
        cast = lambda v: C.cast(C.c_void_p(v or None), capi._dp)
        x, xo, xoo, xd, xdo, a, r = ptrs
        return capi.check(self.lib.goma_gpu_fill(self._ctx, cast(x), cast(xo), cast(xoo), cast(xd), cast(xdo),
                                                 float(delta_t), float(theta), float(time_value),
                                                 float(h_elem_avg), float(U_norm), int(assemble_residual),
                                                 int(assemble_jacobian), cast(a), cast(r),
                                                 self.flags.ctypes.data_as(capi._ip)), "goma_gpu_fill")

    def device_buffers(self) -> capi.DeviceBuffers:
        b = capi.DeviceBuffers()
        capi.check(self.lib.goma_gpu_fill_device_buffers(self._ctx, C.byref(b)), "device_buffers")
        return b

    def fill_device(self, delta_t=0.0, theta=0.0, time_value=0.0, h_elem_avg=0.0, U_norm=0.0,
                    assemble_residual=True, assemble_jacobian=True) -> int:
        """Device-resident assembly: state already in the context's device buffers."""
        return capi.check(self.lib.goma_gpu_fill_device(self._ctx, float(delta_t), float(theta), float(time_value),
                                                        float(h_elem_avg), float(U_norm), int(assemble_residual),
                                                        int(assemble_jacobian),
                                                        self.flags.ctypes.data_as(capi._ip)), "goma_gpu_fill_device")

    def fill_device_async(self, delta_t=0.0, theta=0.0, time_value=0.0, h_elem_avg=0.0, U_norm=0.0,
                          assemble_residual=True, assemble_jacobian=True) -> int:
        """Enqueue the device-resident assembly and return; the value is the cudaEvent_t (as int) recorded behind
        it.  ``fill_wait`` collects the return code and the flags."""
        ev = C.c_void_p()
        capi.check(self.lib.goma_gpu_fill_device_async(self._ctx, float(delta_t), float(theta), float(time_value),
                                                       float(h_elem_avg), float(U_norm), int(assemble_residual),
                                                       int(assemble_jacobian), C.byref(ev)), "goma_gpu_fill_device_async")
        return int(ev.value or 0)

    def fill_wait(self) -> int:
        return capi.check(self.lib.goma_gpu_fill_wait(self._ctx, self.flags.ctypes.data_as(capi._ip)), "goma_gpu_fill_wait")

    def exchange_fence(self, which: int = 0):
        """Owner-side fence: the stream waits until every neighbour has pulled vector ``which`` of the last exchange."""
        capi.check(self.lib.goma_gpu_exchange_fence(self._ctx, int(which)), "goma_gpu_exchange_fence")

    def exchange_status(self) -> int:
        return capi.check(self.lib.goma_gpu_exchange_status(self._ctx), "goma_gpu_exchange_status")

    def global_h_U(self, elem_owned=None):
        """Local sums of global_h_elem_siz / global_velocity_norm from the device-resident ``x``:
        (sum_h, n_elems, sum_v2, n_velocity_unknowns); the caller all-reduces and divides
        (``mm_fill_aux.c:1194-1204``)."""
        out = np.zeros(4)
        eo = None if elem_owned is None else np.ascontiguousarray(elem_owned, np.uint8)
        capi.check(self.lib.goma_gpu_global_h_U(self._ctx, capi._ptr(eo, capi._bp), out.ctypes.data_as(capi._dp)),
                   "goma_gpu_global_h_U")
        return out

    # -- the passes that follow the fill in the Newton loop, on the device-resident system
    def row_sum_scale(self, want_scale=True):
        """row_sum_scaling_scale (``sl_matrix_util.c:441``) of the device-resident matrix and residual, in place.
        Returns (scale[owned unknowns] or None, number of zero rows)."""
        n = C.c_int()
        capi.check(self.lib.goma_gpu_scale_buffer(self._ctx, None, C.byref(n)), "goma_gpu_scale_buffer")
        scale = np.empty(n.value) if want_scale else None
        zr = C.c_int()
        capi.check(self.lib.goma_gpu_row_sum_scale(self._ctx, capi._ptr(scale, capi._dp), C.byref(zr)),
                   "goma_gpu_row_sum_scale")
        return scale, int(zr.value)

    def matvec(self, v):
        """w = A v with the device-resident matrix of the last fill (``goma_gpu_matvec``; the product of the Newton line
        search, ``mm_sol_nonlinear.c:442-449``).  ``v``: torch float64 CUDA tensor over all local unknowns; returns a
        tensor of the same length whose owned rows hold the product (the others are zero)."""
        import torch

        assert v.is_cuda and v.dtype == torch.float64 and v.numel() == self.num_unknowns and v.is_contiguous()
        w = torch.zeros_like(v)
        torch.cuda.synchronize(v.device)
        capi.check(self.lib.goma_gpu_matvec(self._ctx, C.c_void_p(v.data_ptr()), C.c_void_p(w.data_ptr())), "goma_gpu_matvec")
        return w

    def vector_norms(self, which: int = 0):
        """(Loo, L1, L2, index) of the owned part of resid (0), x (1) or xdot (2) -- local values; a distributed
        host combines them with MPI_MAXLOC / MPI_SUM before the square root (``mm_sol_nonlinear.c:3177-3375``)."""
        out = np.zeros(4)
        capi.check(self.lib.goma_gpu_vector_norms(self._ctx, int(which), out.ctypes.data_as(capi._dp)),
                   "goma_gpu_vector_norms")
        return float(out[0]), float(out[1]), float(np.sqrt(out[2])), int(out[3])

    def csr(self, refresh_values=True):
        """CSR view of the owned rows on the device (``goma_gpu_csr_structure`` / ``goma_gpu_csr_values``) as torch
        tensors aliasing the library's buffers: (rowptr int64 [n+1], colind int32 [nnz], values float64 [nnz])."""
        import torch

        h = capi.Csr()
        capi.check(self.lib.goma_gpu_csr_structure(self._ctx, C.byref(self._struct), C.byref(h)), "goma_gpu_csr_structure")
        if refresh_values:
            capi.check(self.lib.goma_gpu_csr_values(self._ctx), "goma_gpu_csr_values")
        dev = torch.device("cuda", self.device)
        return (device_view(h.d_rowptr, h.num_rows + 1, dev, "<i8"), device_view(h.d_colind, int(h.nnz), dev, "<i4"),
                device_view(h.d_values, int(h.nnz), dev, "<f8"))

    def csr_rows(self):
        """CSR layout only: (rowptr int64 [n+1], values float64 [nnz]) aliasing the library's buffers -- no column array
        (``goma_gpu_csr_rows``); ``node_graph`` describes the columns at node level."""
        import torch

        h = capi.Csr()
        capi.check(self.lib.goma_gpu_csr_rows(self._ctx, C.byref(h)), "goma_gpu_csr_rows")
        dev = torch.device("cuda", self.device)
        return device_view(h.d_rowptr, h.num_rows + 1, dev, "<i8"), device_view(h.d_values, int(h.nnz), dev, "<f8")

    def download_system(self):
        """D2H of the device-resident MSR values and residual (after row_sum_scale, say)."""
        import torch

        b = self.device_buffers()
        dev = torch.device("cuda", self.device)
        a = device_view(b.d_a, self.value_count, dev).cpu().numpy()
        r = device_view(b.d_resid, self.num_unknowns, dev).cpu().numpy()
        return a, r

    # -- exchange_dof over NVLink peer memory (dp_comm.setup_peer_exchange wires the ranks together)
    def exchange_export(self) -> bytes:
        h = capi.ExchangeHandles()
        capi.check(self.lib.goma_gpu_exchange_export(self._ctx, C.byref(h)), "goma_gpu_exchange_export")
        return bytes(h)

    def exchange_setup(self, handles, my_slot_at_neighbor, recv_ptr, recv_list, tail_begin):
        n = len(handles)
        arr = (capi.ExchangeHandles * max(n, 1))()
        for k, b in enumerate(handles):
            arr[k] = capi.ExchangeHandles.from_buffer_copy(b)
        slot = np.ascontiguousarray(my_slot_at_neighbor, np.int32)
        rp = np.ascontiguousarray(recv_ptr, np.int32)
        rl = np.ascontiguousarray(recv_list, np.int32)
        ip = lambda v: v.ctypes.data_as(capi._ip)
        capi.check(self.lib.goma_gpu_exchange_setup(self._ctx, n, arr, ip(slot), ip(rp), ip(rl), int(tail_begin)),
                   "goma_gpu_exchange_setup")

    def exchange_dof(self, which: int = 0):
        """Ghost refresh of x (0), xdot (1) or x_old (2): one kernel on the context's stream, no host sync."""
        capi.check(self.lib.goma_gpu_exchange_dof(self._ctx, int(which)), "goma_gpu_exchange_dof")

    def setup_stats(self):
        """Wall seconds of goma_gpu_fill_init: total, validation, uploads, device pattern, tables + records."""
        out = np.zeros(5)
        capi.check(self.lib.goma_gpu_fill_setup_stats(self._ctx, out.ctypes.data_as(capi._dp)), "setup_stats")
        return dict(zip(("total_s", "validate_s", "upload_s", "pattern_s", "records_s"), out.tolist()))

    def last_stats(self):
        ms, n = C.c_double(), C.c_int()
        capi.check(self.lib.goma_gpu_fill_last_stats(self._ctx, C.byref(ms), C.byref(n)), "last_stats")
        return float(ms.value), int(n.value)

    def close(self):
        if self._ctx:
            self.lib.goma_gpu_fill_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class _CudaArray:
    """Minimal __cuda_array_interface__ holder so torch can view a buffer owned by the C library."""

    def __init__(self, ptr, n, typestr="<f8"):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def device_view(ptr, n, device, typestr="<f8"):
    """torch tensor aliasing ``n`` elements at device address ``ptr`` (no copy, no ownership)."""
    import torch

    return torch.as_tensor(_CudaArray(ptr, n, typestr), device=device)


def msr_to_csr(ija, a, n):
    """MSR (diagonal first, reference ``mm_fill_util.c:2865-3031``) -> scipy CSR, for the host Newton loop."""
    import scipy.sparse as sp

    ija = np.asarray(ija)
    rows = np.repeat(np.arange(n), np.diff(ija[: n + 1]))
    cols = ija[n + 1: ija[n]]
    vals = a[n + 1: ija[n]]
    A = sp.csr_matrix((vals, (rows, cols)), shape=(n, n)) + sp.diags(a[:n])
    return A.tocsr()
