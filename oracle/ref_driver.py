"""Run the reference's own matrix_fill_full through ``oracle/_ref/goma_ref_fill``.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  ``goma_ref_fill`` is the
reference's unmodified C code (built by ``oracle/ref_build/build.sh`` from
``/root/reference/src``) behind a small driver; this module writes its work
directory (deck, .mat, mesh.bin, state.bin), runs it and parses the binary output.
"""
from __future__ import annotations

import os
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_EXE = os.path.join(HERE, "_ref", "goma_ref_fill")


def ref_available() -> bool:
    return os.path.isfile(REF_EXE) and os.access(REF_EXE, os.X_OK)


def write_mesh(mesh, path):
    ns_ids = sorted(mesh.node_sets)
    ns_ptr = [0]
    chunks = []
    for k in ns_ids:
        chunks.append(np.asarray(mesh.node_sets[k], np.int32) + 1)
        ns_ptr.append(ns_ptr[-1] + len(chunks[-1]))
    ns_nodes = np.concatenate(chunks).astype(np.int32) if chunks else np.zeros(0, np.int32)
    with open(path, "wb") as f:
        f.write(np.array([mesh.dim, mesh.num_nodes, mesh.num_elems, mesh.npe, len(ns_ids), len(ns_nodes)],
                         np.int32).tobytes())
        f.write(mesh.elem_type.encode().ljust(32, b"\0"))
        for d in range(mesh.dim):
            f.write(np.ascontiguousarray(mesh.coords[d], np.float64).tobytes())
        f.write((mesh.conn + 1).astype(np.int32).tobytes())
        f.write(np.array(ns_ids, np.int32).tobytes())
        f.write(np.array(ns_ptr, np.int32).tobytes())
        f.write(ns_nodes.tobytes())
        # element blocks (optional trailer): count, then the number of (consecutive) elements of each block
        eb = getattr(mesh, "elem_block", None)
        ss = getattr(mesh, "side_sets", None) or {}
        if eb is not None:
            eb = np.asarray(eb)
            if (np.diff(eb) < 0).any():
                raise ValueError("elements must be ordered block by block (EXODUS II stores them that way)")
            counts = np.bincount(eb)
            f.write(np.array([len(counts)], np.int32).tobytes())
            f.write(counts.astype(np.int32).tobytes())
        elif ss:
            f.write(np.array([0], np.int32).tobytes())
        # side sets (second optional trailer): id -> (elements, EXODUS side numbers 1..), both 0-based here / 1-based on file
        if ss:
            ids = sorted(ss)
            ptr = np.cumsum([0] + [len(ss[k][0]) for k in ids]).astype(np.int32)
            f.write(np.array([len(ids)], np.int32).tobytes())
            f.write(np.array(ids, np.int32).tobytes())
            f.write(ptr.tobytes())
            f.write((np.concatenate([np.asarray(ss[k][0]) for k in ids]) + 1).astype(np.int32).tobytes())
            f.write(np.concatenate([np.asarray(ss[k][1]) for k in ids]).astype(np.int32).tobytes())


def write_workdir(problem, workdir):
    os.makedirs(workdir, exist_ok=True)
    with open(os.path.join(workdir, "input"), "w") as f:
        f.write(problem.deck())
    for m in range(getattr(problem, "num_materials", 1)):
        with open(os.path.join(workdir, problem.mat_name(m) + ".mat"), "w") as f:
            f.write(problem.mat_file(m))
    write_mesh(problem.mesh, os.path.join(workdir, "mesh.bin"))


def _run(workdir, *args, timeout=3600):
    p = subprocess.run([REF_EXE, workdir, *args], capture_output=True, text=True, timeout=timeout)
    if p.returncode != 0:
        raise RuntimeError(f"goma_ref_fill failed ({p.returncode}):\n{p.stdout[-3000:]}\n{p.stderr[-3000:]}")
    return p.stdout


def run_map(problem, workdir=None):
    """Unknown map, MSR pattern, Dirichlet table and Inter_Mask as the reference builds them."""
    with tempfile.TemporaryDirectory() as tmp:
        wd = workdir or tmp
        write_workdir(problem, wd)
        _run(wd, "map")
        b = open(os.path.join(wd, "map.bin"), "rb").read()
    h = np.frombuffer(b, np.int32, 8)
    nu, N, nnzp, nn = (int(v) for v in h[:4])
    o = 32
    out = {"num_unknowns": nu, "N": N, "nnz_plus": nnzp, "num_nodes": nn, "num_elems": int(h[4]),
           "pspg": int(h[6]), "max_species": int(h[7])}
    out["first_unknown"] = np.frombuffer(b, np.int32, nn + 1, o).copy(); o += 4 * (nn + 1)
    out["ija"] = np.frombuffer(b, np.int32, nnzp + 1, o).copy(); o += 4 * (nnzp + 1)
    out["idv"] = np.frombuffer(b, np.int32, 3 * nu, o).reshape(nu, 3).copy(); o += 12 * nu
    out["x_dirichlet"] = np.frombuffer(b, np.float64, nu, o).copy(); o += 8 * nu
    out["dbc"] = np.frombuffer(b, np.int32, nu, o).copy(); o += 4 * nu
    out["inter_mask"] = np.frombuffer(b, np.int32, 100, o).reshape(10, 10).copy()
    return out


def run_fill(problem, states, delta_t=0.0, theta=0.0, time=0.0, assemble_jacobian=True,
             preset_dirichlet=False, h_elem_avg=-1.0, U_norm=-1.0, nrep=1, workdir=None, timeout=3600, post=False):
    """``states``: list of dicts with x and optional x_old, x_older, xdot, xdot_old.

    Returns a list of dicts: err, flags, a (MSR values, length nnz_plus+1), resid, x (after the
    optional Dirichlet preset), best_s / mean_s wall time of matrix_fill_full.
    """
    n = len(states[0]["x"])
    with tempfile.TemporaryDirectory() as tmp:
        wd = workdir or tmp
        write_workdir(problem, wd)
        with open(os.path.join(wd, "state.bin"), "wb") as f:
            f.write(np.array([n, len(states), int(assemble_jacobian), int(preset_dirichlet)], np.int32).tobytes())
            f.write(np.array([delta_t, theta, time, h_elem_avg, U_norm], np.float64).tobytes())
            z = np.zeros(n)
            for s in states:
                for key in ("x", "x_old", "x_older", "xdot", "xdot_old"):
                    f.write(np.ascontiguousarray(s.get(key, z), np.float64).tobytes())
        stdout = _run(wd, "fill", str(nrep), *(["post"] if post else []), timeout=timeout)
        b = open(os.path.join(wd, "fill_out.bin"), "rb").read()
        pb = open(os.path.join(wd, "post_out.bin"), "rb").read() if post else None
    nu, nnzp, ns, N = (int(v) for v in np.frombuffer(b, np.int32, 4))
    o = 16
    res = []
    for _ in range(ns):
        fl = np.frombuffer(b, np.int32, 4, o); o += 16
        tm = np.frombuffer(b, np.float64, 4, o); o += 32
        x = np.frombuffer(b, np.float64, nu, o).copy(); o += 8 * nu
        a = np.frombuffer(b, np.float64, nnzp + 1, o).copy(); o += 8 * (nnzp + 1)
        r = np.frombuffer(b, np.float64, nu, o).copy(); o += 8 * nu
        res.append({"err": int(fl[0]), "neg_elem_volume": int(fl[1]), "neg_lub_height": int(fl[2]),
                    "zero_detJ": int(fl[3]), "best_s": float(tm[0]), "mean_s": float(tm[1]),
                    "h_elem_avg": float(tm[2]), "U_norm": float(tm[3]), "x": x, "a": a, "resid": r, "N": N,
                    "stdout": stdout})
    if post:  # row_sum_scaling_scale + norms of the scaled residual, by the reference's own functions
        o = 0
        for r_ in res:
            nrm = np.frombuffer(pb, np.float64, 4, o); o += 32
            r_["post_norms"] = nrm.copy()
            r_["post_scale"] = np.frombuffer(pb, np.float64, nu, o).copy(); o += 8 * nu
            r_["post_a"] = np.frombuffer(pb, np.float64, nnzp + 1, o).copy(); o += 8 * (nnzp + 1)
            r_["post_resid"] = np.frombuffer(pb, np.float64, nu, o).copy(); o += 8 * nu
    return res
