"""The drop-in boundary from the reference's own host language: a plain C program (examples/c_host) that sees only
include/goma_gpu_fill.h and the shared library -- no Python, no torch -- assembles a fixture and must reproduce the
reference's matrix_fill_full output."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from goma_b200 import capi
from tests.cases import case_state

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "examples", "c_host", "fill_from_c.c")


def _build(tmp_path):
    exe = str(tmp_path / "fill_from_c")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), SRC,
                           "-L", os.path.join(ROOT, "goma_b200"), "-lgoma_gpu_fill",
                           "-Wl,-rpath," + os.path.join(ROOT, "goma_b200"), "-o", exe])
    return exe


def test_c_host_compiles_and_links_against_the_abi(built, tmp_path):
    """The header is valid C99 and every entry point the C host uses resolves in libgoma_gpu_fill.so."""
    exe = _build(tmp_path)
    p = subprocess.run([exe], capture_output=True, text=True)
    assert p.returncode == 2 and "usage" in p.stderr


@pytest.mark.gpu
def test_c_host_reproduces_reference_fixture(built, tmp_path):
    exe = _build(tmp_path)
    name = "c2_hex27_ns"
    p, kw, st = case_state(name)
    g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    pst, keep = capi.make_problem_struct(p)
    with open(tmp_path / "problem.bin", "wb") as f:
        f.write(bytes(pst))
        f.write(np.ascontiguousarray(keep["conn"], np.int32).tobytes())
        for d in range(p.dim):
            f.write(np.ascontiguousarray(keep["coords"][d], np.float64).tobytes())
        f.write(np.ascontiguousarray(keep["first"], np.int32)[: p.mesh.num_nodes].tobytes())
        f.write(np.ascontiguousarray(keep["kind"], np.uint8)[: p.mesh.num_nodes].tobytes())
        f.write(np.ascontiguousarray(keep["dbc_flag"], np.uint8)[: pst.num_unknowns].tobytes())
        f.write(np.ascontiguousarray(keep["dbc_value"], np.float64)[: pst.num_unknowns].tobytes())
        f.write(np.ascontiguousarray(st["x"], np.float64).tobytes())
    out = subprocess.run([exe, str(tmp_path / "problem.bin"), str(tmp_path / "out.bin")], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    b = open(tmp_path / "out.bin", "rb").read()
    err, f0, f1, f2 = np.frombuffer(b, np.int32, 4)
    nnz_plus = int(np.frombuffer(b, np.int64, 1, 16)[0])
    a = np.frombuffer(b, np.float64, nnz_plus + 1, 24)
    r = np.frombuffer(b, np.float64, len(g["resid"]), 24 + 8 * (nnz_plus + 1))
    assert err == 0 and not (f0 or f1 or f2) and nnz_plus == len(g["ija"])
    assert np.abs(a - g["a"]).max() / np.abs(g["a"]).max() < 1e-12
    assert np.abs(r - g["resid"]).max() / np.abs(g["resid"]).max() < 1e-12
