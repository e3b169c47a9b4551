"""CPU suite (-m "not gpu"): the oracle against the reference's fixtures, host logic, C-ABI exports."""
import ctypes
import os
import re

import numpy as np
import pytest

from goma_b200 import capi
from goma_b200.mesh import ELEM_TABLE, box_mesh
from goma_b200.problem import Problem
from oracle import port, ref_driver
from tests.cases import GOLDEN_CASES, case_state, build_case

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOL = 1e-12  # north_star: residual and Jacobian entries agree to 1e-12 relative in fp64


def golden(name):
    return np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))


def rel_err(got, ref):
    return np.abs(got - ref).max() / max(np.abs(ref).max(), 1e-300)


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_port_oracle_matches_reference_fixture(built, name):
    """oracle/fill_port.c reproduces what the reference's matrix_fill_full returned."""
    p, kw, st = case_state(name)
    g = golden(name)
    np.testing.assert_array_equal(g["state_x"], st["x"])  # the fixture was made from this very state
    h, U = (p.global_h_elem_siz(), p.global_velocity_norm(st["x"])) if p.pspg else (0.0, 0.0)
    if p.pspg:  # host restatement of global_h_elem_siz / global_velocity_norm vs the reference's values
        assert abs(h - float(g["h_elem_avg"])) < 1e-14 and abs(U - float(g["U_norm"])) < 1e-14
    rc, a, r = port.port_fill(p, g["ija"], st, delta_t=kw.get("delta_t", 0.0), theta=kw.get("theta", 0.0),
                              h_elem_avg=h, U_norm=U)
    assert rc == 0
    assert rel_err(a, g["a"]) < TOL
    assert rel_err(r, g["resid"]) < TOL


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_unknown_map_and_sparsity_bit_exact(built, name):
    """Unknown ordering, MSR graph and Dirichlet table equal the reference's, bit for bit."""
    p, kw, st = case_state(name)
    g = golden(name)
    first, node_kind, kinds = p.unknown_map()
    np.testing.assert_array_equal(first, g["first_unknown"])
    ija = capi.pattern_msr(p)  # host-only entry of the product library
    np.testing.assert_array_equal(ija, g["ija"])
    is_dbc, value, hard = p.dirichlet_table()
    np.testing.assert_array_equal(is_dbc == 1, g["dbc"] >= 0)
    # find_and_set_Dirichlet's preset of x (hard-set cards, deck order), restated in preset_dirichlet
    xm = p.preset_dirichlet(np.full(len(st["x"]), -7.77e77))
    np.testing.assert_array_equal(xm, g["x_dirichlet"])
    # Inter_Mask rows restated in Problem.inter_mask
    ids = [0, 1, 2, 3, 4, 9]
    for r in ids:
        for c in ids:
            if g["inter_mask"][r, r] and g["inter_mask"][c, c]:
                assert bool(g["inter_mask"][r, c]) == Problem.inter_mask(r, c), (r, c)


def test_c_abi_exports_every_declared_symbol(built):
    hdr = open(os.path.join(ROOT, "include", "goma_gpu_fill.h")).read()
    declared = set(re.findall(r"\b(goma_gpu_\w+)\s*\(", hdr))
    declared -= {"goma_gpu_ghost_tail"}  # mentioned in prose only
    lib = ctypes.CDLL(capi.LIB_PATH)
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, missing
    assert set(capi.EXPORTED) <= declared


def test_struct_layout_matches_header(built):
    """ctypes mirror and the C struct agree on size (catches drift between capi.py and the header)."""
    src = '#include "goma_gpu_fill.h"\n#include <stdio.h>\nint main(){printf("%zu %zu", sizeof(struct goma_gpu_problem), sizeof(struct goma_gpu_device_buffers));}'
    import subprocess
    import tempfile

    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(src)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "t.c"), "-o", os.path.join(d, "t")])
        out = subprocess.check_output([os.path.join(d, "t")]).decode().split()
    assert int(out[0]) == ctypes.sizeof(capi.GomaGpuProblem)
    assert int(out[1]) == ctypes.sizeof(capi.DeviceBuffers)


def test_gpu_path_fails_loudly_without_device(built):
    """No CPU fallback: without a CUDA device init must raise, not compute."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from goma_b200.matrix_fill import MatrixFill

    p, _ = build_case("c1_quad9_ns")
    with pytest.raises(capi.GomaGpuError, match="no CUDA device"):
        MatrixFill(p)


def test_unsupported_physics_is_refused(built):
    p, _ = build_case("c1_quad9_ns")
    p.n_species = 7
    st, keep = capi.make_problem_struct(Problem(p.mesh))
    st.num_species = 7
    nnz = ctypes.c_longlong()
    rc = capi.load_library().goma_gpu_pattern_msr(ctypes.byref(st), ctypes.byref(nnz), None)
    assert rc == -2 and b"num_species" in capi.load_library().goma_gpu_last_error()


@pytest.mark.parametrize("et,n", [("QUAD4", (3, 2)), ("QUAD9", (3, 2)), ("HEX8", (2, 3, 2)), ("HEX27", (2, 1, 2))])
def test_mesh_generator(et, n):
    m = box_mesh(et, n, perturb=0.1, seed=7)
    dim, offs, order, _ = ELEM_TABLE[et]
    assert m.num_elems == int(np.prod(n)) and m.npe == len(offs)
    assert m.num_nodes == int(np.prod([order * k + 1 for k in n]))
    assert len(np.unique(m.conn)) == m.num_nodes
    # every element has positive volume at its centroid: corner ordering is right-handed
    c = m.coords[:, m.conn[:, :4 if dim == 2 else 8]]
    if dim == 2:
        a = c[:, :, 1] - c[:, :, 0]
        b = c[:, :, 3] - c[:, :, 0]
        assert np.all(a[0] * b[1] - a[1] * b[0] > 0)
    else:
        a, b, d = (c[:, :, k] - c[:, :, 0] for k in (1, 3, 4))
        assert np.all(np.einsum("ie,ie->e", np.cross(a.T, b.T).T, d) > 0)


def test_empty_node_set_and_single_element(built):
    """Ragged/edge inputs: one element, a Dirichlet card on a node set with no matching unknown."""
    from goma_b200.problem import Dirichlet

    m = box_mesh("QUAD9", (1, 1))
    p = Problem(m, bcs=[Dirichlet("P", 1, 0.0)])  # x-min nodes carry no pressure dof
    is_dbc, _, _ = p.dirichlet_table()
    assert is_dbc.sum() == 0
    ija = capi.pattern_msr(p)
    n = int(p.unknown_map()[0][-1])
    assert n == 9 * 2 + 3 and ija[0] == n + 1 and len(ija) == n + 1 + n * (n - 1)


@pytest.mark.skipif(not ref_driver.ref_available(), reason="oracle/_ref not built (needs /root/reference)")
def test_reference_live_matches_fixture():
    """Where the reference binary is present, re-run it and confirm the committed fixture."""
    name = "c1_quad9_ns"
    p, kw, st = case_state(name)
    g = golden(name)
    res = ref_driver.run_fill(p, [st])[0]
    np.testing.assert_array_equal(res["a"], g["a"])
    np.testing.assert_array_equal(res["resid"], g["resid"])


def test_c1_newton_converges_with_the_oracle_assembly(built):
    """BASELINE.json configs[0] (C1) through the CPU oracle: Newton on the 40x10 quad9 channel converges
    quadratically to a developed (Poiseuille-like) profile -- the reference case the GPU path is compared to."""
    from tests.newton_util import channel_problem, newton

    p = channel_problem()
    ija = capi.pattern_msr(p)
    first, node_kind, kinds = p.unknown_map()
    x0 = p.preset_dirichlet(np.zeros(int(first[-1])))
    x, it, norms = newton(lambda v: port.port_fill(p, ija, {"x": v}), ija, x0)
    assert 2 <= it <= 8 and norms[-1] < 1e-10 and norms[-1] < norms[0] * 1e-9
    mid = np.nonzero((np.abs(p.mesh.coords[0] - 2.0) < 1e-12) & (np.abs(p.mesh.coords[1] - 0.5) < 1e-12))[0][0]
    assert 1.4 < x[first[mid]] < 1.6  # fully developed plane Poiseuille flow: 1.5 x mean velocity


def test_post_fill_oracle_matches_reference_fixture():
    """oracle/post_fill.py (row_sum_scale_MSR + Loo/L1/L2 norms) == the reference's own functions, bit for bit,
    on the system the reference assembled for c1_quad9_ns (tests/golden/make_golden_post.py)."""
    from oracle import post_fill

    g, q = golden("c1_quad9_ns"), golden("post_c1_quad9_ns")
    n = len(g["resid"])
    a, b, scale = post_fill.row_sum_scale_msr(n, g["a"], g["ija"], g["resid"])
    np.testing.assert_array_equal(scale, q["scale"])
    np.testing.assert_array_equal(b, q["resid"])
    np.testing.assert_array_equal(a[: len(q["a"]) - 1], q["a"][:-1])
    loo, l1, l2, k = post_fill.norms(b, n)
    assert (loo, l1, l2, float(k)) == tuple(q["norms"])


def test_header_is_valid_c_and_cxx(tmp_path):
    """include/goma_gpu_fill.h compiles as C99 and as C++17 (extern "C" guards) on its own."""
    import subprocess

    hdr = os.path.join(ROOT, "include", "goma_gpu_fill.h")
    c = tmp_path / "t.c"
    c.write_text('#include "%s"\nint main(void) { struct goma_gpu_problem p; (void)p; return 0; }\n' % hdr)
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-fsyntax-only", str(c)])
    cxx = tmp_path / "t.cpp"
    cxx.write_text('#include "%s"\nint main() { goma_gpu_csr a{}; const char *(*f)(void) = goma_gpu_last_error; (void)a; (void)f; return 0; }\n' % hdr)
    subprocess.check_call(["g++", "-std=c++17", "-Wall", "-Werror", "-fsyntax-only", str(cxx)])
