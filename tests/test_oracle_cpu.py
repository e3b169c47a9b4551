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


def _init_rc(st):
    """goma_gpu_pattern_msr runs the same validate() as goma_gpu_fill_init and needs no device."""
    nnz = ctypes.c_longlong()
    lib = capi.load_library()
    rc = lib.goma_gpu_pattern_msr(ctypes.byref(st), ctypes.byref(nnz), None)
    return rc, lib.goma_gpu_last_error().decode()


def test_inconsistent_unknown_map_is_refused(built):
    """validate(): everything the pattern builder and the record kernel index with is checked first -- a node kind
    out of range, a first_unknown that is not the running sum, a variable missing on a node kind (a multi-material
    layout) and a bad Dirichlet flag are errors (-2), not out-of-bounds reads."""
    p, _ = build_case("c3_hex27_boussinesq")
    st, keep = capi.make_problem_struct(p)
    assert _init_rc(st)[0] == 0
    keep["kind"][5] = 3
    rc, msg = _init_rc(st)
    assert rc == -2 and "node_kind" in msg
    keep["kind"][5] = 0
    keep["first"][7] += 1
    rc, msg = _init_rc(st)
    assert rc == -2 and "first_unknown" in msg
    keep["first"][7] -= 1
    st.kind_slot[0][capi.SLOTS["T"]] = -1  # temperature not defined on the plain nodes
    rc, msg = _init_rc(st)
    assert rc == -2 and "temperature" in msg
    st.kind_slot[0][capi.SLOTS["T"]] = 3
    keep["dbc_flag"][0] = 9
    rc, msg = _init_rc(st)
    assert rc == -2 and "dbc_flag" in msg
    keep["dbc_flag"][0] = 0
    keep["conn"][0, 0] = 10 ** 6
    rc, msg = _init_rc(st)
    assert rc == -2 and "connectivity" in msg


def test_multi_block_and_multi_material_meshes_are_refused(built):
    """The ABI carries one element type and one material (find_elemblock_index / Matilda[ebn],
    src/mm_fill.c:224-235, are not restated): a host that states two blocks or two materials is refused."""
    p, _ = build_case("c1_quad9_ns")
    st, keep = capi.make_problem_struct(p)
    st.num_elem_blocks = 2
    rc, msg = _init_rc(st)
    assert rc == -2 and "element block" in msg
    st.num_elem_blocks = 1
    st.num_materials = 2
    rc, msg = _init_rc(st)
    assert rc == -2 and "material" in msg


@pytest.mark.parametrize("et,nz", [("QUAD4", 0), ("QUAD9", 0), ("HEX8", 2), ("HEX27", 1)])
def test_unstructured_mesh_generator(et, nz):
    """star_mesh: valences 3 and 5, positive orientation, unique nodes, boundary node set closed."""
    from collections import Counter

    from goma_b200.mesh import star_mesh

    m = star_mesh(et, refine=1, nz=nz, perturb=0.3, seed=1)
    assert len(np.unique(m.conn)) == m.num_nodes
    n2d = m.num_elems // max(nz, 1)
    val = Counter(m.conn[:n2d, :4].ravel().tolist())
    assert {3, 5} <= set(val.values())
    c = m.conn[:, :4]
    x, y = m.coords[0][c], m.coords[1][c]
    area = 0.5 * sum(x[:, k] * y[:, (k + 1) % 4] - x[:, (k + 1) % 4] * y[:, k] for k in range(4))
    assert (area > 0).all()


@pytest.mark.skipif(not ref_driver.ref_available(), reason="oracle/_ref not built (needs /root/reference)")
def test_unstructured_mesh_pattern_matches_reference_live(built):
    """find_MSR_problem_graph of the reference on an irregular-valence hex8 mesh == the library's host pattern."""
    from goma_b200.mesh import star_mesh
    from goma_b200.problem import Dirichlet

    m = star_mesh("HEX8", refine=2, nz=2, perturb=0.2, seed=4)
    p = Problem(m, interp="Q1Q1", pspg="local", energy=True, ns_source="BOUSSINESQ", bcs=[Dirichlet("U", 1, 0.0)])
    np.testing.assert_array_equal(ref_driver.run_map(p)["ija"][:-1], capi.pattern_msr(p))


@pytest.mark.parametrize("parts,nranks", [((2, 2, 2), 8), (None, 5)])
def test_irregular_partitions_reproduce_global_rows(built, parts, nranks):
    """decompose() on partitions with many neighbours per rank (2x2x2 bricks: 7 each; a ragged random partition):
    owner = lowest rank, one ghost layer, owned rows only -- the rows of all ranks are the global system, and the
    peer-pull lists fill every external tail (dp_map_comm_vec.c:332-461, dp_comm.c:48-102)."""
    from goma_b200.dp_comm import (brick_partition, decompose, peer_exchange_payload, peer_recv_lists,
                                   scattered_partition)
    from goma_b200.matrix_fill import msr_to_csr
    from goma_b200.problem import Dirichlet
    from tests.cases import make_state

    m = box_mesh("HEX8", (4, 4, 4), perturb=0.1, seed=5)
    p = Problem(m, interp="Q1Q1", pspg="global", energy=True, rho=1.1, mu=0.2, ns_source="BOUSSINESQ",
                gravity=(0.1, -0.3, 0.2), bcs=[Dirichlet("U", 1, 1.0), Dirichlet("T", 2, 0.0, relax=1.0)])
    st = make_state(p, seed=3)
    first_g = p.unknown_map()[0]
    ija_g = capi.pattern_msr(p)
    h, U = p.global_h_elem_siz(), p.global_velocity_norm(st["x"])
    rc, a_g, r_g = port.port_fill(p, ija_g, st, h_elem_avg=h, U_norm=U)
    ng = len(r_g)
    A_g = msr_to_csr(ija_g, a_g, ng)
    part = brick_partition(m, parts) if parts else scattered_partition(m, nranks, seed=2)
    subs = decompose(p, part, nranks)
    if parts:
        assert all(len(s.neighbors) == 7 for s in subs)
    everyone = [peer_exchange_payload(s, b"h") for s in subs]
    l2gs = []
    for s in subs:
        first_l = s.problem.unknown_map()[0]
        l2g = np.empty(int(first_l[-1]), np.int64)
        for k, g in enumerate(s.node_global):
            l2g[first_l[k]:first_l[k + 1]] = np.arange(first_g[g], first_g[g + 1])
        l2gs.append(l2g)
    seen = np.zeros(ng, int)
    for s, l2g in zip(subs, l2gs):
        nown, nl = s.num_owned_dofs, len(l2g)
        # ghost refresh through the pull lists
        _, slots, recv_ptr, rl = peer_recv_lists(s, everyone)
        x = st["x"][l2g].copy()
        x[nown:] = np.nan
        for k, q in enumerate(s.neighbors):
            x[nown + recv_ptr[k]: nown + recv_ptr[k + 1]] = st["x"][l2gs[q]][rl[recv_ptr[k]:recv_ptr[k + 1]]]
        np.testing.assert_array_equal(x, st["x"][l2g])
        ija = capi.pattern_msr(s.problem)
        rc, a, r = port.port_fill(s.problem, ija, {"x": x}, h_elem_avg=h, U_norm=U, num_owned_nodes=s.num_owned_nodes)
        assert rc == 0
        lrows = np.repeat(np.arange(nl), np.diff(ija[:nl + 1]))
        lcols, vals = ija[nl + 1:], a[nl + 1:len(ija)]
        own = lrows < nown
        ref = np.asarray(A_g[l2g[lrows[own]], l2g[lcols[own]]]).ravel()
        assert (not own.any() or np.abs(vals[own] - ref).max() < 1e-13) and not vals[~own].any()
        np.testing.assert_allclose(a[:nown], a_g[l2g[:nown]], rtol=0, atol=1e-13)
        np.testing.assert_allclose(r[:nown], r_g[l2g[:nown]], rtol=0, atol=1e-13)
        seen[l2g[:nown]] += 1
    assert (seen == 1).all()


def test_msr_matvec_oracle_equals_dense_product():
    """oracle/post_fill.py::msr_matvec (the published Aztec DMSR product, restated) against a dense matrix rebuilt from the
    reference's own graph and values of a fixture: the same numbers up to the summation order."""
    from oracle import post_fill

    g = dict(np.load(os.path.join(os.path.dirname(__file__), "golden", "c1_quad9_ns.npz")))
    ija, a = np.asarray(g["ija"], np.int64), g["a"]
    n = len(g["resid"])
    dense = np.zeros((n, n))
    dense[np.arange(n), np.arange(n)] = a[:n]
    rows = np.repeat(np.arange(n), np.diff(ija[: n + 1]))
    dense[rows, ija[n + 1: ija[n]]] = a[n + 1: ija[n]]
    v = np.random.default_rng(3).standard_normal(n)
    w = post_fill.msr_matvec(n, a, ija, v)
    np.testing.assert_allclose(w, dense @ v, rtol=0, atol=1e-12 * np.abs(dense).sum(1).max() * np.abs(v).max())


def test_f3_integrated_bcs_change_only_the_rows_of_their_side_set_nodes():
    """SURVEY.md §8f-3 groundwork.  The fixture is the reference's own assembly of the C4 strip with KINEMATIC + CAPILLARY
    on the free surface (side set 4) and VELO_NORMAL on the wall below (side set 3), parsed from a real deck by the
    reference and applied by its apply_integrated_bc (mm_fill.c:2945-3033) with the rotation of the equations.
    Everything OUTSIDE the rows of the nodes on those two side sets equals the volumetric assembly of the restatement
    (what the GPU path reproduces); inside them the integrated conditions replace or add to it -- those rows are what
    §8f-3 still has to build.  The C-ABI marshalling refuses such a problem instead of assembling it without them."""
    import dataclasses

    p, kw, st = case_state("f3_quad9_free_surface")
    g = golden("f3_quad9_free_surface")
    vol = dataclasses.replace(p, extra_bc_cards=[])
    ija = capi.pattern_msr(vol)
    np.testing.assert_array_equal(ija, g["ija"])  # the integrated conditions add no couplings to the graph
    rc, a, r = port.port_fill(vol, ija, st)
    assert rc == 0
    n = len(r)
    first, kind, kinds = p.unknown_map()
    bc_nodes = np.unique(np.concatenate([p.mesh.node_sets[3], p.mesh.node_sets[4]]))
    bc_rows = np.zeros(n, bool)
    for nd in bc_nodes:
        bc_rows[first[nd]:first[nd + 1]] = True
    rows = np.repeat(np.arange(n), np.diff(np.asarray(ija[: n + 1], np.int64)))
    scale = np.abs(g["a"][:n]).copy()
    np.maximum.at(scale, rows, np.abs(g["a"][n + 1: int(ija[n])]))
    dd = np.abs(a[:n] - g["a"][:n])
    do = np.abs(a[n + 1: int(ija[n])] - g["a"][n + 1: int(ija[n])])
    free = ~bc_rows
    assert (dd[free] <= 1e-12 * scale[free]).all() and (do[free[rows]] <= 1e-12 * scale[rows][free[rows]]).all()
    assert (np.abs(r - g["resid"])[free] <= 1e-12 * np.abs(g["resid"]).max()).all()
    changed = np.zeros(n, bool)
    changed[np.nonzero(dd > 1e-9 * np.maximum(scale, 1e-300))[0]] = True
    changed[rows[do > 1e-9 * np.maximum(scale[rows], 1e-300)]] = True
    assert changed.any() and not changed[free].any()
    assert np.abs(g["a"]).max() > 1e11  # BIG_PENALTY rows of the strongly integrated conditions (rf_bc_const.h)
    with pytest.raises(capi.GomaGpuError, match="integrated boundary conditions"):
        capi.make_problem_struct(p)
