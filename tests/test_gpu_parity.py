"""GPU suite (-m gpu): the CUDA path, called through the C ABI, against the oracle.

Tolerance (BASELINE.json north_star): residual and Jacobian entries agree to 1e-12 relative in
fp64 (summation order differs); sparsity pattern and unknown ordering bit-exact.
"""
import os

import numpy as np
import pytest

from goma_b200 import capi
from goma_b200.matrix_fill import MatrixFill, msr_to_csr
from goma_b200.mesh import box_mesh
from goma_b200.problem import Dirichlet, Problem
from oracle import port, ref_driver
from tests.cases import GOLDEN_CASES, case_state, make_state

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOL = 1e-12


def golden(name):
    return np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))


def rel_err(got, ref):
    return np.abs(got - ref).max() / max(np.abs(ref).max(), 1e-300)


def assert_close(got, ref, what, ija=None):
    """Vectors: max |delta| <= TOL * max |ref|.  MSR matrices (``ija`` given): per ROW, SURVEY.md §7 --
    |delta_ij| <= TOL * ||row_i||_inf, so that a row whose entries are 1e-4 of the global maximum is held to
    1e-12 of ITS scale; rows the reference leaves zero (ghost rows) must be exactly zero."""
    if ija is None:
        e = rel_err(got, ref)
        assert e < TOL, f"{what}: max rel err {e:.3e}"
        return
    ija = np.asarray(ija, np.int64)
    n = int(ija[0]) - 1
    cnt = np.diff(ija[: n + 1])
    nnz_plus = int(ija[n])
    rows = np.repeat(np.arange(n), cnt)
    rowmax = np.abs(ref[:n]).copy()
    np.maximum.at(rowmax, rows, np.abs(ref[n + 1: nnz_plus]))
    dd = np.abs(got[:n] - ref[:n])
    do = np.abs(got[n + 1: nnz_plus] - ref[n + 1: nnz_plus])
    bad_d = dd > TOL * rowmax
    bad_o = do > TOL * rowmax[rows]
    if bad_d.any() or bad_o.any():
        worst = max((dd / np.maximum(rowmax, 1e-300)).max(), (do / np.maximum(rowmax[rows], 1e-300)).max() if len(do) else 0.0)
        raise AssertionError(f"{what}: {int(bad_d.sum() + bad_o.sum())} entries beyond {TOL:g} of their row scale "
                             f"(worst {worst:.3e})")


@pytest.mark.parametrize("scatter", [0, 1, 2], ids=["atomic", "coloured", "first_touch"])
@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_gpu_matches_reference_fixture(built, name, scatter):
    """CUDA assembly == the reference's own matrix_fill_full output (committed fixture)."""
    p, kw, st = case_state(name)
    g = golden(name)
    mf = MatrixFill(p, ija=g["ija"])  # the library also checks its derived graph against the host's
    mf.set_option("scatter", scatter)
    assert mf.nnz_plus == len(g["ija"])
    np.testing.assert_array_equal(mf.export_msr(), g["ija"])
    h, U = (p.global_h_elem_siz(), p.global_velocity_norm(st["x"])) if p.pspg else (0.0, 0.0)
    err, a, r = mf.matrix_fill_full(st["x"], st.get("x_old"), st.get("x_older"), st.get("xdot"), st.get("xdot_old"),
                                    delta_t=kw.get("delta_t", 0.0), theta=kw.get("theta", 0.0),
                                    time_value=kw.get("time", 0.0), h_elem_avg=h, U_norm=U)
    assert err == 0 and not mf.flags.any()
    assert_close(a, g["a"], "Jacobian", ija=g["ija"])
    assert_close(r, g["resid"], "residual")
    # residual-only call (modified Newton, mm_sol_nonlinear.c:1226-1231) leaves the same residual
    err, _, r2 = mf.matrix_fill_full(st["x"], st.get("x_old"), st.get("x_older"), st.get("xdot"), st.get("xdot_old"),
                                     delta_t=kw.get("delta_t", 0.0), theta=kw.get("theta", 0.0),
                                     h_elem_avg=h, U_norm=U, assemble_jacobian=False)
    assert_close(r2, g["resid"], "residual-only")
    mf.close()


@pytest.mark.parametrize("et,n,energy", [("HEX27", (5, 4, 3), True), ("HEX27", (4, 4, 4), False),
                                         ("QUAD9", (17, 9), True), ("QUAD9", (40, 10), False)])
def test_gpu_matches_port_oracle_seeded(built, et, n, energy):
    """Larger seeded cases than the fixtures, against the CPU restatement."""
    m = box_mesh(et, n, perturb=0.1, seed=11)
    dim = m.dim
    bcs = [Dirichlet("U", 1, 1.0), Dirichlet("V", 3, 0.0, relax=1.0), Dirichlet("U", 4, 0.0)]
    if energy:
        bcs += [Dirichlet("T", 1, 1.0), Dirichlet("T", 2, 0.0, relax=1.0)]
    p = Problem(m, energy=energy, rho=1.2, mu=0.03, k=0.05, Cp=1.3, beta=0.7, Tref=0.1,
                gravity=(0.1, -1.0, 0.2)[:3], ns_source="BOUSSINESQ" if energy else "CONSTANT",
                heat_source=0.3, bcs=bcs)
    st = make_state(p, seed=5)
    ija = capi.pattern_msr(p)
    rc, a_ref, r_ref = port.port_fill(p, ija, st)
    assert rc == 0
    mf = MatrixFill(p)
    for scatter in (0, 1, 2, 2):
        mf.set_option("scatter", scatter)
        err, a, r = mf.matrix_fill_full(st["x"])
        assert err == 0
        assert_close(a, a_ref, f"Jacobian scatter={scatter}", ija=ija)
        assert_close(r, r_ref, f"residual scatter={scatter}")
    mf.close()


def test_gpu_ghost_rows_are_not_written(built):
    """Rows of external (ghost) nodes stay zero: load_lec's owned_ledof test (mm_fill.c:5374)."""
    m = box_mesh("QUAD9", (6, 5), perturb=0.1, seed=3)
    p = Problem(m, bcs=[Dirichlet("U", 1, 1.0)])
    st = make_state(p, seed=9)
    owned = m.num_nodes - 40
    ija = capi.pattern_msr(p)
    rc, a_ref, r_ref = port.port_fill(p, ija, st, num_owned_nodes=owned)
    mf = MatrixFill(p, num_owned_nodes=owned)
    err, a, r = mf.matrix_fill_full(st["x"])
    assert_close(a, a_ref, "Jacobian", ija=ija)
    assert_close(r, r_ref, "residual")
    first = p.unknown_map()[0]
    assert not r[first[owned]:].any() and not a[first[owned]:len(r)].any()
    mf.close()


def test_gpu_jacobian_is_derivative_of_residual(built):
    """The reference's own self-check (numerical_jacobian, mm_numjac.c:716): J d == dR/dx . d.
    Size-independent property, run at a size the CPU oracle would take minutes for."""
    m = box_mesh("HEX27", (12, 12, 12), perturb=0.08, seed=2)
    p = Problem(m, energy=True, rho=1.0, mu=0.02, k=0.03, Cp=1.1, beta=0.5, Tref=0.0, gravity=(0, 0, -1.0),
                ns_source="BOUSSINESQ")
    st = make_state(p, seed=4)
    mf = MatrixFill(p)
    n = mf.num_unknowns
    err, a, r0 = mf.matrix_fill_full(st["x"])
    A = msr_to_csr(mf.export_msr(), a, n)
    rng = np.random.default_rng(0)
    d = rng.normal(size=n)
    eps = 1e-6
    _, _, rp = mf.matrix_fill_full(st["x"] + eps * d, assemble_jacobian=False)
    rp = rp.copy()
    _, _, rm = mf.matrix_fill_full(st["x"] - eps * d, assemble_jacobian=False)
    fd = (rp - rm) / (2 * eps)
    jd = A @ d
    assert np.abs(fd - jd).max() / np.abs(jd).max() < 1e-7
    mf.close()


def test_gpu_stokes_residual_is_affine(built):
    """With advection off the residual is affine in x: R(x) - R(0) == J x (exactly, to round-off)."""
    m = box_mesh("HEX27", (8, 8, 8), perturb=0.05, seed=6)
    p = Problem(m, rho=1.0, mu=0.5, etm_momentum=(0.0, 0.0, 1.0, 1.0, 1.0, 0.0), gravity=(0.1, 0.2, -0.3))
    st = make_state(p, seed=8)
    mf = MatrixFill(p)
    n = mf.num_unknowns
    _, a, r = mf.matrix_fill_full(st["x"])
    A = msr_to_csr(mf.export_msr(), a, n)
    _, _, r0 = mf.matrix_fill_full(np.zeros(n), assemble_jacobian=False)
    lhs = r - r0
    rhs = A @ st["x"]
    assert np.abs(lhs - rhs).max() / np.abs(rhs).max() < 1e-11
    mf.close()


@pytest.mark.parametrize("pspg,transient", [("local", False), ("global", True)])
def test_gpu_c5_hex8_matches_port_oracle(built, pspg, transient):
    """Config C5 physics (hex8 Q1/Q1 PSPG + energy + 2 species) on a seeded mesh larger than the fixtures."""
    m = box_mesh("HEX8", (7, 6, 5), perturb=0.12, seed=31)
    bcs = [Dirichlet("U", 1, 1.0), Dirichlet("V", 1, 0.0), Dirichlet("W", 5, 0.0), Dirichlet("T", 1, 1.0),
           Dirichlet("T", 2, 0.0, relax=1.0), Dirichlet("Y", 3, 0.7, species=1), Dirichlet("P", 7, 0.0)]
    p = Problem(m, interp="Q1Q1", energy=True, n_species=2, rho=0.9, mu=0.02, k=0.1, Cp=1.2, beta=0.6, Tref=0.1,
                diffusivity=(0.03, 0.2, 1.0, 1.0), gravity=(0.0, 0.3, -1.0), ns_source="BOUSSINESQ", heat_source=0.2,
                pspg=pspg, ps_scaling=0.1, transient=transient, etm_momentum=(1.0, 1.0, 1.0, 1.0, 1.0, 0.0),
                etm_energy=(1.0,) * 5, etm_species=(1.0,) * 5, bcs=bcs)
    st = make_state(p, seed=17, transient=transient, delta_t=0.02, theta=0.5)
    kw = dict(delta_t=0.02, theta=0.5) if transient else {}
    h, U = p.global_h_elem_siz(), p.global_velocity_norm(st["x"])
    ija = capi.pattern_msr(p)
    rc, a_ref, r_ref = port.port_fill(p, ija, st, h_elem_avg=h, U_norm=U, **kw)
    assert rc == 0
    mf = MatrixFill(p)
    for scatter in (0, 2):
        mf.set_option("scatter", scatter)
        err, a, r = mf.matrix_fill_full(st["x"], st.get("x_old"), st.get("x_older"), st.get("xdot"), st.get("xdot_old"),
                                        h_elem_avg=h, U_norm=U, **kw)
        assert err == 0
        assert_close(a, a_ref, f"Jacobian scatter={scatter}", ija=ija)
        assert_close(r, r_ref, f"residual scatter={scatter}")
    mf.close()


@pytest.mark.skipif(not ref_driver.ref_available(), reason="oracle/_ref binary not present on this box")
def test_gpu_matches_reference_live(built):
    """Where the reference binary travelled with the snapshot: run it live on a fresh case."""
    m = box_mesh("HEX27", (3, 3, 3), perturb=0.1, seed=21)
    p = Problem(m, energy=True, rho=0.9, mu=0.04, k=0.2, Cp=2.0, beta=0.3, Tref=0.5, gravity=(0.3, 0.0, -1.0),
                ns_source="BOUSSINESQ", heat_source=0.1,
                bcs=[Dirichlet("U", 5, 0.0), Dirichlet("V", 5, 0.0), Dirichlet("W", 5, 0.0), Dirichlet("U", 6, 1.0),
                     Dirichlet("T", 1, 1.0), Dirichlet("T", 2, 0.0, relax=1.0), Dirichlet("P", 7, 0.0)])
    st = make_state(p, seed=13)
    ref = ref_driver.run_fill(p, [st])[0]
    mf = MatrixFill(p)
    err, a, r = mf.matrix_fill_full(st["x"])
    assert err == ref["err"] == 0
    assert_close(a, ref["a"], "Jacobian vs live reference", ija=mf.export_msr())
    assert_close(r, ref["resid"], "residual vs live reference")
    mf.close()


def _ale_problem(et, n, energy=False, transient=False, seed=41):
    m = box_mesh(et, n, perturb=0.08, seed=seed)
    bcs = [Dirichlet("U", 1, 1.0), Dirichlet("V", 1, 0.0), Dirichlet("DX", 1, 0.0), Dirichlet("DY", 1, 0.0),
           Dirichlet("DX", 3, 0.01, relax=1.0), Dirichlet("U", 4, 0.0), Dirichlet("DY", 4, 0.0)]
    if m.dim == 3:
        bcs += [Dirichlet("DZ", 5, 0.0), Dirichlet("W", 6, 0.5, relax=1.0)]
    if energy:
        bcs += [Dirichlet("T", 1, 1.0), Dirichlet("T", 2, 0.0, relax=1.0)]
    kw = dict(etm_momentum=(1.0, 1.0, 1.0, 1.0, 1.0, 0.0), etm_energy=(1.0,) * 5, etm_mesh=(1.0,) * 5) if transient else {}
    return Problem(m, ale=True, energy=energy, transient=transient, rho=1.1, mu=0.4, k=0.06, Cp=1.3, beta=0.7, Tref=0.2,
                   gravity=(0.2, -0.3, 0.1), ns_source="BOUSSINESQ" if energy else "CONSTANT", heat_source=0.4,
                   lame_mu=0.8, lame_lambda=1.9, bcs=bcs, **kw)


@pytest.mark.parametrize("et,n,energy,transient", [("QUAD9", (20, 8), False, False), ("QUAD9", (9, 7), True, True),
                                                   ("HEX27", (4, 3, 3), False, False), ("HEX27", (3, 3, 2), True, True)])
def test_gpu_c4_ale_matches_port_oracle(built, et, n, energy, transient):
    """Config C4 (ALE pseudo-solid mesh): all mesh-sensitivity blocks J_d_d, J_m_d, J_c_d, J_e_d on seeded
    meshes larger than the fixtures, against the CPU restatement (itself pinned to the reference)."""
    p = _ale_problem(et, n, energy, transient)
    st = make_state(p, seed=23, transient=transient, delta_t=0.02, theta=0.5)
    kw = dict(delta_t=0.02, theta=0.5) if transient else {}
    ija = capi.pattern_msr(p)
    rc, a_ref, r_ref = port.port_fill(p, ija, st, **kw)
    assert rc == 0
    mf = MatrixFill(p)
    for scatter in (0, 1, 2, 2):
        mf.set_option("scatter", scatter)
        err, a, r = mf.matrix_fill_full(st["x"], st.get("x_old"), st.get("x_older"), st.get("xdot"), st.get("xdot_old"), **kw)
        assert err == 0 and not mf.flags.any()
        assert_close(a, a_ref, f"Jacobian scatter={scatter}", ija=ija)
        assert_close(r, r_ref, f"residual scatter={scatter}")
    mf.close()


def test_gpu_c4_ale_jacobian_is_derivative_of_residual(built):
    """numerical_jacobian-style check (mm_numjac.c:716) of the full ALE block structure at a size the CPU
    oracle would need minutes for: J d == (R(x + eps d) - R(x - eps d)) / (2 eps)."""
    p = _ale_problem("HEX27", (8, 8, 8), energy=True)
    st = make_state(p, seed=4)
    mf = MatrixFill(p)
    n = mf.num_unknowns
    err, a, _ = mf.matrix_fill_full(st["x"])
    assert err == 0
    A = msr_to_csr(mf.export_msr(), a, n)
    d = np.random.default_rng(0).normal(size=n)
    first, node_kind, kinds = p.unknown_map()
    eps = 1e-7
    _, _, rp = mf.matrix_fill_full(st["x"] + eps * d, assemble_jacobian=False)
    rp = rp.copy()
    _, _, rm = mf.matrix_fill_full(st["x"] - eps * d, assemble_jacobian=False)
    fd, jd = (rp - rm) / (2 * eps), A @ d
    is_dbc, _, hard = p.dirichlet_table()
    keep = ~((is_dbc == 1) & (hard == 1))  # hard-set rows: residual 0, diagonal 1 (bc_dirich.c:134-135)
    assert np.abs(fd - jd)[keep].max() / np.abs(jd).max() < 1e-6
    mf.close()


def test_gpu_c4_ale_inverted_element_is_a_domain_failure(built):
    """A displacement field that folds an element makes belly_flop raise neg_elem_volume and matrix_fill_full
    return -1 (mm_fill_solid.c:659-663,811-815; mm_fill.c:285-311); the GPU entry reports the same."""
    p = _ale_problem("QUAD9", (6, 4))
    st = make_state(p, seed=3)
    first, node_kind, kinds = p.unknown_map()
    x = st["x"].copy()
    nd = p.mesh.conn[7, 2]  # drag one corner node across its element
    off = kinds[node_kind[nd]].index("DX")
    x[first[nd] + off] -= 1.2 * (2.0 / 6)
    x[first[nd] + off + 1] -= 1.2 * (1.0 / 4)
    rc, _, _ = port.port_fill(p, capi.pattern_msr(p), {"x": x})
    assert rc == -1
    mf = MatrixFill(p)
    err, _, _ = mf.matrix_fill_full(x)
    assert err == -1 and mf.flags[0] == 1
    mf.close()


def test_gpu_c1_newton_iterations_and_solution_match_oracle(built):
    """BASELINE.json configs[0] (C1): the full Newton solve of the 40x10 quad9 channel, once with the CUDA
    fill and once with the CPU oracle as the assembly: same iteration count, converged solutions agree to
    1e-10 (north_star), and at the converged state the assembled systems agree to 1e-12."""
    from tests.newton_util import channel_problem, newton

    p = channel_problem()
    ija = capi.pattern_msr(p)
    n = int(p.unknown_map()[0][-1])
    x0 = p.preset_dirichlet(np.zeros(n))
    mf = MatrixFill(p)
    xg, itg, ng = newton(lambda x: mf.matrix_fill_full(x), ija, x0)
    xo, ito, no = newton(lambda x: port.port_fill(p, ija, {"x": x}), ija, x0)
    assert itg == ito and 2 <= itg <= 8, (itg, ito, ng, no)
    assert np.abs(xg - xo).max() <= 1e-10 * max(1.0, np.abs(xo).max())
    # the flow developed: parabolic-ish profile with centreline velocity > inflow value at mid-channel
    first, node_kind, kinds = p.unknown_map()
    mid = np.nonzero((np.abs(p.mesh.coords[0] - 2.0) < 1e-12) & (np.abs(p.mesh.coords[1] - 0.5) < 1e-12))[0][0]
    assert xg[first[mid]] > 1.2
    _, a, r = mf.matrix_fill_full(xg)
    _, a_ref, r_ref = port.port_fill(p, ija, {"x": xg})
    assert_close(a, a_ref, "Jacobian at the converged state", ija=ija)
    assert np.abs(r - r_ref).max() < 1e-12
    if ref_driver.ref_available():  # the reference's own assembly at the GPU-converged state
        ref = ref_driver.run_fill(p, [{"x": xg}])[0]
        assert_close(a, ref["a"], "Jacobian vs live reference at the converged state", ija=ija)
        assert np.abs(r - ref["resid"]).max() < 1e-12
    mf.close()


def test_gpu_global_h_and_velocity_norm(built):
    """goma_gpu_global_h_U == the host restatement of global_h_elem_siz / global_velocity_norm
    (mm_fill_aux.c:1128-1207, :612-680) that the PSPG parity tests feed to both sides."""
    m = box_mesh("HEX8", (7, 6, 5), perturb=0.12, seed=31)
    p = Problem(m, interp="Q1Q1", pspg="global", energy=True, ns_source="BOUSSINESQ")
    st = make_state(p, seed=17)
    mf = MatrixFill(p)
    mf.matrix_fill_full(st["x"], h_elem_avg=0.1, U_norm=1.0, assemble_jacobian=False)  # x into HBM
    sh, ne, sv, nv = mf.global_h_U()
    assert ne == m.num_elems and nv == 3 * m.num_nodes
    assert abs(sh / ne - p.global_h_elem_siz()) < 1e-13
    assert abs(sv / nv - p.global_velocity_norm(st["x"])) < 1e-13 * max(1.0, sv / nv)
    owned = np.zeros(m.num_elems, np.uint8)
    owned[::3] = 1
    sh2, ne2, _, _ = mf.global_h_U(owned)
    assert ne2 == owned.sum() and abs(sh2 - p.global_h_elem_siz(owned.astype(bool))) < 1e-12
    m2 = box_mesh("QUAD9", (9, 7), perturb=0.1, seed=3)
    p2 = Problem(m2)
    st2 = make_state(p2, seed=2)
    mf2 = MatrixFill(p2)
    mf2.matrix_fill_full(st2["x"], assemble_jacobian=False)
    sh, ne, sv, nv = mf2.global_h_U()
    assert abs(sh / ne - p2.global_h_elem_siz()) < 1e-13 and abs(sv / nv - p2.global_velocity_norm(st2["x"])) < 1e-13
    mf.close()
    mf2.close()


def test_gpu_row_sum_scaling_and_norms_match_reference(built):
    """row_sum_scaling_scale and the residual norms on the device-resident system ==
    the reference's own row_sum_scale_MSR / Loo_norm / L1_norm / L2_norm output (committed fixture)."""
    p, kw, st = case_state("c1_quad9_ns")
    q = golden("post_c1_quad9_ns")
    mf = MatrixFill(p)
    err, _, _ = mf.matrix_fill_full(st["x"])
    assert err == 0
    scale, zero_rows = mf.row_sum_scale()
    a, r = mf.download_system()
    assert zero_rows == 0
    assert_close(scale, q["scale"], "scale")
    assert_close(a[:-1], q["a"][:-1], "scaled matrix")
    assert_close(r, q["resid"], "scaled residual")
    loo, l1, l2, k = mf.vector_norms(0)
    assert abs(loo - q["norms"][0]) < 1e-12 and abs(l1 - q["norms"][1]) < 1e-11 and abs(l2 - q["norms"][2]) < 1e-12
    assert k == int(q["norms"][3])
    mf.close()


def test_gpu_row_sum_scaling_matches_oracle_seeded(built):
    """Larger seeded hex27 system with ghost rows: only owned rows are scaled; against oracle/post_fill.py."""
    from oracle import post_fill

    m = box_mesh("HEX27", (3, 3, 2), perturb=0.1, seed=5)
    p = Problem(m, energy=True, rho=1.2, mu=0.03, k=0.05, Cp=1.3, ns_source="BOUSSINESQ", gravity=(0, 0, -1.0),
                bcs=[Dirichlet("U", 1, 1.0), Dirichlet("T", 2, 0.0, relax=1.0)])
    st = make_state(p, seed=6)
    owned = m.num_nodes - 30
    n_owned = int(p.unknown_map()[0][owned])
    mf = MatrixFill(p, num_owned_nodes=owned)
    ija = mf.export_msr()
    _, a0, r0 = mf.matrix_fill_full(st["x"])
    a_ref, r_ref, s_ref = post_fill.row_sum_scale_msr(n_owned, a0, ija, r0)
    scale, zero_rows = mf.row_sum_scale()
    a, r = mf.download_system()
    assert zero_rows == 0 and len(scale) == n_owned
    assert_close(scale, s_ref[:n_owned], "scale")
    assert_close(a[:-1], a_ref[:-1], "scaled matrix")
    assert_close(r, r_ref, "scaled residual")
    loo, l1, l2, k = mf.vector_norms(0)
    ref = post_fill.norms(r_ref, n_owned)
    assert abs(loo - ref[0]) < 1e-12 * ref[0] + 1e-300 and abs(l2 - ref[2]) < 1e-12 * ref[2] and k == ref[3]
    mf.close()


@pytest.mark.parametrize("et,n", [("HEX27", (1, 1, 1)), ("QUAD9", (1, 1)), ("HEX8", (1, 1, 1)), ("QUAD4", (2, 1))])
def test_gpu_single_element_and_tiny_meshes(built, et, n):
    """Edge of the size range: one element (one CTA, one colour, every slot a first touch), all scatter modes."""
    m = box_mesh(et, n, perturb=0.0)
    q1 = et in ("HEX8", "QUAD4")
    p = Problem(m, interp="Q1Q1" if q1 else "Q2P1", pspg="local" if q1 else None, rho=1.1, mu=0.3,
                gravity=(0.1, -0.2, 0.3), bcs=[Dirichlet("U", 1, 1.0), Dirichlet("V", 1, 0.0, relax=1.0)])
    st = make_state(p, seed=1)
    ija = capi.pattern_msr(p)
    rc, a_ref, r_ref = port.port_fill(p, ija, st)
    assert rc == 0
    mf = MatrixFill(p)
    for scatter in (0, 1, 2):
        mf.set_option("scatter", scatter)
        err, a, r = mf.matrix_fill_full(st["x"])
        assert err == 0
        assert_close(a, a_ref, f"Jacobian scatter={scatter}", ija=ija)
        assert_close(r, r_ref, f"residual scatter={scatter}")
    mf.close()


def test_gpu_first_touch_fill_is_bit_reproducible_and_idempotent(built):
    """Scatter mode 2 writes every touched slot exactly once per colour order: two fills of the same state give
    the same bits, and a fill after a different state leaves no trace of it (no memset needed)."""
    m = box_mesh("HEX27", (6, 5, 4), perturb=0.1, seed=9)
    p = Problem(m, rho=1.0, mu=0.02, bcs=[Dirichlet("U", 6, 1.0), Dirichlet("W", 5, 0.0, relax=1.0)])
    s1, s2 = make_state(p, seed=1), make_state(p, seed=2)
    mf = MatrixFill(p)
    _, a1, r1 = mf.matrix_fill_full(s1["x"])
    a1, r1 = a1.copy(), r1.copy()
    mf.matrix_fill_full(s2["x"])
    _, a3, r3 = mf.matrix_fill_full(s1["x"])
    np.testing.assert_array_equal(a1, a3)
    np.testing.assert_array_equal(r1, r3)
    # few persistent CTAs (grid_limit) walk the same colour lists: same bits again
    mf.set_option("grid_limit", 3)
    _, a4, r4 = mf.matrix_fill_full(s1["x"])
    np.testing.assert_array_equal(a1, a4)
    np.testing.assert_array_equal(r1, r4)
    mf.close()


def test_gpu_all_rows_ghost_or_dirichlet(built):
    """Degenerate ownership: a rank that owns no node writes nothing; a problem whose velocity is Dirichlet
    everywhere still assembles the pressure rows."""
    m = box_mesh("QUAD9", (4, 3), perturb=0.1, seed=2)
    p = Problem(m, bcs=[Dirichlet("U", 1, 1.0)])
    st = make_state(p, seed=3)
    mf = MatrixFill(p, num_owned_nodes=0)
    err, a, r = mf.matrix_fill_full(st["x"])
    assert err == 0 and not a.any() and not r.any()
    mf.close()
    bcs = [Dirichlet(v, s, 0.0) for s in (1, 2, 3, 4) for v in "UV"]
    p2 = Problem(m, bcs=bcs)
    ija = capi.pattern_msr(p2)
    rc, a_ref, r_ref = port.port_fill(p2, ija, st)
    mf2 = MatrixFill(p2)
    err, a, r = mf2.matrix_fill_full(st["x"])
    assert err == 0
    assert_close(a, a_ref, "Jacobian", ija=ija)
    assert_close(r, r_ref, "residual")
    mf2.close()


def test_gpu_full_size_c2_properties(built):
    """BASELINE.json configs[1] at its full size (100^3 = 1M hex27 elements, 28.4M unknowns, 5.6e9 non-zeros --
    beyond what the CPU oracle or a 32-bit ija can hold): size-independent properties of the assembled system.
      * a uniform velocity field with zero pressure solves the unforced Navier-Stokes equations: residual == 0;
      * with advection off the residual is affine in x: R(x1 + x2) - R(x1) - R(x2) + R(0) == 0;
      * a uniform shift of the velocity changes no momentum residual of the Stokes problem (only gradients enter);
      * the device-resident Jacobian of two fills of the same state is bit-identical (first-touch scatter)."""
    import torch

    from goma_b200.matrix_fill import device_view

    m = box_mesh("HEX27", (100, 100, 100))
    p = Problem(m, rho=1.0, mu=0.01, gravity=(0.0, 0.0, 0.0))
    first, node_kind, kinds = p.unknown_map()
    n = int(first[-1])
    assert n == 28361803
    mf = MatrixFill(p)
    assert mf.nnz_plus > 2 ** 31  # the reference's int ija cannot index this matrix
    vel = np.zeros(n, bool)
    for kind_id, slots in enumerate(kinds):
        nodes = np.nonzero(node_kind == kind_id)[0]
        for name in "UVW":
            vel[first[nodes] + slots.index(name)] = True
    uniform = np.zeros(n)
    uniform[vel] = 0.7
    err, _, r = mf.matrix_fill_full(uniform, assemble_jacobian=False)
    assert err == 0 and np.abs(r).max() < 1e-12
    mf.close()
    # Stokes: affine residual
    p2 = Problem(m, rho=1.0, mu=0.5, etm_momentum=(0.0, 0.0, 1.0, 1.0, 1.0, 0.0), gravity=(0.1, 0.2, -0.3))
    mf = MatrixFill(p2)
    rng = np.random.default_rng(5)
    x1, x2 = rng.normal(size=n), rng.normal(size=n)
    res = lambda x: mf.matrix_fill_full(x, assemble_jacobian=False)[2].copy()
    r0, r1, r2, r12 = res(np.zeros(n)), res(x1), res(x2), res(x1 + x2)
    assert np.abs(r12 - r1 - r2 + r0).max() < 1e-11 * np.abs(r12).max()
    rs = res(x1 + uniform)
    assert np.abs((rs - r1)[vel]).max() < 1e-11 * np.abs(r1).max()
    # device-resident Jacobian: two fills, identical bits
    bufs = mf.device_buffers()
    d_a = device_view(bufs.d_a, mf.nnz_plus + 1, torch.device("cuda", 0))
    mf.matrix_fill_full(x1, assemble_jacobian=False)  # x1 into HBM
    assert mf.fill_device() == 0
    torch.cuda.synchronize()
    s1 = d_a.view(torch.int64).sum().item()
    sa1 = float(d_a.abs().sum().item())
    assert mf.fill_device() == 0
    torch.cuda.synchronize()
    assert d_a.view(torch.int64).sum().item() == s1 and np.isfinite(sa1) and sa1 > 0
    mf.close()


@pytest.mark.parametrize("et,n,energy,ns,ale,transient", [
    ("QUAD9", (9, 6), False, 1, False, False), ("QUAD9", (7, 5), True, 2, False, True),
    ("HEX27", (3, 3, 2), False, 1, False, True), ("HEX27", (3, 2, 2), True, 2, False, False),
    ("QUAD9", (8, 5), False, 1, True, True), ("QUAD9", (6, 5), True, 2, True, False),
    ("HEX27", (3, 2, 2), False, 1, True, False), ("HEX27", (2, 2, 2), True, 2, True, True)])
def test_gpu_q2p1_species_field_sets_match_port_oracle(built, et, n, energy, ns, ale, transient):
    """Q2/P1 with Fickian species (assemble_mass_transport, mm_fill_species.c:194; J_s_s, J_s_v and, on a moving
    mesh, J_s_d :1103-1330), with and without energy and ALE, against the CPU restatement (pinned to the
    reference on fixture q2p1_quad9_species_ale_transient and on live hex27 cases)."""
    m = box_mesh(et, n, perturb=0.08, seed=51)
    bcs = [Dirichlet("U", 1, 1.0), Dirichlet("V", 1, 0.0), Dirichlet("U", 4, 0.0, relax=1.0),
           Dirichlet("Y", 3, 0.7, species=ns - 1), Dirichlet("Y", 2, 0.2, species=0, relax=1.0)]
    if ale:
        bcs += [Dirichlet("DX", 1, 0.0), Dirichlet("DY", 1, 0.0), Dirichlet("DX", 3, 0.01, relax=1.0)]
    if energy:
        bcs += [Dirichlet("T", 1, 1.0)]
    kw = dict(etm_momentum=(1.0, 1.0, 1.0, 1.0, 1.0, 0.0), etm_energy=(1.0,) * 5, etm_species=(1.0,) * 5,
              etm_mesh=(1.0,) * 5) if transient else {}
    p = Problem(m, ale=ale, transient=transient, energy=energy, n_species=ns, diffusivity=(0.05, 0.11, 1.0, 1.0),
                k=0.07, Cp=1.4, beta=0.8, Tref=0.3, ns_source="BOUSSINESQ" if energy else "CONSTANT", heat_source=0.6,
                rho=1.3, mu=0.7, gravity=(0.3, -0.2, 0.1), lame_mu=0.9, lame_lambda=1.7, bcs=bcs, **kw)
    st = make_state(p, seed=29, transient=transient, delta_t=0.02, theta=0.5)
    fkw = dict(delta_t=0.02, theta=0.5) if transient else {}
    ija = capi.pattern_msr(p)
    rc, a_ref, r_ref = port.port_fill(p, ija, st, **fkw)
    assert rc == 0
    mf = MatrixFill(p)
    for scatter in (0, 2):
        mf.set_option("scatter", scatter)
        err, a, r = mf.matrix_fill_full(st["x"], st.get("x_old"), st.get("x_older"), st.get("xdot"), st.get("xdot_old"), **fkw)
        assert err == 0
        assert_close(a, a_ref, f"Jacobian scatter={scatter}", ija=ija)
        assert_close(r, r_ref, f"residual scatter={scatter}")
    mf.close()


def _msr_to_csr_arrays(ija, a, nrows):
    """(rowptr, colind, values) of the first nrows rows with the diagonal merged in, explicit zeros kept."""
    ija = np.asarray(ija, np.int64)
    N = int(ija[0]) - 1
    cnt = np.diff(ija[: nrows + 1])
    rows = np.concatenate([np.repeat(np.arange(nrows), cnt), np.arange(nrows)])
    cols = np.concatenate([ija[ija[0]: ija[nrows]], np.arange(nrows)])
    vals = np.concatenate([a[ija[0]: ija[nrows]], a[:nrows]])
    order = np.lexsort((cols, rows))
    rowptr = np.zeros(nrows + 1, np.int64)
    np.cumsum(cnt + 1, out=rowptr[1:])
    assert N >= nrows
    return rowptr, cols[order].astype(np.int32), vals[order]


@pytest.mark.parametrize("name", ["c3_hex27_boussinesq", "c5_hex8_pspg_global", "c4_quad9_ale"])
def test_gpu_csr_handoff_matches_msr(built, name):
    """the device-resident CSR view (sorted columns, diagonal in place) holds exactly the
    assembled MSR system of the reference fixture; with ghost rows only the owned rows are exported."""
    import torch

    p, kw, st = case_state(name)
    g = golden(name)
    n = len(g["resid"])
    mf = MatrixFill(p)
    h, U = (p.global_h_elem_siz(), p.global_velocity_norm(st["x"])) if p.pspg else (0.0, 0.0)
    err, a, r = mf.matrix_fill_full(st["x"], h_elem_avg=h, U_norm=U)
    assert err == 0
    rowptr, colind, values = (t.cpu().numpy() for t in mf.csr())
    rp, ci, va = _msr_to_csr_arrays(g["ija"], g["a"], n)  # from the REFERENCE's graph and values
    np.testing.assert_array_equal(rowptr, rp)
    np.testing.assert_array_equal(colind, ci)
    assert_close(values, va, "CSR values")
    mf.close()
    # owned rows only
    owned = p.mesh.num_nodes - 7
    n_owned = int(p.unknown_map()[0][owned])
    mf = MatrixFill(p, num_owned_nodes=owned)
    err, a, r = mf.matrix_fill_full(st["x"], h_elem_avg=h, U_norm=U)
    rowptr, colind, values = (t.cpu().numpy() for t in mf.csr())
    assert len(rowptr) == n_owned + 1
    rp, ci, va = _msr_to_csr_arrays(mf.export_msr(), a, n_owned)
    np.testing.assert_array_equal(rowptr, rp)
    np.testing.assert_array_equal(colind, ci)
    assert_close(values, va, "CSR values (owned rows)")
    mf.close()


# ------------------------------------------------------------------ round 2: evidence the round-1 review asked for
@pytest.mark.skipif(not ref_driver.ref_available(), reason="oracle/_ref binary not present on this box")
@pytest.mark.parametrize("n,energy", [(8, False), (8, True), (16, False), (16, True)])
def test_gpu_matches_reference_live_cavity(built, n, energy):
    """parity of configs C2 / C3 against the reference's own matrix_fill_full, run live on the box,
    on 8^3 and 16^3 hex27 sub-problems of the lid-driven cavity (perturbed nodes, seeded state), per-row tolerance."""
    m = box_mesh("HEX27", (n, n, n), perturb=0.1, seed=100 + n)
    bcs = [Dirichlet(v, s, 0.0) for s in (1, 2, 3, 4, 5) for v in "UVW"]
    bcs += [Dirichlet("U", 6, 1.0), Dirichlet("V", 6, 0.0), Dirichlet("W", 6, 0.0), Dirichlet("P", 7, 0.0)]
    kw = {}
    if energy:
        bcs += [Dirichlet("T", 1, 1.0), Dirichlet("T", 2, 0.0)]
        kw = dict(energy=True, k=0.0141, Cp=1.0, beta=1.0, Tref=0.0, gravity=(0.0, 0.0, -1.0), ns_source="BOUSSINESQ")
    p = Problem(m, rho=1.0, mu=0.01, bcs=bcs, **kw)
    st = make_state(p, seed=n)
    ref = ref_driver.run_fill(p, [st])[0]
    mf = MatrixFill(p)
    err, a, r = mf.matrix_fill_full(st["x"])
    assert err == ref["err"] == 0
    assert_close(a, ref["a"], f"Jacobian vs live reference, {n}^3", ija=mf.export_msr())
    assert_close(r, ref["resid"], f"residual vs live reference, {n}^3")
    mf.close()


@pytest.mark.parametrize("et", ["HEX27", "QUAD9", "HEX8"])
def test_gpu_tiny_detJ_is_assembled_like_the_reference(built, et):
    """|detJ| < 1e-10 at every Gauss point (a mesh in micrometre-sized units).  The reference raises zero_detJ only
    inside beer_belly's shell-element branch (mm_fill_util.c:312-344): continuum elements are assembled normally,
    return 0, and so must the GPU path (round 1 wrongly flagged them).  Per-row tolerance: the entries are ~1e-12."""
    dim = 2 if et == "QUAD9" else 3
    L = 3e-5 if dim == 2 else 8e-4  # detJ = prod(h_d / 2) < 2e-11
    m = box_mesh(et, (4, 3) if dim == 2 else (2, 2, 2), hi=(L,) * dim, perturb=0.1, seed=3)
    q1 = et == "HEX8"
    p = Problem(m, interp="Q1Q1" if q1 else "Q2P1", pspg="global" if q1 else None, rho=1.1, mu=0.3,
                gravity=(0.1, -0.2, 0.3 if dim == 3 else 0.0), bcs=[Dirichlet("U", 1, 1.0)])
    st = make_state(p, seed=1)
    ija = capi.pattern_msr(p)
    h, U = (p.global_h_elem_siz(), p.global_velocity_norm(st["x"])) if q1 else (0.0, 0.0)
    rc, a_ref, r_ref = port.port_fill(p, ija, st, h_elem_avg=h, U_norm=U)
    assert rc == 0
    mf = MatrixFill(p)
    err, a, r = mf.matrix_fill_full(st["x"], h_elem_avg=h, U_norm=U)
    assert err == 0 and not mf.flags.any()
    assert_close(a, a_ref, "Jacobian", ija=ija)
    assert_close(r, r_ref, "residual")
    if ref_driver.ref_available():
        ref = ref_driver.run_fill(p, [st])[0]
        assert ref["err"] == 0 and ref["zero_detJ"] == 0
        assert_close(a, ref["a"], "Jacobian vs live reference", ija=ija)
        assert_close(r, ref["resid"], "residual vs live reference")
    mf.close()


def test_gpu_irregular_valence_mesh_all_scatter_modes(built):
    """An unstructured hex27 mesh (vertex valences 3 and 5, extruded) larger than the fixture: the greedy element
    colouring needs more colours than a lattice, node-node lists have irregular lengths; all three scatter modes
    against the CPU restatement (itself pinned to the reference on fixtures irr_*) and, where present, against
    the live reference including the MSR graph."""
    from goma_b200.mesh import star_mesh

    m = star_mesh("HEX27", refine=1, nz=3, perturb=0.3, seed=5)
    bcs = [Dirichlet("U", 1, 0.0), Dirichlet("V", 1, 0.5, relax=1.0), Dirichlet("W", 5, 0.0), Dirichlet("T", 5, 1.0),
           Dirichlet("P", 7, 0.0)]
    p = Problem(m, energy=True, rho=1.1, mu=0.3, k=0.2, Cp=1.3, beta=0.5, Tref=0.1, ns_source="BOUSSINESQ",
                gravity=(0.1, -0.2, 0.3), heat_source=0.2, bcs=bcs)
    st = make_state(p, seed=9)
    ija = capi.pattern_msr(p)
    rc, a_ref, r_ref = port.port_fill(p, ija, st)
    assert rc == 0
    mf = MatrixFill(p)
    np.testing.assert_array_equal(mf.export_msr(), ija)
    for scatter in (0, 1, 2, 2):
        mf.set_option("scatter", scatter)
        err, a, r = mf.matrix_fill_full(st["x"])
        assert err == 0
        assert_close(a, a_ref, f"Jacobian scatter={scatter}", ija=ija)
        assert_close(r, r_ref, f"residual scatter={scatter}")
    if ref_driver.ref_available():
        np.testing.assert_array_equal(ref_driver.run_map(p)["ija"][:-1], ija)
        ref = ref_driver.run_fill(p, [st])[0]
        assert_close(a, ref["a"], "Jacobian vs live reference", ija=ija)
    mf.close()


def test_gpu_async_fill_and_accumulate_option(built):
    """goma_gpu_fill_device_async + goma_gpu_fill_wait give the bits of the synchronous call; with the
    "accumulate" option goma_gpu_fill adds to what the caller holds in a / resid_vector (the reference's +=,
    mm_fill.c:5390,5463) instead of overwriting, and the next plain fill is clean again."""
    p, kw, st = case_state("c3_hex27_boussinesq")
    mf = MatrixFill(p)
    err, a0, r0 = mf.matrix_fill_full(st["x"])
    a0, r0 = a0.copy(), r0.copy()
    ev = mf.fill_device_async()
    assert ev != 0
    assert mf.fill_wait() == 0
    a1, r1 = mf.download_system()
    np.testing.assert_array_equal(a1, a0)
    np.testing.assert_array_equal(r1, r0)
    with pytest.raises(capi.GomaGpuError):
        mf.fill_wait()  # nothing pending
    mf.set_option("accumulate", 1)
    rng = np.random.default_rng(3)
    a_pre, r_pre = rng.normal(size=len(a0)), rng.normal(size=len(r0))
    err, a2, r2 = mf.matrix_fill_full(st["x"], a=a_pre.copy(), resid_vector=r_pre.copy())
    assert err == 0
    ija = mf.export_msr()
    n = len(r0)
    np.testing.assert_allclose(r2, r_pre + r0, rtol=0, atol=1e-12 * np.abs(r0).max())
    keep = np.ones(len(a0), bool)
    keep[n] = False  # a[N] is unused by MSR
    np.testing.assert_allclose(a2[keep], (a_pre + a0)[keep], rtol=0, atol=1e-12 * np.abs(a0).max())
    mf.set_option("accumulate", 0)
    err, a3, r3 = mf.matrix_fill_full(st["x"])
    np.testing.assert_array_equal(a3[keep], a0[keep])
    np.testing.assert_array_equal(r3, r0)
    assert len(ija) == mf.nnz_plus
    mf.close()


def test_gpu_zero_row_scaling_does_not_poison_later_fills(built):
    """A zero row sum makes row_sum_scale_MSR write NaN (0 * inf) into slots no element touches; the reference
    re-zeroes its storage before the next fill (mm_sol_nonlinear.c:1109-1121).  The first-touch scatter never
    rewrites those slots, so the library must re-zero once: the fill after such a scaling is clean."""
    m = box_mesh("QUAD9", (3, 3), perturb=0.1, seed=2)
    # no pressure datum and zero velocity everywhere: continuity rows sum to zero only if div-free; force an exactly
    # zero row instead: all-Dirichlet velocity AND zero viscosity / density make the centroid momentum rows vanish
    p = Problem(m, rho=0.0, mu=0.0, etm_momentum=(0.0, 0.0, 0.0, 0.0, 0.0, 0.0), etm_continuity=(0.0, 0.0))
    st = make_state(p, seed=3)
    mf = MatrixFill(p)
    err, a0, r0 = mf.matrix_fill_full(st["x"])
    assert err == 0 and not a0.any()
    scale, zero_rows = mf.row_sum_scale()
    assert zero_rows == len(r0)
    a1, _ = mf.download_system()
    assert np.isnan(a1).any()
    err, a2, r2 = mf.matrix_fill_full(st["x"])
    assert err == 0 and not np.isnan(a2).any() and not a2.any()
    mf.close()


@pytest.mark.parametrize("name", ["c2_hex27_ns", "c3_hex27_boussinesq", "c5_hex8_pspg_global", "c4_quad9_ale",
                                  "c1_quad9_ns_transient", "irr_hex27_star_bouss", "q2p1_quad9_species_ale_transient",
                                  "mm_hex27_bouss_2mat", "mm_hex8_pspg_2mat_transient"])
@pytest.mark.parametrize("scatter", [0, 2], ids=["atomic", "first_touch"])
def test_gpu_csr_layout_is_assembled_in_place(built, name, scatter):
    """matrix_layout = CSR: the element blocks are scattered straight into the CSR value array (diagonal at its sorted
    position, owned rows only) -- no second copy of the matrix.  The values equal the reference fixture's MSR system
    re-ordered; rowptr / colind equal the reference graph; the residual is the same; with ghost nodes only owned rows
    exist; row-sum scaling works on the CSR array in place."""
    p, kw, st = case_state(name)
    g = golden(name)
    n = len(g["resid"])
    fkw = dict(delta_t=kw.get("delta_t", 0.0), theta=kw.get("theta", 0.0), time_value=kw.get("time", 0.0))
    h, U = (p.global_h_elem_siz(), p.global_velocity_norm(st["x"])) if p.pspg else (0.0, 0.0)
    args = (st["x"], st.get("x_old"), st.get("x_older"), st.get("xdot"), st.get("xdot_old"))
    mf = MatrixFill(p, layout="csr")
    mf.set_option("scatter", scatter)
    rp, ci, va = _msr_to_csr_arrays(g["ija"], g["a"], n)  # from the REFERENCE's graph and values
    assert mf.value_count == len(va)
    err, a, r = mf.matrix_fill_full(*args, h_elem_avg=h, U_norm=U, **fkw)
    assert err == 0 and len(a) == len(va)
    csr_ija = np.concatenate([[n + 1], n + 1 + np.cumsum(np.diff(rp) - 1)])  # per-row scale of the check below
    scale = np.maximum.reduceat(np.abs(va), rp[:-1])
    assert (np.abs(a - va) <= TOL * np.repeat(scale, np.diff(rp))).all(), "CSR values vs reference (per-row tolerance)"
    assert_close(r, g["resid"], "residual")
    rowptr, colind, values = (t.cpu().numpy() for t in mf.csr())
    np.testing.assert_array_equal(rowptr, rp)
    np.testing.assert_array_equal(colind, ci)
    np.testing.assert_array_equal(values, a)  # the CSR view IS the assembled array
    rowptr2, values2 = (t.cpu().numpy() for t in mf.csr_rows())
    np.testing.assert_array_equal(rowptr2, rp)
    np.testing.assert_array_equal(values2, a)
    # row-sum scaling in place on the CSR array == the MSR path's result re-ordered
    mf_m = MatrixFill(p)
    mf_m.matrix_fill_full(*args, h_elem_avg=h, U_norm=U, **fkw)
    s_m, z_m = mf_m.row_sum_scale()
    am, rm = mf_m.download_system()
    s_c, z_c = mf.row_sum_scale()
    ac, rc_ = mf.download_system()
    assert z_m == z_c == 0
    assert_close(s_c, s_m, "scale")
    _, _, vam = _msr_to_csr_arrays(g["ija"], am, n)
    assert_close(ac, vam, "scaled CSR values")
    assert_close(rc_, rm, "scaled residual")
    mf.close()
    mf_m.close()
    assert len(csr_ija) == n + 1
    # owned rows only
    owned = p.mesh.num_nodes - 5
    n_owned = int(p.unknown_map()[0][owned])
    mf = MatrixFill(p, num_owned_nodes=owned, layout="csr")
    mf_m = MatrixFill(p, num_owned_nodes=owned)
    err, a, r = mf.matrix_fill_full(*args, h_elem_avg=h, U_norm=U, **fkw)
    _, am, rm = mf_m.matrix_fill_full(*args, h_elem_avg=h, U_norm=U, **fkw)
    rp, ci, va = _msr_to_csr_arrays(mf_m.export_msr(), am, n_owned)
    assert len(a) == len(va) == rp[-1]
    assert_close(a, va, "CSR values, owned rows")
    assert_close(r, rm, "residual, owned rows")
    mf.close()
    mf_m.close()


@pytest.mark.gpu
@pytest.mark.parametrize("layout", ["msr", "csr"])
@pytest.mark.parametrize("name,chunks", [("c2_hex27_ns", 3), ("c3_hex27_boussinesq", 4), ("c5_hex8_pspg_global", 5),
                                         ("irr_hex27_star_bouss", 2), ("c4_quad9_ale", 64), ("mm_quad9_ale_3mat", 2),
                                         ("mm_hex27_star_2mat", 3)])
def test_gpu_host_stream_chunks_is_bit_identical(built, name, chunks, layout):
    """host_stream_chunks = K: the elements are swept chunk by chunk and finished rows are copied to the host under the
    assembly of the later chunks.  The order of the contributions to a slot changes with the class order, so the
    comparison with the one-sweep fill is at the parity tolerance; two streamed fills are bit-identical; every value
    must have arrived (the host buffer starts as NaN) and equal the device-resident copy."""
    p, kw, st = case_state(name)
    fkw = dict(delta_t=kw.get("delta_t", 0.0), theta=kw.get("theta", 0.0), time_value=kw.get("time", 0.0))
    h, U = (p.global_h_elem_siz(), p.global_velocity_norm(st["x"])) if p.pspg else (0.0, 0.0)
    args = (st["x"], st.get("x_old"), st.get("x_older"), st.get("xdot"), st.get("xdot_old"))
    mf0 = MatrixFill(p, layout=layout)
    _, a0, r0 = mf0.matrix_fill_full(*args, h_elem_avg=h, U_norm=U, **fkw)
    mf0.close()
    mf = MatrixFill(p, layout=layout, host_stream_chunks=chunks)
    a = np.full(mf.value_count, np.nan)
    r = np.full(mf.num_unknowns, np.nan)
    err, a, r = mf.matrix_fill_full(*args, h_elem_avg=h, U_norm=U, a=a, resid_vector=r, **fkw)
    assert err == 0
    assert not np.isnan(r).any()
    assert not np.isnan(a).any(), "a value never reached the host"
    ija = mf.export_msr() if layout == "msr" else None
    if layout == "msr":
        assert_close(a, a0, "streamed vs one sweep", ija=ija)
    else:
        np.testing.assert_allclose(a, a0, rtol=1e-11, atol=1e-11 * np.abs(a0).max())
    assert_close(r, r0, "residual")
    _, nl = mf.last_stats()
    assert nl > 0
    a2 = np.full(mf.value_count, np.nan)
    err, a2, _ = mf.matrix_fill_full(*args, h_elem_avg=h, U_norm=U, a=a2, **fkw)
    np.testing.assert_array_equal(a2, a)
    # the device-resident copy equals what was streamed
    ad, _ = mf.download_system()
    np.testing.assert_array_equal(ad, a)
    mf.close()


def test_gpu_materials_with_equal_constants_equal_one_material(built):
    """Two element blocks whose materials carry the same constants: the launches are split by material, the values
    must equal the one-material fill to the parity tolerance (the order of the contributions to a slot changes) and
    the residual likewise; a material index out of range is refused."""
    import dataclasses

    p, kw, st = case_state("mm_hex27_bouss_2mat")
    same = dataclasses.replace(p, extra_materials=[{}])
    one = p.single_material(0)
    mf2, mf1 = MatrixFill(same), MatrixFill(one)
    _, a2, r2 = mf2.matrix_fill_full(st["x"])
    _, a1, r1 = mf1.matrix_fill_full(st["x"])
    assert_close(a2, a1, "two equal materials vs one", ija=mf1.export_msr())
    assert_close(r2, r1, "residual")
    _, nl2 = mf2.last_stats()
    _, nl1 = mf1.last_stats()
    assert nl2 > nl1  # one launch per (colour, material)
    mf2.close()
    mf1.close()
    bad = dataclasses.replace(p, mesh=dataclasses.replace(p.mesh, elem_block=p.mesh.elem_block + 1))
    with pytest.raises(RuntimeError, match="elem_material"):
        MatrixFill(bad)


@pytest.mark.parametrize("layout", ["msr", "csr"])
@pytest.mark.parametrize("name", ["c1_quad9_ns", "c3_hex27_boussinesq", "c5_hex8_pspg_global", "c4_quad9_ale_energy_transient",
                                  "irr_hex27_star_bouss", "mm_hex8_pspg_2mat_transient"])
def test_gpu_matvec_matches_msr_oracle(built, name, layout):
    """w = A v on the device-resident matrix, columns taken from the node-level neighbour lists (no column-index array):
    equals the MSR product of the Newton line search (mm_sol_nonlinear.c:442-446, oracle/post_fill.py::msr_matvec) on the
    REFERENCE's graph and values -- energy rows (no pressure columns), P1 and equal-order pressure, ghost rows."""
    import torch

    from oracle import post_fill

    p, kw, st = case_state(name)
    g = golden(name)
    n = len(g["resid"])
    fkw = dict(delta_t=kw.get("delta_t", 0.0), theta=kw.get("theta", 0.0), time_value=kw.get("time", 0.0))
    h, U = (p.global_h_elem_siz(), p.global_velocity_norm(st["x"])) if p.pspg else (0.0, 0.0)
    args = (st["x"], st.get("x_old"), st.get("x_older"), st.get("xdot"), st.get("xdot_old"))
    rng = np.random.default_rng(7)
    v = rng.standard_normal(n)
    want = post_fill.msr_matvec(n, g["a"], g["ija"], v)
    rows = np.repeat(np.arange(n), np.diff(np.asarray(g["ija"][: n + 1], np.int64)))
    rowscale = np.abs(g["a"][:n] * v)
    np.add.at(rowscale, rows, np.abs(g["a"][n + 1: int(g["ija"][n])] * v[np.asarray(g["ija"][n + 1: int(g["ija"][n])], np.int64)]))
    mf = MatrixFill(p, layout=layout)
    err, _, _ = mf.matrix_fill_full(*args, h_elem_avg=h, U_norm=U, **fkw)
    assert err == 0
    w = mf.matvec(torch.from_numpy(v).cuda()).cpu().numpy()
    assert (np.abs(w - want) <= 1e-12 * np.maximum(rowscale, 1e-300)).all(), "w = A v vs the MSR product of the reference values"
    mf.set_option("matvec_cap", 20)  # column lists longer than the staging buffer: the path without shared memory
    w = mf.matvec(torch.from_numpy(v).cuda()).cpu().numpy()
    assert (np.abs(w - want) <= 1e-12 * np.maximum(rowscale, 1e-300)).all(), "w = A v, unstaged path"
    mf.close()
    # owned rows only: the rows of ghost nodes are not written
    owned = p.mesh.num_nodes - 5
    n_owned = int(p.unknown_map()[0][owned])
    mf = MatrixFill(p, num_owned_nodes=owned, layout=layout)
    mf.matrix_fill_full(*args, h_elem_avg=h, U_norm=U, **fkw)
    w = mf.matvec(torch.from_numpy(v).cuda()).cpu().numpy()
    assert (np.abs(w[:n_owned] - want[:n_owned]) <= 1e-12 * np.maximum(rowscale[:n_owned], 1e-300)).all()
    assert not w[n_owned:].any()
    mf.close()
