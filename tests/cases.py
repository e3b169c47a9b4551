"""Parity cases shared by the golden-fixture generator and the tests.

Each case is a small instance of one of the BASELINE.json configs: a perturbed structured
mesh, material constants chosen so that every term has a distinct weight, Dirichlet rows of
both kinds (hard-set and relaxed), and a smooth-plus-noise state so that every Jacobian
term is exercised.
"""
from __future__ import annotations

import numpy as np

from goma_b200.mesh import box_mesh, star_mesh
from goma_b200.problem import Dirichlet, Problem

from goma_b200.state import SEED  # noqa: E402


from goma_b200.state import make_state  # noqa: E402,F401  (the seeded synthetic state of SURVEY.md §8d lives in the package)


def _bcs(dim, energy=False):
    bcs = [Dirichlet("U", 1, 1.0), Dirichlet("V", 1, 0.0), Dirichlet("U", 3, 0.0, relax=1.0),
           Dirichlet("V", 3, 0.25, relax=1.0), Dirichlet("U", 4, 0.0), Dirichlet("V", 4, 0.0),
           Dirichlet("V", 2, 0.0)]
    if dim == 3:
        bcs += [Dirichlet("W", 5, 0.0), Dirichlet("W", 6, 0.5, relax=1.0), Dirichlet("W", 1, 0.0)]
    if energy:
        bcs += [Dirichlet("T", 1, 1.0), Dirichlet("T", 2, 0.0, relax=1.0)]
    return bcs


def build_case(name):
    """Returns (problem, fill kwargs)."""
    if name == "c1_quad9_ns":
        m = box_mesh("QUAD9", (6, 4), lo=(0, 0), hi=(2, 1), perturb=0.15, seed=1)
        p = Problem(m, rho=1.3, mu=0.7, gravity=(0.3, -0.2, 0.0), bcs=_bcs(2) + [Dirichlet("P", 7, 0.5)])
        return p, {}
    if name == "c1_quad9_ns_transient":
        m = box_mesh("QUAD9", (5, 3), lo=(0, 0), hi=(2, 1), perturb=0.15, seed=2)
        p = Problem(m, rho=1.3, mu=0.7, gravity=(0.3, -0.2, 0.0), transient=True,
                    etm_momentum=(1.0, 1.0, 1.0, 1.0, 1.0, 0.0), bcs=_bcs(2))
        return p, {"delta_t": 0.01, "theta": 0.5, "time": 0.2}
    if name == "c2_hex27_ns":
        m = box_mesh("HEX27", (2, 2, 2), perturb=0.12, seed=3)
        p = Problem(m, rho=1.0, mu=0.01, gravity=(0.0, 0.0, -0.4), bcs=_bcs(3))
        return p, {}
    if name == "c3_hex27_boussinesq":
        m = box_mesh("HEX27", (2, 2, 2), perturb=0.12, seed=4)
        p = Problem(m, energy=True, rho=1.1, mu=0.05, k=0.07, Cp=1.4, beta=0.8, Tref=0.3,
                    gravity=(0.0, 0.1, -1.0), ns_source="BOUSSINESQ", heat_source=0.6, bcs=_bcs(3, True))
        return p, {}
    if name == "c3_quad9_bouss_transient":
        m = box_mesh("QUAD9", (4, 4), perturb=0.15, seed=5)
        p = Problem(m, energy=True, rho=1.1, mu=0.05, k=0.07, Cp=1.4, beta=0.8, Tref=0.3,
                    gravity=(0.2, -1.0, 0.0), ns_source="BOUSS", heat_source=0.6, transient=True,
                    etm_momentum=(1.0, 1.0, 1.0, 1.0, 1.0, 0.0), etm_energy=(1.0, 1.0, 1.0, 1.0, 1.0),
                    bcs=_bcs(2, True))
        return p, {"delta_t": 0.02, "theta": 0.0, "time": 0.1}
    if name.startswith("c5_"):
        bcs = [Dirichlet("U", 1, 1.0), Dirichlet("V", 1, 0.0), Dirichlet("U", 4, 0.0, relax=1.0), Dirichlet("T", 1, 1.0),
               Dirichlet("T", 2, 0.0, relax=1.0), Dirichlet("Y", 3, 0.7, species=1), Dirichlet("Y", 4, 0.2, relax=1.0),
               Dirichlet("P", 7, 0.0)]
        common = dict(interp="Q1Q1", energy=True, n_species=2, rho=1.2, mu=0.3, k=0.2, Cp=1.5, beta=0.4, Tref=0.2,
                      diffusivity=(0.05, 0.11, 1.0, 1.0), gravity=(0.1, -0.2, -1.0), ns_source="BOUSSINESQ",
                      heat_source=0.3, ps_scaling=0.1)
        if name == "c5_hex8_pspg_local_transient":
            m = box_mesh("HEX8", (3, 2, 2), perturb=0.15, seed=3)
            p = Problem(m, pspg="local", transient=True, etm_momentum=(1.0, 1.0, 1.0, 1.0, 1.0, 0.0),
                        etm_energy=(1.0,) * 5, etm_species=(1.0,) * 5, bcs=bcs + [Dirichlet("W", 5, 0.0)], **common)
            return p, {"delta_t": 0.05, "theta": 0.0, "time": 0.3}
        if name == "c5_hex8_pspg_global":
            m = box_mesh("HEX8", (3, 3, 2), perturb=0.15, seed=6)
            p = Problem(m, pspg="global", bcs=bcs + [Dirichlet("W", 5, 0.0)], **common)
            return p, {}
        if name == "c5_quad4_pspg_local":
            m = box_mesh("QUAD4", (5, 4), perturb=0.15, seed=8)
            p = Problem(m, pspg="local", bcs=bcs, **common)
            return p, {}
    if name.startswith("c4_"):
        # ALE pseudo-solid mesh (ARBITRARY / NONLINEAR), volumetric assembly only: every mesh-sensitivity block
        bcs2 = [Dirichlet("U", 1, 1.0), Dirichlet("V", 1, 0.0), Dirichlet("DX", 1, 0.0), Dirichlet("DY", 1, 0.0),
                Dirichlet("DX", 3, 0.01, relax=1.0), Dirichlet("DY", 3, 0.0), Dirichlet("U", 4, 0.0),
                Dirichlet("V", 4, 0.0, relax=1.0)]
        mat = dict(ale=True, rho=1.3, mu=0.7, gravity=(0.3, -0.2, 0.1), lame_mu=0.9, lame_lambda=1.7)
        if name == "c4_quad9_ale":
            m = box_mesh("QUAD9", (6, 3), lo=(0, 0), hi=(2, 1), perturb=0.1, seed=11)
            return Problem(m, bcs=bcs2, **mat), {}
        if name == "c4_hex27_ale":
            m = box_mesh("HEX27", (2, 2, 1), perturb=0.1, seed=12)
            bcs3 = bcs2 + [Dirichlet("DZ", 5, 0.0), Dirichlet("W", 6, 0.5, relax=1.0), Dirichlet("DZ", 6, 0.02, relax=1.0)]
            return Problem(m, bcs=bcs3, **mat), {}
        if name == "c4_quad9_ale_energy_transient":
            m = box_mesh("QUAD9", (4, 3), lo=(0, 0), hi=(2, 1), perturb=0.1, seed=13)
            p = Problem(m, transient=True, energy=True, k=0.07, Cp=1.4, beta=0.8, Tref=0.3, ns_source="BOUSSINESQ",
                        heat_source=0.6, etm_momentum=(1.0, 1.0, 1.0, 1.0, 1.0, 0.0), etm_energy=(1.0,) * 5,
                        etm_mesh=(1.0, 1.0, 1.0, 1.0, 1.0), bcs=bcs2 + [Dirichlet("T", 1, 1.0)], **mat)
            return p, {"delta_t": 0.02, "theta": 0.5, "time": 0.1}
    if name == "c2_hex27_ns_transient":
        m = box_mesh("HEX27", (2, 2, 1), perturb=0.12, seed=31)
        p = Problem(m, rho=1.0, mu=0.01, gravity=(0.0, 0.0, -0.4), transient=True,
                    etm_momentum=(1.0, 1.0, 1.0, 1.0, 1.0, 0.0), bcs=_bcs(3))
        return p, {"delta_t": 0.005, "theta": 0.5, "time": 0.05}
    if name == "c4_hex27_ale_energy_transient":
        m = box_mesh("HEX27", (2, 1, 2), perturb=0.1, seed=32)
        bcs = [Dirichlet("U", 1, 1.0), Dirichlet("V", 1, 0.0), Dirichlet("DX", 1, 0.0), Dirichlet("DY", 1, 0.0),
               Dirichlet("DZ", 5, 0.0), Dirichlet("DX", 3, 0.01, relax=1.0), Dirichlet("W", 6, 0.5, relax=1.0),
               Dirichlet("T", 1, 1.0), Dirichlet("T", 2, 0.0, relax=1.0)]
        p = Problem(m, ale=True, transient=True, energy=True, k=0.07, Cp=1.4, beta=0.8, Tref=0.3, ns_source="BOUSSINESQ",
                    heat_source=0.6, etm_momentum=(1.0, 1.0, 1.0, 1.0, 1.0, 0.0), etm_energy=(1.0,) * 5,
                    etm_mesh=(1.0,) * 5, rho=1.3, mu=0.7, gravity=(0.3, -0.2, 0.1), lame_mu=0.9, lame_lambda=1.7, bcs=bcs)
        return p, {"delta_t": 0.02, "theta": 0.0, "time": 0.1}
    if name == "c5_hex8_ns_pspg_local":
        m = box_mesh("HEX8", (3, 2, 3), perturb=0.15, seed=33)
        bcs = [Dirichlet("U", 1, 1.0), Dirichlet("V", 1, 0.0), Dirichlet("W", 5, 0.0), Dirichlet("U", 4, 0.0, relax=1.0),
               Dirichlet("P", 7, 0.0)]
        p = Problem(m, interp="Q1Q1", pspg="local", ps_scaling=0.2, rho=1.2, mu=0.3, gravity=(0.1, -0.2, -1.0), bcs=bcs)
        return p, {}
    if name == "q2p1_quad9_species_ale_transient":
        # Q2/P1 with energy and two Fickian species on a moving mesh: J_s_v, J_s_d, J_e_d and the v - xdot_mesh
        # convection velocity in every transport equation
        m = box_mesh("QUAD9", (4, 3), lo=(0, 0), hi=(2, 1), perturb=0.1, seed=21)
        bcs = [Dirichlet("U", 1, 1.0), Dirichlet("V", 1, 0.0), Dirichlet("DX", 1, 0.0), Dirichlet("DY", 1, 0.0),
               Dirichlet("DX", 3, 0.01, relax=1.0), Dirichlet("U", 4, 0.0), Dirichlet("T", 1, 1.0),
               Dirichlet("Y", 3, 0.7, species=1), Dirichlet("Y", 2, 0.2, species=0, relax=1.0)]
        p = Problem(m, ale=True, transient=True, energy=True, n_species=2, diffusivity=(0.05, 0.11, 1.0, 1.0),
                    etm_momentum=(1.0, 1.0, 1.0, 1.0, 1.0, 0.0), etm_energy=(1.0,) * 5, etm_species=(1.0,) * 5,
                    etm_mesh=(1.0,) * 5, k=0.07, Cp=1.4, beta=0.8, Tref=0.3, ns_source="BOUSSINESQ", heat_source=0.6,
                    rho=1.3, mu=0.7, gravity=(0.3, -0.2, 0.1), lame_mu=0.9, lame_lambda=1.7, bcs=bcs)
        return p, {"delta_t": 0.02, "theta": 0.5, "time": 0.1}
    if name.startswith("irr_"):
        # unstructured meshes: vertex valences 3 and 5 (node-node lists of irregular length, greedy colouring with
        # more colours than a lattice needs) -- what an arbitrary Exodus file gives the reference
        if name == "irr_quad9_star_ns":
            m = star_mesh("QUAD9", refine=1, perturb=0.3, seed=1)
            bcs = [Dirichlet("U", 1, 0.0), Dirichlet("V", 1, 0.5, relax=1.0), Dirichlet("P", 7, 0.0)]
            return Problem(m, rho=1.1, mu=0.3, gravity=(0.1, -0.2, 0.0), bcs=bcs), {}
        if name == "irr_hex27_star_bouss":
            m = star_mesh("HEX27", refine=0, nz=2, perturb=0.3, seed=2)
            bcs = [Dirichlet("U", 1, 0.0), Dirichlet("V", 1, 0.5, relax=1.0), Dirichlet("W", 5, 0.0),
                   Dirichlet("T", 5, 1.0), Dirichlet("T", 6, 0.0, relax=1.0), Dirichlet("P", 7, 0.0)]
            return Problem(m, energy=True, rho=1.1, mu=0.3, k=0.2, Cp=1.3, beta=0.5, Tref=0.1, ns_source="BOUSSINESQ",
                           gravity=(0.1, -0.2, 0.3), heat_source=0.2, bcs=bcs), {}
        if name == "irr_hex8_star_pspg":
            m = star_mesh("HEX8", refine=1, nz=2, perturb=0.3, seed=3)
            bcs = [Dirichlet("U", 1, 0.0), Dirichlet("V", 1, 0.5, relax=1.0), Dirichlet("W", 5, 0.0),
                   Dirichlet("T", 5, 1.0), Dirichlet("Y", 6, 0.3, species=1), Dirichlet("P", 7, 0.0)]
            return Problem(m, interp="Q1Q1", pspg="local", ps_scaling=0.1, energy=True, n_species=2, rho=1.2, mu=0.3,
                           k=0.2, Cp=1.5, beta=0.4, Tref=0.2, diffusivity=(0.05, 0.11, 1.0, 1.0),
                           gravity=(0.1, -0.2, -1.0), ns_source="BOUSSINESQ", heat_source=0.3, bcs=bcs), {}
    if name.startswith("mm_"):
        # several element blocks, one material each (mp_glob[Matilda[ebn]], mm_fill.c:224-235): same equations,
        # every constant different; blocks hold consecutive elements as in an EXODUS II file
        def blocks(m, nb):
            m.elem_block = (np.arange(m.num_elems) * nb // m.num_elems).astype(np.int32)
            return m

        if name == "mm_hex27_bouss_2mat":
            m = blocks(box_mesh("HEX27", (2, 2, 3), perturb=0.12, seed=11), 2)
            return Problem(m, energy=True, rho=1.1, mu=0.05, k=0.07, Cp=1.4, beta=0.8, Tref=0.3, gravity=(0.0, 0.1, -1.0),
                           ns_source="BOUSSINESQ", heat_source=0.6, bcs=_bcs(3, True),
                           extra_materials=[dict(rho=2.3, mu=0.4, k=0.31, Cp=0.9, beta=0.25, Tref=-0.2,
                                                 gravity=(0.3, -0.2, -0.5), ns_source="BOUSS", heat_source=-0.4)]), {}
        if name == "mm_quad9_ale_3mat":
            m = blocks(box_mesh("QUAD9", (6, 3), lo=(0, 0), hi=(2, 1), perturb=0.1, seed=12), 3)
            bcs = _bcs(2) + [Dirichlet("DX", 1, 0.0), Dirichlet("DY", 1, 0.0), Dirichlet("DY", 4, 0.02, relax=1.0)]
            return Problem(m, ale=True, rho=1.3, mu=0.7, gravity=(0.3, -0.2, 0.0), lame_mu=0.9, lame_lambda=1.7, bcs=bcs,
                           extra_materials=[dict(rho=0.6, mu=1.9, gravity=(0.0, -1.0, 0.0), lame_mu=2.5, lame_lambda=0.4),
                                            dict(rho=3.0, mu=0.2, lame_mu=0.3, lame_lambda=3.1)]), {}
        if name == "mm_hex8_pspg_2mat_transient":
            m = blocks(box_mesh("HEX8", (3, 3, 2), perturb=0.12, seed=13), 2)
            bcs = [Dirichlet("U", 1, 1.0), Dirichlet("V", 1, 0.0), Dirichlet("U", 4, 0.0, relax=1.0), Dirichlet("T", 1, 1.0),
                   Dirichlet("Y", 2, 0.3, species=1), Dirichlet("P", 7, 0.0)]
            return Problem(m, interp="Q1Q1", pspg="local", ps_scaling=0.1, energy=True, n_species=2, transient=True,
                           etm_momentum=(1.0, 1.0, 1.0, 1.0, 1.0, 0.0), etm_energy=(1.0,) * 5, etm_species=(1.0,) * 5,
                           rho=1.2, mu=0.3, k=0.2, Cp=1.5, beta=0.4, Tref=0.2, diffusivity=(0.05, 0.11, 1.0, 1.0),
                           gravity=(0.1, -0.2, -1.0), ns_source="BOUSSINESQ", heat_source=0.3, bcs=bcs,
                           extra_materials=[dict(rho=0.4, mu=1.1, k=0.9, Cp=0.6, beta=0.1, Tref=0.0,
                                                 diffusivity=(0.4, 0.02, 1.0, 1.0), gravity=(0.0, 0.0, -0.3),
                                                 heat_source=0.0)]), {"delta_t": 0.02, "theta": 0.5, "time": 0.1}
        if name == "mm_hex27_star_2mat":
            m = blocks(star_mesh("HEX27", refine=0, nz=2, perturb=0.3, seed=14), 2)
            bcs = [Dirichlet("U", 1, 0.0), Dirichlet("V", 1, 0.5, relax=1.0), Dirichlet("W", 5, 0.0), Dirichlet("P", 7, 0.0)]
            return Problem(m, rho=1.1, mu=0.3, gravity=(0.1, -0.2, 0.3), bcs=bcs,
                           extra_materials=[dict(rho=4.0, mu=0.02, gravity=(0.0, 0.0, -2.0))]), {}
    if name == "f3_quad9_free_surface":
        # SURVEY.md §8f-3, reference side only: the C4 strip with a free surface on top (KINEMATIC on the mesh normal,
        # CAPILLARY traction) and a slip wall below (VELO_NORMAL) -- integrated conditions on side sets, applied by the
        # reference's apply_integrated_bc (mm_fill.c:2945-3033, bc_integ.c) incl. the rotation of the equations.  The
        # GPU path does not assemble them yet; the fixture pins WHICH rows they change and what they must become.
        import dataclasses

        p, kw = build_case("c4_quad9_ale")
        return dataclasses.replace(p, bcs=[b for b in p.bcs if b.ns_id != 4],
                                   extra_bc_cards=["BC = KINEMATIC SS 4 0.", "BC = CAPILLARY SS 4 1.0 0.0 0.0",
                                                   "BC = VELO_NORMAL SS 3 0.0"]), kw
    raise KeyError(name)


# fixtures of what the GPU path does not assemble yet (reference output only)
REFERENCE_ONLY_CASES = ["f3_quad9_free_surface"]

GOLDEN_CASES = ["mm_hex27_bouss_2mat", "mm_quad9_ale_3mat", "mm_hex8_pspg_2mat_transient", "mm_hex27_star_2mat",
                "irr_quad9_star_ns", "irr_hex27_star_bouss", "irr_hex8_star_pspg","c1_quad9_ns", "c1_quad9_ns_transient", "c2_hex27_ns", "c3_hex27_boussinesq",
                "c3_quad9_bouss_transient", "c5_hex8_pspg_local_transient", "c5_hex8_pspg_global",
                "c5_quad4_pspg_local", "c4_quad9_ale", "c4_hex27_ale", "c4_quad9_ale_energy_transient",
                "q2p1_quad9_species_ale_transient", "c2_hex27_ns_transient", "c4_hex27_ale_energy_transient",
                "c5_hex8_ns_pspg_local"]


def case_state(name):
    p, kw = build_case(name)
    st = make_state(p, transient=p.transient, delta_t=kw.get("delta_t", 0.01), theta=kw.get("theta", 0.0))
    return p, kw, st
