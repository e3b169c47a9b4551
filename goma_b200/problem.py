"""Host-side problem description for the matrix_fill hot path.

In a real run the Goma host owns all of this state in process globals
(``pd_glob``, ``mp_glob``, ``upd``, ``Nodes[]``, ``BC_Types[]`` ... SURVEY.md
App. C) filled from its input deck; the GPU entry only receives a plain-C
snapshot of it (``include/goma_gpu_fill.h``: ``struct goma_gpu_problem``).
This module is the Python mirror of that snapshot for the synthetic configs of
BASELINE.json.  It restates, for the in-scope physics only, the reference's

* variable ids (``include/rf_fem_const.h:174-200``),
* unknown numbering: node-major, variables in increasing id inside a node,
  MASS_FRACTION expanded to ``ns`` entries, P1 pressure = dim+1 dofs on the
  element centroid node (``mm_unknown_map.c:758-971``, ``el_elm_info.c:925-940``),
* interaction mask rows (``mm_unknown_map.c:1226-1290,1692-,2276-``),
* Dirichlet table semantics (``bc_dirich.c:44-151``, ``mm_bc.c`` find_and_set_Dirichlet),

and can also print itself as a Goma input deck + ``.mat`` file so the very same
problem can be run through the reference (the oracle) for parity.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional

import numpy as np

from .mesh import Mesh

# include/rf_fem_const.h:174-200
VELOCITY1, VELOCITY2, VELOCITY3, TEMPERATURE, MASS_FRACTION = 0, 1, 2, 3, 4
MESH_DISPLACEMENT1, MESH_DISPLACEMENT2, MESH_DISPLACEMENT3 = 5, 6, 7
PRESSURE = 9
MAX_CONC = 4

# field slots used on the device: one "slot" per scalar unknown family
# (species expanded), in the order they appear inside a node.
SLOT_NAMES = ["U", "V", "W", "T", "Y0", "Y1", "Y2", "Y3", "DX", "DY", "DZ", "P"]


@dataclass
class Dirichlet:
    """``BC = {U|V|W|T|P|DX|DY|DZ} NS <id> <value> [relax]`` or ``BC = Y NS <id> <w> <value>``.

    ``relax is None`` is the reference's ``BC_relax == -1`` hard-set case: ``x`` is
    preset by find_and_set_Dirichlet and the row residual is 0
    (``bc_dirich.c:134-135``); otherwise the residual is ``x - value`` (``:137-139``).
    """

    var: str
    ns_id: int
    value: float
    relax: Optional[float] = None
    species: int = 0


@dataclass
class Problem:
    mesh: Mesh
    interp: str = "Q2P1"  # "Q2P1" (QUAD9/HEX27) or "Q1Q1" (QUAD4/HEX8, needs PSPG)
    energy: bool = False
    n_species: int = 0
    ale: bool = False  # pseudo-solid ARBITRARY mesh motion (mesh1..dim equations)
    # material (CONSTANT models only)
    rho: float = 1.0
    mu: float = 1.0
    k: float = 1.0
    Cp: float = 1.0
    beta: float = 1.0  # Volume Expansion
    Tref: float = 0.0
    diffusivity: tuple = (1.0, 1.0, 1.0, 1.0)
    gravity: tuple = (0.0, 0.0, 0.0)  # Navier-Stokes Source vector
    ns_source: str = "CONSTANT"  # or "BOUSS"
    heat_source: float = 0.0
    lame_mu: float = 1.0
    lame_lambda: float = 1.0
    # equation term multipliers, reference order of the EQ card
    etm_momentum: tuple = (0.0, 1.0, 1.0, 1.0, 1.0, 0.0)  # mass adv bnd diff src porous
    etm_continuity: tuple = (1.0, 0.0)  # div(adv) src
    etm_energy: tuple = (0.0, 1.0, 1.0, 1.0, 1.0)  # mass adv bnd diff src
    etm_species: tuple = (0.0, 1.0, 1.0, 1.0, 1.0)
    etm_mesh: tuple = (0.0, 0.0, 1.0, 1.0, 0.0)  # mass adv bnd diff src
    transient: bool = False
    pspg: Optional[str] = None  # None | "global" | "local"
    ps_scaling: float = 0.1
    bcs: list = field(default_factory=list)
    # materials 1.. of a multi-block mesh (``mesh.elem_block``): dicts overriding the constants above (rho, mu, k, Cp,
    # beta, Tref, diffusivity, gravity, ns_source, heat_source, lame_mu, lame_lambda); material 0 = the fields above.
    # Same equations in every material (mp_glob[Matilda[ebn]] picked per element block, mm_fill.c:224-235).
    extra_materials: list = field(default_factory=list)
    # raw ``BC = ...`` cards of conditions the GPU path does not assemble (integrated conditions on side sets):
    # written into the deck for the reference driver of the oracle; the C-ABI marshalling refuses such a problem
    extra_bc_cards: list = field(default_factory=list)

    MATERIAL_KEYS = ("rho", "mu", "k", "Cp", "beta", "Tref", "diffusivity", "gravity", "ns_source", "heat_source",
                     "lame_mu", "lame_lambda")

    @property
    def num_materials(self) -> int:
        return 1 + len(self.extra_materials)

    def material(self, m: int) -> dict:
        """Constants of material ``m`` (0 = the Problem's own fields)."""
        d = {k: getattr(self, k) for k in self.MATERIAL_KEYS}
        if m > 0:
            unknown = set(self.extra_materials[m - 1]) - set(self.MATERIAL_KEYS)
            if unknown:
                raise ValueError(f"unknown material constants {sorted(unknown)}")
            d.update(self.extra_materials[m - 1])
        return d

    def single_material(self, m: int):
        """The same problem with material ``m``'s constants everywhere (no extra materials)."""
        import dataclasses

        mesh = dataclasses.replace(self.mesh, elem_block=None)
        return dataclasses.replace(self, mesh=mesh, extra_materials=[], **self.material(m))

    # ------------------------------------------------------------------ layout
    @property
    def dim(self) -> int:
        return self.mesh.dim

    def node_slots(self):
        """Slot names carried by every node that has the main (phi) interpolation."""
        s = ["U", "V", "W"][: self.dim]
        if self.energy:
            s.append("T")
        s += [f"Y{w}" for w in range(self.n_species)]
        if self.ale:
            s += ["DX", "DY", "DZ"][: self.dim]
        return s

    @property
    def n_pressure_dofs(self) -> int:
        return self.dim + 1 if self.interp == "Q2P1" else 1

    def unknown_map(self):
        """Returns (first_unknown[num_nodes+1], node_kind[num_nodes], kinds).

        ``kinds`` is a list of slot-name lists; a node of kind k carries
        ``len(kinds[k])`` consecutive unknowns starting at first_unknown[node].
        P1 pressure appears as ["P", "P", ...] (dim+1 entries) on centroid nodes.
        """
        m = self.mesh
        base = self.node_slots()
        if self.interp == "Q2P1":
            assert m.elem_type in ("QUAD9", "HEX27")
            cen = 8 if m.elem_type == "QUAD9" else 20
            kinds = [base, base + ["P"] * (self.dim + 1)]
            node_kind = np.zeros(m.num_nodes, np.int32)
            node_kind[m.conn[:, cen]] = 1
        elif self.interp == "Q1Q1":
            assert m.elem_type in ("QUAD4", "HEX8")
            kinds = [base + ["P"]]
            node_kind = np.zeros(m.num_nodes, np.int32)
        else:
            raise ValueError(self.interp)
        nunk = np.array([len(k) for k in kinds], np.int64)[node_kind]
        first = np.zeros(m.num_nodes + 1, np.int64)
        np.cumsum(nunk, out=first[1:])
        return first, node_kind, kinds

    @staticmethod
    def slot_var(slot: str) -> int:
        return {"U": 0, "V": 1, "W": 2, "T": 3, "DX": 5, "DY": 6, "DZ": 7, "P": 9}.get(slot, MASS_FRACTION)

    @staticmethod
    def inter_mask(row_var: int, col_var: int) -> bool:
        """Inter_Mask restricted to {v,T,Y,d,P}: only energy rows skip P
        (``mm_unknown_map.c`` R_ENERGY case has no PRESSURE entry)."""
        return not (row_var == TEMPERATURE and col_var == PRESSURE)

    # --------------------------------------------------------------- Dirichlet
    def dirichlet_table(self):
        """Per-unknown Dirichlet data: (is_dbc[n], value[n], hard[n]).

        Restates ``Nodes[]->DBC`` + ``BC_Types[].BC_Data_Float[0]`` + ``BC_relax``
        as consumed by ``put_dirichlet_in_matrix`` (``bc_dirich.c:86-140``).  A later
        BC card on the same unknown overrides an earlier one.
        """
        first, node_kind, kinds = self.unknown_map()
        n = int(first[-1])
        is_dbc = np.zeros(n, np.int32)
        value = np.zeros(n, np.float64)
        hard = np.zeros(n, np.int32)
        for bc in self.bcs:
            slot = f"Y{bc.species}" if bc.var == "Y" else bc.var
            nodes = self.mesh.node_sets[bc.ns_id]
            for kind_id, slots in enumerate(kinds):
                if slot not in slots:
                    continue
                off = slots.index(slot)  # first dof of that variable on the node
                sel = nodes[node_kind[nodes] == kind_id]
                idx = first[sel] + off
                is_dbc[idx] = 1
                value[idx] = bc.value
                hard[idx] = 1 if bc.relax is None else 0
        return is_dbc, value, hard

    def preset_dirichlet(self, x, xdot=None):
        """find_and_set_Dirichlet (``mm_bc.c:480-585``): every hard-set card (no relax float)
        overwrites ``x`` with its value and zeroes ``xdot``, card by card in deck order -- so at a
        node shared by two node sets the value of the last *hard* card stays in ``x`` even when
        a later relaxed card owns the row."""
        first, node_kind, kinds = self.unknown_map()
        for bc in self.bcs:
            if bc.relax is not None:
                continue
            slot = f"Y{bc.species}" if bc.var == "Y" else bc.var
            nodes = self.mesh.node_sets[bc.ns_id]
            for kind_id, slots in enumerate(kinds):
                if slot in slots:
                    sel = nodes[node_kind[nodes] == kind_id]
                    idx = first[sel] + slots.index(slot)
                    x[idx] = bc.value
                    if xdot is not None:
                        xdot[idx] = 0.0
        return x

    # ------------------------------------------------- PSPG global norms (host side)
    def global_h_elem_siz(self, elem_owned=None) -> float:
        """global_h_elem_siz (``mm_fill_aux.c:1128-1207``): mean over the (owned) elements of
        sqrt(sum_p hsquared[p] / dim), hsquared from face-centroid differences (h_elem_siz, ``:844``).
        In a distributed run this is the local sum; the host all-reduces (SUM) and divides by the
        global element count."""
        m = self.mesh
        X = m.coords[:, m.conn]  # [dim, ne, npe]
        if m.dim == 2:
            h0 = 0.5 * (X[:, :, 1] + X[:, :, 2]) - 0.5 * (X[:, :, 0] + X[:, :, 3])
            h1 = 0.5 * (X[:, :, 0] + X[:, :, 1]) - 0.5 * (X[:, :, 2] + X[:, :, 3])
            hsq = (h0 ** 2).sum(0) + (h1 ** 2).sum(0)
        else:
            f = lambda a, b, c, d: 0.25 * (X[:, :, a] + X[:, :, b] + X[:, :, c] + X[:, :, d])
            p1, p2, p3 = f(0, 1, 2, 3), f(1, 2, 5, 6), f(2, 3, 6, 7)
            p4, p5, p6 = f(0, 1, 4, 5), f(0, 3, 4, 7), f(4, 5, 6, 7)
            hsq = ((p2 - p5) ** 2).sum(0) + ((p3 - p4) ** 2).sum(0) + ((p1 - p6) ** 2).sum(0)
        h = np.sqrt(hsq / m.dim)
        if elem_owned is not None:
            h = h[elem_owned]
        return float(h.sum() / m.num_elems) if elem_owned is None else float(h.sum())

    def global_velocity_norm(self, x, num_owned_nodes=None) -> float:
        """global_velocity_norm (``mm_fill_aux.c:612-680``): mean of the squared velocity unknowns over
        the owned nodes (note: no square root in the reference)."""
        first, node_kind, kinds = self.unknown_map()
        nown = self.mesh.num_nodes if num_owned_nodes is None else num_owned_nodes
        tot, cnt = 0.0, 0
        for kind_id, slots in enumerate(kinds):
            nodes = np.nonzero(node_kind[:nown] == kind_id)[0]
            for name in ("U", "V", "W")[: self.dim]:
                v = x[first[nodes] + slots.index(name)]
                tot += float((v * v).sum())
                cnt += len(v)
        return tot / cnt

    # ------------------------------------------------------------------- decks
    def deck(self) -> str:
        """Goma problem-description file for this problem (cards per SURVEY.md App. C)."""
        q = "Q2" if self.interp == "Q2P1" else "Q1"
        pq = "P1" if self.interp == "Q2P1" else "Q1"
        L = ["FEM File Specifications", "FEM file = mesh.exoII", "Output EXODUS II file = out.exoII",
             "GUESS file = contin.dat", "SOLN file = soln.dat", "Write intermediate results = no", "",
             "General Specifications", "Output Level = 0", "Debug = 0", "Initial Guess = zero", "",
             "Time Integration Specifications"]
        if self.transient:
            L += ["Time integration = transient", "delta_t = 0.01", "Maximum number of time steps = 1",
                  "Maximum time = 1.0", "Minimum time step = 1e-9", "Time step parameter = 0.0",
                  "Time step error = 0.01 0 1 1 1 1 1 1", "Printing Frequency = 1"]
        else:
            L += ["Time integration = steady"]
        L += ["", "Solver Specifications", "Solution Algorithm = lu", "Matrix storage format = msr",
              "Number of Newton Iterations = 10", "Newton correction factor = 1",
              "Normalized Residual Tolerance = 1e-10"]
        if self.pspg:
            L += [f"Pressure Stabilization = {'yes' if self.pspg == 'global' else 'local'}",
                  f"Pressure Stabilization Scaling = {self.ps_scaling!r}"]
        L += ["", "Boundary Condition Specifications", f"Number of BC = {len(self.bcs) + len(self.extra_bc_cards)}"]
        for bc in self.bcs:
            relax = "" if bc.relax is None else f" {bc.relax!r}"
            if bc.var == "Y":
                L.append(f"BC = Y NS {bc.ns_id} {bc.species} {bc.value!r}{relax}")
            else:
                L.append(f"BC = {bc.var} NS {bc.ns_id} {bc.value!r}{relax}")
        L += list(self.extra_bc_cards)
        L += ["END OF BC", "", "Problem Description", f"Number of Materials = {self.num_materials}"]
        eqs = []
        fm = lambda t: " ".join(repr(float(v)) for v in t)
        if self.ale:
            for a in range(self.dim):
                eqs.append(f"EQ = mesh{a+1} {q} D{a+1} {q} {fm(self.etm_mesh)}")
        for a in range(self.dim):
            eqs.append(f"EQ = momentum{a+1} {q} U{a+1} {q} {fm(self.etm_momentum)}")
        if self.energy:
            eqs.append(f"EQ = energy {q} T {q} {fm(self.etm_energy)}")
        if self.n_species:
            eqs.append(f"EQ = species_bulk {q} Y {q} {fm(self.etm_species)}")
        eqs.append(f"EQ = continuity {pq} P {pq} {fm(self.etm_continuity)}")
        for m in range(self.num_materials):  # one MAT section per element block, the same equations in each
            L += [f"MAT = {self.mat_name(m)} {m + 1}", "Coordinate System = CARTESIAN", "Element Mapping = isoparametric",
                  "Mesh Motion = ARBITRARY", f"Number of bulk species = {self.n_species}",
                  f"Number of EQ = {len(eqs)}"] + eqs + ["END OF EQ"]
        L += ["END OF MAT", "", "Post Processing Specifications", "Stream Function = no", ""]
        return "\n".join(L)

    def mat_name(self, m: int) -> str:
        return "fluid" if m == 0 else f"fluid{m + 1}"

    def mat_file(self, m: int = 0) -> str:
        if m > 0 or self.extra_materials:
            return self.single_material(m).mat_file()
        g = self.gravity
        L = ["---Physical Properties", f"Density = CONSTANT {self.rho!r}",
             "---Mechanical Properties and Constitutive Equations",
             "Solid Constitutive Equation = NONLINEAR", "Convective Lagrangian Velocity = NONE",
             f"Lame MU = CONSTANT {self.lame_mu!r}", f"Lame LAMBDA = CONSTANT {self.lame_lambda!r}",
             "Stress Free Solvent Vol Frac = CONSTANT 0.", "Liquid Constitutive Equation = NEWTONIAN",
             f"Viscosity = CONSTANT {self.mu!r}", "Momentum Weight Function = GALERKIN",
             "Polymer Constitutive Equation = NOPOLYMER", "---Thermal Properties",
             "Heat Flux Model = USER" if False else "",
             f"Conductivity = CONSTANT {self.k!r}", f"Heat Capacity = CONSTANT {self.Cp!r}",
             f"Volume Expansion = CONSTANT {self.beta!r}", f"Reference Temperature = CONSTANT {self.Tref!r}",
             "Liquidus Temperature = CONSTANT 1.", "Solidus Temperature = CONSTANT 1.",
             "Energy Weight Function = GALERKIN",
             "---Electrical Properties", "Electrical Conductivity = CONSTANT 1.",
             "---Microstructure Properties", "Media Type = CONTINUOUS", "---Species Properties"]
        if self.n_species:
            L += ["Diffusion Constitutive Equation = FICKIAN", "Species Weight Function = GALERKIN"]
            for w in range(self.n_species):
                L += [f"Diffusivity = CONSTANT {w} {self.diffusivity[w]!r}",
                      f"Latent Heat Vaporization = CONSTANT {w} 0.",
                      f"Latent Heat Fusion = CONSTANT {w} 0.",
                      f"Vapor Pressure = CONSTANT {w} 0.",
                      f"Species Volume Expansion = CONSTANT {w} 0.",
                      f"Reference Concentration = CONSTANT {w} 0."]
        else:
            L += ["Diffusion Constitutive Equation = NONE"]
        L += ["---Source Terms",
              f"Navier-Stokes Source = {self.ns_source} {g[0]!r} {g[1]!r} {g[2]!r}",
              "Solid Body Source = CONSTANT 0. 0. 0.", "Mass Source = CONSTANT 0.",
              f"Heat Source = CONSTANT {self.heat_source!r}"]
        for w in range(max(self.n_species, 1)):
            L += [f"Species Source = CONSTANT {w} 0."]
        L += ["Current Source = CONSTANT 0.", ""]
        return "\n".join(l for l in L if l is not None)
