"""Structured synthetic meshes in Exodus II conventions.

The reference reads unstructured Exodus II files (``rd_exo.c:99``); its hot path
only ever sees the resulting arrays: nodal coordinates ``Coor[dim][node]``,
the concatenated connectivity ``Proc_Elem_Connect`` (``rd_mesh.c:476-507``) and
node sets for Dirichlet conditions.  This module generates exactly those arrays
for axis-aligned boxes so that the same mesh can be handed to the reference
(through the oracle driver) and to the GPU path.

Local node order follows the reference's shape functions: QUAD9 / HEX27 in
Exodus-PATRAN order (``rf_shape.c:361-400`` and ``:1105-1200``), HEX8/QUAD4 the
usual counter-clockwise corners (``rf_shape.c:185``, ``:698``).  Global nodes are
numbered x-fastest on the (refined) lattice, elements x-fastest as well.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

# (s,t[,u]) lattice offsets in {0,1,2} of each local node
_QUAD9 = [(0, 0), (2, 0), (2, 2), (0, 2), (1, 0), (2, 1), (1, 2), (0, 1), (1, 1)]
_QUAD4 = [(0, 0), (1, 0), (1, 1), (0, 1)]
_HEX8 = [(0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0), (0, 0, 1), (1, 0, 1), (1, 1, 1), (0, 1, 1)]
_HEX27 = [
    (0, 0, 0), (2, 0, 0), (2, 2, 0), (0, 2, 0), (0, 0, 2), (2, 0, 2), (2, 2, 2), (0, 2, 2),
    (1, 0, 0), (2, 1, 0), (1, 2, 0), (0, 1, 0),
    (0, 0, 1), (2, 0, 1), (2, 2, 1), (0, 2, 1),
    (1, 0, 2), (2, 1, 2), (1, 2, 2), (0, 1, 2),
    (1, 1, 1), (1, 1, 0), (1, 1, 2), (0, 1, 1), (2, 1, 1), (1, 0, 1), (1, 2, 1),
]

ELEM_TABLE = {
    # name: (dim, local lattice offsets, lattice order, exodus type string)
    "QUAD9": (2, _QUAD9, 2, "QUAD9"),
    "QUAD4": (2, _QUAD4, 1, "QUAD4"),
    "HEX27": (3, _HEX27, 2, "HEX27"),
    "HEX8": (3, _HEX8, 1, "HEX8"),
}


@dataclass
class Mesh:
    """Plain arrays, 0-based.  ``conn`` is ``[num_elems, nodes_per_elem]``."""

    elem_type: str
    dim: int
    coords: np.ndarray  # [dim, num_nodes] float64 (the reference's Coor[dim][node])
    conn: np.ndarray  # [num_elems, npe] int32
    node_sets: dict = field(default_factory=dict)  # id -> sorted int32 node list
    shape: tuple = ()  # elements per direction
    lattice: tuple = ()  # nodes per direction

    @property
    def num_nodes(self) -> int:
        return self.coords.shape[1]

    @property
    def num_elems(self) -> int:
        return self.conn.shape[0]

    @property
    def npe(self) -> int:
        return self.conn.shape[1]


def box_mesh(elem_type: str, n, lo=None, hi=None, perturb: float = 0.0, seed: int = 0) -> Mesh:
    """Axis-aligned box of ``n`` elements per direction.

    Node sets (ids as used by the synthetic decks): 1/2 = x min/max, 3/4 = y
    min/max, 5/6 = z min/max, 7 = one pressure-datum node (centroid node of
    element 0 for QUAD9/HEX27, node 0 otherwise).  ``perturb`` > 0 displaces
    interior nodes randomly by that fraction of the lattice spacing (keeps
    detJ > 0 for perturb < 0.25) to exercise non-constant Jacobians.
    """
    dim, offs, order, _ = ELEM_TABLE[elem_type]
    n = tuple(int(v) for v in n)
    assert len(n) == dim
    lo = np.zeros(dim) if lo is None else np.asarray(lo, float)
    hi = np.ones(dim) if hi is None else np.asarray(hi, float)
    lat = tuple(order * k + 1 for k in n)
    axes = [np.linspace(lo[d], hi[d], lat[d]) for d in range(dim)]
    if dim == 2:
        Y, X = np.meshgrid(axes[1], axes[0], indexing="ij")
        coords = np.stack([X.ravel(), Y.ravel()])
    else:
        Z, Y, X = np.meshgrid(axes[2], axes[1], axes[0], indexing="ij")
        coords = np.stack([X.ravel(), Y.ravel(), Z.ravel()])
    strides = np.array([1, lat[0], lat[0] * lat[1]][:dim], dtype=np.int64)

    # element origins on the lattice, x-fastest
    idx = np.indices(n[::-1]).reshape(dim, -1)[::-1]  # [dim, num_elems] (ix, iy, iz)
    origin = (order * idx * strides[:, None]).sum(0)
    off = np.array([sum(o[d] * strides[d] for d in range(dim)) for o in offs], dtype=np.int64)
    conn = (origin[:, None] + off[None, :]).astype(np.int32)

    # lattice index of each node, for node sets and perturbation
    nid = np.arange(coords.shape[1])
    li = [(nid // strides[d]) % lat[d] for d in range(dim)]
    node_sets = {}
    for d in range(dim):
        node_sets[2 * d + 1] = nid[li[d] == 0].astype(np.int32)
        node_sets[2 * d + 2] = nid[li[d] == lat[d] - 1].astype(np.int32)
    datum = conn[0, 8 if elem_type == "QUAD9" else 20] if order == 2 else 0
    node_sets[7] = np.array([datum], dtype=np.int32)

    if perturb > 0.0:
        rng = np.random.default_rng(seed)
        interior = np.ones(coords.shape[1], bool)
        for d in range(dim):
            interior &= (li[d] > 0) & (li[d] < lat[d] - 1)
        for d in range(dim):
            h = (hi[d] - lo[d]) / (lat[d] - 1)
            coords[d, interior] += perturb * h * rng.uniform(-1, 1, interior.sum())
    return Mesh(elem_type, dim, np.ascontiguousarray(coords), conn, node_sets, n, lat)
