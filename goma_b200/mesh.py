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

from typing import Optional

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
    # element block of each element (exo->eb: blocks hold consecutive elements, so the array is non-decreasing);
    # None = one block.  Block b carries material b of the Problem (Matilda[ebn] == ebn in the decks written here).
    elem_block: Optional[np.ndarray] = None
    # side sets (exo->ss_*): id -> (elements [k] 0-based, EXODUS II side numbers [k] 1-based); the reference's integrated
    # boundary conditions (BC = ... SS <id>) hang off them.  Only the reference driver of the oracle consumes them so far.
    side_sets: dict = field(default_factory=dict)

    @property
    def num_elem_blocks(self) -> int:
        return 1 if self.elem_block is None else int(self.elem_block.max()) + 1

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
    # side sets with the ids of the node sets 1..6: the element sides on that face of the box (EXODUS II side numbers of
    # QUAD: 1 bottom, 2 right, 3 top, 4 left; HEX: 1 y-min, 2 x-max, 3 y-max, 4 x-min, 5 z-min, 6 z-max)
    side_of = {2: {1: 4, 2: 2, 3: 1, 4: 3}, 3: {1: 4, 2: 2, 3: 1, 4: 3, 5: 5, 6: 6}}[dim]
    side_sets = {}
    eid = np.arange(conn.shape[0])
    for d in range(dim):
        for hi_side in (0, 1):
            sel = eid[idx[d] == (n[d] - 1 if hi_side else 0)]
            side_sets[2 * d + 1 + hi_side] = (sel.astype(np.int32), np.full(len(sel), side_of[2 * d + 1 + hi_side], np.int32))
    m = Mesh(elem_type, dim, np.ascontiguousarray(coords), conn, node_sets, n, lat)
    m.side_sets = side_sets
    return m


def patch_mesh(elem_type: str, verts, quads, refine: int = 0, nz: int = 0, height: float = 1.0) -> Mesh:
    """Unstructured mesh from a quad4 topology: ``verts`` [nv, 2], ``quads`` [nq, 4] counter-clockwise.

    What the reference reads from an arbitrary Exodus file (``rd_exo.c:99``) rather than from a lattice: vertex
    valences other than 4 (three or five quads round a point), so node-node lists of irregular length and an
    element colouring that a lattice formula cannot give.  ``refine`` splits every quad into 2x2 that many
    times; QUAD9 adds the unique edge-midpoint and centre nodes; ``nz`` > 0 extrudes the plane mesh into
    ``nz`` layers of HEX8 / HEX27 elements (Exodus node order, ``rf_shape.c:698,1105``).
    Node sets: 1 = boundary of the plane mesh (all layers), 5 / 6 = bottom / top layer (3-D), 7 = pressure datum.
    """
    verts = [tuple(map(float, v)) for v in verts]
    quads = [tuple(int(k) for k in q) for q in quads]
    for _ in range(refine):
        mids, nq = {}, []

        def mid(a, b):
            key = (min(a, b), max(a, b))
            if key not in mids:
                verts.append(tuple(0.5 * (verts[a][d] + verts[b][d]) for d in range(2)))
                mids[key] = len(verts) - 1
            return mids[key]

        for q in quads:
            m = [mid(q[k], q[(k + 1) % 4]) for k in range(4)]
            verts.append(tuple(0.25 * sum(verts[v][d] for v in q) for d in range(2)))
            c = len(verts) - 1
            nq += [(q[0], m[0], c, m[3]), (m[0], q[1], m[1], c), (c, m[1], q[2], m[2]), (m[3], c, m[2], q[3])]
        quads = nq
    order = 2 if elem_type in ("QUAD9", "HEX27") else 1
    pts = list(verts)
    conn2 = []
    edge_count = {}
    for q in quads:
        for k in range(4):
            key = (min(q[k], q[(k + 1) % 4]), max(q[k], q[(k + 1) % 4]))
            edge_count[key] = edge_count.get(key, 0) + 1
    if order == 1:
        conn2 = [list(q) for q in quads]
        emid = {}
    else:
        emid = {}
        for q in quads:
            row = list(q)
            for k in range(4):
                a, b = q[k], q[(k + 1) % 4]
                key = (min(a, b), max(a, b))
                if key not in emid:
                    pts.append(tuple(0.5 * (verts[a][d] + verts[b][d]) for d in range(2)))
                    emid[key] = len(pts) - 1
                row.append(emid[key])
            pts.append(tuple(0.25 * sum(verts[v][d] for v in q) for d in range(2)))
            row.append(len(pts) - 1)
            conn2.append(row)
    bnd = set()
    for (a, b), cnt in edge_count.items():
        if cnt == 1:
            bnd.update((a, b))
            if (a, b) in emid:
                bnd.add(emid[(a, b)])
    pts = np.array(pts, float)
    conn2 = np.array(conn2, np.int32)
    n2 = len(pts)
    if nz == 0:
        et = "QUAD9" if order == 2 else "QUAD4"
        assert elem_type == et
        sets = {1: np.array(sorted(bnd), np.int32), 7: np.array([conn2[0, 8] if order == 2 else 0], np.int32)}
        return Mesh(et, 2, np.ascontiguousarray(pts.T), conn2, sets, (len(conn2),), (n2,))
    et = "HEX27" if order == 2 else "HEX8"
    assert elem_type == et
    nlay = order * nz + 1
    zs = np.linspace(0.0, height, nlay)
    coords = np.stack([np.tile(pts[:, 0], nlay), np.tile(pts[:, 1], nlay), np.repeat(zs, n2)])
    q_of = {o: k for k, o in enumerate(_QUAD9 if order == 2 else _QUAD4)}
    offs = _HEX27 if order == 2 else _HEX8
    conn = np.empty((len(conn2) * nz, len(offs)), np.int32)
    for lz in range(nz):
        for k, (s, t, u) in enumerate(offs):
            conn[lz * len(conn2):(lz + 1) * len(conn2), k] = conn2[:, q_of[(s, t)]] + (order * lz + u) * n2
    b = np.array(sorted(bnd), np.int64)
    sets = {1: np.sort((b[None, :] + n2 * np.arange(nlay)[:, None]).ravel()).astype(np.int32),
            5: np.arange(n2, dtype=np.int32), 6: (np.arange(n2) + n2 * (nlay - 1)).astype(np.int32),
            7: np.array([conn[0, 20] if order == 2 else 0], np.int32)}
    for k in (2, 3, 4):
        sets[k] = np.zeros(0, np.int32)
    return Mesh(et, 3, np.ascontiguousarray(coords), conn, sets, (len(conn),), (coords.shape[1],))


def star_mesh(elem_type: str, refine: int = 1, nz: int = 0, seed: int = 0, perturb: float = 0.0) -> Mesh:
    """Two fans of quads sharing an edge region: five quads round one interior vertex (valence 5) and three round
    another (valence 3) -- the smallest plane mesh with both irregular valences."""
    import math

    verts, quads = [(0.0, 0.0)], []
    # valence-5 fan round vertex 0: corners v_k on the unit circle, mid-edge points m_k between them
    ring = []
    for k in range(5):
        a0, a1 = 2 * math.pi * k / 5, 2 * math.pi * (k + 0.5) / 5
        verts.append((math.cos(a0), math.sin(a0)))
        verts.append((0.9 * math.cos(a1), 0.9 * math.sin(a1)))
        ring += [len(verts) - 2, len(verts) - 1]
    for k in range(5):
        v, m_prev, m = ring[2 * k], ring[(2 * k - 1) % 10], ring[2 * k + 1]
        quads.append((0, m_prev, v, m))
    # valence-3 fan attached on the right: centre c3, sharing the edge (v_0, m_0)... built as a separate hexagon
    cx = 2.2
    verts.append((cx, 0.0))
    c3 = len(verts) - 1
    hexr = []
    for k in range(6):
        a = 2 * math.pi * k / 6 + math.pi
        verts.append((cx + 1.0 * math.cos(a), math.sin(a)))
        hexr.append(len(verts) - 1)
    for k in range(3):
        quads.append((c3, hexr[(2 * k - 1) % 6], hexr[2 * k], hexr[(2 * k + 1) % 6]))
    # bridge quad between the fans: v_0 = ring[0] at (1, 0) and hexagon vertex hexr[0] at (cx - 1, 0)
    m0, m9 = ring[1], ring[9]
    quads.append((ring[0], m9, hexr[1], hexr[0]))
    quads.append((ring[0], hexr[0], hexr[5], m0))
    m = patch_mesh(elem_type, verts, quads, refine=refine, nz=nz)
    if perturb > 0.0:
        rng = np.random.default_rng(seed)
        free = np.ones(m.num_nodes, bool)
        free[m.node_sets[1]] = False
        m.coords[:2, free] += perturb * 0.1 * rng.uniform(-1, 1, (2, int(free.sum())))
    return m
