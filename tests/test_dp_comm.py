"""Domain decomposition + ghost exchange: host logic on CPU (gloo, world_size 2) -- no GPU needed."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from goma_b200 import capi
from goma_b200.dp_comm import decompose, exchange_dof, slab_partition, slab_subdomain
from goma_b200.mesh import box_mesh
from goma_b200.problem import Dirichlet, Problem
from oracle import port
from tests.cases import make_state


def _problem(et="QUAD9", n=(6, 4), energy=False):
    m = box_mesh(et, n, perturb=0.1, seed=5)
    bcs = [Dirichlet("U", 1, 1.0), Dirichlet("V", 3, 0.0, relax=1.0), Dirichlet("U", 2, 0.0), Dirichlet("P", 7, 0.3)]
    return Problem(m, energy=energy, rho=1.1, mu=0.2, gravity=(0.1, -0.3, 0.2), bcs=bcs)


def _local_to_global_dofs(sub, first_g):
    first_l = sub.problem.unknown_map()[0]
    out = np.empty(int(first_l[-1]), np.int64)
    for k, g in enumerate(sub.node_global):
        out[first_l[k]:first_l[k + 1]] = np.arange(first_g[g], first_g[g + 1])
    return out


@pytest.mark.parametrize("et,n,nranks,nmat", [("QUAD9", (6, 4), 2, 1), ("QUAD9", (9, 3), 3, 1), ("HEX27", (4, 2, 2), 2, 1),
                                              ("QUAD9", (6, 4), 3, 2)])
def test_subdomain_assembly_reproduces_global_rows(built, et, n, nranks, nmat):
    """Each rank assembles all its local elements but loads only owned rows (mm_fill.c:5374); the
    owned rows of all ranks together are exactly the global matrix and residual.  With two materials (element blocks
    cut across the slabs) every local element keeps the material of its block."""
    p = _problem(et, n)
    if nmat > 1:
        import dataclasses

        p.mesh.elem_block = (np.arange(p.mesh.num_elems) % n[0] * nmat // n[0]).astype(np.int32)  # blocks = x ranges
        p = dataclasses.replace(p, extra_materials=[dict(rho=2.7, mu=0.9, gravity=(0.0, -1.0, 0.0))])
    st = make_state(p, seed=3)
    first_g = p.unknown_map()[0]
    ija_g = capi.pattern_msr(p)
    rc, a_g, r_g = port.port_fill(p, ija_g, st)
    ng = len(r_g)
    gmat = {}
    rows = np.repeat(np.arange(ng), np.diff(ija_g[:ng + 1]))
    for k, (r, c) in enumerate(zip(rows, ija_g[ng + 1:])):
        gmat[(int(r), int(c))] = a_g[ng + 1 + k]
    subs = decompose(p, slab_partition(p.mesh, nranks), nranks)
    assert sum(s.num_owned_nodes for s in subs) == p.mesh.num_nodes
    seen_rows = np.zeros(ng, int)
    for sub in subs:
        l2g = _local_to_global_dofs(sub, first_g)
        xl = {"x": st["x"][l2g]}
        ija = capi.pattern_msr(sub.problem)
        rc, a, r = port.port_fill(sub.problem, ija, xl, num_owned_nodes=sub.num_owned_nodes)
        assert rc == 0
        nl = len(r)
        nown = sub.num_owned_dofs
        np.testing.assert_allclose(r[:nown], r_g[l2g[:nown]], rtol=0, atol=1e-13)
        assert not r[nown:].any()
        np.testing.assert_allclose(a[:nown], a_g[l2g[:nown]], rtol=0, atol=1e-13)
        lrows = np.repeat(np.arange(nl), np.diff(ija[:nl + 1]))
        for k, (lr, lc) in enumerate(zip(lrows, ija[nl + 1:])):
            v = a[nl + 1 + k]
            if lr < nown:
                assert abs(v - gmat[(int(l2g[lr]), int(l2g[lc]))]) < 1e-13
            else:
                assert v == 0.0
        seen_rows[l2g[:nown]] += 1
    assert (seen_rows == 1).all()  # every global row is owned exactly once


def _exchange_worker(rank, world, port_no, et, n, energy, result):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        p = _problem(et, n, energy)
        st = make_state(p, seed=7)
        first_g = p.unknown_map()[0]
        sub = decompose(p, slab_partition(p.mesh, world), world)[rank]
        l2g = _local_to_global_dofs(sub, first_g)
        x = torch.from_numpy(st["x"][l2g].copy())
        x[sub.num_owned_dofs:] = float("nan")  # stale ghosts
        exchange_dof(x, sub)
        ok = bool(np.array_equal(x.numpy(), st["x"][l2g]))
        result[rank] = ok
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("et,n,world,energy", [("QUAD9", (6, 4), 2, False), ("HEX27", (4, 2, 2), 2, True)])
def test_exchange_dof_gloo(built, et, n, world, energy):
    """exchange_dof (dp_comm.c:48-102) over gloo: the external tail ends up holding the owners' values."""
    import random

    port_no = 29500 + random.randint(0, 2000)
    mgr = mp.Manager()
    result = mgr.dict()
    mp.spawn(_exchange_worker, args=(world, port_no, et, n, energy, result), nprocs=world, join=True)
    assert all(result[r] for r in range(world))


@pytest.mark.parametrize("et,n,nranks", [("QUAD9", 3, 3), ("HEX27", 2, 2), ("HEX27", 2, 3)])
def test_slab_subdomain_matches_general_decomposition(et, n, nranks):
    """The direct slab constructor used by the weak-scaling bench == decompose() of the global mesh."""
    dim = 2 if et == "QUAD9" else 3
    shape = (nranks * n,) + (n,) * (dim - 1)
    hi = (float(nranks),) + (1.0,) * (dim - 1)
    gm = box_mesh(et, shape, lo=(0.0,) * dim, hi=hi)
    bcs = [Dirichlet("U", 1, 1.0), Dirichlet("V", 2, 0.0, relax=1.0), Dirichlet("U", 3, 0.0), Dirichlet("P", 7, 0.3)]
    mk = lambda mesh: Problem(mesh, rho=1.1, mu=0.2, bcs=bcs)
    subs = decompose(mk(gm), slab_partition(gm, nranks), nranks)
    for r in range(nranks):
        a, b = subs[r], slab_subdomain(mk, n, r, nranks, et)
        assert (a.num_owned_nodes, a.num_internal_nodes, a.neighbors) == (b.num_owned_nodes, b.num_internal_nodes, b.neighbors)
        np.testing.assert_array_equal(a.node_global, b.node_global)
        np.testing.assert_allclose(a.problem.mesh.coords, b.problem.mesh.coords, atol=1e-14)
        # same elements (the element order may differ: compare as sets of node tuples)
        sa = {tuple(row) for row in a.problem.mesh.conn.tolist()}
        sb = {tuple(row) for row in b.problem.mesh.conn.tolist()}
        assert sa == sb
        np.testing.assert_array_equal(a.list_dof_send, b.list_dof_send)
        np.testing.assert_array_equal(a.ptr_dof_send, b.ptr_dof_send)
        np.testing.assert_array_equal(a.num_dofs_recv, b.num_dofs_recv)
        for k in a.problem.mesh.node_sets:
            np.testing.assert_array_equal(a.problem.mesh.node_sets[k], b.problem.mesh.node_sets[k])
        np.testing.assert_array_equal(a.problem.dirichlet_table()[0], b.problem.dirichlet_table()[0])


@pytest.mark.parametrize("et,n,nranks", [("QUAD9", (9, 3), 3), ("HEX27", (4, 2, 2), 2), ("HEX27", (6, 2, 2), 3)])
def test_peer_pull_lists_fill_the_external_tail(et, n, nranks):
    """The lists handed to goma_gpu_exchange_setup: pulling ``x_neighbour[recv_list]`` into the external tail gives
    every ghost dof its owner's value (what exchange_dof, dp_comm.c:48-102, achieves with send/receive)."""
    from goma_b200.dp_comm import peer_exchange_payload, peer_recv_lists

    p = _problem(et, n)
    first_g = p.unknown_map()[0]
    xg = np.random.default_rng(1).normal(size=int(first_g[-1]))
    subs = decompose(p, slab_partition(p.mesh, nranks), nranks)
    everyone = [peer_exchange_payload(s, b"h%d" % s.rank) for s in subs]
    l2g = [_local_to_global_dofs(s, first_g) for s in subs]
    for s in subs:
        handles, slots, recv_ptr, rl = peer_recv_lists(s, everyone)
        assert handles == [b"h%d" % q for q in s.neighbors]
        x = xg[l2g[s.rank]].copy()
        x[s.num_owned_dofs:] = np.nan  # stale ghosts
        for k, q in enumerate(s.neighbors):
            assert subs[q].neighbors[slots[k]] == s.rank
            xq = xg[l2g[q]]
            lst = rl[recv_ptr[k]:recv_ptr[k + 1]]
            assert (lst < subs[q].num_owned_dofs).all()  # only owned values travel
            x[s.num_owned_dofs + recv_ptr[k]: s.num_owned_dofs + recv_ptr[k + 1]] = xq[lst]
        np.testing.assert_array_equal(x, xg[l2g[s.rank]])


@pytest.mark.parametrize("et,n,cols", [("HEX27", 3, [0, 1, 3]), ("HEX8", 5, [0, 2, 3, 5]), ("QUAD9", 7, [0, 2, 4, 5, 7])])
def test_uneven_slab_split_matches_general_decomposition(et, n, cols):
    """Strong scaling: ONE n^dim box cut into x-slabs of unequal width (bench.py --scaling strong) == decompose()."""
    dim = 2 if et == "QUAD9" else 3
    nranks = len(cols) - 1
    gm = box_mesh(et, (n,) * dim)
    q1 = et == "HEX8"
    bcs = [Dirichlet("U", 1, 1.0), Dirichlet("V", 2, 0.0, relax=1.0), Dirichlet("U", 3, 0.0), Dirichlet("P", 7, 0.3)]
    mk = lambda mesh: Problem(mesh, interp="Q1Q1" if q1 else "Q2P1", pspg="global" if q1 else None, rho=1.1, mu=0.2, bcs=bcs)
    ix = np.arange(gm.num_elems) % n
    part = np.searchsorted(np.asarray(cols[1:]), ix, side="right")
    subs = decompose(mk(gm), part, nranks)
    for r in range(nranks):
        a, b = subs[r], slab_subdomain(mk, n, r, nranks, et, cols=cols, x_len=1.0)
        assert (a.num_owned_nodes, a.num_internal_nodes, a.neighbors) == (b.num_owned_nodes, b.num_internal_nodes, b.neighbors)
        np.testing.assert_array_equal(a.node_global, b.node_global)
        np.testing.assert_allclose(a.problem.mesh.coords, b.problem.mesh.coords, atol=1e-14)
        assert {tuple(row) for row in a.problem.mesh.conn.tolist()} == {tuple(row) for row in b.problem.mesh.conn.tolist()}
        np.testing.assert_array_equal(a.list_dof_send, b.list_dof_send)
        np.testing.assert_array_equal(a.ptr_dof_send, b.ptr_dof_send)
        np.testing.assert_array_equal(a.num_dofs_recv, b.num_dofs_recv)
        # (decompose marks an element by the lowest owner of its nodes, rd_dpi.c:322-336; the slab constructor by
        # the column it lies in: both count every element exactly once over the ranks)
        assert int(b.elem_owned.sum()) == (cols[r + 1] - cols[r]) * n ** (dim - 1)
        np.testing.assert_array_equal(a.problem.dirichlet_table()[0], b.problem.dirichlet_table()[0])


@pytest.mark.parametrize("et,n,parts", [("HEX27", 4, (2, 2, 2)), ("HEX8", 5, (2, 2, 2)), ("HEX27", 3, (3, 1, 2)),
                                        ("QUAD9", 6, (2, 3)), ("HEX8", 4, (2, 2, 1))])
def test_brick_subdomain_matches_general_decomposition(et, n, parts):
    """The direct brick constructor of the strong-scaling bench (up to 26 neighbours per rank) == decompose() of the
    global mesh along brick_partition: ownership, local order, send / receive lists, node sets."""
    from goma_b200.dp_comm import brick_partition, brick_subdomain

    dim = 2 if et == "QUAD9" else 3
    nranks = int(np.prod(parts))
    gm = box_mesh(et, (n,) * dim)
    q1 = et == "HEX8"
    bcs = [Dirichlet("U", 1, 1.0), Dirichlet("V", 2, 0.0, relax=1.0), Dirichlet("U", 3, 0.0), Dirichlet("V", 4, 0.5),
           Dirichlet("P", 7, 0.3)] + ([Dirichlet("W", 6, 0.2)] if dim == 3 else [])
    mk = lambda mesh: Problem(mesh, interp="Q1Q1" if q1 else "Q2P1", pspg="global" if q1 else None, rho=1.1, mu=0.2, bcs=bcs)
    subs = decompose(mk(gm), brick_partition(gm, parts), nranks)
    for r in range(nranks):
        a, b = subs[r], brick_subdomain(mk, n, r, parts, et)
        assert (a.num_owned_nodes, a.num_internal_nodes, a.neighbors) == (b.num_owned_nodes, b.num_internal_nodes, b.neighbors)
        np.testing.assert_array_equal(a.node_global, b.node_global)
        np.testing.assert_allclose(a.problem.mesh.coords, b.problem.mesh.coords, atol=1e-14)
        assert {tuple(row) for row in a.problem.mesh.conn.tolist()} == {tuple(row) for row in b.problem.mesh.conn.tolist()}
        np.testing.assert_array_equal(a.list_dof_send, b.list_dof_send)
        np.testing.assert_array_equal(a.ptr_dof_send, b.ptr_dof_send)
        np.testing.assert_array_equal(a.num_dofs_recv, b.num_dofs_recv)
        for k in a.problem.mesh.node_sets:
            np.testing.assert_array_equal(a.problem.mesh.node_sets[k], b.problem.mesh.node_sets[k])
        np.testing.assert_array_equal(a.problem.dirichlet_table()[0], b.problem.dirichlet_table()[0])
    assert sum(int(brick_subdomain(mk, n, r, parts, et).elem_owned.sum()) for r in range(nranks)) == gm.num_elems
