"""GPU suite: goma_gpu_exchange_dof (exchange_dof, src/dp_comm.c:48-102) and the sub-domain assembly on the CUDA path.

One process per rank as in a real run, but all ranks share cuda:0 (the GPU test box has one device): the peer
pointers come from CUDA IPC handles exactly as between two GPUs, the flag protocol (publish epoch / acquire /
pull) is the same, only the wire is HBM instead of NVLink.  Set-up plumbing (handle exchange) goes over gloo.
"""
import os
import random

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from goma_b200.dp_comm import brick_partition, decompose, scattered_partition, slab_partition
from goma_b200.mesh import box_mesh
from goma_b200.problem import Dirichlet, Problem
from tests.cases import make_state

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]


def _problem(et, n, energy):
    m = box_mesh(et, n, perturb=0.1, seed=5)
    q1 = et in ("HEX8", "QUAD4")
    bcs = [Dirichlet("U", 1, 1.0), Dirichlet("V", 3, 0.0, relax=1.0), Dirichlet("U", 2, 0.0), Dirichlet("P", 7, 0.3)]
    if energy:
        bcs += [Dirichlet("T", 1, 1.0), Dirichlet("T", 2, 0.0, relax=1.0)]
    return Problem(m, interp="Q1Q1" if q1 else "Q2P1", pspg="global" if q1 else None, energy=energy, rho=1.1, mu=0.2,
                   k=0.3, Cp=1.2, beta=0.4, ns_source="BOUSSINESQ" if energy else "CONSTANT", gravity=(0.1, -0.3, 0.2),
                   bcs=bcs)


def _partition(kind, mesh, world):
    if kind == "slab":
        return slab_partition(mesh, world)
    if kind == "brick":
        return brick_partition(mesh, (2, 2, 2) if world == 8 else (2, 2, 1) if mesh.dim == 3 else (2, 2))
    return scattered_partition(mesh, world, seed=3)


def _l2g(sub, first_g):
    first_l = sub.problem.unknown_map()[0]
    out = np.empty(int(first_l[-1]), np.int64)
    for k, g in enumerate(sub.node_global):
        out[first_l[k]:first_l[k + 1]] = np.arange(first_g[g], first_g[g + 1])
    return out


def _worker(rank, world, port_no, et, n, energy, kind, skip_rank, result):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from goma_b200 import capi
        from goma_b200.dp_comm import setup_peer_exchange
        from goma_b200.matrix_fill import MatrixFill, device_view, msr_to_csr
        from oracle import port

        dev = torch.device("cuda", 0)
        p = _problem(et, n, energy)
        st = make_state(p, seed=7)
        first_g = p.unknown_map()[0]
        subs = decompose(p, _partition(kind, p.mesh, world), world)
        sub = subs[rank]
        l2g = _l2g(sub, first_g)
        nown = sub.num_owned_dofs
        mf = MatrixFill(sub.problem, device=0, num_owned_nodes=sub.num_owned_nodes)
        if skip_rank is not None:
            mf.set_option("exchange_timeout_ms", 1500)
        setup_peer_exchange(mf, sub)
        bufs = mf.device_buffers()
        nl = mf.num_unknowns
        out = {"neighbors": len(sub.neighbors)}
        vecs = {0: bufs.d_x, 1: bufs.d_xdot, 2: bufs.d_x_old}
        rng = np.random.default_rng(11)
        for which in (0, 1, 2):
            # owners hold the truth, ghosts hold NaN; the exchange must leave the owners' values in the tail
            xg = st["x"] if which == 0 else rng.normal(size=len(st["x"]))
            xl = xg[l2g].copy()
            xl[nown:] = np.nan
            d = device_view(vecs[which], nl, dev)
            d.copy_(torch.from_numpy(xl))
            torch.cuda.synchronize()
            dist.barrier()
            if skip_rank is None or rank != skip_rank:
                mf.exchange_dof(which)
            if skip_rank is not None:
                if which == 0:
                    try:
                        mf.exchange_status()
                        out["timeout_error"] = False
                    except capi.GomaGpuError as ex:
                        out["timeout_error"] = "never published" in str(ex)
                    break
                continue
            torch.cuda.synchronize()
            out[f"exchange{which}"] = bool(np.array_equal(d.cpu().numpy(), xg[l2g]))
            # nobody overwrites a vector while a neighbour may still be pulling from it: the owner-side fence (instead
            # of a host barrier), then a second round with new values straight away
            mf.exchange_fence(which)
            torch.cuda.synchronize()
            mf.exchange_status()
            xg2 = xg + 1.0 + which
            xl = xg2[l2g].copy()
            xl[nown:] = np.nan
            d.copy_(torch.from_numpy(xl))
            torch.cuda.synchronize()
            mf.exchange_dof(which)
            torch.cuda.synchronize()
            out[f"exchange{which}"] = out[f"exchange{which}"] and bool(np.array_equal(d.cpu().numpy(), xg2[l2g]))
            mf.exchange_fence(which)
            torch.cuda.synchronize()
            if which == 0:  # the fill below assembles the original state
                d.copy_(torch.from_numpy(xg[l2g].copy()))
                torch.cuda.synchronize()
            dist.barrier()
        if skip_rank is None:
            # the fill on the exchanged state: owned rows of all ranks == the global system
            ija_g = capi.pattern_msr(p)
            h, U = (p.global_h_elem_siz(), p.global_velocity_norm(st["x"])) if p.pspg else (0.0, 0.0)
            rc, a_g, r_g = port.port_fill(p, ija_g, st, h_elem_avg=h, U_norm=U)
            ng = len(r_g)
            A_g = msr_to_csr(ija_g, a_g, ng)
            assert mf.fill_device(h_elem_avg=h, U_norm=U) == 0
            a, r = mf.download_system()
            ija = mf.export_msr()
            lrows = np.repeat(np.arange(nl), np.diff(ija[:nl + 1]))
            lcols = ija[nl + 1:]
            vals = a[nl + 1:len(ija)]
            own = lrows < nown
            ref = np.asarray(A_g[l2g[lrows[own]], l2g[lcols[own]]]).ravel()
            scale = np.abs(a_g).max()
            out["offdiag"] = float(np.abs(vals[own] - ref).max() / scale)
            out["ghost_rows_zero"] = bool(not vals[~own].any() and not a[nown:nl].any() and not r[nown:].any())
            out["diag"] = float(np.abs(a[:nown] - a_g[l2g[:nown]]).max() / scale)
            out["resid"] = float(np.abs(r[:nown] - r_g[l2g[:nown]]).max() / np.abs(r_g).max())
            out["rows"] = l2g[:nown].tolist()
        dist.barrier()
        mf.close()
        result[rank] = out
    finally:
        dist.destroy_process_group()


def _spawn(world, et, n, energy, kind, skip_rank=None):
    mgr = mp.Manager()
    result = mgr.dict()
    port_no = 29500 + random.randint(0, 2000)
    mp.spawn(_worker, args=(world, port_no, et, n, energy, kind, skip_rank, result), nprocs=world, join=True)
    return {r: result[r] for r in range(world)}


@pytest.mark.parametrize("world,et,n,energy,kind", [
    (2, "HEX27", (4, 2, 2), True, "slab"),
    (3, "QUAD9", (9, 4), False, "slab"),
    (4, "HEX8", (4, 4, 3), True, "scattered"),
    (8, "HEX27", (4, 4, 4), False, "brick"),
])
def test_gpu_exchange_dof_and_subdomain_assembly(built, world, et, n, energy, kind):
    """exchange_dof over CUDA IPC peer memory fills every ghost dof with its owner's value (x, xdot, x_old), and
    the CUDA fill of every sub-domain, owned rows only (mm_fill.c:5374), reassembles the global system: slabs
    (2 neighbours), a ragged partition, and 2x2x2 bricks where every rank has 7 neighbours
    (dp_map_comm_vec.c:332-461)."""
    res = _spawn(world, et, n, energy, kind)
    p = _problem(et, n, energy)
    ng = int(p.unknown_map()[0][-1])
    seen = np.zeros(ng, int)
    for r in range(world):
        o = res[r]
        assert o["exchange0"] and o["exchange1"] and o["exchange2"], (r, o)
        assert o["ghost_rows_zero"], r
        assert o["offdiag"] < 1e-12 and o["diag"] < 1e-12 and o["resid"] < 1e-12, (r, o)
        seen[np.asarray(o["rows"], int)] += 1
    assert (seen == 1).all()  # every global row is owned, and written, exactly once
    if kind == "brick" and world == 8:
        assert all(res[r]["neighbors"] == 7 for r in range(world))


def test_gpu_exchange_dof_mismatched_collective_times_out(built):
    """A neighbour that never calls exchange_dof: the bounded wait raises an error instead of hanging the GPU."""
    res = _spawn(2, "QUAD9", (6, 4), False, "slab", skip_rank=1)
    assert res[0]["timeout_error"] is True
