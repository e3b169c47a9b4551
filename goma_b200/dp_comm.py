"""Domain decomposition and the ghost-dof exchange of the assembly path.

Mirrors what the reference builds for a distributed run and uses around every fill:

* element partition -> node owners (an element's nodes go to the lowest rank touching them;
  ``rd_dpi.c:322-336``), one layer of ghost elements so that every owned row is complete
  (``dp_ghost.cpp:71``), local node order = internal, boundary, external
  (``mm_unknown_map.c:853-891``) with the external nodes contiguous per owner
  (``dp_map_comm_vec.c:472-483``);
* the send lists ``list_dof_send`` / ``ptr_dof_send`` (``dp_map_comm_vec.c:332-422``) and
* ``exchange_dof(cx, dpi, x, imtrx)`` (``dp_comm.c:48-102``): gather the send unknowns, post one
  send and one receive per neighbour, receive straight into the contiguous external tail of ``x``.

The transport is ``torch.distributed`` point-to-point: NCCL over NVLink on device tensors in a
multi-GPU run (the gather runs in ``goma_gpu_pack_dofs``), gloo on CPU tensors in the tests.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from .mesh import Mesh
from .problem import Problem


@dataclass
class Subdomain:
    rank: int
    problem: Problem  # local problem: mesh in local numbering, same physics / BC cards
    num_owned_nodes: int  # internal + boundary
    num_internal_nodes: int
    node_global: np.ndarray  # local node -> global node
    elem_global: np.ndarray  # local element -> global element
    elem_owned: np.ndarray  # bool: element's owner rank (min of node owners) == rank
    neighbors: list = field(default_factory=list)  # neighbour ranks, ascending
    # dof-level lists (reference: list_dof_send / ptr_dof_send, cx[p].num_dofs_recv)
    list_dof_send: np.ndarray = None
    ptr_dof_send: np.ndarray = None
    num_dofs_recv: np.ndarray = None

    @property
    def num_unknowns(self) -> int:
        return int(self.problem.unknown_map()[0][-1])

    @property
    def num_owned_dofs(self) -> int:
        return int(self.problem.unknown_map()[0][self.num_owned_nodes])


def _sub_problem(problem: Problem, mesh: Mesh) -> Problem:
    import dataclasses

    return dataclasses.replace(problem, mesh=mesh)


def decompose(problem: Problem, elem_rank: np.ndarray, nranks: int):
    """Split ``problem`` along the element partition ``elem_rank`` into ``nranks`` sub-domains."""
    m = problem.mesh
    elem_rank = np.asarray(elem_rank, np.int64)
    nn, ne, npe = m.num_nodes, m.num_elems, m.npe
    node_owner = np.full(nn, nranks, np.int64)
    np.minimum.at(node_owner, m.conn.ravel(), np.repeat(elem_rank, npe))
    first_g, kind_g, kinds = problem.unknown_map()
    ndof_node = np.diff(first_g)

    # which ranks see each node: a rank sees node n iff it owns a node of an element containing n
    owner_of_conn = node_owner[m.conn]  # [ne, npe]
    subs = []
    # per rank: local elements = elements with at least one owned node
    local_elems = [np.nonzero((owner_of_conn == r).any(axis=1))[0] for r in range(nranks)]
    local_nodes = [np.unique(m.conn[le]) for le in local_elems]
    # a node is "boundary" on its owner if some other rank sees it
    seen_by_other = np.zeros(nn, bool)
    for r in range(nranks):
        ln = local_nodes[r]
        seen_by_other[ln[node_owner[ln] != r]] = True

    for r in range(nranks):
        ln = local_nodes[r]
        own = node_owner[ln] == r
        internal = ln[own & ~seen_by_other[ln]]
        boundary = ln[own & seen_by_other[ln]]
        ext = ln[~own]
        ext = ext[np.lexsort((ext, node_owner[ext]))]  # by owner, then global id
        order = np.concatenate([internal, boundary, ext])
        g2l = np.full(nn, -1, np.int64)
        g2l[order] = np.arange(len(order))
        le = local_elems[r]
        conn = g2l[m.conn[le]].astype(np.int32)
        node_sets = {}
        for k, nodes in m.node_sets.items():
            loc = g2l[nodes]
            node_sets[k] = np.sort(loc[loc >= 0]).astype(np.int32)
        lmesh = Mesh(m.elem_type, m.dim, np.ascontiguousarray(m.coords[:, order]), conn, node_sets, m.shape, m.lattice)
        if m.elem_block is not None:  # the material of an element travels with it (mp_glob[Matilda[ebn]])
            lmesh.elem_block = np.ascontiguousarray(m.elem_block[le])
        sub = Subdomain(rank=r, problem=_sub_problem(problem, lmesh), num_owned_nodes=len(internal) + len(boundary),
                        num_internal_nodes=len(internal), node_global=order, elem_global=le,
                        elem_owned=owner_of_conn[le].min(axis=1) == r)
        subs.append(sub)

    # communication lists: rank r receives its external nodes owned by p (sorted by global id);
    # rank p sends exactly those nodes, in the same order
    for r, sub in enumerate(subs):
        first_l = sub.problem.unknown_map()[0]
        ext = sub.node_global[sub.num_owned_nodes:]
        owners = node_owner[ext]
        nbr_recv = set(np.unique(owners).tolist())
        nbr_send = {q for q in range(nranks) if q != r and
                    (node_owner[subs[q].node_global[subs[q].num_owned_nodes:]] == r).any()}
        sub.neighbors = sorted(nbr_recv | nbr_send)
        send_lists, recv_counts = [], []
        for p in sub.neighbors:
            want = subs[p].node_global[subs[p].num_owned_nodes:]
            want = want[node_owner[want] == r]  # already sorted by global id inside one owner
            g2l = {int(g): k for k, g in enumerate(sub.node_global[:sub.num_owned_nodes])}
            loc = np.array([g2l[int(g)] for g in want], np.int64)
            dofs = [np.arange(first_l[k], first_l[k + 1]) for k in loc]
            send_lists.append(np.concatenate(dofs).astype(np.int32) if dofs else np.zeros(0, np.int32))
            recv_counts.append(int(ndof_node[ext[owners == p]].sum()))
        sub.list_dof_send = np.concatenate(send_lists).astype(np.int32) if send_lists else np.zeros(0, np.int32)
        sub.ptr_dof_send = np.concatenate([[0], np.cumsum([len(s) for s in send_lists])]).astype(np.int64)
        sub.num_dofs_recv = np.array(recv_counts, np.int64)
    return subs


def slab_partition(mesh: Mesh, nranks: int) -> np.ndarray:
    """Element partition into ``nranks`` slabs along x (the synthetic stand-in for brkfix/METIS)."""
    nx = mesh.shape[0]
    ix = np.arange(mesh.num_elems) % nx
    return (ix * nranks) // nx


def brick_partition(mesh: Mesh, parts) -> np.ndarray:
    """Element partition of a structured box into ``parts`` = (px, py[, pz]) bricks: in 3-D every brick of a
    2 x 2 x 2 split touches all seven others (they share the centre node), the neighbourhood structure a
    METIS k-way partition produces (``metis_decomp.c:449-466``) rather than the two neighbours of a slab."""
    dim = len(mesh.shape)
    parts = tuple(parts) + (1,) * (dim - len(parts))
    idx = np.arange(mesh.num_elems)
    rank = np.zeros(mesh.num_elems, np.int64)
    mult = 1
    for d in range(dim):
        i_d = idx % mesh.shape[d]
        idx = idx // mesh.shape[d]
        rank += ((i_d * parts[d]) // mesh.shape[d]) * mult
        mult *= parts[d]
    return rank


def scattered_partition(mesh: Mesh, nranks: int, seed: int = 0, block: int = 2) -> np.ndarray:
    """A deliberately ragged element partition (random ranks per small block of elements): many neighbours per
    rank, non-contiguous sub-domains, ranks that own nodes only through ghost elements -- the worst a graph
    partitioner could hand to ``decompose``."""
    rng = np.random.default_rng(seed)
    nblk = (mesh.num_elems + block - 1) // block
    r = rng.integers(0, nranks, nblk)
    r[:nranks] = np.arange(nranks)  # every rank owns something
    return np.repeat(r, block)[: mesh.num_elems]


def slab_subdomain(make_problem, n, rank: int, nranks: int, elem_type: str = "HEX27", cols=None,
                   x_len: float = None) -> Subdomain:
    """Sub-domain ``rank`` of an ``nranks*n x n x n``-element box cut into x-slabs, built directly (the
    global mesh is never formed: this is how the weak-scaling bench gets 1M elements per GPU).

    ``cols`` (``nranks + 1`` increasing column boundaries) cuts a box of ``cols[-1]`` columns unevenly instead --
    the strong-scaling split of ONE ``n^3`` cavity (``cols[-1] == n``, ``x_len = 1``) over the ranks.

    ``make_problem(mesh)`` returns the Problem on the local mesh.  Ownership and ordering follow
    :func:`decompose` exactly (checked against it in the tests): shared planes belong to the lower
    rank, the rank keeps one ghost element column on its high-x side, external nodes are ordered
    by owner then global id.
    """
    from .mesh import ELEM_TABLE, box_mesh

    dim, _, order, _ = ELEM_TABLE[elem_type]
    if cols is None:
        cols = [r * n for r in range(nranks + 1)]
    total_cols = int(cols[-1])
    if x_len is None:
        x_len = total_cols / n
    c0 = int(cols[rank])
    n_own_cols = int(cols[rank + 1]) - c0
    c1 = int(cols[rank + 1]) + (1 if rank < nranks - 1 else 0)
    ncol = c1 - c0
    shape = (ncol,) + (n,) * (dim - 1)
    lo = (c0 * x_len / total_cols,) + (0.0,) * (dim - 1)
    hi = (c1 * x_len / total_cols,) + (1.0,) * (dim - 1)
    m = box_mesh(elem_type, shape, lo=lo, hi=hi)
    LX = m.lattice[0]
    lx = np.arange(m.num_nodes) % LX
    has_left = rank > 0
    has_right = rank < nranks - 1
    own_hi = order * n_own_cols  # highest owned local plane
    cat = np.zeros(m.num_nodes, np.int8)  # 0 internal, 1 boundary, 2 external-left, 3 external-right
    if has_left:
        cat[lx == 0] = 2
        cat[(lx >= 1) & (lx <= order)] = 1  # inside the lower rank's ghost column
    if has_right:
        cat[lx == own_hi] = 1
        cat[lx > own_hi] = 3
    perm = np.argsort(cat, kind="stable")  # local x-fastest order == global id order inside a category
    inv = np.empty_like(perm)
    inv[perm] = np.arange(len(perm))
    node_sets = {}
    for k, nodes in m.node_sets.items():
        if (k == 1 and has_left) or (k == 2 and has_right) or (k == 7 and rank != 0):
            nodes = nodes[:0]
        node_sets[k] = np.sort(inv[nodes]).astype(np.int32)
    lmesh = Mesh(elem_type, dim, np.ascontiguousarray(m.coords[:, perm]), inv[m.conn].astype(np.int32), node_sets,
                 shape, m.lattice)
    problem = make_problem(lmesh)
    n_int = int((cat == 0).sum())
    n_own = n_int + int((cat == 1).sum())
    # global ids (of the nranks*n-long box) for cross-checks
    GLX = order * total_cols + 1
    strides_l = np.array([1, LX, LX * m.lattice[1] if dim == 3 else 0][:dim])
    node_global = np.zeros(m.num_nodes, np.int64)
    rem = np.arange(m.num_nodes)
    gstr = [1, GLX, GLX * m.lattice[1] if dim == 3 else 0]
    for d in reversed(range(dim)):
        idx = rem // strides_l[d]
        rem = rem - idx * strides_l[d]
        node_global += (idx + (order * c0 if d == 0 else 0)) * gstr[d]
    sub = Subdomain(rank=rank, problem=problem, num_owned_nodes=n_own, num_internal_nodes=n_int,
                    node_global=node_global[perm], elem_global=np.zeros(0, np.int64),
                    elem_owned=(np.arange(m.num_elems) % ncol) < n_own_cols)
    first_l = problem.unknown_map()[0]

    def dofs_of(local_nodes_sorted):
        f0, f1 = first_l[local_nodes_sorted], first_l[local_nodes_sorted + 1]
        if len(f0) == 0:
            return np.zeros(0, np.int32)
        cnt = f1 - f0
        out = np.repeat(f0, cnt) + (np.arange(cnt.sum()) - np.repeat(np.cumsum(cnt) - cnt, cnt))
        return out.astype(np.int32)

    sends, recvs = [], []
    sub.neighbors = []
    if has_left:
        sub.neighbors.append(rank - 1)
        sends.append(dofs_of(np.sort(inv[np.nonzero((lx >= 1) & (lx <= order))[0]])))
        ext_left = inv[np.nonzero(lx == 0)[0]]
        recvs.append(int((first_l[ext_left + 1] - first_l[ext_left]).sum()))
    if has_right:
        sub.neighbors.append(rank + 1)
        sends.append(dofs_of(np.sort(inv[np.nonzero(lx == own_hi)[0]])))
        ext_right = inv[np.nonzero(lx > own_hi)[0]]
        recvs.append(int((first_l[ext_right + 1] - first_l[ext_right]).sum()))
    sub.list_dof_send = np.concatenate(sends).astype(np.int32) if sends else np.zeros(0, np.int32)
    sub.ptr_dof_send = np.concatenate([[0], np.cumsum([len(v) for v in sends])]).astype(np.int64)
    sub.num_dofs_recv = np.array(recvs, np.int64)
    return sub


def brick_subdomain(make_problem, n, rank: int, parts, elem_type: str = "HEX27") -> Subdomain:
    """Sub-domain ``rank`` of ONE ``n^3``-element unit cube cut into ``parts = (px, py, pz)`` bricks, built directly
    (the strong-scaling bench at 1M elements: the global mesh is never formed).  Same rules as :func:`decompose`
    (checked against it in the tests): a node belongs to the lowest rank touching it, i.e. to the brick whose
    half-open lattice interval (lo, hi] contains it in every direction; a rank keeps one ghost element layer on each
    of its high sides; local order internal, boundary, external (by owner, then global id)."""
    from .mesh import ELEM_TABLE, box_mesh

    dim, _, order, _ = ELEM_TABLE[elem_type]
    parts = tuple(parts) + (1,) * (dim - len(parts))
    nranks = int(np.prod(parts))
    b = []
    rr = rank
    for d in range(dim):
        b.append(rr % parts[d])
        rr //= parts[d]
    cuts = [[-((-k * n) // parts[d]) for k in range(parts[d] + 1)] for d in range(dim)]  # == brick_partition
    c0 = [cuts[d][b[d]] for d in range(dim)]
    c1 = [cuts[d][b[d] + 1] for d in range(dim)]
    has_low = [b[d] > 0 for d in range(dim)]
    has_high = [b[d] < parts[d] - 1 for d in range(dim)]
    ncol = [c1[d] - c0[d] + (1 if has_high[d] else 0) for d in range(dim)]
    lo = tuple(c0[d] / n for d in range(dim))
    hi = tuple((c0[d] + ncol[d]) / n for d in range(dim))
    m = box_mesh(elem_type, tuple(ncol), lo=lo, hi=hi)
    lat = m.lattice
    nid = np.arange(m.num_nodes)
    l = []
    stride = 1
    for d in range(dim):
        l.append((nid // stride) % lat[d])
        stride *= lat[d]
    own_hi = [order * (c1[d] - c0[d]) for d in range(dim)]
    owned = np.ones(m.num_nodes, bool)
    bnd = np.zeros(m.num_nodes, bool)
    owner_b = []
    for d in range(dim):
        low_face = has_low[d] & (l[d] == 0)
        beyond = has_high[d] & (l[d] > own_hi[d])
        owned &= ~low_face & ~beyond
        bnd |= (has_low[d] & (l[d] <= order)) | (has_high[d] & (l[d] == own_hi[d]))
        owner_b.append(np.where(low_face, b[d] - 1, np.where(beyond, b[d] + 1, b[d])))
    # a node outside r's half-open box belongs to the brick that contains it in EVERY direction
    owner = np.zeros(m.num_nodes, np.int64)
    mult = 1
    for d in range(dim):
        owner += owner_b[d] * mult
        mult *= parts[d]
    GL = [order * n + 1 for _ in range(dim)]
    gid = np.zeros(m.num_nodes, np.int64)
    mult = 1
    for d in range(dim):
        gid += (order * c0[d] + l[d]) * mult
        mult *= GL[d]
    cat = np.where(owned, np.where(bnd, 1, 0), 2)
    key_owner = np.where(owned, 0, owner)
    perm = np.lexsort((gid, key_owner, cat))  # category, then owner (externals), then global id
    inv = np.empty_like(perm)
    inv[perm] = np.arange(len(perm))
    node_sets = {}
    for k, nodes in m.node_sets.items():
        if 1 <= k <= 2 * dim:
            d, high = (k - 1) // 2, (k - 1) % 2 == 1
            if (high and c0[d] + ncol[d] != n) or (not high and c0[d] != 0):  # the local face is not a global one
                nodes = nodes[:0]
        if k == 7 and rank != 0:
            nodes = nodes[:0]
        node_sets[k] = np.sort(inv[nodes]).astype(np.int32)
    lmesh = Mesh(elem_type, dim, np.ascontiguousarray(m.coords[:, perm]), inv[m.conn].astype(np.int32), node_sets,
                 tuple(ncol), m.lattice)
    problem = make_problem(lmesh)
    n_int = int((cat == 0).sum())
    n_own = n_int + int((cat == 1).sum())
    eidx = np.arange(m.num_elems)
    e_owned = np.ones(m.num_elems, bool)
    stride = 1
    for d in range(dim):
        e_owned &= ((eidx // stride) % ncol[d]) < (c1[d] - c0[d])
        stride *= ncol[d]
    sub = Subdomain(rank=rank, problem=problem, num_owned_nodes=n_own, num_internal_nodes=n_int, node_global=gid[perm],
                    elem_global=np.zeros(0, np.int64), elem_owned=e_owned)
    first_l = problem.unknown_map()[0]

    def dofs_of(local_nodes):
        f0, f1 = first_l[local_nodes], first_l[local_nodes + 1]
        if len(f0) == 0:
            return np.zeros(0, np.int32)
        cnt = f1 - f0
        out = np.repeat(f0, cnt) + (np.arange(cnt.sum()) - np.repeat(np.cumsum(cnt) - cnt, cnt))
        return out.astype(np.int32)

    # what each neighbour brick q sees of my owned nodes: my owned nodes inside q's local lattice box
    import itertools

    glat = [order * c0[d] + l[d] for d in range(dim)]
    sends, recvs, nbrs = {}, {}, set()
    for off in itertools.product((-1, 0, 1), repeat=dim):
        if not any(off):
            continue
        qb = [b[d] + off[d] for d in range(dim)]
        if any(qb[d] < 0 or qb[d] >= parts[d] for d in range(dim)):
            continue
        q = 0
        mult = 1
        for d in range(dim):
            q += qb[d] * mult
            mult *= parts[d]
        inside = owned.copy()
        for d in range(dim):
            q0 = order * cuts[d][qb[d]]
            q1 = order * (cuts[d][qb[d] + 1] + (1 if qb[d] < parts[d] - 1 else 0))
            inside &= (glat[d] >= q0) & (glat[d] <= q1)
        nodes = np.nonzero(inside)[0]
        if len(nodes):
            nodes = nodes[np.argsort(gid[nodes])]
            sends[q] = dofs_of(inv[nodes])
            nbrs.add(q)
        ext_q = np.nonzero(~owned & (owner == q))[0]
        if len(ext_q):
            recvs[q] = int((first_l[inv[ext_q] + 1] - first_l[inv[ext_q]]).sum())
            nbrs.add(q)
    sub.neighbors = sorted(nbrs)
    send_lists = [sends.get(q, np.zeros(0, np.int32)) for q in sub.neighbors]
    sub.list_dof_send = np.concatenate(send_lists).astype(np.int32) if send_lists else np.zeros(0, np.int32)
    sub.ptr_dof_send = np.concatenate([[0], np.cumsum([len(v) for v in send_lists])]).astype(np.int64)
    sub.num_dofs_recv = np.array([recvs.get(q, 0) for q in sub.neighbors], np.int64)
    assert nranks >= 1
    return sub


def exchange_dof(x, sub: Subdomain, group=None, pack=None):
    """Refresh the external (ghost) tail of the local vector ``x`` (torch tensor, CPU or CUDA).

    ``pack(x, list) -> buffer`` may be supplied to run the gather on the device through the C ABI;
    the default is a torch ``index_select``.  Message layout = the reference's: one contiguous send
    block per neighbour (``ptr_dof_send``), one contiguous receive block per neighbour in the tail.
    """
    import torch
    import torch.distributed as dist

    if not sub.neighbors:
        return x
    cache = sub.__dict__.setdefault("_idx_cache", {})
    idx = cache.get(x.device)
    if idx is None:  # the send list is static: upload it once per device
        idx = cache[x.device] = torch.as_tensor(sub.list_dof_send, dtype=torch.long, device=x.device)
    send = pack(x, idx) if pack is not None else x.index_select(0, idx)
    ops = []
    tail = sub.num_owned_dofs
    for k, p in enumerate(sub.neighbors):
        s0, s1 = int(sub.ptr_dof_send[k]), int(sub.ptr_dof_send[k + 1])
        nrecv = int(sub.num_dofs_recv[k])
        if s1 > s0:
            ops.append(dist.P2POp(dist.isend, send[s0:s1], p, group))
        if nrecv:
            ops.append(dist.P2POp(dist.irecv, x[tail:tail + nrecv], p, group))
        tail += nrecv
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    return x


def allreduce_flags(flags, group=None):
    """The scalar all-reduces at the end of matrix_fill_full (``mm_fill.c:271-281``), fused into one small tensor and
    reduced with MAX: the three domain-failure flags are 0/1 (MPI_MAX in the reference), and an error count reduced
    with MAX still answers the only question the caller asks of it -- did any rank fail."""
    import torch
    import torch.distributed as dist

    t = torch.as_tensor(flags, dtype=torch.int32)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return t


def peer_exchange_payload(sub: Subdomain, handles: bytes = b"") -> dict:
    """What a rank publishes once for the peer-memory exchange: its IPC handles and, per neighbour, its
    ``list_dof_send`` block (``dp_map_comm_vec.c:224-461``)."""
    return {"handles": handles, "neighbors": [int(p) for p in sub.neighbors],
            "send": {int(p): np.asarray(sub.list_dof_send[int(sub.ptr_dof_send[k]):int(sub.ptr_dof_send[k + 1])], np.int32)
                     for k, p in enumerate(sub.neighbors)}}


def peer_recv_lists(sub: Subdomain, everyone):
    """From every rank's payload: (neighbour handles, my slot in each neighbour's list, recv_ptr, recv_list) where
    ``recv_list`` holds, per neighbour and in the order of this rank's external tail, the dof indices IN THE
    NEIGHBOUR'S NUMBERING to read -- the neighbour's send block for this rank: the pairing the reference's
    send/receive establishes (``dp_comm.c:77-96``)."""
    handles, slots, recv_ptr, recv_list = [], [], [0], []
    for k, p in enumerate(sub.neighbors):
        other = everyone[int(p)]
        handles.append(other["handles"])
        slots.append(other["neighbors"].index(sub.rank))
        block = other["send"][sub.rank]
        if len(block) != int(sub.num_dofs_recv[k]):
            raise RuntimeError(f"rank {sub.rank}: neighbour {p} sends {len(block)} dofs, {int(sub.num_dofs_recv[k])} expected")
        recv_list.append(block)
        recv_ptr.append(recv_ptr[-1] + len(block))
    rl = np.concatenate(recv_list).astype(np.int32) if recv_list else np.zeros(0, np.int32)
    return handles, slots, recv_ptr, rl


def setup_peer_exchange(mf, sub: Subdomain, group=None):
    """Wire the ranks' GPU contexts together for ``MatrixFill.exchange_dof`` (``goma_gpu_exchange_*``).
    Host plumbing only, once per problem (``torch.distributed.all_gather_object`` as the set-up transport)."""
    import torch.distributed as dist

    everyone = [None] * dist.get_world_size(group)
    dist.all_gather_object(everyone, peer_exchange_payload(sub, mf.exchange_export()), group=group)
    handles, slots, recv_ptr, rl = peer_recv_lists(sub, everyone)
    mf.exchange_setup(handles, slots, recv_ptr, rl, sub.num_owned_dofs)
    dist.barrier(group=group)  # every flag block exists and is zero before the first epoch is published
