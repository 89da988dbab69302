"""Multi-GPU: contiguous element-block partition of the canonical global mesh + ghost refresh.

One process per GPU.  Rank g of G owns the contiguous element range [g*Nel/G, (g+1)*Nel/G) of the
canonical numbering (element e = face*ne^2 + ey*ne + ex) -- free of the reference's 6*n^2 rank
constraint (README.md:32, scr/Setup.py:25-29) -- and with it the edges and faces whose global ids
fall into that element's block (scr/Proc2.py:105-123).  Operators are applied owner-computes: every
rank holds, besides its owned elements, the west/south neighbour elements of its owned elements as
read-only halo elements, so that outputs on owned DOFs are complete and need NO reduction.  What
the reference does with VecScatter(gtol_1, ADD_VALUES, SCATTER_REVERSE) after its matrix-free
assembly (eul/Assembly.cpp:2194-2195) and with the forward INSERT scatter before it
(eul/Euler_2.cpp:1455-1456) collapses into ONE exchange: a ghost refresh of the INPUT field,
implemented as pack kernel -> NCCL send/recv over NVLink (torch.distributed) -> unpack kernel.

The partition logic is pure numpy (tested on CPU with world_size-2 gloo); only DistributedEngine
touches the GPU.
"""
import numpy as np

from .engine import Engine, SUBSET_BOUNDARY, SUBSET_INTERIOR
from .mesh import Basis


def element_range(nel, rank, world):
    return (rank * nel) // world, ((rank + 1) * nel) // world


def owner_rank_of_element(e, nel, world):
    """inverse of element_range (vectorised)"""
    e = np.asarray(e, dtype=np.int64)
    r = (e * world) // nel
    # correct the integer-division edge cases
    lo = (r * nel) // world
    hi = ((r + 1) * nel) // world
    r = np.where(e < lo, r - 1, np.where(e >= hi, r + 1, r))
    return r


def west_south_neighbours(mesh):
    """nbr[e, 0] / nbr[e, 1]: the element across the west / south side of e (-1 if none)."""
    if getattr(mesh, "_ws_nbr", None) is not None:
        return mesh._ws_nbr
    p, nel = mesh.p, mesh.nel
    n1e = p * (p + 1)
    dofs = np.concatenate([mesh.el1x.ravel(), mesh.el1y.ravel()]).astype(np.int64)
    elem = np.concatenate([np.repeat(np.arange(nel), n1e), np.repeat(np.arange(nel), n1e)])
    order = np.argsort(dofs, kind="stable")
    ds, es = dofs[order], elem[order]
    # every edge of a closed mesh is used by exactly two (element, slot) pairs
    first = np.searchsorted(ds, np.arange(mesh.N1), side="left")
    last = np.searchsorted(ds, np.arange(mesh.N1), side="right")
    nbr = -np.ones((nel, 2), dtype=np.int64)
    e_ids = np.arange(nel)
    for side, d in ((0, mesh.el1x[:, 0].astype(np.int64)), (1, mesh.el1y[:, 0].astype(np.int64))):
        a, b = es[first[d]], es[np.minimum(last[d] - 1, len(es) - 1)]
        cnt = last[d] - first[d]
        other = np.where(a == e_ids, b, a)
        # a periodic one-element-wide mesh makes an element its own neighbour (both uses are e itself)
        nbr[:, side] = np.where(cnt >= 2, other, -1)
    mesh._ws_nbr = nbr
    return nbr


def node_min_element(mesh):
    """For every global node the lowest-numbered element around it (its owner element: nodes follow elements)."""
    if getattr(mesh, "_node_min_el", None) is None:
        nme = np.full(mesh.N0, mesh.nel, dtype=np.int64)
        np.minimum.at(nme, mesh.el0.ravel().astype(np.int64), np.repeat(np.arange(mesh.nel, dtype=np.int64), mesh.el0.shape[1]))
        mesh._node_min_el = nme
    return mesh._node_min_el


def node_elements(mesh):
    """CSR node -> elements around it: (ptr[N0 + 1], elems)."""
    if getattr(mesh, "_node_els", None) is None:
        nodes = mesh.el0.ravel().astype(np.int64)
        els = np.repeat(np.arange(mesh.nel, dtype=np.int64), mesh.el0.shape[1])
        order = np.argsort(nodes, kind="stable")
        ptr = np.searchsorted(nodes[order], np.arange(mesh.N0 + 1))
        mesh._node_els = (ptr, els[order])
    return mesh._node_els


class Partition:
    """Local view of rank `rank` of `world`: owned + halo elements, owned-first local DOF numbering,
    ghost lists grouped by owner rank.  All index arrays are numpy int32/int64."""

    def __init__(self, mesh, rank, world):
        p = mesh.p
        self.p, self.rank, self.world = p, rank, world
        nel = mesh.nel
        e0, e1 = element_range(nel, rank, world)
        self.e0, self.e1 = e0, e1
        owned = np.arange(e0, e1, dtype=np.int64)
        nbr = west_south_neighbours(mesh)[e0:e1].ravel()
        nbr = nbr[nbr >= 0]
        halo_ws = np.setdiff1d(np.unique(nbr), owned)
        # 0-forms: a node belongs to the rank of the lowest-numbered element around it; for the operators that SUM over
        # the elements around a node (M0h, E01, M0h_up) every element around an owned node is held as a halo element too
        nme = node_min_element(mesh)
        self.node_owner_of = lambda ids: owner_rank_of_element(nme[ids], nel, world)
        own_nodes = np.nonzero((nme >= e0) & (nme < e1))[0]
        nptr, nels = node_elements(mesh)
        around = np.unique(np.concatenate([nels[nptr[n]:nptr[n + 1]] for n in own_nodes])) if len(own_nodes) else np.zeros(0, np.int64)
        halo_node = np.setdiff1d(around, owned)
        halo = np.union1d(halo_ws, halo_node)
        self.halo_ws = halo_ws
        # owned elements that read no row owned elsewhere (interior) come first, the others (boundary) last: the
        # engine's interior / boundary subsets are then index ranges, and the fused ghost-refresh kernel walks the
        # elements in storage order (same rule as mimsem_gpu_set_ghosts)
        b1_, b2_ = 2 * p * p, p * p
        def foreign(e):
            d1 = np.concatenate([mesh.el1x[e], mesh.el1y[e]], axis=1).astype(np.int64) // b1_
            d2 = mesh.el2[e].astype(np.int64) // b2_
            return ((d1 < e0) | (d1 >= e1)).any(axis=1) | ((d2 < e0) | (d2 >= e1)).any(axis=1)
        bnd = foreign(owned)
        ws_own = west_south_neighbours(mesh)[e0:e1]
        for side in (0, 1):
            n = ws_own[:, side]
            ok = n >= 0
            bnd[ok] |= foreign(n[ok])
        owned = np.concatenate([owned[~bnd], owned[bnd]])
        self.n_interior = int((~bnd).sum())
        self.elements = np.concatenate([owned, halo])          # local element -> global element
        self.nel_owned, self.nel_total = len(owned), len(owned) + len(halo)
        L = self.elements
        b1, b2 = 2 * p * p, p * p

        def local_numbering(ids, block, needed=None):
            ids = np.unique(ids.astype(np.int64))
            own = ids[(ids // block >= e0) & (ids // block < e1)] if block else ids
            ghost = np.setdiff1d(ids, own) if block else ids[:0]
            if needed is not None:
                # ghosts some kernel actually reads come first (they are the ones exchanged), the rest after
                need = np.intersect1d(ghost, needed)
                ghost = np.concatenate([need, np.setdiff1d(ghost, need)])
                return np.concatenate([own, ghost]), len(own), len(need)
            return np.concatenate([own, ghost]), len(own), len(ghost)

        # 1-form rows the kernels read: every edge of an owned element and, of a west / south halo element, the
        # edge family ACROSS its far line (all y-normal edges behind an east column, all x-normal edges behind a
        # north row -- the WOTH / SOTH slots of the tile kernel); the halo element's remaining edges are never read
        ws = west_south_neighbours(mesh)
        need = [mesh.el1x[e0:e1].ravel(), mesh.el1y[e0:e1].ravel()]
        for side, shared in ((0, mesh.el1x[e0:e1, 0]), (1, mesh.el1y[e0:e1, 0])):
            n = ws[e0:e1, side]
            ok = n >= 0
            nn, sh = n[ok], shared[ok].astype(np.int64)
            far_is_east = (mesh.el1x[nn].astype(np.int64) == sh[:, None]).any(axis=1)
            need.append(mesh.el1y[nn[far_is_east]].ravel())
            need.append(mesh.el1x[nn[~far_is_east]].ravel())
        needed1 = np.unique(np.concatenate(need).astype(np.int64))

        # 1-forms: edge g is owned by element g // (2 p^2); 2-forms: face g by element g // p^2
        self.g1, self.n1_owned, n1_need = local_numbering(np.concatenate([mesh.el1x[L].ravel(), mesh.el1y[L].ravel()]), b1, needed1)
        self.n1_halo = self.n1_owned + n1_need          # rows [n1_owned, n1_halo) are refreshed from their owners
        # 2-forms the element kernels read: the faces of the west / south halo elements (M1h's neighbour coefficient, E12)
        needed2 = np.unique(mesh.el2[np.concatenate([owned, halo_ws])].ravel().astype(np.int64))
        self.g2, self.n2_owned, n2_need = local_numbering(mesh.el2[L].ravel(), b2, needed2)
        self.n2_halo = self.n2_owned + n2_need
        # nodes: owned first (ascending), then the ghosts grouped by owner rank
        ids0 = np.unique(mesh.el0[L].ravel().astype(np.int64))
        own0 = ids0[(nme[ids0] >= e0) & (nme[ids0] < e1)]
        gh0 = np.setdiff1d(ids0, own0)
        gh0 = gh0[np.lexsort((gh0, self.node_owner_of(gh0)))] if len(gh0) else gh0
        self.g0, self.n0_owned = np.concatenate([own0, gh0]), len(own0)
        self.gq, _, _ = local_numbering(mesh.elq[L].ravel(), 0)
        self.n0, self.n1, self.n2, self.nq = len(self.g0), len(self.g1), len(self.g2), len(self.gq)

        def to_local(gids, table):
            order = np.argsort(gids, kind="stable")
            t = table.astype(np.int64)
            out = order[np.searchsorted(gids[order], t)]
            assert np.array_equal(gids[out], t)
            return out.astype(np.int32)

        self.el1x = to_local(self.g1, mesh.el1x[L])
        self.el1y = to_local(self.g1, mesh.el1y[L])
        self.el2 = to_local(self.g2, mesh.el2[L])
        self.el0 = to_local(self.g0, mesh.el0[L])
        self.elq = to_local(self.gq, mesh.elq[L])
        # ghosts grouped by owner rank (ascending global id inside a group): the rows the element kernels read ...
        self.recv = {0: self._group(self.g0[self.n0_owned:], 0, nel, self.n0_owned),
                     1: self._group(self.g1[self.n1_owned:self.n1_halo], b1, nel, self.n1_owned),
                     2: self._group(self.g2[self.n2_owned:self.n2_halo], b2, nel, self.n2_owned)}
        # ... and ALL ghost rows (the node-sum operators M0h, E01, M0h_up read every element around an owned node)
        self.recv_ext = {1: self._group(self.g1[self.n1_owned:], b1, nel, self.n1_owned),
                         2: self._group(self.g2[self.n2_owned:], b2, nel, self.n2_owned)}

    def _group(self, ghosts, block, nel, offset):
        owner = owner_rank_of_element(ghosts // block, nel, self.world) if block else self.node_owner_of(ghosts)
        out = {}
        for q in np.unique(owner):
            sel = np.nonzero(owner == q)[0]
            out[int(q)] = dict(local=(sel + offset).astype(np.int32), glob=ghosts[sel])
        return out

    def owned_global(self, space):
        return {0: self.g0[:self.n0_owned], 1: self.g1[:self.n1_owned], 2: self.g2[:self.n2_owned]}[space]

    def n_owned(self, space):
        return {0: self.n0_owned, 1: self.n1_owned, 2: self.n2_owned}[space]


def send_lists(mesh, rank, world):
    """For every peer q: the LOCAL (owned) ids on `rank` that q holds as ghosts, in q's receive order.
    Every rank can build every partition, so no index exchange is needed at setup."""
    me = Partition(mesh, rank, world)
    out = {0: {}, 1: {}, 2: {}}
    ext = {1: {}, 2: {}}
    for q in range(world):
        if q == rank:
            continue
        other = Partition(mesh, q, world)
        for space in (0, 1, 2):
            own = me.owned_global(space)          # ascending: local id of an owned DOF = its position
            for table, dst in ((other.recv, out), (getattr(other, "recv_ext", {}), ext)):
                grp = table.get(space, {}).get(rank) if space in dst else None
                if grp is None:
                    continue
                loc = np.searchsorted(own, grp["glob"])
                assert np.array_equal(own[loc], grp["glob"])
                dst[space][q] = loc.astype(np.int32)
    me.sends_ext = ext
    return me, out


class DistributedEngine:
    """Engine of one rank + the ghost-refresh plan.  Fields are local column-layout tensors with
    n_space rows (owned rows are complete after an apply; ghost rows are refreshed by exchange())."""

    SUPPORTED = ("M1", "M1h", "M2", "M2h", "K", "E21", "E12", "M0", "M0h", "E10", "E01", "R", "R_up", "M0h_up", "UtQW")

    def __init__(self, mesh, thick, rank, world, device, max_levels=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank, self.world = rank, world
        self.part, sends = send_lists(mesh, rank, world)
        P = self.part
        eng = Engine(device)
        eng.set_basis(Basis(mesh.p, mesh.m))
        # sums over the elements around a node run in GLOBAL element order: bitwise equal to the single-GPU result
        keys = np.ascontiguousarray(P.elements, dtype=np.int32)
        from .lib import check, _ip
        check(eng.L.mimsem_gpu_set_element_keys(eng._h, len(keys), keys.ctypes.data_as(_ip)))
        eng.set_topo(P.el0, P.el1x, P.el1y, P.el2, P.elq, P.n0, P.n1, P.n2, P.nq, nel_owned=P.nel_owned, mode=0)
        eng.set_option("n0_owned", P.n0_owned)
        self.n_interior, self.n_boundary = eng.set_ghosts(P.n1_owned, P.n2_owned)
        assert self.n_interior == P.n_interior, "partition and engine disagree on the interior / boundary split"
        eng.set_geom(mesh.J[P.elements], mesh.det[P.elements])
        if thick is not None:
            eng.set_thickness(np.ascontiguousarray(thick[:, P.gq]))
        self.engine = eng
        self.device = device
        dev = "cuda:%d" % device
        def build_plan(space, recv, send):
            perm = eng.permutation(space).astype(np.int64)
            plan = []
            for q in sorted(set(recv) | set(send)):
                r, s = recv.get(q), send.get(q)
                rrows = torch.from_numpy(perm[r["local"]].astype(np.int32)).to(dev) if r is not None else None
                srows = torch.from_numpy(perm[s].astype(np.int32)).to(dev) if s is not None else None
                plan.append((q, srows, rrows))
            return plan
        # the rows the element kernels read (peer-to-peer inboxes) and, for the node-sum operators, ALL ghost rows (NCCL)
        self.plan = {space: build_plan(space, P.recv[space], sends[space]) for space in (0, 1, 2)}
        self.plan_ext = {space: build_plan(space, P.recv_ext[space], P.sends_ext[space]) for space in (1, 2)}
        self._bufs = {}
        self.comm_stream = torch.cuda.Stream(device=device, priority=-1)   # halo kernels get SM slots ahead of the bulk kernel
        self.overlap = True
        # levels per ghost row the halo inboxes are allocated for: explicit, else the thickness table's, else 1
        self.nk_max = int(max_levels) if max_levels else (1 if thick is None else int(thick.shape[0]))
        eng.set_option("halo_max_levels", self.nk_max)
        self.p2p = None
        self.graph_safe = False
        self._inbox = {}
        import os
        self.fused = os.environ.get("MIMSEM_FUSED_HALO", "1") != "0"
        # upper limit of push CTAs per fused launch (0: none).  Launches of a burst overlap, the push is off the critical
        # path there, and every push CTA holds a tile's worth of shared memory while it waits for NVLink
        self.push_ctas_cap = int(os.environ.get("MIMSEM_PUSH_CTAS", "0"))
        # fused M1: data + flag (default), or self-validating 16-byte cells without fence and flag (MIMSEM_HALO_LL=1;
        # measured slower at 2 and 4 GPUs, see profiles/r02_summary.md)
        self.ll = os.environ.get("MIMSEM_HALO_LL", "0") != "0"
        if world > 1 and os.environ.get("MIMSEM_HALO", "p2p") == "p2p":
            self._setup_p2p(P, sends)

    # ---------------------------------------------------------------- peer-to-peer halo (no NCCL on the data path)
    MAXP = 16
    NBUF = 3      # inbox copies per space; the push / pull kernels and the fused M1 launch follow the same rule (epoch e -> copy e % NBUF)

    def _setup_p2p(self, P, sends):
        """Allocate this rank's inbox / flag buffer, exchange IPC handles and layouts (control plane: torch.distributed
        object collectives), map every peer's buffer and build the push / pull descriptor arrays."""
        import ctypes as C
        torch, dist, eng = self.torch, self.dist, self.engine
        MAXP, nk = self.MAXP, self.nk_max
        spaces = (0, 1, 2)
        nsp = len(spaces)
        recv_peers = {s: sorted(P.recv[s]) for s in spaces}
        send_peers = {s: sorted(sends[s]) for s in spaces}
        assert all(len(v) <= MAXP for v in list(recv_peers.values()) + list(send_peers.values()))
        hdr_bytes = (2 * nsp + 1) * MAXP * 8           # flags[nsp][MAXP], acks[nsp][MAXP], acks of the in-band (LL) 1-form inbox [MAXP]
        # inbox of a space: [NBUF copies][ghost rows of the space, in ghost order][nk]; a peer's share is the run of
        # rows it owns (ghosts are sorted by global id, owners hold contiguous id ranges)
        layout = {}                                    # (space, peer) -> (slot, first inbox row, nrows)
        region = {}                                    # space -> (offset in bytes, total ghost rows)
        off = hdr_bytes
        for si, s in enumerate(spaces):
            n_own = P.n_owned(s)
            row = 0
            for slot, q in enumerate(recv_peers[s]):
                loc = P.recv[s][q]["local"].astype(np.int64)
                assert np.array_equal(loc, n_own + row + np.arange(len(loc))), "ghost rows of a peer must be one run"
                layout[(s, q)] = (slot, row, len(loc))
                row += len(loc)
            # per-copy stride in doubles, rounded up to an even number: every inbox copy starts 16-byte aligned
            stride = row * nk + ((row * nk) & 1)
            region[s] = (off, row, stride)
            off += self.NBUF * stride * 8
        # in-band (LL) inbox of the 1-form space for the fused M1 launch: [NBUF][ghost rows][nk] 16-byte cells
        off = (off + 15) // 16 * 16
        ll_cells = region[1][1] * nk
        ll_region = (off, ll_cells)
        off += self.NBUF * ll_cells * 16
        # reduction area of the partitioned CG (mimsem_gpu_solve_M1_dist): [2 parities][world][3 sums][64 levels] cells
        red_off = off
        off += 2 * self.world * 3 * 64 * 16
        total = max(off, hdr_bytes + 16)
        base = C.c_void_p()
        handle = C.create_string_buffer(64)
        from .lib import check
        check(eng.L.mimsem_gpu_ipc_alloc(eng._h, total, C.byref(base), handle))
        mine = dict(handle=handle.raw, layout=layout, region=region, ll_region=ll_region, red_off=red_off,
                    send_slot={(s, q): i for s in spaces for i, q in enumerate(send_peers[s])})
        everyone = [None] * self.world
        dist.all_gather_object(everyone, mine)
        peer_base = {}
        for q in range(self.world):
            if q == self.rank:
                continue
            if True:   # every peer is mapped: the CG reduction is all-to-all even where no ghost row is shared
                ptr = C.c_void_p()
                check(eng.L.mimsem_gpu_ipc_open(eng._h, everyone[q]["handle"], C.byref(ptr)))
                peer_base[q] = ptr.value
        dt = np.dtype([("rows", "<u8"), ("nrows", "<i4"), ("row0", "<i4"), ("inbox", "<u8"), ("stride", "<i8"), ("signal", "<u8"),
                       ("wait", "<u8")])
        assert dt.itemsize == 48
        dev = "cuda:%d" % self.device
        my = base.value
        keep = []
        plans = {}
        for si, s in enumerate(spaces):
            perm = eng.permutation(s).astype(np.int64)
            push = np.zeros(len(send_peers[s]), dtype=dt)
            for i, q in enumerate(send_peers[s]):
                rows = torch.from_numpy(perm[sends[s][q]].astype(np.int32)).to(dev)
                keep.append(rows)
                slot_on_q, row0_on_q, n_on_q = everyone[q]["layout"][(s, self.rank)]
                off_on_q, nghost_on_q, stride_on_q = everyone[q]["region"][s]
                assert n_on_q == rows.numel()
                push[i] = (rows.data_ptr(), rows.numel(), row0_on_q, peer_base[q] + off_on_q, stride_on_q,
                           peer_base[q] + (si * MAXP + slot_on_q) * 8,            # flag on q
                           my + ((nsp + si) * MAXP + i) * 8)                       # ack from q, in my memory
            pull = np.zeros(len(recv_peers[s]), dtype=dt)
            off_b, nghost, stride_b = region[s]
            for i, q in enumerate(recv_peers[s]):
                rows = torch.from_numpy(perm[P.recv[s][q]["local"]].astype(np.int32)).to(dev)
                keep.append(rows)
                slot, row0, n = layout[(s, q)]
                ack_slot_on_q = everyone[q]["send_slot"][(s, self.rank)]
                pull[i] = (rows.data_ptr(), n, row0, my + off_b, stride_b,
                           peer_base[q] + ((nsp + si) * MAXP + ack_slot_on_q) * 8,   # ack on q
                           my + (si * MAXP + slot) * 8)                                # flag from q, in my memory
            dpush = torch.from_numpy(push.view(np.uint8).copy()).to(dev)
            dpull = torch.from_numpy(pull.view(np.uint8).copy()).to(dev)
            epochs = torch.zeros(2, dtype=torch.int64, device=dev)   # [push counter, pull counter]
            plans[s] = (len(push), dpush, len(pull), dpull, epochs)
            self._inbox[s] = (my + off_b, stride_b, int(sum(int(r["nrows"]) for r in push)))
        # descriptors of the in-band protocol (1-forms only): same rows, the peers' cell inboxes, separate ack words
        s = 1
        perm = eng.permutation(s).astype(np.int64)
        push = np.zeros(len(send_peers[s]), dtype=dt)
        for i, q in enumerate(send_peers[s]):
            rows = torch.from_numpy(perm[sends[s][q]].astype(np.int32)).to(dev)
            keep.append(rows)
            slot_on_q, row0_on_q, n_on_q = everyone[q]["layout"][(s, self.rank)]
            off_on_q, cells_on_q = everyone[q]["ll_region"]
            push[i] = (rows.data_ptr(), rows.numel(), row0_on_q, peer_base[q] + off_on_q, cells_on_q, 0, my + (2 * nsp * MAXP + i) * 8)
        pull = np.zeros(len(recv_peers[s]), dtype=dt)
        for i, q in enumerate(recv_peers[s]):
            slot, row0, n = layout[(s, q)]
            ack_slot_on_q = everyone[q]["send_slot"][(s, self.rank)]
            pull[i] = (0, n, row0, my + ll_region[0], ll_region[1], peer_base[q] + (2 * nsp * MAXP + ack_slot_on_q) * 8, 0)
        self._ll = dict(npush=len(push), dpush=torch.from_numpy(push.view(np.uint8).copy()).to(dev), npull=len(pull),
                        dpull=torch.from_numpy(pull.view(np.uint8).copy()).to(dev), epochs=torch.zeros(2, dtype=torch.int64, device=dev),
                        inbox=my + ll_region[0], stride=ll_region[1], push_rows=int(sum(int(r["nrows"]) for r in push)))
        areas = np.array([(my if q == self.rank else peer_base[q]) + everyone[q]["red_off"] for q in range(self.world)], dtype=np.uint64)
        self._red = dict(areas=torch.from_numpy(areas.view(np.int64).copy()).to(dev), seq=torch.zeros(1, dtype=torch.int64, device=dev))
        err = torch.zeros(1, dtype=torch.int32, device=dev)
        self.p2p = dict(plans=plans, err=err, keep=keep, base=base, peer_base=peer_base)
        self.graph_safe = True
        torch.cuda.synchronize(self.device)
        dist.barrier()   # every inbox is allocated and zeroed before anybody pushes

    def halo_error(self):
        """True if a p2p exchange timed out waiting for a peer."""
        return bool(self.p2p is not None and int(self.p2p["err"].item()) != 0)

    def _check_levels(self, field):
        if field.shape[1] > self.nk_max:
            from .lib import MimsemError
            raise MimsemError("ghost refresh of %d levels, but the halo inboxes hold %d (DistributedEngine(max_levels=...))"
                              % (field.shape[1], self.nk_max))

    def push(self, field, space):
        self._check_levels(field)
        npush, dpush, _, _, epochs = self.p2p["plans"][space]
        eng = self.engine
        from .lib import check
        check(eng.L.mimsem_gpu_halo_push(eng._h, npush, dpush.data_ptr(), field.shape[1], field.shape[1], self.NBUF, field.data_ptr(),
                                         epochs.data_ptr(), self.p2p["err"].data_ptr(), eng._stream()))

    def pull(self, field, space):
        self._check_levels(field)
        _, _, npull, dpull, epochs = self.p2p["plans"][space]
        eng = self.engine
        from .lib import check
        check(eng.L.mimsem_gpu_halo_pull(eng._h, npull, dpull.data_ptr(), field.shape[1], field.shape[1], self.NBUF, field.data_ptr(),
                                         epochs.data_ptr() + 8, self.p2p["err"].data_ptr(), eng._stream()))

    # sizes / plumbing shared with Engine
    def space_sizes(self, op):
        return self.engine.space_sizes(op)

    @property
    def launch_count(self):
        return self.engine.launch_count

    def halo_bytes(self, space, nlev):
        return sum((0 if s is None else s.numel()) for _, s, _ in self.plan[space]) * nlev * 8

    def exchange(self, field, space, ext=False):
        """Ghost refresh of a local field.  p2p mode: push kernel (stores into the peers' inboxes) + pull kernel;
        otherwise pack -> NCCL send/recv -> unpack.  ext: ALL ghost rows of the space (what the operators that sum over
        the elements around a node read), always through NCCL -- those are not on the hot path."""
        if self.p2p is not None and not ext:
            self.push(field, space)
            self.pull(field, space)
            return field
        eng, torch, dist = self.engine, self.torch, self.dist
        nlev = field.shape[1]
        st = eng._stream()
        ops, unpack = [], []
        for q, srows, rrows in (self.plan_ext if ext else self.plan)[space]:
            if srows is not None:
                sb = self._buf(("s", space, q, nlev, ext), srows.numel() * nlev, field.device)
                eng.L.mimsem_gpu_gather_rows(eng._h, srows.numel(), nlev, nlev, srows.data_ptr(), field.data_ptr(), sb.data_ptr(), st)
                ops.append(dist.P2POp(dist.isend, sb, q))
            if rrows is not None:
                rb = self._buf(("r", space, q, nlev, ext), rrows.numel() * nlev, field.device)
                ops.append(dist.P2POp(dist.irecv, rb, q))
                unpack.append((rrows, rb))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        for rrows, rb in unpack:
            eng.L.mimsem_gpu_scatter_rows(eng._h, rrows.numel(), nlev, nlev, rrows.data_ptr(), rb.data_ptr(), field.data_ptr(), st)
        return field

    def _buf(self, key, n, device):
        b = self._bufs.get(key)
        if b is None or b.numel() < n:
            b = self.torch.empty(n, dtype=self.torch.float64, device=device)
            self._bufs[key] = b
        return b[:n]

    # which inputs of an operator are read through ghost rows: (space, "min" = the rows the element kernels read |
    # "ext" = every ghost row) for the field x, the coefficient c and the advecting velocity u1.  2-form operators are
    # element-local; M0 is diagonal (its weights are summed once, from the elements this rank holds around its nodes).
    NEEDS = {"M1": dict(x=(1, "min")), "M1h": dict(x=(1, "min"), c=(2, "min")), "M2": {}, "M2h": {}, "K": dict(x=(1, "min"), c=(1, "min")),
             "E21": dict(x=(1, "min")), "E12": dict(x=(2, "min")), "UtQW": dict(x=(2, "min"), c=(1, "min")),
             "M0": {}, "M0h": dict(c=(2, "ext")), "E10": dict(x=(0, "min")), "E01": dict(x=(1, "ext")),
             "R": dict(x=(1, "min"), c=(0, "min")), "R_up": dict(x=(1, "min"), c=(0, "min"), u1=(1, "min")),
             "M0h_up": dict(x=(0, "min"), c=(2, "ext"), u1=(1, "ext"))}

    def apply(self, op, x, coeff=None, out=None, exchange=True, flags=0, x_next=None, pipeline_last=False, **kw):
        """Ghost refresh of the inputs + local apply.  For the element kernels the refresh runs on a side stream
        while the interior elements (no ghost reads) are computed; boundary elements follow once it has landed.
        x_next (fused M1 only): software pipelining over independent applies -- this call pushes the boundary rows of
        x_next, the input of the NEXT call, and consumes what the previous call (or prologue_push) sent for x;
        pipeline_last=True ends such a sequence (consume only)."""
        if op not in self.SUPPORTED:
            raise NotImplementedError("operator %s" % op)
        torch = self.torch
        sin, sout, sc = self.engine.SPACES[op]
        need = self.NEEDS[op] if exchange else {}
        fields = dict(x=x, c=coeff, u1=kw.get("u1"))
        todo = [(fields[k], sp, how == "ext") for k, (sp, how) in need.items() if fields[k] is not None]
        do_x = "x" in need
        do_c = "c" in need and coeff is not None
        if out is None:
            out = self.engine.zeros(self.engine.space_sizes(op)[1], x.shape[1])
        if not todo:
            return self.engine.apply(op, x, coeff=coeff, out=out, flags=flags, **kw)
        if op not in ("M1", "M1h", "K"):
            # not on the hot path: refresh, then apply (E21 / E12 / UtQW and the 0-form family)
            for f, sp, ext in todo:
                self.exchange(f, sp, ext=ext)
            return self.engine.apply(op, x, coeff=coeff, out=out, flags=flags, **kw)
        if op == "M1" and self.p2p is not None and self.fused and self._fused_ok(x, kw):
            return self._apply_m1_fused(x, out, flags, x_next=x_next, mode=3 if pipeline_last else None, **kw)
        if x_next is not None or pipeline_last:
            raise NotImplementedError("pipelined ghost refresh needs the fused M1 path")
        overlap = self.overlap and self.n_interior > 0
        if overlap and self.p2p is not None:
            # side stream: push kernels (store into the peers' inboxes over NVLink) and pull kernels (wait for the
            # peers' flags, fill my ghost rows); main stream: interior elements meanwhile, boundary elements after
            main = torch.cuda.current_stream(self.device)
            side = self.comm_stream
            ready = torch.cuda.Event()
            ready.record(main)
            side.wait_event(ready)
            with torch.cuda.stream(side):
                if do_x:
                    self.push(x, sin)
                if do_c:
                    self.push(coeff, sc)
                if do_x:
                    self.pull(x, sin)
                if do_c:
                    self.pull(coeff, sc)
                done = torch.cuda.Event()
                done.record(side)
            self.engine.apply(op, x, coeff=coeff, out=out, flags=flags | SUBSET_INTERIOR, **kw)
            main.wait_event(done)
            self.engine.apply(op, x, coeff=coeff, out=out, flags=flags | SUBSET_BOUNDARY, **kw)
            return out
        if not overlap:
            if do_x:
                self.exchange(x, sin)
            if do_c:
                self.exchange(coeff, sc)
            return self.engine.apply(op, x, coeff=coeff, out=out, flags=flags, **kw)
        main = torch.cuda.current_stream(self.device)
        ready = torch.cuda.Event()
        ready.record(main)
        with torch.cuda.stream(self.comm_stream):
            self.comm_stream.wait_event(ready)
            if do_x:
                self.exchange(x, sin)
            if do_c:
                self.exchange(coeff, sc)
            done = torch.cuda.Event()
            done.record(self.comm_stream)
        self.engine.apply(op, x, coeff=coeff, out=out, flags=flags | SUBSET_INTERIOR, **kw)
        main.wait_event(done)
        self.engine.apply(op, x, coeff=coeff, out=out, flags=flags | SUBSET_BOUNDARY, **kw)
        return out

    def _fused_ok(self, x, kw):
        nlev = x.shape[1]
        return nlev % 2 == 0 and nlev <= 64 and nlev <= self.nk_max and self.engine.p >= 2

    def prologue_push(self, x):
        """Start a pipelined sequence: push the boundary rows of x (the first input) without applying anything.
        Collective; the caller must make sure every rank has finished it (stream sync + barrier) before the first
        pipelined apply if an earlier sequence already raised the flags of this epoch."""
        self._apply_m1_fused(x, x, 0, x_next=x, mode=2)

    def _apply_m1_fused(self, x, out, flags, lev0=0, scale=1.0, tpow=0, x_next=None, mode=None):
        """ONE launch: push CTAs + interior tiles + boundary tiles that stage their ghost rows from the inbox
        (mimsem_gpu_apply_M1_halo).  Shares the epoch / flag / ack words of the 1-form space with push()/pull():
        both epoch counters advance together so the two mechanisms can be mixed."""
        from .lib import check
        eng = self.engine
        npush, dpush, npull, dpull, epochs = self.p2p["plans"][1]
        inbox, stride, push_rows = self._inbox[1]
        nlev = x.shape[1]
        push_ctas = max(1, min(148, push_rows // 16))    # ~16 rows per CTA: one pass with four loads in flight per thread
        if self.push_ctas_cap:
            push_ctas = max(1, min(push_ctas, self.push_ctas_cap))
        if mode is None:
            mode = 0 if x_next is None else 1
        xp = x.data_ptr() if x_next is None else x_next.data_ptr()
        if self.ll:
            L = self._ll
            push_ctas = max(1, min(148, L["push_rows"] // 16))
            check(eng.L.mimsem_gpu_apply_M1_halo_ll(eng._h, lev0, nlev, nlev, scale, tpow, flags, x.data_ptr(), out.data_ptr(), xp, mode,
                                                    L["npush"], L["dpush"].data_ptr(), L["npull"], L["dpull"].data_ptr(), L["inbox"],
                                                    L["stride"], self.NBUF, push_ctas, L["epochs"].data_ptr(),
                                                    self.p2p["err"].data_ptr(), eng._stream()))
            return out
        check(eng.L.mimsem_gpu_apply_M1_halo(eng._h, lev0, nlev, nlev, scale, tpow, flags, x.data_ptr(), out.data_ptr(), xp, mode,
                                             npush, dpush.data_ptr(), npull, dpull.data_ptr(), inbox, stride, self.NBUF, push_ctas,
                                             epochs.data_ptr(), self.p2p["err"].data_ptr(), eng._stream()))
        return out

    def solve(self, op, b, out=None, lev0=0, scale=1.0, tpow=0, flags=0, rtol=1e-13, maxit=200):
        """x = M1^-1 b on the partitioned mesh (collective): Jacobi-PCG whose operator is the fused ghost-refresh + M1
        launch and whose dot products are completed over peer memory (mimsem_gpu_solve_M1_dist).  b, x: local fields
        (owned rows matter).  Returns (x, iterations, worst relative residual) -- identical on every rank."""
        import ctypes as C
        from .lib import check, MimsemError, HaloDesc, ReduceDesc
        if op != "M1":
            raise MimsemError("partitioned solve: M1 (M0 is diagonal: Engine.solve('M0') on the owned rows)")
        if self.p2p is None:
            raise MimsemError("the partitioned solve needs the peer-to-peer halo (MIMSEM_HALO=p2p)")
        eng = self.engine
        nlev = b.shape[1]
        self._check_levels(b)
        if out is None:
            out = eng.zeros(eng.n1, nlev)
        if self.ll:
            L = self._ll
            hd = HaloDesc(L["npush"], L["dpush"].data_ptr(), L["npull"], L["dpull"].data_ptr(), L["inbox"], L["stride"], self.NBUF,
                          max(1, min(148, L["push_rows"] // 16)), L["epochs"].data_ptr(), self.p2p["err"].data_ptr(), 1)
        else:
            npush, dpush, npull, dpull, epochs = self.p2p["plans"][1]
            inbox, stride, push_rows = self._inbox[1]
            hd = HaloDesc(npush, dpush.data_ptr(), npull, dpull.data_ptr(), inbox, stride, self.NBUF, max(1, min(148, push_rows // 16)),
                          epochs.data_ptr(), self.p2p["err"].data_ptr(), 0)
        rd = ReduceDesc(self.world, self.rank, self._red["areas"].data_ptr(), self._red["seq"].data_ptr(), self.p2p["err"].data_ptr())
        it = C.c_int(0)
        rr = C.c_double(0.0)
        check(eng.L.mimsem_gpu_solve_M1_dist(eng._h, lev0, nlev, nlev, scale, tpow, flags, b.data_ptr(), out.data_ptr(), rtol, maxit,
                                             C.byref(it), C.byref(rr), C.addressof(hd), C.addressof(rd), eng._stream()))
        return out, it.value, rr.value

    def capture(self, op, x, coeff=None, out=None, **kw):
        """Capture one full step (pack, NCCL ghost refresh, unpack, interior and boundary kernels) into a CUDA
        graph -- the launch-bound inner loop of a time step -- and return (replay, out).  All ranks must capture
        and replay the same sequence."""
        torch = self.torch
        if out is None:
            out = self.engine.zeros(self.engine.space_sizes(op)[1], x.shape[1])
        # warm up outside the capture (allocates the pack buffers, creates NCCL channels)
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for _ in range(2):
                self.apply(op, x, coeff=coeff, out=out, **kw)
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            self.apply(op, x, coeff=coeff, out=out, **kw)
        return graph.replay, out

    def capture_burst(self, op, xs, coeffs, outs, nsteps, pipelined=False, **kw):
        """Capture `nsteps` consecutive steps, step i on slot i % len(xs), into ONE CUDA graph and return its replay.  With
        the engine option "pdl_independent" the launches inside carry programmatic dependencies: a step starts while the
        previous one drains (the slots must then be independent fields).  pipelined (fused M1): step i also pushes the
        boundary rows of the NEXT step's input (the last step those of xs[0], so that replays chain); the caller starts the
        sequence with prologue_push(xs[0]) + a barrier.  All ranks must capture and replay alike."""
        torch = self.torch
        n = len(xs)
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for i in range(2 * n):
                self.apply(op, xs[i % n], coeff=None if coeffs is None else coeffs[i % n], out=outs[i % n], **kw)
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        # fused M1 launches of one graph form a burst: each works on its own epoch offset and counter slot, so that under
        # programmatic dependent launch they overlap completely (push of step i+1 during step i); see HaloFused
        as_burst = (op == "M1" and self.p2p is not None and self.fused and not self.ll and 1 < nsteps <= 32
                    and self._fused_ok(xs[0], kw))
        graph = torch.cuda.CUDAGraph()
        try:
            with torch.cuda.graph(graph):
                for i in range(nsteps):
                    if as_burst:
                        self.engine.set_option("halo_burst_len", nsteps)
                        self.engine.set_option("halo_burst_pos", i)
                    nxt = {}
                    if pipelined:
                        nxt = {"x_next": xs[(i + 1) % n if i < nsteps - 1 else 0]}
                    self.apply(op, xs[i % n], coeff=None if coeffs is None else coeffs[i % n], out=outs[i % n], **nxt, **kw)
        finally:
            self.engine.set_option("halo_burst_len", 0)
            self.engine.set_option("halo_burst_pos", 0)
        return graph.replay

    # test / IO helpers ------------------------------------------------------------------
    def scatter_from_global(self, levels_global, space):
        """numpy (nlev, N_space) global field -> local column tensor with owned AND ghost rows filled."""
        g = {0: self.part.g0, 1: self.part.g1, 2: self.part.g2}[space]
        loc = np.ascontiguousarray(levels_global[:, g])
        t = self.torch.from_numpy(loc).to("cuda:%d" % self.device)
        return self.engine.to_columns(t, space)

    def owned_to_global(self, field, space, out_global):
        """write the owned rows of a local column tensor into a numpy (nlev, N_space) global array"""
        loc = self.engine.to_levels(field, space).cpu().numpy()
        n_own = self.part.n_owned(space)
        out_global[:, self.part.owned_global(space)] = loc[:, :n_own]
        return out_global
