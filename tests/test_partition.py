"""CPU tests of the N>1 host logic: element-block partition, owned-first local numbering, ghost lists,
and the ghost refresh itself over a world_size-2 gloo group (no GPU, no CUDA library calls)."""
import os
import socket

import numpy as np
import pytest

import mimsem_b200 as mb
from mimsem_b200.parallel import Partition, element_range, owner_rank_of_element, send_lists, west_south_neighbours


@pytest.mark.parametrize("kind,p,ne,world", [("sphere", 3, 4, 2), ("sphere", 3, 4, 8), ("sphere", 4, 2, 3), ("box", 3, 4, 2), ("box", 3, 4, 4)])
def test_partition_covers_everything_once(kind, p, ne, world):
    mesh = mb.Mesh(kind, p, ne)
    seen1 = np.zeros(mesh.N1, int)
    seen2 = np.zeros(mesh.N2, int)
    for r in range(world):
        P = Partition(mesh, r, world)
        e0, e1 = element_range(mesh.nel, r, world)
        assert np.array_equal(np.sort(P.elements[:P.nel_owned]), np.arange(e0, e1))   # interior first, boundary last
        assert np.all(owner_rank_of_element(np.arange(e0, e1), mesh.nel, world) == r)
        seen1[P.g1[:P.n1_owned]] += 1
        seen2[P.g2[:P.n2_owned]] += 1
        # local tables reproduce the global ones
        assert np.array_equal(P.g1[P.el1x], mesh.el1x[P.elements])
        assert np.array_equal(P.g1[P.el1y], mesh.el1y[P.elements])
        assert np.array_equal(P.g2[P.el2], mesh.el2[P.elements])
        assert np.array_equal(P.gq[P.elq], mesh.elq[P.elements])
        # every west / south neighbour of an owned element is local (owner-computes needs no output reduction)
        nb = west_south_neighbours(mesh)[e0:e1]
        assert np.all(np.isin(nb[nb >= 0], P.elements))
        # ghosts are owned by the rank the plan says
        for space, block in ((1, 2 * p * p), (2, p * p)):
            for q, grp in P.recv[space].items():
                assert q != r
                assert np.all(owner_rank_of_element(grp["glob"] // block, mesh.nel, world) == q)
    assert np.all(seen1 == 1) and np.all(seen2 == 1)


@pytest.mark.parametrize("kind,p,ne,world", [("sphere", 3, 4, 2), ("sphere", 3, 4, 8), ("sphere", 4, 2, 3), ("box", 3, 4, 4)])
def test_exchanged_rows_cover_everything_the_kernels_read(kind, p, ne, world):
    """Only the ghost rows in [n1_owned, n1_halo) are refreshed.  Re-derive, from the mesh tables alone, every 1-form row
    the M1 tile plan of an owned element stages (own edges, east column, north row, and the other edge family of the west /
    south neighbour across its far line) and check that each is owned or refreshed; and that the interior elements,
    stored first, read no foreign row at all."""
    mesh = mb.Mesh(kind, p, ne)
    ws = west_south_neighbours(mesh)
    b1, b2 = 2 * p * p, p * p
    for r in range(world):
        P = Partition(mesh, r, world)
        refreshed = set(P.g1[:P.n1_halo].tolist())
        owned2 = set(P.g2[:P.n2_owned].tolist()) | {int(g) for q in P.recv[2].values() for g in q["glob"]}
        for le in range(P.nel_owned):
            e = int(P.elements[le])
            reads = set(mesh.el1x[e].tolist()) | set(mesh.el1y[e].tolist())
            faces = set(mesh.el2[e].tolist())
            for side, shared in ((0, int(mesh.el1x[e, 0])), (1, int(mesh.el1y[e, 0]))):
                n = int(ws[e, side])
                if n < 0:
                    continue
                reads |= set(mesh.el1y[n].tolist()) if shared in set(mesh.el1x[n].tolist()) else set(mesh.el1x[n].tolist())
                faces |= set(mesh.el2[n].tolist())          # M1(h): the neighbour's coefficient block
            assert reads <= refreshed, (r, e, sorted(reads - refreshed)[:5])
            assert faces <= owned2
            if le < P.n_interior:
                assert all(P.e0 <= d // b1 < P.e1 for d in reads) and all(P.e0 <= f // b2 < P.e1 for f in faces)
        # the exchange is smaller than "every edge of every halo element"
        assert P.n1_halo <= P.n1


def test_send_lists_match_receive_lists():
    mesh = mb.Mesh("sphere", 3, 4)
    world = 4
    parts, sends = zip(*[send_lists(mesh, r, world) for r in range(world)])
    for r in range(world):
        for space in (1, 2):
            for q, grp in parts[r].recv[space].items():
                s = sends[q][space][r]                       # what q sends to r, as q-local owned ids
                assert np.array_equal(parts[q].owned_global(space)[s], grp["glob"])


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mesh = mb.Mesh("sphere", 3, 4)
        part, sends = send_lists(mesh, rank, world)
        rng = np.random.default_rng(0)
        ok = True
        # the rows the element kernels read (spaces 0, 1, 2) and, for the node-sum operators, ALL ghost rows (ext)
        for space, N, ext in ((0, mesh.N0, False), (1, mesh.N1, False), (2, mesh.N2, False), (1, mesh.N1, True), (2, mesh.N2, True)):
            glob = rng.uniform(-1, 1, (N, 3))                 # same on every rank
            g = {0: part.g0, 1: part.g1, 2: part.g2}[space]
            n_own = part.n_owned(space)
            recv = (part.recv_ext if ext else part.recv)[space]
            send = (part.sends_ext if ext else sends)[space]
            loc = torch.zeros((len(g), 3), dtype=torch.float64)
            loc[:n_own] = torch.from_numpy(glob[g[:n_own]])   # ghosts start as zero
            ops, unpack = [], []
            for peer in sorted(set(recv) | set(send)):
                if peer in send:
                    sb = loc[torch.from_numpy(send[peer].astype(np.int64))].contiguous()
                    ops.append(dist.P2POp(dist.isend, sb, peer))
                if peer in recv:
                    rows = torch.from_numpy(recv[peer]["local"].astype(np.int64))
                    rb = torch.empty((len(rows), 3), dtype=torch.float64)
                    ops.append(dist.P2POp(dist.irecv, rb, peer))
                    unpack.append((rows, rb))
            for w in dist.batch_isend_irecv(ops):
                w.wait()
            for rows, rb in unpack:
                loc[rows] = rb
            # minimal plans fill the rows some element kernel reads (rows beyond n*_halo belong to halo elements but are
            # only read by the node-sum operators); the ext plans and the node plan fill every ghost row
            n_filled = len(g) if (ext or space == 0) else {1: part.n1_halo, 2: part.n2_halo}[space]
            ok = ok and bool(np.array_equal(loc.numpy()[:n_filled], glob[g][:n_filled]))
            ok = ok and (ext or space != 1 or n_filled < len(g))
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


def test_ghost_refresh_gloo_world2():
    """The exchange plan, executed with gloo send/recv on CPU tensors, fills every ghost row with the owner's value."""
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    res = [q.get(timeout=120) for _ in procs]
    for pr in procs:
        pr.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


@pytest.mark.parametrize("kind,p,ne,world", [("sphere", 3, 4, 2), ("sphere", 4, 6, 4), ("sphere", 3, 6, 8), ("box", 3, 6, 3)])
def test_cpp_partition_equals_python_partition(kind, p, ne, world):
    """The C++ host layer (mimsem_b200/host/Partition.cpp, libmimsem_host.so) builds the same subdomains, numberings and
    ghost / send lists as parallel.py, array for array."""
    import ctypes as C
    import os
    import mimsem_b200 as mb
    from mimsem_b200.parallel import send_lists
    lib = C.CDLL(os.path.join(os.path.dirname(mb.__file__), "libmimsem_host.so"))
    lib.mimsem_host_partition_create.argtypes = [C.c_int] * 5 + [C.POINTER(C.c_void_p)]
    lib.mimsem_host_partition_array.restype = C.c_int64
    lib.mimsem_host_partition_array.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p]
    lib.mimsem_host_partition_sizes.argtypes = [C.c_void_p, C.c_void_p]
    lib.mimsem_host_partition_destroy.argtypes = [C.c_void_p]
    mesh = mb.Mesh(kind, p, ne)

    def arr(h, name):
        n = lib.mimsem_host_partition_array(h, name.encode(), None)
        assert n >= 0, name
        out = np.zeros(n, dtype=np.int64)
        lib.mimsem_host_partition_array(h, name.encode(), out.ctypes.data_as(C.c_void_p))
        return out
    for rank in range(world):
        P, sends = send_lists(mesh, rank, world)
        h = C.c_void_p()
        assert lib.mimsem_host_partition_create(0 if kind == "sphere" else 1, p, ne, rank, world, C.byref(h)) == 0
        sz = np.zeros(12, dtype=np.int64)
        lib.mimsem_host_partition_sizes(h, sz.ctypes.data_as(C.c_void_p))
        assert list(sz) == [P.nel_owned, P.nel_total, P.n_interior, P.n0, P.n1, P.n2, P.nq, P.n0_owned, P.n1_owned, P.n2_owned,
                            P.n1_halo, P.n2_halo]
        for name in ("elements", "g0", "g1", "g2", "gq", "el0", "el1x", "el1y", "el2", "elq"):
            assert np.array_equal(arr(h, name), np.asarray(getattr(P, name)).ravel()), (rank, name)
        for space in (0, 1, 2):
            for q in range(world):
                r = P.recv[space].get(q)
                assert np.array_equal(arr(h, "recv%d_%d" % (space, q)), r["local"] if r is not None else np.zeros(0)), (rank, space, q)
                s = sends[space].get(q)
                assert np.array_equal(arr(h, "send%d_%d" % (space, q)), s if s is not None else np.zeros(0)), (rank, space, q)
        lib.mimsem_host_partition_destroy(h)


def test_cpp_file_rendezvous_three_processes(tmp_path):
    """The control-plane transport of the C++ multi-GPU host layer between plain processes of one box (FileComm in
    mimsem_b200/host/DistEngine.cpp: allgather and barrier through files): three processes, forty rounds of changing size,
    every contribution verified by every rank (mimsem_b200/host/filecomm_check.cpp).  No GPU."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    host = os.path.join(root, "mimsem_b200", "host")
    subprocess.run(["make", "-C", host], check=True, stdout=subprocess.DEVNULL)
    exe = os.path.join(host, "build", "filecomm_check")
    procs = [subprocess.Popen([exe, str(tmp_path / "rdv")], env=dict(os.environ, MIMSEM_RANK=str(r), MIMSEM_WORLD="3"),
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(3)]
    outs = [p.communicate(timeout=120)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs) and all("ok" in o for o in outs), outs
