"""Host-side wrappers: basis tables, reference-rank topology maps, the canonical global mesh."""
import ctypes as C

import numpy as np

from .lib import load_library, check, _dp, _ip, _lp, _vp

SPHERE, BOX = 0, 1
_KIND = {"sphere": SPHERE, "box": BOX, SPHERE: SPHERE, BOX: BOX}


def _p(a, t):
    return a.ctypes.data_as(t)


class Basis:
    """GaussLobatto / LagrangeNode / LagrangeEdge tabulations (reference: eul/Basis.cpp)."""

    def __init__(self, p, m=None):
        L = load_library()
        m = p if m is None else m
        self.p, self.m = p, m
        self.x = np.zeros(m + 1)
        self.w = np.zeros(m + 1)
        check(L.mimsem_basis_gll(m, _p(self.x, _dp), _p(self.w, _dp)))
        self.ljxi = np.zeros((m + 1, p + 1))
        self.ejxi = np.zeros((m + 1, p))
        check(L.mimsem_basis_tables(p, m, _p(self.ljxi, _dp), _p(self.ejxi, _dp)))

    def elmat(self, which):
        """ElMats tabulation: 'U','V','W','P' as (quad point, dof) matrices, 'Q' the Wii diagonal."""
        p, m = self.p, self.m
        q2 = (m + 1) ** 2
        idx = "UVWPQ".index(which)
        ncol = [p * (p + 1), p * (p + 1), p * p, (p + 1) ** 2, 1][idx]
        A = np.zeros((q2, ncol))
        check(load_library().mimsem_elmat(idx, p, m, _p(A, _dp)))
        return A[:, 0].copy() if which == "Q" else A


def patch_topology(kind, order, ne, nprocs, rank):
    """loc0, loc1x, loc1y, loc2 and local_sizes of one reference MPI rank (scr/Proc2.py, scr/ProcBox.py)."""
    L = load_library()
    k = _KIND[kind]
    sz = np.zeros(8, dtype=np.int32)
    check(L.mimsem_topo_patch_sizes(k, order, ne, nprocs, rank, _p(sz, _ip)))
    loc0 = np.zeros(sz[0], dtype=np.int32)
    loc1x = np.zeros(sz[1], dtype=np.int32)
    loc1y = np.zeros(sz[2], dtype=np.int32)
    loc2 = np.zeros(sz[3], dtype=np.int32)
    check(L.mimsem_topo_patch(k, order, ne, nprocs, rank, _p(loc0, _ip), _p(loc1x, _ip), _p(loc1y, _ip), _p(loc2, _ip)))
    return dict(loc0=loc0, loc1x=loc1x, loc1y=loc1y, loc2=loc2, local_sizes=sz[4:8].copy())


def write_input(kind, p, ne, nprocs, directory, m=None):
    """Write the reference's input/*.txt set into `directory` (scr/Setup.py formats)."""
    check(load_library().mimsem_topo_write_input(_KIND[kind], p, p if m is None else m, ne, nprocs, directory.encode()))


class Mesh:
    """Canonical global mesh: element -> DOF tables, Jacobians, quadrature-point coordinates."""

    def __init__(self, kind, p, ne, m=None, signed_det=False):
        L = load_library()
        self.kind = _KIND[kind]
        h = _vp()
        check(L.mimsem_mesh_create(self.kind, p, p if m is None else m, ne, int(signed_det), C.byref(h)))
        self._h = h
        sz = np.zeros(8, dtype=np.int64)
        check(L.mimsem_mesh_sizes(h, _p(sz, _lp)))
        self.p, self.m, self.ne, self.nel, self.N0, self.N1, self.N2, self.NQ = [int(v) for v in sz]
        p, m = self.p, self.m
        self.el0 = np.zeros((self.nel, (p + 1) ** 2), dtype=np.int32)
        self.el1x = np.zeros((self.nel, p * (p + 1)), dtype=np.int32)
        self.el1y = np.zeros((self.nel, p * (p + 1)), dtype=np.int32)
        self.el2 = np.zeros((self.nel, p * p), dtype=np.int32)
        self.elq = np.zeros((self.nel, (m + 1) ** 2), dtype=np.int32)
        check(L.mimsem_mesh_tables(h, _p(self.el0, _ip), _p(self.el1x, _ip), _p(self.el1y, _ip), _p(self.el2, _ip),
                                   _p(self.elq, _ip)))
        self.J = np.zeros((self.nel, (m + 1) ** 2, 4))
        self.det = np.zeros((self.nel, (m + 1) ** 2))
        check(L.mimsem_mesh_geometry(h, _p(self.J, _dp), _p(self.det, _dp)))
        self.xyz = np.zeros((self.NQ, 3))
        check(L.mimsem_mesh_coords(h, _p(self.xyz, _dp)))

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                load_library().mimsem_mesh_destroy(self._h)
                self._h = None
        except Exception:
            pass
