"""ORACLE-ONLY (test infrastructure): ctypes binding of oracle/_ref/libref_{eul,src,box}.so.

Those libraries are the reference's own, unmodified hot-path sources compiled against the serial
rank-emulating PETSc/MPI shim (oracle/Makefile, oracle/ref_driver.cpp).  This module is the "O1"
oracle of SURVEY.md section 8c: reference constructor + assemble(...) on every emulated rank,
merged to one CSR matrix, applied with a CSR SpMV (== PETSc MatMult up to summation order).

May be imported only from tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REFDIR = os.path.join(HERE, "_ref")

OPS = dict(Umat=0, Wmat=1, Pmat=2, Uhmat=3, WtQUmat=4, E10=5, E01=6, E21=7, E12=8, Pmat_h=9,
           Whmat=10, RotMat=11, Phmat_up=12, RotMat_up=13, Ut_mat=14, Ut_mat_h=15, UtQWmat=16, WtQdUdz_mat=17, Umat_ray=18,
           WtQmat=19, UtQmat=20, PtQmat=21)

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_lp = C.POINTER(C.c_long)


def _d(a):
    return a.ctypes.data_as(_dp) if a is not None else None


def available(variant="eul"):
    return os.path.exists(os.path.join(REFDIR, "libref_%s.so" % variant))


def mesh_dir(kind, p, ne, nprocs):
    return os.path.join(REFDIR, "meshes", "%s_p%d_ne%d_np%d" % (kind, p, ne, nprocs))


class Reference:
    """One emulated `mpirun -np nranks` instance of the reference (variant = 'eul' | 'src' | 'box')."""

    def __init__(self, variant, meshdir, nranks, nk=1, nthreads=None):
        self.variant = variant
        self.lib = C.CDLL(os.path.join(REFDIR, "libref_%s.so" % variant))
        L = self.lib
        L.ref_open.restype = C.c_void_p
        L.ref_open.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int]
        L.ref_close.argtypes = [C.c_void_p]
        L.ref_info.argtypes = [C.c_void_p, C.c_int, _ip]
        L.ref_get_loc.argtypes = [C.c_void_p, C.c_int, C.c_int, _ip]
        L.ref_get_geom.argtypes = [C.c_void_p, C.c_int, _dp, _dp]
        L.ref_get_coords.argtypes = [C.c_void_p, C.c_int, _dp]
        L.ref_geom_nl.argtypes = [C.c_void_p, C.c_int]
        L.ref_get_basis.argtypes = [C.c_void_p, _dp, _dp, _dp, _dp]
        L.ref_assemble.restype = C.c_long
        L.ref_assemble.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_int, _dp, _dp, _dp,
                                   C.c_double, C.c_double, C.c_int, _dp]
        L.ref_csr_shape.argtypes = [C.c_void_p, _lp]
        L.ref_csr_get.argtypes = [C.c_void_p, _lp, _ip, _dp]
        L.ref_spmv.restype = C.c_double
        L.ref_spmv.argtypes = [C.c_void_p, _dp, _dp, C.c_int]
        if variant != "src":
            L.ref_set_thick.argtypes = [C.c_void_p, C.c_int, _dp]
        if variant == "eul":
            L.ref_uvec_apply.restype = C.c_double
            L.ref_uvec_apply.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_int, _dp, _dp, C.c_int]
            L.ref_geom_interp.argtypes = [C.c_void_p, C.c_int, C.c_int, _dp, _dp]
            L.ref_init_topog.argtypes = [C.c_void_p, C.c_int, _dp]
            L.ref_uvec_assemble_hu.restype = C.c_double
            L.ref_uvec_assemble_hu.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_int, _dp, _dp, _dp, _dp, C.c_int]
        self.nthreads = nthreads or min(os.cpu_count() or 1, nranks)
        self.nranks = nranks
        self.nk = nk
        cwd = os.getcwd()
        try:
            self.h = L.ref_open(os.path.abspath(meshdir).encode(), nranks, nk, self.nthreads)
        finally:
            os.chdir(cwd)  # ref_open chdir()s into the mesh directory (the reference reads input/ relative to cwd)
        if not self.h:
            raise RuntimeError("ref_open failed for %s" % meshdir)
        info = np.zeros(16, dtype=np.int32)
        L.ref_info(self.h, 0, info.ctypes.data_as(_ip))
        (self.p, self.nelsx, self.m, self.n0, self.n1x, self.n1y, self.n2, _, _, _,
         self.N0, self.N1, self.N2, self.n0q) = [int(v) for v in info[:14]]
        self.secs = np.zeros(2)

    def close(self):
        if self.h:
            self.lib.ref_close(self.h)
            self.h = None

    def info(self, rank):
        info = np.zeros(16, dtype=np.int32)
        self.lib.ref_info(self.h, rank, info.ctypes.data_as(_ip))
        keys = "elOrd nElsX quadOrd n0 n1x n1y n2 n0l n1l n2l nDofs0G nDofs1G nDofs2G n0q".split()
        return dict(zip(keys, [int(v) for v in info[:14]]))

    def loc(self, rank, which):
        idx = dict(loc0=0, loc1x=1, loc1y=2, loc2=3, loc1=4, locq=5)[which]
        i = self.info(rank)
        n = dict(loc0=i["n0"], loc1x=i["n1x"], loc1y=i["n1y"], loc2=i["n2"], loc1=i["n1x"] + i["n1y"],
                 locq=i["n0q"])[which]
        out = np.zeros(n, dtype=np.int32)
        self.lib.ref_get_loc(self.h, rank, idx, out.ctypes.data_as(_ip))
        return out

    def geom(self, rank):
        nel = self.nelsx * self.nelsx
        q2 = (self.m + 1) ** 2
        det = np.zeros((nel, q2))
        J = np.zeros((nel, q2, 2, 2))
        self.lib.ref_get_geom(self.h, rank, _d(det), _d(J))
        return det, J

    def coords(self, rank):
        nl = self.lib.ref_geom_nl(self.h, rank)
        x = np.zeros((nl, 3))
        self.lib.ref_get_coords(self.h, rank, _d(x))
        return x

    def basis(self):
        m, n = self.m, self.p
        x = np.zeros(m + 1); w = np.zeros(m + 1)
        l = np.zeros((m + 1, n + 1)); e = np.zeros((m + 1, n))
        self.lib.ref_get_basis(self.h, _d(x), _d(w), _d(l), _d(e))
        return x, w, l, e

    def set_thick(self, rank, thick):
        """thick[nk][n0q] in the rank's local quadrature-point numbering (eul) / node numbering (box)."""
        thick = np.ascontiguousarray(thick, dtype=np.float64)
        assert thick.shape == (self.nk, self.n0q), (thick.shape, self.nk, self.n0q)
        self.lib.ref_set_thick(self.h, rank, _d(thick))

    def assemble(self, op, lev=0, scale=1.0, flag=True, c2=None, c1=None, c0=None, tau=0.0, dt=0.0):
        """Reference ctor + assemble on all ranks, merged; returns scipy.sparse.csr_matrix."""
        import scipy.sparse as sp
        c2 = None if c2 is None else np.ascontiguousarray(c2, dtype=np.float64)
        c1 = None if c1 is None else np.ascontiguousarray(c1, dtype=np.float64)
        c0 = None if c0 is None else np.ascontiguousarray(c0, dtype=np.float64)
        nnz = self.lib.ref_assemble(self.h, OPS[op], lev, scale, int(bool(flag)), _d(c2), _d(c1), _d(c0),
                                    tau, dt, self.nthreads, _d(self.secs))
        if nnz < 0:
            raise ValueError("operator %s not available in variant %s" % (op, self.variant))
        shp = np.zeros(3, dtype=np.int64)
        self.lib.ref_csr_shape(self.h, shp.ctypes.data_as(_lp))
        indptr = np.zeros(shp[0] + 1, dtype=np.int64)
        indices = np.zeros(shp[2], dtype=np.int32)
        data = np.zeros(shp[2])
        self.lib.ref_csr_get(self.h, indptr.ctypes.data_as(_lp), indices.ctypes.data_as(_ip), _d(data))
        return sp.csr_matrix((data, indices, indptr), shape=(int(shp[0]), int(shp[1])))

    def assemble_only(self, op, lev=0, scale=1.0, flag=True, c2=None, c1=None, c0=None):
        """Like assemble() but leaves the CSR inside the library (for timing); returns (assemble_s, merge_s)."""
        nnz = self.lib.ref_assemble(self.h, OPS[op], lev, scale, int(bool(flag)), _d(c2), _d(c1), _d(c0),
                                    0.0, 0.0, self.nthreads, _d(self.secs))
        if nnz < 0:
            raise ValueError(op)
        return float(self.secs[0]), float(self.secs[1])

    def spmv(self, x, nthreads=None):
        """y = A x with the library-resident CSR of the last assemble; returns (y, seconds)."""
        shp = np.zeros(3, dtype=np.int64)
        self.lib.ref_csr_shape(self.h, shp.ctypes.data_as(_lp))
        x = np.ascontiguousarray(x, dtype=np.float64)
        assert x.shape[0] == shp[1]
        y = np.zeros(shp[0])
        s = self.lib.ref_spmv(self.h, _d(x), _d(y), nthreads or self.nthreads)
        return y, s

    def uvec_apply(self, x, lev=0, scale=1.0, vert_scale=True):
        """The reference's own matrix-free M1*x (Uvec::assemble + reverse ADD scatter); eul only."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.zeros(self.N1)
        s = self.lib.ref_uvec_apply(self.h, lev, scale, int(bool(vert_scale)), _d(x), _d(y), self.nthreads)
        return y, s

    def uvec_assemble_hu(self, us, hs, facs, lev=0, scale=1.0):
        """The reference's own Uvec::assemble_hu accumulated over (velocity, density, factor) triples and scattered
        (eul/Assembly.cpp:2198-2279, driven as eul/HorizSolve.cpp:298-306 does); eul only.  Returns the global 1-form."""
        us = np.ascontiguousarray(us, dtype=np.float64)
        hs = np.ascontiguousarray(hs, dtype=np.float64)
        facs = np.ascontiguousarray(facs, dtype=np.float64)
        assert us.shape == (len(facs), self.N1) and hs.shape == (len(facs), self.N2)
        y = np.zeros(self.N1)
        self.lib.ref_uvec_assemble_hu(self.h, lev, scale, len(facs), _d(us), _d(hs), _d(facs), _d(y), self.nthreads)
        return y

    def geom_interp(self, rank, which, vec):
        """The reference's Geom::interp0 / interp1_l / interp2_l / interp1_g / interp2_g (which = 0..4) of one rank at all
        quadrature points (eul/Geom.cpp:328-417); eul only."""
        vec = np.ascontiguousarray(vec, dtype=np.float64)
        nel, q2 = self.nelsx ** 2, (self.m + 1) ** 2
        out = np.zeros((nel, q2, 2) if which in (1, 3) else (nel, q2))
        self.lib.ref_geom_interp(self.h, rank, which, _d(vec), _d(out))
        return out

    def init_topog(self, rank):
        """Geom::initTopog with the UMJS14 level function over a smooth hill (see oracle/ref_driver.cpp); returns thick[nk][n0q]."""
        out = np.zeros((self.nk, self.n0q))
        self.lib.ref_init_topog(self.h, rank, _d(out))
        return out
