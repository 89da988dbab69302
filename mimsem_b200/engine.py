"""Engine: ctypes wrapper of the mimsem_gpu_* device engine.  Device memory and streams come from
torch (plumbing only); every numerical step is a kernel of libmimsem_gpu.so."""
import ctypes as C

import numpy as np

from .lib import load_library, check, MimsemError, _dp, _ip, _lp, _vp
from .mesh import Basis

OPS = dict(M1=0, M2=1, M0=2, M1h=3, K=4, M2h=5, M0h=6, E10=10, E01=11, E21=12, E12=13)
FIXED_LEVEL = 1
THICK_MEAN = 8       # apply_M1: thickness factor = mean thickness of levels lev, lev+1 (Ut_mat::assemble)
SUBSET_INTERIOR = 2
SUBSET_BOUNDARY = 4


def _np_ptr(a, t):
    return a.ctypes.data_as(t)


class Engine:
    """One subdomain on one GPU.

    Fields are torch float64 CUDA tensors in column layout, shape (ndof, nlev), contiguous
    (element [dof, k] at dof*ld + k): see include/mimsem_gpu.h.
    """

    def __init__(self, device=0):
        import torch
        if not torch.cuda.is_available():
            raise MimsemError("no CUDA device: mimsem_b200 has no CPU fallback")
        self.torch = torch
        self.L = load_library()
        self.device = device
        h = _vp()
        check(self.L.mimsem_gpu_create(device, C.byref(h)))
        self._h = h
        self.nk = 0

    @classmethod
    def from_mesh(cls, mesh, device=0, thick=None):
        """Whole global mesh as a single subdomain (owner-computes, no halo)."""
        eng = cls(device)
        b = Basis(mesh.p, mesh.m)
        eng.set_basis(b)
        eng.set_topo(mesh.el0, mesh.el1x, mesh.el1y, mesh.el2, mesh.elq, mesh.N0, mesh.N1, mesh.N2, mesh.NQ)
        eng.set_geom(mesh.J, mesh.det)
        if thick is not None:
            eng.set_thickness(thick)
        return eng

    def set_option(self, name, value):
        """tuning / test knob of this context (include/mimsem_gpu.h: mimsem_gpu_set_option)"""
        check(self.L.mimsem_gpu_set_option(self._h, name.encode(), int(value)))

    def close(self):
        if getattr(self, "_h", None):
            self.L.mimsem_gpu_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---------------------------------------------------------------- setup
    def set_basis(self, basis):
        self.p, self.m = basis.p, basis.m
        w = np.ascontiguousarray(basis.w)
        l = np.ascontiguousarray(basis.ljxi)
        e = np.ascontiguousarray(basis.ejxi)
        check(self.L.mimsem_gpu_set_basis(self._h, basis.p, basis.m, _np_ptr(w, _dp), _np_ptr(l, _dp), _np_ptr(e, _dp)))

    def set_topo(self, el0, el1x, el1y, el2, elq, n0, n1, n2, nq, nel_owned=None, mode=0):
        arrs = [np.ascontiguousarray(a, dtype=np.int32) for a in (el0, el1x, el1y, el2, elq)]
        nel_total = arrs[0].shape[0]
        nel_owned = nel_total if nel_owned is None else nel_owned
        check(self.L.mimsem_gpu_set_topo(self._h, nel_total, nel_owned, n0, n1, n2, nq, mode,
                                         *[_np_ptr(a, _ip) for a in arrs]))
        self.n0, self.n1, self.n2, self.nq = n0, n1, n2, nq
        self.nel_total, self.nel_owned = nel_total, nel_owned

    def set_ghosts(self, n1_owned, n2_owned):
        """rows >= n*_owned (caller numbering) are ghosts; returns (n_interior, n_boundary) owned elements."""
        cnt = np.zeros(2, dtype=np.int32)
        check(self.L.mimsem_gpu_set_ghosts(self._h, n1_owned, n2_owned, _np_ptr(cnt, _ip)))
        return int(cnt[0]), int(cnt[1])

    def set_geom(self, J, det):
        J = np.ascontiguousarray(J, dtype=np.float64)
        det = np.ascontiguousarray(det, dtype=np.float64)
        assert J.size == self.nel_total * (self.m + 1) ** 2 * 4 and det.size == self.nel_total * (self.m + 1) ** 2
        check(self.L.mimsem_gpu_set_geom(self._h, _np_ptr(J, _dp), _np_ptr(det, _dp)))

    def set_thickness(self, thick):
        """thick[nk][nq] (level-major, as the reference's Geom::thick)."""
        if thick is None:
            check(self.L.mimsem_gpu_set_thickness(self._h, 0, None))
            self.nk = 0
            return
        thick = np.ascontiguousarray(thick, dtype=np.float64)
        assert thick.ndim == 2 and thick.shape[1] == self.nq, thick.shape
        check(self.L.mimsem_gpu_set_thickness(self._h, thick.shape[0], _np_ptr(thick, _dp)))
        self.nk = thick.shape[0]

    # ---------------------------------------------------------------- helpers
    def space_sizes(self, op):
        n0, n1, n2 = self.n0, self.n1, self.n2
        if op in ("R", "R_up"):
            return (n1, n1, n0)
        if op == "M0h_up":
            return (n0, n0, n2)
        if op == "UtQW":
            return (n2, n1, n1)
        return {"M1": (n1, n1, 0), "M1h": (n1, n1, n2), "M2": (n2, n2, 0), "M2h": (n2, n2, n2), "M0": (n0, n0, 0),
                "M0h": (n0, n0, n2), "K": (n1, n2, n1), "E10": (n0, n1, 0), "E01": (n1, n0, 0), "E21": (n1, n2, 0),
                "E12": (n2, n1, 0)}[op]

    def _stream(self):
        return self.torch.cuda.current_stream(self.device).cuda_stream

    def empty(self, n, nlev):
        return self.torch.empty((n, nlev), dtype=self.torch.float64, device="cuda:%d" % self.device)

    def zeros(self, n, nlev):
        return self.torch.zeros((n, nlev), dtype=self.torch.float64, device="cuda:%d" % self.device)

    def _chk(self, t, n, nlev, name):
        if t.dtype != self.torch.float64 or not t.is_cuda or not t.is_contiguous() or tuple(t.shape) != (n, nlev):
            raise MimsemError("%s must be a contiguous float64 CUDA tensor of shape (%d, %d), got %s %s"
                              % (name, n, nlev, tuple(t.shape), t.dtype))

    def to_columns(self, levels, space):
        """(nlev, n) per-level device tensor in the caller's DOF numbering -> (n, nlev) engine column layout.
        space = 0, 1, 2 for nodes / edges / faces (-1: plain transpose)."""
        nlev, n = levels.shape
        out = self.empty(n, nlev)
        check(self.L.mimsem_gpu_levels_to_columns(self._h, space, n, nlev, nlev, levels.data_ptr(), out.data_ptr(),
                                                  self._stream()))
        return out

    def to_levels(self, cols, space):
        n, nlev = cols.shape
        out = self.torch.empty((nlev, n), dtype=self.torch.float64, device=cols.device)
        check(self.L.mimsem_gpu_columns_to_levels(self._h, space, n, nlev, nlev, cols.data_ptr(), out.data_ptr(),
                                                  self._stream()))
        return out

    def apply_ray(self, x, exner, exner_s, dt, out=None, lev0=0, scale=1.0):
        """Umat_ray::assemble(lev, scale, dt, exner_k, exner_s) + MatMult for all levels (eul/Assembly.cpp:1846-1979):
        exner = the Exner-pressure 2-form in column layout (n2, nlev), exner_s = its level-0 column, one value per face (n2,)."""
        nlev = x.shape[1]
        self._chk(x, self.n1, nlev, "x")
        self._chk(exner, self.n2, nlev, "exner")
        if exner_s.dtype != self.torch.float64 or exner_s.numel() != self.n2 or not exner_s.is_contiguous():
            raise MimsemError("exner_s: one float64 per face")
        if out is None:
            out = self.zeros(self.n1, nlev)
        check(self.L.mimsem_gpu_apply_M1ray(self._h, lev0, nlev, nlev, scale, dt, exner.data_ptr(), exner_s.data_ptr(), x.data_ptr(),
                                            out.data_ptr(), self._stream()))
        return out

    def permutation(self, space):
        """row of each caller-numbered DOF in the engine's column layout."""
        n = (self.n0, self.n1, self.n2)[space]
        perm = np.zeros(n, dtype=np.int32)
        check(self.L.mimsem_gpu_form_permutation(self._h, space, _np_ptr(perm, _ip)))
        return perm

    SPACES = {"M1": (1, 1, None), "M1h": (1, 1, 2), "M2": (2, 2, None), "M2h": (2, 2, 2), "M0": (0, 0, None),
              "M0h": (0, 0, 2), "K": (1, 2, 1), "E10": (0, 1, None), "E01": (1, 0, None), "E21": (1, 2, None),
              "E12": (2, 1, None), "R": (1, 1, 0), "R_up": (1, 1, 0), "M0h_up": (0, 0, 2), "UtQW": (2, 1, 1)}

    # ---------------------------------------------------------------- applies (device resident)
    def apply(self, op, x, coeff=None, out=None, lev0=0, scale=1.0, tpow=0, flags=0, u1=None, tau=0.0):
        """R / R_up / M0h_up (RotMat, RotMat_up, Phmat::assemble_up): coeff = q0 (0-form) resp. h2 (2-form),
        u1 = advecting 1-form velocity, tau = fac*dt."""
        nin, nout, ncoef = self.space_sizes(op)
        nlev = x.shape[1]
        self._chk(x, nin, nlev, "x")
        if out is None:
            # rows a subdomain does not own are not written by the kernels
            out = self.zeros(nout, nlev) if self.nel_owned != self.nel_total else self.empty(nout, nlev)
        self._chk(out, nout, nlev, "out")
        if ncoef:
            if coeff is None:
                raise MimsemError("operator %s needs a coefficient field" % op)
            self._chk(coeff, ncoef, nlev, "coeff")
        st = self._stream()
        h, L = self._h, self.L
        xp, yp = x.data_ptr(), out.data_ptr()
        if op in ("R_up", "M0h_up"):
            if u1 is None:
                raise MimsemError("operator %s needs the advecting velocity u1" % op)
            self._chk(u1, self.n1, nlev, "u1")
            fn = getattr(L, "mimsem_gpu_apply_" + op)
            check(fn(h, lev0, nlev, nlev, scale, tpow, flags, coeff.data_ptr(), u1.data_ptr(), tau, xp, yp, st))
        elif op == "UtQW":
            # UtQWmat::assemble(u1, scale) + MatMult: coeff = u1 (1-form), x = 2-form
            check(L.mimsem_gpu_apply_UtQW(h, nlev, nlev, scale, coeff.data_ptr(), xp, yp, st))
        elif op == "R":
            check(L.mimsem_gpu_apply_R(h, lev0, nlev, nlev, scale, tpow, flags, coeff.data_ptr(), xp, yp, st))
        elif op in ("M1", "M2", "M0"):
            fn = getattr(L, "mimsem_gpu_apply_" + op)
            check(fn(h, lev0, nlev, nlev, scale, tpow, flags, xp, yp, st))
        elif op in ("M1h", "M2h", "M0h", "K"):
            fn = getattr(L, "mimsem_gpu_apply_" + op)
            check(fn(h, lev0, nlev, nlev, scale, tpow, flags, coeff.data_ptr(), xp, yp, st))
        else:
            check(L.mimsem_gpu_apply_incidence(h, OPS[op] - 10, nlev, nlev, xp, yp, st))
        return out

    # ---------------------------------------------------------------- mass-matrix solves
    def solve(self, op, b, out=None, lev0=0, scale=1.0, tpow=0, flags=0, rtol=1e-13, maxit=200):
        """x = M^-1 b for op in ("M1", "M0") with the matrix that apply(op, ..., same arguments) applies.
        Returns (x, iterations, worst relative residual); M0 is diagonal (0 iterations)."""
        n = self.n1 if op == "M1" else self.n0
        nlev = b.shape[1]
        self._chk(b, n, nlev, "b")
        if out is None:
            out = self.empty(n, nlev)
        self._chk(out, n, nlev, "out")
        if op in ("M2", "M2h"):
            # WmatInv / WhmatInv: element-local dense solves
            raise MimsemError("use solve_m2")
        if op == "M0":
            check(self.L.mimsem_gpu_solve_M0(self._h, lev0, nlev, nlev, scale, tpow, flags, b.data_ptr(), out.data_ptr(), self._stream()))
            return out, 0, 0.0
        if op != "M1":
            raise MimsemError("solve: operator must be M1 or M0")
        it = C.c_int(0)
        rr = C.c_double(0.0)
        check(self.L.mimsem_gpu_solve_M1(self._h, lev0, nlev, nlev, scale, tpow, flags, b.data_ptr(), out.data_ptr(), rtol, maxit,
                                         C.byref(it), C.byref(rr), self._stream()))
        return out, it.value, rr.value

    def diag(self, op, nlev, lev0=0, scale=1.0, tpow=0, flags=0, coeff=None):
        """diagonal of M1, or of M0 / M0(h) (= Pvec::assemble / Phvec::assemble: the lumped 0-form mass vector)."""
        if op in ("M0", "M0h"):
            out = self.zeros(self.n0, nlev)
            if op == "M0h":
                self._chk(coeff, self.n2, nlev, "coeff")
            check(self.L.mimsem_gpu_diag_M0(self._h, lev0, nlev, nlev, scale, tpow, flags, None if op == "M0" else coeff.data_ptr(),
                                            out.data_ptr(), self._stream()))
            return out
        if op != "M1":
            raise MimsemError("diag: M1, M0 or M0h")
        out = self.zeros(self.n1, nlev)
        check(self.L.mimsem_gpu_diag_M1(self._h, lev0, nlev, nlev, scale, tpow, flags, out.data_ptr(), self._stream()))
        return out

    def solve_m2(self, b, coeff=None, out=None, lev0=0, scale=1.0, tpow=0, flags=0):
        """x = M2^-1 b (coeff: M2(rho)^-1 b) -- WmatInv / WhmatInv::assemble + MatMult (eul/Assembly.cpp:1658-1800)."""
        nlev = b.shape[1]
        self._chk(b, self.n2, nlev, "b")
        if out is None:
            out = self.zeros(self.n2, nlev) if self.nel_owned != self.nel_total else self.empty(self.n2, nlev)
        if coeff is not None:
            self._chk(coeff, self.n2, nlev, "coeff")
        check(self.L.mimsem_gpu_solve_M2(self._h, lev0, nlev, nlev, scale, tpow, flags, None if coeff is None else coeff.data_ptr(),
                                         b.data_ptr(), out.data_ptr(), self._stream()))
        return out

    def pc_bjacobi(self, r, out=None, lev0=0, scale=1.0, tpow=0, flags=0):
        """z = blockdiag(M1)^-1 r, one block per owned element: PCBJACOBI with one block per element, as the reference
        sets it on ksp1 (eul/HorizSolve.cpp:77-84)."""
        nlev = r.shape[1]
        self._chk(r, self.n1, nlev, "r")
        if out is None:
            out = self.zeros(self.n1, nlev)
        check(self.L.mimsem_gpu_pc_bjacobi_M1(self._h, lev0, nlev, nlev, scale, tpow, flags, r.data_ptr(), out.data_ptr(), self._stream()))
        return out

    def to_vertical(self, cols):
        """L2Vecs::HorizToVert: (n2, nlev) column layout -> (nel_owned, nlev * p^2) per-element vertical vectors."""
        nlev = cols.shape[1]
        self._chk(cols, self.n2, nlev, "cols")
        out = self.torch.empty((self.nel_owned, nlev * self.p * self.p), dtype=self.torch.float64, device=cols.device)
        check(self.L.mimsem_gpu_columns_to_vertical(self._h, nlev, nlev, cols.data_ptr(), out.data_ptr(), self._stream()))
        return out

    def from_vertical(self, vert, out=None):
        """L2Vecs::VertToHoriz: (nel_owned, nlev * p^2) -> (n2, nlev) column layout."""
        nlev = vert.shape[1] // (self.p * self.p)
        if out is None:
            out = self.zeros(self.n2, nlev)
        check(self.L.mimsem_gpu_vertical_to_columns(self._h, nlev, nlev, vert.data_ptr(), out.data_ptr(), self._stream()))
        return out

    def capture(self, op, x, coeff=None, out=None, **kw):
        """Capture one apply into a CUDA graph; returns (replay, out)."""
        torch = self.torch
        if out is None:
            out = self.empty(self.space_sizes(op)[1], x.shape[1])
        self.apply(op, x, coeff=coeff, out=out, **kw)
        torch.cuda.synchronize(self.device)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            self.apply(op, x, coeff=coeff, out=out, **kw)
        return graph.replay, out

    # ---------------------------------------------------------------- end to end with host buffers
    def apply_host(self, op, x, coeff=None, lev0=0, scale=1.0, tpow=0, flags=0, out=None):
        """x: numpy (nlev, nin) in the reference's per-level layout; returns numpy (nlev, nout)."""
        nin, nout, ncoef = self.space_sizes(op)
        x = np.ascontiguousarray(x, dtype=np.float64)
        nlev = x.shape[0]
        assert x.shape == (nlev, nin), x.shape
        if out is None:
            out = np.empty((nlev, nout))
        cp = None
        if ncoef:
            coeff = np.ascontiguousarray(coeff, dtype=np.float64)
            assert coeff.shape == (nlev, ncoef)
            cp = _np_ptr(coeff, _dp)
        check(self.L.mimsem_gpu_apply_host(self._h, OPS[op], lev0, nlev, scale, tpow, flags, cp, _np_ptr(x, _dp),
                                           _np_ptr(out, _dp)))
        return out

    def incidence_csr(self, which):
        """The +-1 stencil of E10/E01/E21/E12 as a scipy CSR matrix over local indices."""
        import scipy.sparse as sp
        w = OPS[which] - 10
        sz = np.zeros(3, dtype=np.int64)
        check(self.L.mimsem_gpu_incidence_csr(self._h, w, _np_ptr(sz, _lp), None, None, None))
        indptr = np.zeros(sz[0] + 1, dtype=np.int64)
        indices = np.zeros(sz[2], dtype=np.int32)
        vals = np.zeros(sz[2])
        check(self.L.mimsem_gpu_incidence_csr(self._h, w, _np_ptr(sz, _lp), _np_ptr(indptr, _lp), _np_ptr(indices, _ip),
                                              _np_ptr(vals, _dp)))
        return sp.csr_matrix((vals, indices, indptr), shape=(int(sz[0]), int(sz[1])))

    @property
    def launch_count(self):
        return int(self.L.mimsem_gpu_launch_count(self._h))
