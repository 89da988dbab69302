"""ORACLE (test infrastructure, "O2" of SURVEY.md section 8c) -- NOT part of the product path.

A plain numpy/scipy CPU restatement of the reference's horizontal operator path, following the
reference's own algorithm: tabulate the bases, form the DENSE per-element matrices
U^T diag(Q) U exactly as the assemble() methods do, add them into a global sparse matrix
(MatSetValues ADD_VALUES), and apply it with a CSR SpMV (MatMult).  Nothing here is
sum-factorised or matrix-free, so it shares no algebra with the CUDA kernels it checks.

Pinned against the reference itself: tests/test_oracle.py compares every function below with
oracle/_ref (the reference's unmodified sources behind a PETSc shim) and with the golden vectors
under tests/golden/ that were generated from oracle/_ref (tests/golden/make_golden.py).
The reference ships no tests or golden vectors of its own.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
"""
import os

import numpy as np
import scipy.sparse as sp

RAD_SPHERE = 6371220.0  # eul/Geom.cpp:20
BOX_LX = 1000.0         # box/Geom.cpp:20


# --------------------------------------------------------------------------------------------
# Basis  (eul/Basis.cpp)

def gauss_lobatto(n):
    """GaussLobatto::GaussLobatto, eul/Basis.cpp:22-98."""
    s = np.sqrt
    if n == 1:
        x, w = [-1.0, 1.0], [1.0, 1.0]
    elif n == 2:
        x, w = [-1.0, 0.0, 1.0], [1.0 / 3.0, 4.0 / 3.0, 1.0 / 3.0]
    elif n == 3:
        x, w = [-1.0, -s(0.2), s(0.2), 1.0], [1.0 / 6.0, 5.0 / 6.0, 5.0 / 6.0, 1.0 / 6.0]
    elif n == 4:
        x = [-1.0, -s(3.0 / 7.0), 0.0, s(3.0 / 7.0), 1.0]
        w = [0.1, 49.0 / 90.0, 64.0 / 90.0, 49.0 / 90.0, 0.1]
    elif n == 5:
        a = 2.0 * s(7.0) / 21.0
        x = [-1.0, -s(1.0 / 3.0 + a), -s(1.0 / 3.0 - a), s(1.0 / 3.0 - a), s(1.0 / 3.0 + a), 1.0]
        w1, w2 = (14.0 - s(7.0)) / 30.0, (14.0 + s(7.0)) / 30.0
        w = [1.0 / 15.0, w1, w2, w2, w1, 1.0 / 15.0]
    elif n == 6:
        a = 2.0 * s(5.0 / 3.0) / 11.0
        x = [-1.0, -s(5.0 / 11.0 + a), -s(5.0 / 11.0 - a), 0.0, s(5.0 / 11.0 - a), s(5.0 / 11.0 + a), 1.0]
        w1, w2 = (124.0 - 7.0 * s(15.0)) / 350.0, (124.0 + 7.0 * s(15.0)) / 350.0
        w = [1.0 / 21.0, w1, w2, 256.0 / 525.0, w2, w1, 1.0 / 21.0]
    elif n == 7:
        x = [-1.0, -0.871740148509607, -0.591700181433142, -0.209299217902479, 0.209299217902479,
             0.591700181433142, 0.871740148509607, 1.0]
        w = [0.035714285714286, 0.210704227143506, 0.341122692483504, 0.412458794658704, 0.412458794658704,
             0.341122692483504, 0.210704227143506, 0.035714285714286]
    else:
        raise ValueError("invalid gauss-lobatto quadrature order: %d" % n)
    return np.array(x, dtype=np.float64), np.array(w, dtype=np.float64)


def lagrange_eval_q(xn, x, i):
    """LagrangeNode::eval_q, eul/Basis.cpp:183-190."""
    y = 1.0
    for j in range(len(xn)):
        if j != i:
            y *= (x - xn[j]) / (xn[i] - xn[j])
    return y


def lagrange_deriv(xn, x, i):
    """LagrangeNode::evalDeriv, eul/Basis.cpp:192-213."""
    bb = 0.0
    for j in range(len(xn)):
        if j == i:
            continue
        aa = 1.0
        for k in range(len(xn)):
            if k == i or k == j:
                continue
            aa *= (x - xn[k]) / (xn[i] - xn[k])
        bb += aa / (xn[i] - xn[j])
    return bb


def edge_eval(xn, x, i):
    """LagrangeEdge::eval, eul/Basis.cpp:277-286."""
    c = 0.0
    for j in range(i + 1):
        c -= lagrange_deriv(xn, x, j)
    return c


def basis_tables(p, m):
    """ljxi[(m+1),(p+1)], ejxi[(m+1),p]: eul/Basis.cpp:128-136, 241-249."""
    qx, qw = gauss_lobatto(m)
    xn, _ = gauss_lobatto(p)
    ljxi = np.array([[lagrange_eval_q(xn, qx[q], j) for j in range(p + 1)] for q in range(m + 1)])
    ejxi = np.array([[edge_eval(xn, qx[q], i) for i in range(p)] for q in range(m + 1)])
    return qx, qw, ljxi, ejxi


# --------------------------------------------------------------------------------------------
# ElMats  (eul/ElMats.cpp)

def elmats(p, m):
    """U (M1x_j_xy_i :20-45), V (M1y_j_xy_i :55-80), W (M2_j_xy_i :90-112), P (M0_j_xy_i :120-142), Q (Wii :149-186)."""
    qx, qw, L, E = basis_tables(p, m)
    mp1, np1 = m + 1, p + 1
    q2 = mp1 * mp1
    U = np.zeros((q2, p * np1))
    V = np.zeros((q2, p * np1))
    W = np.zeros((q2, p * p))
    P = np.zeros((q2, np1 * np1))
    Q = np.zeros(q2)
    for q in range(q2):
        for j in range(p * np1):
            U[q, j] = L[q % mp1, j % np1] * E[q // mp1, j // np1]
            V[q, j] = E[q % mp1, j % p] * L[q // mp1, j // p]
        for j in range(p * p):
            W[q, j] = E[q % mp1, j % p] * E[q // mp1, j // p]
        for j in range(np1 * np1):
            P[q, j] = L[q % mp1, j % np1] * L[q // mp1, j // np1]
        Q[q] = qw[q % mp1] * qw[q // mp1]
    return dict(U=U, V=V, W=W, P=P, Q=Q, qx=qx, qw=qw, L=L, E=E)


# --------------------------------------------------------------------------------------------
# Topo  (eul/Topo.cpp) -- reads the input/*.txt files the reference reads

def _ints(path):
    with open(path) as f:
        return np.array([int(s) for s in f.read().split()], dtype=np.int64)


class Topo:
    """One emulated rank: eul/Topo.cpp:15-156 (maps) and :200-305 (element index generators)."""

    def __init__(self, inputdir, rank, nprocs, sphere=True):
        res = _ints(os.path.join(inputdir, "grid_res.txt"))
        self.elOrd, self.nElsX = int(res[0]), int(res[1])
        self.nDofsX = self.elOrd * self.nElsX
        self.pi = rank
        self.loc0 = _ints(os.path.join(inputdir, "nodes_%04d.txt" % rank))
        self.loc1x = _ints(os.path.join(inputdir, "edges_x_%04d.txt" % rank))
        self.loc1y = _ints(os.path.join(inputdir, "edges_y_%04d.txt" % rank))
        self.loc2 = _ints(os.path.join(inputdir, "faces_%04d.txt" % rank))
        self.n0, self.n1x, self.n1y, self.n2 = len(self.loc0), len(self.loc1x), len(self.loc1y), len(self.loc2)
        nx2 = nprocs * self.nDofsX * self.nDofsX
        self.nDofs0G = nx2 + (2 if sphere else 0)   # eul/Topo.cpp:113 ; box/Topo.cpp:112
        self.nDofs1G = 2 * nx2
        self.nDofs2G = nx2

    def el_inds(self):
        """elInds0_g / 1x_g / 1y_g / 2_g for every element (ey, ex): eul/Topo.cpp:253-305."""
        p, ne, nx = self.elOrd, self.nElsX, self.nDofsX
        e0 = np.zeros((ne * ne, (p + 1) ** 2), dtype=np.int64)
        e1x = np.zeros((ne * ne, p * (p + 1)), dtype=np.int64)
        e1y = np.zeros((ne * ne, p * (p + 1)), dtype=np.int64)
        e2 = np.zeros((ne * ne, p * p), dtype=np.int64)
        for ey in range(ne):
            for ex in range(ne):
                el = ey * ne + ex
                kk = 0
                for iy in range(p + 1):
                    for ix in range(p + 1):
                        e0[el, kk] = self.loc0[(ey * p + iy) * (nx + 1) + ex * p + ix]
                        kk += 1
                kk = 0
                for iy in range(p):
                    for ix in range(p + 1):
                        e1x[el, kk] = self.loc1x[(ey * p + iy) * (nx + 1) + ex * p + ix]
                        kk += 1
                kk = 0
                for iy in range(p + 1):
                    for ix in range(p):
                        e1y[el, kk] = self.loc1y[(ey * p + iy) * nx + ex * p + ix]
                        kk += 1
                kk = 0
                for iy in range(p):
                    for ix in range(p):
                        # faces: local index + pi*n2 (eul/Topo.cpp:297-301)
                        e2[el, kk] = (ey * p + iy) * nx + ex * p + ix
                        kk += 1
                # NB elInds2_l is element-blocked: (ey*nElsX+ex)*p^2 + k  (eul/Topo.cpp:244-248)
                e2[el, :] = el * p * p + np.arange(p * p) + self.pi * self.n2
        return e0, e1x, e1y, e2


# --------------------------------------------------------------------------------------------
# Geom  (eul/Geom.cpp)

def sphere_geometry(inputdir, rank, topo, m, signed_det=False):
    """Geom::Geom + updateGlobalCoords + initJacobians: eul/Geom.cpp:24-134, 682-741, 245-326.

    Returns det[nel, q2], J[nel, q2, 2, 2], locq[n0q] for one rank."""
    qx, _ = gauss_lobatto(m)
    x = np.loadtxt(os.path.join(inputdir, "geom_%04d.txt" % rank)).reshape(-1, 3).copy()
    locq = _ints(os.path.join(inputdir, "quads_%04d.txt" % rank))
    s = np.stack([np.arctan2(x[:, 1], x[:, 0]), np.arcsin(x[:, 2] / RAD_SPHERE)], axis=1)
    ne = topo.nElsX
    nxq = m * ne
    mp1 = m + 1

    def inds0(ex, ey):
        return np.array([(ey * m + iy) * (nxq + 1) + ex * m + ix for iy in range(mp1) for ix in range(mp1)])

    def rtilde(c, x1, x2):
        return 0.25 * ((1.0 - x1) * (1.0 - x2) * c[0] + (1.0 + x1) * (1.0 - x2) * c[1] + (1.0 + x1) * (1.0 + x2) * c[2]
                       + (1.0 - x1) * (1.0 + x2) * c[3])

    corner_slots = (0, mp1 - 1, mp1 * mp1 - 1, (mp1 - 1) * mp1)
    for ey in range(ne):          # updateGlobalCoords, eul/Geom.cpp:682-724
        for ex in range(ne):
            ii = inds0(ex, ey)
            c = [x[ii[k]].copy() for k in corner_slots]
            for q in range(mp1 * mp1):
                if q in corner_slots:
                    continue
                r = rtilde(c, qx[q % mp1], qx[q // mp1])
                mag = np.sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2])
                x[ii[q]] = RAD_SPHERE * r / mag
                s[ii[q], 0] = np.arctan2(x[ii[q], 1], x[ii[q], 0])
                s[ii[q], 1] = np.arcsin(x[ii[q], 2] / RAD_SPHERE)
    det = np.zeros((ne * ne, mp1 * mp1))
    J = np.zeros((ne * ne, mp1 * mp1, 2, 2))
    for ey in range(ne):          # initJacobians / jacobian, eul/Geom.cpp:726-741, 245-319
        for ex in range(ne):
            el = ey * ne + ex
            ii = inds0(ex, ey)
            c = [x[ii[k]] for k in corner_slots]
            C = np.array(c).T     # 3 x 4
            for q in range(mp1 * mp1):
                x1, x2 = qx[q % mp1], qx[q // mp1]
                lon, lat = s[ii[q]]
                r = rtilde(c, x1, x2)
                rinv = 1.0 / np.sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2])
                A = np.array([[-np.sin(lon), np.cos(lon), 0.0], [0.0, 0.0, 1.0]])
                B = np.array([
                    [np.sin(lon) ** 2 * np.cos(lat) ** 2 + np.sin(lat) ** 2, -0.5 * np.sin(2 * lon) * np.cos(lat) ** 2,
                     -0.5 * np.cos(lon) * np.sin(2 * lat)],
                    [-0.5 * np.sin(2 * lon) * np.cos(lat) ** 2, np.cos(lon) ** 2 * np.cos(lat) ** 2 + np.sin(lat) ** 2,
                     -0.5 * np.sin(lon) * np.sin(2 * lat)],
                    [-np.cos(lon) * np.sin(lat), -np.sin(lon) * np.sin(lat), np.cos(lat)]])
                D = np.array([[-1.0 + x2, -1.0 + x1], [1.0 - x2, -1.0 - x1], [1.0 + x2, 1.0 + x1], [-1.0 - x2, 1.0 - x1]])
                Jq = (A @ B @ C @ D) * (0.25 * RAD_SPHERE * rinv)
                J[el, q] = Jq
                d = Jq[0, 0] * Jq[1, 1] - Jq[0, 1] * Jq[1, 0]
                det[el, q] = d if signed_det else abs(d)   # src/Geom.cpp:251 vs eul/Geom.cpp:325
    return det, J, locq


def box_geometry(topo, nprocs):
    """box/Geom.cpp:132-143, 499-522: constant diagonal Jacobian."""
    npx = int(round(np.sqrt(nprocs)))
    p = topo.elOrd
    h = 0.5 * BOX_LX / (topo.nElsX * npx)
    nel, q2 = topo.nElsX ** 2, (p + 1) ** 2
    J = np.zeros((nel, q2, 2, 2))
    J[:, :, 0, 0] = h
    J[:, :, 1, 1] = h
    return np.full((nel, q2), abs(h * h)), J


# --------------------------------------------------------------------------------------------
# Assembly  (eul/Assembly.cpp) over ALL emulated ranks -> one global CSR matrix

class Oracle:
    """All ranks of one mesh directory; every method returns the reference's assembled global matrix."""

    def __init__(self, meshdir, nprocs, kind="sphere", variant="eul", m=None):
        inputdir = os.path.join(meshdir, "input")
        self.sphere = (kind == "sphere")
        self.variant = variant
        self.nprocs = nprocs
        self.topos = [Topo(inputdir, r, nprocs, self.sphere) for r in range(nprocs)]
        t0 = self.topos[0]
        self.p = t0.elOrd
        self.m = self.p if m is None else m
        self.N0, self.N1, self.N2 = t0.nDofs0G, t0.nDofs1G, t0.nDofs2G
        self.em = elmats(self.p, self.m)
        self.inds = [t.el_inds() for t in self.topos]
        self.det, self.J, self.locq = [], [], []
        for r, t in enumerate(self.topos):
            if self.sphere:
                d, J, lq = sphere_geometry(inputdir, r, t, self.m, signed_det=(variant == "src"))
            else:
                d, J = box_geometry(t, nprocs)
                lq = t.loc0
            self.det.append(d)
            self.J.append(J)
            self.locq.append(np.asarray(lq))
        self.thick = None   # [nk][NQ global]

    def set_thick(self, thick_global):
        """thick[nk][NQ] indexed by GLOBAL quadrature-point id (each rank reads its own copy through locq)."""
        self.thick = np.asarray(thick_global, dtype=np.float64)

    # geom->thickInv[lev][inds_0[ii]] with inds_0 = Geom::elInds0_l (eul/Geom.cpp:799-811)
    def _tinv(self, r, lev):
        t = self.topos[r]
        m, ne = self.m, t.nElsX
        nxq = m * ne
        out = np.zeros((ne * ne, (m + 1) ** 2))
        for ey in range(ne):
            for ex in range(ne):
                loc = np.array([(ey * m + iy) * (nxq + 1) + ex * m + ix for iy in range(m + 1) for ix in range(m + 1)])
                th = self.thick[lev][self.locq[r][loc]]
                out[ey * ne + ex] = 1.0 / th
        return out

    def _metric(self, r):
        J = self.J[r]
        gaa = J[:, :, 0, 0] ** 2 + J[:, :, 1, 0] ** 2                              # eul/Assembly.cpp:103
        gab = J[:, :, 0, 0] * J[:, :, 0, 1] + J[:, :, 1, 0] * J[:, :, 1, 1]          # :104
        gbb = J[:, :, 0, 1] ** 2 + J[:, :, 1, 1] ** 2                              # :105
        return gaa, gab, gbb

    @staticmethod
    def _add(rows, cols, blocks, trip):
        # MatSetValues(M, nr, rows, nc, cols, block, ADD_VALUES) for every element
        nel, nr, nc = blocks.shape
        trip.append((np.repeat(rows, nc, axis=1).ravel(), np.tile(cols, (1, nr)).ravel(), blocks.ravel()))

    @staticmethod
    def _csr(trip, shape):
        r = np.concatenate([t[0] for t in trip])
        c = np.concatenate([t[1] for t in trip])
        v = np.concatenate([t[2] for t in trip])
        return sp.coo_matrix((v, (r, c)), shape=shape).tocsr()

    def _interp2_g(self, r, h2):
        """geom->interp2_g at every quadrature point: (W h_local)/det, eul/Geom.cpp:363-375, 408-417."""
        e2 = self.inds[r][3]
        return (h2[e2] @ self.em["W"].T) / self.det[r]

    def umat(self, lev=0, scale=1.0, tpow=0, h2=None, tpow_h=0):
        """Umat::_assemble (eul/Assembly.cpp:51-153); with h2: Uhmat::assemble (:416-474)."""
        U, V, Q = self.em["U"], self.em["V"], self.em["Q"]
        trip = []
        for r in range(self.nprocs):
            gaa, gab, gbb = self._metric(r)
            c = Q[None, :] * (scale / self.det[r])
            if h2 is not None:
                hi = self._interp2_g(r, h2)
                if tpow_h:
                    hi = hi * self._tinv(r, lev)
                c = hi * c
            if tpow:
                c = c * self._tinv(r, lev)
            Qaa, Qab, Qbb = gaa * c, gab * c, gbb * c
            _, e1x, e1y, _ = self.inds[r]
            self._add(e1x, e1x, np.einsum("qi,eq,qj->eij", U, Qaa, U), trip)
            self._add(e1x, e1y, np.einsum("qi,eq,qj->eij", U, Qab, V), trip)
            self._add(e1y, e1x, np.einsum("qi,eq,qj->eij", V, Qab, U), trip)
            self._add(e1y, e1y, np.einsum("qi,eq,qj->eij", V, Qbb, V), trip)
        return self._csr(trip, (self.N1, self.N1))

    def umat_ray(self, lev, scale, dt, exner_k, exner_s):
        """Umat_ray::assemble(lev, scale, dt, exner_k, exner_s) (eul/Assembly.cpp:1846-1979): the 1-form mass matrix with
        the point weight dt k_v(exner, exner_s) thickInv[lev]; exner = interp2_g(exner_k) thickInv[lev], exner_s =
        interp2_g(exner_s) thickInv[0]; k_v = compute_k_v (:1846-1856) with CP = 1004.5, RD = 287 (:15-16)."""
        CP, RD, sigma_b, k_f = 1004.5, 287.0, 0.7, 1.1574074074074073e-05
        U, V, Q = self.em["U"], self.em["V"], self.em["Q"]
        trip = []
        for r in range(self.nprocs):
            gaa, gab, gbb = self._metric(r)
            ti = self._tinv(r, lev)
            e = self._interp2_g(r, exner_k) * ti
            es = self._interp2_g(r, exner_s) * self._tinv(r, 0)
            sigma = (e / CP) ** (CP / RD) / (es / CP) ** (CP / RD)
            k_v = np.where(sigma < sigma_b, 0.0, k_f * (sigma - sigma_b) / (1.0 - sigma_b)) * dt
            c = Q[None, :] * (scale / self.det[r]) * k_v * ti
            Qaa, Qab, Qbb = gaa * c, gab * c, gbb * c
            _, e1x, e1y, _ = self.inds[r]
            self._add(e1x, e1x, np.einsum("qi,eq,qj->eij", U, Qaa, U), trip)
            self._add(e1x, e1y, np.einsum("qi,eq,qj->eij", U, Qab, V), trip)
            self._add(e1y, e1x, np.einsum("qi,eq,qj->eij", V, Qab, U), trip)
            self._add(e1y, e1y, np.einsum("qi,eq,qj->eij", V, Qbb, V), trip)
        return self._csr(trip, (self.N1, self.N1))

    def wmat(self, lev=0, scale=1.0, tpow=0, rho=None, tpow_rho=0):
        """Wmat::_assemble (eul/Assembly.cpp:341-360); with rho: Whmat::assemble (:1262-1287)."""
        W, Q = self.em["W"], self.em["Q"]
        trip = []
        for r in range(self.nprocs):
            c = Q[None, :] * (scale / self.det[r])
            if rho is not None:
                ri = self._interp2_g(r, rho)
                if tpow_rho:
                    ri = ri * self._tinv(r, lev)
                c = c * ri
            if tpow:
                c = c * self._tinv(r, lev)
            e2 = self.inds[r][3]
            self._add(e2, e2, np.einsum("qi,eq,qj->eij", W, c, W), trip)
        return self._csr(trip, (self.N2, self.N2))

    def pmat(self, lev=0, scale=1.0, tpow=1, h2=None):
        """Pmat::assemble (eul/Assembly.cpp:2021-2036) / assemble_h (:2067-2086); src: tpow=0 (src/Assembly.cpp:324-372)."""
        P, Q = self.em["P"], self.em["Q"]
        trip = []
        for r in range(self.nprocs):
            c = scale * Q[None, :] * self.det[r]
            if tpow:
                c = c * self._tinv(r, lev)
            if h2 is not None:
                c = c * (self._interp2_g(r, h2) * self._tinv(r, lev))
            e0 = self.inds[r][0]
            self._add(e0, e0, np.einsum("qi,eq,qj->eij", P, c, P), trip)
        return self._csr(trip, (self.N0, self.N0))

    def wtqumat(self, u1, lev=0, scale=1.0, tpow=2):
        """WtQUmat::assemble (eul/Assembly.cpp:947-981); src: tpow=0 (src/Assembly.cpp:1172-1218)."""
        U, V, W, Q = self.em["U"], self.em["V"], self.em["W"], self.em["Q"]
        trip = []
        for r in range(self.nprocs):
            J, det = self.J[r], self.det[r]
            _, e1x, e1y, e2 = self.inds[r]
            ul0 = u1[e1x] @ U.T      # interp1_l, eul/Geom.cpp:343-361
            ul1 = u1[e1y] @ V.T
            ux0 = (J[:, :, 0, 0] * ul0 + J[:, :, 0, 1] * ul1) / det   # interp1_g, :377-389
            ux1 = (J[:, :, 1, 0] * ul0 + J[:, :, 1, 1] * ul1) / det
            ti = self._tinv(r, lev) if tpow else 1.0
            if tpow:
                ux0, ux1 = ux0 * ti, ux1 * ti
            c = Q[None, :] * (scale / det)
            Qaa = 0.5 * (ux0 * J[:, :, 0, 0] + ux1 * J[:, :, 1, 0]) * c
            Qab = 0.5 * (ux0 * J[:, :, 0, 1] + ux1 * J[:, :, 1, 1]) * c
            if tpow:
                Qaa, Qab = Qaa * ti, Qab * ti
            self._add(e2, e1x, np.einsum("qi,eq,qj->eij", W, Qaa, U), trip)
            self._add(e2, e1y, np.einsum("qi,eq,qj->eij", W, Qab, V), trip)
        return self._csr(trip, (self.N2, self.N1))

    def _departure_tables(self, r, u1, tau):
        """lx[e,q,j], ly[e,q,j]: LagrangeNode::eval_q at xi_q - tau J^-1 u_g(xi_q)
        (src/Assembly.cpp:531-541 in Phmat::assemble_up, :1811-1820 in RotMat_up::assemble)."""
        U, V = self.em["U"], self.em["V"]
        J, det = self.J[r], self.det[r]
        _, e1x, e1y, _ = self.inds[r]
        mp1, np1 = self.m + 1, self.p + 1
        ul0 = u1[e1x] @ U.T                                            # interp1_l
        ul1 = u1[e1y] @ V.T
        ux0 = (J[:, :, 0, 0] * ul0 + J[:, :, 0, 1] * ul1) / det        # interp1_g, src/Geom.cpp:302-313
        ux1 = (J[:, :, 1, 0] * ul0 + J[:, :, 1, 1] * ul1) / det
        v0 = +J[:, :, 1, 1] * ux0 / det - J[:, :, 0, 1] * ux1 / det    # ux2, src/Assembly.cpp:532-533
        v1 = -J[:, :, 1, 0] * ux0 / det + J[:, :, 0, 0] * ux1 / det
        xq = self.em["qx"]
        xn = gauss_lobatto(self.p)[0]
        q = np.arange(mp1 * mp1)
        ptx = xq[q % mp1][None, :] - tau * v0
        pty = xq[q // mp1][None, :] - tau * v1
        lx = np.stack([lagrange_eval_q(xn, ptx, j) for j in range(np1)], axis=-1)
        ly = np.stack([lagrange_eval_q(xn, pty, j) for j in range(np1)], axis=-1)
        return lx, ly

    def rotmat(self, q0, lev=0, scale=1.0, tpow=0, u1=None, tau=0.0):
        """RotMat::assemble (src/Assembly.cpp:1346-1395; eul/Assembly.cpp:1030-1083 with tpow=2 and scale);
        with u1: RotMat_up::assemble (src/Assembly.cpp:1784-1853), tau = fac*dt."""
        U, V, P, Q = self.em["U"], self.em["V"], self.em["P"], self.em["Q"]
        np1 = self.p + 1
        trip = []
        for r in range(self.nprocs):
            J, det = self.J[r], self.det[r]
            e0, e1x, e1y, _ = self.inds[r]
            if u1 is None:
                vort = q0[e0] @ P.T                                    # interp0
            else:
                lx, ly = self._departure_tables(r, u1, tau)
                jj = np.arange(np1 * np1)
                vort = np.einsum("ej,eqj->eq", q0[e0], lx[:, :, jj % np1] * ly[:, :, jj // np1])
            if tpow:
                vort = vort * self._tinv(r, lev) ** tpow
            dj = J[:, :, 0, 0] * J[:, :, 1, 1] - J[:, :, 0, 1] * J[:, :, 1, 0]
            Qab = vort * (-dj) * Q[None, :] * (scale / det)
            Qba = vort * (+dj) * Q[None, :] * (scale / det)
            self._add(e1x, e1y, np.einsum("qi,eq,qj->eij", U, Qab, V), trip)
            self._add(e1y, e1x, np.einsum("qi,eq,qj->eij", V, Qba, U), trip)
        return self._csr(trip, (self.N1, self.N1))

    def phmat_up(self, u1, h2, tau):
        """Phmat::assemble_up(ul, hl, fac, dt) (src/Assembly.cpp:499-567), tau = fac*dt."""
        P, W, Q = self.em["P"], self.em["W"], self.em["Q"]
        np1 = self.p + 1
        trip = []
        for r in range(self.nprocs):
            e0, _, _, e2 = self.inds[r]
            lx, ly = self._departure_tables(r, u1, tau)
            hx = h2[e2] @ W.T                                          # interp2_l (no determinant)
            jj = np.arange(np1 * np1)
            QP = (hx * Q[None, :])[:, :, None] * lx[:, :, jj % np1] * ly[:, :, jj // np1]
            self._add(e0, e0, np.einsum("qi,eqj->eij", P, QP), trip)
        return self._csr(trip, (self.N0, self.N0))

    # ---- operators of the horizontal-vorticity / vertical-momentum terms ----------------
    def ut_mat(self, lev, scale=1.0):
        """Ut_mat::assemble(lev, scale) (eul/Assembly.cpp:1338-1388): the 1-form mass matrix weighted by the mean
        thickness of levels lev and lev+1 (NOT its inverse)."""
        U, V, Q = self.em["U"], self.em["V"], self.em["Q"]
        trip = []
        for r in range(self.nprocs):
            gaa, gab, gbb = self._metric(r)
            tavg = 0.5 * (1.0 / self._tinv(r, lev) + 1.0 / self._tinv(r, lev + 1))
            c = Q[None, :] * (scale / self.det[r]) * tavg
            _, e1x, e1y, _ = self.inds[r]
            self._add(e1x, e1x, np.einsum("qi,eq,qj->eij", U, gaa * c, U), trip)
            self._add(e1x, e1y, np.einsum("qi,eq,qj->eij", U, gab * c, V), trip)
            self._add(e1y, e1x, np.einsum("qi,eq,qj->eij", V, gab * c, U), trip)
            self._add(e1y, e1y, np.einsum("qi,eq,qj->eij", V, gbb * c, V), trip)
        return self._csr(trip, (self.N1, self.N1))

    def ut_mat_h(self, rho, scale=1.0):
        """Ut_mat::assemble_h(lev, scale, rho) (eul/Assembly.cpp:1390-1440): no thickness factor at all, i.e. exactly
        Uhmat without its 1/thick factors -- umat(h2=rho) with tpow = tpow_h = 0 (the device's apply_M1h with tpow 0)."""
        return self.umat(0, scale, 0, h2=rho, tpow_h=0)

    def wtqdudz_mat(self, u1, scale=1.0):
        """WtQdUdz_mat::assemble(u1, scale) (eul/Assembly.cpp:1581-1640): W^T diag((u_g . J[:,a]) w scale/det) [U V], i.e.
        WtQUmat without the factor 1/2 and without the thickness factors = 2 x wtqumat(tpow = 0) (apply_K, tpow 0)."""
        return 2.0 * self.wtqumat(u1, 0, scale, tpow=0)

    def utqwmat(self, u1, scale=1.0):
        """UtQWmat::assemble(u1, scale) (eul/Assembly.cpp:1490-1538): [U V]^T diag(...) W with interp1_g_t == interp1_g
        (eul/Geom.cpp:392-406), i.e. the transpose of WtQdUdz_mat."""
        return self.wtqdudz_mat(u1, scale).T.tocsr()

    # ---- quadrature-point projections of the initial fields (columns: Geom's global quadrature points) ----
    def _elq_global(self, r):
        """Geom::elInds0_g for every element of rank r (eul/Geom.cpp:813-825): global ids of its (m+1)^2 quadrature points."""
        t = self.topos[r]
        m, ne = self.m, t.nElsX
        nxq = m * ne
        out = np.zeros((ne * ne, (m + 1) ** 2), dtype=np.int64)
        for ey in range(ne):
            for ex in range(ne):
                loc = np.array([(ey * m + iy) * (nxq + 1) + ex * m + ix for iy in range(m + 1) for ix in range(m + 1)])
                out[ey * ne + ex] = self.locq[r][loc]
        return out

    def _nq_global(self):
        return int(max(int(lq.max()) for lq in self.locq)) + 1

    def wtqmat(self):
        """WtQmat::assemble (eul/Assembly.cpp:707-751): W^T diag(w_q) per element, rows = faces, columns = quadrature points."""
        W, Q = self.em["W"], self.em["Q"]
        trip = []
        blk = (W * Q[:, None]).T                                  # [face j][point q]
        for r in range(self.nprocs):
            e2 = self.inds[r][3]
            self._add(e2, self._elq_global(r), np.broadcast_to(blk, (e2.shape[0],) + blk.shape).copy(), trip)
        return self._csr(trip, (self.N2, self._nq_global()))

    def ptqmat(self):
        """PtQmat::assemble (eul/Assembly.cpp:766-808): P^T diag(w_q det_q) per element, rows = nodes."""
        P, Q = self.em["P"], self.em["Q"]
        trip = []
        for r in range(self.nprocs):
            e0 = self.inds[r][0]
            c = Q[None, :] * self.det[r]
            self._add(e0, self._elq_global(r), np.einsum("qj,eq->ejq", P, c), trip)
        return self._csr(trip, (self.N0, self._nq_global()))

    def utqmat(self):
        """UtQmat::assemble (eul/Assembly.cpp:824-902): x-edges U^T diag(w_q) (J00 u_x + J10 u_y), y-edges V^T diag(w_q) (J01 u_x +
        J11 u_y); columns 2 q + c, the two components of a point interleaved."""
        U, V, Q = self.em["U"], self.em["V"], self.em["Q"]
        trip = []
        for r in range(self.nprocs):
            J = self.J[r]
            _, e1x, e1y, _ = self.inds[r]
            q = self._elq_global(r)
            for rows, T, a in ((e1x, U, 0), (e1y, V, 1)):
                for comp in (0, 1):
                    c = Q[None, :] * J[:, :, comp, a]             # J[comp][a]: Qaa = J00, Qab = J10 (x-edges); Qba = J01, Qbb = J11
                    self._add(rows, 2 * q + comp, np.einsum("qj,eq->ejq", T, c), trip)
        return self._csr(trip, (self.N1, 2 * self._nq_global()))

    def e10(self):
        """E10mat::E10mat, eul/Assembly.cpp:1102-1162 (INSERT_VALUES); returns (E10, E01 = -E10^T)."""
        p = self.p
        np1 = p + 1
        rows, cols, vals = [], [], []
        for r in range(self.nprocs):
            e0, e1x, e1y, _ = self.inds[r]
            for el in range(e0.shape[0]):
                for ii in range(p):
                    for jj in range(p):
                        ll = jj * np1 + ii
                        rows += [e1x[el, jj * np1 + ii]] * 2
                        cols += [e0[el, ll], e0[el, ll + np1]]
                        vals += [+1.0, -1.0]
                        rows += [e1y[el, jj * p + ii]] * 2
                        cols += [e0[el, ll], e0[el, ll + 1]]
                        vals += [-1.0, +1.0]
        E10 = sp.coo_matrix((vals, (rows, cols)), shape=(self.N1, self.N0)).tocsr()
        return E10, (-E10.T).tocsr()

    def e21(self):
        """E21mat::E21mat, eul/Assembly.cpp:1170-1220; returns (E21, E12 = -E21^T)."""
        p = self.p
        np1 = p + 1
        rows, cols, vals = [], [], []
        for r in range(self.nprocs):
            _, e1x, e1y, e2 = self.inds[r]
            for el in range(e2.shape[0]):
                for ii in range(p):
                    for jj in range(p):
                        rows += [e2[el, ii * p + jj]] * 4
                        cols += [e1x[el, ii * np1 + jj], e1x[el, ii * np1 + jj + 1], e1y[el, ii * p + jj],
                                 e1y[el, (ii + 1) * p + jj]]
                        vals += [-1.0, +1.0, -1.0, +1.0]
        E21 = sp.coo_matrix((vals, (rows, cols)), shape=(self.N2, self.N1)).tocsr()
        return E21, (-E21.T).tocsr()
