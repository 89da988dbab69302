"""Shared helpers for the parity tests (test infrastructure)."""
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
REF_MESHES = os.path.join(ROOT, "oracle", "_ref", "meshes")

TOL = 1.0e-12   # BASELINE.json north_star: FP64 operator applications match to relative L2 <= 1e-12


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64).ravel()
    b = np.asarray(b, dtype=np.float64).ravel()
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def golden(name):
    return np.load(os.path.join(GOLDEN, name))


def ref_mesh_dir(kind, p, ne, nprocs):
    return os.path.join(REF_MESHES, "%s_p%d_ne%d_np%d" % (kind, p, ne, nprocs))


def have_ref_mesh(kind, p, ne, nprocs):
    return os.path.isdir(os.path.join(ref_mesh_dir(kind, p, ne, nprocs), "input"))


def have_ref_lib(variant):
    return os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libref_%s.so" % variant))


def eul_levels(nk, ztop=30000.0, mu=15.0):
    """Stretched level heights of the reference's baroclinic test case (eul/UMJS14.cpp:39, 124-129)."""
    f = np.arange(nk + 1) / nk
    return ztop * (np.sqrt(mu * f * f + 1.0) - 1.0) / (np.sqrt(mu + 1.0) - 1.0)


def synthetic_thickness(xyz, nk, kind="sphere", ztop=30000.0):
    """thick[nk][NQ]: the reference's level function times a horizontally non-uniform factor
    (1 + 0.1 cos(lat)) that mimics topography."""
    if kind == "sphere":
        dz = np.diff(eul_levels(nk, ztop))
        r = np.linalg.norm(xyz, axis=1)
        lat = np.arcsin(xyz[:, 2] / r)
        return dz[:, None] * (1.0 + 0.1 * np.cos(lat))[None, :]
    dz = np.full(nk, 1500.0 / nk)   # box/Bubble.cpp:25, 40-42
    return dz[:, None] * (1.0 + 0.1 * np.cos(2 * np.pi * xyz[:, 0] / 1000.0))[None, :]


def synthetic_fields(rng, nk, N0, N1, N2, det_mean):
    """Seeded inputs shaped as SURVEY.md section 8d prescribes, per-level layout (nk, n)."""
    return dict(x1=rng.uniform(-1, 1, (nk, N1)), x2=rng.uniform(-1, 1, (nk, N2)), x0=rng.uniform(-1, 1, (nk, N0)),
                h2=rng.uniform(0.5, 1.5, (nk, N2)) * 1.0e4, u1=rng.uniform(-1, 1, (nk, N1)) * det_mean)


def to_cols(engine, a, space):
    """numpy (nlev, n) per-level array (reference numbering) -> engine column-layout device tensor (n, nlev)."""
    import torch
    t = torch.from_numpy(np.ascontiguousarray(a)).to("cuda:%d" % engine.device)
    return engine.to_columns(t, space)


def to_np(engine, t, space):
    """engine column-layout device tensor (n, nlev) -> numpy (nlev, n) in the reference numbering."""
    return engine.to_levels(t, space).cpu().numpy()


def block_jacobi_reference(A, r, nb):
    """blockdiag(A)^-1 r with PETSc's PCBJACOBI blocks for PCBJacobiSetTotalBlocks(pc, N / nb, NULL): equal consecutive
    row blocks of nb rows (eul/HorizSolve.cpp:77-84: nb = 2 p^2, the edges element e owns), dense solves."""
    A = A.tocsr()
    z = np.zeros(A.shape[0])
    for r0 in range(0, A.shape[0], nb):
        rows = np.arange(r0, r0 + nb)
        z[rows] = np.linalg.solve(A[rows][:, rows].toarray(), r[rows])
    return z
