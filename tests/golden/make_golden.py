#!/usr/bin/env python3
"""Generate the golden fixtures under tests/golden/ FROM THE REFERENCE ITSELF.

Sources of truth (this container only; /root/reference does not exist on the GPU box):
  * oracle/_ref/meshes/*      -- written by the reference's own scr/Setup.py / Setup_Box.py
                                 (oracle/gen_meshes.py)
  * oracle/_ref/libref_*.so   -- the reference's unmodified {src,eul,box} hot-path sources behind
                                 the PETSc/MPI shim (oracle/Makefile): constructor + assemble(...)
                                 per emulated rank, merged CSR, SpMV == MatMult.
The reference ships no golden vectors of its own; these are outputs of the
reference run here on seeded inputs.

    make -C oracle && python oracle/gen_meshes.py --big && python tests/golden/make_golden.py
"""
import glob
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import refbind as rb  # noqa: E402

SCALE = 1.0e8


def ints(path):
    return np.loadtxt(path, dtype=np.int64).astype(np.int32).ravel()


def topo_fixture(name, nprocs):
    d = os.path.join(rb.REFDIR, "meshes", name, "input")
    out = {}
    for r in range(nprocs):
        for key, stem in (("loc0", "nodes"), ("loc1x", "edges_x"), ("loc1y", "edges_y"), ("loc2", "faces"),
                          ("sizes", "local_sizes")):
            out["%s_%d" % (key, r)] = ints(os.path.join(d, "%s_%04d.txt" % (stem, r)))
    np.savez_compressed(os.path.join(HERE, "topo_%s.npz" % name), **out)


def topo_digests():
    """sha256 of the int32 little-endian bytes of every topology file of every generated mesh."""
    dig = {}
    for d in sorted(glob.glob(os.path.join(rb.REFDIR, "meshes", "*"))):
        name = os.path.basename(d)
        nprocs = int(name.split("_np")[1])
        h = hashlib.sha256()
        for r in range(nprocs):
            for stem in ("nodes", "edges_x", "edges_y", "faces", "local_sizes"):
                h.update(ints(os.path.join(d, "input", "%s_%04d.txt" % (stem, r))).astype("<i4").tobytes())
        dig[name] = h.hexdigest()
    with open(os.path.join(HERE, "topo_sha256.json"), "w") as f:
        json.dump(dig, f, indent=1, sort_keys=True)


def thickness(rng, nk, nq, base):
    # horizontally non-uniform layers so the per-point table matters
    return base * (1.0 + np.arange(nk))[:, None] * rng.uniform(0.9, 1.1, (nk, nq))


def ops_fixture(variant, kind, p, ne, nprocs, nk, fname, seed):
    rng = np.random.default_rng(seed)
    R = rb.Reference(variant, rb.mesh_dir(kind, p, ne, nprocs), nprocs, nk=nk)
    N0, N1, N2 = R.N0, R.N1, R.N2
    out = dict(p=p, ne=ne, nprocs=nprocs, nk=nk, N0=N0, N1=N1, N2=N2, scale=SCALE if variant != "src" else 1.0)
    scale = out["scale"]
    # geometry + basis as the reference computed them
    dets, Js = zip(*[R.geom(r) for r in range(nprocs)])
    out["det"] = np.concatenate(dets)
    out["J"] = np.concatenate(Js).reshape(-1, (p + 1) ** 2, 4)
    qx, qw, l, e = R.basis()
    out.update(gll_x=qx, gll_w=qw, ljxi=l, ejxi=e)
    if variant != "src":
        nq = N0
        thick = thickness(rng, nk, nq, 100.0)
        out["thick"] = thick
        for r in range(nprocs):
            R.set_thick(r, thick[:, R.loc(r, "locq" if variant == "eul" else "loc0")])
    x1 = rng.uniform(-1, 1, (nk, N1))
    x2 = rng.uniform(-1, 1, (nk, N2))
    x0 = rng.uniform(-1, 1, (nk, N0))
    h2 = rng.uniform(0.5, 1.5, (nk, N2)) * 1.0e4
    u1 = rng.uniform(-1, 1, (nk, N1)) * float(np.mean(out["det"]))
    out.update(x1=x1, x2=x2, x0=x0, h2=h2, u1=u1)
    # drawn AFTER every earlier field so that the older vectors of the fixture are unchanged
    q0 = rng.uniform(-1, 1, (nk, N0)) * 1.0e-4          # potential vorticity
    u1_up = u1 * 1.0e-3                                  # advecting velocity: departure points move ~0.1 element widths
    up_fac, up_dt = 0.5, 300.0                           # UP_TAU (src/SWEqn_Picard.cpp:30) and a time step
    out.update(q0=q0, u1_up=u1_up, up_fac=up_fac, up_dt=up_dt)
    # second velocity / density of the flux diagnostic (drawn last, see above)
    x1b = rng.uniform(-1, 1, (nk, N1))
    h2b = rng.uniform(0.5, 1.5, (nk, N2)) * 1.0e4
    out.update(x1b=x1b, h2b=h2b)

    def run(op, x, **kw):
        ys = []
        for lev in range(nk):
            args = {k: (v[lev] if isinstance(v, np.ndarray) else v) for k, v in kw.items()}
            A = R.assemble(op, lev=lev, scale=scale, **args)
            ys.append(A @ x[lev])
        return np.array(ys)

    if variant == "eul":
        out["y_Umat_vs1"] = run("Umat", x1, flag=True)
        out["y_Umat_vs0"] = run("Umat", x1, flag=False)
        out["y_Wmat_vs1"] = run("Wmat", x2, flag=True)
        out["y_Pmat"] = run("Pmat", x0)
        out["y_Pmat_h"] = run("Pmat_h", x0, c2=h2)
        out["y_Uhmat_cv1"] = run("Uhmat", x1, flag=True, c2=h2)
        out["y_Uhmat_cv0"] = run("Uhmat", x1, flag=False, c2=h2)
        out["y_Whmat_vs1"] = run("Whmat", x2, flag=True, c2=h2)
        out["y_WtQUmat"] = run("WtQUmat", x1, c1=u1)
        out["y_RotMat"] = run("RotMat", x1, c0=q0)
        # operators of the horizontal-vorticity / vertical-momentum terms
        out["y_Ut_mat"] = np.array([R.assemble("Ut_mat", lev=lev, scale=scale) @ x1[lev] for lev in range(nk - 1)])
        out["y_Ut_mat_h"] = run("Ut_mat_h", x1, c2=h2)
        out["y_WtQdUdz_mat"] = run("WtQdUdz_mat", x1, c1=u1)
        out["y_UtQWmat"] = run("UtQWmat", x2, c1=u1)
        # the reference's own matrix-free twin, driven as diagnose_fluxes does (eul/HorizSolve.cpp:298-306):
        # vl = sum of assemble_hu(u_a, h_b, fac), then the reverse ADD scatter
        out["y_Uvec_hu"] = np.array([R.uvec_assemble_hu([x1[lev], x1[lev], x1b[lev], x1b[lev]], [h2[lev], h2b[lev], h2[lev], h2b[lev]],
                                                        [1.0 / 3.0, 1.0 / 6.0, 1.0 / 6.0, 1.0 / 3.0], lev=lev, scale=scale) for lev in range(nk)])
        out["y_Uvec"] = np.array([R.uvec_apply(x1[lev], lev=lev, scale=scale)[0] for lev in range(nk)])
        # Rayleigh friction (Umat_ray, eul/Assembly.cpp:1846-1979): Exner-pressure 2-forms whose ratio to level 0, times the
        # thickness ratio, puts sigma = (p / p_s) on both sides of the 0.7 threshold of compute_k_v (drawn last, see above)
        ex2 = 1.0e3 * (1.0 + 0.03 * rng.uniform(-1, 1, (nk, N2)))
        ex2 *= (np.linspace(1.0, 0.9, nk) * thick.mean(axis=1) / thick[0].mean())[:, None]
        ray_dt = 300.0
        out.update(ex2=ex2, ray_dt=ray_dt)
        out["y_Umat_ray"] = np.array([R.assemble("Umat_ray", lev=lev, scale=scale, c2=ex2[lev], c1=ex2[0], dt=ray_dt) @ x1[lev]
                                      for lev in range(nk)])
    elif variant == "src":
        out["y_Umat"] = run("Umat", x1)
        out["y_Wmat"] = run("Wmat", x2)
        out["y_Pmat"] = run("Pmat", x0)
        out["y_Uhmat"] = run("Uhmat", x1, c2=h2)
        out["y_WtQUmat"] = run("WtQUmat", x1, c1=u1)
        out["y_RotMat"] = run("RotMat", x1, c0=q0)
        out["y_RotMat_up"] = run("RotMat_up", x1, c0=q0, c1=u1_up, tau=up_fac, dt=up_dt)
        out["y_Phmat_up"] = run("Phmat_up", x0, c1=u1_up, c2=h2, tau=up_fac, dt=up_dt)
    else:  # box: Umat/Wmat are assembled once at level 0 in the ctor (box/Assembly.cpp:44-45, 171-172)
        out["y_Umat_M"] = run("Umat", x1, flag=True)
        out["y_Umat_Mo"] = run("Umat", x1, flag=False)
        out["y_Wmat_M"] = run("Wmat", x2, flag=True)
        out["y_Uhmat_cv1"] = run("Uhmat", x1, flag=True, c2=h2)
        out["y_WtQUmat"] = run("WtQUmat", x1, c1=u1)
        out["y_RotMat"] = run("RotMat", x1, c0=q0)
    for nm in ("E10", "E01", "E21", "E12"):
        A = R.assemble(nm)
        out["%s_indptr" % nm] = A.indptr.astype(np.int64)
        out["%s_indices" % nm] = A.indices.astype(np.int32)
        out["%s_data" % nm] = A.data
    R.close()
    np.savez_compressed(os.path.join(HERE, fname), **out)


def geom_fixture(p, ne, nprocs, nk, rank, fname, seed):
    """Geom::interp* and Geom::initTopog of the reference on one rank."""
    rng = np.random.default_rng(seed)
    R = rb.Reference("eul", rb.mesh_dir("sphere", p, ne, nprocs), nprocs, nk=nk)
    i = R.info(rank)
    v0 = rng.uniform(-1, 1, i["n0"])
    v1 = rng.uniform(-1, 1, i["n1x"] + i["n1y"])
    v2 = rng.uniform(-1, 1, i["n2"])
    out = dict(p=p, ne=ne, nprocs=nprocs, nk=nk, rank=rank, v0=v0, v1=v1, v2=v2)
    out["interp0"] = R.geom_interp(rank, 0, v0)
    out["interp1_l"] = R.geom_interp(rank, 1, v1)
    out["interp2_l"] = R.geom_interp(rank, 2, v2)
    out["interp1_g"] = R.geom_interp(rank, 3, v1)
    out["interp2_g"] = R.geom_interp(rank, 4, v2)
    out["thick"] = R.init_topog(rank)
    R.close()
    np.savez_compressed(os.path.join(HERE, fname), **out)


def main():
    geom_fixture(3, 4, 6, 5, 2, "geom_eul_sphere_p3_ne4_rank2.npz", seed=11)
    for name, nprocs in (("sphere_p3_ne4_np6", 6), ("sphere_p3_ne4_np24", 24), ("sphere_p4_ne2_np6", 6),
                         ("sphere_p2_ne2_np6", 6), ("box_p3_ne4_np4", 4), ("box_p3_ne4_np1", 1)):
        topo_fixture(name, nprocs)
    topo_digests()
    ops_fixture("eul", "sphere", 3, 4, 6, 3, "ops_eul_sphere_p3_ne4.npz", seed=0)
    ops_fixture("eul", "sphere", 4, 2, 6, 2, "ops_eul_sphere_p4_ne2.npz", seed=1)
    ops_fixture("src", "sphere", 3, 4, 6, 1, "ops_src_sphere_p3_ne4.npz", seed=2)
    ops_fixture("box", "box", 3, 4, 1, 2, "ops_box_p3_ne4.npz", seed=3)


if __name__ == "__main__":
    main()
