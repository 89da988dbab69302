"""CPU tests: the numpy oracle restatement (oracle/mimsem_oracle.py) is pinned against golden
vectors produced by the reference's own sources (oracle/_ref; tests/golden/make_golden.py) and,
when oracle/_ref is present, against a live run of the reference."""
import os

import numpy as np
import pytest
import scipy.sparse as sp

from helpers import TOL, golden, have_ref_lib, have_ref_mesh, ref_mesh_dir, rel_l2

from oracle import mimsem_oracle as mo

needs_mesh = pytest.mark.skipif(not have_ref_mesh("sphere", 3, 4, 6), reason="oracle/_ref/meshes not generated")


def _csr(g, nm, shape):
    return sp.csr_matrix((g[nm + "_data"], g[nm + "_indices"], g[nm + "_indptr"]), shape=shape)


@needs_mesh
def test_oracle_eul_vs_golden():
    g = golden("ops_eul_sphere_p3_ne4.npz")
    O = mo.Oracle(ref_mesh_dir("sphere", 3, 4, 6), 6, "sphere", "eul")
    O.set_thick(g["thick"])
    s = float(g["scale"])
    nk = int(g["nk"])
    assert rel_l2(np.concatenate(O.det), g["det"]) < 1e-14
    for lev in range(nk):
        assert rel_l2(O.umat(lev, s, 1) @ g["x1"][lev], g["y_Umat_vs1"][lev]) < TOL
        assert rel_l2(O.umat(lev, s, 0) @ g["x1"][lev], g["y_Umat_vs0"][lev]) < TOL
        assert rel_l2(O.wmat(lev, s, 1) @ g["x2"][lev], g["y_Wmat_vs1"][lev]) < TOL
        assert rel_l2(O.pmat(lev, s) @ g["x0"][lev], g["y_Pmat"][lev]) < TOL
        assert rel_l2(O.pmat(lev, s, h2=g["h2"][lev]) @ g["x0"][lev], g["y_Pmat_h"][lev]) < TOL
        assert rel_l2(O.umat(lev, s, 1, h2=g["h2"][lev], tpow_h=1) @ g["x1"][lev], g["y_Uhmat_cv1"][lev]) < TOL
        assert rel_l2(O.umat(lev, s, 1, h2=g["h2"][lev], tpow_h=0) @ g["x1"][lev], g["y_Uhmat_cv0"][lev]) < TOL
        assert rel_l2(O.wmat(lev, s, 1, rho=g["h2"][lev], tpow_rho=1) @ g["x2"][lev], g["y_Whmat_vs1"][lev]) < TOL
        assert rel_l2(O.wtqumat(g["u1"][lev], lev, s) @ g["x1"][lev], g["y_WtQUmat"][lev]) < TOL
        assert rel_l2(O.rotmat(g["q0"][lev], lev, s, 2) @ g["x1"][lev], g["y_RotMat"][lev]) < TOL
        # SURVEY section 8f-2 operators, stated through the ones the device already has:
        #   Ut_mat::assemble_h == Uhmat without thickness factors, WtQdUdz_mat == 2 WtQUmat without thickness factors,
        #   UtQWmat == WtQdUdz_mat^T, Ut_mat::assemble == Umat weighted by the mean thickness of two levels
        assert rel_l2(O.ut_mat_h(g["h2"][lev], s) @ g["x1"][lev], g["y_Ut_mat_h"][lev]) < TOL
        assert rel_l2(O.wtqdudz_mat(g["u1"][lev], s) @ g["x1"][lev], g["y_WtQdUdz_mat"][lev]) < TOL
        assert rel_l2(O.utqwmat(g["u1"][lev], s) @ g["x2"][lev], g["y_UtQWmat"][lev]) < TOL
        if lev < nk - 1:
            assert rel_l2(O.ut_mat(lev, s) @ g["x1"][lev], g["y_Ut_mat"][lev]) < TOL
        # the reference's matrix-free twins (Uvec::assemble, Uvec::assemble_hu as diagnose_fluxes drives it) against its
        # own matrices: vl = 1/3 F(h1) u1 + 1/6 F(h2) u1 + 1/6 F(h1) u2 + 1/3 F(h2) u2
        assert rel_l2(O.umat(lev, s, 1) @ g["x1"][lev], g["y_Uvec"][lev]) < TOL
        F1, F2 = O.umat(lev, s, 1, h2=g["h2"][lev], tpow_h=1), O.umat(lev, s, 1, h2=g["h2b"][lev], tpow_h=1)
        hu = F1 @ (g["x1"][lev] / 3.0 + g["x1b"][lev] / 6.0) + F2 @ (g["x1"][lev] / 6.0 + g["x1b"][lev] / 3.0)
        assert rel_l2(hu, g["y_Uvec_hu"][lev]) < TOL
    E10, E01 = O.e10()
    E21, E12 = O.e21()
    for nm, A in (("E10", E10), ("E01", E01), ("E21", E21), ("E12", E12)):
        assert abs(A - _csr(g, nm, A.shape)).max() == 0.0   # exact +-1 stencils
    assert abs(E21 @ E10).max() == 0.0                      # E21 E10 = 0 as a matrix identity


@needs_mesh
@pytest.mark.parametrize("fname,p,ne", [("ops_eul_sphere_p3_ne4.npz", 3, 4), ("ops_eul_sphere_p4_ne2.npz", 4, 2)])
def test_oracle_rayleigh_friction_vs_golden(fname, p, ne):
    """Umat_ray (eul/Assembly.cpp:1846-1979): the numpy restatement against vectors of the reference's own class."""
    g = golden(fname)
    O = mo.Oracle(ref_mesh_dir("sphere", p, ne, 6), 6, "sphere", "eul")
    O.set_thick(g["thick"])
    s, dt = float(g["scale"]), float(g["ray_dt"])
    for lev in range(int(g["nk"])):
        y = O.umat_ray(lev, s, dt, g["ex2"][lev], g["ex2"][0]) @ g["x1"][lev]
        assert rel_l2(y, g["y_Umat_ray"][lev]) < TOL, (lev, rel_l2(y, g["y_Umat_ray"][lev]))
    # both branches of compute_k_v are exercised: some points sit below the sigma = 0.7 threshold, most above
    assert 0 < (g["y_Umat_ray"] == 0).mean() < 0.5


@needs_mesh
@pytest.mark.parametrize("p,ne", [(3, 4), (4, 2)])
def test_oracle_quadrature_projections_vs_golden(p, ne):
    """WtQmat, UtQmat, PtQmat (eul/Assembly.cpp:707-751, 824-902, 766-808): the numpy restatement against vectors of the
    reference's own classes (tests/golden/make_golden_quadproj.py)."""
    g = golden("quadproj_eul_sphere_p%d_ne%d.npz" % (p, ne))
    O = mo.Oracle(ref_mesh_dir("sphere", p, ne, 6), 6, "sphere", "eul")
    assert rel_l2(O.wtqmat() @ g["xq"], g["y_WtQmat"]) < 1e-14
    assert rel_l2(O.utqmat() @ g["uq"], g["y_UtQmat"]) < 1e-14
    assert rel_l2(O.ptqmat() @ g["xq"], g["y_PtQmat"]) < 1e-14


@needs_mesh
def test_oracle_src_vs_golden():
    g = golden("ops_src_sphere_p3_ne4.npz")
    O = mo.Oracle(ref_mesh_dir("sphere", 3, 4, 6), 6, "sphere", "src")
    assert rel_l2(O.umat() @ g["x1"][0], g["y_Umat"][0]) < TOL
    assert rel_l2(O.wmat() @ g["x2"][0], g["y_Wmat"][0]) < TOL
    assert rel_l2(O.pmat(tpow=0) @ g["x0"][0], g["y_Pmat"][0]) < TOL
    assert rel_l2(O.umat(h2=g["h2"][0]) @ g["x1"][0], g["y_Uhmat"][0]) < TOL
    assert rel_l2(O.wtqumat(g["u1"][0], tpow=0) @ g["x1"][0], g["y_WtQUmat"][0]) < TOL
    # rotational term and the PV-upwinded operators of BASELINE config 2 (src/Assembly.cpp:1346-1395, 1784-1853, 499-567)
    tau = float(g["up_fac"]) * float(g["up_dt"])
    assert rel_l2(O.rotmat(g["q0"][0]) @ g["x1"][0], g["y_RotMat"][0]) < TOL
    assert rel_l2(O.rotmat(g["q0"][0], u1=g["u1_up"][0], tau=tau) @ g["x1"][0], g["y_RotMat_up"][0]) < TOL
    assert rel_l2(O.phmat_up(g["u1_up"][0], g["h2"][0], tau) @ g["x0"][0], g["y_Phmat_up"][0]) < TOL
    # the upwinding is not a no-op on these inputs
    assert rel_l2(g["y_RotMat_up"][0], g["y_RotMat"][0]) > 1e-3


@pytest.mark.skipif(not have_ref_mesh("box", 3, 4, 1), reason="oracle/_ref/meshes not generated")
def test_oracle_box_vs_golden():
    g = golden("ops_box_p3_ne4.npz")
    O = mo.Oracle(ref_mesh_dir("box", 3, 4, 1), 1, "box", "box")
    O.set_thick(g["thick"])
    s = float(g["scale"])
    for lev in range(int(g["nk"])):
        # box/: Umat/Wmat are built once with level-0 thickness (box/Assembly.cpp:44-45)
        assert rel_l2(O.umat(0, s, 1) @ g["x1"][lev], g["y_Umat_M"][lev]) < TOL
        assert rel_l2(O.umat(0, s, 0) @ g["x1"][lev], g["y_Umat_Mo"][lev]) < TOL
        assert rel_l2(O.wmat(0, s, 1) @ g["x2"][lev], g["y_Wmat_M"][lev]) < TOL
        assert rel_l2(O.umat(lev, s, 1, h2=g["h2"][lev], tpow_h=1) @ g["x1"][lev], g["y_Uhmat_cv1"][lev]) < TOL
        assert rel_l2(O.wtqumat(g["u1"][lev], lev, s) @ g["x1"][lev], g["y_WtQUmat"][lev]) < TOL
        assert rel_l2(O.rotmat(g["q0"][lev], lev, s, 2) @ g["x1"][lev], g["y_RotMat"][lev]) < TOL


@pytest.mark.skipif(not (have_ref_lib("eul") and have_ref_mesh("sphere", 4, 2, 6)), reason="oracle/_ref not built")
def test_oracle_vs_live_reference_p4():
    """Live run of the reference's own sources (oracle/_ref) against the restatement, p = 4."""
    from oracle import refbind as rb
    md = ref_mesh_dir("sphere", 4, 2, 6)
    R = rb.Reference("eul", md, 6, nk=2)
    O = mo.Oracle(md, 6, "sphere", "eul")
    rng = np.random.default_rng(7)
    thick = rng.uniform(100, 200, (2, O.N0))
    O.set_thick(thick)
    for r in range(6):
        R.set_thick(r, thick[:, R.loc(r, "locq")])
    x = rng.uniform(-1, 1, O.N1)
    h = rng.uniform(0.5, 1.5, O.N2)
    assert rel_l2(O.umat(1, 1e8, 1) @ x, R.assemble("Umat", 1, 1e8, True) @ x) < TOL
    assert rel_l2(O.umat(1, 1e8, 1, h2=h, tpow_h=1) @ x, R.assemble("Uhmat", 1, 1e8, True, c2=h) @ x) < TOL
    assert rel_l2(O.wtqumat(x, 0, 1e8) @ x, R.assemble("WtQUmat", 0, 1e8, c1=x) @ x) < TOL
    # the reference's own matrix-free twin (Uvec::assemble, eul/Assembly.cpp:2124-2196) agrees with its matrix
    y, _ = R.uvec_apply(x, lev=1, scale=1e8)
    assert rel_l2(y, R.assemble("Umat", 1, 1e8, True) @ x) < TOL
    R.close()


@pytest.mark.skipif(not (have_ref_lib("src") and have_ref_mesh("sphere", 4, 2, 6)), reason="oracle/_ref not built")
def test_oracle_upwinded_operators_vs_live_reference_p4():
    """BASELINE config 2's operators at p = 4: live run of the reference's src/ sources against the restatement
    (RotMat, RotMat_up, Phmat::assemble_up; src/Assembly.cpp:1346-1395, 1784-1853, 499-567)."""
    from oracle import refbind as rb
    md = ref_mesh_dir("sphere", 4, 2, 6)
    R = rb.Reference("src", md, 6, nk=1)
    O = mo.Oracle(md, 6, "sphere", "src")
    rng = np.random.default_rng(17)
    q0 = rng.uniform(-1, 1, O.N0) * 1e-4
    u1 = rng.uniform(-1, 1, O.N1) * float(np.mean(np.abs(np.concatenate(O.det)))) * 1e-3
    h2 = rng.uniform(0.5, 1.5, O.N2) * 1e4
    x1 = rng.uniform(-1, 1, O.N1)
    x0 = rng.uniform(-1, 1, O.N0)
    fac, dt = 0.5, 300.0
    assert rel_l2(O.rotmat(q0) @ x1, R.assemble("RotMat", c0=q0) @ x1) < TOL
    assert rel_l2(O.rotmat(q0, u1=u1, tau=fac * dt) @ x1, R.assemble("RotMat_up", c0=q0, c1=u1, tau=fac, dt=dt) @ x1) < TOL
    assert rel_l2(O.phmat_up(u1, h2, fac * dt) @ x0, R.assemble("Phmat_up", c1=u1, c2=h2, tau=fac, dt=dt) @ x0) < TOL
    R.close()
