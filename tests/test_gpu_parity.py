"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI (ctypes), against
  (1) golden vectors produced by the reference's own sources (tests/golden/*.npz),
  (2) the numpy oracle restatement on seeded inputs at the UMJS14 shape (C3),
  (3) size-independent properties at BASELINE.json's largest shape (C5).
Tolerance: relative L2 <= 1e-12 (BASELINE.json north_star); incidence stencils bit-exact."""
import numpy as np
import pytest
import scipy.sparse as sp

import mimsem_b200 as mb
from helpers import TOL, golden, rel_l2, synthetic_fields, synthetic_thickness, to_cols, to_np

pytestmark = pytest.mark.gpu


def _engine(kind, p, ne, thick=None, signed=False):
    mesh = mb.Mesh(kind, p, ne, signed_det=signed)
    return mesh, mb.Engine.from_mesh(mesh, 0, thick=thick)


def _apply(eng, op, x, coeff=None, u1=None, **kw):
    sin, sout, sc = eng.SPACES[op]
    c = None if coeff is None else to_cols(eng, coeff, sc)
    if u1 is not None:
        kw["u1"] = to_cols(eng, u1, 1)
    return to_np(eng, eng.apply(op, to_cols(eng, x, sin), coeff=c, **kw), sout)


def _apply_per_level(eng, op, x, coeff=None, **kw):
    """The MatShell pattern: one single-column launch per level."""
    out = []
    for lev in range(x.shape[0]):
        c = None if coeff is None else coeff[lev:lev + 1]
        out.append(_apply(eng, op, x[lev:lev + 1], c, lev0=lev, **kw)[0])
    return np.array(out)


@pytest.mark.parametrize("fname,p,ne", [("ops_eul_sphere_p3_ne4.npz", 3, 4), ("ops_eul_sphere_p4_ne2.npz", 4, 2)])
def test_eul_operators_vs_reference_golden(fname, p, ne):
    g = golden(fname)
    mesh, eng = _engine("sphere", p, ne, thick=g["thick"])
    s = float(g["scale"])
    cases = [("M1", "x1", None, 1, "y_Umat_vs1"), ("M1", "x1", None, 0, "y_Umat_vs0"), ("M2", "x2", None, 1, "y_Wmat_vs1"),
             ("M0", "x0", None, 1, "y_Pmat"), ("M0h", "x0", "h2", 2, "y_Pmat_h"), ("M1h", "x1", "h2", 2, "y_Uhmat_cv1"),
             ("M1h", "x1", "h2", 1, "y_Uhmat_cv0"), ("M2h", "x2", "h2", 2, "y_Whmat_vs1"), ("K", "x1", "u1", 2, "y_WtQUmat"),
             ("R", "x1", "q0", 2, "y_RotMat")]
    for op, xk, ck, tpow, yk in cases:
        coeff = None if ck is None else g[ck]
        y = _apply(eng, op, g[xk], coeff, scale=s, tpow=tpow)
        assert rel_l2(y, g[yk]) < TOL, (op, tpow, rel_l2(y, g[yk]))
        y1 = _apply_per_level(eng, op, g[xk], coeff, scale=s, tpow=tpow)
        # one launch over all levels vs one launch per level (the MatShell pattern); the batched launch may take the
        # TMA tile kernel and the single-column launch the register kernel: same arithmetic, different summation order
        assert rel_l2(y1, y) < 1e-14, op
    assert eng.launch_count > 0


def test_src_operators_vs_reference_golden():
    g = golden("ops_src_sphere_p3_ne4.npz")
    mesh, eng = _engine("sphere", 3, 4, signed=True)
    for op, xk, ck, yk in [("M1", "x1", None, "y_Umat"), ("M2", "x2", None, "y_Wmat"), ("M0", "x0", None, "y_Pmat"),
                           ("M1h", "x1", "h2", "y_Uhmat"), ("K", "x1", "u1", "y_WtQUmat")]:
        y = _apply(eng, op, g[xk], None if ck is None else g[ck], scale=1.0, tpow=0)
        assert rel_l2(y, g[yk]) < TOL, (op, rel_l2(y, g[yk]))
    # BASELINE config 2: rotational term and the potential-vorticity-upwinded operators
    tau = float(g["up_fac"]) * float(g["up_dt"])
    y = _apply(eng, "R", g["x1"], g["q0"])
    assert rel_l2(y, g["y_RotMat"]) < TOL, rel_l2(y, g["y_RotMat"])
    y = _apply(eng, "R_up", g["x1"], g["q0"], u1=g["u1_up"], tau=tau)
    assert rel_l2(y, g["y_RotMat_up"]) < TOL, rel_l2(y, g["y_RotMat_up"])
    y = _apply(eng, "M0h_up", g["x0"], g["h2"], u1=g["u1_up"], tau=tau)
    assert rel_l2(y, g["y_Phmat_up"]) < TOL, rel_l2(y, g["y_Phmat_up"])
    # tau = 0 reduces the upwinded operators to the plain ones
    assert rel_l2(_apply(eng, "R_up", g["x1"], g["q0"], u1=g["u1_up"], tau=0.0), g["y_RotMat"]) < TOL
    assert rel_l2(_apply(eng, "M0h_up", g["x0"], g["h2"], u1=g["u1_up"], tau=0.0), _apply(eng, "M0h", g["x0"], g["h2"])) < TOL


def test_box_operators_vs_reference_golden():
    g = golden("ops_box_p3_ne4.npz")
    mesh, eng = _engine("box", 3, 4, thick=g["thick"])
    s = float(g["scale"])
    FL = mb.engine.FIXED_LEVEL
    # box/: Umat, Wmat are assembled once with the level-0 thickness (box/Assembly.cpp:44-45, 171-172)
    assert rel_l2(_apply(eng, "M1", g["x1"], scale=s, tpow=1, flags=FL), g["y_Umat_M"]) < TOL
    assert rel_l2(_apply(eng, "M1", g["x1"], scale=s, tpow=0), g["y_Umat_Mo"]) < TOL
    assert rel_l2(_apply(eng, "M2", g["x2"], scale=s, tpow=1, flags=FL), g["y_Wmat_M"]) < TOL
    assert rel_l2(_apply(eng, "M1h", g["x1"], g["h2"], scale=s, tpow=2), g["y_Uhmat_cv1"]) < TOL
    assert rel_l2(_apply(eng, "K", g["x1"], g["u1"], scale=s, tpow=2), g["y_WtQUmat"]) < TOL
    assert rel_l2(_apply(eng, "R", g["x1"], g["q0"], scale=s, tpow=2), g["y_RotMat"]) < TOL


@pytest.mark.parametrize("fname,kind,p,ne", [("ops_eul_sphere_p3_ne4.npz", "sphere", 3, 4),
                                             ("ops_eul_sphere_p4_ne2.npz", "sphere", 4, 2), ("ops_box_p3_ne4.npz", "box", 3, 4)])
def test_incidence_bit_exact(fname, kind, p, ne):
    g = golden(fname)
    mesh, eng = _engine(kind, p, ne)
    rng = np.random.default_rng(5)
    mats = {}
    for nm in ("E10", "E01", "E21", "E12"):
        A = eng.incidence_csr(nm)
        ref = sp.csr_matrix((g[nm + "_data"], g[nm + "_indices"], g[nm + "_indptr"]), shape=A.shape)
        ref.sort_indices()
        assert np.array_equal(A.indptr, ref.indptr) and np.array_equal(A.indices, ref.indices)
        assert np.array_equal(A.data, ref.data)             # exact +-1 entries, identical sparsity
        mats[nm] = ref
        # integer-valued vectors: every partial sum is exact, so the device result must equal the CSR product exactly
        x = rng.integers(-1000, 1000, (3, A.shape[1])).astype(np.float64)
        y = _apply(eng, nm, x)
        assert np.array_equal(y, (ref @ x.T).T), nm
        xf = rng.uniform(-1, 1, (3, A.shape[1]))
        assert rel_l2(_apply(eng, nm, xf), (ref @ xf.T).T) < TOL
        # 2 and 8 levels take the 16-byte kernels (two / four levels per thread): still exact on integer data
        for nl in (2, 8):
            xi = rng.integers(-1000, 1000, (nl, A.shape[1])).astype(np.float64)
            assert np.array_equal(_apply(eng, nm, xi), (ref @ xi.T).T), (nm, nl)
    # E21 E10 = 0: exact on integer data, composed on the device
    x0 = rng.integers(-1000, 1000, (2, mesh.N0)).astype(np.float64)
    z = eng.apply("E21", eng.apply("E10", to_cols(eng, x0, 0)))
    assert float(z.abs().max()) == 0.0
    assert abs(mats["E21"] @ mats["E10"]).max() == 0.0


def test_umjs14_shape_vs_oracle(tmp_path):
    """C3: eul p=3, 12x12 elements/face, 30 levels; oracle = numpy restatement of the assemble path."""
    from oracle import mimsem_oracle as mo
    p, ne, nk = 3, 12, 30
    mesh = mb.Mesh("sphere", p, ne)
    thick = synthetic_thickness(mesh.xyz, nk)
    eng = mb.Engine.from_mesh(mesh, 0, thick=thick)
    d = tmp_path / "input"
    d.mkdir()
    mb.write_input("sphere", p, ne, 6, str(d))
    O = mo.Oracle(str(tmp_path), 6, "sphere", "eul")
    O.set_thick(thick)
    rng = np.random.default_rng(0)
    f = synthetic_fields(rng, nk, mesh.N0, mesh.N1, mesh.N2, float(mesh.det.mean()))
    s = 1.0e8
    levs = [0, 13, 29]
    yM1 = _apply(eng, "M1", f["x1"], scale=s, tpow=1)
    yM2 = _apply(eng, "M2", f["x2"], scale=s, tpow=1)
    yM0 = _apply(eng, "M0", f["x0"], scale=s, tpow=1)
    yM0h = _apply(eng, "M0h", f["x0"], f["h2"], scale=s, tpow=2)
    yM1h = _apply(eng, "M1h", f["x1"], f["h2"], scale=s, tpow=2)
    yM2h = _apply(eng, "M2h", f["x2"], f["h2"], scale=s, tpow=2)
    yK = _apply(eng, "K", f["x1"], f["u1"], scale=s, tpow=2)
    for lev in levs:
        assert rel_l2(yM1[lev], O.umat(lev, s, 1) @ f["x1"][lev]) < TOL
        assert rel_l2(yM2[lev], O.wmat(lev, s, 1) @ f["x2"][lev]) < TOL
        assert rel_l2(yM0[lev], O.pmat(lev, s) @ f["x0"][lev]) < TOL
        assert rel_l2(yM0h[lev], O.pmat(lev, s, h2=f["h2"][lev]) @ f["x0"][lev]) < TOL
        assert rel_l2(yM1h[lev], O.umat(lev, s, 1, h2=f["h2"][lev], tpow_h=1) @ f["x1"][lev]) < TOL
        assert rel_l2(yM2h[lev], O.wmat(lev, s, 1, rho=f["h2"][lev], tpow_rho=1) @ f["x2"][lev]) < TOL
        assert rel_l2(yK[lev], O.wtqumat(f["u1"][lev], lev, s) @ f["x1"][lev]) < TOL
    # end-to-end entry point with host buffers gives the same bits as the device-resident path
    assert np.array_equal(eng.apply_host("M1", f["x1"], scale=s, tpow=1), yM1)
    assert np.array_equal(eng.apply_host("K", f["x1"], coeff=f["u1"], scale=s, tpow=2), yK)
    assert np.array_equal(eng.apply_host("E21", f["x1"]), _apply(eng, "E21", f["x1"]))


def _oracle(tmp_path, kind, p, ne, nprocs, variant):
    """numpy restatement of the reference's assemble path (O2, pinned to the reference itself in tests/test_oracle.py)
    on mesh files written by the product's own writer (bit-identical to the reference's, tests/test_host.py)."""
    from oracle import mimsem_oracle as mo
    d = tmp_path / "input"
    d.mkdir()
    mb.write_input(kind, p, ne, nprocs, str(d))
    return mo.Oracle(str(tmp_path), nprocs, kind, variant)


def test_c5_benchmark_shape_vs_oracle(tmp_path):
    """C5 = the configuration bench.py quotes (eul p=4, 48x48 elements per face, 60 levels): the kernels the bench
    times -- k_apply_m1_tile<4,*,60,*>, k_apply_k_tma<4,60>, M2, M0 -- against the oracle's assembled matrices on levels
    0, 29 and 59, relative L2 <= 1e-12."""
    p, ne, nk = 4, 48, 60
    mesh = mb.Mesh("sphere", p, ne)
    thick = synthetic_thickness(mesh.xyz, nk)
    eng = mb.Engine.from_mesh(mesh, 0, thick=thick)
    O = _oracle(tmp_path, "sphere", p, ne, 6, "eul")
    O.set_thick(thick)
    rng = np.random.default_rng(0)
    f = synthetic_fields(rng, nk, mesh.N0, mesh.N1, mesh.N2, float(mesh.det.mean()))
    s = 1.0e8
    l0 = eng.launch_count
    yM1 = _apply(eng, "M1", f["x1"], scale=s, tpow=1)
    yM1h = _apply(eng, "M1h", f["x1"], f["h2"], scale=s, tpow=2)
    yK = _apply(eng, "K", f["x1"], f["u1"], scale=s, tpow=2)
    yM2 = _apply(eng, "M2", f["x2"], scale=s, tpow=1)
    yM0 = _apply(eng, "M0", f["x0"], scale=s, tpow=1)
    assert eng.launch_count > l0
    for lev in (0, 29, 59):
        assert rel_l2(yM1[lev], O.umat(lev, s, 1) @ f["x1"][lev]) < TOL, ("M1", lev)
        assert rel_l2(yM1h[lev], O.umat(lev, s, 1, h2=f["h2"][lev], tpow_h=1) @ f["x1"][lev]) < TOL, ("M1h", lev)
        assert rel_l2(yK[lev], O.wtqumat(f["u1"][lev], lev, s) @ f["x1"][lev]) < TOL, ("K", lev)
        assert rel_l2(yM2[lev], O.wmat(lev, s, 1) @ f["x2"][lev]) < TOL, ("M2", lev)
        assert rel_l2(yM0[lev], O.pmat(lev, s) @ f["x0"][lev]) < TOL, ("M0", lev)


def test_c4_box_shape_vs_oracle(tmp_path):
    """C4: box/ p=3, 20x20 elements, 40 levels, including the level-0 thickness of box Umat / Wmat (MIMSEM_FIXED_LEVEL)."""
    p, ne, nk = 3, 20, 40
    mesh = mb.Mesh("box", p, ne)
    thick = synthetic_thickness(mesh.xyz, nk, "box")
    eng = mb.Engine.from_mesh(mesh, 0, thick=thick)
    O = _oracle(tmp_path, "box", p, ne, 1, "box")
    O.set_thick(thick)
    rng = np.random.default_rng(1)
    f = synthetic_fields(rng, nk, mesh.N0, mesh.N1, mesh.N2, float(mesh.det.mean()))
    q0 = rng.uniform(-1, 1, (nk, mesh.N0)) * 1e-4
    s = 1.0e8
    FL = mb.engine.FIXED_LEVEL
    yM = _apply(eng, "M1", f["x1"], scale=s, tpow=1, flags=FL)
    yMo = _apply(eng, "M1", f["x1"], scale=s, tpow=0)
    yW = _apply(eng, "M2", f["x2"], scale=s, tpow=1, flags=FL)
    yF = _apply(eng, "M1h", f["x1"], f["h2"], scale=s, tpow=2)
    yK = _apply(eng, "K", f["x1"], f["u1"], scale=s, tpow=2)
    yR = _apply(eng, "R", f["x1"], q0, scale=s, tpow=2)
    # host-buffer entry point with more levels than one pipeline chunk under MIMSEM_FIXED_LEVEL
    assert np.array_equal(eng.apply_host("M1", f["x1"], scale=s, tpow=1, flags=FL), yM)
    for lev in (0, 19, 39):
        assert rel_l2(yM[lev], O.umat(0, s, 1) @ f["x1"][lev]) < TOL, ("Umat.M", lev)
        assert rel_l2(yMo[lev], O.umat(0, s, 0) @ f["x1"][lev]) < TOL, ("Umat.Mo", lev)
        assert rel_l2(yW[lev], O.wmat(0, s, 1) @ f["x2"][lev]) < TOL, ("Wmat.M", lev)
        assert rel_l2(yF[lev], O.umat(lev, s, 1, h2=f["h2"][lev], tpow_h=1) @ f["x1"][lev]) < TOL, ("Uhmat", lev)
        assert rel_l2(yK[lev], O.wtqumat(f["u1"][lev], lev, s) @ f["x1"][lev]) < TOL, ("WtQUmat", lev)
        assert rel_l2(yR[lev], O.rotmat(q0[lev], lev, s, 2) @ f["x1"][lev]) < TOL, ("RotMat", lev)


def test_c2_galewsky_shape_vs_oracle(tmp_path):
    """C2: src/ p=3, 16x16 elements per face, one level: Uhmat, WtQUmat and the PV-upwinded RotMat_up / Phmat::assemble_up."""
    p, ne = 3, 16
    mesh = mb.Mesh("sphere", p, ne, signed_det=True)
    eng = mb.Engine.from_mesh(mesh, 0)
    O = _oracle(tmp_path, "sphere", p, ne, 6, "src")
    rng = np.random.default_rng(2)
    dm = float(np.abs(mesh.det).mean())
    f = synthetic_fields(rng, 1, mesh.N0, mesh.N1, mesh.N2, dm)
    q0 = rng.uniform(-1, 1, (1, mesh.N0)) * 1e-4
    u_up = rng.uniform(-1, 1, (1, mesh.N1)) * dm * 1e-3
    tau = 0.5 * 300.0
    assert rel_l2(_apply(eng, "M1h", f["x1"], f["h2"])[0], O.umat(h2=f["h2"][0]) @ f["x1"][0]) < TOL
    assert rel_l2(_apply(eng, "K", f["x1"], f["u1"])[0], O.wtqumat(f["u1"][0], tpow=0) @ f["x1"][0]) < TOL
    assert rel_l2(_apply(eng, "R", f["x1"], q0)[0], O.rotmat(q0[0]) @ f["x1"][0]) < TOL
    assert rel_l2(_apply(eng, "R_up", f["x1"], q0, u1=u_up, tau=tau)[0], O.rotmat(q0[0], u1=u_up[0], tau=tau) @ f["x1"][0]) < TOL
    assert rel_l2(_apply(eng, "M0h_up", f["x0"], f["h2"], u1=u_up, tau=tau)[0], O.phmat_up(u_up[0], f["h2"][0], tau) @ f["x0"][0]) < TOL


def test_full_size_properties_c5():
    """C5: eul p=4, 48x48 elements/face, 60 levels (26.5 M 1-form DOF-levels) -- properties that need no oracle."""
    import torch
    p, ne, nk = 4, 48, 60
    mesh = mb.Mesh("sphere", p, ne)
    thick = synthetic_thickness(mesh.xyz, nk)
    eng = mb.Engine.from_mesh(mesh, 0, thick=thick)
    g = torch.Generator(device="cuda").manual_seed(3)
    dev = "cuda:0"
    x = torch.rand((mesh.N1, nk), dtype=torch.float64, device=dev, generator=g) * 2 - 1
    z = torch.rand((mesh.N1, nk), dtype=torch.float64, device=dev, generator=g) * 2 - 1
    h = (torch.rand((mesh.N2, nk), dtype=torch.float64, device=dev, generator=g) + 0.5) * 1e4
    s = 1.0e8
    # M1 is symmetric positive definite, level by level
    Mx = eng.apply("M1", x, scale=s, tpow=1)
    Mz = eng.apply("M1", z, scale=s, tpow=1)
    a, b = (z * Mx).sum(0), (x * Mz).sum(0)
    assert float(((a - b).abs() / b.abs()).max()) < 1e-11
    assert float((x * Mx).sum(0).min()) > 0.0
    # linearity
    lin = eng.apply("M1", 2.0 * x - 3.0 * z, scale=s, tpow=1)
    assert float((lin - (2.0 * Mx - 3.0 * Mz)).norm() / lin.norm()) < 1e-13
    # M1(h) with h == const reduces to const * M1 with one more thickness factor handled by tpow
    hc = torch.full_like(h, 7.0)
    # interp2_g of a constant 2-form is not constant (it divides by det), so test symmetry instead
    Fx = eng.apply("M1h", x, coeff=h, scale=s, tpow=2)
    Fz = eng.apply("M1h", z, coeff=h, scale=s, tpow=2)
    a, b = (z * Fx).sum(0), (x * Fz).sum(0)
    assert float(((a - b).abs() / b.abs()).max()) < 1e-11
    del hc
    # K(u) u' is linear in both arguments: K(x) z == K(z) x  (the metric is symmetric)
    Kxz = eng.apply("K", z, coeff=x, scale=s, tpow=2)
    Kzx = eng.apply("K", x, coeff=z, scale=s, tpow=2)
    assert float((Kxz - Kzx).norm() / Kxz.norm()) < 1e-13
    # M2 symmetric
    w = torch.rand((mesh.N2, nk), dtype=torch.float64, device=dev, generator=g)
    v = torch.rand((mesh.N2, nk), dtype=torch.float64, device=dev, generator=g)
    a, b = (v * eng.apply("M2", w, scale=s, tpow=1)).sum(0), (w * eng.apply("M2", v, scale=s, tpow=1)).sum(0)
    assert float(((a - b).abs() / b.abs()).max()) < 1e-12
    # E21 E10 = 0 exactly on integer-valued data; E12 = -E21^T and E01 = -E10^T as adjoints
    n = torch.randint(-1000, 1000, (mesh.N0, 4), device=dev, generator=g).to(torch.float64)
    assert float(eng.apply("E21", eng.apply("E10", n)).abs().max()) == 0.0
    xi = torch.randint(-100, 100, (mesh.N1, 4), device=dev, generator=g).to(torch.float64)
    fi = torch.randint(-100, 100, (mesh.N2, 4), device=dev, generator=g).to(torch.float64)
    assert torch.equal((fi * eng.apply("E21", xi)).sum(0), -(xi * eng.apply("E12", fi)).sum(0))
    assert torch.equal((xi * eng.apply("E10", n)).sum(0), -(n * eng.apply("E01", xi)).sum(0))
    # global integral: sum of M0's diagonal = area of the sphere (reference probe: ratio 0.99999465 at p=3, ne=4)
    one = torch.ones((mesh.N0, 1), dtype=torch.float64, device=dev)
    area = float(eng.apply("M0", one, scale=1.0, tpow=0).sum())
    assert abs(area / (4 * np.pi * 6371220.0 ** 2) - 1.0) < 1e-6


def test_layout_roundtrip():
    import torch
    mesh, eng = _engine("sphere", 3, 4)
    for space, n in ((0, mesh.N0), (1, mesh.N1), (2, mesh.N2)):
        a = torch.rand((7, n), dtype=torch.float64, device="cuda:0")
        c = eng.to_columns(a, space)
        perm = torch.from_numpy(eng.permutation(space).astype(np.int64)).cuda()
        assert sorted(perm.tolist()) == list(range(n))
        ref = torch.empty_like(c)
        ref[perm] = a.t().contiguous()
        assert torch.equal(c, ref)
        assert torch.equal(eng.to_levels(c, space), a)


def test_error_paths():
    mesh, eng = _engine("sphere", 3, 4)
    import torch
    x = torch.zeros((mesh.N1, 2), dtype=torch.float64, device="cuda:0")
    with pytest.raises(mb.MimsemError):
        eng.apply("M1", x, tpow=1)          # no thickness table set
    with pytest.raises(mb.MimsemError):
        eng.apply("M1h", x)                 # missing coefficient
    with pytest.raises(mb.MimsemError):
        eng.apply("M1", x[:10].contiguous())


@pytest.mark.parametrize("kind,p,ne,nk", [("sphere", 3, 6, 30), ("sphere", 4, 4, 60), ("sphere", 2, 3, 8), ("box", 3, 5, 40),
                                          ("sphere", 4, 16, 60), ("sphere", 3, 20, 30)])   # the last two: many tiles per SM (ring wrap-around)
def test_m1_kernel_variants_agree(kind, p, ne, nk, monkeypatch):
    """The TMA tile kernels (one CTA per tile; persistent and double-buffered), the line-task kernel and the thread-per-element kernel compute the
    same M1 / M1(h) (identical up to FP summation order)."""
    mesh = mb.Mesh(kind, p, ne)
    thick = synthetic_thickness(mesh.xyz, nk, kind)
    rng = np.random.default_rng(11)
    f = synthetic_fields(rng, nk, mesh.N0, mesh.N1, mesh.N2, float(mesh.det.mean()))
    res = {}
    for variant in ("0", "1", "2", "3"):
        monkeypatch.setenv("MIMSEM_M1_VARIANT", variant)
        eng = mb.Engine.from_mesh(mesh, 0, thick=thick)
        res[variant] = (_apply(eng, "M1", f["x1"], scale=1e8, tpow=1), _apply(eng, "M1h", f["x1"], f["h2"], scale=1e8, tpow=2),
                        _apply(eng, "M1", f["x1"], scale=1.0, tpow=0))
        eng.close()
    for v in ("1", "2", "3"):
        for a, b in zip(res[v], res["0"]):
            assert rel_l2(a, b) < 1e-14, (v, rel_l2(a, b))
    # the persistent double-buffered kernel (3) runs the tile kernel's (2) arithmetic: bit for bit the same
    for a, b in zip(res["3"], res["2"]):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("variant", [2, 3])
def test_programmatic_dependent_launch_of_independent_applies(variant):
    """"pdl_independent": the caller promises that consecutive launches do not depend on each other; the tile kernels are
    then launched with programmatic stream serialization (a launch starts while the previous one drains).  Eight
    back-to-back applies on distinct fields, eagerly and from one CUDA graph, must equal the ordinary launches bit for bit."""
    import torch
    mesh = mb.Mesh("sphere", 4, 16)
    nk = 60
    eng = mb.Engine.from_mesh(mesh, 0, thick=synthetic_thickness(mesh.xyz, nk))
    eng.set_option("m1_variant", variant)
    g = torch.Generator(device="cuda").manual_seed(3)
    xs = [torch.rand((mesh.N1, nk), dtype=torch.float64, device="cuda", generator=g) for _ in range(8)]
    hs = [torch.rand((mesh.N2, nk), dtype=torch.float64, device="cuda", generator=g) + 0.5 for _ in range(8)]
    def run():
        out = []
        for x, h in zip(xs, hs):
            out.append(eng.apply("M1", x, scale=1e8, tpow=1))
            out.append(eng.apply("M1h", x, coeff=h, scale=1e8, tpow=2))
            out.append(eng.apply("K", x, coeff=x, scale=1e8, tpow=2))
            out.append(eng.apply("M2", h, scale=1e8, tpow=1))
        torch.cuda.synchronize()
        return out
    ref = run()
    eng.set_option("pdl_independent", 1)
    got = run()
    for a, b in zip(ref, got):
        assert torch.equal(a, b)
    ys = [torch.empty_like(xs[0]) for _ in xs]
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        for x, y in zip(xs, ys):
            eng.apply("M1", x, out=y, scale=1e8, tpow=1)
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=st):
            for x, y in zip(xs, ys):
                eng.apply("M1", x, out=y, scale=1e8, tpow=1)
    for y in ys:
        y.zero_()
    gr.replay()
    torch.cuda.synchronize()
    for i, y in enumerate(ys):
        assert torch.equal(y, ref[4 * i])


@pytest.mark.parametrize("kind,p,ne,nk", [("sphere", 3, 6, 30), ("sphere", 4, 4, 60), ("sphere", 2, 3, 8), ("box", 3, 5, 40),
                                          ("sphere", 5, 2, 20)])
def test_m2_kernel_variants_agree(kind, p, ne, nk, monkeypatch):
    """M2 / M2(rho) (Wmat, Whmat): the TMA tile kernel (default) and the thread-per-element-level kernel agree up to FP
    summation order; the tile kernel is pinned to the reference by the golden-vector and oracle tests above."""
    mesh = mb.Mesh(kind, p, ne)
    thick = synthetic_thickness(mesh.xyz, nk, kind)
    rng = np.random.default_rng(13)
    f = synthetic_fields(rng, nk, mesh.N0, mesh.N1, mesh.N2, float(mesh.det.mean()))
    res = {}
    for variant in ("0", "1"):
        monkeypatch.setenv("MIMSEM_M2_VARIANT", variant)
        eng = mb.Engine.from_mesh(mesh, 0, thick=thick)
        res[variant] = (_apply(eng, "M2", f["x2"], scale=1e8, tpow=1), _apply(eng, "M2h", f["x2"], f["h2"], scale=1e8, tpow=2),
                        _apply(eng, "M2", f["x2"], scale=1.0, tpow=0))
        eng.close()
    for a, b in zip(res["1"], res["0"]):
        assert rel_l2(a, b) < 1e-14, rel_l2(a, b)


@pytest.mark.parametrize("kind,p,ne,nk", [("sphere", 3, 6, 30), ("sphere", 4, 4, 60), ("sphere", 2, 3, 8), ("box", 3, 5, 40),
                                          ("sphere", 5, 2, 20)])
def test_k_kernel_variants_agree(kind, p, ne, nk, monkeypatch):
    """K (WtQUmat): the TMA tile kernel (default) and the thread-per-element-level kernel agree up to FP summation order."""
    mesh = mb.Mesh(kind, p, ne)
    thick = synthetic_thickness(mesh.xyz, nk, kind)
    rng = np.random.default_rng(12)
    f = synthetic_fields(rng, nk, mesh.N0, mesh.N1, mesh.N2, float(mesh.det.mean()))
    res = {}
    for variant in ("0", "1"):
        monkeypatch.setenv("MIMSEM_K_VARIANT", variant)
        eng = mb.Engine.from_mesh(mesh, 0, thick=thick)
        l0 = eng.launch_count
        res[variant] = (_apply(eng, "K", f["x1"], f["u1"], scale=1e8, tpow=2), _apply(eng, "K", f["x1"], f["u1"], scale=1.0, tpow=0))
        assert eng.launch_count > l0
        eng.close()
    for a, b in zip(res["1"], res["0"]):
        assert rel_l2(a, b) < 1e-14, rel_l2(a, b)


def test_mass_matrix_solves_vs_reference_matrix():
    """KSPSolve(ksp1 | ksp0) on the mass matrices.  The device CG (M1) and the pointwise M0 solve
    against a sparse direct solve with the REFERENCE's assembled matrices (numpy restatement, pinned to the reference's
    golden vectors in tests/test_oracle.py), p = 3, 4x4 elements per face, 3 levels (odd: register kernels) and the
    same mesh with 4 levels (even: TMA tile kernel inside the iteration)."""
    import scipy.sparse.linalg as spla
    import tempfile
    import os
    from oracle import mimsem_oracle as mo
    g = golden("ops_eul_sphere_p3_ne4.npz")
    s = float(g["scale"])
    with tempfile.TemporaryDirectory() as tmp:
        os.makedirs(os.path.join(tmp, "input"))
        mb.write_input("sphere", 3, 4, 6, os.path.join(tmp, "input"))
        O = mo.Oracle(tmp, 6, "sphere", "eul")
        for nk in (3, 4):
            thick = np.concatenate([g["thick"], g["thick"][:1] * 1.7])[:nk]
            O.set_thick(thick)
            mesh, eng = _engine("sphere", 3, 4, thick=thick)
            rng = np.random.default_rng(21)
            b1 = rng.uniform(-1, 1, (nk, mesh.N1))
            b0 = rng.uniform(-1, 1, (nk, mesh.N0))
            x1, its, rr = eng.solve("M1", to_cols(eng, b1, 1), scale=s, tpow=1, rtol=1e-14, maxit=300)
            assert its < 300 and rr <= 1e-13, (its, rr)
            x1 = to_np(eng, x1, 1)
            x0 = to_np(eng, eng.solve("M0", to_cols(eng, b0, 0), scale=s, tpow=1)[0], 0)
            d1 = to_np(eng, eng.diag("M1", nk, scale=s, tpow=1), 1)
            for lev in range(nk):
                A1 = O.umat(lev, s, 1).tocsc()
                assert rel_l2(x1[lev], spla.spsolve(A1, b1[lev])) < 1e-11, (nk, lev)
                assert rel_l2(d1[lev], A1.diagonal()) < TOL
                A0 = O.pmat(lev, s).tocsc()
                assert rel_l2(x0[lev], spla.spsolve(A0, b0[lev])) < TOL
            eng.close()


def test_mass_matrix_solve_c5_roundtrip():
    """C5 (p=4, 48x48 per face, 60 levels): x -> M1 x -> solve recovers x; iteration count stays mesh-independent."""
    import torch
    p, ne, nk = 4, 48, 60
    mesh = mb.Mesh("sphere", p, ne)
    eng = mb.Engine.from_mesh(mesh, 0, thick=synthetic_thickness(mesh.xyz, nk))
    gen = torch.Generator(device="cuda:0").manual_seed(3)
    x = torch.rand((mesh.N1, nk), dtype=torch.float64, device="cuda:0", generator=gen) * 2 - 1
    b = eng.apply("M1", x, scale=1e8, tpow=1)
    xs, its, rr = eng.solve("M1", b, scale=1e8, tpow=1, rtol=1e-13, maxit=200)
    assert its < 200 and rr <= 1e-12, (its, rr)
    err = float((xs - x).norm() / x.norm())
    assert err < 1e-10, err


@pytest.mark.parametrize("fname,p,ne", [("ops_eul_sphere_p3_ne4.npz", 3, 4), ("ops_eul_sphere_p4_ne2.npz", 4, 2)])
def test_vorticity_term_operators_vs_reference_golden(fname, p, ne):
    """SURVEY section 8f-2, the operators that map onto kernels the device already has (identities pinned against the
    reference in tests/test_oracle.py): Ut_mat::assemble(lev, scale) = M1 with the mean thickness of levels lev, lev+1;
    Ut_mat::assemble_h(lev, scale, rho) = M1(h) without thickness factors; WtQdUdz_mat::assemble(u1, scale) = 2 K(u1)
    without thickness factors (eul/Assembly.cpp:1338-1440, 1581-1640)."""
    g = golden(fname)
    mesh, eng = _engine("sphere", p, ne, thick=g["thick"])
    s = float(g["scale"])
    nk = int(g["nk"])
    y = _apply(eng, "M1h", g["x1"], g["h2"], scale=s, tpow=0)
    assert rel_l2(y, g["y_Ut_mat_h"]) < TOL, rel_l2(y, g["y_Ut_mat_h"])
    y = _apply(eng, "K", g["x1"], g["u1"], scale=2.0 * s, tpow=0)
    assert rel_l2(y, g["y_WtQdUdz_mat"]) < TOL, rel_l2(y, g["y_WtQdUdz_mat"])
    y = _apply(eng, "M1", g["x1"], scale=s, tpow=1, flags=mb.engine.THICK_MEAN)
    assert rel_l2(y[:nk - 1], g["y_Ut_mat"]) < TOL, rel_l2(y[:nk - 1], g["y_Ut_mat"])
    y1 = _apply_per_level(eng, "M1", g["x1"][:nk - 1], scale=s, tpow=1, flags=mb.engine.THICK_MEAN)
    assert rel_l2(y1, g["y_Ut_mat"]) < TOL


@pytest.mark.parametrize("fname,p,ne", [("ops_eul_sphere_p3_ne4.npz", 3, 4), ("ops_eul_sphere_p4_ne2.npz", 4, 2)])
def test_remaining_coefficient_operators(fname, p, ne, tmp_path):
    """UtQWmat (golden vector of the reference), Pvec / Phvec (= the diagonal of Pmat / Pmat::assemble_h,
    eul/Assembly.cpp:602-681) and WmatInv / WhmatInv (element-block inverse of Wmat / Whmat, eul/Assembly.cpp:1658-1800)
    against the oracle's assembled matrices."""
    import scipy.sparse.linalg as spla
    g = golden(fname)
    mesh, eng = _engine("sphere", p, ne, thick=g["thick"])
    s = float(g["scale"])
    nk = int(g["nk"])
    y = _apply(eng, "UtQW", g["x2"], g["u1"], scale=s)
    assert rel_l2(y, g["y_UtQWmat"]) < TOL, rel_l2(y, g["y_UtQWmat"])
    O = _oracle(tmp_path, "sphere", p, ne, 6, "eul")
    O.set_thick(g["thick"])
    dP = to_np(eng, eng.diag("M0", nk, scale=s, tpow=1), 0)
    dPh = to_np(eng, eng.diag("M0h", nk, scale=s, tpow=2, coeff=to_cols(eng, g["h2"], 2)), 0)
    xw = to_np(eng, eng.solve_m2(to_cols(eng, g["x2"], 2), scale=s, tpow=1), 2)
    # a density-like coefficient (positive interpolant: the blocks are SPD as in the model) ...
    rho = 1.0e4 * (1.0 + 0.1 * np.random.default_rng(4).uniform(-1, 1, g["h2"].shape))
    xwh = to_np(eng, eng.solve_m2(to_cols(eng, g["x2"], 2), coeff=to_cols(eng, rho, 2), scale=s, tpow=2), 2)
    # ... and the golden file's rough one (U(0.5, 1.5) per DOF: some blocks are indefinite; the reference inverts them with
    # full pivoting, eul/LinAlg.cpp:186-260, the device factorisation does not pivot)
    xwr = to_np(eng, eng.solve_m2(to_cols(eng, g["x2"], 2), coeff=to_cols(eng, g["h2"], 2), scale=s, tpow=2), 2)
    for lev in range(nk):
        P0 = O.pmat(lev, s)
        assert abs(P0 - sp.diags(P0.diagonal())).max() < 1e-9 * abs(P0).max()      # m == p: Pmat is diagonal, Pvec its diagonal
        assert rel_l2(dP[lev], P0.diagonal()) < TOL
        assert rel_l2(dPh[lev], O.pmat(lev, s, h2=g["h2"][lev]).diagonal()) < TOL
        assert rel_l2(xw[lev], spla.spsolve(O.wmat(lev, s, 1).tocsc(), g["x2"][lev])) < TOL, ("WmatInv", lev)
        assert rel_l2(xwh[lev], spla.spsolve(O.wmat(lev, s, 1, rho=rho[lev], tpow_rho=1).tocsc(), g["x2"][lev])) < TOL, ("WhmatInv", lev)
        assert rel_l2(xwr[lev], spla.spsolve(O.wmat(lev, s, 1, rho=g["h2"][lev], tpow_rho=1).tocsc(), g["x2"][lev])) < 1e-8, ("WhmatInv, rough", lev)


@pytest.mark.parametrize("fname,p,ne", [("ops_eul_sphere_p3_ne4.npz", 3, 4), ("ops_eul_sphere_p4_ne2.npz", 4, 2)])
def test_matrix_free_twins_vs_reference_uvec(fname, p, ne):
    """Uvec::assemble and Uvec::assemble_hu (eul/Assembly.cpp:2124-2279) -- golden vectors produced by the reference's OWN
    matrix-free routines, driven as diagnose_fluxes drives them (eul/HorizSolve.cpp:298-306: four (velocity, density,
    factor) terms accumulated, then the reverse ADD scatter).  On the device the four terms are two M1(h) applies
    (the form is bilinear): F(h1)(u1/3 + u2/6) + F(h2)(u1/6 + u2/3)."""
    g = golden(fname)
    mesh, eng = _engine("sphere", p, ne, thick=g["thick"])
    s = float(g["scale"])
    assert rel_l2(_apply(eng, "M1", g["x1"], scale=s, tpow=1), g["y_Uvec"]) < TOL
    ya = _apply(eng, "M1h", g["x1"] / 3.0 + g["x1b"] / 6.0, g["h2"], scale=s, tpow=2)
    yb = _apply(eng, "M1h", g["x1"] / 6.0 + g["x1b"] / 3.0, g["h2b"], scale=s, tpow=2)
    assert rel_l2(ya + yb, g["y_Uvec_hu"]) < TOL, rel_l2(ya + yb, g["y_Uvec_hu"])
    # term by term, as the host class Uvec does it
    terms = [(g["x1"], g["h2"], 1 / 3), (g["x1"], g["h2b"], 1 / 6), (g["x1b"], g["h2"], 1 / 6), (g["x1b"], g["h2b"], 1 / 3)]
    y4 = sum(_apply(eng, "M1h", u, h, scale=s * f, tpow=2) for u, h, f in terms)
    assert rel_l2(y4, g["y_Uvec_hu"]) < TOL


@pytest.mark.parametrize("kind,p,ne,variant", [("sphere", 3, 4, "eul"), ("sphere", 4, 2, "eul"), ("box", 3, 4, "box")])
def test_element_block_jacobi_vs_reference_matrix(kind, p, ne, variant, tmp_path):
    """PCBJACOBI with one block per element (PCBJacobiSetTotalBlocks(pc, size * nElsX^2, NULL), eul/HorizSolve.cpp:77-84):
    PETSc cuts the assembled Umat into equal consecutive row blocks of 2 p^2 rows -- the edges element e owns in the
    reference's numbering.  The device tabulates, factorises and solves every block itself; checked against a dense solve
    with the same blocks cut out of the oracle's assembled matrix, and as a preconditioner (PCG iteration counts)."""
    g = golden({"eul": "ops_eul_sphere_p%d_ne%d.npz" % (p, ne), "box": "ops_box_p3_ne4.npz"}[variant])
    mesh, eng = _engine(kind, p, ne, thick=g["thick"])
    s = float(g["scale"])
    nk = int(g["nk"])
    O = _oracle(tmp_path, kind, p, ne, 6 if kind == "sphere" else 1, variant)
    O.set_thick(g["thick"])
    r = g["x1"]
    z = to_np(eng, eng.pc_bjacobi(to_cols(eng, r, 1), scale=s, tpow=1), 1)
    nb = 2 * p * p
    for lev in range(nk):
        A = O.umat(lev, s, 1).tocsr()
        ref = np.zeros(mesh.N1)
        for e in range(mesh.nel):
            rows = np.arange(e * nb, (e + 1) * nb)          # PETSc's equal consecutive blocks == the element's owned edges
            ref[rows] = np.linalg.solve(A[rows][:, rows].toarray(), r[lev][rows])
        assert rel_l2(z[lev], ref) < 1e-11, (lev, rel_l2(z[lev], ref))
    # the blocks are those of the owned edges: every owned edge of element e carries a global id in [e nb, (e + 1) nb)
    own = np.concatenate([mesh.el1x.reshape(mesh.nel, p, p + 1)[:, :, :p].reshape(mesh.nel, -1), mesh.el1y[:, :p * p]], axis=1)
    assert np.array_equal(np.sort(own, axis=1), np.arange(mesh.nel * nb).reshape(mesh.nel, nb))


@pytest.mark.parametrize("fname,p,ne", [("ops_eul_sphere_p3_ne4.npz", 3, 4), ("ops_eul_sphere_p4_ne2.npz", 4, 2)])
def test_rayleigh_friction_vs_reference_golden(fname, p, ne):
    """Umat_ray::assemble(lev, scale, dt, exner_k, exner_s) + MatMult (eul/Assembly.cpp:1846-1979; eul/Euler_2.cpp:1218-1229):
    all levels in one launch against vectors of the reference's own class; level by level gives the same."""
    import torch
    g = golden(fname)
    mesh, eng = _engine("sphere", p, ne, thick=g["thick"])
    s, dt = float(g["scale"]), float(g["ray_dt"])
    exs = torch.from_numpy(np.ascontiguousarray(g["ex2"][0])).cuda()
    y = to_np(eng, eng.apply_ray(to_cols(eng, g["x1"], 1), to_cols(eng, g["ex2"], 2), exs, dt, scale=s), 1)
    assert rel_l2(y, g["y_Umat_ray"]) < TOL, rel_l2(y, g["y_Umat_ray"])
    for lev in range(int(g["nk"])):
        yl = to_np(eng, eng.apply_ray(to_cols(eng, g["x1"][lev:lev + 1], 1), to_cols(eng, g["ex2"][lev:lev + 1], 2), exs, dt, lev0=lev,
                                      scale=s), 1)
        assert rel_l2(yl[0], g["y_Umat_ray"][lev]) < TOL


@pytest.mark.parametrize("kind,p,ne,nk", [("sphere", 3, 4, 30), ("sphere", 4, 3, 60), ("box", 3, 5, 7)])
def test_l2vecs_relabelling_bit_exact(kind, p, ne, nk):
    """L2Vecs::HorizToVert / VertToHoriz (eul/L2Vecs.cpp:55-101): vz[e][k*p^2 + i] = vh[k][elInds2_l(e)[i]], bit for bit."""
    import torch
    mesh, eng = _engine(kind, p, ne)
    rng = np.random.default_rng(9)
    vh = rng.uniform(-1, 1, (nk, mesh.N2))
    vz = np.zeros((mesh.nel, nk * p * p))
    for e in range(mesh.nel):                       # the reference's loop, restated
        for k in range(nk):
            vz[e, k * p * p:(k + 1) * p * p] = vh[k, mesh.el2[e]]
    cols = to_cols(eng, vh, 2)
    got = eng.to_vertical(cols)
    assert np.array_equal(got.cpu().numpy(), vz)
    back = eng.from_vertical(torch.from_numpy(vz).cuda())
    assert np.array_equal(to_np(eng, back, 2), vh)


def test_multi_gpu_partitioned_apply():
    """N>1: element-block partition + NCCL ghost refresh, bitwise equal to the single-GPU result (tests/mp_check.py)."""
    import subprocess
    import sys
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (covered on CPU by tests/test_partition.py with gloo)")
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(min(n, 4)), "--master-addr",
           "127.0.0.1", "--master-port", "29517", os.path.join(root, "tests", "mp_check.py")]
    # both hand-over protocols of the fused M1 launch: in-band 16-byte cells (default) and data + flag
    for ll in ("1", "0"):
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, MIMSEM_HALO_LL=ll))
        assert r.returncode == 0 and "MP_CHECK OK" in r.stdout, (ll, r.stdout[-2000:] + r.stderr[-2000:])


def test_cpp_host_layer_one_process(tmp_path):
    """The C++ host layer (DistEngine) as ONE process on one GPU: the same program as the multi-GPU check with a world of one
    -- every operator and the solve through DistEngine::apply / solve_M1 against a second engine, and a burst of M1 launches
    captured into a CUDA graph through the C ABI's stream / graph helpers (mimsem_gpu_stream_create, _graph_begin / _end /
    _launch) and replayed."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.run(["make", "-C", os.path.join(root, "mimsem_b200", "host")], check=True, stdout=subprocess.DEVNULL)
    exe = os.path.join(root, "mimsem_b200", "host", "build", "host_dist_check")
    r = subprocess.run([exe, "sphere", "3", "4", "30", str(tmp_path)], env=dict(os.environ, MIMSEM_RANK="0", MIMSEM_WORLD="1"),
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "HOST_DIST_CHECK OK" in r.stdout and "burst of 6 x M1, field pair 1" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
def test_multi_gpu_cpp_host_layer(tmp_path):
    """N>1 without Python on the data or the control path: the C++ host layer (mimsem_b200/host/DistEngine, Partition) as N
    plain processes with a file rendezvous; all fifteen operators, the partitioned solve and bursts of fused M1 launches, bitwise equal to one GPU
    (mimsem_b200/host/host_dist_check.cpp)."""
    import os
    import subprocess
    import torch
    n = min(torch.cuda.device_count(), 4)
    if n < 2:
        pytest.skip("needs >= 2 GPUs (the partition itself is compared with parallel.py on CPU: tests/test_partition.py)")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.run(["make", "-C", os.path.join(root, "mimsem_b200", "host")], check=True, stdout=subprocess.DEVNULL)
    exe = os.path.join(root, "mimsem_b200", "host", "build", "host_dist_check")
    for cfg in (["sphere", "3", "6", "30"], ["box", "3", "6", "40"]):
        rdv = tmp_path / ("rdv_" + cfg[0])
        rdv.mkdir()
        procs = [subprocess.Popen([exe] + cfg + [str(rdv)], env=dict(os.environ, MIMSEM_RANK=str(r), MIMSEM_WORLD=str(n)),
                                  stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(n)]
        outs = [p.communicate(timeout=600)[0] for p in procs]
        assert all(p.returncode == 0 for p in procs) and "HOST_DIST_CHECK OK" in outs[0], outs
