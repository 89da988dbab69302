"""The C++ host mirror of the reference's classes (mimsem_b200/host/): Umat, Wmat, Pmat, Uhmat, Whmat, WtQUmat,
E10mat, E21mat used exactly as the reference's solvers use them -- constructor, assemble(...), MatMult on the
public Mat (a MatShell) -- for every rank of an emulated `mpirun -np N`, against the reference's golden vectors."""
import os
import subprocess

import numpy as np
import pytest

from helpers import ROOT, TOL, golden, have_ref_mesh, ref_mesh_dir, rel_l2

HOST = os.path.join(ROOT, "mimsem_b200", "host")
BIN = os.path.join(HOST, "build", "host_apply")


def _build():
    subprocess.run(["make", "-C", HOST], check=True, stdout=subprocess.DEVNULL)


def test_host_library_builds_and_links():
    _build()
    assert os.path.exists(BIN) and os.path.exists(os.path.join(ROOT, "mimsem_b200", "libmimsem_host.so"))


def _run(tmp_path, g, kind, p, ne, nprocs, nk, meshdir="-"):
    _build()
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    np.concatenate([g[k].ravel() for k in ("thick", "x1", "x2", "x0", "h2", "u1")]).astype("<f8").tofile(fin)
    r = subprocess.run([BIN, meshdir, str(kind), str(p), str(ne), str(nprocs), str(nk), fin, fout], capture_output=True, text=True,
                       timeout=300)
    assert r.returncode == 0 and "host_apply ok" in r.stdout, r.stdout + r.stderr
    out = np.fromfile(fout, dtype="<f8")
    N0, N1, N2 = int(g["N0"]), int(g["N1"]), int(g["N2"])
    sizes = [("Umat", N1), ("Wmat", N2), ("Pmat", N0), ("Pmat_h", N0), ("Uhmat", N1), ("Whmat", N2), ("WtQUmat", N2), ("E21", N2),
             ("E12", N1), ("E10", N1), ("E01", N0)]
    per_lev = sum(n for _, n in sizes)
    assert out.size == nk * per_lev
    res = {k: [] for k, _ in sizes}
    for lev in range(nk):
        o = lev * per_lev
        for k, n in sizes:
            res[k].append(out[o:o + n])
            o += n
    return {k: np.array(v) for k, v in res.items()}


@pytest.mark.gpu
@pytest.mark.parametrize("from_files", [False, True])
def test_host_classes_vs_reference_golden(tmp_path, from_files):
    import scipy.sparse as sp
    g = golden("ops_eul_sphere_p3_ne4.npz")
    meshdir = "-"
    if from_files:
        if not have_ref_mesh("sphere", 3, 4, 6):
            pytest.skip("reference-generated mesh files not present")
        meshdir = ref_mesh_dir("sphere", 3, 4, 6)     # Topo / Geom read the reference's own input/*.txt
    res = _run(tmp_path, g, 0, 3, 4, 6, 3, meshdir)
    for k, key in (("Umat", "y_Umat_vs1"), ("Wmat", "y_Wmat_vs1"), ("Pmat", "y_Pmat"), ("Pmat_h", "y_Pmat_h"), ("Uhmat", "y_Uhmat_cv1"),
                   ("Whmat", "y_Whmat_vs1"), ("WtQUmat", "y_WtQUmat")):
        assert rel_l2(res[k], g[key]) < TOL, (k, rel_l2(res[k], g[key]))
    for k, xk in (("E21", "x1"), ("E12", "x2"), ("E10", "x0"), ("E01", "x1")):
        A = sp.csr_matrix((g[k + "_data"], g[k + "_indices"], g[k + "_indptr"]))
        assert rel_l2(res[k], (A @ g[xk].T).T) < TOL, k


@pytest.mark.gpu
def test_host_classes_24_ranks_vs_oracle(tmp_path):
    """np = 24 (2x2 patches per face): a different global edge numbering; oracle = numpy restatement."""
    import mimsem_b200 as mb
    from oracle import mimsem_oracle as mo
    d = tmp_path / "mesh" / "input"
    d.mkdir(parents=True)
    mb.write_input("sphere", 3, 4, 24, str(d))
    O = mo.Oracle(str(tmp_path / "mesh"), 24, "sphere", "eul")
    rng = np.random.default_rng(3)
    nk = 2
    g = dict(N0=O.N0, N1=O.N1, N2=O.N2, thick=rng.uniform(100, 200, (nk, O.N0)), x1=rng.uniform(-1, 1, (nk, O.N1)),
             x2=rng.uniform(-1, 1, (nk, O.N2)), x0=rng.uniform(-1, 1, (nk, O.N0)), h2=rng.uniform(0.5, 1.5, (nk, O.N2)) * 1e4,
             u1=rng.uniform(-1, 1, (nk, O.N1)) * 1e12)
    O.set_thick(g["thick"])
    res = _run(tmp_path, g, 0, 3, 4, 24, nk, str(tmp_path / "mesh"))
    for lev in range(nk):
        assert rel_l2(res["Umat"][lev], O.umat(lev, 1e8, 1) @ g["x1"][lev]) < TOL
        assert rel_l2(res["Uhmat"][lev], O.umat(lev, 1e8, 1, h2=g["h2"][lev], tpow_h=1) @ g["x1"][lev]) < TOL
        assert rel_l2(res["WtQUmat"][lev], O.wtqumat(g["u1"][lev], lev, 1e8) @ g["x1"][lev]) < TOL
        assert rel_l2(res["Pmat"][lev], O.pmat(lev, 1e8) @ g["x0"][lev]) < TOL
        assert rel_l2(res["Wmat"][lev], O.wmat(lev, 1e8, 1) @ g["x2"][lev]) < TOL
    E10, E01 = O.e10()
    E21, E12 = O.e21()
    assert rel_l2(res["E01"], (E01 @ g["x1"].T).T) < TOL and rel_l2(res["E12"], (E12 @ g["x2"].T).T) < TOL
    assert rel_l2(res["E10"], (E10 @ g["x0"].T).T) < TOL and rel_l2(res["E21"], (E21 @ g["x1"].T).T) < TOL


@pytest.mark.gpu
@pytest.mark.parametrize("from_files", [False, True])
def test_host_src_upwinded_classes_vs_reference_golden(tmp_path, from_files):
    """BASELINE config 2: RotMat, RotMat_up and Phmat::assemble / assemble_up with the reference's src/ signatures
    (src/Assembly.h:37-49, 227-250), 6 emulated ranks, against vectors produced by the reference's own sources."""
    _build()
    g = golden("ops_src_sphere_p3_ne4.npz")
    meshdir = "-"
    if from_files:
        if not have_ref_mesh("sphere", 3, 4, 6):
            pytest.skip("reference-generated mesh files not present")
        meshdir = ref_mesh_dir("sphere", 3, 4, 6)
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    np.concatenate([g[k][0] for k in ("x1", "x0", "h2", "q0", "u1_up")]).astype("<f8").tofile(fin)
    r = subprocess.run([BIN + "_src", meshdir, "3", "4", "6", repr(float(g["up_fac"])), repr(float(g["up_dt"])), fin, fout],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "host_apply_src ok" in r.stdout, r.stdout + r.stderr
    out = np.fromfile(fout, dtype="<f8")
    N0, N1 = int(g["N0"]), int(g["N1"])
    assert out.size == 2 * N1 + 2 * N0
    y_R, y_Rup, y_Ph, y_Phup = out[:N1], out[N1:2 * N1], out[2 * N1:2 * N1 + N0], out[2 * N1 + N0:]
    assert rel_l2(y_R, g["y_RotMat"][0]) < TOL, rel_l2(y_R, g["y_RotMat"][0])
    assert rel_l2(y_Rup, g["y_RotMat_up"][0]) < TOL, rel_l2(y_Rup, g["y_RotMat_up"][0])
    assert rel_l2(y_Phup, g["y_Phmat_up"][0]) < TOL, rel_l2(y_Phup, g["y_Phmat_up"][0])
    assert rel_l2(y_Ph, y_Phup) > 1e-6      # the upwinding changes the operator


@pytest.mark.gpu
def test_host_vorticity_term_classes_vs_reference_golden(tmp_path):
    """Ut_mat (assemble, assemble_h) and WtQdUdz_mat with the reference's eul/ signatures (eul/Assembly.h:201-229, 257-280),
    6 emulated ranks, against vectors produced by the reference's own sources."""
    _build()
    g = golden("ops_eul_sphere_p3_ne4.npz")
    nk = int(g["nk"])
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    np.concatenate([g[k].ravel() for k in ("thick", "x1", "h2", "u1")]).astype("<f8").tofile(fin)
    r = subprocess.run([BIN + "_vort", "-", "3", "4", "6", str(nk), fin, fout], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "host_apply_vort ok" in r.stdout, r.stdout + r.stderr
    out = np.fromfile(fout, dtype="<f8")
    N1, N2 = int(g["N1"]), int(g["N2"])
    o = 0
    for lev in range(nk):
        if lev < nk - 1:
            assert rel_l2(out[o:o + N1], g["y_Ut_mat"][lev]) < TOL, ("Ut_mat", lev)
            o += N1
        assert rel_l2(out[o:o + N1], g["y_Ut_mat_h"][lev]) < TOL, ("Ut_mat_h", lev)
        o += N1
        assert rel_l2(out[o:o + N2], g["y_WtQdUdz_mat"][lev]) < TOL, ("WtQdUdz_mat", lev)
        o += N2
    assert o == out.size
