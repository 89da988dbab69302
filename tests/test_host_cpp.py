"""The C++ host mirror of the reference's classes (mimsem_b200/host/): Umat, Wmat, Pmat, Uhmat, Whmat, WtQUmat,
E10mat, E21mat used exactly as the reference's solvers use them -- constructor, assemble(...), MatMult on the
public Mat (a MatShell) -- for every rank of an emulated `mpirun -np N`, against the reference's golden vectors."""
import os
import subprocess

import numpy as np
import pytest

from helpers import ROOT, TOL, block_jacobi_reference, golden, have_ref_mesh, ref_mesh_dir, rel_l2

HOST = os.path.join(ROOT, "mimsem_b200", "host")
BIN = os.path.join(HOST, "build", "host_apply")


def _build():
    subprocess.run(["make", "-C", HOST], check=True, stdout=subprocess.DEVNULL)


def test_compat_krylov_solver_on_cpu():
    """KSPSolve of the compatibility layer (GMRES(30) with restarts, CG, the four preconditioner routes) on small dense
    shell operators: mimsem_b200/host/compat_ksp_check.cpp.  No GPU."""
    _build()
    r = subprocess.run([os.path.join(HOST, "build", "compat_ksp_check")], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "compat_ksp_check ok" in r.stdout, r.stdout + r.stderr


@pytest.mark.parametrize("kind,p,ne,nprocs", [(0, 3, 4, 6), (0, 3, 4, 24), (0, 4, 2, 6), (0, 3, 6, 54), (1, 3, 4, 1), (1, 3, 4, 4)])
def test_block_jacobi_context_tables_on_cpu(kind, p, ne, nprocs):
    """Host half of the element-block Jacobi preconditioner of the Umat shell (PCBJACOBI, eul/HorizSolve.cpp:77-84): every
    rank's patch plus copies of the west / south neighbours across the patch boundary -- on another rank of the cubed
    sphere (across cube seams and inside a face) or at the other end of a periodic box patch; placement checked against
    the point coordinates: mimsem_b200/host/host_pc_tables_check.cpp.  No GPU."""
    _build()
    r = subprocess.run([os.path.join(HOST, "build", "host_pc_tables_check"), str(kind), str(p), str(ne), str(nprocs), "2"],
                       capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "host_pc_tables_check ok" in r.stdout, r.stdout + r.stderr


@pytest.mark.parametrize("p,ne", [(3, 4), (4, 2)])
def test_quadrature_projections_vs_reference_golden(tmp_path, p, ne):
    """WtQmat, UtQmat, PtQmat of the mirror (host loops; the reference initialises its fields with them, eul/Euler_2.cpp:432,
    493, 535) on six emulated ranks against vectors of the reference's own classes (tests/golden/make_golden_quadproj.py), and
    Geom::writeVertToHoriz (eul/Geom.cpp:633-679) against write2, file for file.  No GPU."""
    _build()
    g = golden("quadproj_eul_sphere_p%d_ne%d.npz" % (p, ne))
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    np.concatenate([g["xq"], g["uq"]]).astype("<f8").tofile(fin)
    r = subprocess.run([os.path.join(HOST, "build", "host_quadproj_check"), str(p), str(ne), fin, fout, str(tmp_path)], capture_output=True,
                       text=True, timeout=120)
    assert r.returncode == 0 and "host_quadproj_check ok" in r.stdout, r.stdout + r.stderr
    out = np.fromfile(fout, dtype="<f8")
    N0, N1, N2 = int(g["N0"]), int(g["N1"]), int(g["N2"])
    assert out.size == N0 + N1 + N2 + 2
    assert rel_l2(out[:N2], g["y_WtQmat"]) < 1e-14
    assert rel_l2(out[N2:N2 + N1], g["y_UtQmat"]) < 1e-14
    assert rel_l2(out[N2 + N1:N2 + N1 + N0], g["y_PtQmat"]) < 1e-14
    assert out[-2] == 1.0, "writeVertToHoriz and write2 wrote different files"
    assert out[-1] == 1.0, "the level-less src/ writers (Geom::write0/1/2(Vec, char*, int)) differ from the eul/ writers at unit thickness"


def test_element_tabulations_row_view_on_cpu():
    """src/ and box/ callers index the element tabulations as rows (double** A); the mirror keeps both views in every object
    (host/ElMats.h, -DMIMSEM_ELMATS_ROWS): mimsem_b200/host/elmats_rows_check.cpp, compiled with the flag against the
    library compiled without.  No GPU."""
    _build()
    r = subprocess.run([os.path.join(HOST, "build", "elmats_rows_check")], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and "elmats_rows_check ok" in r.stdout, r.stdout + r.stderr


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
    sizes = [("Umat", N1), ("BJacobi", N1), ("Wmat", N2), ("Pmat", N0), ("Pmat_h", N0), ("Uhmat", N1), ("Whmat", N2), ("WtQUmat", N2), ("E21", N2),
             ("E12", N1), ("E10", N1), ("E01", N0)]
    per_lev = sum(n for _, n in sizes)
    batched = [("Umat", N1), ("Wmat", N2), ("Pmat", N0), ("Uhmat", N1), ("WtQUmat", N2), ("E21", N2)]
    assert out.size == nk * (per_lev + sum(n for _, n in batched))
    res = {k: [] for k, _ in sizes}
    for lev in range(nk):
        o = lev * per_lev
        for k, n in sizes:
            res[k].append(out[o:o + n])
            o += n
    res = {k: np.array(v) for k, v in res.items()}
    # all levels in one device call per operator (MimsemMatMultLevels) against level by level (a launch over several
    # levels may run another kernel variant than a one-level launch: equal to rounding, not to the bit)
    o = nk * per_lev
    for k, n in batched:
        d = rel_l2(out[o:o + nk * n].reshape(nk, n), res[k])
        assert d < 1e-14, ("MimsemMatMultLevels", k, d)
        o += nk * n
    return res


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
    # PCBJACOBI of the Umat shell (one block per element, eul/HorizSolve.cpp:77-84) on six patches: the blocks of the
    # elements along a patch's west / south boundary carry the far-line terms of an element on ANOTHER rank
    if have_ref_mesh("sphere", 3, 4, 6):
        from oracle import mimsem_oracle as mo
        O = mo.Oracle(ref_mesh_dir("sphere", 3, 4, 6), 6, "sphere", "eul")
        O.set_thick(g["thick"])
        for lev in range(3):
            ref = block_jacobi_reference(O.umat(lev, float(g["scale"]), 1), g["x1"][lev], 2 * 3 * 3)
            assert rel_l2(res["BJacobi"][lev], ref) < 1e-11, (lev, rel_l2(res["BJacobi"][lev], ref))


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
        # element blocks in the np = 24 numbering; patches now also meet INSIDE a cube face
        ref = block_jacobi_reference(O.umat(lev, 1e8, 1), g["x1"][lev], 2 * 3 * 3)
        assert rel_l2(res["BJacobi"][lev], ref) < 1e-11, (lev, rel_l2(res["BJacobi"][lev], ref))
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


REFERENCE = os.environ.get("MIMSEM_REFERENCE", "/root/reference")
# what the host mirror replaces inside each directory of the reference (everything else is taken from the reference)
MIRRORED = {"Basis.h", "Topo.h", "Geom.h", "ElMats.h", "Assembly.h", "Basis.cpp", "Topo.cpp", "Geom.cpp", "ElMats.cpp", "Assembly.cpp"}
PETSC_HEADERS = ("petsc.h", "petscis.h", "petscvec.h", "petscmat.h", "petscksp.h", "petscpc.h", "petscviewerhdf5.h", "mpi.h")


@pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, "eul")), reason="the reference sources are only mounted in the build container")
@pytest.mark.parametrize("variant,caller,decls_only", [("eul", "HorizSolve.cpp", False), ("box", "HorizSolve.cpp", False),
                                                       ("eul", "L2Vecs.cpp", False), ("eul", "Euler_2.cpp", True),
                                                       ("eul", "VertSolve.cpp", True), ("eul", "VertOps.cpp", True),
                                                       ("eul", "UMJS14.cpp", True), ("box", "Euler_2.cpp", True),
                                                       ("box", "VertSolve.cpp", True), ("box", "VertOps.cpp", True),
                                                       ("box", "Bubble.cpp", True), ("eul", "HeldSuarez.cpp", True),
                                                       ("src", "Williamson2.cpp", True), ("src", "Williamson5.cpp", True)])
def test_reference_caller_compiles_unchanged_against_the_mirror(tmp_path, variant, caller, decls_only):
    """The drop-in claim, mechanically: the reference's own caller translation unit -- eul/HorizSolve.cpp (constructs Umat,
    Wmat, Pmat, Uhmat, WtQUmat, RotMat, Ut_mat, UtQWmat, Whmat, E10mat, E21mat, Uvec, Wvec, PtQmat; calls assemble(...) /
    assemble_hu(...), reads ->M, ->vl, ->vg, KSPSolve on M1 and M0) and box/HorizSolve.cpp (Umat / Wmat with M and Mo, the
    box signatures of Uvec::assemble_hu and Wvec::assemble_K) -- compiles UNCHANGED against mimsem_b200/host/*.h (the
    reference's other headers stay the reference's), and every symbol it needs from the mirrored classes and from the
    PETSc subset is defined by libmimsem_host.so.  The sources are reached through symbolic links in a scratch directory
    (a quoted #include looks beside the including file first); nothing is copied.
    decls_only: eul/Euler_2.cpp (constructs WtQmat, UtQmat, Umat_ray; MatAXPY of the friction matrix into M1->M; Geom::
    writeVertToHoriz), eul/VertSolve.cpp and the driver eul/UMJS14.cpp also build the vertical solver's ASSEMBLED SeqAIJ
    matrices -- PETSc proper, outside the path and not part of the compatibility layer: for these the PETSc headers
    forward to tests/petsc_decls_only.h (prototypes without bodies) and the symbol check covers the mirrored classes.
    box/ (like src/) keeps its element tabulations as rows (double** A, box/ElMats.h): its callers are compiled with
    -DMIMSEM_ELMATS_ROWS, which names the row view of the mirror's tabulations `A` (host/ElMats.h).  Of src/ only the
    drivers are here (Topo, Geom incl. the level-less write0/1/2): src/SWEqn_Picard.cpp feeds the operator matrices to
    MatMatMult / MatGetRow, i.e. needs assembled matrices."""
    _build()
    src = os.path.join(REFERENCE, variant)
    for f in os.listdir(src):
        if f not in MIRRORED and (f.endswith(".h") or f == caller):
            os.symlink(os.path.join(src, f), str(tmp_path / f))
    for f in PETSC_HEADERS + ("petscsnes.h",):
        (tmp_path / f).write_text('#include "%s"\n' % ("petsc_decls_only.h" if decls_only else "petsc_compat.h"))
    obj = str(tmp_path / "caller.o")
    r = subprocess.run(["g++", "-std=c++11", "-w", "-c"] + (["-DMIMSEM_ELMATS_ROWS"] if variant in ("box", "src") else []) +
                       ["-I", str(tmp_path), "-I", HOST, "-I", os.path.join(ROOT, "tests"), str(tmp_path / caller), "-o", obj],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    undef = subprocess.run(["nm", "-u", "-C", obj], capture_output=True, text=True, check=True).stdout
    need = [l.split(None, 1)[1].strip() for l in undef.splitlines() if l.strip().startswith("U ")]
    lib = os.path.join(ROOT, "mimsem_b200", "libmimsem_host.so")
    have = subprocess.run(["nm", "-D", "-C", "--defined-only", lib], capture_output=True, text=True, check=True).stdout
    defined = {l.split(None, 2)[2].strip() for l in have.splitlines() if len(l.split(None, 2)) == 3}
    classes = ("Umat", "Wmat", "Pmat", "Uhmat", "Whmat", "WtQUmat", "RotMat", "Ut_mat", "UtQWmat", "WtQdUdz_mat", "E10mat", "E21mat", "Uvec",
               "Wvec", "PtQmat", "WtQmat", "UtQmat", "Umat_ray", "Pvec", "Phvec", "WmatInv", "WhmatInv", "Topo", "Geom", "GaussLobatto", "LagrangeNode", "LagrangeEdge",
               "M1x_j_xy_i", "M1y_j_xy_i", "M2_j_xy_i", "M0_j_xy_i", "Wii")
    petsc = () if decls_only else ("Vec", "Mat", "KSP", "PC", "IS", "MPI_Comm_")
    mine = [s for s in need if s.split("::")[0] in classes or (petsc and s.split("(")[0].startswith(petsc) and "::" not in s.split("(")[0])]
    floor = {"HorizSolve.cpp": 20, "Euler_2.cpp": 20}.get(caller, 2)
    assert len(mine) > floor, need                   # the caller really uses the mirrored surface
    missing = [s for s in mine if s not in defined]
    assert not missing, missing


@pytest.mark.gpu
@pytest.mark.parametrize("fname,p,ne", [("ops_eul_sphere_p3_ne4.npz", 3, 4), ("ops_eul_sphere_p4_ne2.npz", 4, 2)])
def test_host_twins_and_remaining_classes_vs_reference(tmp_path, fname, p, ne):
    """Uvec (assemble, assemble_hu exactly as diagnose_fluxes calls it), UtQWmat, Pvec, Phvec, WmatInv, WhmatInv, Umat_ray through the
    C++ mirror on six emulated ranks against the reference's golden vectors / the oracle's matrices, and KSPSolve on the
    Umat shell of the periodic box (GMRES + block Jacobi requested as eul/HorizSolve.cpp:77-84 does)."""
    import scipy.sparse.linalg as spla
    from oracle import mimsem_oracle as mo
    _build()
    g = golden(fname)
    nk = int(g["nk"])
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    np.concatenate([g[k].ravel() for k in ("thick", "x1", "x1b", "x2", "h2", "h2b", "u1", "ex2")]).astype("<f8").tofile(fin)
    r = subprocess.run([BIN + "_twins", str(p), str(ne), str(nk), fin, fout], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "host_apply_twins ok" in r.stdout, r.stdout + r.stderr
    out = np.fromfile(fout, dtype="<f8")
    N0, N1, N2 = int(g["N0"]), int(g["N1"]), int(g["N2"])
    s = float(g["scale"])
    O = mo.Oracle(ref_mesh_dir("sphere", p, ne, 6), 6, "sphere", "eul") if have_ref_mesh("sphere", p, ne, 6) else None
    if O is not None:
        O.set_thick(g["thick"])
    o = 0

    def take(n):
        nonlocal o
        v = out[o:o + n]
        o += n
        return v
    for lev in range(nk):
        assert rel_l2(take(N1), g["y_Uvec"][lev]) < TOL, ("Uvec::assemble", lev)
        assert rel_l2(take(N1), g["y_Uvec_hu"][lev]) < TOL, ("Uvec::assemble_hu", lev)
        assert rel_l2(take(N1), g["y_UtQWmat"][lev]) < TOL, ("UtQWmat", lev)
        pv, phv, wi, whi = take(N0), take(N0), take(N2), take(N2)
        if O is not None:
            assert rel_l2(pv, O.pmat(lev, s).diagonal()) < TOL, ("Pvec", lev)
            assert rel_l2(phv, O.pmat(lev, s, h2=g["h2"][lev]).diagonal()) < TOL, ("Phvec", lev)
            assert rel_l2(wi, spla.spsolve(O.wmat(lev, s, 1).tocsc(), g["x2"][lev])) < TOL, ("WmatInv", lev)
            assert rel_l2(whi, spla.spsolve(O.wmat(lev, s, 1, rho=g["h2b"][lev], tpow_rho=1).tocsc(), g["x2"][lev])) < 1e-8, ("WhmatInv", lev)
        assert rel_l2(take(N1), g["y_Umat_ray"][lev]) < TOL, ("Umat_ray", lev)
        # MatAXPY(M1->M, 1.0, M1ray->M, ...) as eul/Euler_2.cpp:1229 adds the friction to the mass matrix; the next assemble() drops it
        assert rel_l2(take(N1), g["y_Umat_vs1"][lev] + g["y_Umat_ray"][lev]) < TOL, ("Umat + Umat_ray (MatAXPY)", lev)
        assert rel_l2(take(N1), g["y_Umat_vs1"][lev]) < TOL, ("Umat after re-assembly", lev)
    its, err, its_diag = take(3)
    # KSPSolve(ksp0, ...) on the Pmat shell (eul/HorizSolve.cpp:87-96): M0 is diagonal, the block-Jacobi request is its exact inverse
    its0, err0 = take(2)
    assert 1 <= its0 <= 2 and err0 < 1e-13, (its0, err0)
    # the box Pvec (box/Assembly.cpp:357-372): vg = SCALE vg1 = the diagonal 0-form mass matrix of level 0
    nb0 = (3 * 4) ** 2
    vg, vg1 = take(nb0), take(nb0)
    assert rel_l2(vg, 1.0e8 * vg1) < 1e-15 and vg1.min() > 0.0
    if have_ref_mesh("box", 3, 4, 1):
        Ob = mo.Oracle(ref_mesh_dir("box", 3, 4, 1), 1, "box", "box")
        Ob.set_thick(np.array([750.0 * (1.0 + 0.05 * ((np.arange(nb0) * 7 + lev) % 5)) for lev in range(2)]))
        assert rel_l2(vg1, Ob.pmat(0, 1.0).diagonal()) < TOL, "box Pvec::vg1"
    assert o == out.size
    # GMRES(30) + the element blocks of the periodic box (every west / south neighbour sits at the other end of the same
    # patch) against GMRES(30) + the diagonal
    assert 0 < its < its_diag < 200 and err < 1e-11, (its, err, its_diag)


def test_host_geom_interp_topog_and_writers_vs_reference(tmp_path):
    """SURVEY section 8a G4 / G5 and 8f-4 without a GPU: the host Geom's interp0 / interp1_l / interp2_l / interp1_g /
    interp2_g and initTopog against values computed by the REFERENCE's own Geom (tests/golden/geom_*.npz, generated by
    tests/golden/make_golden.py through oracle/_ref), the field writers write0/1/2 (ASCII VecView layout) and the PETSc
    binary Vec format of the restart files (big-endian classid 1211214, length, float64 payload)."""
    _build()
    g = golden("geom_eul_sphere_p3_ne4_rank2.npz")
    p, ne, nprocs, nk, rank = (int(g[k]) for k in ("p", "ne", "nprocs", "nk", "rank"))
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    np.concatenate([g["v0"], g["v1"], g["v2"]]).astype("<f8").tofile(fin)
    r = subprocess.run([os.path.join(HOST, "build", "host_geom_check"), str(p), str(ne), str(nprocs), str(nk), str(rank), fin, fout,
                        str(tmp_path)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "host_geom_check ok" in r.stdout, r.stdout + r.stderr
    out = np.fromfile(fout, dtype="<f8")
    o = 0
    for key in ("interp0", "interp1_l", "interp2_l", "interp1_g", "interp2_g", "thick"):
        ref = g[key]
        got = out[o:o + ref.size].reshape(ref.shape)
        o += ref.size
        assert rel_l2(got, ref) < 1e-14, (key, rel_l2(got, ref))
    assert o == out.size
    # the restart file is PETSc's binary Vec: readable with numpy as big-endian
    raw = (tmp_path / "output" / "rho_001_0007.vec").read_bytes()
    cid, n = np.frombuffer(raw[:8], dtype=">i4")
    assert cid == 1211214 and n == 6 * (p * ne) ** 2
    vals = np.frombuffer(raw[8:], dtype=">f8")
    assert vals.size == n and np.array_equal(vals, 1.0e4 * (1.0 + 0.001 * (np.arange(n) % 97)) + 0.25)
    # write2 with vert_scale: interp2_g(h) / thick at every quadrature point, one value per line after the header
    lines = (tmp_path / "output" / "rho_001_0007.dat").read_text().splitlines()
    body = [float(x) for x in lines if x and x[0] in "-0123456789"]
    assert len(body) == 6 * (p * ne) ** 2 + 2 and min(body) > 0.0
    for nm in ("velocity_x_001_0007.dat", "velocity_y_001_0007.dat", "velocity_001_0007.vec", "vorticity_001_0007.dat"):
        assert (tmp_path / "output" / nm).exists(), nm
