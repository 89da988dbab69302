"""CPU tests (-m "not gpu"): host logic of the product against the golden fixtures generated from
the reference (tests/golden/make_golden.py), the oracle restatement against the same fixtures, and
the C-ABI surface of libmimsem_gpu.so."""
import ctypes
import hashlib
import json
import os
import re

import numpy as np
import pytest

import mimsem_b200 as mb
from mimsem_b200 import lib as mlib
from helpers import GOLDEN, ROOT, golden, have_ref_lib, have_ref_mesh, ref_mesh_dir, rel_l2

TOPO_CASES = [("sphere", 3, 4, 6), ("sphere", 3, 4, 24), ("sphere", 4, 2, 6), ("sphere", 2, 2, 6), ("box", 3, 4, 4),
              ("box", 3, 4, 1)]


def test_library_exports_every_declared_symbol():
    """include/mimsem_gpu.h and the built library agree (no compute call is made here)."""
    hdr = open(os.path.join(ROOT, "include", "mimsem_gpu.h")).read()
    declared = set(re.findall(r"\b(mimsem_[a-z0-9_A-Z]+)\s*\(", hdr))
    declared -= {"mimsem_mesh", "mimsem_gpu_ctx"}
    assert declared == set(mlib.SIGNATURES), declared ^ set(mlib.SIGNATURES)
    lib = ctypes.CDLL(mlib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name


def test_header_is_plain_c(tmp_path):
    """The drop-in boundary is a C ABI: include/mimsem_gpu.h compiles as strict C99 (plain pointers and sizes, no C++ or
    torch types in any signature)."""
    import subprocess
    src = tmp_path / "abi.c"
    src.write_text('#include "mimsem_gpu.h"\nint main(void) { return 0; }\n')
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "include"),
                        str(src)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_no_gpu_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(mb.MimsemError):
        mb.Engine(0)
    L = mb.load_library()
    h = ctypes.c_void_p()
    assert L.mimsem_gpu_create(0, ctypes.byref(h)) != 0
    assert b"no CPU fallback" in L.mimsem_last_error()


@pytest.mark.parametrize("kind,p,ne,nprocs", TOPO_CASES)
def test_topology_bit_exact_vs_reference_files(kind, p, ne, nprocs):
    """Topo maps are bit-exact against what the reference's scr/Setup*.py wrote (golden copy)."""
    g = golden("topo_%s_p%d_ne%d_np%d.npz" % (kind, p, ne, nprocs))
    for r in range(nprocs):
        t = mb.patch_topology(kind, p, ne, nprocs, r)
        for key in ("loc0", "loc1x", "loc1y", "loc2"):
            assert np.array_equal(t[key], g["%s_%d" % (key, r)]), (r, key)
        assert np.array_equal(t["local_sizes"], g["sizes_%d" % r])


@pytest.mark.parametrize("name", ["sphere_p3_ne12_np6", "sphere_p3_ne16_np6", "sphere_p4_ne48_np6", "sphere_p4_ne4_np24",
                                  "box_p3_ne20_np1"])
def test_topology_digest_at_baseline_sizes(name):
    """Same check at BASELINE.json's mesh sizes through a sha256 of all maps (C2, C3, C4, C5)."""
    dig = json.load(open(os.path.join(GOLDEN, "topo_sha256.json")))
    kind, p, ne, nprocs = re.match(r"(\w+)_p(\d+)_ne(\d+)_np(\d+)", name).groups()
    p, ne, nprocs = int(p), int(ne), int(nprocs)
    h = hashlib.sha256()
    for r in range(nprocs):
        t = mb.patch_topology(kind, p, ne, nprocs, r)
        for key in ("loc0", "loc1x", "loc1y", "loc2", "local_sizes"):
            h.update(t[key].astype("<i4").tobytes())
    assert h.hexdigest() == dig[name]


def test_write_input_roundtrip(tmp_path):
    """The input/*.txt writer reproduces the reference's files (integers exactly, coordinates to 1e-15)."""
    d = tmp_path / "input"
    d.mkdir()
    mb.write_input("sphere", 3, 4, 6, str(d))
    g = golden("topo_sphere_p3_ne4_np6.npz")
    for r in range(6):
        for key, stem in (("loc0", "nodes"), ("loc1x", "edges_x"), ("loc1y", "edges_y"), ("loc2", "faces"),
                          ("sizes", "local_sizes")):
            got = np.loadtxt(d / ("%s_%04d.txt" % (stem, r)), dtype=np.int64)
            assert np.array_equal(got, g["%s_%d" % (key, r)])
    assert open(d / "grid_res.txt").read().split() == ["3", "4"]
    assert open(d / "grid_res_quad.txt").read().split() == ["3", "4"]
    if have_ref_mesh("sphere", 3, 4, 6):
        ref = ref_mesh_dir("sphere", 3, 4, 6)
        for r in range(6):
            a = np.loadtxt(d / ("geom_%04d.txt" % r))
            b = np.loadtxt(os.path.join(ref, "input", "geom_%04d.txt" % r))
            assert np.abs(a - b).max() <= 1e-15 * 6371220.0 * 4
            assert np.array_equal(np.loadtxt(d / ("quads_%04d.txt" % r)), np.loadtxt(os.path.join(ref, "input", "quads_%04d.txt" % r)))


@pytest.mark.parametrize("fname,kind,signed", [("ops_eul_sphere_p3_ne4.npz", "sphere", False),
                                               ("ops_eul_sphere_p4_ne2.npz", "sphere", False),
                                               ("ops_src_sphere_p3_ne4.npz", "sphere", True),
                                               ("ops_box_p3_ne4.npz", "box", False)])
def test_basis_and_geometry_vs_reference(fname, kind, signed):
    g = golden(fname)
    p, ne = int(g["p"]), int(g["ne"])
    b = mb.Basis(p)
    assert np.array_equal(b.x, g["gll_x"]) and np.array_equal(b.w, g["gll_w"])
    assert np.abs(b.ejxi - g["ejxi"]).max() <= 4e-16 * np.abs(g["ejxi"]).max()
    assert np.abs(b.ljxi - g["ljxi"]).max() <= 1e-15
    m = mb.Mesh(kind, p, ne, signed_det=signed)
    assert (m.N0, m.N1, m.N2) == (int(g["N0"]), int(g["N1"]), int(g["N2"]))
    assert rel_l2(m.det, g["det"]) < 1e-14 and np.abs(m.det - g["det"]).max() <= 1e-14 * np.abs(g["det"]).max()
    assert np.abs(m.J - g["J"]).max() <= 1e-14 * np.abs(g["J"]).max()


def test_elmats_match_the_oracle_tabulation():
    from oracle import mimsem_oracle as mo
    for p in (2, 3, 4):
        em = mo.elmats(p, p)
        b = mb.Basis(p)
        for k in "UVWPQ":
            assert np.abs(b.elmat(k) - em[k]).max() <= 1e-14 * max(1.0, np.abs(em[k]).max())


def test_gll_weights_sum_to_two():
    """The reference's own sanity check (eul/Basis.cpp:91-97)."""
    for n in range(1, 8):
        b = mb.Basis(n)
        assert abs(b.w.sum() - 2.0) < 1e-8
        assert np.allclose(b.x, -b.x[::-1])
