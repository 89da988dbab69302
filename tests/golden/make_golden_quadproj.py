#!/usr/bin/env python3
"""Golden vectors of the quadrature-point projections the reference initialises its fields with -- WtQmat, UtQmat, PtQmat
(eul/Assembly.cpp:758-830, WtQmat / UtQmat further down; callers eul/Euler_2.cpp:432, 493, 535) -- FROM THE REFERENCE ITSELF:
the unmodified classes behind the PETSc/MPI shim (oracle/_ref/libref_eul.so) on the reference-generated meshes, six emulated
ranks, merged matrix times a seeded quadrature-point vector.  Separate from make_golden.py so that the other fixtures keep
their bytes.

    make -C oracle && python tests/golden/make_golden_quadproj.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import refbind as rb  # noqa: E402


def main():
    for p, ne in ((3, 4), (4, 2)):
        R = rb.Reference("eul", rb.mesh_dir("sphere", p, ne, 6), 6, nk=1)
        W, U, P = R.assemble("WtQmat"), R.assemble("UtQmat"), R.assemble("PtQmat")
        nq = W.shape[1]
        assert U.shape[1] == 2 * nq and P.shape[1] == nq
        rng = np.random.default_rng(11 + p)
        xq = rng.uniform(-1, 1, nq)                 # a scalar field at the quadrature points
        uq = rng.uniform(-1, 1, 2 * nq)             # a vector field: components interleaved per point (eul/Euler_2.cpp:420-431)
        out = dict(p=p, ne=ne, nq=nq, N0=P.shape[0], N1=U.shape[0], N2=W.shape[0], xq=xq, uq=uq, y_WtQmat=W @ xq, y_UtQmat=U @ uq,
                   y_PtQmat=P @ xq)
        R.close()
        np.savez_compressed(os.path.join(HERE, "quadproj_eul_sphere_p%d_ne%d.npz" % (p, ne)), **out)
        print("quadproj p%d ne%d: nq %d, |WtQ x| %.6e |UtQ u| %.6e |PtQ x| %.6e" % (p, ne, nq, np.linalg.norm(out["y_WtQmat"]),
                                                                                 np.linalg.norm(out["y_UtQmat"]), np.linalg.norm(out["y_PtQmat"])))


if __name__ == "__main__":
    main()
