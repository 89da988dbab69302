#!/usr/bin/env python3
"""ORACLE-ONLY (test infrastructure): run the reference's own offline mesh generators.

Runs /root/reference/scr/Setup.py (cubed sphere; scr/Setup.py:11-78) or scr/Setup_Box.py
(doubly periodic box; scr/Setup_Box.py:11-50) UNMODIFIED, from a scratch copy of scr/ (the
scripts write to ``../<proj>/`` relative to the cwd and the reference mount is read-only), and
moves the resulting ``input/*.txt`` under ``oracle/_ref/meshes/<name>/input/``.

``oracle/_ref/`` is git-ignored but travels to the GPU box.  Nothing under /root/reference is
copied into the repository; the scratch copy lives in a TemporaryDirectory.

    python oracle/gen_meshes.py                  # the standard set used by tests / bench
    python oracle/gen_meshes.py sphere 3 4 6     # p ne nprocs   -> sphere_p3_ne4_np6
    python oracle/gen_meshes.py box 3 4 1
"""
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("MIMSEM_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref", "meshes")

# (kind, p, ne, nprocs); quadrature order == p in every BASELINE config
STANDARD = [
    ("sphere", 3, 4, 6),    # C1 (src Williamson2) and the small eul parity mesh
    ("sphere", 3, 4, 24),   # multi-patch-per-face numbering (Topo closed-form check)
    ("sphere", 2, 2, 6),
    ("sphere", 4, 2, 6),
    ("sphere", 4, 4, 24),
    ("sphere", 3, 12, 6),   # C3 (eul UMJS14)
    ("sphere", 3, 16, 6),   # C2 (src Galewsky)
    ("box", 3, 4, 1),
    ("box", 3, 4, 4),
    ("box", 3, 20, 1),      # C4 (box bubble)
]
# C5, with the rank counts the CPU-baseline arm of bench.py picks for 6..23, 24..53, 54..95 and >= 96 host threads
BIG = [("sphere", 4, 48, 6), ("sphere", 4, 48, 24), ("sphere", 4, 48, 54), ("sphere", 4, 48, 96)]


def mesh_name(kind, p, ne, nprocs):
    return "%s_p%d_ne%d_np%d" % (kind, p, ne, nprocs)


def generate(kind, p, ne, nprocs, force=False):
    name = mesh_name(kind, p, ne, nprocs)
    dst = os.path.join(OUT, name)
    if os.path.isdir(os.path.join(dst, "input")) and not force:
        return dst
    if not os.path.isdir(os.path.join(REF, "scr")):
        raise RuntimeError("reference not mounted at %s; cannot generate %s" % (REF, name))
    with tempfile.TemporaryDirectory() as tmp:
        scr = os.path.join(tmp, "scr")
        shutil.copytree(os.path.join(REF, "scr"), scr)
        if kind == "sphere":
            cmd = [sys.executable, "Setup.py", str(p), str(ne), str(nprocs), str(p), "out"]
            proj = os.path.join(tmp, "out")
        else:
            cmd = [sys.executable, "Setup_Box.py", str(p), str(ne), str(nprocs)]
            proj = os.path.join(tmp, "box")
        subprocess.run(cmd, cwd=scr, check=True, stdout=subprocess.DEVNULL)
        os.makedirs(dst, exist_ok=True)
        if os.path.isdir(os.path.join(dst, "input")):
            shutil.rmtree(os.path.join(dst, "input"))
        shutil.move(os.path.join(proj, "input"), os.path.join(dst, "input"))
    return dst


def main(argv):
    if len(argv) >= 5:
        print(generate(argv[1], int(argv[2]), int(argv[3]), int(argv[4]), force=True))
        return
    todo = list(STANDARD)
    if len(argv) > 1 and argv[1] == "--big":
        todo += BIG
    for cfg in todo:
        print(generate(*cfg))


if __name__ == "__main__":
    main(sys.argv)
