#!/usr/bin/env python3
"""End-to-end entry point (mimsem_gpu_apply_host: pinned host buffers in the reference's per-level layout, H2D + D2H inside
the call) under several pipeline chunk sizes, in ONE process: wall time per call and a bitwise comparison with the
device-resident apply for every setting.

    python scripts/tune_e2e.py --op M1 --op M1h --workload C5 --chunks 12,8,6,4
Prints one JSON line per (operator, chunk size)."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--op", action="append", default=[])
    ap.add_argument("--workload", default="C5")
    ap.add_argument("--chunks", default="12,8,6,4")
    ap.add_argument("--reps", type=int, default=8)
    args = ap.parse_args()
    import numpy as np
    import torch
    import bench
    import mimsem_b200 as mb
    from helpers import synthetic_thickness
    kind, p, ne, nk, variant = bench.WORKLOADS[args.workload]
    mesh = mb.Mesh(kind, p, ne, signed_det=(variant == "src"))
    thick = synthetic_thickness(mesh.xyz, nk, kind) if nk > 1 or variant != "src" else None
    eng = mb.Engine.from_mesh(mesh, 0, thick=thick)
    for op in (args.op or ["M1"]):
        nin, nout, ncoef = eng.space_sizes(op)
        tpow = bench.TPOW[op] if thick is not None else 0
        hx = torch.empty((nk, nin), dtype=torch.float64).pin_memory()
        hx.uniform_(-1, 1)
        hy = torch.empty((nk, nout), dtype=torch.float64).pin_memory()
        hc = None
        if ncoef:
            hc = torch.empty((nk, ncoef), dtype=torch.float64).pin_memory()
            hc.uniform_(0.5, 1.5)
        kw = {} if op.startswith("E") else dict(scale=bench.SCALE, tpow=tpow)
        # device-resident result of the same input (all levels, one launch)
        xd = eng.to_columns(hx.cuda(), eng.SPACES[op][0])
        cd = None if hc is None else eng.to_columns(hc.cuda(), eng.SPACES[op][2])
        ref = eng.to_levels(eng.apply(op, xd, coeff=cd, **kw), eng.SPACES[op][1]).cpu().numpy()
        for ch in [int(v) for v in args.chunks.split(",")]:
            eng.set_option("host_chunk", ch)
            call = lambda: eng.apply_host(op, hx.numpy(), coeff=None if hc is None else hc.numpy(), out=hy.numpy(), **kw)
            for _ in range(2):
                call()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(args.reps):
                call()
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / args.reps
            same = bool(np.array_equal(hy.numpy(), ref))
            gb = 8.0 * nk * (nin + ncoef + nout) / 1e9
            print(json.dumps({"op": op, "workload": args.workload, "host_chunk": ch, "ms": dt * 1e3, "gdofs": nout * nk / dt / 1e9,
                              "pcie_gbs_both_ways": gb / dt, "bitwise_vs_device_resident": same}), flush=True)


if __name__ == "__main__":
    main()
