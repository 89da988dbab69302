#!/usr/bin/env python3
"""Time one operator of one workload under several engine option settings in ONE process (the GPU box is paid for by the
minute): ring of field sets larger than L2, CUDA events on the launching stream, 5 warm-up + 20 timed launches per setting.

    python scripts/tune_ops.py --op M1 --workload C5 --sweep m1_min_blocks=4,5,6 --sweep prefetch_ahead=0,296,740
Prints one JSON line per setting (ms per launch, fraction of the measured HBM peak on the algorithmic bytes)."""
import argparse
import itertools
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--op", action="append", default=[])
    ap.add_argument("--workload", default="C5")
    ap.add_argument("--sweep", action="append", default=[], metavar="NAME=v1,v2,...")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--burst", type=int, default=0, help="capture this many consecutive steps in ONE CUDA graph and time replays of it")
    args = ap.parse_args()
    import torch
    import bench
    import mimsem_b200 as mb
    from helpers import synthetic_thickness
    kind, p, ne, nk, variant = bench.WORKLOADS[args.workload]
    mesh = mb.Mesh(kind, p, ne, signed_det=(variant == "src"))
    thick = synthetic_thickness(mesh.xyz, nk, kind) if nk > 1 or variant != "src" else None
    eng = mb.Engine.from_mesh(mesh, 0, thick=thick)
    peak, _ = bench.measured_peak()
    names = [s.split("=")[0] for s in args.sweep]
    values = [[int(v) for v in s.split("=")[1].split(",")] for s in args.sweep]
    dev = "cuda:0"
    for op in (args.op or ["M1"]):
        nin, nout, ncoef = eng.space_sizes(op)
        tpow = bench.TPOW[op] if thick is not None else 0
        ring = max(3, -(-4 * 126_000_000 // (16 * max(nin, nout) * nk)))
        g = torch.Generator(device=dev).manual_seed(1)
        xs = [torch.rand((nin, nk), dtype=torch.float64, device=dev, generator=g) * 2 - 1 for _ in range(ring)]
        ys = [torch.empty((nout, nk), dtype=torch.float64, device=dev) for _ in range(ring)]
        cs = [torch.rand((ncoef, nk), dtype=torch.float64, device=dev, generator=g) + 0.5 for _ in range(ring)] if ncoef else None
        alg, dofs = bench.algorithmic_bytes(op, mesh.nel, (p + 1) ** 2, mesh.N0, mesh.N1, mesh.N2, mesh.NQ, nk)
        for combo in itertools.product(*values) if values else [()]:
            for n, v in zip(names, combo):
                eng.set_option(n, v)
            def step(i):
                j = i % ring
                eng.apply(op, xs[j], coeff=None if cs is None else cs[j], out=ys[j], scale=bench.SCALE, tpow=tpow)
            for i in range(5):
                step(i)
            torch.cuda.synchronize()
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
            ev[0].record()
            for i in range(args.steps):
                step(5 + i)
                ev[i + 1].record()
            torch.cuda.synchronize()
            ms = ev[0].elapsed_time(ev[-1]) / args.steps
            best = min(ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps))
            if args.burst:
                s = torch.cuda.Stream()
                with torch.cuda.stream(s):
                    for i in range(args.burst):
                        step(i)
                    torch.cuda.synchronize()
                    gr = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(gr, stream=s):
                        for i in range(args.burst):
                            step(i)
                for _ in range(2):
                    gr.replay()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                reps = max(1, args.steps // args.burst * 3)
                e0.record()
                for _ in range(reps):
                    gr.replay()
                e1.record()
                torch.cuda.synchronize()
                ms = best = e0.elapsed_time(e1) / (reps * args.burst)
            print(json.dumps({"workload": args.workload, "op": op, "options": dict(zip(names, combo)), "ms": ms, "ms_best": best,
                              "gdofs": dofs / ms / 1e6, "frac": alg / (ms * 1e-3) / 1e9 / peak, "alg_bytes": alg}), flush=True)
        del xs, ys, cs
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
