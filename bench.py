#!/usr/bin/env python3
"""bench.py -- GDOF/s of the FP64 horizontal operator apply (BASELINE.json metric) on N B200s.

A "step" is one pass of the hot path over one batch of synthetic input: one 1-form mass-matrix
application y = M1 x (the reference's Umat::assemble(lev, SCALE, true) + MatMult for EVERY level,
eul/Assembly.cpp:51-153) over all vertical levels in a single launch.  Workload at N = 1 is
BASELINE.json configs[4] ("eul/ baroclinic scaling p=4, 48x48 elems/face, 60 levels"), the only
BASELINE shape that is HBM-bound and the one the metric's roofline target
is stated on; it fits one GPU (0.42 GB per field pair).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--op M1] [--workload C5] [--chain] [--pipelined] [--impl reference]

N > 1 (under torchrun, one rank per GPU): the one C5 mesh is split into contiguous element blocks (STRONG scaling); the
ghost refresh is fused into the M1 launch over NVLink peer memory.  The timed mode is LOCKSTEP -- every launch pushes
the boundary rows of its own input and consumes them, which is what a dependent sequence of applies (a solver
iteration) does; the software-pipelined mode (the launch of step i pushes the rows of step i+1's independent input) is
timed afterwards and reported as the extra field "pipelined".  Before anything is timed, the N-GPU output of step 0 is
compared BITWISE with a single-GPU apply of the gathered input on rank 0; no number is printed if they differ.
--impl reference times the reference's own CPU path (oracle/_ref) with as many emulated MPI ranks as the host has
cores for; it imports nothing of the product.  --workload C1..C4: the other BASELINE shapes (L2-resident, launch-bound);
--chain times the diagnose chain (4 x Uhmat apply + E21) of one time step under ONE CUDA graph instead of a single
operator.  --workload C5_half | C5_quarter | C5_eighth: one GPU, no exchange, 1/2 .. 1/8 of the elements (the
granularity ceiling of strong scaling).

Prints ONE JSON line (rank 0).  See the task contract for the keys.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

WORKLOADS = {
    # name: (kind, p, ne, nk, variant)
    "C1": ("sphere", 3, 4, 1, "src"),
    "C2": ("sphere", 3, 16, 1, "src"),
    "C3": ("sphere", 3, 12, 30, "eul"),
    "C4": ("box", 3, 20, 40, "box"),
    "C5": ("sphere", 4, 48, 60, "eul"),
    # 1/2, 1/4 and 1/8 of C5's elements on ONE GPU (no exchange): the granularity ceiling of strong scaling
    "C5_half": ("sphere", 4, 34, 60, "eul"),
    "C5_quarter": ("sphere", 4, 24, 60, "eul"),
    "C5_eighth": ("sphere", 4, 17, 60, "eul"),
}
SCALE = 1.0e8
TPOW = {"M1": 1, "M1h": 2, "M2": 1, "M0": 1, "K": 2, "E21": 0, "E12": 0, "E10": 0, "E01": 0, "M2h": 2, "M0h": 2, "R": 2,
        "R_up": 0, "M0h_up": 0}
OPS = ["M1", "M1h", "M2", "M0", "K", "E21", "E12", "E10", "E01", "R_up", "M0h_up"]
METRIC = "GDOF/s FP64 horizontal operator apply"
UP_TAU = 0.5 * 300.0   # src/SWEqn_Picard.cpp:30 UP_TAU times a 300 s step


def algorithmic_bytes(op, nel, q2, N0, N1, N2, NQ, nk, thickness=True):
    """Compulsory HBM traffic of one launch: every input and output DOF once, the
    per-point inverse thickness once per unique quadrature point per level (3-D only), coefficient fields once,
    the pre-scaled geometry once per (element, quadrature point).  Returns (bytes, output DOF-levels)."""
    thick = 8 * NQ * nk if thickness else 0
    if op == "M1":
        return 8 * N1 * nk * 2 + thick + 24 * nel * q2, N1 * nk
    if op == "M1h":
        return 8 * N1 * nk * 2 + 8 * N2 * nk + thick + 24 * nel * q2, N1 * nk
    if op == "M2":
        return 8 * N2 * nk * 2 + thick + 8 * nel * q2, N2 * nk
    if op == "M0":
        return 8 * N0 * nk * 2 + thick + 8 * N0, N0 * nk
    if op == "K":
        return 8 * N1 * nk * 2 + 8 * N2 * nk + thick + 24 * nel * q2, N2 * nk
    if op == "E21":
        return 8 * N1 * nk + 8 * N2 * nk, N2 * nk
    if op == "E12":
        return 8 * N2 * nk + 8 * N1 * nk, N1 * nk
    if op == "E10":
        return 8 * N0 * nk + 8 * N1 * nk, N1 * nk
    if op == "E01":
        return 8 * N1 * nk + 8 * N0 * nk, N0 * nk
    if op == "R_up":    # x, y, u1 (1-forms), q0 (0-form), J (4) + det + signed weight per (element, point)
        return 8 * N1 * nk * 3 + 8 * N0 * nk + thick + 48 * nel * q2, N1 * nk
    if op == "M0h_up":  # x, y (0-forms), h2 (2-form), u1 (1-form), J (4) + det per (element, point)
        return 8 * N0 * nk * 2 + 8 * N2 * nk + 8 * N1 * nk + thick + 40 * nel * q2, N0 * nk
    raise ValueError(op)


def ncu_traffic(workload, op):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture (profiles/r0*_traffic.json), or None."""
    for name in ("r02_traffic.json", "r01_traffic.json"):
        try:
            t = json.load(open(os.path.join(ROOT, "profiles", name)))[workload][op]
            return t["dram_bytes_read"] + t["dram_bytes_write"], "ncu --set full capture, profiles/" + name
        except Exception:
            continue
    return None, None


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                pw.append(float(r[2]))
                for nm, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the reference's own sources (oracle/_ref), nothing of the product

def reference_ranks(kind, ne, cores):
    """The reference runs one MPI rank per patch: 6 n^2 ranks on the sphere (n^2 in the box) with n dividing the elements
    per side (README.md:32 of the reference).  Use as many as the host has cores for (at most 96), at least one patch per face."""
    best = 6 if kind == "sphere" else 1
    for n in range(1, ne + 1):
        if ne % n:
            continue
        r = (6 if kind == "sphere" else 1) * n * n
        if r <= min(cores, 96):
            best = max(best, r)
    return best


def reference_mesh_dir(kind, p, ne, nprocs):
    """input/ directory written by the reference's OWN generator scripts (oracle/gen_meshes.py, travels in oracle/_ref);
    falls back to fewer ranks if that rank count was not generated.  Returns (dir, nprocs) or (None, None)."""
    base = 6 if kind == "sphere" else 1
    cands = sorted({nprocs, base} | {base * n * n for n in range(1, 12) if base * n * n <= nprocs}, reverse=True)
    for np_ in cands:
        d = os.path.join(ROOT, "oracle", "_ref", "meshes", "%s_p%d_ne%d_np%d" % (kind, p, ne, np_))
        if os.path.isdir(os.path.join(d, "input")):
            return d, np_
    return None, None


def cpu_reference_sample(workload, nlev_sample, budget_s, seed=0):
    """Time the reference's own CPU path (oracle/_ref: its unmodified sources behind the PETSc shim) on a
    bounded sample of the workload: Umat::assemble + MatMult for `nlev_sample` levels, one emulated MPI
    rank per patch, each on its own host thread.  Meshes, coordinates and geometry are the reference's own."""
    from helpers import synthetic_thickness
    from oracle import refbind as rb
    kind, p, ne, nk, variant = WORKLOADS[workload]
    if not rb.available(variant):
        return None
    ncpu = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    md, nprocs = reference_mesh_dir(kind, p, ne, reference_ranks(kind, ne, ncpu))
    if md is None:
        return None
    cores = min(ncpu, nprocs)
    R = rb.Reference(variant, md, nprocs, nk=nk, nthreads=cores)
    if variant != "src":
        for r in range(nprocs):
            R.set_thick(r, synthetic_thickness(R.coords(r), nk, kind))   # the reference's own Geom::x of that rank
    rng = np.random.default_rng(seed)
    x = rng.uniform(-1, 1, R.N1)
    t_asm = t_mrg = t_mv = 0.0
    done = 0
    t0 = time.time()
    for lev in range(min(nlev_sample, nk)):
        a, m = R.assemble_only("Umat", lev=lev, scale=SCALE if variant != "src" else 1.0, flag=True)
        _, s = R.spmv(x, nthreads=cores)
        t_asm += a
        t_mrg += m
        t_mv += s
        done += 1
        if time.time() - t0 > budget_s:
            break
    # the matrix-free twin the reference also ships (Uvec::assemble, eul/Assembly.cpp:2124-2196)
    t_mf = None
    if variant == "eul":
        _, t_mf = R.uvec_apply(x, lev=0, scale=SCALE)
    R.close()
    dofs = R.N1 * done
    # the shim's triplet merge stands in for PETSc's own insertion work inside MatSetValues/MatAssembly;
    # it is NOT charged to the reference (reported in `sample` only)
    total = t_asm + t_mv
    return {"value": dofs / total / 1e9, "unit": "GDOF/s", "cores": cores, "kind": "reference",
            "sample": "%s: Umat::assemble + MatMult for %d of %d levels (reference sources via PETSc shim, %d emulated MPI ranks on %d "
                      "threads; mesh, coordinates and Jacobians from the reference's own scripts and Geom); assemble %.2fs, SpMV %.3fs counted; "
                      "shim CSR merge %.2fs not counted" % (workload, done, nk, nprocs, cores, t_asm, t_mv, t_mrg),
            "matmult_only_gdofs": dofs / t_mv / 1e9,
            "matrix_free_twin_gdofs": (R.N1 / t_mf / 1e9) if t_mf else None}


def run_reference_arm(args):
    """--impl reference: the reference's CPU implementation of the path on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    kind, p, ne, nk, variant = WORKLOADS[args.workload]
    per_step_levels = min(2, nk)
    times = []
    res = None
    for it in range(args.warmup + args.steps):
        t0 = time.time()
        res = cpu_reference_sample(args.workload, per_step_levels, budget_s=120, seed=it)
        if res is None:
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref not built (reference sources need /root/reference)"}))
            return
        if it >= args.warmup:
            times.append(res["value"])
        if time.time() - t0 > 60 and it >= args.warmup:
            break
    v = float(np.mean(times))
    N1 = (12 if kind == "sphere" else 2) * (p * ne) ** 2
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "GDOF/s", "n_gpus": args.gpus, "steps": len(times),
            "warmup": args.warmup, "ms_per_step": 1e3 * (N1 * per_step_levels / 1e9) / v, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "%s %s/ p=%d %dx%d elems/face %d levels, operator Umat (M1) apply; each step = assemble+MatMult "
                                   "for %d levels" % (args.workload, variant, p, ne, ne, nk, per_step_levels)},
            "cpu_baseline": {"kind": res["kind"], "cores": res["cores"], "sample": res["sample"], "value": v, "unit": "GDOF/s"},
            "e2e": {"value": v, "unit": "GDOF/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------------------------

def parity_vs_one_gpu(eng, mesh, thick, op, x, coeff, y, kw, rank, world, local):
    """N-GPU correctness at the size the numbers are quoted on: gather the owned rows of the input, the coefficient and
    the N-GPU output; rank 0 applies the operator to the gathered input on ONE GPU and compares bit for bit."""
    import torch
    import torch.distributed as dist
    import mimsem_b200 as mb
    sin, sout, sc = eng.engine.SPACES[op]
    nk = x.shape[1]
    sizes = {0: mesh.N0, 1: mesh.N1, 2: mesh.N2}

    def gather(field, space):
        g = np.zeros((nk, sizes[space]))
        eng.owned_to_global(field, space, g)
        t = torch.from_numpy(g).to("cuda:%d" % local)
        dist.all_reduce(t)          # every DOF has exactly one owner: a sum of one value and zeros
        return t

    xg = gather(x, sin)
    cg = gather(coeff, sc) if coeff is not None else None
    yg = gather(y, sout)
    ok = True
    if rank == 0:
        single = mb.Engine.from_mesh(mesh, local, thick=thick)
        y1 = single.to_levels(single.apply(op, single.to_columns(xg, sin), coeff=None if cg is None else single.to_columns(cg, sc), **kw), sout)
        ok = bool(torch.equal(y1, yg))
        single.close()
        del single, y1
    flag = torch.tensor([1 if ok else 0], device="cuda:%d" % local)
    dist.broadcast(flag, 0)
    del xg, cg, yg
    torch.cuda.empty_cache()
    return bool(flag.item())


def run_dependent_chain(args, eng, mesh, nk, scale, world, rank, dev, variant, p, ne):
    """Every kernel of a step depends on the one before it (and the solve on its own dot products): no overlap between
    launches is possible, each pays its ramp, its ghost refresh and its tail.  Stream order, no CUDA graph (the solve reads
    its convergence flags back every four iterations)."""
    import torch
    n1, n2 = eng.space_sizes("M1h")[0], eng.space_sizes("M1h")[2]
    g = torch.Generator(device=dev).manual_seed(99 + rank)
    u = torch.rand((n1, nk), dtype=torch.float64, device=dev, generator=g) * 2 - 1
    rho = torch.rand((n2, nk), dtype=torch.float64, device=dev, generator=g) + 0.5
    F = torch.empty_like(u)
    d = torch.empty((eng.space_sizes("E21")[1], nk), dtype=torch.float64, device=dev)

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()
    its = [0]

    def step():
        eng.apply("M1h", u, coeff=rho, out=F, scale=scale, tpow=2)
        us, it, _ = eng.solve("M1", F, scale=scale, tpow=1, rtol=1e-10, maxit=200)
        its[0] = it
        eng.apply("E21", us, out=d)
    for _ in range(max(2, min(args.warmup, 3))):
        step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    nstep = max(3, min(args.steps, 10))
    e0.record()
    for _ in range(nstep):
        step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / nstep
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0])
        if eng.halo_error():
            raise SystemExit("rank %d: a ghost refresh timed out" % rank)
    if rank == 0:
        dofs = mesh.N1 * nk
        print(json.dumps({"metric": "dependent chain M1h -> solve_M1 -> E21 (GDOF/s of the 1-form field per chain)", "value": dofs / (ms * 1e-3) / 1e9,
                          "unit": "GDOF/s", "n_gpus": world, "steps": nstep, "ms_per_step": ms, "cg_iterations": its[0],
                          "config": {"workload": "%s: %s p=%d, %dx%d elems/face, %d levels" % (args.workload, variant, p, ne, ne, nk),
                                     "launch": "stream order, no CUDA graph; solve: batched Jacobi-PCG, rtol 1e-10, dot products over peer memory"}}))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="mimsem_b200")
    ap.add_argument("--op", default="M1", choices=OPS)
    ap.add_argument("--workload", default="C5", choices=sorted(WORKLOADS))
    ap.add_argument("--chain", action="store_true", help="time the diagnose chain (4 x M1h + E21) under one CUDA graph instead of --op")
    ap.add_argument("--dependent-chain", action="store_true",
                    help="time the DEPENDENT chain F = M1h(rho) u -> u' = M1^-1 F (PCG, rtol 1e-10) -> div = E21 u' (the mass flux of "
                         "diagnose_fluxes, eul/HorizSolve.cpp:298-310) instead of --op; prints its own JSON line")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-sustained", action="store_true")
    ap.add_argument("--no-pdl", action="store_true",
                    help="1 GPU: one CUDA graph per step and ordinary stream order between steps (default: graphs of several consecutive "
                         "steps, launched with programmatic dependent launch -- the steps of the ring are independent applies)")
    ap.add_argument("--no-parity", action="store_true", help="N > 1: skip the bitwise comparison with a single-GPU apply")
    ap.add_argument("--opt", action="append", default=[], metavar="NAME=VALUE",
                    help="engine tuning knob (mimsem_gpu_set_option), e.g. --opt m1_min_blocks=5; recorded in config")
    ap.add_argument("--pipelined", action="store_true",
                    help="N > 1: time ONLY the software-pipelined ghost refresh (independent inputs) as the headline")
    ap.add_argument("--lockstep", action="store_true", help="(default since round 2; kept for old command lines)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch
    import mimsem_b200 as mb
    from helpers import synthetic_thickness

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = "cuda:%d" % local

    kind, p, ne, nk, variant = WORKLOADS[args.workload]
    flat = variant == "src"              # 2-D shallow water: no thickness, unit scale, signed det J
    mesh = mb.Mesh(kind, p, ne, signed_det=flat)
    thick = None if flat else synthetic_thickness(mesh.xyz, nk, kind)
    scale = 1.0 if flat else SCALE
    if world > 1:
        from mimsem_b200.parallel import DistributedEngine
        eng = DistributedEngine(mesh, thick, rank, world, local)
    else:
        eng = mb.Engine.from_mesh(mesh, local, thick=thick)
    # 1 GPU: consecutive steps are independent applies (ring of distinct fields), so a launch may start while the CTAs of
    # the previous one retire (programmatic dependent launch); plain M1 then runs best as the persistent ring kernel
    burst = not args.no_graph and not args.no_pdl and not args.dependent_chain and (world == 1 or (args.op == "M1" and not args.chain))
    auto_opts = []
    if burst:
        auto_opts.append("pdl_independent=1")   # (plain M1 on >= 4000 elements then runs as the persistent ring kernel: the engine's default rule)
    for kv in auto_opts + args.opt:
        name, val = kv.split("=")
        (eng if world == 1 else eng.engine).set_option(name, int(val))
    if args.dependent_chain:
        run_dependent_chain(args, eng, mesh, nk, scale, world, rank, dev, variant, p, ne)
        return
    op = "M1h" if args.chain else args.op
    if world > 1 and (args.chain or op not in eng.SUPPORTED):
        raise SystemExit("operator %s is not available on more than one GPU" % ("chain" if args.chain else op))
    nin, nout, ncoef = eng.space_sizes(op)
    tpow = 0 if flat else TPOW[op]
    need_u = op in ("R_up", "M0h_up")
    n1_local = eng.space_sizes("M1")[0]

    # ring of distinct field sets: every step reads and writes buffers that were last touched
    # >= RING-1 steps ago; one field pair (0.42 GB on C5) already exceeds the 126 MB L2
    # (at N GPUs the per-rank slice shrinks, so the ring grows until it covers >= 4x the 126 MB L2)
    # computed from GLOBAL sizes: every rank must capture and replay the same number of steps (the ghost refresh is collective)
    n_glob = max(mesh.N1 if op not in ("M2", "M0") else (mesh.N2 if op == "M2" else mesh.N0), 1)
    RING = min(64, max(3, -(-4 * 126_000_000 // (16 * max(n_glob // world, 1) * nk))))
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    xs = [torch.rand((nin, nk), dtype=torch.float64, device=dev, generator=g) * 2 - 1 for _ in range(RING)]
    ys = [torch.empty((nout, nk), dtype=torch.float64, device=dev) for _ in range(RING)]
    cs = None
    if ncoef:
        cs = [torch.rand((ncoef, nk), dtype=torch.float64, device=dev, generator=g) + 0.5 for _ in range(RING)]
    us = None
    if need_u:
        umag = float(np.abs(mesh.det).mean()) * 1e-3
        us = [(torch.rand((n1_local, nk), dtype=torch.float64, device=dev, generator=g) * 2 - 1) * umag for _ in range(RING)]
    chain_extra = None
    if args.chain:
        # the other three (coefficient, field) pairs of one diagnose step and the divergence of the first flux
        n2 = eng.space_sizes("E21")[1]
        chain_extra = [([torch.rand((nin, nk), dtype=torch.float64, device=dev, generator=g) for _ in range(3)],
                        [torch.rand((ncoef, nk), dtype=torch.float64, device=dev, generator=g) + 0.5 for _ in range(3)],
                        [torch.empty((nout, nk), dtype=torch.float64, device=dev) for _ in range(3)],
                        torch.empty((n2, nk), dtype=torch.float64, device=dev)) for _ in range(RING)]

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def kw_of(j):
        if op.startswith("E"):
            return {}
        kw = dict(scale=scale, tpow=tpow)
        if need_u:
            kw.update(u1=us[j], tau=UP_TAU)
        return kw

    def apply_slot(j, **extra):
        eng.apply(op, xs[j], coeff=None if cs is None else cs[j], out=ys[j], **kw_of(j), **extra)
        if args.chain:
            fx, fc, fy, dv = chain_extra[j]
            for i in range(3):
                eng.apply(op, fx[i], coeff=fc[i], out=fy[i], **kw_of(j))
            eng.apply("E21", ys[j], out=dv)

    # kernels of this library per step (graph replays bypass the library's launch counter)
    l0 = eng.launch_count
    apply_slot(0)
    barrier()
    launches_per_step = eng.launch_count - l0

    # N > 1: the N-GPU result of this very configuration must equal the single-GPU result bit for bit
    parity = None
    if world > 1 and not args.no_parity:
        ok = parity_vs_one_gpu(eng, mesh, thick, op, xs[0], None if cs is None else cs[0], ys[0], kw_of(0), rank, world, local)
        parity = {"bitwise_vs_1gpu": ok, "what": "owned rows of step 0's output on %d GPUs vs a one-GPU apply of the gathered input (rank 0), %s %s"
                                                 % (world, args.workload, op)}
        if not ok:
            raise SystemExit("rank %d: the %d-GPU output differs from the single-GPU output -- refusing to print a number" % (rank, world))

    fused = world > 1 and op == "M1" and getattr(eng, "p2p", None) is not None and eng.fused
    if world > 1 and burst and not (fused and getattr(eng, "graph_safe", False)):
        burst = False
        eng.engine.set_option("pdl_independent", 0)
        auto_opts = []

    def measure_burst(pipelined=False):
        """1 GPU: W warm-up + exactly K timed steps, replayed from CUDA graphs that hold several consecutive steps each
        (slot = step % RING), so that programmatic dependent launch can overlap a launch with the tail of the previous one."""
        B = RING * max(1, 18 // RING)

        def capture(nsteps):
            if world > 1:
                class _G:   # same interface as a CUDAGraph
                    replay = staticmethod(eng.capture_burst(op, xs, cs, ys, nsteps, pipelined=pipelined, **kw_of(0)))
                return _G
            st = torch.cuda.Stream()
            with torch.cuda.stream(st):
                for i in range(min(nsteps, RING)):
                    apply_slot(i % RING)
                torch.cuda.synchronize()
                gr = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gr, stream=st):
                    for i in range(nsteps):
                        apply_slot(i % RING)
            return gr
        g_full = capture(B)
        n_full, n_rem = args.steps // B, args.steps % B
        g_rem = capture(n_rem) if n_rem else None
        if pipelined:
            barrier()
            eng.prologue_push(xs[0])
            barrier()
        for _ in range(-(-args.warmup // B)):
            g_full.replay()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n_full):
            g_full.replay()
        if g_rem is not None:
            g_rem.replay()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1) / args.steps
        if world > 1:
            import torch.distributed as dist
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t[0])
        sustained_ms = None
        if not args.no_sustained:
            n_rep = max(1, min(100000 if world == 1 else 4000, int(1.0e3 / max(ms, 1e-3))) // B)
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record()
            for _ in range(n_rep):
                g_full.replay()
            s1.record()
            barrier()
            sus = s0.elapsed_time(s1) / (n_rep * B)
            if world > 1:
                import torch.distributed as dist
                t = torch.tensor([sus], dtype=torch.float64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                sus = float(t[0])
            sustained_ms = (sus, n_rep * B)
        if pipelined:
            # end of the pipelined sequence: consume what the last step pushed, push nothing
            eng.apply(op, xs[0], coeff=None if cs is None else cs[0], out=ys[0], pipeline_last=True, **kw_of(0))
            barrier()
        return ms, ms, launches_per_step * args.steps, sustained_ms, True

    def measure(pipelined):
        """W warm-up + K timed steps (CUDA-graph replays of the captured step, one graph per ring slot)."""
        nxt = (lambda j: {"x_next": xs[(j + 1) % RING]}) if pipelined else (lambda j: {})
        if pipelined:
            eng.prologue_push(xs[0])
            barrier()
        replays = None
        if burst:
            return measure_burst(pipelined)
        if not args.no_graph and (world == 1 or getattr(eng, 'graph_safe', False)):
            try:
                # (each capture warms up with real applies of its slot; after the last slot a pipelined sequence
                #  holds the boundary rows of xs[0], the input of the first step)
                if world == 1:
                    replays = []
                    for j in range(RING):
                        apply_slot(j)
                        torch.cuda.synchronize()
                        gr = torch.cuda.CUDAGraph()
                        with torch.cuda.graph(gr):
                            apply_slot(j)
                        replays.append(gr.replay)
                else:
                    replays = [eng.capture(op, xs[j], coeff=None if cs is None else cs[j], out=ys[j], **kw_of(j), **nxt(j))[0]
                               for j in range(RING)]
            except Exception as exc:  # fall back to eager launches
                if rank == 0:
                    print("graph capture failed (%r); timing eager launches" % (exc,), file=sys.stderr)
                replays = None

        def step(i):
            j = i % RING
            if replays is not None:
                replays[j]()
            else:
                apply_slot(j, **nxt(j))

        barrier()
        for i in range(args.warmup):
            step(i)
        barrier()
        l_before = eng.launch_count
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
        barrier()
        ev[0].record()
        for i in range(args.steps):
            step(args.warmup + i)
            ev[i + 1].record()
        barrier()
        total_ms = ev[0].elapsed_time(ev[-1])
        per = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
        launches = launches_per_step * args.steps if replays is not None else eng.launch_count - l_before
        if world > 1:
            import torch.distributed as dist
            t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            total_ms = float(t[0])
        # sustained leg: the same step replayed back to back for about 1 s (clocks settle, power rises).  The count is
        # derived from the REDUCED time (every rank must replay the same number of collective steps) and is a multiple
        # of the ring so that a pipelined sequence ends where it began
        sustained_ms = None
        if not args.no_sustained:
            n_sus = int(max(args.steps, min(100000 if world == 1 else 4000, 1.0e3 / max(total_ms / args.steps, 1e-3))))
            n_sus = -(-n_sus // RING) * RING
            first = args.warmup + args.steps
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            e0.record()
            for i in range(n_sus):
                step(first + i)
            e1.record()
            barrier()
            sus = e0.elapsed_time(e1) / n_sus
            if world > 1:
                import torch.distributed as dist
                t = torch.tensor([sus], dtype=torch.float64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                sus = float(t[0])
            sustained_ms = (sus, n_sus)
        return total_ms / args.steps, float(np.mean(per)), launches, sustained_ms, replays is not None

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    headline_pipelined = bool(args.pipelined and fused)
    ms_per_step, kern_ms, launches, sustained, graphed = measure(headline_pipelined)
    clocks = sampler.stop() if rank == 0 else None
    second = None
    strict = None
    if fused and burst and not headline_pipelined:
        barrier()
        second = measure(True)[0]
    if fused and burst:
        # the same launches in ordinary stream order, one graph per step: what a chain of DEPENDENT applies gets
        barrier()
        eng.engine.set_option("pdl_independent", 0)
        burst = False
        strict = measure(False)[0]
        burst = True
    elif fused and not headline_pipelined:
        barrier()
        second = measure(True)[0]


    q2 = (p + 1) ** 2
    if args.chain:
        a1, d1 = algorithmic_bytes("M1h", mesh.nel, q2, mesh.N0, mesh.N1, mesh.N2, mesh.NQ, nk, thickness=not flat)
        a2, d2 = algorithmic_bytes("E21", mesh.nel, q2, mesh.N0, mesh.N1, mesh.N2, mesh.NQ, nk, thickness=not flat)
        alg_bytes, out_dofs = 4 * a1 + a2, 4 * d1 + d2
    else:
        alg_bytes, out_dofs = algorithmic_bytes(op, mesh.nel, q2, mesh.N0, mesh.N1, mesh.N2, mesh.NQ, nk, thickness=not flat)
    value = out_dofs / (ms_per_step * 1e-3) / 1e9
    peak, peak_src = measured_peak()
    achieved = (alg_bytes / world) / (kern_ms * 1e-3) / 1e9

    # end to end through the C ABI with HOST buffers (pinned), H2D + D2H inside the timed region
    e2e = None
    if not args.no_e2e and not args.chain and not need_u:
        # N > 1: every rank passes its ghosted local vector (the reference's VecCreateSeq(topo->n1) convention), so the
        # host call needs no exchange; the aggregate is all ranks' owned output DOFs over the slowest rank's time
        heng = eng if world == 1 else eng.engine
        hx = torch.empty((nk, nin), dtype=torch.float64).pin_memory()
        hx.uniform_(-1, 1)
        hy = torch.empty((nk, nout), dtype=torch.float64).pin_memory()
        hc = None
        if ncoef:
            hc = torch.empty((nk, ncoef), dtype=torch.float64).pin_memory()
            hc.uniform_(0.5, 1.5)
        n_e2e = max(3, min(args.steps, 10))
        hkw = {} if op.startswith("E") else dict(scale=scale, tpow=tpow)
        for _ in range(2):
            heng.apply_host(op, hx.numpy(), coeff=None if hc is None else hc.numpy(), out=hy.numpy(), **hkw)
        barrier()
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            heng.apply_host(op, hx.numpy(), coeff=None if hc is None else hc.numpy(), out=hy.numpy(), **hkw)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / n_e2e
        h2d, d2h = 8 * nk * (nin + ncoef), 8 * nk * nout
        if world > 1:
            import torch.distributed as dist
            tt = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt = float(tt[0])
            tb = torch.tensor([float(h2d), float(d2h)], dtype=torch.float64, device=dev)
            dist.all_reduce(tb)
            h2d, d2h = int(tb[0]), int(tb[1])
        e2e = {"value": out_dofs / dt / 1e9, "unit": "GDOF/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "ms_per_step": dt * 1e3,
               "api": "mimsem_gpu_apply_host (per-level host layout in, per-level host layout out)" +
                      ("" if world == 1 else "; one call per rank on its ghosted local vectors, bytes summed over ranks")}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            cpu = cpu_reference_sample(args.workload, 4, budget_s=25)
        except Exception as exc:  # the checker must never take the measurement down
            cpu = {"error": repr(exc)}

    if world > 1 and eng.halo_error():
        raise SystemExit("rank %d: a peer-to-peer ghost refresh timed out (ranks out of step?) -- the measurement is void" % rank)
    if rank == 0:
        what = "diagnose chain: 4 x M1h (Uhmat) + E21 under one CUDA graph" if args.chain else "operator %s over all levels in one launch" % op
        traffic, traffic_src = ncu_traffic(args.workload, "chain" if args.chain else op) if world == 1 else (None, None)
        roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes / world, "kernel_ms": kern_ms}
        if sustained:
            roof["sustained_frac"] = (alg_bytes / world) / (sustained[0] * 1e-3) / 1e9 / peak
            roof["sustained_ms"] = sustained[0]
            roof["sustained_steps"] = sustained[1]
        mode = "none (1 GPU)"
        if world > 1:
            if fused:
                mode = "fused into the M1 launch over NVLink peer memory; " + (
                    "push of step i+1's input overlapped with step i (ring of independent inputs)" if headline_pipelined
                    else "push and consume in the same launch")
            else:
                mode = "push / pull kernels over NVLink peer memory on a side stream"
        line = {"metric": METRIC, "value": value, "unit": "GDOF/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": "%s: %s p=%d, %dx%d elems/face, %d levels; %s (Nel=%d, out DOF-levels=%d)"
                                       % (args.workload, variant, p, ne, ne, nk, what, mesh.nel, out_dofs),
                           "cache": "ring of %d distinct field sets per GPU (%.1f MB each, %.0f MB in total per GPU vs 126 MB L2)" % (RING, 16e-6 * nin * nk, RING * 16e-6 * nin * nk),
                           "parallelism": "element-block x%d" % world, "cuda_graph": graphed, "options": auto_opts + args.opt,
                           "launch": ("CUDA graphs of up to %d consecutive steps; the steps are independent applies (ring), launched with "
                                      "programmatic dependent launch: a step starts while the CTAs of the previous one retire" % (RING * max(1, 18 // RING))
                                      if burst else "one CUDA graph replay (or eager launch) per step, stream order between steps"),
                           "ghost_refresh": mode},
                "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks}
        if parity is not None:
            line["parity_check"] = parity
        if strict is not None:
            line["dependent_applies"] = {"value": out_dofs / (strict * 1e-3) / 1e9, "unit": "GDOF/s", "ms_per_step": strict,
                                         "note": "the same fused launches in ordinary stream order (one CUDA graph per step, no overlap between "
                                                 "consecutive launches): the rate of a chain of dependent applies"}
        if second is not None:
            line["pipelined"] = {"value": out_dofs / (second * 1e-3) / 1e9, "unit": "GDOF/s", "ms_per_step": second,
                                 "note": "ghost rows of step i+1's INDEPENDENT input pushed during step i; not the headline"}
        print(json.dumps(line))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
