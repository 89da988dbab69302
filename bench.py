#!/usr/bin/env python3
"""bench.py -- GDOF/s of the FP64 horizontal operator apply (BASELINE.json metric) on N B200s.

A "step" is one pass of the hot path over one batch of synthetic input: one 1-form mass-matrix
application y = M1 x (the reference's Umat::assemble(lev, SCALE, true) + MatMult for EVERY level,
eul/Assembly.cpp:51-153) over all vertical levels in a single launch.  Workload at N = 1 is
BASELINE.json configs[4] ("eul/ baroclinic scaling p=4, 48x48 elems/face, 60 levels"), the only
BASELINE shape that is HBM-bound and the one the metric's roofline target
is stated on; it fits one GPU (0.42 GB per field pair).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--op M1] [--workload C5] [--lockstep] [--impl reference]

N > 1 (under torchrun, one rank per GPU): the one C5 mesh is split into contiguous element blocks (strong scaling); the
ghost refresh is fused into the M1 launch over NVLink peer memory and, by default, software-pipelined over the ring of
independent inputs (the launch of step i pushes the boundary rows of step i+1's input); --lockstep pushes and consumes
in the same launch.  --impl reference times the reference's own CPU path (oracle/_ref) with as many emulated MPI ranks
as the host has cores for.  --workload C5_half | C5_quarter | C5_eighth: one GPU, no exchange, 1/2 .. 1/8 of the
elements (the granularity ceiling of strong scaling, profiles/r01_session2_summary.md section 8).

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
    "C5": ("sphere", 4, 48, 60, "eul"),
    "C3": ("sphere", 3, 12, 30, "eul"),
    "C4": ("box", 3, 20, 40, "box"),
    # 1/2, 1/4 and 1/8 of C5's elements on ONE GPU (no exchange): the granularity ceiling of strong scaling
    "C5_half": ("sphere", 4, 34, 60, "eul"),
    "C5_quarter": ("sphere", 4, 24, 60, "eul"),
    "C5_eighth": ("sphere", 4, 17, 60, "eul"),
}
SCALE = 1.0e8
TPOW = {"M1": 1, "M1h": 2, "M2": 1, "M0": 1, "K": 2, "E21": 0, "E12": 0, "E10": 0, "E01": 0, "M2h": 2, "M0h": 2}
METRIC = "GDOF/s FP64 horizontal operator apply"


def algorithmic_bytes(op, nel, q2, N0, N1, N2, NQ, nk):
    """Compulsory HBM traffic of one launch: every input and output DOF once, the
    per-point inverse thickness once per unique quadrature point per level, coefficient fields once,
    the pre-scaled geometry once per (element, quadrature point)."""
    thick = 8 * NQ * nk
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
    raise ValueError(op)


def ncu_traffic(workload, op):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture (profiles/r01_traffic.json), or None."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))[workload][op]
        return t["dram_bytes_read"] + t["dram_bytes_write"]
    except Exception:
        return None


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
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for nm, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


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


def reference_mesh_dir(kind, p, ne, tmp_root, nprocs=None):
    """input/ directory for the reference's sources: the reference-generated one when it travelled with
    the repo, else written by the product's own writer (bit-identical maps, coordinates to 1e-15)."""
    if nprocs is None:
        nprocs = 6 if kind == "sphere" else 1
    d = os.path.join(ROOT, "oracle", "_ref", "meshes", "%s_p%d_ne%d_np%d" % (kind, p, ne, nprocs))
    if os.path.isdir(os.path.join(d, "input")):
        return d, nprocs, "reference-generated"
    import mimsem_b200 as mb
    d = os.path.join(tmp_root, "mesh_%s_p%d_ne%d_np%d" % (kind, p, ne, nprocs))
    os.makedirs(os.path.join(d, "input"), exist_ok=True)
    mb.write_input(kind, p, ne, nprocs, os.path.join(d, "input"))
    return d, nprocs, "written by mimsem_topo_write_input"


def cpu_reference_sample(workload, nlev_sample, budget_s, seed=0):
    """Time the reference's own CPU path (oracle/_ref: its unmodified sources behind the PETSc shim) on a
    bounded sample of the workload: Umat::assemble + MatMult for `nlev_sample` levels, one emulated MPI
    rank per cube face on its own host thread."""
    import tempfile
    from helpers import synthetic_thickness
    from oracle import refbind as rb
    import mimsem_b200 as mb
    kind, p, ne, nk, variant = WORKLOADS[workload]
    if not rb.available(variant):
        return None
    tmp = tempfile.mkdtemp(prefix="mimsem_bench_")
    ncpu = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    md, nprocs, how = reference_mesh_dir(kind, p, ne, tmp, reference_ranks(kind, ne, ncpu))
    cores = min(ncpu, nprocs)
    R = rb.Reference(variant, md, nprocs, nk=nk, nthreads=cores)
    mesh = mb.Mesh(kind, p, ne)
    thick = synthetic_thickness(mesh.xyz, nk, kind)
    for r in range(nprocs):
        R.set_thick(r, thick[:, R.loc(r, "locq" if variant == "eul" else "loc0")])
    rng = np.random.default_rng(seed)
    x = rng.uniform(-1, 1, R.N1)
    t_asm = t_mrg = t_mv = 0.0
    done = 0
    t0 = time.time()
    for lev in range(nlev_sample):
        a, m = R.assemble_only("Umat", lev=lev, scale=SCALE, flag=True)
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
                      "threads; mesh %s); assemble %.2fs, SpMV %.3fs counted; shim CSR merge %.2fs not counted" % (workload, done, nk, nprocs, cores, how, t_asm, t_mv, t_mrg),
            "matmult_only_gdofs": dofs / t_mv / 1e9,
            "matrix_free_twin_gdofs": (R.N1 / t_mf / 1e9) if t_mf else None}


def run_reference_arm(args):
    """--impl reference: the reference's CPU implementation of the path on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    kind, p, ne, nk, variant = WORKLOADS[args.workload]
    per_step_levels = 2
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
    N1 = 12 * (p * ne) ** 2
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "GDOF/s", "n_gpus": args.gpus, "steps": len(times),
            "warmup": args.warmup, "ms_per_step": 1e3 * (N1 * per_step_levels / 1e9) / v, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "%s eul/ p=%d %dx%d elems/face %d levels, operator Umat (M1) apply; each step = assemble+MatMult "
                                   "for %d levels" % (args.workload, p, ne, ne, nk, per_step_levels)},
            "cpu_baseline": {"kind": res["kind"], "cores": res["cores"], "sample": res["sample"], "value": v, "unit": "GDOF/s"},
            "e2e": {"value": v, "unit": "GDOF/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="mimsem_b200")
    ap.add_argument("--op", default="M1", choices=["M1", "M1h", "M2", "M0", "K", "E21", "E12"])
    ap.add_argument("--workload", default="C5", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--opt", action="append", default=[], metavar="NAME=VALUE",
                    help="engine tuning knob (mimsem_gpu_set_option), e.g. --opt m1_min_blocks=5; recorded in config")
    ap.add_argument("--lockstep", action="store_true", help="N > 1: push and consume the ghost rows in the same launch (no pipelining over the ring)")
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
    mesh = mb.Mesh(kind, p, ne)
    thick = synthetic_thickness(mesh.xyz, nk, kind)
    if world > 1:
        from mimsem_b200.parallel import DistributedEngine
        eng = DistributedEngine(mesh, thick, rank, world, local)
    else:
        eng = mb.Engine.from_mesh(mesh, local, thick=thick)
    for kv in args.opt:
        name, val = kv.split("=")
        (eng if world == 1 else eng.engine).set_option(name, int(val))
    op = args.op
    nin, nout, ncoef = eng.space_sizes(op)
    tpow = TPOW[op]

    # ring of distinct field sets: every step reads and writes buffers that were last touched
    # >= RING-1 steps ago; one field pair (0.42 GB on C5) already exceeds the 126 MB L2
    # (at N GPUs the per-rank slice shrinks, so the ring grows until it covers >= 4x the 126 MB L2)
    # computed from GLOBAL sizes: every rank must capture and replay the same number of steps (the ghost refresh is collective)
    n_glob = {"M1": mesh.N1, "M1h": mesh.N1, "K": mesh.N1, "E21": mesh.N1, "M2": mesh.N2, "E12": mesh.N2, "M0": mesh.N0}[op]
    RING = max(3, -(-4 * 126_000_000 // (16 * (n_glob // world) * nk)))
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    xs = [torch.rand((nin, nk), dtype=torch.float64, device=dev, generator=g) * 2 - 1 for _ in range(RING)]
    ys = [torch.empty((nout, nk), dtype=torch.float64, device=dev) for _ in range(RING)]
    cs = None
    if ncoef:
        cs = [torch.rand((ncoef, nk), dtype=torch.float64, device=dev, generator=g) + 0.5 for _ in range(RING)]

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    # kernels of this library per step (graph replays bypass the library's launch counter)
    l0 = eng.launch_count
    eng.apply(op, xs[0], coeff=None if cs is None else cs[0], out=ys[0], scale=SCALE, tpow=tpow)
    barrier()
    launches_per_step = eng.launch_count - l0

    # N > 1, M1: the ghost refresh is software-pipelined over the ring of independent inputs -- the launch of step i
    # pushes the boundary rows of step i+1's input and consumes what step i-1 pushed (one push + one apply per step,
    # as without pipelining; only the dependency distance changes).  --lockstep pushes and consumes in the same launch.
    pipelined = world > 1 and op == "M1" and not args.lockstep and getattr(eng, "p2p", None) is not None and eng.fused
    nxt = (lambda j: {"x_next": xs[(j + 1) % RING]}) if pipelined else (lambda j: {})
    if pipelined:
        eng.prologue_push(xs[0])
        barrier()

    # the step (ghost refresh + kernels) is captured once per ring slot into a CUDA graph and replayed:
    # at 4-8 GPUs a step is tens of microseconds, i.e. launch-bound without graphs
    replays = None
    if not args.no_graph and (world == 1 or getattr(eng, 'graph_safe', False)):
        try:
            # (each capture warms up with two real applies of its slot; after the last slot the pipeline holds the
            #  boundary rows of xs[0], the input of the first step)
            replays = [eng.capture(op, xs[j], coeff=None if cs is None else cs[j], out=ys[j], scale=SCALE, tpow=tpow, **nxt(j))[0]
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
            eng.apply(op, xs[j], coeff=None if cs is None else cs[j], out=ys[j], scale=SCALE, tpow=tpow, **nxt(j))

    barrier()
    for i in range(args.warmup):
        step(i)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = eng.launch_count
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    ev[0].record()
    for i in range(args.steps):
        step(args.warmup + i)
        ev[i + 1].record()
    barrier()
    total_ms = ev[0].elapsed_time(ev[-1])
    per_launch_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    launches = eng.launch_count - launches0
    if replays is not None:
        launches = launches_per_step * args.steps
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t[0])
    ms_per_step = total_ms / args.steps

    q2 = (p + 1) ** 2
    alg_bytes, out_dofs = algorithmic_bytes(op, mesh.nel, q2, mesh.N0, mesh.N1, mesh.N2, mesh.NQ, nk)
    value = out_dofs / (ms_per_step * 1e-3) / 1e9
    peak, peak_src = measured_peak()
    kern_ms = float(np.mean(per_launch_ms))
    achieved = (alg_bytes / world) / (kern_ms * 1e-3) / 1e9

    # end to end through the C ABI with HOST buffers (pinned), H2D + D2H inside the timed region
    e2e = None
    if not args.no_e2e:
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
        for _ in range(2):
            heng.apply_host(op, hx.numpy(), coeff=None if hc is None else hc.numpy(), scale=SCALE, tpow=tpow, out=hy.numpy())
        barrier()
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            heng.apply_host(op, hx.numpy(), coeff=None if hc is None else hc.numpy(), scale=SCALE, tpow=tpow, out=hy.numpy())
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / n_e2e
        h2d, d2h = 8 * nk * (nin + ncoef), 8 * nk * nout
        if world > 1:
            import torch.distributed as dist
            tt = torch.tensor([dt, -float(h2d), -float(d2h)], dtype=torch.float64, device=dev)
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
        line = {"metric": METRIC, "value": value, "unit": "GDOF/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if world > 1 else "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": "%s: %s p=%d, %dx%d elems/face, %d levels; operator %s over all levels in one launch "
                                       "(Nel=%d, out DOF-levels=%d)" % (args.workload, variant, p, ne, ne, nk, op, mesh.nel, out_dofs),
                           "cache": "ring of %d distinct field sets per GPU (%.0f MB each, %.0f MB in total per GPU vs 126 MB L2)" % (RING, 16e-6 * nin * nk, RING * 16e-6 * nin * nk),
                           "parallelism": "element-block x%d" % world, "cuda_graph": replays is not None, "options": args.opt,
                           "ghost_refresh": ("none (1 GPU)" if world == 1 else
                                             ("fused into the M1 launch over NVLink peer memory; push of step i+1's input overlapped with step i (ring of independent inputs)"
                                              if pipelined else "fused into the M1 launch over NVLink peer memory; push and consume in the same launch"))},
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": (ncu_traffic(args.workload, op) if world == 1 else None), "traffic_source": "ncu --set full capture, profiles/r01_traffic.json",
                             "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes / world,
                             "kernel_ms": kern_ms},
                "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks}
        print(json.dumps(line))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
