"""Multi-GPU parity check, run under torchrun (one rank per GPU, NCCL):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mp_check.py
Every rank applies the partitioned operators to its element block (ghost rows start as zero, so the NCCL
ghost refresh is exercised); the owned rows are gathered and compared, on rank 0, with the reference's
golden vectors and with the single-GPU engine -- which must agree BITWISE (owner-computes gather)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import mimsem_b200 as mb  # noqa: E402
from mimsem_b200.parallel import DistributedEngine  # noqa: E402
from helpers import TOL, golden, rel_l2, synthetic_fields, synthetic_thickness, to_cols, to_np  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    failures = []
    cases = [("golden", "sphere", 3, 4, 3), ("synthetic", "sphere", 4, 6, 60), ("synthetic", "sphere", 3, 6, 30), ("synthetic", "box", 3, 6, 40)]
    for what, kind, p, ne, nk in cases:
        mesh = mb.Mesh(kind, p, ne)
        if what == "golden":
            g = golden("ops_eul_sphere_p3_ne4.npz")
            thick, f = g["thick"], {k: g[k] for k in ("x1", "x2", "x0", "h2", "u1", "q0")}
        else:
            thick = synthetic_thickness(mesh.xyz, nk, kind)
            f = synthetic_fields(np.random.default_rng(5), nk, mesh.N0, mesh.N1, mesh.N2, float(mesh.det.mean()))
            f["q0"] = np.random.default_rng(6).uniform(-1, 1, (nk, mesh.N0)) * 1e-4
        deng = DistributedEngine(mesh, thick, rank, world, local)
        single = mb.Engine.from_mesh(mesh, local, thick=thick) if rank == 0 else None
        # all fourteen operators of the path (+ UtQW): 1- and 2-form operators, then the 0-form family (node-partitioned:
        # M0, M0h, E01, M0h_up sum over the elements around a node; E10, R, R_up read node values of their elements)
        f["uu"] = f["u1"] * 1.0e-3
        ops = [("M1", "x1", None, 1), ("M1h", "x1", "h2", 2), ("M2", "x2", None, 1), ("M2h", "x2", "h2", 2), ("K", "x1", "u1", 2),
               ("E21", "x1", None, 0), ("E12", "x2", None, 0), ("UtQW", "x2", "u1", 0),
               ("M0", "x0", None, 1), ("M0h", "x0", "h2", 2), ("E10", "x0", None, 0), ("E01", "x1", None, 0), ("R", "x1", "q0", 2),
               ("R_up", "x1", "q0", 0), ("M0h_up", "x0", "h2", 0)]
        for op, xk, ck, tpow in ops:
            sin, sout, sc = deng.engine.SPACES[op]
            x = deng.scatter_from_global(f[xk], sin)
            n_own_in = deng.part.n_owned(sin)
            perm_in = torch.from_numpy(deng.engine.permutation(sin).astype(np.int64)).cuda()
            x[perm_in[n_own_in:]] = 0.0                       # ghosts must come from the exchange
            c = None
            if ck is not None:
                c = deng.scatter_from_global(f[ck], sc)
                n_own_c = deng.part.n_owned(sc)
                perm_c = torch.from_numpy(deng.engine.permutation(sc).astype(np.int64)).cuda()
                c[perm_c[n_own_c:]] = 0.0
            kw = dict(scale=1e8, tpow=tpow) if op[0] != "E" else {}
            if op == "UtQW":
                kw = dict(scale=1e8)
            skw = dict(kw)
            if op in ("R_up", "M0h_up"):
                u = deng.scatter_from_global(f["uu"], 1)
                perm_u = torch.from_numpy(deng.engine.permutation(1).astype(np.int64)).cuda()
                u[perm_u[deng.part.n1_owned:]] = 0.0
                kw.update(u1=u, tau=150.0)
            y = deng.apply(op, x, coeff=c, **kw)
            N = {0: mesh.N0, 1: mesh.N1, 2: mesh.N2}[sout]
            yg = np.zeros((f[xk].shape[0], N))
            deng.owned_to_global(y, sout, yg)
            t = torch.from_numpy(yg).cuda()
            dist.all_reduce(t)
            if rank == 0:
                yg = t.cpu().numpy()
                cs = None if ck is None else to_cols(single, f[ck], sc)
                if op in ("R_up", "M0h_up"):
                    skw.update(u1=to_cols(single, f["uu"], 1), tau=150.0)
                ys = to_np(single, single.apply(op, to_cols(single, f[xk], sin), coeff=cs, **skw), sout)
                if not np.array_equal(yg, ys):
                    failures.append((what, kind, p, ne, op, "differs from single GPU", rel_l2(yg, ys)))
                if what == "golden":
                    key = {"M1": "y_Umat_vs1", "M1h": "y_Uhmat_cv1", "M2": "y_Wmat_vs1", "M2h": "y_Whmat_vs1", "K": "y_WtQUmat"}.get(op)
                    if key and rel_l2(yg, g[key]) >= TOL:
                        failures.append((what, op, "golden", rel_l2(yg, g[key])))
        # software-pipelined ghost refresh (fused M1): prologue, then each apply pushes the NEXT input
        if deng.p2p is not None and deng.fused and nk % 2 == 0:
            xa = deng.scatter_from_global(f["x1"], 1)
            xb = deng.scatter_from_global(f["x1"][::-1].copy() * 0.5, 1)
            perm1 = torch.from_numpy(deng.engine.permutation(1).astype(np.int64)).cuda()
            xa[perm1[deng.part.n1_owned:]] = 0.0
            xb[perm1[deng.part.n1_owned:]] = 0.0
            torch.cuda.synchronize(); dist.barrier()
            deng.prologue_push(xa)
            torch.cuda.synchronize(); dist.barrier()
            outs = [deng.apply("M1", xa, x_next=xb, scale=1e8, tpow=1), deng.apply("M1", xb, x_next=xa, scale=1e8, tpow=1),
                    deng.apply("M1", xa, pipeline_last=True, scale=1e8, tpow=1)]
            for yl, xg in zip(outs, (f["x1"], f["x1"][::-1].copy() * 0.5, f["x1"])):
                yg = np.zeros((nk, mesh.N1))
                deng.owned_to_global(yl, 1, yg)
                t = torch.from_numpy(yg).cuda()
                dist.all_reduce(t)
                if rank == 0:
                    ys = to_np(single, single.apply("M1", to_cols(single, xg, 1), scale=1e8, tpow=1), 1)
                    if not np.array_equal(t.cpu().numpy(), ys):
                        failures.append((what, kind, p, ne, "pipelined M1 differs from single GPU", rel_l2(t.cpu().numpy(), ys)))
            # back in the ordinary mode: push and consume in one call
            yl = deng.apply("M1", xb, scale=1e8, tpow=1)
            yg = np.zeros((nk, mesh.N1))
            deng.owned_to_global(yl, 1, yg)
            t = torch.from_numpy(yg).cuda()
            dist.all_reduce(t)
            if rank == 0:
                ys = to_np(single, single.apply("M1", to_cols(single, f["x1"][::-1].copy() * 0.5, 1), scale=1e8, tpow=1), 1)
                if not np.array_equal(t.cpu().numpy(), ys):
                    failures.append((what, kind, p, ne, "M1 after a pipelined sequence differs from single GPU"))
        # back-to-back M1 (fused launch) and K (push / pull kernels) on the SAME 1-form space with no host synchronisation
        # in between: the two ghost-refresh mechanisms must not overwrite each other's inbox copies
        if deng.p2p is not None and nk % 2 == 0:
            x1 = deng.scatter_from_global(f["x1"], 1)
            u1 = deng.scatter_from_global(f["u1"], 1)
            perm1 = torch.from_numpy(deng.engine.permutation(1).astype(np.int64)).cuda()
            x1[perm1[deng.part.n1_owned:]] = 0.0
            u1[perm1[deng.part.n1_owned:]] = 0.0
            for it in range(12):
                ya = deng.apply("M1", x1, scale=1e8, tpow=1)
                yb = deng.apply("K", x1, coeff=u1, scale=1e8, tpow=2)
            for yl, op, sp_ in ((ya, "M1", 1), (yb, "K", 2)):
                N = {1: mesh.N1, 2: mesh.N2}[sp_]
                yg = np.zeros((nk, N))
                deng.owned_to_global(yl, sp_, yg)
                tt = torch.from_numpy(yg).cuda()
                dist.all_reduce(tt)
                if rank == 0:
                    cs = to_cols(single, f["u1"], 1) if op == "K" else None
                    ys = to_np(single, single.apply(op, to_cols(single, f["x1"], 1), coeff=cs, scale=1e8, tpow=1 if op == "M1" else 2), sp_)
                    if not np.array_equal(tt.cpu().numpy(), ys):
                        failures.append((what, kind, p, ne, "back-to-back M1 / K: %s differs from single GPU" % op))
        # programmatic dependent launch: six fused steps over three INDEPENDENT fields in one CUDA graph; a step starts while
        # the previous one drains (interior tiles), its push CTAs and boundary tiles wait for the previous launch
        if deng.p2p is not None and deng.fused and getattr(deng, "graph_safe", False) and nk % 2 == 0:
            xsl, refs = [], []
            perm1 = torch.from_numpy(deng.engine.permutation(1).astype(np.int64)).cuda()
            for i in range(3):
                xi = deng.scatter_from_global(f["x1"] * (1.0 + 0.25 * i), 1)
                xi[perm1[deng.part.n1_owned:]] = 0.0
                xsl.append(xi)
                refs.append(deng.apply("M1", xi, scale=1e8, tpow=1).clone())
            outs = [torch.empty_like(r) for r in refs]
            deng.engine.set_option("pdl_independent", 1)
            replay = deng.capture_burst("M1", xsl, None, outs, 6, scale=1e8, tpow=1)
            for o in outs:
                o.zero_()
            for _ in range(3):
                replay()
            torch.cuda.synchronize()
            deng.engine.set_option("pdl_independent", 0)
            n_own = deng.part.n1_owned
            for i in range(3):
                rows = perm1[:n_own]
                if not torch.equal(outs[i][rows], refs[i][rows]):
                    failures.append((what, kind, p, ne, "fused M1 under programmatic dependent launch differs (slot %d, rank %d)" % (i, rank)))
            # the same burst with the ghost rows of step i+1 pushed during step i (pipelined), replays chained
            deng.engine.set_option("pdl_independent", 1)
            replay = deng.capture_burst("M1", xsl, None, outs, 6, pipelined=True, scale=1e8, tpow=1)
            for o in outs:
                o.zero_()
            torch.cuda.synchronize()
            dist.barrier()
            deng.prologue_push(xsl[0])
            torch.cuda.synchronize()
            dist.barrier()
            for _ in range(3):
                replay()
            last = deng.apply("M1", xsl[0], scale=1e8, tpow=1, pipeline_last=True)
            torch.cuda.synchronize()
            deng.engine.set_option("pdl_independent", 0)
            for i in range(3):
                rows = perm1[:n_own]
                if not torch.equal(outs[i][rows], refs[i][rows]):
                    failures.append((what, kind, p, ne, "pipelined burst differs (slot %d, rank %d)" % (i, rank)))
            if not torch.equal(last[perm1[:n_own]], refs[0][perm1[:n_own]]):
                failures.append((what, kind, p, ne, "consume-only apply after a pipelined burst differs (rank %d)" % rank))
        # partitioned mass-matrix solve: x -> b = M1 x (partitioned) -> CG over peer memory recovers x; every rank
        # stops at the same iteration with the same residual
        if deng.p2p is not None and deng.fused and nk % 2 == 0:
            xt = deng.scatter_from_global(f["x1"], 1)
            bb = deng.apply("M1", xt, scale=1e8, tpow=1)
            xs, its, rr = deng.solve("M1", bb, scale=1e8, tpow=1, rtol=1e-13, maxit=300)
            stat = torch.tensor([float(its), rr], dtype=torch.float64, device="cuda")
            lo, hi = stat.clone(), stat.clone()
            dist.all_reduce(lo, op=dist.ReduceOp.MIN)
            dist.all_reduce(hi, op=dist.ReduceOp.MAX)
            if not torch.equal(lo, hi):
                failures.append((what, kind, p, ne, "partitioned solve: ranks disagree on iterations / residual", lo.tolist(), hi.tolist()))
            xg = np.zeros((nk, mesh.N1))
            deng.owned_to_global(xs, 1, xg)
            tt = torch.from_numpy(xg).cuda()
            dist.all_reduce(tt)
            if rank == 0:
                err = rel_l2(tt.cpu().numpy(), f["x1"])
                if not (its < 300 and err < 1e-10):
                    failures.append((what, kind, p, ne, "partitioned solve", its, rr, err))
        if rank == 0:
            print("case", what, kind, p, ne, nk, "world", world, "halo bytes/rank (1-form)", deng.halo_bytes(1, f["x1"].shape[0]), flush=True)
        if deng.halo_error():
            failures.append(("halo timeout", rank))
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("MP_CHECK", "FAIL %r" % failures if failures else "OK")
    sys.exit(1 if failures else 0)


if __name__ == "__main__":
    main()
