"""Diagnostic: time the pieces of one distributed M1 step (torchrun)."""
import os, sys
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import mimsem_b200 as mb
from mimsem_b200.parallel import DistributedEngine
from mimsem_b200.engine import SUBSET_INTERIOR, SUBSET_BOUNDARY
from helpers import synthetic_thickness
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
mesh = mb.Mesh("sphere", 4, 48); nk = 60
thick = synthetic_thickness(mesh.xyz, nk)
d = DistributedEngine(mesh, thick, rank, world, local)
e = d.engine
x = torch.rand((e.n1, nk), dtype=torch.float64, device="cuda"); y = torch.empty_like(x)
def timeit(f, n=20):
    for _ in range(3): f()
    torch.cuda.synchronize(); dist.barrier()
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): f()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3
res = {}
if os.environ.get("MIMSEM_DEBUG"):
    g_step = d.capture("M1", x, out=y, scale=1e8, tpow=1)[0]
    print("rank", rank, "MIMSEM_DEBUG", os.environ["MIMSEM_DEBUG"], "step_graph", round(timeit(g_step, 100), 1), flush=True)
    dist.barrier(); dist.destroy_process_group(); sys.exit(0)
res["interior"] = timeit(lambda: e.apply("M1", x, out=y, scale=1e8, tpow=1, flags=SUBSET_INTERIOR))
res["boundary"] = timeit(lambda: e.apply("M1", x, out=y, scale=1e8, tpow=1, flags=SUBSET_BOUNDARY))
res["all_local"] = timeit(lambda: e.apply("M1", x, out=y, scale=1e8, tpow=1))
res["exchange"] = timeit(lambda: d.exchange(x, 1))
res["step"] = timeit(lambda: d.apply("M1", x, out=y, scale=1e8, tpow=1))
import time
def graphed(f):
    f(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        f()
    return g.replay
dist.barrier()
g_local = graphed(lambda: e.apply("M1", x, out=y, scale=1e8, tpow=1))
res["all_local_graph"] = timeit(g_local)
g_step = d.capture("M1", x, out=y, scale=1e8, tpow=1)[0]
res["step_graph"] = timeit(g_step)
res["step_graph_200"] = timeit(g_step, 200)
torch.cuda.synchronize(); dist.barrier()
t0 = time.perf_counter()
for _ in range(200): g_step()
res["cpu_issue_per_replay"] = (time.perf_counter() - t0) / 200 * 1e6
torch.cuda.synchronize()
print("rank", rank, "owned", e.nel_owned, "total", e.nel_total, "interior", d.n_interior, "boundary", d.n_boundary, "n1", e.n1,
      "halo_bytes", d.halo_bytes(1, nk), {k: round(v, 1) for k, v in res.items()}, "us; halo err", d.halo_error(), flush=True)
dist.barrier(); dist.destroy_process_group()
