"""Diagnostic (torchrun): per-step times of the fused ghost-refresh + M1 launch and, for one launch, the phase stamps of
the push CTAs and the boundary tiles (globaltimer)."""
import os, sys
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import mimsem_b200 as mb
from mimsem_b200.parallel import DistributedEngine
from helpers import synthetic_thickness
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
mesh = mb.Mesh("sphere", 4, 48); nk = 60
d = DistributedEngine(mesh, synthetic_thickness(mesh.xyz, nk), rank, world, local)
e = d.engine
x = torch.rand((e.n1, nk), dtype=torch.float64, device="cuda"); y = torch.empty_like(x)
NS = int(os.environ.get("NSTEPS", "40"))
for _ in range(3): d.apply("M1", x, out=y, scale=1e8, tpow=1)
torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(NS + 1)]
ev[0].record()
for i in range(NS):
    d.apply("M1", x, out=y, scale=1e8, tpow=1)
    ev[i + 1].record()
torch.cuda.synchronize()
t = np.array([ev[i].elapsed_time(ev[i + 1]) * 1e3 for i in range(NS)])
msg = "rank %d steps us: median %.1f min %.1f max %.1f n>1ms %d err %s" % (rank, np.median(t), t.min(), t.max(), int((t > 1000).sum()), d.halo_error())
# one instrumented launch
push_ctas = max(1, min(148, d._inbox[1][2] // 16))
buf = torch.zeros((e.nel_owned + push_ctas + 8, 6), dtype=torch.int64, device="cuda")
os.environ["MIMSEM_DBG_TIMES"] = str(buf.data_ptr())
dist.barrier(); torch.cuda.synchronize()
d.apply("M1", x, out=y, scale=1e8, tpow=1)
torch.cuda.synchronize()
del os.environ["MIMSEM_DBG_TIMES"]
b = buf.cpu().numpy().astype(np.float64)
tiles, push = b[:e.nel_owned], b[e.nel_owned:e.nel_owned + push_ctas]
t0 = min(tiles[:, 0].min(), push[:, 0].min())
ni = d.n_interior
msg += "\n   span %.1f | push CTAs: start %.1f..%.1f, ack-wait %.1f, copy %.1f, fence %.1f, end by %.1f" % (
    (tiles[:, 3].max() - t0) / 1e3, (push[:, 0].min() - t0) / 1e3, (push[:, 0].max() - t0) / 1e3, (push[:, 1] - push[:, 0]).max() / 1e3,
    (push[:, 2] - push[:, 1]).max() / 1e3, (push[:, 3] - push[:, 2]).max() / 1e3, (push[:, 4].max() - t0) / 1e3)
bt = tiles[ni:]
msg += "\n   boundary tiles: first start %.1f, issue(wait) mean %.2f max %.2f, lifetime mean %.2f ; interior lifetime mean %.2f, last interior end %.1f" % (
    (bt[:, 0].min() - t0) / 1e3, (bt[:, 1] - bt[:, 0]).mean() / 1e3, (bt[:, 1] - bt[:, 0]).max() / 1e3, (bt[:, 3] - bt[:, 0]).mean() / 1e3,
    (tiles[:ni, 3] - tiles[:ni, 0]).mean() / 1e3, (tiles[:ni, 3].max() - t0) / 1e3)
print(msg, flush=True)
dist.barrier(); dist.destroy_process_group()
