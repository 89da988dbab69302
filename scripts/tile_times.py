"""Diagnostic: per-phase latency of the M1 TMA tile kernel (globaltimer stamps per CTA)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import mimsem_b200 as mb
from helpers import synthetic_thickness
mesh = mb.Mesh("sphere", 4, 48); nk = 60
eng = mb.Engine.from_mesh(mesh, 0, thick=synthetic_thickness(mesh.xyz, nk))
x = torch.rand((mesh.N1, nk), dtype=torch.float64, device="cuda"); y = torch.empty_like(x)
buf = torch.zeros((mesh.nel, 6), dtype=torch.int64, device="cuda")
for _ in range(3): eng.apply("M1", x, out=y, scale=1e8, tpow=1)
torch.cuda.synchronize()
os.environ["MIMSEM_DBG_TIMES"] = str(buf.data_ptr())
eng.apply("M1", x, out=y, scale=1e8, tpow=1)
torch.cuda.synchronize()
t = buf.cpu().numpy().astype(np.float64)
t0 = t[:, 0].min()
d = np.diff(t[:, :5], axis=1)
print("kernel span us", (t[:, 4].max() - t0) / 1e3)
print("phase means ns: setup+issue %.0f  wait %.0f  compute %.0f  store %.0f ; lifetime %.0f" % (*d.mean(0), (t[:, 4] - t[:, 0]).mean()))
print("phase p90   ns:", np.percentile(d, 90, axis=0), "lifetime p90", np.percentile(t[:, 4] - t[:, 0], 90))
# concurrency: average number of CTAs alive
ev = np.concatenate([np.stack([t[:, 0], np.ones(len(t))], 1), np.stack([t[:, 4], -np.ones(len(t))], 1)])
ev = ev[np.argsort(ev[:, 0])]
alive = np.cumsum(ev[:, 1]); dt = np.diff(ev[:, 0])
print("mean CTAs alive", (alive[:-1] * dt).sum() / dt.sum(), "of", 148 * 4)
