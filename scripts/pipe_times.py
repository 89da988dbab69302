"""Diagnostic (MIMSEM_DIAG build: make -C mimsem_b200/csrc DIAG=-DMIMSEM_DIAG BUILD=build_diag OUT=../libmimsem_gpu_diag.so;
run with MIMSEM_GPU_LIB=mimsem_b200/libmimsem_gpu_diag.so): per-tile phase stamps of the persistent M1 kernel."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import mimsem_b200 as mb
from helpers import synthetic_thickness
op = sys.argv[1] if len(sys.argv) > 1 else "M1"
mesh = mb.Mesh("sphere", 4, 48); nk = 60
eng = mb.Engine.from_mesh(mesh, 0, thick=synthetic_thickness(mesh.xyz, nk))
eng.set_option("m1_variant", 3)
for a in sys.argv[2:]:
    n, v = a.split("="); eng.set_option(n, int(v))
x = torch.rand((mesh.N1, nk), dtype=torch.float64, device="cuda"); y = torch.empty_like(x)
h = torch.rand((mesh.N2, nk), dtype=torch.float64, device="cuda") + 0.5
kw = dict(scale=1e8, tpow=1) if op == "M1" else dict(coeff=h, scale=1e8, tpow=2)
NCTA = 148
buf = torch.zeros((NCTA, 128, 8), dtype=torch.int64, device="cuda")
for _ in range(3): eng.apply(op, x, out=y, **kw)
torch.cuda.synchronize()
eng.set_option("diag_times", buf.data_ptr())
eng.apply(op, x, out=y, **kw)
torch.cuda.synchronize()
t = buf.cpu().numpy().astype(np.float64)
t0 = t[t > 0].min()
print(op, sys.argv[2:], "kernel span us", (t.max() - t0) / 1e3)
J = slice(8, 80)   # steady-state jobs
st0, st1, st2, m0, m1, m2 = (t[:, J, i] for i in range(6))
print("stager: start->copies issued %.0f ns (incl. wait for an empty buffer), ->far operands written %.0f ns" % ((st1 - st0).mean(), (st2 - st0).mean()))
print("main:   wait for full %.0f ns, contraction %.0f ns" % ((m1 - m0).mean(), (m2 - m1).mean()))
per = np.diff(t[:, 8:81, 5], axis=1)
print("job completion period per CTA %.0f ns (p10 %.0f p90 %.0f)" % (per.mean(), np.percentile(per, 10), np.percentile(per, 90)))
print("copies issued -> contraction sees the buffer full: %.0f ns ; far operands written -> full seen: %.0f ns" % ((m1 - st1).mean(), (m1 - st2).mean()))
print("first job: stager start -> full %.0f ns" % (t[:, 0, 4] - t[:, 0, 0]).mean())
