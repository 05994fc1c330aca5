import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import casingsimulations_b200 as cs
from casingsimulations_b200 import _lib
nth = int(sys.argv[1]); csz = float(sys.argv[2])
mp = cs.model.CasingInHalfspace(src_a=np.r_[0.0, 0.0, 0.0], src_b=np.r_[1000.0, 0.0, 0.0], mur_casing=100.0)
mg = cs.CasingMeshGenerator(modelParameters=mp, npadx=10, npadz=25, csz=csz, hy=np.ones(nth) * 2 * np.pi / nth)
m = mg.mesh
print("mesh", tuple(int(v) for v in m.vnC), flush=True)
ctx = _lib.Context(0)
ctx.set_mesh(m.hx, m.hy, m.hz, m.x0[2])
ctx.paint(mp.paint_params())
op = ctx.op(_lib.KIND_FACE_J)
rhs = torch.randn(op.n, dtype=torch.float64, device=ctx.device)
for dt in (1e-6, 1e-5, 1e-4, 1e-3, 1e-2):
    op.factor([1.0 / dt])
    try:
        x = op.solve(rhs, rtol=1e-10, maxit=60)
        y = op.apply(x, shifts=1.0 / dt)
        print("dt", dt, "iters", op.info["iters"], "relres", op.info["relres"], "bwd", op.info["backward_err"], "raw res", float(torch.linalg.norm(y - rhs) / torch.linalg.norm(rhs)), "|x|", float(torch.linalg.norm(x)), flush=True)
    except Exception as e:
        print("dt", dt, "FAILED", str(e)[:200], op.info, flush=True)
ctx.close()
