"""two-front elimination with REAL factors (TDEM shifts): residual of op.solve on a face / cell operator"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import casingsimulations_b200 as cs
from casingsimulations_b200 import _lib
nth = int(sys.argv[1]); npadz = int(sys.argv[2]); csz = float(sys.argv[3])
mp = cs.model.CasingInHalfspace(src_a=np.r_[0.0, 0.0, 0.0], src_b=np.r_[1000.0, 0.0, 0.0], mur_casing=100.0)
mg = cs.CasingMeshGenerator(modelParameters=mp, npadx=10, npadz=npadz, csz=csz, hy=np.ones(nth) * 2 * np.pi / nth)
m = mg.mesh
ctx = _lib.Context(0)
ctx.set_mesh(m.hx, m.hy, m.hz, m.x0[2])
ctx.paint(mp.paint_params())
print("mesh", tuple(int(v) for v in m.vnC), "CS_TWIST", os.environ.get("CS_TWIST"))
for kind, name, shift in ((_lib.KIND_DC_GROUNDED, "dc", 0.0), (_lib.KIND_FACE_J, "face_j", 1e6), (_lib.KIND_FACE_J, "face_j", 1e5), (_lib.KIND_FACE_J, "face_j", 1e2)):
    op = ctx.op(kind)
    op.factor([shift])
    g = torch.Generator(device="cpu").manual_seed(3)
    x = torch.randn(op.n, dtype=torch.float64, generator=g).to(ctx.device)
    b = op.apply(x, shift)
    try:
        y = op.solve(b)
        r = b - op.apply(y, shift)
        print(name, "n", op.n, "residual", float(torch.linalg.norm(r) / torch.linalg.norm(b)), "err", float(torch.linalg.norm(y - x) / torch.linalg.norm(x)), "nan", bool(torch.isnan(y).any()))
    except Exception as ex:
        print(name, "FAILED", str(ex)[:200])
ctx.close()
