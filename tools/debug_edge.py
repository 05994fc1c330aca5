import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import CASING, paint_params, small_casing_mesh
from test_gpu_parity import oracle_matrix, rnd
from oracle.cyl_oracle import OracleCylMesh, paint_casing_in_halfspace
from casingsimulations_b200 import _lib
ctx = _lib.Context(0)
nth = 4
hx, hy, hz, z0 = small_casing_mesh(nth)
om = OracleCylMesh([hx, hy, hz], [0, 0, z0])
ctx.set_mesh(hx, hy, hz, z0)
sig, mu = paint_casing_in_halfspace(om, **CASING)
ctx.paint(paint_params(CASING))
for kind, code in (("edge_h", 2), ("edge_e", 3)):
    op = ctx.op(code)
    for s in (0.0, 3.0):
        A = oracle_matrix(om, sig, mu, kind, s)
        xx = np.stack([rnd(op.n, False, 10), rnd(op.n, False, 20)])
        yy = op.apply(ctx.to_device(xx), shifts=s).cpu().numpy()
        print("vec0 rel", np.linalg.norm(yy[0] - A @ xx[0]) / np.linalg.norm(A @ xx[0]))
        x, y = xx[1], yy[1]
        ref = A @ x
        nx, nz = om.nx, om.nz
        blocks = {"Ex": (0, om.nEx), "Ey": (om.nEx, om.nEx + om.nEy), "Ez": (om.nEx + om.nEy, om.nE)}
        print(kind, "s", s, "total rel", np.linalg.norm(y - ref) / np.linalg.norm(ref))
        for name, (a, b) in blocks.items():
            d = np.abs(y[a:b] - ref[a:b]); sc = np.abs(ref[a:b]) + 1e-300
            relc = d / sc
            bad = np.nonzero(relc > 1e-9)[0]
            print("  ", name, "rel", np.linalg.norm(d) / np.linalg.norm(ref[a:b]), "nbad", len(bad), "of", b - a)
            for q in bad[:6]:
                if name == "Ez":
                    k, r = divmod(q, nx * nth + 1)
                    print("      idx", q, "k", k, "axis" if r == 0 else ("i %d j %d" % ((r - 1) % nx, (r - 1) // nx)), y[a + q], ref[a + q])
                else:
                    print("      idx", q, "i", q % nx, "j", (q // nx) % nth, "k", q // (nx * nth), y[a + q], ref[a + q])
