"""two-front elimination across the frequency range at the bench size (109x32x1394): iterations / residual / fallbacks"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import bench
from casingsimulations_b200 import _lib
mp, mg, src = bench.build_inputs(bench.MESH_OURS, [1.0])
m = mg.mesh
ctx = _lib.Context(0)
ctx.set_mesh(m.hx, m.hy, m.hz, m.x0[2])
se = ctx.to_device(np.real(src.s_e))
ctx.paint(mp.paint_params())
for f in (0.01, 0.1, 100.0, 1e4):
    try:
        u = ctx.solve_fdem("h", np.r_[f], se)
        print(f, "Hz:", {k: round(float(ctx.info[k]), 3) if k != "relres" else float(ctx.info[k]) for k in ("iters", "relres", "ms_factor", "ms_solve")}, flush=True)
    except Exception as ex:
        print(f, "Hz: FAILED", str(ex)[:160], flush=True)
ctx.close()
