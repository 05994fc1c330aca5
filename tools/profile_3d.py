"""One 3-D FDEM-h casing solve (paint -> assemble -> factor -> solve) for ncu:
    python tools/profile_3d.py NTH CSZ NPADZ [NFREQ]
Prints the library's per-phase timers; the launch list / --set full captures of this command are summarised under
profiles/ (the per-step kernel grids depend on nx and nth only, so a short mesh in z shows the same kernels as C3)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import casingsimulations_b200 as cs
from casingsimulations_b200 import _lib

nth = int(sys.argv[1]); csz = float(sys.argv[2]); npadz = int(sys.argv[3])
nfreq = int(sys.argv[4]) if len(sys.argv) > 4 else 1
mp = cs.model.CasingInHalfspace(src_a=np.r_[0.0, 0.0, 0.0], src_b=np.r_[1000.0, 0.0, 0.0], mur_casing=100.0)
mg = cs.CasingMeshGenerator(modelParameters=mp, npadx=10, npadz=npadz, csz=csz, hy=np.ones(nth) * 2 * np.pi / nth)
m = mg.mesh
th = m.vectorCCy[nth // 2]
mp.src_a, mp.src_b = np.r_[0.0, th, 0.0], np.r_[1000.0, th, 0.0]
src = cs.sources.TopCasingSrc(modelParameters=mp, meshGenerator=mg, physics="fdem")
ctx = _lib.Context(0)
ctx.set_mesh(m.hx, m.hy, m.hz, m.x0[2])
se_d = ctx.to_device(src.s_e)
freqs = np.logspace(0, 1, nfreq) if nfreq > 1 else np.r_[1.0]
for rep in range(int(os.environ.get("REPS", "1"))):
    ctx.paint(mp.paint_params())
    u = ctx.solve_fdem("h", freqs, se_d, rtol=1e-10)
    torch.cuda.synchronize()
    print("mesh", tuple(int(v) for v in m.vnC), "cells", m.nC, "edges", m.nE, {k: round(float(v), 3) for k, v in ctx.info.items()}, flush=True)
ctx.close()
