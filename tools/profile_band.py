"""One factorisation + one preconditioner solve of nsys C2-size FDEM-h systems: the short program
that ncu profiles for the band kernels (k_lu_panel, k_lu_update, k_band_solve)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import casingsimulations_b200 as cs
from casingsimulations_b200 import _lib

nsys = int(sys.argv[1]) if len(sys.argv) > 1 else 8
mp = cs.model.CasingInHalfspace(src_a=np.r_[0.0, 0.0, 0.0], src_b=np.r_[1000.0, 0.0, 0.0])
mg = cs.CasingMeshGenerator(modelParameters=mp, npadx=10, npadz=20, csz=1.5)
m = mg.mesh
ctx = _lib.Context(0)
ctx.set_mesh(m.hx, m.hy, m.hz, m.x0[2])
ctx.paint(mp.paint_params())
op = ctx.op(_lib.KIND_EDGE_H)
freqs = np.logspace(-1, 3, nsys)
import time
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    op.factor(2j * np.pi * freqs)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    rhs = torch.randn((nsys, op.n), dtype=torch.complex128, device=ctx.device)
    x = op.solve(rhs, shift_of_vec=np.arange(nsys), rtol=1e-8)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print("nsys", nsys, "factor ms", (t1 - t0) * 1e3, "solve ms", (t2 - t1) * 1e3, op.info)
ctx.close()
