"""BASELINE config 4: TDEM (default j formulation, backward Euler) on the 3-D mesh 109 x 32 x 460
(1 604 480 cells), 200 steps over 5 distinct dt, TopCasingSrc at a cell-centre azimuth.
Usage: python tools/run_c4_tdem.py [nsteps_per_dt]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import casingsimulations_b200 as cs
from casingsimulations_b200 import _lib

nper = int(sys.argv[1]) if len(sys.argv) > 1 else 40
nth = int(sys.argv[2]) if len(sys.argv) > 2 else 32
steps = [(1e-6, nper), (1e-5, nper), (1e-4, nper), (1e-3, nper), (1e-2, nper)]
mp = cs.model.CasingInHalfspace(src_a=np.r_[0.0, 0.0, 0.0], src_b=np.r_[1000.0, 0.0, 0.0], mur_casing=100.0, timeSteps=steps)
mg = cs.CasingMeshGenerator(modelParameters=mp, npadx=10, npadz=25, csz=2.5, hy=np.ones(nth) * 2 * np.pi / nth)
m = mg.mesh
th = m.vectorCCy[nth // 2 + (1 if nth > 1 else 0) - 1]
mp.src_a, mp.src_b = np.r_[0.0, th, 0.0], np.r_[1000.0, th, 0.0]
src = cs.sources.TopCasingSrc(modelParameters=mp, meshGenerator=mg, physics="tdem")
src.validate()
print("mesh", tuple(int(v) for v in m.vnC), "cells", m.nC, "faces", m.nF, "steps", len(mp.timeSteps), flush=True)
ctx = _lib.Context(0)
ctx.set_mesh(m.hx, m.hy, m.hz, m.x0[2])
ctx.paint(mp.paint_params())
se_d = ctx.to_device(src.s_e)
torch.cuda.synchronize(); t0 = time.perf_counter()
out = ctx.tdem_run("j", mp.timeSteps, se_d, rtol=1e-10)
torch.cuda.synchronize(); t1 = time.perf_counter()
print("TDEM wall s", t1 - t0, {k: round(float(v), 3) for k, v in ctx.info.items()}, flush=True)
print("peak memory GB (torch)", torch.cuda.max_memory_allocated() / 1e9, "free/total GB", [round(v / 1e9, 1) for v in torch.cuda.mem_get_info()])
# size-independent checks: galvanic initial state D (j0 + s_e) = 0 and monotone decay of the ohmic energy
j0 = out[0]
div = ctx.div(j0 + se_d)
print("|D (j0 + s_e)| / |D s_e| =", float(torch.linalg.norm(div) / torch.linalg.norm(ctx.div(se_d))))
alpha = ctx.face_inner_product(ctx.paint(mp.paint_params())[0], True, False)
en = [float((out[t] * alpha * out[t]).sum()) for t in range(1, out.shape[0])]
print("energy monotone:", all(en[i + 1] <= en[i] * (1 + 1e-9) for i in range(len(en) - 1)), "first/last", en[0], en[-1])
print("energy per step:", " ".join("%.4e" % e for e in en[:60]))
dj = [float(torch.linalg.norm(ctx.div(out[t])) / torch.linalg.norm(ctx.div(se_d))) for t in range(1, out.shape[0], max(1, out.shape[0] // 12))]
print("|D j^n| / |D s_e| along the run:", " ".join("%.2e" % v for v in dj))
iz = ctx.casing_currents(out[0] + se_d * 0, mp.casing_a, mp.casing_b, mp.casing_z)[1]
zn = m.vectorNz
k = int(np.argmin(np.abs(zn + 5.0)))
print("DC casing current 5 m below the top (A):", float(iz[k]))
ctx.close()
