"""3-D FDEM-h casing solve on one B200 (the north-star headline: per-frequency 3-D solve vs the
host direct solver).  Usage: python tools/bench_3d_fdem.py NTH CSZ NPADZ [cpu]
Prints GPU assemble / factor / solve times per frequency and, with 'cpu', the oracle's
scipy-assembly + SuperLU time for ONE frequency on the same mesh (may take minutes / tens of GB)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import casingsimulations_b200 as cs
from casingsimulations_b200 import _lib

nth = int(sys.argv[1]); csz = float(sys.argv[2]); npadz = int(sys.argv[3])
do_cpu = len(sys.argv) > 4 and sys.argv[4] == "cpu"
mp = cs.model.CasingInHalfspace(src_a=np.r_[0.0, 0.0, 0.0], src_b=np.r_[1000.0, 0.0, 0.0], mur_casing=100.0, freqs=np.r_[1.0, 10.0])
mg = cs.CasingMeshGenerator(modelParameters=mp, npadx=10, npadz=npadz, csz=csz, hy=np.ones(nth) * 2 * np.pi / nth)
m = mg.mesh
th = m.vectorCCy[nth // 2]
mp.src_a, mp.src_b = np.r_[0.0, th, 0.0], np.r_[1000.0, th, 0.0]
src = cs.sources.TopCasingSrc(modelParameters=mp, meshGenerator=mg, physics="fdem")
se = src.s_e
print("mesh", tuple(m.vnC), "cells", m.nC, "edges", m.nE, flush=True)
ctx = _lib.Context(0)
ctx.set_mesh(m.hx, m.hy, m.hz, m.x0[2])
se_d = ctx.to_device(se)
freqs = np.r_[1.0, 10.0][:int(os.environ.get('NFREQ', '2'))]
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ctx.paint(mp.paint_params())
    u = ctx.solve_fdem("h", freqs, se_d, rtol=1e-10)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    print("GPU rep", rep, "wall s per frequency", (t1 - t0) / len(freqs), {k: round(float(v), 3) for k, v in ctx.info.items()}, flush=True)
j = ctx.fields_fdem("h", "j", freqs[0], u[0], se_d)
ix, iz = ctx.casing_currents(j, mp.casing_a, mp.casing_b, mp.casing_z)
kz = int(np.argmin(np.abs(m.vectorNz + 2.0 * csz)))          # node two cells below the casing top
print("vertical casing current two cells below the casing top (A):", complex(iz[kz].cpu()))
if do_cpu:
    from oracle.cyl_oracle import OracleCylMesh, OracleFDEM, paint_casing_in_halfspace
    t0 = time.perf_counter()
    om = OracleCylMesh([m.hx, m.hy, m.hz], m.x0)
    sig, mu = paint_casing_in_halfspace(om, sigma_back=mp.sigma_back, sigma_air=mp.sigma_air, sigma_casing=mp.sigma_casing,
                                        sigma_inside=mp.sigma_inside, mur_casing=mp.mur_casing, casing_a=mp.casing_a,
                                        casing_b=mp.casing_b, casing_z=tuple(mp.casing_z))
    P = OracleFDEM(om, sig, mu, "h")
    t1 = time.perf_counter()
    h = P.solve(freqs[0], se)
    t2 = time.perf_counter()
    print("CPU oracle: assembly s", t1 - t0, "SuperLU factor+solve s (one frequency)", t2 - t1, flush=True)
    jg = j.cpu().numpy(); jc = P.j_from(freqs[0], h, se)
    A, rhs = P.getA(freqs[0]), P.getRHS(freqs[0], se)
    d = np.sqrt(np.abs(A.diagonal()))
    print("scaled residual: gpu", np.linalg.norm((A @ u[0].cpu().numpy() - rhs) / d) / np.linalg.norm(rhs / d),
          "cpu", np.linalg.norm((A @ h - rhs) / d) / np.linalg.norm(rhs / d), "j rel diff gpu vs cpu", np.linalg.norm(jg - jc) / np.linalg.norm(jc))
ctx.close()
