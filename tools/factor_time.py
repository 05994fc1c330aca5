"""times cs_op_factor alone (CS_FACTOR_LOG=1 prints the device time of every call): python tools/factor_time.py NTH CSZ NPADZ [REPS]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import casingsimulations_b200 as cs
from casingsimulations_b200 import _lib
nth = int(sys.argv[1]); csz = float(sys.argv[2]); npadz = int(sys.argv[3]); reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
mp = cs.model.CasingInHalfspace(src_a=np.r_[0.0, 0.0, 0.0], src_b=np.r_[1000.0, 0.0, 0.0], mur_casing=100.0)
mg = cs.CasingMeshGenerator(modelParameters=mp, npadx=10, npadz=npadz, csz=csz, hy=np.ones(nth) * 2 * np.pi / nth)
m = mg.mesh
ctx = _lib.Context(0)
ctx.set_mesh(m.hx, m.hy, m.hz, m.x0[2])
ctx.paint(mp.paint_params())
op = ctx.op(_lib.KIND_EDGE_H)
for r in range(reps):
    op.factor([2j * np.pi])
ctx.sync()
ctx.close()
