"""Runs only the stencil applies (K3 DC fp64, K4 edge-form complex128, K4' face-form fp64) on the
109x32x1394 mesh a few times: the short program that ncu profiles (B200_PROFILING.md recipe)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import casingsimulations_b200 as cs
from casingsimulations_b200 import _lib

nth = int(sys.argv[1]) if len(sys.argv) > 1 else 32
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
mp = cs.model.CasingInHalfspace(src_a=np.r_[0.0, 0.0, 0.0], src_b=np.r_[1000.0, 0.0, 0.0], mur_casing=100.0)
hy = np.r_[2 * np.pi] if nth == 1 else np.ones(nth) * 2 * np.pi / nth
mg = cs.CasingMeshGenerator(modelParameters=mp, npadx=10, npadz=25, csz=0.75, hy=hy)
m = mg.mesh
ctx = _lib.Context(0)
ctx.set_mesh(m.hx, m.hy, m.hz, m.x0[2])
ctx.paint(mp.paint_params())
peak = 6547.5
for kind, cplx, bpc, name in ((_lib.KIND_EDGE_H, True, 144.0, "edge_op c128"), (_lib.KIND_DC, False, 40.0, "dc_apply f64"),
                              (_lib.KIND_FACE_J, False, 96.0, "face_op f64"), (_lib.KIND_EDGE_H, False, 96.0, "edge_op f64")):
    op = ctx.op(kind, plan=False)
    x = torch.randn(op.n, dtype=torch.complex128 if cplx else torch.float64, device=ctx.device)
    torch.cuda.synchronize()
    ms = op.apply_timed(x, 2j * np.pi if cplx else 1e3, reps)
    if nth == 1:
        bpc = {144.0: 56.0, 40.0: 32.0, 96.0: 40.0}[bpc] if not (kind == _lib.KIND_FACE_J) else 64.0
    gbs = bpc * ctx.nC / (ms * 1e-3) / 1e9
    print("{:14s} nC={} {:8.1f} us  {:7.1f} GB/s  {:.3f} of measured peak".format(name, ctx.nC, ms * 1e3, gbs, gbs / peak))
ctx.close()
