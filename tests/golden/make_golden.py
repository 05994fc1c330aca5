"""Generates tests/golden/small_casing.npz from the CPU oracle (refined "truth" solves).

The reference itself cannot be imported in this container (SimPEG / discretize /
pymatsolver absent), so these vectors pin the ORACLE against
regressions and let the GPU tests run without re-solving on the CPU; they are not
outputs of the reference.  Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from conftest import CASING, small_casing_mesh  # noqa: E402
from oracle.cyl_oracle import (OracleCylMesh, OracleDC, OracleFDEM, OracleTDEM, paint_casing_in_halfspace)  # noqa: E402


def source_wire(om):
    g = om.gridFz
    w = np.zeros(om.nFz)
    sel = (g[:, 0] < om.hx.min()) & (g[:, 2] > -2.6) & (g[:, 2] < 0.1)
    if not om.isSymmetric:
        sel &= g[:, 1] < om.hy[0]
    w[sel] = 1.0
    return np.r_[np.zeros(om.nFx + om.nFy), w] / om.area


out = {}
for nth in (1, 4):
    hx, hy, hz, z0 = small_casing_mesh(nth)
    om = OracleCylMesh([hx, hy, hz], [0, 0, z0])
    sig, mu = paint_casing_in_halfspace(om, **CASING)
    se = source_wire(om)
    tag = "n%d_" % nth
    out[tag + "sigma"], out[tag + "mu"], out[tag + "s_e"] = sig, mu, se
    out[tag + "MfRho"] = om.getFaceInnerProduct(sig, invProp=True).diagonal()
    out[tag + "MeMu"] = om.getEdgeInnerProduct(mu).diagonal()
    a = np.array([[0.03, np.pi / 4 if nth > 1 else 0.0, -0.25]])
    b = np.array([[2.0, 5 * np.pi / 4 if nth > 1 else 0.0, -0.25]])
    out[tag + "dc_a"], out[tag + "dc_b"] = a, b
    out[tag + "dc_phi"] = OracleDC(om, sig).fields(a, b, refine=True)[:, 0]
    P = OracleFDEM(om, sig, mu, "h")
    h = P.solve(100.0, se, refine=True)
    out[tag + "fdem_h_100Hz"] = h
    out[tag + "fdem_j_100Hz"] = P.j_from(100.0, h, se)
    T = OracleTDEM(om, sig, mu, [(1e-5, 3), (1e-4, 3)])
    out[tag + "tdem_j"] = T.fields(se, refine=True)
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "small_casing.npz"), **out)
print({k: v.shape for k, v in out.items()})
