"""Generates tests/golden/small_casing.npz from the CPU oracle: solutions of the EXACTLY assembled discrete systems
(oracle.exact_mesh: every operator in np.longdouble from the same fp64 inputs, long-double iterative refinement).

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
from oracle.cyl_oracle import (OracleCylMesh, OracleDC, OracleFDEM, OracleTDEM, paint_casing_in_halfspace, exact_mesh,
                               casing_currents, rhs_rounding_sensitivity)  # noqa: E402


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
    ox = exact_mesh(om)
    out[tag + "dc_phi"] = OracleDC(ox, sig).fields(a, b, refine=True)[:, 0]
    P = OracleFDEM(ox, sig, mu, "h")
    for freq, name in ((1.0, "1Hz"), (100.0, "100Hz")):
        h = P.solve(freq, se, refine=True)
        j = np.asarray(P.j_from(freq, h, se), dtype=complex)
        out[tag + "fdem_h_" + name] = h
        out[tag + "fdem_j_" + name] = j
        out[tag + "fdem_e_" + name] = np.asarray(P.e_from(freq, h, se), dtype=complex)
        out[tag + "fdem_Iz_" + name] = casing_currents(j, om, CASING["casing_a"], CASING["casing_b"], CASING["casing_z"])["z"][1]
        # what one fp64 rounding of the right-hand side alone does to the plain L2 norm of j (see oracle docstring)
        out[tag + "fdem_j_sens_" + name] = rhs_rounding_sensitivity(P.getA(freq), P.getRHS(freq, se), lambda x: P.j_from(freq, x, se))
    T = OracleTDEM(ox, sig, mu, [(1e-5, 3), (1e-4, 3)])
    out[tag + "tdem_j"] = T.fields(se, refine=True)
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "small_casing.npz"), **out)
print({k: np.shape(v) for k, v in out.items()})
print({k: float(v) for k, v in out.items() if 'sens' in k})
