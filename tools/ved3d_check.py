"""Known-answer check of the 3-D cylindrical discretisation, GPU only: a vertical electric dipole placed OFF the
mesh axis in a wholespace, FDEM-h at 1 kHz / 10 kHz, against Ward & Hohmann eq. 2.40.  The comparison region
contains the mesh axis, so the run also tells which axis conventions (axis_weight_edge, axis_weight_face --
SURVEY A.4 OQ1, [3P-recall]) agree with the physics.  usage: python tools/ved3d_check.py [nth]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from casingsimulations_b200 import _lib
from oracle.cyl_oracle import OracleCylMesh, mesh_tensor
from test_oracle import ved_wholespace

def run(nth=32, conventions=((1.0, 0.5), (0.5, 0.5), (1.0, 1.0), (0.5, 1.0)), freqs=(1e3, 1e4), verbose=True, mur=1.0):
    """returns {(awe, awf): {(freq, region): relative L2 error over all three components}}"""
    hx = mesh_tensor([(1.0, 40), (1.0, 18, 1.35)])
    hz = mesh_tensor([(1.0, 18, -1.35), (1.0, 80), (1.0, 18, 1.35)])
    hy = np.ones(nth) * 2 * np.pi / nth
    m = OracleCylMesh([hx, hy, hz], [0, 0, -hz.sum() / 2])
    sigma = 0.1
    i0, j0 = 10, nth // 4                                     # source column: r in (10, 11), theta cell j0
    gz3 = m.gridFz.reshape(tuple(m.vnFz) + (3,), order="F")
    r0, th0 = gz3[i0, j0, 0, 0], gz3[i0, j0, 0, 1]
    x0, y0 = r0 * np.cos(th0), r0 * np.sin(th0)
    w = np.zeros(tuple(m.vnFz))
    ksel = np.where((gz3[i0, j0, :, 2] > -2.5) & (gz3[i0, j0, :, 2] < 2.5))[0]
    w[i0, j0, ksel] = 1.0
    zsrc = gz3[i0, j0, ksel, 2]
    se = np.r_[np.zeros(m.nFx + m.nFy), w.ravel(order="F")] / m.area

    def analytic(grid, comp, freq):
        rho, th, z = grid[:, 0], grid[:, 1], grid[:, 2]
        dx, dy = rho * np.cos(th) - x0, rho * np.sin(th) - y0
        rp = np.maximum(np.hypot(dx, dy), 1e-12)
        # (a permeable wholespace mu_r behaves like sigma -> sigma mu_r in k^2 with the 1/sigma prefactor kept)
        ez = sum(ved_wholespace(rp, z, zs, sigma * mur, freq)[0] for zs in zsrc) * mur
        er = sum(ved_wholespace(rp, z, zs, sigma * mur, freq)[1] for zs in zsrc) * mur
        ex, ey = er * dx / rp, er * dy / rp
        if comp == "z":
            return ez
        if comp == "r":
            return ex * np.cos(th) + ey * np.sin(th)
        return -ex * np.sin(th) + ey * np.cos(th)

    out = {}
    for awe, awf in conventions:
        ctx = _lib.Context(0)
        ctx.set_mesh(hx, hy, hz, m.x0[2], axis_weight_edge=awe, axis_weight_face=awf)
        ctx.set_model(ctx.to_device(sigma * np.ones(m.nC)), ctx.to_device(mur * 4e-7 * np.pi * np.ones(m.nC)))
        se_d = ctx.to_device(se, complex_=False)
        u = ctx.solve_fdem("h", np.asarray(freqs, dtype=float), se_d)
        res = {"relres": ctx.info["relres"]}
        for k, freq in enumerate(freqs):
            e = ctx.fields_fdem("h", "e", freq, u[k], se_d).cpu().numpy()
            parts = {"r": (e[:m.nFx], m.gridFx), "t": (e[m.nFx:m.nFx + m.nFy], m.gridFy), "z": (e[m.nFx + m.nFy:], m.gridFz)}
            for region, rmax in (("axis", 4.0), ("shell", None)):
                num = den = 0.0
                for comp, (val, grid) in parts.items():
                    ana = analytic(grid, comp, freq)
                    x, y = grid[:, 0] * np.cos(grid[:, 1]), grid[:, 0] * np.sin(grid[:, 1])
                    R = np.sqrt((x - x0) ** 2 + (y - y0) ** 2 + grid[:, 2] ** 2)
                    sel = (R > 6) & (R < 25)
                    if rmax is not None:
                        sel &= grid[:, 0] < rmax
                    num += np.sum(np.abs(val[sel] - ana[sel]) ** 2)
                    den += np.sum(np.abs(ana[sel]) ** 2)
                res[(freq, region)] = float(np.sqrt(num / den))
        ctx.close()
        out[(awe, awf)] = res
        if verbose:
            print("axis_weight_edge %.1f axis_weight_face %.1f:" % (awe, awf), res)
    return out


if __name__ == "__main__":
    run(int(sys.argv[1]) if len(sys.argv) > 1 else 32, mur=float(sys.argv[2]) if len(sys.argv) > 2 else 1.0,
        freqs=tuple(float(f) for f in sys.argv[3:]) or (1e3, 1e4))
