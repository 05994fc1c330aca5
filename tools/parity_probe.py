"""GPU solution vs the EXACTLY assembled oracle system (oracle.exact_mesh + long-double refinement) and vs the
fp64-assembled oracle's refined / plain SuperLU solutions, on the small casing meshes of tests/conftest.py."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import CASING, paint_params, small_casing_mesh
from casingsimulations_b200 import _lib
from oracle.cyl_oracle import *

rel = lambda u, v: float(np.linalg.norm(np.asarray(u) - np.asarray(v)) / np.linalg.norm(np.asarray(v)))
ctx = _lib.Context(0)
for nth in (1, 4):
    hx, hy, hz, z0 = small_casing_mesh(nth)
    om = OracleCylMesh([hx, hy, hz], [0, 0, z0]); ox = exact_mesh(om)
    ctx.set_mesh(hx, hy, hz, z0)
    sig, mu = paint_casing_in_halfspace(om, **CASING)
    ctx.paint(paint_params(CASING))
    # DC
    a = np.array([[0.03, np.pi / 4 if nth > 1 else 0.0, -0.25]]); b = np.array([[2.0, 5 * np.pi / 4 if nth > 1 else 0.0, -0.25]])
    Pd, Px = OracleDC(om, sig), OracleDC(ox, sig)
    phi = ctx.solve_dc(ctx.to_device(Pd.getRHS(a, b).T.copy())).cpu().numpy().T
    x_ex, x_ref, x_lu = Px.fields(a, b, refine=True), Pd.fields(a, b, refine=True), Pd.fields(a, b)
    print("nth", nth, "DC   phi: gpu-vs-exact %.2e | asmtruth-vs-exact %.2e | splu-vs-exact %.2e | j gpu-vs-exact %.2e" % (
        rel(phi, x_ex), rel(x_ref, x_ex), rel(x_lu, x_ex), rel(Px.j(phi), Px.j(x_ex))), flush=True)
    g = om.gridFz; w = np.zeros(om.nFz)
    sel = (g[:, 0] < om.hx.min()) & (g[:, 2] > -2.6) & (g[:, 2] < 0.1)
    if nth > 1: sel &= g[:, 1] < hy[0]
    w[sel] = 1.0
    se = np.r_[np.zeros(om.nFx + om.nFy), w] / om.area
    P, PX = OracleFDEM(om, sig, mu, "h"), OracleFDEM(ox, sig, mu, "h")
    freqs = np.r_[0.1, 1.0, 100.0, 1e4]
    u = ctx.solve_fdem("h", freqs, ctx.to_device(se)).cpu().numpy()
    print("   fdem info", {k: ctx.info[k] for k in ("iters", "relres", "dd_steps", "dd_corr")})
    for f, freq in enumerate(freqs):
        h_x, h_lu = PX.solve(freq, se, refine=True), P.solve(freq, se)
        jx = PX.j_from(freq, h_x, se)
        line = "   FDEM-h %8g Hz: h gpu-vs-exact %.2e (splu %.2e) | j gpu %.2e (splu %.2e) | e gpu %.2e" % (
            freq, rel(u[f], h_x), rel(h_lu, h_x), rel(PX.j_from(freq, u[f], se), jx), rel(P.j_from(freq, h_lu, se), jx),
            rel(PX.e_from(freq, u[f], se), PX.e_from(freq, h_x, se)))
        print(line, flush=True)
    # TDEM
    steps = [(1e-5, 3), (1e-4, 3)]
    Pt, PtX = OracleTDEM(om, sig, mu, steps), OracleTDEM(ox, sig, mu, steps)
    F = ctx.tdem_run("j", Pt.time_steps, ctx.to_device(se)).cpu().numpy().T
    Fx, Flu = PtX.fields(se, refine=True), Pt.fields(se)
    print("   TDEM-j: gpu-vs-exact %.2e | splu-vs-exact %.2e" % (rel(F, Fx), rel(Flu, Fx)), flush=True)
ctx.close()
