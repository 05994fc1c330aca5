"""two-front vs one-front elimination on the reference-arm mesh (109x4x40, p = 330): parity against the exactly assembled
system at several frequencies.  Run once per CS_TWIST value."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import bench
from casingsimulations_b200 import _lib
from oracle.cyl_oracle import OracleCylMesh, OracleFDEM, paint_casing_in_halfspace, exact_mesh, energy_rel, casing_currents
freqs = [0.1, 1.0, 10.0, 100.0, 1e4]
mp, mg, src = bench.build_inputs(bench.MESH_REF, freqs)
m = mg.mesh
ctx = _lib.Context(0)
ctx.set_mesh(m.hx, m.hy, m.hz, m.x0[2])
se = ctx.to_device(np.real(src.s_e))
ctx.paint(mp.paint_params())
u = ctx.solve_fdem("h", mp.freqs, se)
info = dict(ctx.info)
om = OracleCylMesh([m.hx, m.hy, m.hz], m.x0)
sig, mu = paint_casing_in_halfspace(om, sigma_back=mp.sigma_back, sigma_air=mp.sigma_air, sigma_casing=mp.sigma_casing,
                                    sigma_inside=mp.sigma_inside, mur_casing=mp.mur_casing, casing_a=mp.casing_a,
                                    casing_b=mp.casing_b, casing_z=tuple(mp.casing_z), surface_z=mp.surface_z)
PX = OracleFDEM(exact_mesh(om), sig, mu, "h")
P64 = OracleFDEM(om, sig, mu, "h")
se_np = np.real(src.s_e)
rho = np.asarray(om.getFaceInnerProduct(sig, invProp=True).diagonal())
cas = (mp.casing_a, mp.casing_b, tuple(mp.casing_z))
out = {"CS_TWIST": os.environ.get("CS_TWIST"), "iters": info.get("iters"), "relres": info.get("relres")}
for k, f in enumerate(freqs):
    h_ref = PX.solve(f, se_np, refine=True)
    j_ref = np.asarray(PX.j_from(f, h_ref, se_np), dtype=complex)
    j = ctx.fields_fdem("h", "j", f, u[k], se).cpu().numpy()
    h64 = P64.solve(f, se_np)
    j64 = np.asarray(P64.j_from(f, h64, se_np), dtype=complex)
    iz, iz_ref, iz64 = (casing_currents(v, om, *cas)["z"][1] for v in (j, j_ref, j64))
    out["%g Hz" % f] = {"gpu j_energy": float(energy_rel(j, j_ref, rho)), "gpu Iz": float(np.linalg.norm(iz - iz_ref) / np.linalg.norm(iz_ref)),
                        "fp64 SuperLU j_energy": float(energy_rel(j64, j_ref, rho)), "fp64 SuperLU Iz": float(np.linalg.norm(iz64 - iz_ref) / np.linalg.norm(iz_ref))}
print(json.dumps(out))
