"""How the TDEM-e per-step error against the exactly assembled systems depends on the Krylov tolerance (experiment)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
sys.path.insert(0, os.path.join(sys.path[0], "tests"))
from test_gpu_parity import setup, rel
from casingsimulations_b200 import _lib
from oracle.cyl_oracle import OracleTDEM, exact_mesh, energy_rel
ctx = _lib.Context(0)
om, sig, mu, *_ = setup(ctx, 4)
rng = np.random.default_rng(21)
se = np.zeros(om.nE)
ezi = om.nEx + om.nEy + np.arange(om.nEz)
se[ezi[rng.choice(om.nEz, 10, replace=False)]] = rng.standard_normal(10)
steps = [(1e-5, 3), (1e-4, 3)]
P = OracleTDEM(om, sig, mu, steps, form="e")
Fref = OracleTDEM(exact_mesh(om), sig, mu, steps, form="e").fields(se, refine=True)
wgt = np.asarray(P.MeSigma.diagonal())
for rtol in (1e-10, 1e-12, 1e-14):
    try:
        F = ctx.tdem_run("e", P.time_steps, ctx.to_device(se), rtol=rtol, maxit=200).cpu().numpy().T
    except Exception as ex:
        print(rtol, "raised", str(ex)[:200]); continue
    print(rtol, "iters", ctx.info.get("iters"), "relres", ctx.info.get("relres"),
          ["%.1e/%.1e" % (rel(F[:, t], Fref[:, t]), energy_rel(F[:, t], Fref[:, t], wgt)) for t in range(Fref.shape[1])])
