"""Casing current / charge integration (kept API of casingSimulations/physics.py:10-83).
The masked, area-weighted sums over (r, theta) run on the GPU (kernel K9); only
the final selection of z levels strictly inside the casing happens on the host."""
import numpy as np

from . import _lib

_ctx_cache = {}


def _ctx_for(mesh, device=0):
    key = (device, mesh.hx.tobytes(), mesh.hy.tobytes(), mesh.hz.tobytes(), float(mesh.x0[2]))
    if key not in _ctx_cache:
        _ctx_cache.clear()
        ctx = _lib.Context(device)
        ctx.set_mesh(mesh.hx, mesh.hy, mesh.hz, mesh.x0[2])
        _ctx_cache[key] = ctx
    return _ctx_cache[key]


def casing_currents(j, mesh, model_parameters, device=0):
    """Current (A) carried by the casing wall: radial leak-off I_x on cell-centre
    z levels and vertical current I_z on z nodes.  ``j`` is (nF, 1) as in the
    reference (physics.py:49-54 require exactly one column)."""
    ctx = _ctx_for(mesh, device)
    j = np.asarray(j)
    assert j.ndim == 2 and j.shape[1] == 1, "j must be (nF, 1)"
    jd = ctx.to_device(j[:, 0])
    cz = model_parameters.casing_z
    ix, iz = ctx.casing_currents(jd, float(model_parameters.casing_a), float(model_parameters.casing_b), cz)
    ix, iz = ix.cpu().numpy()[:, None], iz.cpu().numpy()[:, None]
    ix_inds = (mesh.vectorCCz > cz[0]) & (mesh.vectorCCz < cz[1])
    iz_inds = (mesh.vectorNz > cz[0]) & (mesh.vectorNz < cz[1])
    return {"x": (mesh.vectorCCz[ix_inds], ix[ix_inds]), "z": (mesh.vectorNz[iz_inds], iz[iz_inds])}


def casing_charges(charge, mesh, model_parameters, device=0):
    """Charge per z level within one fine cell of the casing (physics.py:67-83);
    unlike the reference the caller's array is not modified."""
    ctx = _ctx_for(mesh, device)
    cz = model_parameters.casing_z
    c = np.asarray(charge, dtype=float).reshape(-1)
    out = ctx.casing_charges(ctx.to_device(c), float(model_parameters.casing_b), cz).cpu().numpy()
    z_inds = (mesh.vectorCCz > cz[0]) & (mesh.vectorCCz < cz[1])
    return mesh.vectorCCz[z_inds], out[z_inds]
