"""theta-slice helpers that pin the vector layout (kept API of
casingSimulations/utils.py:28-96) and the JSON loader."""
import numpy as np

from .base import load_properties  # noqa: F401  (utils.py:11-24)
from .mesh import CylMesh


def _mkvc(x, ndim=1):
    v = np.asarray(x).reshape(-1, order="F")
    return v if ndim == 1 else v[:, None]


def face3DthetaSlice(mesh3D, j3D, theta_ind=0):
    """(x, z) face components of one theta column, as a 2-D simulation would
    return them (utils.py:28-49)."""
    if mesh3D.isSymmetric:
        return j3D
    j3D = np.asarray(j3D)
    jx = j3D[:mesh3D.nFx].reshape(tuple(mesh3D.vnFx), order="F")
    jz = j3D[mesh3D.nFx + mesh3D.nFy:].reshape(tuple(mesh3D.vnFz), order="F")
    return np.vstack([_mkvc(jx[:, theta_ind, :], 2), _mkvc(jz[:, theta_ind, :], 2)])


def edge3DthetaSlice(mesh3D, h3D, theta_ind=0):
    """theta edge component of one theta column (utils.py:52-67)."""
    h3D = np.asarray(h3D)
    hy = h3D[mesh3D.nEx:mesh3D.nEx + mesh3D.nEy].reshape(tuple(mesh3D.vnEy), order="F")
    return _mkvc(hy[:, theta_ind, :])


def ccv3DthetaSlice(mesh3D, v3D, theta_ind=0):
    """cell-centred vector (3 stacked components) on one theta column (utils.py:70-87)."""
    v3D = np.asarray(v3D)
    nC, shape = mesh3D.nC, tuple(mesh3D.vnC)
    comps = [v3D[c * nC:(c + 1) * nC].reshape(shape, order="F") for c in range(3)]
    return np.vstack([_mkvc(c[:, theta_ind, :], 2) for c in comps])


def mesh2d_from_3d(mesh):
    """utils.py:89-96"""
    return CylMesh([mesh.hx, 1.0, mesh.hz], x0=mesh.x0)
