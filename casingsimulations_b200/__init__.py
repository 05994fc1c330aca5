"""casingsimulations_b200: B200-native replacement of the solve path that
simpeg-research/casingSimulations hands to SimPEG + Pardiso.

Same entry points as the reference package (casingSimulations/__init__.py):
``model``, ``CylMeshGenerator``, ``CasingMeshGenerator``, ``sources``, ``run``,
``casing_currents``, ``casing_charges``, the theta-slice helpers and
``load_properties``; plotting / viewer modules are out of scope.
"""
from . import model
from .mesh import CylMeshGenerator, CasingMeshGenerator, TensorMeshGenerator, CylMesh, meshTensor
from .physics import casing_currents, casing_charges
from . import sources
from . import run
from .solver import Solver
from .utils import (load_properties, edge3DthetaSlice, face3DthetaSlice, ccv3DthetaSlice, mesh2d_from_3d,
                    writeSimulationPy, loadSimulationResults)
from .info import __version__, __author__, __license__
