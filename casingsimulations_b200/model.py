"""Model parameter objects and painters (kept API of casingSimulations/model.py).

Each model knows how to paint sigma / mu_r on cell centres with the strict
inequalities of the reference (model.py:249, 290-293, 514-529, 538).  The numpy
painters below serve the reference's public ``model.sigma(mesh)`` API; inside
``run.Simulation*`` the CasingIn{Whole,Half}space / SingleLayer family is painted
directly on the GPU by kernel K1 (``paint_params``) and never leaves the device.
"""
import numpy as np

from .base import BaseCasing
from .mesh import meshTensor

MU_0 = 4e-7 * np.pi

SIMULATION_PARAMETERS_FILENAME = "ModelParameters.json"
SIGMA_BACK = 1e-2
SIGMA_AIR = 1e-6
SIGMA_CASING = 5.5e6
MUR = 1.0
CASING_L = 1000
CASING_D = 10e-2
CASING_T = 1e-2


class Wholespace(BaseCasing):
    """model.py:56-215 (survey parameters mixed in)"""
    _props = {"filename": SIMULATION_PARAMETERS_FILENAME, "sigma_back": SIGMA_BACK, "mur_back": MUR,
              "freqs": None, "timeSteps": None, "src_a": None, "src_b": None}
    _array_props = ("freqs", "src_a", "src_b")

    def __init__(self, filename=None, **kwargs):
        if filename is not None:
            kwargs["filename"] = filename
        self._init_props(**kwargs)

    def _check_sigma_back(self, v):
        if v < 0:
            raise ValueError("sigma_back must be >= 0 (model.py:123-127)")
        return v

    def _check_mur_back(self, v):
        if v < 0:
            raise ValueError("mur_back must be >= 0")
        return v

    def _check_timeSteps(self, v):
        if v is None:
            return v
        if isinstance(v, list):                      # TimeStepArray, model.py:46-53
            v = meshTensor([tuple(t) if isinstance(t, (list, tuple)) else t for t in v])
        return np.asarray(v, dtype=float)

    def __str__(self):
        return self.info

    @property
    def info_survey(self):
        info = "\n ---- Survey ---- "
        if self.freqs is not None:
            info += "\n   {:1.0f} frequencies. min: {:1.1e} Hz, max: {:1.1e} Hz".format(
                len(self.freqs), self.freqs.min(), self.freqs.max())
        if self.timeSteps is not None:
            info += ("\n   {:1.0f} time steps. min time step: {:1.1e} s, max time step: {:1.1e} s. "
                     "Total time: {:1.1e} s".format(len(self.timeSteps), self.timeSteps.min(),
                                                    self.timeSteps.max(), self.timeSteps.sum()))
        return info

    @property
    def info_model(self):
        return ("\n ---- Model ---- \n\n  background: \n    - conductivity: {:1.1e} S/m"
                "\n    - permeability: {:1.1f} mu_0".format(self.sigma_back, self.mur_back))

    @property
    def info(self):
        info = self.info_survey + "\n\n" + self.info_model
        if hasattr(self, "info_casing"):
            info += "\n\n" + self.info_casing
        return info

    def skin_depth(self, sigma=None, mu=None, f=None):
        sigma = self.sigma_back if sigma is None else sigma
        mu = MU_0 if mu is None else mu
        f = self.freqs if f is None else f
        return np.sqrt(2.0 / (2.0 * np.pi * f * mu * sigma))

    def diffusion_distance(self, t=None, sigma=None, mu=None):
        sigma = self.sigma_back if sigma is None else sigma
        mu = MU_0 if mu is None else mu
        t = self.timeSteps.sum() if t is None else t
        return np.sqrt(2 * t / (mu * sigma))

    # painters -----------------------------------------------------------------
    def sigma(self, mesh):
        return self.sigma_back * np.ones(mesh.nC)

    def mur(self, mesh):
        return self.mur_back * np.ones(mesh.nC)

    def mu(self, mesh):
        return MU_0 * self.mur(mesh)

    def paint_params(self):
        """The 18 doubles consumed by the GPU painter (csrc/cs_kernels.cu k_paint),
        or None when the model has features K1 does not know (targets, flaws,
        arbitrary layer stacks) and must be painted through sigma()/mur()."""
        if type(self).sigma is not Wholespace.sigma or type(self).mur is not Wholespace.mur:
            return None
        p = np.zeros(18)
        p[0], p[11] = self.sigma_back, self.mur_back
        return p


class Halfspace(Wholespace):
    """model.py:218-260"""
    _props = {"sigma_air": SIGMA_AIR, "surface_z": 0.0}

    def ind_air(self, mesh):
        return mesh.gridCC[:, 2] > self.surface_z

    def sigma(self, mesh):
        sigma = super(Halfspace, self).sigma(mesh)
        sigma[self.ind_air(mesh)] = self.sigma_air
        return sigma

    def paint_params(self):
        if type(self).sigma is not Halfspace.sigma or type(self).mur is not Wholespace.mur:
            return None
        p = np.zeros(18)
        p[0], p[11] = self.sigma_back, self.mur_back
        p[1], p[2], p[3] = 1.0, self.sigma_air, self.surface_z
        return p


class SingleLayer(Halfspace):
    """model.py:263-303"""
    _props = {"sigma_layer": SIGMA_BACK, "layer_z": np.r_[-CASING_L, -CASING_L * 0.9]}
    _array_props = ("layer_z",)

    def ind_layer(self, mesh):
        z = mesh.gridCC[:, 2]
        return (z < self.layer_z[1]) & (z > self.layer_z[0])

    def sigma(self, mesh):
        sigma = super(SingleLayer, self).sigma(mesh)
        sigma[self.ind_layer(mesh)] = self.sigma_layer
        return sigma

    def paint_params(self):
        if type(self).sigma is not SingleLayer.sigma or type(self).mur is not Wholespace.mur:
            return None
        p = np.zeros(18)
        p[0], p[11] = self.sigma_back, self.mur_back
        p[1], p[2], p[3] = 1.0, self.sigma_air, self.surface_z
        p[14], p[15], p[16], p[17] = 1.0, self.sigma_layer, self.layer_z[0], self.layer_z[1]
        return p


class Layers(Halfspace):
    """model.py:306-340"""
    _props = {"sigma_layers": [SIGMA_BACK], "layer_tops": [0.0]}

    def sigma(self, mesh):
        sigma = super(Layers, self).sigma(mesh)
        z = mesh.gridCC[:, 2]
        for top, sig in zip(self.layer_tops, self.sigma_layers):
            sigma[z < top] = sig
        return sigma

    def paint_params(self):
        return None


class TargetMixin(object):
    """model.py:343-396"""
    _props = {"target_radius": np.r_[0.0, 25.0], "target_z": np.r_[-925.0, -900.0],
              "target_theta": np.r_[0.0, 2 * np.pi], "sigma_target": SIGMA_BACK}
    _array_props = ("target_radius", "target_z", "target_theta")

    def indx_target(self, mesh):
        r = mesh.gridCC[:, 0]
        return (r >= self.target_radius[0]) & (r <= self.target_radius[1])

    def indy_target(self, mesh):
        t = mesh.gridCC[:, 1]
        return (t >= self.target_theta[0]) & (t <= self.target_theta[1])

    def indz_target(self, mesh):
        z = mesh.gridCC[:, 2]
        return (z >= self.target_z[0]) & (z <= self.target_z[1])

    def ind_target(self, mesh):
        return self.indx_target(mesh) & self.indy_target(mesh) & self.indz_target(mesh)

    def add_sigma_target(self, mesh, sigma):
        sigma[self.ind_target(mesh)] = self.sigma_target
        return sigma


class TargetInHalfspace(TargetMixin, Halfspace):
    def sigma(self, mesh):
        return self.add_sigma_target(mesh, Halfspace.sigma(self, mesh))

    def paint_params(self):
        return None


class CasingMixin(object):
    """model.py:405-580"""
    _props = {"sigma_casing": SIGMA_CASING, "sigma_inside": SIGMA_BACK, "mur_casing": MUR, "mur_inside": MUR,
              "casing_top": 0.0, "casing_l": CASING_L, "casing_d": CASING_D, "casing_t": CASING_T}

    @property
    def info_casing(self):
        return ("\n ---- Casing ---- \n\n  properties: \n    - conductivity: {:1.1e} S/m\n    - permeability: {:1.1f} mu_0"
                "\n    - conductivity inside: {:1.1e} S/m\n\n  geometry: \n    - casing top: {:1.1f} m"
                "\n    - casing length: {:1.1f} m\n    - casing diameter: {:1.1e} m\n    - casing thickness: {:1.1e} m".format(
                    self.sigma_casing, self.mur_casing, self.sigma_inside, self.casing_top, self.casing_l,
                    self.casing_d, self.casing_t))

    @property
    def casing_r(self):
        return self.casing_d / 2.0

    @property
    def casing_a(self):
        return self.casing_r - self.casing_t / 2.0

    @property
    def casing_b(self):
        return self.casing_r + self.casing_t / 2.0

    @property
    def casing_z(self):
        return np.r_[-self.casing_l, 0.0] + self.casing_top

    def indx_casing(self, mesh):
        r = mesh.gridCC[:, 0]
        return (r > self.casing_a) & (r < self.casing_b)

    def indz_casing(self, mesh):
        z = mesh.gridCC[:, 2]
        return (z > self.casing_z[0]) & (z < self.casing_z[1])

    def indx_inside(self, mesh):
        return mesh.gridCC[:, 0] < self.casing_a

    def ind_casing(self, mesh):
        return self.indx_casing(mesh) & self.indz_casing(mesh)

    def ind_inside(self, mesh):
        return self.indx_inside(mesh) & self.indz_casing(mesh)

    def add_sigma_casing(self, mesh, sigma):
        sigma[self.ind_casing(mesh)] = self.sigma_casing      # casing first,
        sigma[self.ind_inside(mesh)] = self.sigma_inside      # then the fluid inside (model.py:566-567)
        return sigma

    def add_mur_casing(self, mesh, mur):
        mur[self.ind_casing(mesh)] = self.mur_casing
        mur[self.ind_inside(mesh)] = self.mur_inside
        return mur

    def _casing_params(self, p):
        p[4], p[5], p[6] = 1.0, self.sigma_casing, self.sigma_inside
        p[7], p[8], p[9], p[10] = self.casing_a, self.casing_b, self.casing_z[0], self.casing_z[1]
        p[12], p[13] = self.mur_casing, self.mur_inside
        return p


class FlawedCasingMixin(CasingMixin):
    """model.py:583-668"""
    _props = {"flaw_r": np.r_[0.0, 0.0], "flaw_theta": np.r_[0.0, 2 * np.pi], "flaw_z": np.r_[0.0, 0.0],
              "sigma_flaw": SIGMA_CASING, "mur_flaw": MUR}
    _array_props = ("flaw_r", "flaw_theta", "flaw_z")

    def indices_flaw(self, mesh):
        cc = mesh.gridCC
        return ((cc[:, 0] >= self.flaw_r[0]) & (cc[:, 0] <= self.flaw_r[1]) &
                (cc[:, 1] >= self.flaw_theta[0]) & (cc[:, 1] <= self.flaw_theta[1]) &
                (cc[:, 2] >= self.flaw_z[0]) & (cc[:, 2] <= self.flaw_z[1]))

    def add_sigma_casing(self, mesh, sigma):
        sigma = CasingMixin.add_sigma_casing(self, mesh, sigma)
        sigma[self.indices_flaw(mesh)] = self.sigma_flaw
        return sigma

    def add_mur_casing(self, mesh, mur):
        mur = CasingMixin.add_mur_casing(self, mesh, mur)
        mur[self.indices_flaw(mesh)] = self.mur_flaw
        return mur


def _bg_params(cls):
    """background painter parameters evaluated on a casing subclass instance"""
    def f(self):
        p = np.zeros(18)
        p[0], p[11] = self.sigma_back, self.mur_back
        if isinstance(self, Halfspace):
            p[1], p[2], p[3] = 1.0, self.sigma_air, self.surface_z
        if isinstance(self, SingleLayer):
            p[14], p[15], p[16], p[17] = 1.0, self.sigma_layer, self.layer_z[0], self.layer_z[1]
        return p
    return f


class CasingInWholespace(CasingMixin, Wholespace):
    """model.py:670-692"""

    def sigma(self, mesh):
        return self.add_sigma_casing(mesh, Wholespace.sigma(self, mesh))

    def mur(self, mesh):
        return self.add_mur_casing(mesh, Wholespace.mur(self, mesh))

    def paint_params(self):
        return self._casing_params(_bg_params(Wholespace)(self))


class CasingInHalfspace(CasingMixin, Halfspace):
    """model.py:695-717"""

    def sigma(self, mesh):
        return self.add_sigma_casing(mesh, Halfspace.sigma(self, mesh))

    def mur(self, mesh):
        return self.add_mur_casing(mesh, Wholespace.mur(self, mesh))

    def paint_params(self):
        return self._casing_params(_bg_params(Halfspace)(self))


class FlawedCasingInHalfspace(FlawedCasingMixin, Halfspace):
    """model.py:720-723"""

    def sigma(self, mesh):
        return self.add_sigma_casing(mesh, Halfspace.sigma(self, mesh))

    def mur(self, mesh):
        return self.add_mur_casing(mesh, Wholespace.mur(self, mesh))

    def paint_params(self):
        return None


class CasingInHalfspaceWithTarget(CasingMixin, TargetInHalfspace):
    """model.py:725-738 (the reference paints no casing permeability here)"""

    def sigma(self, mesh):
        return self.add_sigma_casing(mesh, TargetInHalfspace.sigma(self, mesh))

    def mur(self, mesh):
        return Wholespace.mur(self, mesh)

    def paint_params(self):
        return None


class CasingInSingleLayer(CasingMixin, SingleLayer):
    """model.py:740-762"""

    def sigma(self, mesh):
        return self.add_sigma_casing(mesh, SingleLayer.sigma(self, mesh))

    def mur(self, mesh):
        return self.add_mur_casing(mesh, Wholespace.mur(self, mesh))

    def paint_params(self):
        return self._casing_params(_bg_params(SingleLayer)(self))


class FlawedCasingInSingleLayer(FlawedCasingMixin, SingleLayer):
    """model.py:765-768"""

    def sigma(self, mesh):
        return self.add_sigma_casing(mesh, SingleLayer.sigma(self, mesh))

    def mur(self, mesh):
        return self.add_mur_casing(mesh, Wholespace.mur(self, mesh))

    def paint_params(self):
        return None


class CasingInLayers(CasingMixin, Layers):
    """model.py:771-794"""

    def sigma(self, mesh):
        return self.add_sigma_casing(mesh, Layers.sigma(self, mesh))

    def mur(self, mesh):
        return self.add_mur_casing(mesh, Wholespace.mur(self, mesh))

    def paint_params(self):
        return None


class FlawedCasingInLayers(FlawedCasingMixin, Layers):
    """model.py:797-800"""

    def sigma(self, mesh):
        return self.add_sigma_casing(mesh, Layers.sigma(self, mesh))

    def mur(self, mesh):
        return self.add_mur_casing(mesh, Wholespace.mur(self, mesh))

    def paint_params(self):
        return None


class Wires(object):
    """SimPEG.maps.Wires stand-in (model.py:857-868): named slices of the model vector."""

    def __init__(self, *args):
        self._slices, start = {}, 0
        for name, n in args:
            self._slices[name] = slice(start, start + n)
            start += n
        self.nP = start

    def __getattr__(self, name):
        sl = self.__dict__.get("_slices", {})
        if name in sl:
            s = sl[name]
            return lambda m: np.asarray(m)[s]
        raise AttributeError(name)


class PhysicalProperties(object):
    """model.py:808-868 (plotting omitted)"""

    def __init__(self, meshGenerator, modelParameters):
        self.meshGenerator = meshGenerator
        self.mesh = meshGenerator.mesh
        self.modelParameters = modelParameters

    @property
    def mur(self):
        if getattr(self, "_mur", None) is None:
            self._mur = self.modelParameters.mur(self.mesh)
        return self._mur

    @property
    def mu(self):
        return MU_0 * self.mur

    @property
    def sigma(self):
        if getattr(self, "_sigma", None) is None:
            self._sigma = self.modelParameters.sigma(self.mesh)
        return self._sigma

    @property
    def model(self):
        return np.hstack([self.sigma, self.mu])

    @property
    def wires(self):
        if getattr(self, "_wires", None) is None:
            self._wires = Wires(("sigma", self.mesh.nC), ("mu", self.mesh.nC))
        return self._wires
