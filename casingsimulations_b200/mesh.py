"""Mesh generators (kept API of casingSimulations/mesh.py) and the minimal
cylindrical mesh value object that stands in for ``discretize.CylMesh``.

Only what the hot path needs is here: 1-D width tables, counts, grids and
measures.  Differential operators and inner products live on the GPU
(csrc/cs_kernels.cu); nothing in this module does linear algebra.
"""
import numpy as np

from .base import BaseCasing


def meshTensor(value):
    """discretize.utils.meshTensor as used at mesh.py:46-61, 448-452, 534-537 and
    model.py:50-53: ``(h, n)`` is n cells of width h, ``(h, n, f)`` expands
    geometrically (h*f, h*f^2, ...), negative f mirrors the expansion."""
    if isinstance(value, np.ndarray):
        return value.astype(float)
    pieces = []
    for item in value:
        if np.isscalar(item):
            pieces.append(np.r_[float(item)])
        elif len(item) == 2:
            pieces.append(np.full(int(item[1]), float(item[0])))
        elif len(item) == 3:
            h, n, fac = float(item[0]), int(item[1]), float(item[2])
            grown = h * np.abs(fac) ** np.arange(1, n + 1, dtype=float)
            pieces.append(grown[::-1] if fac < 0 else grown)
        else:
            raise ValueError("meshTensor entries are h, (h, n) or (h, n, factor)")
    return np.concatenate(pieces) if pieces else np.zeros(0)


class CylMesh(object):
    """Value object for a (r, theta, z) tensor mesh, discretize.CylMesh naming
    (created at mesh.py:87-99).  Ordering: i = r fastest, then theta, then z."""

    def __init__(self, h, x0=None):
        self.hx = np.asarray(h[0], dtype=float)
        hy = h[1]
        self.hy = np.r_[2 * np.pi] if np.isscalar(hy) else np.atleast_1d(np.asarray(hy, dtype=float))
        self.hz = np.asarray(h[2], dtype=float)
        self.x0 = np.zeros(3) if x0 is None else np.asarray(x0, dtype=float)
        if abs(self.hy.sum() - 2 * np.pi) > 1e-6:
            raise AssertionError("hy must sum to 2*pi")
        nx, nt, nz = len(self.hx), len(self.hy), len(self.hz)
        self.vnC = np.r_[nx, nt, nz]
        self.nC = nx * nt * nz
        self.isSymmetric = nt == 1
        self.dim = 3
        if self.isSymmetric:
            self.vnFx, self.vnFy, self.vnFz = np.r_[nx, 1, nz], np.r_[nx, 0, nz], np.r_[nx, 1, nz + 1]
            self.vnEx, self.vnEy, self.vnEz = np.r_[nx, 0, nz + 1], np.r_[nx, 1, nz + 1], None
            self.vnF = np.r_[nx * nz, 0, nx * (nz + 1)]
            self.vnE = np.r_[0, nx * (nz + 1), 0]
        else:
            self.vnFx, self.vnFy, self.vnFz = np.r_[nx, nt, nz], np.r_[nx, nt, nz], np.r_[nx, nt, nz + 1]
            self.vnEx, self.vnEy, self.vnEz = np.r_[nx, nt, nz + 1], np.r_[nx, nt, nz + 1], None
            self.vnF = np.r_[nx * nt * nz, nx * nt * nz, nx * nt * (nz + 1)]
            self.vnE = np.r_[nx * nt * (nz + 1), nx * nt * (nz + 1), nz * (nx * nt + 1)]
        self.nFx, self.nFy, self.nFz = (int(v) for v in self.vnF)
        self.nEx, self.nEy, self.nEz = (int(v) for v in self.vnE)
        self.nF, self.nE = int(self.vnF.sum()), int(self.vnE.sum())
        self._rnodes = np.r_[0.0, np.cumsum(self.hx)]

    # 1-D vectors
    @property
    def vectorNx(self):
        return self._rnodes[1:]

    @property
    def vectorCCx(self):
        return 0.5 * (self._rnodes[1:] + self._rnodes[:-1])

    @property
    def vectorNy(self):
        return np.r_[0.0, np.cumsum(self.hy)[:-1]]

    @property
    def vectorCCy(self):
        return np.r_[0.0] if self.isSymmetric else self.vectorNy + 0.5 * self.hy

    @property
    def vectorNz(self):
        return self.x0[2] + np.r_[0.0, np.cumsum(self.hz)]

    @property
    def vectorCCz(self):
        n = self.vectorNz
        return 0.5 * (n[1:] + n[:-1])

    @staticmethod
    def _grid(x, y, z):
        X, Y, Z = np.meshgrid(x, y, z, indexing="ij")
        return np.c_[X.reshape(-1, order="F"), Y.reshape(-1, order="F"), Z.reshape(-1, order="F")]

    def _cached(self, key, fn):
        cache = self.__dict__.setdefault("_cache", {})
        if key not in cache:
            cache[key] = fn()
        return cache[key]

    @property
    def gridCC(self):
        return self._cached("cc", lambda: self._grid(self.vectorCCx, self.vectorCCy, self.vectorCCz))

    cell_centers = gridCC

    @property
    def gridFx(self):
        return self._cached("fx", lambda: self._grid(self.vectorNx, self.vectorCCy, self.vectorCCz))

    @property
    def gridFy(self):
        if self.isSymmetric:
            return np.zeros((0, 3))
        return self._cached("fy", lambda: self._grid(self.vectorCCx, self.vectorNy, self.vectorCCz))

    @property
    def gridFz(self):
        return self._cached("fz", lambda: self._grid(self.vectorCCx, self.vectorCCy, self.vectorNz))

    @property
    def gridEx(self):
        if self.isSymmetric:
            return np.zeros((0, 3))
        return self._cached("ex", lambda: self._grid(self.vectorCCx, self.vectorNy, self.vectorNz))

    @property
    def gridEy(self):
        return self._cached("ey", lambda: self._grid(self.vectorNx, self.vectorCCy, self.vectorNz))

    # measures
    @property
    def _ring(self):
        return 0.5 * (self._rnodes[1:] ** 2 - self._rnodes[:-1] ** 2)

    @property
    def vol(self):
        return np.kron(self.hz, np.kron(self.hy, self._ring))

    @property
    def area(self):
        return self._cached("area", self._area)

    def _area(self):
        ax = np.kron(self.hz, np.kron(self.hy, self.vectorNx))
        az = np.kron(np.ones(len(self.hz) + 1), np.kron(self.hy, self._ring))
        if self.isSymmetric:
            return np.r_[ax, az]
        ay = np.kron(self.hz, np.kron(np.ones(len(self.hy)), self.hx))
        return np.r_[ax, ay, az]

    @property
    def edge(self):
        ey = np.kron(np.ones(len(self.hz) + 1), np.kron(self.hy, self.vectorNx))
        if self.isSymmetric:
            return ey
        ex = np.kron(np.ones(len(self.hz) + 1), np.kron(np.ones(len(self.hy)), self.hx))
        ez = np.kron(self.hz, np.r_[1.0, np.ones(len(self.hx) * len(self.hy))])
        return np.r_[ex, ey, ez]


def closestPoints(mesh, pts, gridLoc="CC"):
    """discretize.utils.closestPoints: index of the grid point nearest (raw
    (r, theta, z) Euclidean distance) to each query point (sources.py:435, 449)."""
    grid = getattr(mesh, "grid" + gridLoc)
    pts = np.atleast_2d(np.asarray(pts, dtype=float))
    return np.array([int(np.argmin(((grid - p) ** 2).sum(axis=1))) for p in pts])


def pad_for_casing_and_data(casing_b=None, csx1=2.5e-3, csx2=25, pfx1=1.3, pfx2=1.5, domain_x=1000, npadx=10):
    """Radial widths for casing problems (mesh.py:32-63): a uniform block of csx1
    cells through the casing, geometric growth rescaled so the first two blocks
    end at exactly 3*csx2, a uniform block of csx2 cells out to domain_x, then
    npadx cells of padding."""
    n_fine = int(np.ceil(casing_b / csx1 + 2))
    n_grow = int(np.floor(np.log(csx2 / csx1) / np.log(pfx1)))
    fine = meshTensor([(csx1, n_fine)])
    grow = meshTensor([(csx1, n_grow, pfx1)])
    inner_extent = 3                       # hard-coded "dx1 = 3" of mesh.py:53
    grow = grow * ((inner_extent * csx2 - fine.sum()) / grow.sum())
    n_coarse = int(np.ceil((domain_x - inner_extent) / csx2))
    coarse = meshTensor([(csx2, n_coarse)])
    pad = meshTensor([(csx2, npadx, pfx2)])
    return np.hstack([fine, grow, coarse, pad])


class BaseMeshGenerator(BaseCasing):
    """mesh.py:67-109"""
    _props = {"filename": "MeshParameters.json", "modelParameters": None}
    _instance_props = ("modelParameters",)

    @property
    def mesh(self):
        if getattr(self, "_mesh", None) is None:
            object.__setattr__(self, "_mesh", CylMesh([self.hx, self.hy, self.hz], x0=self.x0))
        return self._mesh

    def copy(self):
        cpy = super(BaseMeshGenerator, self).copy()
        cpy.modelParameters = self.modelParameters
        return cpy

    def serialize(self):
        out = super(BaseMeshGenerator, self).serialize()
        if self.__dict__.get("domain_z_user", None) is not None:
            out["domain_z"] = self.domain_z_user
        return out


class TensorMeshGenerator(BaseMeshGenerator):
    """mesh.py:112-323 -- Cartesian comparison meshes are outside the hot path
    (CylMesh only)."""

    def __init__(self, **kwargs):
        raise NotImplementedError("TensorMeshGenerator is out of scope of the B200 hot path (CylMesh only)")


class BaseCylMixin(object):
    """mesh.py:326-492 (plotting omitted)"""
    _props = {"csz": 25.0, "hy": np.r_[2 * np.pi], "nca": 5, "ncb": 5, "pfz": 1.5, "npadx": 23, "npadz": 38,
              "domain_x": 1000.0}
    _array_props = ("hy",)

    def _check_hy(self, value):
        if value is not None and abs(np.sum(value) - 2 * np.pi) >= 1e-6:
            raise AssertionError("hy must sum to 2*pi (mesh.py:408-411)")
        return value

    @property
    def x0(self):
        nbelow = self.npadz + self.ncz - self.nca
        return np.r_[0.0, 0.0, -np.sum(self.hz[:nbelow])]

    @property
    def domain_z(self):
        # an explicit user value (mesh.py domain_z.setter) wins and must survive the cache invalidation that every
        # public assignment triggers: it is kept under a public name, outside the '_'-prefixed derived quantities
        user = self.__dict__.get("domain_z_user", None)
        if user is not None:
            return user
        mp = self.modelParameters
        dz = np.absolute(mp.src_b[2] - mp.src_a[2])
        if getattr(mp, "casing_z", None) is not None:
            dz = max(np.absolute(mp.casing_z[1] - mp.casing_z[0]), dz)
        return dz

    @domain_z.setter
    def domain_z(self, value):
        object.__setattr__(self, "domain_z_user", None if value is None else float(value))

    @property
    def ncy(self):
        return len(self.hy)

    @property
    def ncz(self):
        return int(np.ceil(self.domain_z / self.csz)) + self.nca + self.ncb

    @property
    def hz(self):
        if getattr(self, "_hz", None) is None:
            object.__setattr__(self, "_hz", meshTensor([(self.csz, self.npadz, -self.pfz), (self.csz, self.ncz),
                                                        (self.csz, self.npadz, self.pfz)]))
        return self._hz

    def create_2D_mesh(self):
        mesh2D = self.copy()
        mesh2D.hy = np.r_[2 * np.pi]
        return mesh2D


class CylMeshGenerator(BaseCylMixin, BaseMeshGenerator):
    """Simple 3-D cylindrical mesh (mesh.py:495-538)."""
    _props = {"csx": 25.0, "pfx": 1.5, "nch": 10}

    @property
    def ncx(self):
        return int(np.ceil(self.domain_x / self.csx) + self.nch)

    @property
    def hx(self):
        if getattr(self, "_hx", None) is None:
            object.__setattr__(self, "_hx", meshTensor([(self.csx, self.ncx), (self.csx, self.npadx, self.pfx)]))
        return self._hx


class CasingMeshGenerator(BaseCylMixin, BaseMeshGenerator):
    """Mesh refined around the casing wall (mesh.py:541-582)."""
    _props = {"csx1": 2.5e-3, "csx2": 25.0, "pfx1": 1.3, "pfx2": 1.5}

    @property
    def hx(self):
        if getattr(self, "_hx", None) is None:
            object.__setattr__(self, "_hx", pad_for_casing_and_data(
                self.modelParameters.casing_b, self.csx1, self.csx2, self.pfx1, self.pfx2, self.domain_x, self.npadx))
        return self._hx
