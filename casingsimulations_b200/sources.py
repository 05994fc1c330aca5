"""Source geometries (kept API of casingSimulations/sources.py).

Every source paints +-1 on a handful of mesh faces (wire segments and
electrodes) and divides by the face area, which yields the face current density
``s_e`` of a 1 A line current (sources.py:219-239, 364-379, 565-583, 693-712,
817-836).  The masks are host-side numpy on the 1-D grids (microseconds); the
vector is uploaded once and becomes the right-hand side on the GPU.
"""
import numpy as np

from .base import BaseCasing, memo_property
from .mesh import closestPoints


class RawVec_e(object):
    """Stand-in for SimPEG fdem.sources.RawVec_e([], freq, s_e) (sources.py:121)."""

    def __init__(self, rxList, freq, s_e):
        self.rxList, self.freq, self.s_e = rxList, float(freq), np.asarray(s_e)

    frequency = property(lambda self: self.freq)


class RawVec_Grounded(object):
    """Stand-in for SimPEG tdem.sources.RawVec_Grounded([], s_e) (sources.py:125):
    step-off waveform, DC initial state."""

    def __init__(self, rxList, s_e):
        self.rxList, self.s_e = rxList, np.asarray(s_e)


class BaseCasingSrc(BaseCasing):
    """sources.py:19-127"""
    _props = {"filename": "Source.json", "modelParameters": None, "meshGenerator": None, "physics": None,
              "src_a": None, "src_b": None}
    _array_props = ("src_a", "src_b")
    _instance_props = ("modelParameters", "meshGenerator")

    def __init__(self, **kwargs):
        self._init_props(**kwargs)
        if self.src_a is None:
            self.src_a = self.modelParameters.src_a
        if self.src_b is None:
            self.src_b = self.modelParameters.src_b
        assert self.src_a[1] == self.src_b[1], "non y-axis aligned sources have not been implemented"

    def _check_physics(self, v):
        if v is not None and v.lower() not in ("fdem", "tdem"):
            raise ValueError("physics must be 'fdem' or 'tdem'")
        return v

    @property
    def mesh(self):
        return self.meshGenerator.mesh

    @property
    def casing_a(self):
        return self.modelParameters.casing_a

    @property
    def freqs(self):
        return self.modelParameters.freqs

    @property
    def srcList(self):
        if getattr(self, "_srcList", None) is None:
            if self.physics.lower() == "fdem":
                lst = [RawVec_e([], f, self.s_e.astype("complex")) for f in self.freqs]
                for src in lst:
                    # the sources of one casing source differ in frequency only: the solver groups them by this tag and
                    # uploads the shared REAL vector once (comparing / de-interleaving 233 MB complex copies costs 0.1 s)
                    src._shared_s_e = self.s_e
            elif self.physics == "tdem":
                lst = [RawVec_Grounded([], self.s_e)]
            else:
                raise ValueError("physics must be fdem or tdem")
            object.__setattr__(self, "_srcList", lst)
        return self._srcList

    # helpers shared by the wire geometries ---------------------------------------
    def _theta_window(self, grid, theta, half_width):
        """open interval test on the azimuth, skipped on a symmetric mesh"""
        if getattr(self.mesh, "isSymmetric", False):
            return np.ones(grid.shape[0], dtype=bool)
        return (grid[:, 1] > theta - half_width) & (grid[:, 1] < theta + half_width)

    def _assemble(self, s_x, s_z):
        mesh = self.mesh
        s_y = np.zeros(mesh.vnF[1])
        return np.hstack([s_x, s_y, s_z]) / mesh.area

    def _assert_single(self, grid, mask, cols, what):
        # checked once per parameter set: the record lives among the '_'-prefixed derived quantities that
        # BaseCasing._invalidate drops when a public parameter is assigned (a boolean take on a 4.8 M-row grid is 12 ms,
        # six of them per validate())
        done = self.__dict__.get("_memo_assert_single")
        if done is None:
            done = set()
            object.__setattr__(self, "_memo_assert_single", done)
        if (what, cols) in done:
            return
        idx = np.flatnonzero(mask)
        names = {0: "x", 1: "y", 2: "z"}
        for c in cols:
            assert len(np.unique(grid[idx, c])) == 1, "the {} has more than one {}-location".format(what, names[c])
        done.add((what, cols))


class HorizontalElectricDipole(BaseCasingSrc):
    """sources.py:130-272"""

    def __init__(self, **kwargs):
        super(HorizontalElectricDipole, self).__init__(**kwargs)
        assert self.src_a[2] == self.src_b[2], "z locations must be the same for a HED"

    @memo_property
    def src_a_closest(self):
        # the reference indexes gridFx with an 'Fz' closest-point index (sources.py:149)
        return self.mesh.gridFx[closestPoints(self.mesh, self.src_a, "Fz"), :][0]

    @memo_property
    def src_b_closest(self):
        return self.mesh.gridFx[closestPoints(self.mesh, self.src_b, "Fz"), :][0]

    @memo_property
    def surface_wire(self):
        mesh, g = self.mesh, self.mesh.gridFx
        lo, hi = min(self.src_a[0], self.src_b[0]), max(self.src_a[0], self.src_b[0])
        dz = mesh.hz.min() / 2.0
        mask = (g[:, 0] <= hi) & (g[:, 0] >= lo) & (g[:, 2] > self.src_b[2] - dz) & (g[:, 2] < self.src_b[2] + dz)
        return mask & self._theta_window(g, self.src_b[1], mesh.hy.min() / 2.0)

    @property
    def surface_wire_direction(self):
        return -1.0 if self.src_a[0] < self.src_b[0] else 1.0

    @memo_property
    def s_e(self):
        mesh = self.mesh
        s_x, s_z = np.zeros(mesh.vnF[0]), np.zeros(mesh.vnF[2])
        s_x[self.surface_wire] = self.surface_wire_direction
        return self._assemble(s_x, s_z)

    def _validate_wire(self):
        self._assert_single(self.mesh.gridFx, self.surface_wire, (1, 2), "surface wire")


class VerticalElectricDipole(BaseCasingSrc):
    """sources.py:275-413"""

    def __init__(self, **kwargs):
        super(VerticalElectricDipole, self).__init__(**kwargs)
        assert all(self.src_a[:2] == self.src_b[:2]), "src_a and src_b must have the same horizontal location"

    @memo_property
    def src_a_closest(self):
        return self.mesh.gridFz[closestPoints(self.mesh, self.src_a, "Fz"), :][0]

    @memo_property
    def src_b_closest(self):
        return self.mesh.gridFz[closestPoints(self.mesh, self.src_b, "Fz"), :][0]

    @memo_property
    def wire_in_borehole(self):
        mesh, g = self.mesh, self.mesh.gridFz
        dx, dz = mesh.hx.min() / 2.0, 0.5 * mesh.hz.min()
        xa = self.src_a_closest[0]
        zlo, zhi = min(self.src_a[2], self.src_b[2]), max(self.src_a[2], self.src_b[2])
        mask = (g[:, 0] < xa + dx) & (g[:, 0] > xa - dx) & (g[:, 2] >= zlo - dz) & (g[:, 2] <= zhi + dz)
        return mask & self._theta_window(g, self.src_a[1], mesh.hy.min() / 2.0)

    @memo_property
    def _wire_direction(self):
        # the reference reads an attribute it never defines (sources.py:374); current
        # flows from src_a to src_b, matching the other sources' sign convention
        return -1.0 if self.src_a[2] > self.src_b[2] else 1.0

    @memo_property
    def s_e(self):
        mesh = self.mesh
        s_x, s_z = np.zeros(mesh.vnF[0]), np.zeros(mesh.vnF[2])
        s_z[self.wire_in_borehole] = self._wire_direction
        return self._assemble(s_x, s_z)

    def _validate_wire(self):
        self._assert_single(self.mesh.gridFz, self.wire_in_borehole, (0, 1), "wire in borehole")


class DownHoleTerminatingSrc(BaseCasingSrc):
    """A wire down the borehole axis that ends down-hole, a surface wire and a
    return electrode (sources.py:416-643)."""

    @memo_property
    def src_a_closest(self):
        return self.mesh.gridFz[closestPoints(self.mesh, self.src_a, "Fz"), :][0]

    @memo_property
    def src_b_closest(self):
        return self.mesh.gridFz[closestPoints(self.mesh, self.src_b, "Fz"), :][0]

    @memo_property
    def wire_in_borehole(self):
        mesh, g = self.mesh, self.mesh.gridFz
        dx, xa = mesh.hx.min() / 2.0, self.src_a_closest[0]
        mask = ((g[:, 0] < xa + dx) & (g[:, 0] > xa - dx) &
                (g[:, 2] >= self.src_a[2] - 0.5 * mesh.hz.min()) & (g[:, 2] < self.src_b[2] + 1.5 * mesh.hz.min()))
        return mask & self._theta_window(g, self.src_a[1], mesh.hy.min() / 2.0)

    @memo_property
    def surface_wire(self):
        mesh, g = self.mesh, self.mesh.gridFx
        xa, xb = self.src_a_closest[0], self.src_b_closest[0]
        mask = ((g[:, 0] <= max(xa, xb)) & (g[:, 0] >= min(xa, xb)) &
                (g[:, 2] > mesh.hz.min()) & (g[:, 2] <= 2 * mesh.hz.min()))
        return mask & self._theta_window(g, self.src_b[1], mesh.hy.min() / 2.0)

    @memo_property
    def surface_electrode(self):
        mesh, g = self.mesh, self.mesh.gridFz
        dx, xb = mesh.hx.min() / 2.0, self.src_b_closest[0]
        mask = ((g[:, 0] > xb - dx) & (g[:, 0] < xb + dx) &
                (g[:, 2] >= self.src_b_closest[2] - mesh.hz.min()) & (g[:, 2] < 2 * mesh.hz.min()))
        return mask & self._theta_window(g, self.src_b[1], mesh.hy.min() / 2.0)

    @property
    def surface_wire_direction(self):
        return -1.0 if self.src_a[0] < self.src_b[0] else 1.0

    @memo_property
    def s_e(self):
        mesh = self.mesh
        s_x, s_z = np.zeros(mesh.vnF[0]), np.zeros(mesh.vnF[2])
        s_z[self.wire_in_borehole] = -1.0
        s_x[self.surface_wire] = self.surface_wire_direction
        s_z[self.surface_electrode] = 1.0
        return self._assemble(s_x, s_z)

    def _validate_wire(self):
        self._assert_single(self.mesh.gridFz, self.surface_electrode, (0, 1), "surface electrode")
        self._assert_single(self.mesh.gridFx, self.surface_wire, (1, 2), "surface wire")
        self._assert_single(self.mesh.gridFz, self.wire_in_borehole, (0, 1), "wire in borehole")


class DownHoleCasingSrc(DownHoleTerminatingSrc):
    """Down-hole electrode coupled to the casing wall (sources.py:647-748)."""

    @memo_property
    def downhole_electrode(self):
        mesh, g = self.mesh, self.mesh.gridFx
        za = self.src_a_closest[2]
        mask = (g[:, 0] <= self.casing_a) & (g[:, 2] <= za) & (g[:, 2] > za - mesh.hz.min())
        return mask & self._theta_window(g, self.src_a[1], mesh.hy.min() / 2.0)

    @memo_property
    def s_e(self):
        mesh = self.mesh
        s_x, s_z = np.zeros(mesh.vnF[0]), np.zeros(mesh.vnF[2])
        s_z[self.wire_in_borehole] = -1.0
        s_x[self.downhole_electrode] = 1.0
        s_x[self.surface_wire] = -1.0
        s_z[self.surface_electrode] = 1.0
        return self._assemble(s_x, s_z)

    def _validate_wire_more(self):
        self._assert_single(self.mesh.gridFx, self.downhole_electrode, (1, 2), "downhole electrode")


class SurfaceGroundedSrc(DownHoleTerminatingSrc):
    """Two surface electrodes joined by a surface wire (sources.py:751-897)."""

    @memo_property
    def positive_electrode(self):
        mesh, g = self.mesh, self.mesh.gridFz
        a = self.src_a_closest
        mask = (g[:, 0] == a[0]) & (g[:, 2] >= a[2]) & (g[:, 2] < 1.5 * mesh.hz.min())
        return mask & self._theta_window(g, self.src_a[1], mesh.hy.min())

    @memo_property
    def s_e(self):
        mesh = self.mesh
        s_x, s_z = np.zeros(mesh.vnF[0]), np.zeros(mesh.vnF[2])
        s_z[self.positive_electrode] = -1.0
        s_x[self.surface_wire] = self.surface_wire_direction
        s_z[self.surface_electrode] = 1.0
        return self._assemble(s_x, s_z)

    def _validate_wire(self):
        self._assert_single(self.mesh.gridFz, self.surface_electrode, (0, 1), "surface electrode")
        self._assert_single(self.mesh.gridFz, self.positive_electrode, (0, 1), "tophole electrode")
        self._assert_single(self.mesh.gridFx, self.surface_wire, (1, 2), "surface wire")


class TopCasingSrc(SurfaceGroundedSrc):
    """Positive electrode on the top of the casing wall (sources.py:900-910); the
    reference moves src_a[0] onto the casing in place, shared array included."""

    def __init__(self, **kwargs):
        super(TopCasingSrc, self).__init__(**kwargs)
        self.src_a[0] = self.casing_a + self.mesh.hx.min() / 2.0
        self._invalidate("src_a")


class SourceList(BaseCasing):
    """sources.py:948-973"""
    _props = {"filename": "SourceList.json", "sources": None}

    @property
    def srcList(self):
        if getattr(self, "_srcList", None) is None:
            lst = []
            for src in self.sources:
                lst += src.srcList
            object.__setattr__(self, "_srcList", lst)
        return self._srcList
