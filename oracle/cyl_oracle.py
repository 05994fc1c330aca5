"""CPU ORACLE -- test infrastructure only, never the product path.

numpy + scipy restatement of the arithmetic that casingSimulations delegates
to un-vendored third-party packages (discretize.CylMesh, SimPEG problems,
pymatsolver / SimPEG.SolverLU).  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it.

PARITY UNPINNED: the reference ships no golden vectors and its dependencies
(SimPEG 0.14.x, discretize ~0.6, pymatsolver) are not installable here, so
this restatement is pinned only by (i) the reference's own self-consistency
tests (tests/test_cyl2D3D.py:202-359, re-run in tests/test_oracle.py),
(ii) discrete identities and (iii) closed-form solutions (DC dipole and the
complex field of a vertical electric dipole in a wholespace, tests/test_oracle.py).
Every [3P-recall]
convention below is anchored to the reference call site that triggers it.

Reference anchors
-----------------
mesh creation             mesh.py:87-99, 518, 563  (discretize.CylMesh([hx,hy,hz], x0))
meshTensor                mesh.py:46-61, 448-452, 534-537; model.py:50-53
vector layout             utils.py:28-87  ([Fx|Fy|Fz], [Ex|Ey|Ez], Fortran order)
DC problem                run.py:336-352  (Problem3D_CC, bc_type='Dirichlet')
FDEM problem              run.py:237-257  (Problem3D_{e,b,h,j}); tests/test_cyl2D3D.py:130-153
TDEM problem              run.py:284-301  (Problem3D_{e,b,h,j}, timeSteps, RawVec_Grounded)
solve                     run.py:17-24, 205 (Pardiso, fallback SolverLU == scipy splu)
post-processing           physics.py:10-83
"""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

MU_0 = 4e-7 * np.pi
EPSILON_0 = 8.8541878128e-12


# --------------------------------------------------------------------------
# discretize.utils.meshTensor  [3P-recall]; call sites mesh.py:46-61,448-452
# --------------------------------------------------------------------------
def mesh_tensor(spec):
    """(h, n) -> n cells of width h; (h, n, f): f>0 grows h*f**(1..n),
    f<0 the same widths reversed.  ndarray input is passed through."""
    if isinstance(spec, np.ndarray):
        return spec.astype(float)
    out = []
    for v in spec:
        if np.isscalar(v):
            out.append(np.r_[float(v)])
            continue
        if len(v) == 2:
            out.append(np.ones(int(v[1])) * float(v[0]))
        else:
            h, n, f = float(v[0]), int(v[1]), float(v[2])
            w = h * np.abs(f) ** (np.arange(n) + 1.0)
            out.append(w[::-1] if f < 0 else w)
    return np.hstack(out) if out else np.zeros(0)


def _sdiag(v):
    return sp.diags(np.asarray(v).ravel(), 0, format="csr")


def _ddx(n):
    """n x (n+1) forward difference."""
    return sp.diags([-np.ones(n), np.ones(n)], [0, 1], shape=(n, n + 1), format="csr")


def _ddx_periodic(n):
    """n x n difference d[j] = x[(j+1)%n] - x[j]."""
    D = sp.lil_matrix((n, n))
    for j in range(n):
        D[j, j] += -1.0
        D[j, (j + 1) % n] += 1.0
    return D.tocsr()


def _av(n):
    """n x (n+1) two-point average."""
    return sp.diags([0.5 * np.ones(n), 0.5 * np.ones(n)], [0, 1], shape=(n, n + 1), format="csr")


def _av_periodic(n):
    """n x n: cell j <- nodes j and (j+1)%n."""
    A = sp.lil_matrix((n, n))
    for j in range(n):
        A[j, j] += 0.5
        A[j, (j + 1) % n] += 0.5
    return A.tocsr()


def _kron3(a, b, c):
    return sp.kron(sp.kron(a, b), c, format="csr")


class OracleCylMesh(object):
    """discretize.CylMesh restated (all [3P-recall]).

    Index convention: i = r (fastest), j = theta, k = z; Fortran ordering;
    vectors are [x|y|z] concatenations (utils.py:41-47, 63-67).

    axis_weight_edge / axis_weight_face: radial averaging weight that the
    innermost cell gives to the Ey edge / Fx face at r_1 (A.4, open question 1).
    """

    def __init__(self, h, x0=None, axis_weight_edge=1.0, axis_weight_face=0.5, dtype=float):
        # dtype=np.longdouble assembles every measure, operator and inner product in extended precision from the same
        # fp64 inputs: the "exact" discrete system of exact_mesh() / the parity truth (see solve_refined)
        self.dtype = dtype
        self.hx = np.asarray(h[0], dtype=dtype).copy()
        self.hy = np.atleast_1d(np.asarray(h[1], dtype=dtype)).copy()
        self.hz = np.asarray(h[2], dtype=dtype).copy()
        self.x0 = np.zeros(3, dtype=dtype) if x0 is None else np.asarray(x0, dtype=dtype).copy()
        assert abs(self.hy.sum() - 2 * np.pi) < 1e-6
        self.nx, self.nth, self.nz = len(self.hx), len(self.hy), len(self.hz)
        self.isSymmetric = self.nth == 1
        self.axis_weight_edge = float(axis_weight_edge)
        self.axis_weight_face = float(axis_weight_face)
        nx, nth, nz = self.nx, self.nth, self.nz
        self.vnC = np.r_[nx, nth, nz]
        self.nC = nx * nth * nz
        if self.isSymmetric:
            self.vnF = np.r_[nx * nz, 0, nx * (nz + 1)]
            self.vnE = np.r_[0, nx * (nz + 1), 0]
            self.vnFx, self.vnFy, self.vnFz = (nx, 1, nz), (nx, 0, nz), (nx, 1, nz + 1)
            self.vnEx, self.vnEy, self.vnEz = (nx, 0, nz + 1), (nx, 1, nz + 1), None
            self.nN = 0
        else:
            self.vnF = np.r_[nx * nth * nz, nx * nth * nz, nx * nth * (nz + 1)]
            self.vnE = np.r_[nx * nth * (nz + 1), nx * nth * (nz + 1), nz * (nx * nth + 1)]
            self.vnFx, self.vnFy, self.vnFz = (nx, nth, nz), (nx, nth, nz), (nx, nth, nz + 1)
            self.vnEx, self.vnEy, self.vnEz = (nx, nth, nz + 1), (nx, nth, nz + 1), None
            self.nN = (nz + 1) * (nx * nth + 1)
        self.nFx, self.nFy, self.nFz = [int(v) for v in self.vnF]
        self.nEx, self.nEy, self.nEz = [int(v) for v in self.vnE]
        self.nF, self.nE = int(self.vnF.sum()), int(self.vnE.sum())
        # 1-D tables
        self.vectorNx_full = np.r_[0.0, np.cumsum(self.hx)]          # nx+1 incl. axis
        self.vectorNx = self.vectorNx_full[1:] if True else None
        self.vectorCCx = 0.5 * (self.vectorNx_full[1:] + self.vectorNx_full[:-1])
        self.vectorNy = np.r_[0.0, np.cumsum(self.hy)[:-1]]
        self.vectorCCy = np.r_[0.0] if self.isSymmetric else self.vectorNy + 0.5 * self.hy
        self.vectorNz = self.x0[2] + np.r_[0.0, np.cumsum(self.hz)]
        self.vectorCCz = 0.5 * (self.vectorNz[1:] + self.vectorNz[:-1])

    # ---- grids (Fortran order, i fastest) ---------------------------------
    @staticmethod
    def _ndgrid(x, y, z):
        X, Y, Z = np.meshgrid(x, y, z, indexing="ij")
        return np.c_[X.ravel(order="F"), Y.ravel(order="F"), Z.ravel(order="F")]

    @property
    def gridCC(self):
        return self._ndgrid(self.vectorCCx, self.vectorCCy, self.vectorCCz)

    @property
    def gridFx(self):
        return self._ndgrid(self.vectorNx, self.vectorCCy, self.vectorCCz)

    @property
    def gridFy(self):
        if self.isSymmetric:
            return np.zeros((0, 3))
        return self._ndgrid(self.vectorCCx, self.vectorNy, self.vectorCCz)

    @property
    def gridFz(self):
        return self._ndgrid(self.vectorCCx, self.vectorCCy, self.vectorNz)

    @property
    def gridEy(self):
        return self._ndgrid(self.vectorNx, self.vectorCCy, self.vectorNz)

    @property
    def gridEx(self):
        if self.isSymmetric:
            return np.zeros((0, 3))
        return self._ndgrid(self.vectorCCx, self.vectorNy, self.vectorNz)

    @property
    def gridEz(self):
        if self.isSymmetric:
            return np.zeros((0, 3))
        out = []
        reg = self._ndgrid(self.vectorNx, self.vectorNy, np.r_[0.0])
        for k in range(self.nz):
            out.append(np.r_[0.0, 0.0, self.vectorCCz[k]][None, :])
            blk = reg.copy()
            blk[:, 2] = self.vectorCCz[k]
            out.append(blk)
        return np.vstack(out)

    # ---- measures (A.2) ----------------------------------------------------
    def _rings(self):
        rn = self.vectorNx_full
        return 0.5 * (rn[1:] ** 2 - rn[:-1] ** 2)

    @property
    def vol(self):
        return np.kron(self.hz, np.kron(self.hy, self._rings()))

    @property
    def area(self):
        aFx = np.kron(self.hz, np.kron(self.hy, self.vectorNx))
        aFz = np.kron(np.ones(self.nz + 1), np.kron(self.hy, self._rings()))
        if self.isSymmetric:
            return np.r_[aFx, aFz]
        aFy = np.kron(self.hz, np.kron(np.ones(self.nth), self.hx))
        return np.r_[aFx, aFy, aFz]

    @property
    def edge(self):
        lEy = np.kron(np.ones(self.nz + 1), np.kron(self.hy, self.vectorNx))
        if self.isSymmetric:
            return lEy
        lEx = np.kron(np.ones(self.nz + 1), np.kron(np.ones(self.nth), self.hx))
        lEz = np.kron(self.hz, np.r_[1.0, np.ones(self.nx * self.nth)])
        return np.r_[lEx, lEy, lEz]

    # ---- Ez deflation helper ------------------------------------------------
    def _ez_full_to_deflated(self):
        """(nEz_deflated x nEz_full) selector/sum: full ordering is the
        (nx+1) x nth x nz Fortran array of edges at r nodes 0..nx, theta nodes
        0..nth-1; all axis entries (0, j, k) collapse onto one edge per k that
        comes first in its k-block (A.2, open question 2)."""
        nx, nth, nz = self.nx, self.nth, self.nz
        rows, cols = [], []
        for k in range(nz):
            base = k * (nx * nth + 1)
            for j in range(nth):
                for i in range(nx + 1):
                    full = i + (nx + 1) * (j + nth * k)
                    if i == 0:
                        rows.append(base)
                    else:
                        rows.append(base + 1 + (i - 1) + nx * j)
                    cols.append(full)
        P = sp.csr_matrix((np.ones(len(rows)), (rows, cols)),
                          shape=(nz * (nx * nth + 1), (nx + 1) * nth * nz))
        return P

    # ---- topological operators ----------------------------------------------
    def _Dtopo(self):
        nx, nth, nz = self.nx, self.nth, self.nz
        dr = _ddx(nx)[:, 1:]                      # drop the absent r=0 face
        Ix, It, Iz = sp.identity(nx), sp.identity(nth), sp.identity(nz)
        Dx = _kron3(Iz, It, dr)
        Dz = _kron3(_ddx(nz), It, Ix)
        if self.isSymmetric:
            return sp.hstack([Dx, Dz], format="csr")
        Dy = _kron3(Iz, _ddx_periodic(nth), Ix)
        return sp.hstack([Dx, Dy, Dz], format="csr")

    def _Ctopo(self):
        nx, nth, nz = self.nx, self.nth, self.nz
        Ix, It = sp.identity(nx), sp.identity(nth)
        Iz, Izn = sp.identity(nz), sp.identity(nz + 1)
        dz = _ddx(nz)                              # nz x (nz+1)
        if self.isSymmetric:
            # Ey -> [Fx; Fz]  (A.3)
            CFx = -_kron3(dz, It, Ix)
            drn = _ddx(nx)[:, 1:]                  # value(r_{i+1}) - value(r_i), r_0 dropped
            CFz = _kron3(Izn, It, drn)
            return sp.vstack([CFx, CFz], format="csr")
        dth = _ddx_periodic(nth)
        P = self._ez_full_to_deflated()            # deflated <- full
        # full Ez arrays: (nx+1, nth, nz)
        Ixn = sp.identity(nx + 1)
        sel_outer = sp.identity(nx + 1, format="csr")[1:, :]   # pick r_{i+1}
        dr_full = _ddx(nx)                                      # nx x (nx+1): E(r_{i+1}) - E(r_i)
        # Fx: +dtheta Ez(r_{i+1}) - dz Ey
        Fx_Ez = _kron3(Iz, dth, sel_outer) @ P.T
        Fx_Ey = -_kron3(dz, It, Ix)
        Fx_Ex = sp.csr_matrix((self.nFx, self.nEx))
        # Fy: +dz Ex - dr Ez
        Fy_Ex = _kron3(dz, It, Ix)
        Fy_Ez = -_kron3(Iz, It, dr_full) @ P.T
        Fy_Ey = sp.csr_matrix((self.nFy, self.nEy))
        # Fz: +dr Ey (r_0 term absent) - dtheta Ex
        drn = _ddx(nx)[:, 1:]
        Fz_Ey = _kron3(Izn, It, drn)
        Fz_Ex = -_kron3(Izn, dth, Ix)
        Fz_Ez = sp.csr_matrix((self.nFz, self.nEz))
        return sp.bmat([[Fx_Ex, Fx_Ey, Fx_Ez],
                        [Fy_Ex, Fy_Ey, Fy_Ez],
                        [Fz_Ex, Fz_Ey, Fz_Ez]], format="csr")

    def _Gtopo(self):
        """nodes -> edges (3-D only).  Node ordering mirrors Ez: per z-node
        level the axis node first, then (r_{i+1}, theta_j) i-fastest."""
        assert not self.isSymmetric
        nx, nth, nz = self.nx, self.nth, self.nz
        npl = nx * nth + 1

        def nid(i, j, k):          # i in 0..nx (0 = axis)
            return k * npl + (0 if i == 0 else 1 + (i - 1) + nx * (j % nth))
        rows, cols, vals = [], [], []
        e = 0
        for k in range(nz + 1):                    # Ex (i, jnode, knode): r_i -> r_{i+1}
            for j in range(nth):
                for i in range(nx):
                    rows += [e, e]; cols += [nid(i + 1, j, k), nid(i, j, k)]; vals += [1.0, -1.0]; e += 1
        for k in range(nz + 1):                    # Ey (r_{i+1}, jcell, knode)
            for j in range(nth):
                for i in range(nx):
                    rows += [e, e]; cols += [nid(i + 1, j + 1, k), nid(i + 1, j, k)]; vals += [1.0, -1.0]; e += 1
        for k in range(nz):                        # Ez: axis first then regular
            rows += [e, e]; cols += [nid(0, 0, k + 1), nid(0, 0, k)]; vals += [1.0, -1.0]; e += 1
            for j in range(nth):
                for i in range(nx):
                    rows += [e, e]; cols += [nid(i + 1, j, k + 1), nid(i + 1, j, k)]; vals += [1.0, -1.0]; e += 1
        return sp.csr_matrix((vals, (rows, cols)), shape=(self.nE, self.nN))

    # ---- differential operators (A.3) ---------------------------------------
    @property
    def faceDiv(self):
        return _sdiag(1.0 / self.vol) @ self._Dtopo() @ _sdiag(self.area)

    @property
    def edgeCurl(self):
        return _sdiag(1.0 / self.area) @ self._Ctopo() @ _sdiag(self.edge)

    @property
    def nodalGrad(self):
        return _sdiag(1.0 / self.edge) @ self._Gtopo()

    # ---- averaging / inner products (A.4) -------------------------------------
    def _avR(self, w_axis):
        """nx x nx: cell i <- radial faces/edges at r_{i+1} (column i) and r_i
        (column i-1); the r=0 column is dropped; entry [0,0] = w_axis."""
        A = _av(self.nx)[:, 1:].tolil()
        A[0, 0] = w_axis
        return A.tocsr()

    def face_average_transpose(self):
        """[AvFx^T; AvFy^T; AvFz^T] (nF x nC), each component a 2-point average."""
        nx, nth, nz = self.nx, self.nth, self.nz
        Ix, It, Iz = sp.identity(nx), sp.identity(nth), sp.identity(nz)
        AvFx = _kron3(Iz, It, self._avR(self.axis_weight_face))
        AvFz = _kron3(_av(nz), It, Ix)
        if self.isSymmetric:
            return sp.vstack([AvFx.T, AvFz.T], format="csr")
        AvFy = _kron3(Iz, _av_periodic(nth), Ix)
        return sp.vstack([AvFx.T, AvFy.T, AvFz.T], format="csr")

    def edge_average_transpose(self):
        """[AvEx^T; AvEy^T; AvEz^T] (nE x nC), each component a 4-point average."""
        nx, nth, nz = self.nx, self.nth, self.nz
        Ix, It = sp.identity(nx), sp.identity(nth)
        avz = _av(nz)
        AvEy = _kron3(avz, It, self._avR(self.axis_weight_edge))
        if self.isSymmetric:
            return AvEy.T.tocsr()
        avt = _av_periodic(nth)
        AvEx = _kron3(avz, avt, Ix)
        AvEz_full = _kron3(sp.identity(nz), avt, _av(nx))      # nC x full Ez
        AvEz = AvEz_full @ self._ez_full_to_deflated().T        # nC x deflated
        return sp.vstack([AvEx.T, AvEy.T, AvEz.T], format="csr")

    # ---- vector averages to cell centres (view.py:300-308: mesh.aveF2CCV, mesh.aveE2CCV) [3P-recall] ----
    @property
    def aveF2CCV(self):
        """block diagonal of the face -> cell averages of each component: (2|3) nC x nF"""
        nx, nth, nz = self.nx, self.nth, self.nz
        Ix, It, Iz = sp.identity(nx), sp.identity(nth), sp.identity(nz)
        AvFx = _kron3(Iz, It, self._avR(self.axis_weight_face))
        AvFz = _kron3(_av(nz), It, Ix)
        if self.isSymmetric:
            return sp.block_diag([AvFx, AvFz], format="csr")
        return sp.block_diag([AvFx, _kron3(Iz, _av_periodic(nth), Ix), AvFz], format="csr")

    @property
    def aveE2CCV(self):
        """block diagonal of the edge -> cell averages of each component: (1|3) nC x nE"""
        return sp.block_diag([b.T for b in self._edge_average_blocks()], format="csr")

    def _edge_average_blocks(self):
        full = self.edge_average_transpose()
        if self.isSymmetric:
            return [full]
        return [full[:self.nEx], full[self.nEx:self.nEx + self.nEy], full[self.nEx + self.nEy:]]

    def getFaceInnerProduct(self, prop=None, invProp=False, invMat=False):
        p = np.ones(self.nC, dtype=self.dtype) if prop is None else np.asarray(prop, dtype=self.dtype)
        if invProp:
            p = 1.0 / p
        d = self.face_average_transpose() @ (self.vol * p)
        return _sdiag(1.0 / d if invMat else d)

    def getEdgeInnerProduct(self, prop=None, invProp=False, invMat=False):
        p = np.ones(self.nC, dtype=self.dtype) if prop is None else np.asarray(prop, dtype=self.dtype)
        if invProp:
            p = 1.0 / p
        d = self.edge_average_transpose() @ (self.vol * p)
        return _sdiag(1.0 / d if invMat else d)


# --------------------------------------------------------------------------
# closestPoints  [3P-recall]; call sites sources.py:435,449, DC Dipole run.py:346
# --------------------------------------------------------------------------
def closest_points(grid, pts):
    pts = np.atleast_2d(np.asarray(pts, dtype=float))
    out = np.zeros(pts.shape[0], dtype=int)
    for n, p in enumerate(pts):
        out[n] = int(np.argmin(((grid - p[None, :]) ** 2).sum(axis=1)))
    return out


# --------------------------------------------------------------------------
# model painter restated: model.py:190-215 (background), 242-260 (air),
# 507-580 (casing then inside, strict inequalities)
# --------------------------------------------------------------------------
def paint_casing_in_halfspace(mesh, sigma_back=1e-2, sigma_air=1e-6, sigma_casing=5.5e6,
                              sigma_inside=1e-2, mur_back=1.0, mur_casing=1.0, mur_inside=1.0,
                              casing_a=0.045, casing_b=0.055, casing_z=(-1000.0, 0.0),
                              surface_z=0.0, halfspace=True):
    cc = mesh.gridCC
    r, z = cc[:, 0], cc[:, 2]
    sigma = sigma_back * np.ones(mesh.nC)
    mur = mur_back * np.ones(mesh.nC)
    if halfspace:
        sigma[z > surface_z] = sigma_air
    inz = (z > casing_z[0]) & (z < casing_z[1])
    cas = (r > casing_a) & (r < casing_b) & inz
    ins = (r < casing_a) & inz
    sigma[cas] = sigma_casing
    sigma[ins] = sigma_inside
    mur[cas] = mur_casing
    mur[ins] = mur_inside
    return sigma, MU_0 * mur


# --------------------------------------------------------------------------
# the solver the reference falls back to: SimPEG.SolverLU == scipy splu
# (run.py:19-24); A.8
# --------------------------------------------------------------------------
class SolverLU(object):
    def __init__(self, A, **kwargs):
        self.A = A.tocsc()
        self.lu = spla.splu(self.A)

    def __mul__(self, b):
        b = np.asarray(b)
        if b.ndim == 1:
            return self.lu.solve(b.astype(self.A.dtype))
        out = np.empty(b.shape, dtype=np.result_type(self.A.dtype, b.dtype))
        for c in range(b.shape[1]):
            out[:, c] = self.lu.solve(b[:, c].astype(self.A.dtype))
        return out

    solve = __mul__

    def clean(self):
        self.lu = None


def exact_mesh(mesh):
    """The same mesh with every measure / operator / inner product assembled in np.longdouble (eps 1e-19) from the same
    fp64 inputs.  Problems built on it (OracleDC / OracleFDEM / OracleTDEM with refine=True) give the solution of the
    EXACT discrete system.  Why it matters: an fp64-assembled matrix (the reference's included) carries entry roundings
    of 1e-16 relative to the LARGE conductances on its diagonal -- next to the 1e13 coefficient contrast of a casing
    model that is a spurious leak comparable to the true small couplings, and it moves the solution by 1e-6 (DC) to
    1e-4 (3-D curl-curl).  A matrix-free difference-form operator (the GPU path) does not have that error, so it is
    compared with THIS truth at a fixed tolerance; the fp64-assembled oracle's own distance from it is reported."""
    return OracleCylMesh([mesh.hx, mesh.hy, mesh.hz], mesh.x0, mesh.axis_weight_edge, mesh.axis_weight_face,
                         dtype=np.longdouble)


def solve_refined(A, b, iters=10, lu=None):
    """"Truth" for parity tests: SuperLU solution polished by iterative refinement
    with the residual b - A x accumulated in extended precision (np.longdouble).
    If A itself holds long-double entries (exact_mesh) they are used as they are: the result is then the solution of
    the exactly assembled system.
    The casing systems are so ill-conditioned (coefficient contrast 1e13, curl-curl
    null space) that a plain fp64 direct solve -- the reference's Pardiso / SolverLU
    included -- carries a relative error of 1e-7 (DC) up to 1e-4 (3-D FDEM j at 1 Hz);
    returns (x_refined, x_plain_splu)."""
    A = A.tocsr()
    A.sort_indices()
    b = np.asarray(b)
    cplx = np.iscomplexobj(A.data) or np.iscomplexobj(b)
    ld = np.clongdouble if cplx else np.longdouble
    wd = complex if cplx else float
    if lu is None:
        lu = spla.splu(A.tocsc().astype(wd))
    x0 = lu.solve(b.astype(wd))
    data = A.data.astype(ld)
    assert np.all(np.diff(A.indptr) > 0)
    x, bl = x0.astype(ld), b.astype(ld)
    global LAST_REFINE_RESIDUAL
    dsc = 1.0 / np.sqrt(np.abs(np.asarray(A.diagonal(), dtype=wd)))      # Jacobi scaling of the residual norm
    bn = float(np.linalg.norm(np.asarray(b, dtype=wd) * dsc))
    for _ in range(iters):
        r = bl - np.add.reduceat(data * x[A.indices], A.indptr[:-1])
        x = x + lu.solve(r.astype(wd)).astype(ld)
    r = bl - np.add.reduceat(data * x[A.indices], A.indptr[:-1])
    # extended-precision residual (Jacobi-scaled norm) of the returned vector: <= 1e-12 when the refinement converged (it diverges when
    # cond * eps >= 1, e.g. 3-D curl-curl at 0.1 Hz); callers that need a truth must check it
    LAST_REFINE_RESIDUAL = float(np.linalg.norm(np.asarray(r, dtype=wd) * dsc)) / bn if bn > 0 else 0.0
    return x.astype(wd), x0


LAST_REFINE_RESIDUAL = 0.0


def energy_rel(u, ref, weight):
    """relative error in the weighted (energy) norm sqrt(sum weight |.|^2): with weight = diag(MfRho) on face
    currents j it is the dissipated-power norm -- the norm in which the fp64 data of a casing problem determine the
    solution (current loops inside the steel wall cost almost no energy and are invisible in it)."""
    u, ref, w = np.asarray(u), np.asarray(ref), np.asarray(weight, dtype=float)
    return float(np.sqrt(np.sum(w * np.abs(np.asarray(u - ref, dtype=complex)) ** 2) /
                         np.sum(w * np.abs(np.asarray(ref, dtype=complex)) ** 2)))


def rhs_rounding_sensitivity(A_exact, b_exact, quantity):
    """How far one fp64 rounding of the right-hand side alone (relative 1.1e-16) moves quantity(x): the solution
    of the exactly assembled system for b_exact vs for fl64(b_exact).  On 3-D casing systems at low frequency the rounding
    error has a component along the curl-curl null space that the regularisation i w MeMu amplifies by 1/(w mu): plain
    L2 norms of h and j move by 1e-2 / 1e-5 at 1 Hz while the energy norm and the casing currents do not move
    (1e-10 / 1e-16).  No fp64 implementation -- the reference's included -- can agree with another one beyond this."""
    cd = np.clongdouble if np.iscomplexobj(b_exact) or np.iscomplexobj(A_exact.data) else np.longdouble
    wd = complex if cd is np.clongdouble else float
    x1 = solve_refined(A_exact, b_exact, iters=12)[0]
    x2 = solve_refined(A_exact, np.asarray(np.asarray(b_exact, dtype=wd), dtype=cd), iters=12)[0]
    q1, q2 = np.asarray(quantity(x1), dtype=wd), np.asarray(quantity(x2), dtype=wd)
    return float(np.linalg.norm(q2 - q1) / np.linalg.norm(q1))


# --------------------------------------------------------------------------
# DC  (A.5; Problem3D_CC, Dirichlet; run.py:339-352)
# --------------------------------------------------------------------------
class OracleDC(object):
    def __init__(self, mesh, sigma):
        self.mesh, self.sigma = mesh, np.asarray(sigma, dtype=getattr(mesh, "dtype", float))
        self.Dt = mesh._Dtopo() @ _sdiag(mesh.area)             # vol * faceDiv
        self.MfRho = mesh.getFaceInnerProduct(1.0 / self.sigma)
        self.MfRhoI = _sdiag(1.0 / self.MfRho.diagonal())

    def getA(self):
        return (self.Dt @ self.MfRhoI @ self.Dt.T).tocsr()

    def getRHS(self, src_a, src_b, current=1.0):
        src_a, src_b = np.atleast_2d(src_a), np.atleast_2d(src_b)
        q = np.zeros((self.mesh.nC, src_a.shape[0]))
        cc = self.mesh.gridCC
        for s in range(src_a.shape[0]):
            q[closest_points(cc, src_a[s])[0], s] += current
            q[closest_points(cc, src_b[s])[0], s] -= current
        return q

    def fields(self, src_a, src_b, refine=False):
        q = self.getRHS(src_a, src_b)
        if refine:
            A = self.getA()
            lu = spla.splu(A.tocsc().astype(float))
            return np.stack([solve_refined(A, q[:, s], lu=lu)[0] for s in range(q.shape[1])], axis=1)
        Ainv = SolverLU(self.getA())
        phi = Ainv * q
        return phi

    def j(self, phi):
        return self.MfRhoI @ (self.Dt.T @ phi)

    def e(self, phi):
        Mf = self.mesh.getFaceInnerProduct()
        return _sdiag(1.0 / Mf.diagonal()) @ (self.MfRho @ self.j(phi))

    def charge(self, phi):
        return EPSILON_0 * (self.mesh.vol[:, None] * (self.mesh.faceDiv @ self.e(phi)).reshape(self.mesh.nC, -1)).reshape(np.shape(phi))


# --------------------------------------------------------------------------
# FDEM  (A.6; e^{+i w t}; s_m = 0; s_e on faces, sources.py:119-123)
# --------------------------------------------------------------------------
class OracleFDEM(object):
    def __init__(self, mesh, sigma, mu, form="h"):
        self.mesh, self.form = mesh, form
        dt_ = getattr(mesh, "dtype", float)
        self.sigma, self.mu = np.asarray(sigma, dt_), np.asarray(mu, dt_)
        self.C = mesh.edgeCurl
        m = mesh
        self.MfRho = m.getFaceInnerProduct(1.0 / self.sigma)
        self.MeMu = m.getEdgeInnerProduct(self.mu)
        self.MfMui = m.getFaceInnerProduct(1.0 / self.mu)
        self.MeSigma = m.getEdgeInnerProduct(self.sigma)
        self.MeMuI = _sdiag(1.0 / self.MeMu.diagonal())
        self.MeSigmaI = _sdiag(1.0 / self.MeSigma.diagonal())
        self.Mf = m.getFaceInnerProduct()
        self.Me = m.getEdgeInnerProduct()

    def getA(self, freq):
        w = 2 * np.pi * freq
        C = self.C
        if self.form == "h":
            return (C.T @ self.MfRho @ C + 1j * w * self.MeMu).tocsr()
        if self.form == "e":
            return (C.T @ self.MfMui @ C + 1j * w * self.MeSigma).tocsr()
        if self.form == "j":
            n = self.mesh.nF
            return (self.MfRho @ (C @ self.MeMuI @ C.T @ self.MfRho + 1j * w * sp.identity(n))).tocsr()
        if self.form == "b":
            n = self.mesh.nF
            return (self.MfMui @ (C @ self.MeSigmaI @ C.T @ self.MfMui + 1j * w * sp.identity(n))).tocsr()
        raise ValueError(self.form)

    def getRHS(self, freq, s_e):
        """s_e: face current density (RawVec_e hands it over as-is for h/j,
        integrated with Me... only for edge-based e/b forms when given on edges;
        casingSimulations always supplies FACE vectors, sources.py:236)."""
        w = 2 * np.pi * freq
        s_e = np.asarray(s_e, dtype=complex)
        C = self.C
        if self.form == "h":
            return C.T @ (self.MfRho @ s_e)
        if self.form == "j":
            return self.MfRho @ (-1j * w * s_e)
        # E-B formulations: s_e is an EDGE vector (SimPEG RawVec_e, integrate=False)
        if self.form == "e":
            return -1j * w * s_e
        if self.form == "b":
            return self.MfMui @ (C @ (self.MeSigmaI @ s_e))
        raise ValueError(self.form)

    def solve(self, freq, s_e, refine=False):
        A = self.getA(freq)
        if refine:
            return solve_refined(A, self.getRHS(freq, s_e))[0]
        Ainv = SolverLU(A)
        u = Ainv * self.getRHS(freq, s_e)
        Ainv.clean()
        return u

    # derived fields (A.6 last column)
    def e_from_eb(self, freq, u, s_e):
        """E-B formulations: e on edges"""
        if self.form == "e":
            return u
        return self.MeSigmaI @ (self.C.T @ (self.MfMui @ u) - np.asarray(s_e, dtype=complex))

    def b_from_eb(self, freq, u, s_e):
        if self.form == "b":
            return u
        return -(self.C @ u) / (1j * 2 * np.pi * freq)

    def j_from_eb(self, freq, u, s_e):
        return _sdiag(1.0 / self.Me.diagonal()) @ (self.MeSigma @ self.e_from_eb(freq, u, s_e))

    def j_from(self, freq, u, s_e):
        if self.form == "h":
            return self.C @ u - np.asarray(s_e, dtype=complex)
        return u

    def h_from(self, freq, u, s_e):
        if self.form == "h":
            return u
        w = 2 * np.pi * freq
        return self.MeMuI @ (-(self.C.T @ (self.MfRho @ u))) / (1j * w)

    def e_from(self, freq, u, s_e):
        j = self.j_from(freq, u, s_e)
        return _sdiag(1.0 / self.Mf.diagonal()) @ (self.MfRho @ j)

    def b_from(self, freq, u, s_e):
        h = self.h_from(freq, u, s_e)
        return _sdiag(1.0 / self.Me.diagonal()) @ (self.MeMu @ h)


# --------------------------------------------------------------------------
# TDEM backward Euler (A.7; run.py:284-301; step-off RawVec_Grounded)
# --------------------------------------------------------------------------
class OracleTDEM(object):
    def __init__(self, mesh, sigma, mu, time_steps, form="j", ground_first_cell=True):
        self.mesh, self.form = mesh, form
        dt_ = getattr(mesh, "dtype", float)
        self.sigma, self.mu = np.asarray(sigma, dt_), np.asarray(mu, dt_)
        self.time_steps = mesh_tensor(time_steps) if isinstance(time_steps, list) else np.asarray(time_steps, float)
        self.C = mesh.edgeCurl
        self.MfRho = mesh.getFaceInnerProduct(1.0 / self.sigma)
        self.MfRhoI = _sdiag(1.0 / self.MfRho.diagonal())
        self.MeMuI = _sdiag(1.0 / mesh.getEdgeInnerProduct(self.mu).diagonal())
        self.Dt = mesh._Dtopo() @ _sdiag(mesh.area)
        self.ground_first_cell = ground_first_cell
        self.dt_threshold = 1e-8
        if form == "e":                              # E-B formulation on the edges (A.7), 3-D meshes
            self.MfMui = mesh.getFaceInnerProduct(1.0 / self.mu)
            self.MeSigma = mesh.getEdgeInnerProduct(self.sigma)
            self.G = mesh.nodalGrad

    def getAdc(self):
        A = (self.Dt @ self.MfRhoI @ self.Dt.T).tolil()
        if self.ground_first_cell:
            A[0, 0] = A[0, 0] + 1.0            # [3P-recall] SimPEG tdem getAdc
        return A.tocsr()

    def j_initial(self, s_e, refine=False):
        if refine:
            phi = solve_refined(self.getAdc(), self.Dt @ s_e)[0]
        else:
            phi = SolverLU(self.getAdc()) * (self.Dt @ s_e)
        return -(self.MfRhoI @ (self.Dt.T @ phi)), phi

    def getAdiag(self, dt):
        n = self.mesh.nF
        return (self.MfRho @ (self.C @ self.MeMuI @ self.C.T @ self.MfRho + sp.identity(n) / dt)).tocsr()

    # ---- e formulation (A.7): s_e is an EDGE vector --------------------------------------------------
    def getAdc_nodal(self):
        A = (self.G.T @ self.MeSigma @ self.G).tolil()
        A[0, 0] = A[0, 0] + 1.0                    # grounded at node 0 (the axis node of the lowest level), like getAdc
        return A.tocsr()

    def e_initial(self, s_e, refine=False):
        rhs = self.G.T @ s_e
        A = self.getAdc_nodal()
        phi = solve_refined(A, rhs)[0] if refine else SolverLU(A) * np.asarray(rhs, dtype=float)
        return -(self.G @ phi), phi

    def getAdiag_e(self, dt):
        return (self.C.T @ self.MfMui @ self.C + self.MeSigma / dt).tocsr()

    def fields_e(self, s_e, refine=False):
        """eSolution (nE x (nT+1)), step-off: (C^T MfMui C + MeSigma/dt) e^{n+1} = MeSigma e^n/dt - (s^{n+1} - s^n)/dt"""
        assert self.form == "e"
        s_e = np.asarray(s_e, dtype=getattr(self.mesh, "dtype", float))
        nT = len(self.time_steps)
        F = np.zeros((self.mesh.nE, nT + 1))
        F[:, 0] = np.asarray(self.e_initial(s_e, refine)[0], dtype=float)
        A, lu, dt_prev = None, None, None
        for t, dt in enumerate(self.time_steps):
            if lu is None or abs(dt - dt_prev) > self.dt_threshold:
                A = self.getAdiag_e(dt)
                lu = spla.splu(A.tocsc().astype(float)); dt_prev = dt
            rhs = self.MeSigma @ F[:, t] / dt + (s_e / dt if t == 0 else 0.0 * s_e)
            F[:, t + 1] = solve_refined(A, rhs, lu=lu)[0] if refine else lu.solve(np.asarray(rhs, dtype=float))
        return F

    def fields(self, s_e, refine=False):
        """returns jSolution (nF x (nT+1)); step-off waveform: s_e(t=0) = s_e,
        s_e(t>0) = 0."""
        if self.form == "e":
            return self.fields_e(s_e, refine)
        assert self.form == "j"
        s_e = np.asarray(s_e, dtype=float)
        nT = len(self.time_steps)
        F = np.zeros((self.mesh.nF, nT + 1))
        F[:, 0], _ = self.j_initial(s_e, refine)
        A, lu, dt_prev = None, None, None
        for t, dt in enumerate(self.time_steps):
            if lu is None or abs(dt - dt_prev) > self.dt_threshold:
                A = self.getAdiag(dt)
                lu = spla.splu(A.tocsc().astype(float)); dt_prev = dt
            se_n1 = 0.0 * s_e
            se_n = s_e if t == 0 else 0.0 * s_e
            rhs = self.MfRho @ (-(se_n1 - se_n) / dt)
            Asub_x = -(self.MfRho @ F[:, t]) / dt
            if refine:
                F[:, t + 1] = solve_refined(A, rhs - Asub_x, lu=lu)[0]
            else:
                F[:, t + 1] = lu.solve(np.asarray(rhs - Asub_x, dtype=float))
        return F


# --------------------------------------------------------------------------
# physics.casing_currents / casing_charges restated (physics.py:10-83)
# --------------------------------------------------------------------------
def casing_currents(j, mesh, casing_a, casing_b, casing_z):
    j = np.asarray(j).reshape(mesh.nF, -1)
    gFx, gFz = mesh.gridFx, mesh.gridFz
    fx = (gFx[:, 0] >= casing_a) & (gFx[:, 0] <= casing_b) & (gFx[:, 2] <= casing_z[1]) & (gFx[:, 2] >= casing_z[0])
    fz = (gFz[:, 0] >= casing_a) & (gFz[:, 0] <= casing_b) & (gFz[:, 2] <= casing_z[1]) & (gFz[:, 2] >= casing_z[0])
    jA = mesh.area[:, None] * j
    jx = (jA[:mesh.nFx, 0] * fx).reshape(mesh.vnFx, order="F")
    jz = (jA[mesh.nFx + mesh.nFy:, 0] * fz).reshape(mesh.vnFz, order="F")
    ix, iz = jx.sum(0).sum(0), jz.sum(0).sum(0)
    ix_inds = (mesh.vectorCCz > casing_z[0]) & (mesh.vectorCCz < casing_z[1])
    iz_inds = (mesh.vectorNz > casing_z[0]) & (mesh.vectorNz < casing_z[1])
    return {"x": (mesh.vectorCCz[ix_inds], ix[ix_inds]), "z": (mesh.vectorNz[iz_inds], iz[iz_inds])}


def casing_charges(charge, mesh, casing_b, casing_z):
    cc = mesh.gridCC
    inds = ((cc[:, 0] >= -casing_b - mesh.hx.min()) & (cc[:, 0] <= casing_b + mesh.hx.min()) &
            (cc[:, 2] < casing_z[1]) & (cc[:, 2] > casing_z[0]))
    c = np.array(charge, dtype=float).ravel().copy()
    c[~inds] = 0.0
    c = c.reshape(tuple(mesh.vnC), order="F").sum(0).sum(0)
    zi = (mesh.vectorCCz > casing_z[0]) & (mesh.vectorCCz < casing_z[1])
    return mesh.vectorCCz[zi], c[zi]
