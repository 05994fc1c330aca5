"""SimPEG / pymatsolver solver plugin protocol: ``Ainv = Solver(A, **opts)``,
``x = Ainv * rhs``, ``Ainv.clean()`` (injected at run.py:246, 294, 343;
tests/test_cyl2D3D.py:134, 140).

``A`` is an assembled scipy.sparse matrix of one of the cylindrical-mesh systems
(cell, edge or face unknowns in the reference layout).  The solver needs the
mesh to know the (r, theta, z) structure: pass ``mesh=`` (any object with hx, hy,
hz, x0) in the solver options.  The matrix is uploaded in CSR form, the
theta-mode band matrices are probed through the GPU SpMV, factored and used as
the preconditioner of the same Krylov driver as the matrix-free path.  There is
no CPU fallback: without the mesh hint, or if the system does not fit one of the
three layouts, construction raises.
"""
import numpy as np

from . import _lib


class Solver(object):
    def __init__(self, A, mesh=None, rtol=1e-10, maxit=50, device=0, check_accuracy=False, accuracy_tol=1e-6, **kwargs):
        import scipy.sparse as sp
        if mesh is None:
            raise _lib.CasingB200Error(
                "casingsimulations_b200.Solver needs the cylindrical mesh: Solver(A, mesh=mesh). "
                "There is no generic sparse fallback (and no CPU fallback) by design.")
        if not sp.issparse(A):
            raise TypeError("A must be a scipy.sparse matrix")
        self.A = A.tocsr()
        self.A.sort_indices()
        self.mesh, self.rtol, self.maxit = mesh, rtol, maxit
        self.check_accuracy, self.accuracy_tol = check_accuracy, accuracy_tol
        n = self.A.shape[0]
        assert self.A.shape[0] == self.A.shape[1], "A must be square"
        self.ctx = _lib.Context(device)
        self.ctx.set_mesh(mesh.hx, mesh.hy, mesh.hz, mesh.x0[2])
        if n == self.ctx.nC:
            family = 0
        elif n == self.ctx.nE:
            family = 1
        elif n == self.ctx.nF:
            family = 2
        else:
            self.ctx.close()
            raise _lib.CasingB200Error("matrix size {} matches neither nC, nE nor nF of the mesh".format(n))
        self.is_complex = np.iscomplexobj(self.A.data)
        self.op = self.ctx.op_from_csr(self.A, family)
        self._factored = False

    def factor(self):
        if not self._factored:
            self.op.factor([0.0])
            self._factored = True

    def _solve(self, rhs):
        if not isinstance(rhs, np.ndarray):
            raise TypeError("Can only multiply by a numpy array.")
        n = self.A.shape[0]
        if rhs.shape[0] != n:
            raise AssertionError("rhs has the wrong size")
        self.factor()
        b = rhs.reshape(n, -1)
        cplx = self.is_complex or np.iscomplexobj(b)
        if cplx and not self.is_complex:
            # real matrix, complex right-hand side: two real solves
            xr = self._solve(np.ascontiguousarray(b.real))
            xi = self._solve(np.ascontiguousarray(b.imag))
            return (xr.reshape(n, -1) + 1j * xi.reshape(n, -1)).reshape(rhs.shape)
        bd = self.ctx.to_device(np.ascontiguousarray(b.T), complex_=self.is_complex)
        x = self.op.solve(bd, rtol=self.rtol, maxit=self.maxit).cpu().numpy().T
        self.info = self.op.info
        return np.ascontiguousarray(x).reshape(rhs.shape)

    def __mul__(self, rhs):
        return self._solve(rhs)

    solve = __mul__

    def clean(self):
        if getattr(self, "ctx", None) is not None:
            self.ctx.close()
            self.ctx = None

    def __del__(self):
        try:
            self.clean()
        except Exception:
            pass
