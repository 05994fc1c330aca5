"""ctypes binding of libcasing_b200.so (C ABI in include/casing_b200.h).

PyTorch is used only to own device buffers; every computation happens in the
hand-written sm_100a kernels behind the C ABI.  There is NO CPU fallback: if
the shared library is missing or no CUDA device is present, calls raise.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcasing_b200.so")

KIND_DC, KIND_DC_GROUNDED, KIND_EDGE_H, KIND_EDGE_E, KIND_FACE_J, KIND_FACE_B, KIND_NODAL_E = range(7)
INFO_LEN = 16
INFO_NAMES = ["iters", "relres", "ms_assemble", "ms_factor", "ms_solve", "n_apply", "n_launch", "bytes", "converged", "backward_err",
              "dd_steps", "dd_corr"]

_lib = None


class CasingB200Error(RuntimeError):
    pass


def build_library(verbose=False):
    """Compile the CUDA sources in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    import subprocess
    script = os.path.join(_HERE, "csrc", "build.sh")
    res = subprocess.run(["bash", script], capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout)
        print(res.stderr)
    if res.returncode != 0:
        raise CasingB200Error("nvcc build of libcasing_b200.so failed:\n" + res.stderr[-4000:])
    return LIB_PATH


def load_library():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise CasingB200Error(
            "libcasing_b200.so is not built (run `python -c 'import __graft_entry__ as g; g.build()'`); "
            "this package has no CPU fallback")
    lib = ctypes.CDLL(LIB_PATH)
    c = ctypes
    vp, dp, ip, ll = c.c_void_p, c.POINTER(c.c_double), c.POINTER(c.c_int), c.c_longlong
    sig = {
        "cs_last_error": (c.c_char_p, []),
        "cs_launch_count": (ll, []),
        "cs_ctx_create": (c.c_int, [c.c_int, vp, c.POINTER(vp)]),
        "cs_ctx_destroy": (c.c_int, [vp]),
        "cs_ctx_sync": (c.c_int, [vp]),
        "cs_fp64_peak": (c.c_int, [vp, dp]),
        "cs_ctx_set_refine": (c.c_int, [vp, c.c_int]),
        "cs_comm_unique_id": (c.c_int, [vp]),
        "cs_comm_init": (c.c_int, [vp, vp, c.c_int, c.c_int, c.c_int, c.c_int]),
        "cs_comm_local_range": (c.c_int, [vp, ip]),
        "cs_comm_halo": (c.c_int, [vp, c.c_int, vp, c.c_int, c.c_int]),
        "cs_mesh_set": (c.c_int, [vp, c.c_int, c.c_int, c.c_int, dp, dp, dp, c.c_double, c.c_double, c.c_double]),
        "cs_mesh_counts": (c.c_int, [vp, c.POINTER(ll)]),
        "cs_model_paint": (c.c_int, [vp, dp, vp, vp]),
        "cs_model_set": (c.c_int, [vp, vp, vp]),
        "cs_face_inner_product": (c.c_int, [vp, vp, c.c_int, c.c_int, vp]),
        "cs_edge_inner_product": (c.c_int, [vp, vp, c.c_int, c.c_int, vp]),
        "cs_curl": (c.c_int, [vp, vp, vp, c.c_int]),
        "cs_curlT": (c.c_int, [vp, vp, vp, c.c_int]),
        "cs_div": (c.c_int, [vp, vp, vp, c.c_int]),
        "cs_divT": (c.c_int, [vp, vp, vp, c.c_int]),
        "cs_grad": (c.c_int, [vp, vp, vp, c.c_int]),
        "cs_gradT": (c.c_int, [vp, vp, vp, c.c_int]),
        "cs_ave_f2ccv": (c.c_int, [vp, vp, vp, c.c_int]),
        "cs_ave_e2ccv": (c.c_int, [vp, vp, vp, c.c_int]),
        "cs_op_get": (c.c_int, [vp, c.c_int, c.POINTER(vp)]),
        "cs_op_get_noplan": (c.c_int, [vp, c.c_int, c.POINTER(vp)]),
        "cs_op_apply_timed": (c.c_int, [vp, dp, vp, vp, c.c_int, c.c_int, c.c_int, dp]),
        "cs_op_from_csr": (c.c_int, [vp, c.c_int, ll, ll, c.POINTER(ll), ip, vp, c.c_int, c.POINTER(vp)]),
        "cs_op_size": (ll, [vp]),
        "cs_op_apply": (c.c_int, [vp, dp, vp, vp, c.c_int, c.c_int]),
        "cs_op_factor": (c.c_int, [vp, dp, c.c_int, c.c_int]),
        "cs_op_solve": (c.c_int, [vp, vp, vp, c.c_int, ip, c.c_int, c.c_double, c.c_int, dp]),
        "cs_op_clean": (c.c_int, [vp]),
        "cs_solve_dc": (c.c_int, [vp, vp, vp, c.c_int, c.c_double, c.c_int, dp]),
        "cs_solve_fdem": (c.c_int, [vp, c.c_int, dp, c.c_int, vp, vp, c.c_double, c.c_int, dp]),
        "cs_tdem_run": (c.c_int, [vp, c.c_int, dp, c.c_int, vp, vp, c.c_double, c.c_int, dp]),
        "cs_tdem_run_multi": (c.c_int, [vp, c.c_int, dp, c.c_int, vp, c.c_int, vp, c.c_double, c.c_int, dp]),
        "cs_fields_fdem": (c.c_int, [vp, c.c_int, c.c_int, c.c_double, vp, vp, vp]),
        "cs_fields_dc": (c.c_int, [vp, c.c_int, vp, vp, c.c_int]),
        "cs_casing_currents": (c.c_int, [vp, vp, c.c_int, c.c_double, c.c_double, c.c_double, c.c_double, vp, vp]),
        "cs_casing_charges": (c.c_int, [vp, vp, c.c_double, c.c_double, c.c_double, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


EXPORTED_SYMBOLS = [
    "cs_last_error", "cs_launch_count", "cs_ctx_create", "cs_ctx_destroy", "cs_ctx_sync", "cs_fp64_peak", "cs_ctx_set_refine",
    "cs_comm_unique_id", "cs_comm_init", "cs_comm_local_range", "cs_comm_halo", "cs_mesh_set",
    "cs_mesh_counts", "cs_model_paint", "cs_model_set", "cs_face_inner_product", "cs_edge_inner_product",
    "cs_curl", "cs_curlT", "cs_div", "cs_divT", "cs_grad", "cs_gradT", "cs_ave_f2ccv", "cs_ave_e2ccv", "cs_op_get", "cs_op_get_noplan", "cs_op_apply_timed", "cs_op_from_csr", "cs_op_size", "cs_op_apply", "cs_op_factor",
    "cs_op_solve", "cs_op_clean", "cs_solve_dc", "cs_solve_fdem", "cs_tdem_run", "cs_tdem_run_multi", "cs_fields_fdem",
    "cs_fields_dc", "cs_casing_currents", "cs_casing_charges",
]


def check(rc):
    if rc != 0:
        raise CasingB200Error(load_library().cs_last_error().decode("utf-8", "replace"))


def _dptr(a):
    return np.ascontiguousarray(a, dtype=np.float64).ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise CasingB200Error("no CUDA device: casingsimulations_b200 has no CPU fallback")
    return torch


def _tptr(t):
    return ctypes.c_void_p(t.data_ptr())


class Op(object):
    """One assembled operator (weights + theta-mode band matrices) of a Context."""

    def __init__(self, ctx, kind, handle=None):
        self.ctx, self.kind = ctx, kind
        if handle is None:
            handle = ctypes.c_void_p()
            ctx._enter()
            check(ctx.lib.cs_op_get(ctx.h, kind, ctypes.byref(handle)))
            ctx._exit()
        self.h = handle
        self.n = int(ctx.lib.cs_op_size(handle))

    def apply(self, x, shifts=None):
        """y = A(s) x; x torch tensor (nvec, n) or (n,), float64 or complex128."""
        torch = _torch()
        x2 = x.reshape(-1, self.n).contiguous()
        nvec = x2.shape[0]
        y = torch.empty_like(x2)
        sh = np.zeros(2 * nvec)
        if shifts is not None:
            s = np.atleast_1d(np.asarray(shifts, dtype=complex))
            s = np.broadcast_to(s, (nvec,))
            sh[0::2], sh[1::2] = s.real, s.imag
        self.ctx._enter()
        check(self.ctx.lib.cs_op_apply(self.h, _dptr(sh), _tptr(x2), _tptr(y), nvec, int(x2.is_complex())))
        self.ctx._exit()
        return y.reshape(x.shape)

    def apply_timed(self, x, shift=0.0, reps=20):
        """average device time (ms) of one apply launch, CUDA events on the library stream"""
        torch = _torch()
        x2 = x.reshape(-1, self.n).contiguous()
        y = torch.empty_like(x2)
        s = complex(shift)
        sh = np.tile(np.array([s.real, s.imag]), x2.shape[0])
        ms = ctypes.c_double(0.0)
        self.ctx._enter()
        check(self.ctx.lib.cs_op_apply_timed(self.h, _dptr(sh), _tptr(x2), _tptr(y), x2.shape[0], int(x2.is_complex()),
                                             int(reps), ctypes.byref(ms)))
        self.ctx._exit()
        return ms.value

    def factor(self, shifts=(0.0,)):
        s = np.atleast_1d(np.asarray(shifts, dtype=complex))
        sh = np.zeros(2 * len(s))
        sh[0::2], sh[1::2] = s.real, s.imag
        cplx = int(np.any(s.imag != 0))
        self.ctx._enter()
        check(self.ctx.lib.cs_op_factor(self.h, _dptr(sh), len(s), cplx))
        self.ctx._exit()
        self._complex_shift = bool(cplx)

    def solve(self, rhs, shift_of_vec=None, rtol=1e-10, maxit=50):
        torch = _torch()
        r2 = rhs.reshape(-1, self.n).contiguous()
        nvec = r2.shape[0]
        x = torch.empty_like(r2)
        info = np.zeros(INFO_LEN)
        sov = None
        if shift_of_vec is not None:
            sov = np.ascontiguousarray(shift_of_vec, dtype=np.int32).ctypes.data_as(ctypes.POINTER(ctypes.c_int))
        self.ctx._enter()
        rc = self.ctx.lib.cs_op_solve(self.h, _tptr(r2), _tptr(x), nvec, sov, int(r2.is_complex()), rtol, maxit,
                                      info.ctypes.data_as(ctypes.POINTER(ctypes.c_double)))
        self.ctx._exit()
        self.info = dict(zip(INFO_NAMES, info))
        check(rc)
        return x.reshape(rhs.shape)

    def clean(self):
        self.ctx._enter()
        check(self.ctx.lib.cs_op_clean(self.h))
        self.ctx._exit()


class Context(object):
    """A solver context bound to one CUDA device (mirrors cs_ctx)."""

    def __init__(self, device=0, stream=None):
        """stream: raw cudaStream_t (int) to run on, e.g. torch.cuda.Stream().cuda_stream;
        None creates a private non-blocking stream."""
        torch = _torch()
        self.lib = load_library()
        self.device = torch.device("cuda", device)
        # The library runs on ONE explicit stream.  Torch allocations / copies made by the
        # caller live on torch's current stream, so every wrapper orders the two streams
        # (wait_stream in, wait_stream out) -- see _enter / _exit.
        if stream is None:
            self.tstream = torch.cuda.Stream(device=self.device)
        else:
            self.tstream = torch.cuda.ExternalStream(stream, device=self.device)
        h = ctypes.c_void_p()
        check(self.lib.cs_ctx_create(device, ctypes.c_void_p(self.tstream.cuda_stream), ctypes.byref(h)))
        self.h = h
        self.info = {}

    def _enter(self):
        torch = _torch()
        cur = torch.cuda.current_stream(self.device)
        if cur.cuda_stream != self.tstream.cuda_stream:
            self.tstream.wait_stream(cur)

    def _exit(self):
        torch = _torch()
        cur = torch.cuda.current_stream(self.device)
        if cur.cuda_stream != self.tstream.cuda_stream:
            cur.wait_stream(self.tstream)

    def close(self):
        if getattr(self, "h", None):
            self.lib.cs_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        self._enter()
        check(self.lib.cs_ctx_sync(self.h))
        self._exit()

    def launch_count(self):
        return int(self.lib.cs_launch_count())

    # -- z-slab decomposition of one system over several GPUs (one process per GPU) ------------------
    @staticmethod
    def comm_unique_id():
        """128-byte NCCL id drawn by rank 0; hand it to the other ranks (e.g. torch.distributed.broadcast)"""
        buf = ctypes.create_string_buffer(128)
        check(load_library().cs_comm_unique_id(buf))
        return buf.raw

    def comm_init(self, unique_id, rank, nranks, z_lo, z_hi):
        """this context owns the global cell levels [z_lo, z_hi); call before set_mesh (which then takes the global mesh)"""
        buf = ctypes.create_string_buffer(unique_id, 128) if unique_id is not None else None
        check(self.lib.cs_comm_init(self.h, buf, int(rank), int(nranks), int(z_lo), int(z_hi)))
        self.slab = (int(z_lo), int(z_hi))

    def local_range(self):
        """global cell levels [k0, k1) the local mesh covers (owned levels plus one ghost layer below)"""
        kk = (ctypes.c_int * 2)()
        check(self.lib.cs_comm_local_range(self.h, kk))
        return int(kk[0]), int(kk[1])

    def halo(self, vec, family):
        """refresh the ghost levels of a local vector (family 0 cells, 1 edges, 2 faces) in place"""
        n = {0: self.nC, 1: self.nE, 2: self.nF}[family]
        v2 = vec.reshape(-1, n)
        assert v2.is_contiguous()
        self._enter()
        rc = self.lib.cs_comm_halo(self.h, family, _tptr(v2), v2.shape[0], int(v2.is_complex()))
        self._exit()
        check(rc)
        return vec

    def set_refine(self, max_dd_steps):
        """iterative refinement of edge-form solves with a double-double residual (0 = off, the default)"""
        check(self.lib.cs_ctx_set_refine(self.h, int(max_dd_steps)))

    def fp64_peak_tflops(self):
        """DFMA throughput of this device (TFLOP/s), measured live: the FP64 roofline denominator"""
        v = ctypes.c_double(0.0)
        self._enter()
        rc = self.lib.cs_fp64_peak(self.h, ctypes.byref(v))
        self._exit()
        check(rc)
        return v.value

    # -- mesh / model -------------------------------------------------------
    def set_mesh(self, hx, hy, hz, z0, axis_weight_edge=1.0, axis_weight_face=0.5):
        hx, hy, hz = [np.ascontiguousarray(np.atleast_1d(h), dtype=np.float64) for h in (hx, hy, hz)]
        self.nx, self.nth, self.nz = len(hx), len(hy), len(hz)
        self._enter()
        check(self.lib.cs_mesh_set(self.h, self.nx, self.nth, self.nz, _dptr(hx), _dptr(hy), _dptr(hz), float(z0),
                                   float(axis_weight_edge), float(axis_weight_face)))
        self._exit()
        cnt = (ctypes.c_longlong * 7)()
        check(self.lib.cs_mesh_counts(self.h, cnt))
        self.nC, self.nFx, self.nFy, self.nFz, self.nEx, self.nEy, self.nEz = [int(v) for v in cnt]
        self.nF, self.nE = self.nFx + self.nFy + self.nFz, self.nEx + self.nEy + self.nEz

    def empty(self, n, complex_=False, cols=None):
        torch = _torch()
        shape = (n,) if cols is None else (cols, n)
        return torch.empty(shape, dtype=torch.complex128 if complex_ else torch.float64, device=self.device)

    def to_device(self, a, complex_=None):
        torch = _torch()
        a = np.asarray(a)
        if complex_ is None:
            complex_ = np.iscomplexobj(a)
        a = np.ascontiguousarray(a, dtype=np.complex128 if complex_ else np.float64)
        return torch.from_numpy(a).to(self.device)

    def paint(self, params):
        """params: the 18 doubles of k_paint.  Returns (sigma, mu) device tensors."""
        sigma, mu = self.empty(self.nC), self.empty(self.nC)
        self._enter()
        check(self.lib.cs_model_paint(self.h, _dptr(np.asarray(params, dtype=np.float64)), _tptr(sigma), _tptr(mu)))
        self._exit()
        return sigma, mu

    def set_model(self, sigma, mu):
        sigma = sigma if hasattr(sigma, "data_ptr") else self.to_device(sigma, False)
        mu = mu if hasattr(mu, "data_ptr") else self.to_device(mu, False)
        assert sigma.numel() == self.nC and mu.numel() == self.nC
        self._enter()
        check(self.lib.cs_model_set(self.h, _tptr(sigma), _tptr(mu)))
        self._exit()

    # -- building blocks ----------------------------------------------------
    def face_inner_product(self, prop=None, invProp=False, invMat=False):
        out = self.empty(self.nF)
        p = ctypes.c_void_p(0) if prop is None else _tptr(prop)
        self._enter()
        check(self.lib.cs_face_inner_product(self.h, p, int(invProp), int(invMat), _tptr(out)))
        self._exit()
        return out

    def edge_inner_product(self, prop=None, invProp=False, invMat=False):
        out = self.empty(self.nE)
        p = ctypes.c_void_p(0) if prop is None else _tptr(prop)
        self._enter()
        check(self.lib.cs_edge_inner_product(self.h, p, int(invProp), int(invMat), _tptr(out)))
        self._exit()
        return out

    def _unary(self, fn, x, nin, nout):
        assert x.numel() == nin
        y = self.empty(nout, x.is_complex())
        x = x.contiguous()
        self._enter()
        rc = fn(self.h, _tptr(x), _tptr(y), int(x.is_complex()))
        self._exit()
        check(rc)
        return y

    def curl(self, xE):
        return self._unary(self.lib.cs_curl, xE, self.nE, self.nF)

    def curlT(self, xF):
        return self._unary(self.lib.cs_curlT, xF, self.nF, self.nE)

    def div(self, xF):
        return self._unary(self.lib.cs_div, xF, self.nF, self.nC)

    def divT(self, xC):
        return self._unary(self.lib.cs_divT, xC, self.nC, self.nF)

    @property
    def nN(self):
        """nodes of a 3-D mesh: per z-node level the axis node, then nx*nth nodes"""
        return (self.nz + 1) * (self.nx * self.nth + 1)

    def grad(self, xN):
        """mesh.nodalGrad * xN (3-D meshes)"""
        return self._unary(self.lib.cs_grad, xN, self.nN, self.nE)

    def gradT(self, xE):
        """mesh.nodalGrad.T * xE (3-D meshes)"""
        return self._unary(self.lib.cs_gradT, xE, self.nE, self.nN)

    def ave_f2ccv(self, xF):
        """mesh.aveF2CCV * xF: cell-centred vector, blocks [x | y | z] ([x | z] on a symmetric mesh)"""
        return self._unary(self.lib.cs_ave_f2ccv, xF, self.nF, self.nC * (2 if self.nth == 1 else 3))

    def ave_e2ccv(self, xE):
        """mesh.aveE2CCV * xE: cell-centred vector, blocks [x | y | z] ([y] on a symmetric mesh)"""
        return self._unary(self.lib.cs_ave_e2ccv, xE, self.nE, self.nC * (1 if self.nth == 1 else 3))

    def op(self, kind, plan=True):
        if plan:
            return Op(self, kind)
        h = ctypes.c_void_p()
        self._enter()
        check(self.lib.cs_op_get_noplan(self.h, kind, ctypes.byref(h)))
        self._exit()
        return Op(self, kind, handle=h)

    def op_from_csr(self, A, family):
        """Operator from an assembled scipy.sparse CSR matrix (Solver(A) plugin path)."""
        A = A.tocsr()
        cplx = np.iscomplexobj(A.data)
        indptr = np.ascontiguousarray(A.indptr, dtype=np.int64)
        indices = np.ascontiguousarray(A.indices, dtype=np.int32)
        vals = np.ascontiguousarray(A.data, dtype=np.complex128 if cplx else np.float64)
        h = ctypes.c_void_p()
        self._enter()
        check(self.lib.cs_op_from_csr(self.h, family, A.shape[0], A.nnz,
                                      indptr.ctypes.data_as(ctypes.POINTER(ctypes.c_longlong)),
                                      indices.ctypes.data_as(ctypes.POINTER(ctypes.c_int)),
                                      vals.ctypes.data_as(ctypes.c_void_p), int(cplx), ctypes.byref(h)))
        self._exit()
        return Op(self, -1, handle=h)

    # -- physics-level drivers ----------------------------------------------
    def _info(self, info):
        self.info = dict(zip(INFO_NAMES, info))

    def solve_dc(self, q, rtol=1e-10, maxit=50):
        """q: (nrhs, nC) device tensor -> phi (nrhs, nC)."""
        torch = _torch()
        q2 = q.reshape(-1, self.nC).contiguous()
        phi = torch.empty_like(q2)
        info = np.zeros(INFO_LEN)
        self._enter()
        rc = self.lib.cs_solve_dc(self.h, _tptr(q2), _tptr(phi), q2.shape[0], rtol, maxit,
                                  info.ctypes.data_as(ctypes.POINTER(ctypes.c_double)))
        self._exit()
        self._info(info)
        check(rc)
        return phi.reshape(q.shape)

    def solve_fdem(self, form, freqs, s_e, rtol=1e-10, maxit=50):
        """s_e: real face vector (device); returns (nfreq, n) complex solutions."""
        freqs = np.ascontiguousarray(np.atleast_1d(freqs), dtype=np.float64)
        n = self.nE if form in ("h", "e") else self.nF
        nsrc = self.nE if form in ("e", "b") else self.nF        # E-B forms take an EDGE source vector
        if s_e.numel() != nsrc:
            raise CasingB200Error("the {} formulation takes a source vector on {} ({} entries), got {}".format(
                form, "edges" if form in ("e", "b") else "faces", nsrc, s_e.numel()))
        u = self.empty(n, True, cols=len(freqs))
        info = np.zeros(INFO_LEN)
        self._enter()
        rc = self.lib.cs_solve_fdem(self.h, ord(form), _dptr(freqs), len(freqs), _tptr(s_e.contiguous()), _tptr(u),
                                    rtol, maxit, info.ctypes.data_as(ctypes.POINTER(ctypes.c_double)))
        self._exit()
        self._info(info)
        check(rc)
        return u

    def tdem_run(self, form, dts, s_e, rtol=1e-10, maxit=50):
        """s_e: (nF,) -> (nsteps+1, nF), or (nsrc, nF) -> (nsteps+1, nsrc, nF): all sources share the factorisations"""
        dts = np.ascontiguousarray(dts, dtype=np.float64)
        multi = s_e.dim() == 2
        n = self.nE if form == "e" else self.nF                  # e: EDGE source vectors and fields
        if s_e.shape[-1] != n:
            raise CasingB200Error("the {} formulation takes source vectors on {} ({} entries), got {}".format(
                form, "edges" if form == "e" else "faces", n, s_e.shape[-1]))
        se2 = s_e.reshape(-1, n).contiguous()
        nsrc = se2.shape[0]
        torch = _torch()
        out = torch.empty((len(dts) + 1, nsrc, n), dtype=torch.float64, device=self.device)
        info = np.zeros(INFO_LEN)
        self._enter()
        rc = self.lib.cs_tdem_run_multi(self.h, ord(form), _dptr(dts), len(dts), _tptr(se2), nsrc, _tptr(out),
                                        rtol, maxit, info.ctypes.data_as(ctypes.POINTER(ctypes.c_double)))
        self._exit()
        self._info(info)
        check(rc)
        return out if multi else out[:, 0, :]

    def fields_fdem(self, form, which, freq, u, s_e):
        if form in ("e", "b"):
            n = self.nF if which == "b" else self.nE              # E-B forms: e, j on edges, b on faces
        else:
            n = self.nF if which in ("j", "e") else self.nE
        out = self.empty(n, True)
        self._enter()
        check(self.lib.cs_fields_fdem(self.h, ord(form), ord(which), float(freq), _tptr(u.contiguous()),
                                      _tptr(s_e.contiguous()), _tptr(out)))
        self._exit()
        return out

    def fields_dc(self, which, phi):
        phi2 = phi.reshape(-1, self.nC).contiguous()
        n = self.nC if which == "charge" else self.nF
        out = self.empty(n, False, cols=phi2.shape[0])
        code = {"j": "j", "e": "e", "charge": "c"}[which]
        self._enter()
        check(self.lib.cs_fields_dc(self.h, ord(code), _tptr(phi2), _tptr(out), phi2.shape[0]))
        self._exit()
        return out

    def casing_currents(self, j, casing_a, casing_b, casing_z):
        j = j.contiguous()
        ix = self.empty(self.nz, j.is_complex())
        iz = self.empty(self.nz + 1, j.is_complex())
        self._enter()
        check(self.lib.cs_casing_currents(self.h, _tptr(j), int(j.is_complex()), casing_a, casing_b,
                                          float(casing_z[0]), float(casing_z[1]), _tptr(ix), _tptr(iz)))
        self._exit()
        return ix, iz

    def casing_charges(self, charge, casing_b, casing_z):
        out = self.empty(self.nz)
        self._enter()
        check(self.lib.cs_casing_charges(self.h, _tptr(charge.contiguous()), casing_b, float(casing_z[0]),
                                         float(casing_z[1]), _tptr(out)))
        self._exit()
        return out
