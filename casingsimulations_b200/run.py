"""Simulation wrappers (kept API of casingSimulations/run.py:35-352).

``SimulationDC / SimulationFDEM / SimulationTDEM`` keep the reference's
constructor keywords, side effects (mkdir, JSON parameter files, ``np.save`` of
the ``<f>Solution`` array) and the indexable fields object, but ``prob.fields``
runs entirely on the B200 behind the C ABI: K1/K2 assembly, matrix-free
operators, theta-mode banded factorization and the preconditioned Krylov solve.
There is no CPU solver path.
"""
import os
import time

import numpy as np

from . import _lib
from .base import BaseCasing
from .model import PhysicalProperties, MU_0
from .solver import Solver
from .sources import SourceList


class _Survey(object):
    def __init__(self, srcList):
        self.srcList = list(srcList)
        self.source_list = self.srcList

    @property
    def freqs(self):
        out = []
        for s in self.srcList:
            if s.freq not in out:
                out.append(s.freq)
        return out

    nSrc = property(lambda self: len(self.srcList))


Survey = _Survey


class DipoleSrc(object):
    """dc.sources.Dipole([], a, b) stand-in (run.py:345-348): +I at the cell centre
    closest to a, -I at the one closest to b."""

    def __init__(self, rxList, locA, locB, current=1.0):
        self.rxList, self.loc, self.current = rxList, [np.asarray(locA, float), np.asarray(locB, float)], current


class Fields(object):
    """Indexable like SimPEG Fields objects: ``f[src | [srcs] | :, name]`` and, for
    TDEM, ``f[:, name, tInd]`` (usage: run.py:211, tests/test_cyl2D3D.py:212,266)."""

    def __init__(self, problem, solution, srcList, times=None):
        self.problem = self.simulation = self.prob = problem
        self._sol = solution              # numpy (n, nSrc) or (n, nSrc, nT+1)
        self.srcList = list(srcList)
        self.survey = problem.survey
        self.times = times
        self._cache = {}

    def _src_indices(self, sel):
        if isinstance(sel, slice):
            return list(range(len(self.srcList)))[sel]
        if isinstance(sel, (list, tuple)):
            return [self.srcList.index(s) for s in sel]
        return [self.srcList.index(sel)]

    def __getitem__(self, key):
        if not isinstance(key, tuple):
            key = (key,)
        sel = key[0] if len(key) > 0 else slice(None)
        name = key[1] if len(key) > 1 else self.problem._solution_name
        tind = key[2] if len(key) > 2 else slice(None)
        idx = self._src_indices(sel)
        out = self.problem._field(self, name, idx)
        if out.ndim == 3:
            out = out[:, :, tind]
            if out.ndim == 3 and out.shape[1] == 1:
                out = out[:, 0, :]
        return out


class _Problem(object):
    """Base of the SimPEG problem stand-ins: owns the GPU context."""
    _solution_name = None

    def __init__(self, mesh, Solver=Solver, verbose=False, device=0, paint_params=None, solver_opts=None, **kw):
        self.mesh, self.Solver, self.verbose, self.device = mesh, Solver, verbose, device
        self.paint_params = paint_params
        self.solver_opts = solver_opts or {}
        self.rtol = self.solver_opts.get("rtol", 1e-10)
        self.maxit = self.solver_opts.get("maxit", 50)
        self.survey = None
        self._ctx = None
        self.info = {}

    def pair(self, survey):
        self.survey = survey

    @property
    def ctx(self):
        if self._ctx is None:
            ctx = _lib.Context(self.device)
            m = self.mesh
            ctx.set_mesh(m.hx, m.hy, m.hz, m.x0[2], self.solver_opts.get("axis_weight_edge", 1.0),
                         self.solver_opts.get("axis_weight_face", 0.5))
            if self.solver_opts.get("dd_refine", 0):
                ctx.set_refine(self.solver_opts["dd_refine"])
            self._ctx = ctx
        return self._ctx

    def _set_model(self, m):
        ctx = self.ctx
        if m is None:
            assert self.paint_params is not None, "no model vector and no GPU paint parameters"
            self._sigma_d, self._mu_d = ctx.paint(self.paint_params)
        else:
            m = np.asarray(m, dtype=float)
            nC = self.mesh.nC
            assert m.size == 2 * nC, "model vector must be [sigma; mu] (model.py:848-868)"
            self._sigma_d, self._mu_d = ctx.to_device(m[:nC]), ctx.to_device(m[nC:])
            ctx.set_model(self._sigma_d, self._mu_d)

    def clean(self):
        if self._ctx is not None:
            self._ctx.close()
            self._ctx = None
        self._model_on_device = False

    # ---- fields rebuilt from a saved solution (utils.loadSimulationResults) ----
    def load_fields(self, solution, m=None, paint_params=None):
        """Fields object around a solution array read from ``fields.npy`` (run.py:208-212).  Nothing touches the GPU
        until a derived field (j, e, ...) is asked for; then the model is set and the solution uploaded once."""
        self._pending_model = (m, paint_params)
        self._model_on_device = False
        f = self._fields_from_array(np.asarray(solution))
        f._loaded = True
        return f

    def _ensure_model(self):
        if not getattr(self, "_model_on_device", False) and getattr(self, "_pending_model", None) is not None:
            m, pp = self._pending_model
            if pp is not None:
                self.paint_params = pp
            self._set_model(m)
            self._model_on_device = True


class Problem3D_CC(_Problem):
    """dc.Problem3D_CC(mesh, sigmaMap=..., bc_type='Dirichlet', Solver=...) (run.py:339-344)"""
    _solution_name = "phiSolution"

    def fields(self, m=None):
        self._set_model(m)
        ctx, mesh = self.ctx, self.mesh
        srcs = self.survey.srcList
        q = np.zeros((len(srcs), mesh.nC))
        from .mesh import closestPoints
        for s, src in enumerate(srcs):
            ia = closestPoints(mesh, src.loc[0], "CC")[0]
            ib = closestPoints(mesh, src.loc[1], "CC")[0]
            q[s, ia] += src.current
            q[s, ib] -= src.current
        phi = ctx.solve_dc(ctx.to_device(q), rtol=self.rtol, maxit=self.maxit)
        self.info = dict(ctx.info)
        f = Fields(self, phi.cpu().numpy().T.copy(), srcs)
        f._phi_d = phi
        return f

    def _fields_from_array(self, sol):
        srcs = self.survey.srcList
        assert sol.size == self.mesh.nC * len(srcs), "fields file does not match mesh / number of sources"
        sol = sol.reshape(self.mesh.nC, -1)
        return Fields(self, sol, srcs)

    def _field(self, f, name, idx):
        if name in ("phiSolution", "phi"):
            return f._sol[:, idx]
        ctx = self.ctx
        if getattr(f, "_phi_d", None) is None:                    # loaded from disk: upload once
            self._ensure_model()
            f._phi_d = ctx.to_device(np.ascontiguousarray(np.real(f._sol).T))
        phi = f._phi_d[idx].contiguous()
        if name in ("j", "e", "charge"):
            return ctx.fields_dc(name, phi).cpu().numpy().T.copy()
        if name == "charge_density":
            return ctx.fields_dc("charge", phi).cpu().numpy().T / self.mesh.vol[:, None]
        raise KeyError(name)


class _ProblemFDEM(_Problem):
    form = None

    @property
    def _solution_name(self):
        return self.form + "Solution"

    def fields(self, m=None):
        self._set_model(m)
        ctx = self.ctx
        torch = _lib._torch()
        srcs = self.survey.srcList
        n = self.mesh.nE if self.form in ("h", "e") else self.mesh.nF
        nsrc_vec = self.mesh.nE if self.form in ("e", "b") else self.mesh.nF
        for s in srcs:
            if np.size(s.s_e) != nsrc_vec:
                # what SimPEG does with such a source is a shape error deep inside getRHS; say it up front
                raise NotImplementedError(
                    "Problem3D_{f}: source vectors must live on {w} ({n} entries, got {g}).  casingSimulations sources are face "
                    "vectors (sources.py:236), which only the h and j formulations accept; the e / b formulations take an "
                    "edge vector (RawVec_e of the E-B formulation).".format(f=self.form, w="edges" if self.form in ("e", "b") else "faces",
                                                                          n=nsrc_vec, g=np.size(s.s_e)))
        # one contiguous column per source (Fortran order, like the (n, nSrc) arrays SimPEG fields hold): rows of a C-ordered
        # (nSrc, n) array, handed out transposed
        sol_c = np.empty((len(srcs), n), dtype=complex)
        sol = sol_c.T
        # sources that share one s_e vector (casingSimulations: one source per frequency,
        # sources.py:119-123) go to the GPU as one batched multi-frequency solve
        groups = []
        for i, s in enumerate(srcs):
            key = getattr(s, "_shared_s_e", None)           # set by BaseCasingSrc.srcList: same parent vector
            for g in groups:
                if (key is not None and g[2] is key) or g[0] is s.s_e or \
                        (key is None and g[2] is None and g[0].shape == s.s_e.shape and np.array_equal(g[0], s.s_e)):
                    g[1].append(i)
                    break
            else:
                groups.append((s.s_e, [i], key))
        self._se_d, self._u_d = {}, {}
        info = {}
        for se, members, key in groups:
            se_d = ctx.to_device(np.real(key if key is not None else se), complex_=False)
            freqs = np.array([srcs[i].freq for i in members])
            u = ctx.solve_fdem(self.form, freqs, se_d, rtol=self.rtol, maxit=self.maxit)
            for key, val in ctx.info.items():
                info[key] = max(info.get(key, 0.0), val) if key in ("relres", "iters") else info.get(key, 0.0) + val
            lo, hi = members[0], members[-1] + 1
            if members == list(range(lo, hi)):
                # device -> the caller's fresh array in ONE copy (a pinned staging buffer plus a host copy into fresh
                # pages cost twice the page traffic: 0.13 s for the 467 MB of the bench's two frequencies)
                torch.from_numpy(sol_c[lo:hi]).copy_(u)
            else:
                stage = getattr(self, "_stage", None)
                if stage is None or stage.shape != u.shape:
                    stage = self._stage = torch.empty(u.shape, dtype=u.dtype, pin_memory=True)
                stage.copy_(u)
                uh = stage.numpy()
                for k, i in enumerate(members):
                    sol_c[i] = uh[k]
            for k, i in enumerate(members):
                self._se_d[i], self._u_d[i] = se_d, u[k]
        self.info = info
        return Fields(self, sol, srcs)

    def _fields_from_array(self, sol):
        srcs = self.survey.srcList
        n = self.mesh.nE if self.form in ("h", "e") else self.mesh.nF
        assert sol.size == n * len(srcs), "fields file does not match mesh / number of sources"
        sol = sol.reshape(n, -1).astype(complex)
        self._se_d, self._u_d = {}, {}
        return Fields(self, sol, srcs)

    def _field(self, f, name, idx):
        if name == self._solution_name or name == self.form:
            return f._sol[:, idx]
        ctx = self.ctx
        if getattr(f, "_loaded", False):                          # loaded from disk: upload what is asked for
            self._ensure_model()
            for i in idx:
                if i not in self._u_d:
                    self._u_d[i] = ctx.to_device(np.ascontiguousarray(f._sol[:, i]), complex_=True)
                    self._se_d[i] = ctx.to_device(np.real(f.srcList[i].s_e), complex_=False)
        cols = [ctx.fields_fdem(self.form, name, f.srcList[i].freq, self._u_d[i], self._se_d[i]).cpu().numpy() for i in idx]
        return np.stack(cols, axis=1)


class Problem3D_h(_ProblemFDEM):
    form = "h"


class Problem3D_j(_ProblemFDEM):
    form = "j"


class Problem3D_e(_ProblemFDEM):
    """fdem.Problem3D_e (run.py:240-242 with formulation='e'): C^T MfMui C + i w MeSigma, rhs -i w s_e with s_e an EDGE
    vector; fields e, j (edges), b (faces).  3-D meshes only."""
    form = "e"


class Problem3D_b(_ProblemFDEM):
    """fdem.Problem3D_b: MfMui (C MeSigmaI C^T MfMui + i w), rhs MfMui C MeSigmaI s_e with s_e an EDGE vector"""
    form = "b"


class _ProblemTDEM(_Problem):
    form = "j"

    @property
    def _solution_name(self):
        return self.form + "Solution"

    def __init__(self, mesh, timeSteps=None, **kw):
        super(_ProblemTDEM, self).__init__(mesh, **kw)
        from .mesh import meshTensor
        self.timeSteps = meshTensor(timeSteps) if isinstance(timeSteps, list) else np.asarray(timeSteps, dtype=float)
        self.time_steps = self.timeSteps

    @property
    def times(self):
        return np.r_[0.0, np.cumsum(self.timeSteps)]

    def fields(self, m=None):
        self._set_model(m)
        ctx = self.ctx
        srcs = self.survey.srcList
        n = self.mesh.nE if self.form == "e" else self.mesh.nF
        for s in srcs:
            if np.size(s.s_e) != n:
                raise NotImplementedError(
                    "tdem Problem3D_{f}: source vectors must live on {w} ({n} entries, got {g}).  casingSimulations sources are "
                    "face vectors (sources.py:125, 236), which the j formulation accepts; the e formulation takes an edge "
                    "vector.".format(f=self.form, w="edges" if self.form == "e" else "faces", n=n, g=np.size(s.s_e)))
        # all sources of the survey at once: (n x nSrc) right-hand sides sharing every factorisation
        se = np.ascontiguousarray(np.stack([np.real(s.s_e) for s in srcs]))
        out = ctx.tdem_run(self.form, self.timeSteps, ctx.to_device(se, complex_=False), rtol=self.rtol, maxit=self.maxit)
        sol = np.ascontiguousarray(out.cpu().numpy().transpose(2, 1, 0))             # (n, nSrc, nT+1)
        info = dict(ctx.info)
        self.info = info
        return Fields(self, sol, srcs, times=self.times)

    def _fields_from_array(self, sol):
        srcs = self.survey.srcList
        nT = len(self.timeSteps) + 1
        if sol.ndim == 2:                                         # one source: Fields.__getitem__ dropped the source axis
            sol = sol[:, None, :]
        n = self.mesh.nE if self.form == "e" else self.mesh.nF
        assert sol.shape == (n, len(srcs), nT), "fields file does not match sources / time steps"
        return Fields(self, np.real(sol), srcs, times=self.times)

    def _field(self, f, name, idx):
        if name in (self.form + "Solution", self.form):
            return f._sol[:, idx, :]
        if self.form == "e":
            raise KeyError(name)
        if name == "e":      # e = Mf^-1 MfRho j
            ctx = self.ctx
            self._ensure_model()
            w = (ctx.face_inner_product(self._sigma_d, True, False) / ctx.face_inner_product()).cpu().numpy()
            return f._sol[:, idx, :] * w[:, None, None]
        raise KeyError(name)


class Problem3D_j_tdem(_ProblemTDEM):
    """tdem.Problem3D_j (run.py:287-296), face sources"""
    form = "j"


class Problem3D_e_tdem(_ProblemTDEM):
    """tdem.Problem3D_e: C^T MfMui C + MeSigma/dt on the edges, nodal DC initial state, EDGE source vectors"""
    form = "e"


class BaseSimulation(BaseCasing):
    """run.py:35-216"""
    _props = {"fields_filename": "fields.npy", "filename": "simulationParameters.json", "num_threads": 1,
              "modelParameters": None, "meshGenerator": None, "src": None, "srcList": None, "verbose": False,
              "formulation": None, "device": 0, "solver_opts": None}
    _instance_props = ("modelParameters", "meshGenerator", "src", "srcList")
    physics = None
    _formulations = ()

    def __init__(self, **kwargs):
        self._init_props(**kwargs)
        if not os.path.isdir(self.directory):
            os.mkdir(self.directory)
        if getattr(self.meshGenerator, "modelParameters", None) is None:
            self.meshGenerator.modelParameters = self.modelParameters
        if self.src is not None:
            if getattr(self.src, "modelParameters", None) is None:
                self.src.modelParameters = self.modelParameters
            if getattr(self.src, "meshGenerator", None) is None:
                self.src.meshGenerator = self.meshGenerator
            self.src.physics = self.physics
        if self.srcList is not None:
            for s in self.srcList.sources:
                s.physics = self.physics

    def _invalidate(self, name):
        pass    # simulations keep their lazily built problem / fields

    def _check_formulation(self, v):
        if v is not None and self._formulations and v not in self._formulations:
            raise ValueError("formulation must be one of {}".format(self._formulations))
        return v

    def _observe_src(self, value):
        if value is not None and self.__dict__.get("srcList", None) is None:
            object.__setattr__(self, "srcList", SourceList(sources=[value]))   # run.py:106-108

    @property
    def physprops(self):
        if getattr(self, "_physprops", None) is None:
            object.__setattr__(self, "_physprops", PhysicalProperties(self.meshGenerator, self.modelParameters))
        return self._physprops

    @property
    def survey(self):
        if getattr(self, "_survey", None) is None:
            self.prob
        return self._survey

    @property
    def mesh(self):
        return self.meshGenerator.mesh

    def serialize(self):
        out = BaseCasing.serialize(self)
        out.pop("device", None)
        return out

    def fields(self):
        if getattr(self, "_fields", None) is None:
            object.__setattr__(self, "_fields", self.run())
        return self._fields

    def _gpu_model(self):
        """None => paint on the GPU (K1); otherwise the [sigma; mu] vector of model.py:848-868."""
        pp = getattr(self.modelParameters, "paint_params", lambda: None)()
        return (None, pp) if pp is not None else (self.physprops.model, None)

    def run(self, save=True, verbose=False):
        print("Validating parameters...")
        self.validate()
        for obj in (self.modelParameters, self.meshGenerator, self.src):
            if obj is not None:
                obj.validate()
        sim_mesh = self.meshGenerator.mesh
        print("      max x: {}, min z: {}, max z: {}, nC: {}".format(
            sim_mesh.vectorNx.max(), sim_mesh.vectorNz.min(), sim_mesh.vectorNz.max(), sim_mesh.nC))
        if save:
            self.save()
        prb = self.prob
        if verbose is True:
            prb.verbose = True
        print("Starting {}".format(type(self).__name__))
        t = time.time()
        print("Using {} Solver".format(prb.Solver))
        m, pp = self._gpu_model()
        prb.paint_params = pp
        fields = prb.fields(m)
        print("   ... Done. Elapsed time : {}".format(time.time() - t))
        if save:
            np.save("/".join([self.directory, self.fields_filename]), fields[:, "{}Solution".format(self.formulation)])
        object.__setattr__(self, "_fields", fields)
        return fields


class SimulationFDEM(BaseSimulation):
    """run.py:219-263"""
    _props = {"formulation": "h"}
    _formulations = ("e", "b", "h", "j")
    physics = "fdem"

    @property
    def prob(self):
        if getattr(self, "_prob", None) is None:
            klass = {"h": Problem3D_h, "j": Problem3D_j, "e": Problem3D_e, "b": Problem3D_b}[self.formulation]
            some = self.srcList.srcList if self.srcList is not None else (self.src.srcList if self.src is not None else [])
            if self.formulation in ("e", "b") and some and np.size(some[0].s_e) != self.meshGenerator.mesh.nE:
                raise NotImplementedError(
                    "casingSimulations sources are face vectors (sources.py:236); only the h and j formulations accept "
                    "them.  The e / b formulations (Problem3D_e / Problem3D_b) take an edge source vector.")
            prob = klass(self.meshGenerator.mesh, Solver=Solver, verbose=self.verbose, device=self.device,
                         solver_opts=self.solver_opts)
            if self.srcList is not None:
                survey = _Survey(self.srcList.srcList)
            elif self.src is not None:
                survey = _Survey(self.src.srcList)
            else:
                raise Exception("one of src, srcList must be set")
            prob.survey = survey
            object.__setattr__(self, "_prob", prob)
            object.__setattr__(self, "_survey", survey)
        return self._prob


class SimulationTDEM(BaseSimulation):
    """run.py:266-307"""
    _props = {"formulation": "j"}
    _formulations = ("e", "b", "h", "j")
    physics = "tdem"

    @property
    def prob(self):
        if getattr(self, "_prob", None) is None:
            if self.formulation not in ("j", "e"):
                raise NotImplementedError("the time-domain path implements the j (default; face sources, sources.py:125) and e "
                                          "(edge sources) formulations")
            some = self.srcList.srcList
            if self.formulation == "e" and some and np.size(some[0].s_e) != self.meshGenerator.mesh.nE:
                raise NotImplementedError(
                    "casingSimulations sources are face vectors (sources.py:125, 236); only the j formulation accepts them.  "
                    "The e formulation (Problem3D_e) takes an edge source vector.")
            klass = Problem3D_e_tdem if self.formulation == "e" else Problem3D_j_tdem
            prob = klass(self.meshGenerator.mesh, timeSteps=self.modelParameters.timeSteps, Solver=Solver,
                         verbose=self.verbose, device=self.device, solver_opts=self.solver_opts)
            survey = _Survey(self.srcList.srcList)
            prob.survey = survey
            object.__setattr__(self, "_prob", prob)
            object.__setattr__(self, "_survey", survey)
        return self._prob


class SimulationDC(BaseSimulation):
    """run.py:310-352"""
    _props = {"src_a": None, "src_b": None, "fields_filename": "fieldsDC.npy", "formulation": "phi"}
    _formulations = ()
    physics = "dc"

    def __init__(self, **kwargs):
        super(SimulationDC, self).__init__(**kwargs)
        self.src_a = np.atleast_2d(np.asarray(self.src_a, dtype=float))
        self.src_b = np.atleast_2d(np.asarray(self.src_b, dtype=float))
        prob = Problem3D_CC(self.meshGenerator.mesh, Solver=Solver, device=self.device, solver_opts=self.solver_opts)
        srcs = [DipoleSrc([], self.src_a[i, :], self.src_b[i, :]) for i in range(self.src_a.shape[0])]
        survey = _Survey(srcs)
        prob.survey = survey
        object.__setattr__(self, "_srcList", srcs)
        object.__setattr__(self, "_prob", prob)
        object.__setattr__(self, "_survey", survey)

    def serialize(self):
        out = BaseSimulation.serialize(self)
        for k in ("src_a", "src_b"):
            if isinstance(out.get(k), np.ndarray):
                out[k] = out[k].tolist()
        return out

    @property
    def prob(self):
        return self._prob
