"""Host-side logic of the facade (no GPU): mesh generators, model painters, sources,
serialization, and that the C-ABI library loads and exports every declared symbol."""
import ctypes
import os
import re
import warnings

import numpy as np
import pytest

import casingsimulations_b200 as cs
from casingsimulations_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _casing_setup(nth=1, **kw):
    mp = cs.model.CasingInHalfspace(src_a=np.r_[0.0, 0.0, 0.0], src_b=np.r_[1000.0, 0.0, 0.0], freqs=np.r_[1.0, 10.0])
    hy = np.r_[2 * np.pi] if nth == 1 else np.ones(nth) * 2 * np.pi / nth
    args = dict(npadx=10, npadz=20, csz=1.5)
    args.update(kw)
    mg = cs.CasingMeshGenerator(modelParameters=mp, hy=hy, **args)
    return mp, mg


def test_mesh_sizes_match_survey_appendix_b():
    mp, mg = _casing_setup()
    m = mg.mesh
    assert tuple(m.vnC) == (109, 1, 717) and (m.nC, m.nF, m.nE) == (78153, 156415, 78262)
    mp3, mg3 = _casing_setup(32, npadz=25, csz=0.75)
    assert tuple(mg3.mesh.vnC) == (109, 32, 1394) and mg3.mesh.nC == 4862272
    assert mg3.mesh.nF == 14590304 and mg3.mesh.nE == 14595186
    w = cs.model.Wholespace(src_a=np.r_[0.0, 0.0, -9.0], src_b=np.r_[0.0, 0.0, -1.0], freqs=np.r_[0.1, 1.0, 2.0], sigma_back=0.1)
    g2 = cs.CylMeshGenerator(modelParameters=w, npadx=11, npadz=26, csz=2.0)
    g3 = cs.CylMeshGenerator(modelParameters=w, hy=np.ones(4) * 2 * np.pi / 4.0, csz=2.0, npadx=11, npadz=26)
    assert (g2.mesh.nC, g2.mesh.nF, g2.mesh.nE) == (4026, 8113, 4087)          # tests/test_cyl2D3D.py:99-105
    assert (g3.mesh.nC, g3.mesh.nF, g3.mesh.nE) == (16104, 48556, 48866)
    assert np.isclose(g2.mesh.hz.min(), 2.0) and np.isclose(g2.mesh.hx.min(), 25.0)
    assert abs(mg.mesh.vectorNz[mg.npadz + mg.ncz - mg.nca]) < 1e-9             # z = 0 is a node (mesh.py:368-377)
    with pytest.raises(AssertionError):
        cs.CylMeshGenerator(modelParameters=w, hy=np.ones(4))                    # must sum to 2 pi (mesh.py:408-411)


def test_domain_z_override_is_kept():
    """mesh.py domain_z.setter: an explicit value wins over the model-derived extent, as attribute and as keyword,
    and survives copy() / later parameter changes (the cache invalidation must not drop it)."""
    mp = cs.model.CasingInHalfspace(src_a=np.r_[0.0, 0.0, 0.0], src_b=np.r_[1000.0, 0.0, 0.0])
    mg = cs.CasingMeshGenerator(modelParameters=mp, csz=10.0)
    assert mg.domain_z == 1000.0 and mg.ncz == 110
    mg.domain_z = 500.0
    assert mg.domain_z == 500.0 and mg.ncz == 60 and len(mg.hz) == 60 + 2 * mg.npadz
    mg.csz = 5.0                                     # another assignment: the override stays
    assert mg.domain_z == 500.0 and mg.ncz == 110
    mg2 = cs.CasingMeshGenerator(modelParameters=mp, domain_z=500.0, csz=10.0)
    assert mg2.domain_z == 500.0 and mg2.ncz == 60 and mg2.copy().domain_z == 500.0
    g3 = cs.CylMeshGenerator(modelParameters=mp, domain_z=200.0, csz=10.0)
    assert g3.ncz == 30


def test_mesh_geometry_matches_oracle():
    from oracle.cyl_oracle import OracleCylMesh
    mp, mg = _casing_setup(4, npadx=3, npadz=4, csz=50.0)
    m = mg.mesh
    om = OracleCylMesh([m.hx, m.hy, m.hz], m.x0)
    for name in ("gridCC", "gridFx", "gridFy", "gridFz", "gridEx", "gridEy", "area", "vol", "edge"):
        assert np.array_equal(getattr(m, name), getattr(om, name)), name
    assert np.array_equal(cs.meshTensor([(1.0, 3, -1.5), (1.0, 2)]), np.r_[1.5 ** 3, 1.5 ** 2, 1.5, 1, 1])


def test_model_defaults_and_predicates():
    """mirrors tests/test_model.py:21-51"""
    w = cs.model.Wholespace()
    assert w.sigma_back == 1e-2 and w.mur_back == 1.0
    with pytest.raises(ValueError):
        cs.model.Wholespace(sigma_back=-1.0)
    with pytest.raises(Exception):
        cs.model.Wholespace(not_a_parameter=1.0)
    mp, mg = _casing_setup()
    m = mg.mesh
    sig = mp.sigma(m)
    z, r = m.gridCC[:, 2], m.gridCC[:, 0]
    assert np.all(sig[z > 0] == mp.sigma_air)
    cas = (r > mp.casing_a) & (r < mp.casing_b) & (z > -1000) & (z < 0)
    assert np.all(sig[cas] == 5.5e6) and cas.sum() == 4 * 667
    assert np.all(sig[(r < mp.casing_a) & (z > -1000) & (z < 0)] == mp.sigma_inside)
    assert np.all(mp.mur(m) == 1.0) and np.allclose(mp.mu(m), 4e-7 * np.pi)
    assert np.isclose(mp.casing_a, 0.045) and np.isclose(mp.casing_b, 0.055) and np.allclose(mp.casing_z, [-1000, 0])
    with warnings.catch_warnings(record=True) as wlist:
        warnings.simplefilter("always")
        cs.model.Wholespace(version="0.0.0")
        assert any("version" in str(x.message) for x in wlist)
    pp = cs.model.PhysicalProperties(mg, mp)
    assert pp.model.shape == (2 * m.nC,) and np.array_equal(pp.wires.sigma(pp.model), sig)


def test_gpu_paint_params_reproduce_numpy_painters():
    """the 18 doubles handed to kernel K1 encode exactly the numpy painter (checked by emulating K1)"""
    def emulate(p, m):
        r, z = m.gridCC[:, 0], m.gridCC[:, 2]
        s, mr = np.full(m.nC, p[0]), np.full(m.nC, p[11])
        if p[1]:
            s[z > p[3]] = p[2]
        if p[14]:
            s[(z < p[17]) & (z > p[16])] = p[15]
        if p[4]:
            inz = (z > p[9]) & (z < p[10])
            c, i = inz & (r > p[7]) & (r < p[8]), inz & (r < p[7])
            s[c], mr[c] = p[5], p[12]
            s[i], mr[i] = p[6], p[13]
        return s, mr
    src = dict(src_a=np.r_[0.0, 0.0, 0.0], src_b=np.r_[1000.0, 0.0, 0.0])
    models = [cs.model.Wholespace(**src), cs.model.Halfspace(**src), cs.model.SingleLayer(sigma_layer=1.0, **src),
              cs.model.CasingInWholespace(mur_casing=80.0, **src), cs.model.CasingInHalfspace(mur_casing=100.0, sigma_inside=2.0, **src),
              cs.model.CasingInSingleLayer(sigma_layer=0.5, **src)]
    mg = cs.CasingMeshGenerator(modelParameters=models[4], npadx=4, npadz=6, csz=20.0)
    for mp in models:
        s, mr = emulate(mp.paint_params(), mg.mesh)
        assert np.array_equal(s, mp.sigma(mg.mesh)) and np.array_equal(mr, mp.mur(mg.mesh)), type(mp).__name__
    for klass in (cs.model.CasingInLayers, cs.model.FlawedCasingInHalfspace, cs.model.TargetInHalfspace,
                  cs.model.CasingInHalfspaceWithTarget):
        assert klass(**src).paint_params() is None                              # painted through sigma()/mur()
    flawed = cs.model.FlawedCasingInHalfspace(flaw_r=np.r_[0.04, 0.06], flaw_z=np.r_[-500.0, -490.0], sigma_flaw=1.0, **src)
    sf = flawed.sigma(mg.mesh)
    assert (sf == 1.0).sum() > 0


def test_top_casing_source():
    mp, mg = _casing_setup()
    src = cs.sources.TopCasingSrc(modelParameters=mp, meshGenerator=mg, physics="fdem")
    m = mg.mesh
    assert np.isclose(src.src_a[0], mp.casing_a + m.hx.min() / 2.0) and mp.src_a is src.src_a   # sources.py:910
    src.validate()
    cur = np.round(src.s_e * m.area, 12)
    assert set(np.unique(cur)) == {-1.0, 0.0, 1.0}
    # Kirchhoff: the wire path is closed except at the two grounding points
    from oracle.cyl_oracle import OracleCylMesh
    om = OracleCylMesh([m.hx, m.hy, m.hz], m.x0)
    d = np.round(om._Dtopo() @ cur, 9)
    assert np.count_nonzero(d) == 2 and np.isclose(d.sum(), 0.0)
    lst = src.srcList
    assert len(lst) == 2 and lst[0].freq == 1.0 and lst[0].s_e.dtype == complex
    # the sources of one survey carry their shared real parent vector (the solver groups them by it instead of comparing)
    assert lst[0]._shared_s_e is lst[1]._shared_s_e and lst[0]._shared_s_e.dtype == float
    assert np.array_equal(lst[0]._shared_s_e, np.real(lst[1].s_e))
    # the wire checks are memoised per parameter set and dropped with the other derived quantities on assignment
    assert ("surface wire", (1, 2)) in src._memo_assert_single
    src.src_b = np.r_[900.0, 0.0, 0.0]
    assert "_memo_assert_single" not in src.__dict__
    src.validate()
    assert ("surface wire", (1, 2)) in src._memo_assert_single
    src.src_b = np.r_[1000.0, 0.0, 0.0]
    # 3-D: theta must be a cell-centre azimuth
    mp3, mg3 = _casing_setup(8, npadx=4, npadz=6, csz=20.0)
    th = mg3.mesh.vectorCCy[3]
    mp3.src_a, mp3.src_b = np.r_[0.0, th, 0.0], np.r_[1000.0, th, 0.0]
    s3 = cs.sources.TopCasingSrc(modelParameters=mp3, meshGenerator=mg3, physics="tdem")
    s3.validate()
    om3 = OracleCylMesh([mg3.mesh.hx, mg3.mesh.hy, mg3.mesh.hz], mg3.mesh.x0)
    d3 = np.round(om3._Dtopo() @ (s3.s_e * mg3.mesh.area), 9)
    assert np.count_nonzero(d3) == 2
    assert len(s3.srcList) == 1


def test_other_sources_are_closed_wire_paths():
    from oracle.cyl_oracle import OracleCylMesh
    mp, mg = _casing_setup(1, npadx=6, npadz=8, csz=5.0)
    om = OracleCylMesh([mg.mesh.hx, mg.mesh.hy, mg.mesh.hz], mg.mesh.x0)
    mp.src_a, mp.src_b = np.r_[0.0, 0.0, -500.0], np.r_[500.0, 0.0, 0.0]
    for klass in (cs.sources.DownHoleTerminatingSrc, cs.sources.DownHoleCasingSrc):
        s = klass(modelParameters=mp, meshGenerator=mg, physics="fdem", src_a=mp.src_a.copy(), src_b=mp.src_b.copy())
        s.validate()
        d = np.round(om._Dtopo() @ (s.s_e * mg.mesh.area), 9)
        assert np.count_nonzero(d) == 2 and np.isclose(d.sum(), 0.0), klass.__name__
    with pytest.raises(AssertionError):
        cs.sources.TopCasingSrc(modelParameters=mp, meshGenerator=mg, src_a=np.r_[0, 0.1, 0.0], src_b=np.r_[10, 0.0, 0.0])


def test_serialization_round_trip(tmp_path):
    mp, mg = _casing_setup()
    mp.directory = str(tmp_path)
    mp.save()
    back = cs.load_properties(str(tmp_path / mp.filename))
    assert type(back) is cs.model.CasingInHalfspace and back.sigma_casing == mp.sigma_casing
    assert np.array_equal(back.freqs, mp.freqs) and np.array_equal(back.src_b, mp.src_b)
    mg.directory = str(tmp_path)
    mg.save()
    mgb = cs.load_properties(str(tmp_path / mg.filename))
    assert np.array_equal(mgb.hx, mg.hx) and np.array_equal(mgb.hz, mg.hz) and np.array_equal(mgb.x0, mg.x0)
    cp = mg.copy()
    assert cp.modelParameters is mg.modelParameters and np.array_equal(cp.hz, mg.hz)
    m2 = mg.create_2D_mesh()
    assert m2.mesh.isSymmetric
    ts = cs.model.CasingInHalfspace(timeSteps=[(1e-6, 4), (1e-5, 4)])
    assert len(ts.timeSteps) == 8 and np.isclose(ts.timeSteps.sum(), 4.4e-5)


def test_theta_slices():
    mp, mg = _casing_setup(4, npadx=3, npadz=4, csz=50.0)
    m = mg.mesh
    j = np.arange(m.nF, dtype=float)
    sl = cs.face3DthetaSlice(m, j, theta_ind=2)
    jx = j[:m.nFx].reshape(tuple(m.vnFx), order="F")
    assert sl.shape == (m.vnFx[0] * m.vnFx[2] + m.vnFz[0] * m.vnFz[2], 1) and sl[0, 0] == jx[0, 2, 0]
    h = np.arange(m.nE, dtype=float)
    assert cs.edge3DthetaSlice(m, h, 1)[0] == h[m.nEx + m.vnEy[0]]
    assert cs.mesh2d_from_3d(m).isSymmetric


def test_library_exports_every_declared_symbol():
    path = _lib.LIB_PATH
    assert os.path.exists(path), "run __graft_entry__.build() first"
    lib = ctypes.CDLL(path)
    header = open(os.path.join(ROOT, "include", "casing_b200.h")).read()
    declared = set(re.findall(r"\b(cs_[a-z_A-Z0-9]+)\s*\(", header))
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), name
    assert declared == set(_lib.EXPORTED_SYMBOLS)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.CasingB200Error):
        _lib.Context(0)
    mp, mg = _casing_setup(1, npadx=3, npadz=4, csz=50.0)
    src = cs.sources.TopCasingSrc(modelParameters=mp, meshGenerator=mg)
    sim = cs.run.SimulationFDEM(modelParameters=mp, meshGenerator=mg, src=src, directory="/tmp/cs_b200_nogpu")
    assert sim.formulation == "h" and len(sim.survey.srcList) == 2
    with pytest.raises(_lib.CasingB200Error):
        sim.run(save=False)
    import scipy.sparse as sp
    with pytest.raises(_lib.CasingB200Error):
        cs.Solver(sp.identity(4, format="csr"))                # no mesh hint -> raises, never falls back


def test_load_simulation_results_and_write_simulation_py(tmp_path):
    """utils.py:98-253: a study directory written by Simulation*.save() / np.save can be re-read;
    loading needs no GPU."""
    mp, mg = _casing_setup(npadx=3, npadz=4, csz=50.0)
    d = str(tmp_path / "study")
    src = cs.sources.TopCasingSrc(modelParameters=mp, meshGenerator=mg, directory=d)
    rng = np.random.default_rng(3)
    # FDEM: (nE, nFreq) complex
    sim = cs.run.SimulationFDEM(modelParameters=mp, meshGenerator=mg, src=src, directory=d)
    sim.save()
    sol = rng.standard_normal((mg.mesh.nE, 2)) + 1j * rng.standard_normal((mg.mesh.nE, 2))
    np.save(os.path.join(d, sim.fields_filename), sol)
    back = cs.loadSimulationResults(directory=d)
    assert type(back) is cs.run.SimulationFDEM and back.formulation == "h"
    assert np.array_equal(back.meshGenerator.mesh.hx, mg.mesh.hx) and back.modelParameters.sigma_casing == mp.sigma_casing
    f = back.fields()
    assert np.array_equal(f[:, "hSolution"], sol)
    assert np.array_equal(f[back.survey.srcList[1], "hSolution"][:, 0], sol[:, 1])
    # TDEM: one source -> (nF, nT + 1) as written by run.py:208-212
    mpt = cs.model.CasingInHalfspace(src_a=np.r_[0.0, 0.0, 0.0], src_b=np.r_[1000.0, 0.0, 0.0],
                                     timeSteps=[(1e-6, 3), (1e-5, 2)])
    mgt = cs.CasingMeshGenerator(modelParameters=mpt, npadx=3, npadz=4, csz=50.0)
    dt_ = str(tmp_path / "tdem")
    srct = cs.sources.TopCasingSrc(modelParameters=mpt, meshGenerator=mgt, directory=dt_)
    simt = cs.run.SimulationTDEM(modelParameters=mpt, meshGenerator=mgt, src=srct, directory=dt_)
    simt.save()
    solt = rng.standard_normal((mgt.mesh.nF, 6))
    np.save(os.path.join(dt_, simt.fields_filename), solt)
    backt = cs.loadSimulationResults(directory=dt_)
    assert np.array_equal(backt.fields()[:, "jSolution", :], solt)
    assert np.array_equal(backt.fields()[:, "jSolution", 2][:, 0], solt[:, 2])
    # DC
    dd = str(tmp_path / "dc")
    simd = cs.run.SimulationDC(modelParameters=mp, meshGenerator=mg, directory=dd, src_a=np.r_[0.05, 0.0, -10.0],
                               src_b=np.r_[500.0, 0.0, -10.0])
    simd.save()
    sold = rng.standard_normal((mg.mesh.nC, 1))
    np.save(os.path.join(dd, simd.fields_filename), sold)
    backd = cs.loadSimulationResults(directory=dd, fields=simd.fields_filename)
    assert np.array_equal(backd.fields()[:, "phiSolution"], sold)
    with pytest.raises(AssertionError):
        np.save(os.path.join(dd, simd.fields_filename), rng.standard_normal((mg.mesh.nC + 1, 1)))
        cs.loadSimulationResults(directory=dd, fields=simd.fields_filename)
    # writeSimulationPy: same script layout as the reference
    mp.directory = mg.directory = src.directory = d
    mp.save(); mg.save(); src.save()
    cs.writeSimulationPy(modelParameters=mp.filename, meshGenerator=mg.filename, src=src.filename, physics="FDEM", directory=d)
    text = open(os.path.join(d, "simulation.py")).read()
    compile(text, "simulation.py", "exec")
    for needle in ("casingSimulations.run.SimulationFDEM(", "fields = sim.run()", "mesh2D.hy = np.r_[2*np.pi]",
                   "casingSimulations.run.SimulationDC(", "src_a = sim.src.src_a_closest - np.r_[0., 0., csz/2.]"):
        assert needle in text
    cs.writeSimulationPy(physics="TDEM", directory=d, simulation_filename="s2.py", includeDC=False, include2D=False)
    t2 = open(os.path.join(d, "s2.py")).read()
    assert "SimulationTDEM" in t2 and "SimulationDC" not in t2 and "sim2D" not in t2


def _compile_c_example(out):
    import shutil
    import subprocess
    if shutil.which("gcc") is None or not os.path.exists("/usr/local/cuda/include/cuda_runtime.h"):
        pytest.skip("gcc or the CUDA headers are not available")
    libdir = os.path.join(ROOT, "casingsimulations_b200")
    cmd = ["gcc", "-O2", "-I", os.path.join(ROOT, "include"), "-I", "/usr/local/cuda/include",
           os.path.join(ROOT, "examples", "c_abi_dc.c"), "-L", libdir, "-lcasing_b200", "-L", "/usr/local/cuda/lib64",
           "-lcudart", "-lm", "-Wl,-rpath," + libdir, "-o", out]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    return out


def test_c_abi_header_compiles_and_links_from_plain_c(tmp_path):
    """the boundary is a C ABI: include/casing_b200.h is valid C and a C client links against the library"""
    exe = _compile_c_example(str(tmp_path / "c_abi_dc"))
    assert os.path.exists(exe)


def test_unsupported_formulations_raise_explicitly(tmp_path):
    """run.py:226-230, 273-277 accept formulation in e/b/h/j, but casingSimulations' sources are FACE vectors
    (sources.py:236, RawVec_e(s_e)): SimPEG's e / b problems integrate s_e with the EDGE inner product and cannot take
    them (a size mismatch nE vs nF).  The facade says so when the problem is built instead of failing inside a kernel."""
    mp = cs.model.CasingInHalfspace(directory=str(tmp_path), src_a=np.r_[0.0, 0.0, 0.0], src_b=np.r_[1000.0, 0.0, 0.0], freqs=np.r_[1.0],
                                    timeSteps=[(1e-5, 2)])
    mg = cs.CasingMeshGenerator(directory=str(tmp_path), modelParameters=mp, npadx=4, npadz=6, csz=20.0)
    for form in ("e", "b"):
        src = cs.sources.TopCasingSrc(directory=str(tmp_path), modelParameters=mp, meshGenerator=mg, physics="fdem")
        sim = cs.run.SimulationFDEM(directory=str(tmp_path), modelParameters=mp, meshGenerator=mg, src=src, formulation=form)
        with pytest.raises(NotImplementedError, match="face vectors"):
            sim.prob
    with pytest.raises(ValueError):
        cs.run.SimulationFDEM(directory=str(tmp_path), modelParameters=mp, meshGenerator=mg, src=src, formulation="x")
    src = cs.sources.TopCasingSrc(directory=str(tmp_path), modelParameters=mp, meshGenerator=mg, physics="tdem")
    for form in ("b", "h"):
        sim = cs.run.SimulationTDEM(directory=str(tmp_path), modelParameters=mp, meshGenerator=mg, src=src, formulation=form)
        with pytest.raises(NotImplementedError):
            sim.prob
