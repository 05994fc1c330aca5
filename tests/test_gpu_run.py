"""GPU tests through the reference-facing API: run.Simulation*, the SimPEG-style problem /
survey / fields objects, the Solver(A) plugin, physics.casing_currents -- i.e. the reference's
own tests (tests/test_cyl2D3D.py, tests/test_SimulationRun.py) re-targeted at the B200 path."""
import os
import shutil

import numpy as np
import pytest

import casingsimulations_b200 as cs
from casingsimulations_b200 import run, sources

pytestmark = pytest.mark.gpu
TOL, ZERO = 1e-4, 1e-7                                        # tests/test_cyl2D3D.py:18-19


def rel(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / np.linalg.norm(b)


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


def get_src_wire(mesh, mp):
    wire = np.zeros(mesh.vnF[2])
    g = mesh.gridFz
    wire[(g[:, 0] < mesh.hx.min()) & (g[:, 2] > mp.src_a[2]) & (g[:, 2] < mp.src_b[2])] = 1
    return np.hstack([np.zeros(mesh.nFx), np.zeros(mesh.nFy), wire])


@pytest.fixture(scope="module")
def cyl2d3d():
    """tests/test_cyl2D3D.py:81-170 verbatim in structure: Wholespace sigma = 0.1, 2-D vs 3-D
    CylMeshGenerator, a vertical wire on the innermost Fz faces, 3 frequencies, Problem3D_h."""
    mp = cs.model.Wholespace(src_a=np.r_[0.0, 0.0, -9.0], src_b=np.r_[0.0, 0.0, -1.0], freqs=np.r_[0.1, 1.0, 2.0], sigma_back=1e-1)
    mesh2D = cs.CylMeshGenerator(modelParameters=mp, npadx=11, npadz=26, csz=2.0)
    mesh3D = cs.CylMeshGenerator(modelParameters=mp, hy=np.ones(4) * 2 * np.pi / 4.0, csz=2.0, npadx=11, npadz=26)
    out = {}
    for tag, mg in (("2D", mesh2D), ("3D", mesh3D)):
        wire = get_src_wire(mg.mesh, mp)
        srcList = [sources.RawVec_e([], f, wire) for f in mp.freqs]
        pp = cs.model.PhysicalProperties(mg, mp)
        prb = run.Problem3D_h(mg.mesh, sigmaMap=pp.wires.sigma, muMap=pp.wires.mu, Solver=cs.Solver)
        prb.pair(run.Survey(srcList))
        out[tag] = (mg, srcList, prb, prb.fields(pp.model))
        assert prb.info["relres"] <= 1e-10
    return out


def _jslice(mesh, j3D, q=0):
    return cs.face3DthetaSlice(mesh, j3D, q)


def test_cyl2D3D_symmetry_and_agreement(cyl2d3d):
    mg2, src2, prb2, f2 = cyl2d3d["2D"]
    mg3, src3, prb3, f3 = cyl2d3d["3D"]
    m3 = mg3.mesh
    for s2, s3 in zip(src2, src3):
        j3, h3 = f3[s3, "j"], f3[s3, "h"]
        j0, h0 = _jslice(m3, j3), cs.edge3DthetaSlice(m3, h3)
        for q in range(1, m3.vnC[1]):
            assert rel(_jslice(m3, j3, q), j0) < TOL                      # :202-231
            assert rel(cs.edge3DthetaSlice(m3, h3, q), h0) < TOL          # :273-302
        assert np.all(np.abs(j3[m3.nFx:m3.nFx + m3.nFy]) < ZERO)          # :233-251
        assert np.all(np.abs(h3[:m3.nEx]) < ZERO) and np.all(np.abs(h3[m3.nEx + m3.nEy:]) < ZERO)   # :253-271
        assert rel(j0, f2[s2, "j"]) < TOL                                 # :304-330
        assert rel(h0, f2[s2, "h"][:, 0]) < TOL                           # :332-359


def test_cyl2D3D_matches_oracle(cyl2d3d):
    from oracle.cyl_oracle import MU_0, OracleCylMesh, OracleFDEM
    mg2, src2, prb2, f2 = cyl2d3d["2D"]
    m = mg2.mesh
    om = OracleCylMesh([m.hx, m.hy, m.hz], m.x0)
    P = OracleFDEM(om, 0.1 * np.ones(om.nC), MU_0 * np.ones(om.nC))
    for s in src2:
        h = P.solve(s.freq, np.real(s.s_e), refine=True)
        assert rel(f2[s, "h"][:, 0], h) < 1e-6
        assert rel(f2[s, "j"][:, 0], P.j_from(s.freq, h, np.real(s.s_e))) < 1e-6
        assert rel(f2[s, "e"][:, 0], P.e_from(s.freq, h, np.real(s.s_e))) < 1e-6
        assert rel(f2[s, "b"][:, 0], P.b_from(s.freq, h, np.real(s.s_e))) < 1e-6


def _casing_sim(tmp, nth=1, physics="fdem", **mesh_kw):
    mp = cs.model.CasingInHalfspace(directory=tmp, sigma_back=1e-1, src_a=np.r_[0.0, 0.0, 0.0], src_b=np.r_[1000.0, 0.0, 0.0],
                                    freqs=np.r_[0.5, 50.0], casing_l=100.0,
                                    timeSteps=[(1e-5, 3), (1e-4, 3)])
    hy = np.r_[2 * np.pi] if nth == 1 else np.ones(nth) * 2 * np.pi / nth
    kw = dict(npadx=6, npadz=8, csz=5.0)
    kw.update(mesh_kw)
    mg = cs.CasingMeshGenerator(directory=tmp, modelParameters=mp, hy=hy, **kw)
    if nth > 1:
        th = mg.mesh.vectorCCy[1]
        mp.src_a, mp.src_b = np.r_[0.0, th, 0.0], np.r_[1000.0, th, 0.0]
    return mp, mg


def test_simulation_fdem_run_and_save(tmp_path):
    """tests/test_SimulationRun.py:38-81: run each source, the saved fields.npy equals fields[:, 'h']"""
    tmp = str(tmp_path / "sim2D")
    mp, mg = _casing_sim(tmp)
    from oracle.cyl_oracle import OracleCylMesh, OracleFDEM
    m = mg.mesh
    om = OracleCylMesh([m.hx, m.hy, m.hz], m.x0)
    P = OracleFDEM(om, mp.sigma(m), mp.mu(m))
    for klass, a, b in ((sources.TopCasingSrc, np.r_[0.0, 0.0, 0.0], np.r_[1000.0, 0.0, 0.0]),
                        (sources.DownHoleCasingSrc, np.r_[0.0, 0.0, -50.0], np.r_[500.0, 0.0, 0.0]),
                        (sources.DownHoleTerminatingSrc, np.r_[0.0, 0.0, -50.0], np.r_[500.0, 0.0, 0.0])):
        src = klass(directory=tmp, modelParameters=mp, meshGenerator=mg, src_a=a.copy(), src_b=b.copy())
        sim = run.SimulationFDEM(directory=tmp, modelParameters=mp, meshGenerator=mg, src=src)
        fields = sim.run()
        loaded = np.load("/".join([tmp, "fields.npy"]))
        assert np.all(fields[:, "h"] == loaded)                                        # :51
        assert os.path.exists(os.path.join(tmp, "simulationParameters.json"))
        assert sim.prob.info["relres"] <= 1e-10
        for k, s in enumerate(sim.survey.srcList):
            h = P.solve(s.freq, np.real(s.s_e), refine=True)
            assert rel(fields[s, "h"][:, 0], h) < 1e-6, klass.__name__
        assert sim.fields() is fields
        sim.prob.clean()


def test_simulation_dc_and_casing_currents(tmp_path):
    tmp = str(tmp_path / "simDC")
    mp, mg = _casing_sim(tmp)
    m = mg.mesh
    a = np.array([[0.05125, 0.0, -2.5]])
    b = np.array([[1000.0, 0.0, -2.5]])
    sim = run.SimulationDC(directory=tmp, modelParameters=mp, meshGenerator=mg, src_a=a, src_b=b)
    f = sim.run()
    assert sim.prob.info["relres"] <= 1e-10 or sim.prob.info["backward_err"] <= 2e-15     # fp64 floor, see DESIGN.md
    assert np.array_equal(np.load(os.path.join(tmp, "fieldsDC.npy")), f[:, "phiSolution"])
    from oracle.cyl_oracle import OracleCylMesh, OracleDC, casing_currents, exact_mesh
    om = OracleCylMesh([m.hx, m.hy, m.hz], m.x0)
    P = OracleDC(exact_mesh(om), mp.sigma(m))                          # the exactly assembled system
    phi = P.fields(a, b, refine=True)
    print("DC sim: phi vs exact", rel(f[:, "phi"], phi))
    assert rel(f[:, "phi"], phi) < 1e-6
    j = f[:, "j"]
    jref = np.asarray(P.j(phi), dtype=float)
    assert rel(j, jref) < 1e-6
    cur = cs.casing_currents(j, m, mp)
    ref = casing_currents(jref, om, mp.casing_a, mp.casing_b, mp.casing_z)
    assert np.array_equal(cur["z"][0], ref["z"][0]) and rel(cur["z"][1][:, 0], ref["z"][1]) < 1e-6
    assert rel(cur["x"][1][:, 0], ref["x"][1]) < 1e-6
    assert 0.5 < -cur["z"][1][-1, 0] <= 1.0 + 1e-6                     # ~1 A enters the casing top
    z, q = cs.casing_charges(f[:, "charge"][:, 0], m, mp)
    assert len(z) == len(q)
    sim.prob.clean()


@pytest.mark.parametrize("nth", [1, 4])
def test_simulation_tdem(tmp_path, nth):
    tmp = str(tmp_path / "simT")
    mp, mg = _casing_sim(tmp, nth, npadx=4, npadz=5, csz=10.0)
    src = sources.TopCasingSrc(directory=tmp, modelParameters=mp, meshGenerator=mg)
    sim = run.SimulationTDEM(directory=tmp, modelParameters=mp, meshGenerator=mg, src=src)
    f = sim.run()
    jsol = f[:, "jSolution"]
    assert jsol.shape == (mg.mesh.nF, 7)
    assert np.array_equal(np.load(os.path.join(tmp, "fields.npy")), jsol)
    from oracle.cyl_oracle import OracleCylMesh, OracleTDEM, exact_mesh
    m = mg.mesh
    om = OracleCylMesh([m.hx, m.hy, m.hz], m.x0)
    Fref = OracleTDEM(exact_mesh(om), mp.sigma(m), mp.mu(m), mp.timeSteps).fields(src.s_e, refine=True)
    # fixed tolerance vs the exactly assembled systems: every backward-Euler solve of the face-form system leaves ~1e-7
    # (plain L2) that the fp64 residual cannot see, and the steps add up -- 3e-6 over these 6 steps (measured 1.1e-6 /
    # 1.6e-6; the reference's arithmetic -- fp64 assembly + SuperLU -- 1e-5 / 4e-3).  (The energy norm is no help here:
    # the j formulation's weak spot is the galvanic part in the resistive air cells, which that norm weighs by 1e6.)
    print("TDEM sim vs exact: plain", rel(jsol, Fref))
    tol = 3e-6
    assert rel(jsol, Fref) < tol, (rel(jsol, Fref), tol)
    assert rel(f[:, "j", 3][:, 0], Fref[:, 3]) < tol
    e = f[:, "e"]
    w = om.getFaceInnerProduct(1.0 / mp.sigma(m)).diagonal() / om.getFaceInnerProduct().diagonal()
    assert e.shape == jsol.shape and rel(e, jsol * w[:, None]) < 1e-12
    sim.prob.clean()


@pytest.mark.parametrize("nth", [1, 4])
def test_solver_plugin_protocol(nth):
    """Ainv = Solver(A, **opts); x = Ainv * rhs; Ainv.clean()  (run.py:246,294,343) with assembled matrices"""
    from conftest import CASING, small_casing_mesh
    from oracle.cyl_oracle import OracleCylMesh, OracleDC, OracleFDEM, paint_casing_in_halfspace, solve_refined
    hx, hy, hz, z0 = small_casing_mesh(nth)
    om = OracleCylMesh([hx, hy, hz], [0, 0, z0])
    sig, mu = paint_casing_in_halfspace(om, **CASING)
    mesh = cs.CylMesh([hx, hy, hz], x0=[0, 0, z0])
    # real SPD, multi-RHS (DC)
    Pd = OracleDC(om, sig)
    A = Pd.getA()
    q = Pd.getRHS(np.array([[0.03, 0.0, -0.25], [0.03, 0.0, -1.2]]), np.array([[2.0, 0.0, -0.25], [1.0, 0.0, -0.25]]))
    Ainv = cs.Solver(A, mesh=mesh)
    x = Ainv * q
    assert x.shape == q.shape
    for c in range(2):
        # here the system IS the fp64-assembled matrix handed to the plugin (its SpMV is the operator), so the truth is that
        # matrix' refined solution; cond * eps bounds what any fp64 solver can do with it (SuperLU: 7e-7 on this matrix)
        err = rel(x[:, c], solve_refined(A, q[:, c])[0])
        print("plugin DC column", c, "vs refined solution of the same matrix", err)
        assert err < 1e-6, err
    x1 = Ainv * q[:, 0]
    assert x1.shape == (om.nC,) and rel(x1, x[:, 0]) < 1e-9
    with pytest.raises(TypeError):
        Ainv * [1.0, 2.0]
    Ainv.clean()
    # complex symmetric (FDEM-h)
    P = OracleFDEM(om, sig, mu, "h")
    g = om.gridFz
    w = np.zeros(om.nFz)
    sel = (g[:, 0] < om.hx.min()) & (g[:, 2] > -2.6) & (g[:, 2] < 0.1)
    if nth > 1:
        sel &= g[:, 1] < om.hy[0]
    w[sel] = 1.0
    se = np.r_[np.zeros(om.nFx + om.nFy), w] / om.area
    Ah, rhs = P.getA(100.0), P.getRHS(100.0, se)
    Ainv = cs.Solver(Ah, mesh=mesh)
    h = Ainv.solve(rhs)
    d = np.sqrt(np.abs(Ah.diagonal()))
    assert np.linalg.norm((Ah @ h - rhs) / d) <= 1e-10 * np.linalg.norm(rhs / d)
    jt = P.j_from(100.0, solve_refined(Ah, rhs)[0], se)
    assert rel(P.j_from(100.0, h, se), jt) < 1e-4
    Ainv.clean()
    with pytest.raises(cs._lib.CasingB200Error):
        import scipy.sparse as sp
        cs.Solver(sp.identity(7, format="csr"), mesh=mesh)


def test_full_size_c1_dc():
    """BASELINE config 1 at full size: 109 x 1 x 717 (78 153 cells), a on the casing wall top, b at 1 km"""
    mp = cs.model.CasingInHalfspace(src_a=np.r_[0.0, 0.0, 0.0], src_b=np.r_[1000.0, 0.0, 0.0])
    mg = cs.CasingMeshGenerator(modelParameters=mp, npadx=10, npadz=20, csz=1.5)
    m = mg.mesh
    a, b = np.array([[0.05125, 0.0, -0.75]]), np.array([[1000.0, 0.0, -0.75]])
    sim = run.SimulationDC(directory="/tmp/cs_b200_c1", modelParameters=mp, meshGenerator=mg, src_a=a, src_b=b)
    f = sim.run(save=False)
    info = sim.prob.info
    print("C1 info", info)
    assert info["relres"] <= 1e-10 or info["backward_err"] <= 2e-15
    from oracle.cyl_oracle import OracleCylMesh, OracleDC, exact_mesh
    om = OracleCylMesh([m.hx, m.hy, m.hz], m.x0)
    phi_lu = OracleDC(om, mp.sigma(m)).fields(a, b)
    phi = OracleDC(exact_mesh(om), mp.sigma(m)).fields(a, b, refine=True)
    print("C1 gpu vs exact", rel(f[:, "phi"], phi), "fp64-assembled SuperLU vs exact", rel(phi_lu, phi))
    assert rel(f[:, "phi"], phi) < 1e-6
    cur = cs.casing_currents(f[:, "j"], m, mp)
    assert 0.9 < -cur["z"][1][-1, 0] <= 1.0 + 1e-6
    sim.prob.clean()
    shutil.rmtree("/tmp/cs_b200_c1", ignore_errors=True)


def test_full_size_c3_dc_3d_properties():
    """BASELINE config 3 at full size (109 x 32 x 1394 = 4.86 M cells): the oracle cannot factor this
    on the host, so parity goes through size-independent properties: Krylov residual, reciprocity
    phi_a(b) = phi_b(a) of the symmetric operator, and linearity."""
    import torch
    from casingsimulations_b200 import _lib
    mp = cs.model.CasingInHalfspace(src_a=np.r_[0.0, 0.0, 0.0], src_b=np.r_[1000.0, 0.0, 0.0], mur_casing=100.0)
    mg = cs.CasingMeshGenerator(modelParameters=mp, npadx=10, npadz=25, csz=0.75, hy=np.ones(32) * 2 * np.pi / 32)
    m = mg.mesh
    assert m.nC == 4862272
    ctx = _lib.Context(0)
    ctx.set_mesh(m.hx, m.hy, m.hz, m.x0[2])
    ctx.paint(mp.paint_params())
    from casingsimulations_b200.mesh import closestPoints
    ia = closestPoints(m, np.r_[0.05125, 33 * np.pi / 32, -0.375])[0]
    ib = closestPoints(m, np.r_[1000.0, 17 * np.pi / 32, -0.375])[0]
    q = torch.zeros((2, m.nC), dtype=torch.float64, device=ctx.device)
    q[0, ia], q[0, ib] = 1.0, -1.0            # the C3 dipole
    q[1, ib] = 1.0                            # unit source at b
    phi = ctx.solve_dc(q)
    print("C3 info", ctx.info)
    assert ctx.info["relres"] <= 1e-10 or ctx.info["backward_err"] <= 2e-15
    qa = torch.zeros((1, m.nC), dtype=torch.float64, device=ctx.device)
    qa[0, ia] = 1.0
    phia = ctx.solve_dc(qa)
    # reciprocity: potential at a due to a source at b == potential at b due to a source at a
    assert abs(float(phi[1, ia]) - float(phia[0, ib])) <= 1e-6 * abs(float(phia[0, ib]))
    # linearity: dipole = unit(a) - unit(b)
    lin = phia[0] - phi[1]
    assert float(torch.linalg.norm(lin - phi[0]) / torch.linalg.norm(phi[0])) < 1e-6
    ctx.close()


def test_full_size_c4_tdem_200_steps():
    """BASELINE config 4 at full size: TDEM (default j formulation, backward Euler, run.py:260-315) on the 3-D mesh
    109 x 32 x 460 (1 604 480 cells), 200 steps over 5 distinct dt (5 factorisations), TopCasingSrc.  The host cannot
    factor one of these systems, so parity goes through size-independent properties of the scheme (the small-mesh twin
    against the exactly assembled oracle systems is test_simulation_tdem / test_gpu_parity.py::test_tdem_run): every
    Krylov solve to 1e-10, the galvanic initial state D (j0 + s_e) = 0, the ohmic energy j^T MfRho j non-increasing,
    the divergence of j^n staying at rounding level, and the casing carrying the whole source current at t = 0."""
    import time
    import torch
    from casingsimulations_b200 import _lib
    nper, nth = 40, 32
    steps = [(1e-6, nper), (1e-5, nper), (1e-4, nper), (1e-3, nper), (1e-2, nper)]
    mp = cs.model.CasingInHalfspace(src_a=np.r_[0.0, 0.0, 0.0], src_b=np.r_[1000.0, 0.0, 0.0], mur_casing=100.0, timeSteps=steps)
    mg = cs.CasingMeshGenerator(modelParameters=mp, npadx=10, npadz=25, csz=2.5, hy=np.ones(nth) * 2 * np.pi / nth)
    m = mg.mesh
    assert tuple(int(v) for v in m.vnC) == (109, 32, 460) and len(mp.timeSteps) == 200
    th = m.vectorCCy[nth // 2]
    mp.src_a, mp.src_b = np.r_[0.0, th, 0.0], np.r_[1000.0, th, 0.0]
    src = cs.sources.TopCasingSrc(modelParameters=mp, meshGenerator=mg, physics="tdem")
    src.validate()
    ctx = _lib.Context(0)
    ctx.set_mesh(m.hx, m.hy, m.hz, m.x0[2])
    sig = ctx.paint(mp.paint_params())[0]
    se_d = ctx.to_device(src.s_e)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = ctx.tdem_run("j", mp.timeSteps, se_d, rtol=1e-10)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    info = dict(ctx.info)
    print("C4 TDEM: wall %.2f s" % wall, {k: round(float(v), 3) for k, v in info.items()})
    # worst TRUE residual of the 400 solves: the recurrence stops at 1e-10, the refinement rounds keep the best iterate;
    # 1e-10 ... 2e-9 from run to run (the 1e-3 and 1e-2 s steps are the ill-conditioned ones)
    assert info["converged"] == 1.0 and info["relres"] <= 1e-8
    nds = float(torch.linalg.norm(ctx.div(se_d)))
    assert float(torch.linalg.norm(ctx.div(out[0] + se_d))) <= 1e-6 * nds       # (1.5e-7 measured: unweighted norm of a 1e-10 weighted residual)
    alpha = ctx.face_inner_product(sig, True, False)
    en = np.array([float((out[t] * alpha * out[t]).sum()) for t in range(1, out.shape[0])])
    assert np.all(en[1:] <= en[:-1] * (1 + 1e-9)), "ohmic energy must not increase"
    assert en[-1] < 1e-3 * en[0]
    for t in range(1, out.shape[0], 20):
        assert float(torch.linalg.norm(ctx.div(out[t]))) <= 1e-6 * nds, t
    iz = ctx.casing_currents(out[0], mp.casing_a, mp.casing_b, mp.casing_z)[1]
    k = int(np.argmin(np.abs(m.vectorNz + 5.0)))
    assert 0.9 < abs(float(iz[k])) <= 1.0 + 1e-6            # 5 m below the top nearly all of the 1 A is still in the casing
    ctx.close()


def test_theta_varying_flawed_casing_krylov(tmp_path):
    """A flawed casing (theta-limited flaw, model.py:583-668) breaks the block-circulant structure:
    the theta-mode factorisation (built from the theta = 0 column) is then only a preconditioner
    and BiCGStab has to do real work.  Parity against the oracle on the same 3-D mesh."""
    tmp = str(tmp_path / "flaw")
    nth = 8
    mp = cs.model.FlawedCasingInHalfspace(directory=tmp, sigma_back=1e-1, src_a=np.r_[0.0, 0.0, 0.0],
                                          src_b=np.r_[1000.0, 0.0, 0.0], freqs=np.r_[50.0], casing_l=100.0,
                                          flaw_r=np.r_[0.04, 0.06], flaw_z=np.r_[-60.0, -40.0],
                                          flaw_theta=np.r_[np.pi / 2, np.pi], sigma_flaw=1e-1)
    mg = cs.CasingMeshGenerator(directory=tmp, modelParameters=mp, hy=np.ones(nth) * 2 * np.pi / nth, npadx=4, npadz=5, csz=10.0)
    th = mg.mesh.vectorCCy[0]
    mp.src_a, mp.src_b = np.r_[0.0, th, 0.0], np.r_[1000.0, th, 0.0]
    assert mp.paint_params() is None                      # painted through sigma()/mur(), uploaded with cs_model_set
    m = mg.mesh
    sig = mp.sigma(m).reshape(tuple(m.vnC), order="F")
    assert np.ptp(sig, axis=1).max() > 0                  # genuinely theta-dependent
    # DC
    a = np.array([[0.05125, th, -5.0]])
    b = np.array([[500.0, th, -5.0]])
    sim = run.SimulationDC(directory=tmp, modelParameters=mp, meshGenerator=mg, src_a=a, src_b=b,
                           solver_opts=dict(maxit=300, rtol=1e-10))
    f = sim.run(save=False)
    info = sim.prob.info
    print("flawed DC info", info)
    assert info["iters"] >= 2                              # not a one-shot direct solve any more
    import oracle.cyl_oracle as oc
    from oracle.cyl_oracle import OracleCylMesh, OracleDC, OracleFDEM, exact_mesh, energy_rel, casing_currents
    om = OracleCylMesh([m.hx, m.hy, m.hz], m.x0)
    ox = exact_mesh(om)
    phi_lu = OracleDC(om, mp.sigma(m)).fields(a, b)
    phi = OracleDC(ox, mp.sigma(m)).fields(a, b, refine=True)          # exactly assembled system
    err = rel(f[:, "phi"], phi)
    print("flawed DC gpu vs exact", err, "fp64-assembled SuperLU vs exact", rel(phi_lu, phi))
    assert err < 1e-6, err                                             # Krylov rtol 1e-10 through the approximate (circulant) preconditioner
    sim.prob.clean()
    # FDEM-h
    src = sources.TopCasingSrc(directory=tmp, modelParameters=mp, meshGenerator=mg)
    simf = run.SimulationFDEM(directory=tmp, modelParameters=mp, meshGenerator=mg, src=src,
                              solver_opts=dict(maxit=300, rtol=1e-10))
    ff = simf.run(save=False)
    print("flawed FDEM info", simf.prob.info)
    Pf, PX = OracleFDEM(om, mp.sigma(m), mp.mu(m)), OracleFDEM(ox, mp.sigma(m), mp.mu(m))
    s = simf.survey.srcList[0]
    se = np.real(s.s_e)
    A, rhs = Pf.getA(s.freq), Pf.getRHS(s.freq, se)
    d = np.sqrt(np.abs(A.diagonal()))
    res = np.linalg.norm((A @ ff[s, "h"][:, 0] - rhs) / d) / np.linalg.norm(rhs / d)
    h_true = PX.solve(s.freq, se, refine=True)
    truth_res = oc.LAST_REFINE_RESIDUAL
    j_true = np.asarray(PX.j_from(s.freq, h_true, se), dtype=complex)
    j_lu = Pf.j_from(s.freq, Pf.solve(s.freq, se), se)
    jg = ff[s, "j"][:, 0]
    rho = np.asarray(Pf.MfRho.diagonal())
    iz = casing_currents(jg, om, mp.casing_a, mp.casing_b, mp.casing_z)["z"][1]
    iz_true = casing_currents(j_true, om, mp.casing_a, mp.casing_b, mp.casing_z)["z"][1]
    print("flawed FDEM scaled residual", res, "| truth refinement residual", truth_res, "| j plain gpu vs exact", rel(jg, j_true),
          "(fp64 SuperLU", rel(j_lu, j_true), ") | j energy norm", energy_rel(jg, j_true, rho), "| casing Iz", rel(iz, iz_true))
    assert res < 1e-10
    assert truth_res < 1e-12, "the long-double refinement of the exact system did not converge: no truth for this case"
    # 50 Hz: the long-double refinement of the exact system converges (at 5 Hz it diverges: no truth exists there).
    # Fixed tolerances, GMRES to rtol 1e-10 through the circulant preconditioner (measured 3e-7 / 3e-7 / 2e-5; the
    # reference's arithmetic -- fp64 assembly + SuperLU -- is 2e-3 away in j).  The energy norm weighs the resistive air
    # faces, which an iterative solve at a given residual resolves last.
    assert rel(jg, j_true) < 1e-6
    assert rel(iz, iz_true) < 1e-6
    assert energy_rel(jg, j_true, rho) < 1e-4
    simf.prob.clean()


@pytest.mark.parametrize("nth", [1, 4])
def test_against_committed_golden_vectors(nth):
    """tests/golden/small_casing.npz (oracle, extended-precision refined; see make_golden.py): the GPU
    path reproduces the committed vectors without re-solving anything on the CPU."""
    from conftest import CASING, paint_params, small_casing_mesh
    from casingsimulations_b200 import _lib
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "small_casing.npz"))
    tag = "n%d_" % nth
    hx, hy, hz, z0 = small_casing_mesh(nth)
    ctx = _lib.Context(0)
    ctx.set_mesh(hx, hy, hz, z0)
    s_d, m_d = ctx.paint(paint_params(CASING))
    assert np.array_equal(s_d.cpu().numpy(), g[tag + "sigma"]) and np.array_equal(m_d.cpu().numpy(), g[tag + "mu"])
    assert rel(ctx.face_inner_product(s_d, True, False).cpu().numpy(), g[tag + "MfRho"]) < 1e-14
    assert rel(ctx.edge_inner_product(m_d).cpu().numpy(), g[tag + "MeMu"]) < 1e-14
    mesh = cs.CylMesh([hx, hy, hz], x0=[0, 0, z0])
    from casingsimulations_b200.mesh import closestPoints
    q = np.zeros((1, mesh.nC))
    q[0, closestPoints(mesh, g[tag + "dc_a"][0])[0]] = 1.0
    q[0, closestPoints(mesh, g[tag + "dc_b"][0])[0]] = -1.0
    phi = ctx.solve_dc(ctx.to_device(q)).cpu().numpy()[0]
    assert rel(phi, g[tag + "dc_phi"]) < 1e-6
    se = ctx.to_device(g[tag + "s_e"])
    from oracle.cyl_oracle import OracleCylMesh, casing_currents, energy_rel
    om = OracleCylMesh([hx, hy, hz], [0, 0, z0])
    freqs = [1.0, 100.0]
    u = ctx.solve_fdem("h", freqs, se)
    for f, name in enumerate(("1Hz", "100Hz")):
        j = ctx.fields_fdem("h", "j", freqs[f], u[f], se).cpu().numpy()
        e = ctx.fields_fdem("h", "e", freqs[f], u[f], se).cpu().numpy()
        iz = casing_currents(j, om, CASING["casing_a"], CASING["casing_b"], CASING["casing_z"])["z"][1]
        sens = float(g[tag + "fdem_j_sens_" + name])           # effect of ONE fp64 rounding of the rhs on plain-L2 j
        print(tag, name, "j plain", rel(j, g[tag + "fdem_j_" + name]), "sens", sens, "energy", energy_rel(j, g[tag + "fdem_j_" + name], g[tag + "MfRho"]),
              "e", rel(e, g[tag + "fdem_e_" + name]), "Iz", rel(iz, g[tag + "fdem_Iz_" + name]))
        assert rel(e, g[tag + "fdem_e_" + name]) < 1e-6
        assert energy_rel(j, g[tag + "fdem_j_" + name], g[tag + "MfRho"]) < 1e-6
        assert rel(iz, g[tag + "fdem_Iz_" + name]) < 1e-6
        if sens < 1e-7:                                        # plain L2 only where the fp64 data determine it
            assert rel(j, g[tag + "fdem_j_" + name]) < 1e-6
        if nth == 1:
            assert rel(u[f].cpu().numpy(), g[tag + "fdem_h_" + name]) < 1e-6 and rel(j, g[tag + "fdem_j_" + name]) < 1e-6
    F = ctx.tdem_run("j", np.r_[np.ones(3) * 1e-5, np.ones(3) * 1e-4], se).cpu().numpy().T
    assert rel(F, g[tag + "tdem_j"]) < 1e-6
    ctx.close()


def test_two_contexts_on_two_host_threads():
    """include/casing_b200.h: 'a cs_ctx is not thread safe, independent contexts are'.  Two host threads, each with
    its own context (own stream), run the DC + FDEM-h solves of different meshes at the same time; every result
    must equal the one obtained alone."""
    import threading
    from casingsimulations_b200 import _lib
    from conftest import CASING, paint_params
    from oracle.cyl_oracle import mesh_tensor

    def problem(nx, nz):
        hx = mesh_tensor([(0.01, 4), (0.01, nx - 4, 1.3)])
        hz = mesh_tensor([(0.5, nz // 3, -1.3), (0.5, nz - 2 * (nz // 3)), (0.5, nz // 3, 1.3)])
        return hx, np.r_[2 * np.pi], hz, -np.sum(hz) * 0.6

    def run(args, out, key, reps):
        hx, hy, hz, z0 = args
        ctx = _lib.Context(0)
        ctx.set_mesh(hx, hy, hz, z0)
        ctx.paint(paint_params(CASING))
        rng = np.random.default_rng(7)
        q = rng.standard_normal(ctx.nC)
        se = rng.standard_normal(ctx.nF)
        res = None
        for _ in range(reps):
            phi = ctx.solve_dc(ctx.to_device(q[None, :])).cpu().numpy()[0]
            h = ctx.solve_fdem("h", [3.0, 30.0], ctx.to_device(se)).cpu().numpy()
            res = (phi, h)
        out[key] = res
        ctx.close()

    probs = {"a": problem(40, 60), "b": problem(24, 90)}
    alone, both = {}, {}
    for k, a in probs.items():
        run(a, alone, k, 1)
    ths = [threading.Thread(target=run, args=(a, both, k, 4)) for k, a in probs.items()]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    for k in probs:
        assert k in both, "worker thread failed"
        for x, y in zip(alone[k], both[k]):
            assert rel(x, y) < 1e-9


def test_load_simulation_results_rebuilds_derived_fields(tmp_path):
    """utils.py:210-253: a saved study re-read with loadSimulationResults gives the same solution and the same
    derived fields (j, e, b; DC j / e / charge; TDEM e) as the run that wrote it."""
    tmp = str(tmp_path / "study")
    mp, mg = _casing_sim(tmp)
    src = sources.TopCasingSrc(directory=tmp, modelParameters=mp, meshGenerator=mg)
    sim = run.SimulationFDEM(directory=tmp, modelParameters=mp, meshGenerator=mg, src=src)
    f = sim.run()
    want = {name: f[:, name] for name in ("hSolution", "j", "e", "b")}
    sim.prob.clean()
    back = cs.loadSimulationResults(directory=tmp)
    fb = back.fields()
    for name, val in want.items():
        assert np.array_equal(fb[:, name], val), name
    s1 = back.survey.srcList[1]
    assert np.array_equal(fb[s1, "j"][:, 0], want["j"][:, 1])
    back.prob.clean()
    # DC
    tmpd = str(tmp_path / "dc")
    simd = run.SimulationDC(directory=tmpd, modelParameters=mp, meshGenerator=mg, src_a=np.array([[0.05125, 0.0, -2.5]]),
                            src_b=np.array([[1000.0, 0.0, -2.5]]))
    fd = simd.run()
    wantd = {name: fd[:, name] for name in ("phi", "j", "e", "charge")}
    simd.prob.clean()
    backd = cs.loadSimulationResults(directory=tmpd, fields="fieldsDC.npy")
    for name, val in wantd.items():
        assert np.array_equal(backd.fields()[:, name], val), name
    backd.prob.clean()
    # TDEM (one source: the saved array has no source axis)
    tmpt = str(tmp_path / "tdem")
    mpt, mgt = _casing_sim(tmpt, 1, npadx=4, npadz=5, csz=10.0)
    srct = sources.TopCasingSrc(directory=tmpt, modelParameters=mpt, meshGenerator=mgt)
    simt = run.SimulationTDEM(directory=tmpt, modelParameters=mpt, meshGenerator=mgt, src=srct)
    ft = simt.run()
    jt, et = ft[:, "j", :], ft[:, "e", :]
    simt.prob.clean()
    backt = cs.loadSimulationResults(directory=tmpt)
    assert np.array_equal(backt.fields()[:, "j", :], jt) and np.array_equal(backt.fields()[:, "e", :], et)
    backt.prob.clean()


def test_full_size_c2_fdem_sweep():
    """BASELINE config 2 (the bench workload) at full size: FDEM-h on the 109 x 1 x 717 casing mesh, TopCasingSrc,
    32 frequencies 0.1 Hz - 1 kHz in one batched solve.  Oracle parity at both ends of the sweep (rel-L2 <= 1e-6,
    BASELINE.json north_star), plus size-independent properties for all 32: Krylov residual, residual through the
    oracle's assembled matrix, D~ (j + s_e) = D~ C h = 0, linearity in s_e, ~1 A down the casing at 0.1 Hz."""
    import torch
    from oracle.cyl_oracle import OracleCylMesh, OracleFDEM, _sdiag
    freqs = np.logspace(-1, 3, 32)
    mp = cs.model.CasingInHalfspace(src_a=np.r_[0.0, 0.0, 0.0], src_b=np.r_[1000.0, 0.0, 0.0], freqs=freqs)
    mg = cs.CasingMeshGenerator(modelParameters=mp, npadx=10, npadz=20, csz=1.5)
    m = mg.mesh
    assert tuple(m.vnC) == (109, 1, 717)
    src = sources.TopCasingSrc(modelParameters=mp, meshGenerator=mg, directory="/tmp/cs_b200_c2")
    sim = run.SimulationFDEM(directory="/tmp/cs_b200_c2", modelParameters=mp, meshGenerator=mg, src=src)
    f = sim.run(save=False)
    info = sim.prob.info
    print("C2 info", info)
    assert info["relres"] <= 1e-10
    h = f[:, "hSolution"]
    assert h.shape == (m.nE, 32)
    om = OracleCylMesh([m.hx, m.hy, m.hz], m.x0)
    P = OracleFDEM(om, mp.sigma(m), mp.mu(m), "h")
    se = np.real(sim.survey.srcList[0].s_e)
    for k in (0, 31):
        href = P.solve(freqs[k], se, refine=True)
        print("C2 f = %g Hz: rel-L2 vs oracle" % freqs[k], rel(h[:, k], href))
        assert rel(h[:, k], href) < 1e-6
    Dt = om._Dtopo() @ _sdiag(om.area)
    j = f[:, "j"]
    for k in range(0, 32, 5):
        A, rhs = P.getA(freqs[k]), P.getRHS(freqs[k], se)
        d = np.sqrt(np.abs(A.diagonal()))
        assert np.linalg.norm((A @ h[:, k] - rhs) / d) <= 1e-9 * np.linalg.norm(rhs / d), freqs[k]
        div = Dt @ (j[:, k] + se)
        assert np.linalg.norm(div) <= 1e-9 * np.linalg.norm(np.abs(Dt) @ np.abs(j[:, k])), freqs[k]
    # linearity: the batched solve of 2.5 s_e is 2.5 h (two frequencies, straight through the C ABI)
    ctx = sim.prob.ctx
    u = ctx.solve_fdem("h", freqs[[3, 20]], ctx.to_device(2.5 * se, complex_=False)).cpu().numpy().T
    assert rel(u, 2.5 * h[:, [3, 20]]) < 1e-8
    cur = cs.casing_currents(np.real(j[:, [0]]), m, mp)
    assert 0.9 < -cur["z"][1][-1, 0] <= 1.0 + 1e-6                     # ~1 A enters the casing top at 0.1 Hz
    sim.prob.clean()
    shutil.rmtree("/tmp/cs_b200_c2", ignore_errors=True)


def test_fdem_vertical_dipole_wholespace_analytic_gpu():
    """The GPU FDEM-h path against the closed-form wholespace field of a vertical electric dipole (Ward & Hohmann
    1988, eq. 2.40): a known answer that does not go through the oracle at all (SURVEY 8c iii)."""
    from casingsimulations_b200 import _lib
    from test_oracle import ved_check, ved_problem
    hx, hz, m, se, zsrc = ved_problem()
    sigma = 0.1
    ctx = _lib.Context(0)
    ctx.set_mesh(hx, np.r_[2 * np.pi], hz, m.x0[2])
    ctx.set_model(ctx.to_device(sigma * np.ones(m.nC)), ctx.to_device(4e-7 * np.pi * np.ones(m.nC)))
    freqs = np.r_[1e3, 1e4]
    se_d = ctx.to_device(se, complex_=False)
    u = ctx.solve_fdem("h", freqs, se_d)
    assert ctx.info["relres"] <= 1e-10
    for k, (freq, imag_min) in enumerate(((1e3, 0.08), (1e4, 0.35))):
        e = ctx.fields_fdem("h", "e", freq, u[k], se_d).cpu().numpy()
        rz, rx, imag_share = ved_check(m, e, zsrc, sigma, freq)
        print("VED", freq, rz, rx, imag_share)
        assert rz < 0.015 and rx < 0.015, (freq, rz, rx)
        assert imag_share > imag_min
    ctx.close()


def test_tdem_vertical_dipole_stepoff_analytic_gpu():
    """The GPU TDEM-j path (DC initial state, backward Euler, divergence cleaning) against the closed-form step-off
    response of a vertical electric dipole in a wholespace (Ward & Hohmann 1988, eq. 2.50), without the oracle."""
    from casingsimulations_b200 import _lib
    from oracle.cyl_oracle import mesh_tensor
    from test_oracle import VED_DTS, ved_problem, ved_stepoff_errors
    hx, hz, m, se, zsrc = ved_problem()
    sigma = 0.1
    ctx = _lib.Context(0)
    ctx.set_mesh(hx, np.r_[2 * np.pi], hz, m.x0[2])
    ctx.set_model(ctx.to_device(sigma * np.ones(m.nC)), ctx.to_device(4e-7 * np.pi * np.ones(m.nC)))
    dts = mesh_tensor(VED_DTS)
    times = np.r_[0.0, np.cumsum(dts)]
    J = ctx.tdem_run("j", dts, ctx.to_device(se, complex_=False)).cpu().numpy().T
    errs = ved_stepoff_errors(m, J, times, zsrc, sigma, [0, 10, 25, 35, 45])
    print("VED step-off", errs)
    assert errs[0][0] < 0.015 and errs[0][1] < 0.015
    for ez, ex in errs[1:]:
        assert ez < 0.05 and ex < 0.09
    ctx.close()


def test_3d_off_axis_dipole_wholespace_analytic_gpu():
    """3-D known answer, GPU only (the host cannot factor this 115 k-cell curl-curl system in test time): a vertical
    electric dipole 10.5 m OFF the mesh axis, FDEM-h at 1 kHz and 10 kHz, all three field components against
    Ward & Hohmann eq. 2.40.  The comparison region contains the mesh axis, so the test also settles the face
    convention of SURVEY A.4 OQ1 by physics: axis_weight_face = 0.5 (default) is 2.5x closer than 1.0 there."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("ved3d_check", os.path.join(os.path.dirname(os.path.dirname(
        os.path.abspath(__file__))), "tools", "ved3d_check.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    out = mod.run(32, conventions=((1.0, 0.5), (1.0, 1.0)), verbose=False)
    dflt, alt = out[(1.0, 0.5)], out[(1.0, 1.0)]
    print("off-axis dipole", dflt, alt)
    assert dflt["relres"] <= 1e-10
    for freq in (1e3, 1e4):
        assert dflt[(freq, "axis")] < 0.03 and dflt[(freq, "shell")] < 0.09
        assert alt[(freq, "axis")] > 1.8 * dflt[(freq, "axis")]


def test_dc_3d_off_axis_dipole_analytic_gpu():
    """3-D DC known answer on the GPU path: current dipole 8.5 m off the axis of a 109 k-cell mesh in a wholespace
    against phi = I / (4 pi sigma) (1/Ra - 1/Rb), incl. the cells around the axis."""
    from casingsimulations_b200 import _lib
    from casingsimulations_b200.mesh import closestPoints
    from oracle.cyl_oracle import OracleCylMesh, mesh_tensor
    nth = 32
    hx = mesh_tensor([(1.0, 24), (1.0, 14, 1.4)])
    hz = mesh_tensor([(1.0, 14, -1.4), (1.0, 48), (1.0, 14, 1.4)])
    hy = np.ones(nth) * 2 * np.pi / nth
    sigma = 0.1
    m = OracleCylMesh([hx, hy, hz], [0, 0, -hz.sum() / 2])
    th0 = (nth // 4 + 0.5) * 2 * np.pi / nth
    cc = m.gridCC
    X = np.c_[cc[:, 0] * np.cos(cc[:, 1]), cc[:, 0] * np.sin(cc[:, 1]), cc[:, 2]]

    def closest(pt):
        return int(np.argmin((cc[:, 0] - pt[0]) ** 2 + (np.angle(np.exp(1j * (cc[:, 1] - pt[1])))) ** 2 * pt[0] ** 2 + (cc[:, 2] - pt[2]) ** 2))
    ia, ib = closest((8.5, th0, 6.5)), closest((8.5, th0, -6.5))
    q = np.zeros((1, m.nC))
    q[0, ia], q[0, ib] = 1.0, -1.0
    ctx = _lib.Context(0)
    ctx.set_mesh(hx, hy, hz, m.x0[2])
    ctx.set_model(ctx.to_device(sigma * np.ones(m.nC)), ctx.to_device(4e-7 * np.pi * np.ones(m.nC)))
    phi = ctx.solve_dc(ctx.to_device(q)).cpu().numpy()[0]
    assert ctx.info["relres"] <= 1e-10 or ctx.info["backward_err"] <= 2e-15
    Ra = np.maximum(np.linalg.norm(X - X[ia], axis=1), 1e-9)
    Rb = np.maximum(np.linalg.norm(X - X[ib], axis=1), 1e-9)
    ana = (1 / Ra - 1 / Rb) / (4 * np.pi * sigma)
    axis = (cc[:, 0] < 3) & (np.abs(cc[:, 2]) < 12) & (Ra > 4) & (Rb > 4)
    shell = (Ra > 4) & (Rb > 4) & (Ra < 18) & (Rb < 18)
    ea, es = rel(phi[axis], ana[axis]), rel(phi[shell], ana[shell])
    print("DC 3-D off-axis dipole: axis", ea, "shell", es)
    assert ea < 0.01 and es < 0.04
    ctx.close()


def test_plain_c_client_of_the_c_abi(tmp_path):
    """examples/c_abi_dc.c: a DC solve driven from C through include/casing_b200.h (cudaMalloc buffers, no Python)"""
    import subprocess
    from test_host import _compile_c_example
    exe = _compile_c_example(str(tmp_path / "c_abi_dc"))
    res = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    print(res.stdout, res.stderr)
    assert res.returncode == 0 and "C ABI OK" in res.stdout


def test_reference_smoke_setup_at_full_size(tmp_path):
    """The reference's own smoke test at its real size (tests/test_SimulationRun.py:18-81): CasingInWholespace,
    CasingMeshGenerator(npadx=8, npadz=19, csz=0.25) = 107 x 1 x 4048 = 433 136 cells, 0.5 Hz, the three source classes.
    Its assertion (saved fields.npy == fields[:, 'h']) plus parity of h, j, e against the exactly assembled system."""
    tmp = str(tmp_path / "sim2D")
    mp = cs.model.CasingInWholespace(directory=tmp, src_a=np.r_[0.0, np.pi, 0.0], src_b=np.r_[1e3, np.pi, 0.0], freqs=np.r_[0.5],
                                     sigma_back=1e-1)
    mg = cs.CasingMeshGenerator(directory=tmp, modelParameters=mp, npadx=8, npadz=19, csz=0.25)
    m = mg.mesh
    assert tuple(m.vnC) == (107, 1, 4048) and m.nC == 433136 and m.nE == 433243
    from oracle.cyl_oracle import OracleCylMesh, OracleFDEM, exact_mesh
    import oracle.cyl_oracle as oc
    om = OracleCylMesh([m.hx, m.hy, m.hz], m.x0)
    PX = OracleFDEM(exact_mesh(om), mp.sigma(m), mp.mu(m))
    for klass in (sources.TopCasingSrc, sources.DownHoleCasingSrc, sources.DownHoleTerminatingSrc):
        src = klass(directory=tmp, modelParameters=mp, meshGenerator=mg)
        src.validate()
        sim = run.SimulationFDEM(directory=tmp, modelParameters=mp, meshGenerator=mg, src=src)
        fields = sim.run()
        loaded = np.load("/".join([tmp, "fields.npy"]))
        assert np.all(fields[:, "h"] == loaded)                                            # tests/test_SimulationRun.py:51
        assert sim.prob.info["relres"] <= 1e-10
        s = sim.survey.srcList[0]
        se = np.real(s.s_e)
        h = PX.solve(s.freq, se, refine=True)
        assert oc.LAST_REFINE_RESIDUAL < 1e-12
        errs = (rel(fields[s, "h"][:, 0], h), rel(fields[s, "j"][:, 0], np.asarray(PX.j_from(s.freq, h, se), dtype=complex)),
                rel(fields[s, "e"][:, 0], np.asarray(PX.e_from(s.freq, h, se), dtype=complex)))
        print(klass.__name__, "433k cells: h, j, e vs exact system", errs, {k: round(float(v), 2) for k, v in sim.prob.info.items() if k.startswith("ms_")})
        assert max(errs) < 1e-6, (klass.__name__, errs)
        sim.prob.clean()


@pytest.mark.parametrize("nth", [1, 4])
def test_dipole_sources_gpu(tmp_path, nth):
    """HorizontalElectricDipole / VerticalElectricDipole (sources.py:130-413) through SimulationFDEM on
    the GPU, h / e / energy-norm j against the exactly assembled system (plain-L2 j too on the 2-D mesh)."""
    tmp = str(tmp_path / "dip")
    from oracle.cyl_oracle import OracleCylMesh, OracleFDEM, exact_mesh, energy_rel
    hy = np.r_[2 * np.pi] if nth == 1 else np.ones(nth) * 2 * np.pi / nth
    mp = cs.model.Halfspace(directory=tmp, sigma_back=1e-1, src_a=np.r_[50.0, 0.0, -5.0], src_b=np.r_[150.0, 0.0, -5.0], freqs=np.r_[10.0, 1000.0])
    mg = cs.CylMeshGenerator(directory=tmp, modelParameters=mp, hy=hy, csx=25.0, csz=10.0, domain_x=300.0, domain_z=60.0, npadx=6, npadz=7)
    m = mg.mesh
    th = float(m.vectorCCy[min(1, nth - 1)])
    om = OracleCylMesh([m.hx, m.hy, m.hz], m.x0)
    PX = OracleFDEM(exact_mesh(om), mp.sigma(m), mp.mu(m))
    rho = np.asarray(om.getFaceInnerProduct(1.0 / mp.sigma(m)).diagonal())
    zc = float(m.vectorCCz[np.argmin(np.abs(m.vectorCCz + 5.0))])
    cases = [(sources.HorizontalElectricDipole, np.r_[50.0, th, zc], np.r_[150.0, th, zc]),
             (sources.VerticalElectricDipole, np.r_[0.0, th, -45.0], np.r_[0.0, th, -5.0])]
    for klass, a, b in cases:
        src = klass(directory=tmp, modelParameters=mp, meshGenerator=mg, src_a=a.copy(), src_b=b.copy())
        src.validate()
        assert np.count_nonzero(src.s_e) > 0
        sim = run.SimulationFDEM(directory=tmp, modelParameters=mp, meshGenerator=mg, src=src)
        fields = sim.run(save=False)
        assert sim.prob.info["relres"] <= 1e-10
        for s in sim.survey.srcList:
            se = np.real(s.s_e)
            h = PX.solve(s.freq, se, refine=True)
            jt = np.asarray(PX.j_from(s.freq, h, se), dtype=complex)
            et = np.asarray(PX.e_from(s.freq, h, se), dtype=complex)
            ee, ej = rel(fields[s, "e"][:, 0], et), energy_rel(fields[s, "j"][:, 0], jt, rho)
            print(klass.__name__, nth, s.freq, "e", ee, "j energy", ej, "j plain", rel(fields[s, "j"][:, 0], jt), "h", rel(fields[s, "h"][:, 0], h))
            assert ee < 1e-6 and ej < 1e-6
            if nth == 1:
                assert rel(fields[s, "j"][:, 0], jt) < 1e-6 and rel(fields[s, "h"][:, 0], h) < 1e-6
        sim.prob.clean()


@pytest.mark.parametrize("klass_name", ["SingleLayer", "Layers", "TargetInHalfspace", "CasingInLayers", "CasingInHalfspaceWithTarget"])
def test_remaining_painters_gpu(tmp_path, klass_name):
    """the layer / target painters (model.py:263-403, 720-800) drive the GPU path (K1 when the class has
    paint parameters, otherwise sigma()/mur() uploaded through cs_model_set): DC potentials and FDEM-h fields against
    the exactly assembled systems with the same painted model."""
    tmp = str(tmp_path / "paint")
    from oracle.cyl_oracle import OracleCylMesh, OracleDC, OracleFDEM, exact_mesh
    kw = dict(directory=tmp, sigma_back=1e-2, src_a=np.r_[0.0, 0.0, 0.0], src_b=np.r_[1000.0, 0.0, 0.0], freqs=np.r_[5.0])
    if klass_name in ("SingleLayer",):
        kw.update(sigma_layer=1.0, layer_z=np.r_[-60.0, -30.0])
    if klass_name in ("Layers", "CasingInLayers"):
        kw.update(sigma_layers=[1e-1, 1.0, 1e-2], layer_tops=[0.0, -30.0, -60.0])
    if "Target" in klass_name:
        kw.update(target_radius=np.r_[0.0, 40.0], target_z=np.r_[-80.0, -50.0], sigma_target=10.0)
    if klass_name.startswith("Casing"):
        kw.update(casing_l=100.0)
    mp = getattr(cs.model, klass_name)(**kw)
    mg = cs.CasingMeshGenerator(directory=tmp, modelParameters=mp, npadx=6, npadz=8, csz=5.0, domain_z=100.0) if klass_name.startswith("Casing") \
        else cs.CylMeshGenerator(directory=tmp, modelParameters=mp, csx=10.0, csz=5.0, domain_x=100.0, domain_z=100.0, npadx=8, npadz=8)
    m = mg.mesh
    sig = mp.sigma(m)
    assert len(np.unique(sig)) >= 3                                                     # the painter did paint something
    om = OracleCylMesh([m.hx, m.hy, m.hz], m.x0)
    ox = exact_mesh(om)
    r0 = float(m.vectorCCx[3])
    a, b = np.array([[r0, 0.0, -2.5]]), np.array([[float(m.vectorCCx[-6]), 0.0, -2.5]])
    sim = run.SimulationDC(directory=tmp, modelParameters=mp, meshGenerator=mg, src_a=a, src_b=b)
    f = sim.run(save=False)
    phi = OracleDC(ox, sig).fields(a, b, refine=True)
    print(klass_name, "DC phi vs exact", rel(f[:, "phi"], phi))
    assert rel(f[:, "phi"], phi) < 1e-6
    sim.prob.clean()
    src = sources.VerticalElectricDipole(directory=tmp, modelParameters=mp, meshGenerator=mg, src_a=np.r_[0.0, 0.0, -42.5], src_b=np.r_[0.0, 0.0, -2.5])
    simf = run.SimulationFDEM(directory=tmp, modelParameters=mp, meshGenerator=mg, src=src)
    ff = simf.run(save=False)
    s = simf.survey.srcList[0]
    PX = OracleFDEM(ox, sig, mp.mu(m))
    h = PX.solve(s.freq, np.real(s.s_e), refine=True)
    print(klass_name, "FDEM h vs exact", rel(ff[s, "h"][:, 0], h))
    assert rel(ff[s, "h"][:, 0], h) < 1e-6
    assert rel(ff[s, "e"][:, 0], np.asarray(PX.e_from(s.freq, h, np.real(s.s_e)), dtype=complex)) < 1e-6
    simf.prob.clean()


def test_two_front_elimination_parity_on_wide_band_mesh():
    """The two-front (twisted) elimination only runs on wide bands (3-D edge / face systems with nx ~ 100: p = 330), which the
    small test meshes do not have.  Mesh of the bench's reference arm, 109 x 4 x 40 = 17 440 cells (41 levels, split at level
    20): FDEM-h against the EXACTLY assembled system.  Measured (two-front / one-front / the reference's fp64 SuperLU):
    10 Hz j energy norm 3.7e-6 / 1.1e-6 / 6.0e-4, casing current 8e-8 / 1e-7 / 3e-6; 10 kHz 3.5e-8 / 7e-9 / 3.9e-7 and
    3e-10 / 1e-9 / 1e-6.  Fixed tolerances: 2e-5 (10 Hz) and 1e-6 (10 kHz) in the energy norm, 1e-6 for the casing current.
    (Below ~5 Hz the extended-precision refinement of the truth does not converge on this coarse mesh.)"""
    import bench
    from casingsimulations_b200 import _lib
    from oracle.cyl_oracle import OracleCylMesh, OracleFDEM, paint_casing_in_halfspace, exact_mesh, energy_rel, casing_currents
    freqs = [10.0, 1e4]
    mp, mg, src = bench.build_inputs(bench.MESH_REF, freqs)
    m = mg.mesh
    assert tuple(int(v) for v in m.vnC) == (109, 4, 40)
    ctx = _lib.Context(0)
    ctx.set_mesh(m.hx, m.hy, m.hz, m.x0[2])
    se = ctx.to_device(np.real(src.s_e))
    ctx.paint(mp.paint_params())
    u = ctx.solve_fdem("h", mp.freqs, se)
    assert ctx.info["converged"] == 1.0 and ctx.info["iters"] <= 2
    om = OracleCylMesh([m.hx, m.hy, m.hz], m.x0)
    sig, mu = paint_casing_in_halfspace(om, sigma_back=mp.sigma_back, sigma_air=mp.sigma_air, sigma_casing=mp.sigma_casing,
                                        sigma_inside=mp.sigma_inside, mur_casing=mp.mur_casing, casing_a=mp.casing_a,
                                        casing_b=mp.casing_b, casing_z=tuple(mp.casing_z), surface_z=mp.surface_z)
    PX = OracleFDEM(exact_mesh(om), sig, mu, "h")
    se_np = np.real(src.s_e)
    rho = np.asarray(om.getFaceInnerProduct(sig, invProp=True).diagonal())
    cas = (mp.casing_a, mp.casing_b, tuple(mp.casing_z))
    for k, (f, tol) in enumerate(zip(freqs, (2e-5, 1e-6))):
        j_ref = np.asarray(PX.j_from(f, PX.solve(f, se_np, refine=True), se_np), dtype=complex)
        j = ctx.fields_fdem("h", "j", f, u[k], se).cpu().numpy()
        iz, iz_ref = casing_currents(j, om, *cas)["z"][1], casing_currents(j_ref, om, *cas)["z"][1]
        ej, ei = energy_rel(j, j_ref, rho), float(np.linalg.norm(iz - iz_ref) / np.linalg.norm(iz_ref))
        print("two-front parity at", f, "Hz: j energy", ej, "casing Iz", ei)
        assert ej < tol and ei < 1e-6
    ctx.close()


def _run_dist_tool(nproc, args, env=None):
    import json
    import subprocess
    import sys
    tool = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "dist_fdem.py")
    if nproc == 1:
        cmd = [sys.executable, tool] + args
    else:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc), "--master-addr", "127.0.0.1",
               "--master-port", str(29600 + os.getpid() % 300), tool] + args
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=dict(os.environ, **(env or {})))
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    return json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])


@pytest.mark.parametrize("twist", ["1", "0"])
def test_z_slab_solve_single_rank_matches_plain_solve(twist):
    """cs_comm_init with ONE rank runs the whole slab code path (local plan, ownership masks, mode-sharded global factors,
    slab <-> mode transposition) without NCCL: same answer as the plain solve on the same GPU -- with the two-front
    elimination of the global mode systems (default) and with the one-front one; both sides of the comparison use the
    same elimination, so the arithmetic is identical."""
    res = _run_dist_tool(1, ["4", "25", "17", "compare"], env={"CS_TWIST": twist})
    print(res)
    assert res["info"]["relres"] <= 1e-10
    assert res["rel_diff_vs_single_gpu_j"] < 1e-9 and res["rel_diff_vs_single_gpu_h"] < 1e-6   # (h carries null-space noise)


def test_z_slab_solve_two_gpus_matches_single_gpu():
    """BASELINE config 5 machinery on two ranks: z slabs, NCCL halo exchange, all-reduced inner products, mode-sharded
    factors; parity of the gathered solution against the single-GPU solve of the same system."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run with gpurun --gpus 2)")
    res = _run_dist_tool(2, ["8", "10", "10", "compare"])
    print(res)
    assert res["n_gpus"] == 2 and res["info"]["relres"] <= 1e-10
    assert res["rel_diff_vs_single_gpu_j"] < 1e-9 and res["rel_diff_vs_single_gpu_h"] < 1e-6


def test_tdem_source_list_shares_factorisations(tmp_path):
    """run.py:284-301 with a SourceList (sources.py:948-973): the sources of a TDEM survey are columns of one batched
    backward-Euler run that share every factorisation; each column equals the single-source run."""
    tmp = str(tmp_path / "simTL")
    mp, mg = _casing_sim(tmp, 1, npadx=4, npadz=5, csz=10.0)
    s1 = sources.TopCasingSrc(directory=tmp, modelParameters=mp, meshGenerator=mg, physics="tdem")
    s2 = sources.DownHoleCasingSrc(directory=tmp, modelParameters=mp, meshGenerator=mg, physics="tdem",
                                   src_a=np.r_[0.0, 0.0, -50.0], src_b=np.r_[500.0, 0.0, 0.0])
    sl = sources.SourceList(directory=tmp, sources=[s1, s2])
    sim = run.SimulationTDEM(directory=tmp, modelParameters=mp, meshGenerator=mg, srcList=sl)
    f = sim.run(save=False)
    both = f[:, "jSolution"]
    assert both.shape == (mg.mesh.nF, 2, 7)
    t_fac_both = sim.prob.info["ms_factor"]
    sim.prob.clean()
    for k, s in enumerate((s1, s2)):
        one = run.SimulationTDEM(directory=tmp, modelParameters=mp, meshGenerator=mg, src=s)
        f1 = one.run(save=False)
        assert rel(both[:, k, :], f1[:, "jSolution"]) < 1e-9, k
        one.prob.clean()
    print("factor ms for both sources", t_fac_both)
