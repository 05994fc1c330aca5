"""Pins the CPU oracle (oracle/cyl_oracle.py): the reference ships no golden vectors, so the
acceptance checks of SURVEY.md section 8c are (i) the reference's own 2-D / 3-D consistency test
(tests/test_cyl2D3D.py:202-359) re-run on the oracle, (ii) discrete identities, (iii) analytic
solutions, plus the committed golden fixtures that guard against regressions."""
import os

import numpy as np
import pytest

from conftest import CASING, small_casing_mesh
from oracle.cyl_oracle import (MU_0, OracleCylMesh, OracleDC, OracleFDEM, OracleTDEM, casing_currents, mesh_tensor,
                               paint_casing_in_halfspace, solve_refined)

TOL, ZERO = 1e-4, 1e-7          # tests/test_cyl2D3D.py:18-19


def test_mesh_tensor():
    assert np.allclose(mesh_tensor([(2.0, 3)]), [2, 2, 2])
    assert np.allclose(mesh_tensor([(1.0, 2, 2.0)]), [2, 4])
    assert np.allclose(mesh_tensor([(1.0, 2, -2.0)]), [4, 2])


@pytest.mark.parametrize("nth", [1, 4])
def test_discrete_identities(nth):
    hx, hy, hz, z0 = small_casing_mesh(nth)
    m = OracleCylMesh([hx, hy, hz], [0, 0, z0])
    assert len(m.area) == m.nF and len(m.edge) == m.nE and len(m.vol) == m.nC
    assert abs(m.faceDiv @ m.edgeCurl).max() < 1e-10 * abs(m.edgeCurl).max()       # div curl = 0
    if nth > 1:
        assert abs(m.edgeCurl @ m.nodalGrad).max() < 1e-10 * abs(m.edgeCurl).max()  # curl grad = 0
        assert m.gridEz.shape[0] == m.nEz
    sig, mu = paint_casing_in_halfspace(m, **CASING)
    A = OracleDC(m, sig).getA()
    assert abs(A - A.T).max() < 1e-12 * abs(A).max()
    Ah = OracleFDEM(m, sig, mu, "h").getA(10.0)
    assert abs(Ah - Ah.T).max() < 1e-12 * abs(Ah).max()                             # complex symmetric
    # volumes / areas: total volume and the outer cylinder surface
    R, H = hx.sum(), hz.sum()
    assert np.isclose(m.vol.sum(), np.pi * R ** 2 * H)
    outer = m.area[:m.nFx].reshape(m.vnFx, order="F")[-1].sum()
    assert np.isclose(outer, 2 * np.pi * R * H)


def _ref_test_mesh(nth):
    """tests/test_cyl2D3D.py:87-105: CylMeshGenerator(csz=2, npadx=11, npadz=26), src z in [-9, -1]"""
    hx = mesh_tensor([(25.0, int(np.ceil(1000 / 25.0) + 10)), (25.0, 11, 1.5)])
    ncz = int(np.ceil(8.0 / 2.0)) + 5 + 5
    hz = mesh_tensor([(2.0, 26, -1.5), (2.0, ncz), (2.0, 26, 1.5)])
    hy = np.r_[2 * np.pi] if nth == 1 else np.ones(nth) * 2 * np.pi / nth
    return OracleCylMesh([hx, hy, hz], np.r_[0, 0, -np.sum(hz[:26 + ncz - 5])])


def _wire(m):
    g = m.gridFz
    w = np.zeros(m.nFz)
    w[(g[:, 0] < m.hx.min()) & (g[:, 2] > -9.0) & (g[:, 2] < -1.0)] = 1
    return np.r_[np.zeros(m.nFx + m.nFy), w]


def test_reference_cyl2D3D_consistency():
    """the reference's six assertions, at its own sizes (4 026 / 16 104 cells), one frequency"""
    m2, m3 = _ref_test_mesh(1), _ref_test_mesh(4)
    assert (m2.nC, m2.nF, m2.nE) == (4026, 8113, 4087) and (m3.nC, m3.nF, m3.nE) == (16104, 48556, 48866)
    freq = 1.0
    P2 = OracleFDEM(m2, 0.1 * np.ones(m2.nC), MU_0 * np.ones(m2.nC))
    P3 = OracleFDEM(m3, 0.1 * np.ones(m3.nC), MU_0 * np.ones(m3.nC))
    h2, h3 = P2.solve(freq, _wire(m2)), P3.solve(freq, _wire(m3))
    j2, j3 = P2.j_from(freq, h2, _wire(m2)), P3.j_from(freq, h3, _wire(m3))
    h3y = h3[m3.nEx:m3.nEx + m3.nEy].reshape(m3.vnEy, order="F")
    j3x = j3[:m3.nFx].reshape(m3.vnFx, order="F")
    j3z = j3[m3.nFx + m3.nFy:].reshape(m3.vnFz, order="F")

    def jslice(q):
        return np.r_[j3x[:, q, :].ravel(order="F"), j3z[:, q, :].ravel(order="F")]
    for q in range(1, 4):                                                   # :202-231, :273-302
        assert np.linalg.norm(jslice(0) - jslice(q)) / np.linalg.norm(jslice(0)) < TOL
        assert np.linalg.norm(h3y[:, 0] - h3y[:, q]) / np.linalg.norm(h3y[:, 0]) < TOL
    assert np.all(np.abs(j3[m3.nFx:m3.nFx + m3.nFy]) < ZERO)               # :233-251 (magnitudes)
    assert np.all(np.abs(h3[:m3.nEx]) < ZERO) and np.all(np.abs(h3[m3.nEx + m3.nEy:]) < ZERO)   # :253-271
    assert np.linalg.norm(jslice(0) - j2) / np.linalg.norm(j2) < TOL       # :304-330
    assert np.linalg.norm(h3y[:, 0].ravel(order="F") - h2) / np.linalg.norm(h2) < TOL   # :332-359


def test_dc_point_source_analytic():
    """phi = I / (4 pi sigma) (1/Ra - 1/Rb) away from the electrodes and the boundary"""
    hx = mesh_tensor([(1.0, 40), (1.0, 18, 1.4)])
    hz = mesh_tensor([(1.0, 18, -1.4), (1.0, 80), (1.0, 18, 1.4)])
    m = OracleCylMesh([hx, np.r_[2 * np.pi], hz], [0, 0, -hz.sum() / 2])
    sigma = 0.1
    P = OracleDC(m, sigma * np.ones(m.nC))
    a, b = np.array([[0.0, 0.0, 10.5]]), np.array([[0.0, 0.0, -10.5]])
    phi = P.fields(a, b)[:, 0]
    cc = m.gridCC
    ia, ib = np.argmax(P.getRHS(a, b)[:, 0]), np.argmin(P.getRHS(a, b)[:, 0])
    Ra = np.sqrt(cc[:, 0] ** 2 + (cc[:, 2] - cc[ia, 2]) ** 2)
    Rb = np.sqrt(cc[:, 0] ** 2 + (cc[:, 2] - cc[ib, 2]) ** 2)
    ana = 1.0 / (4 * np.pi * sigma) * (1 / Ra - 1 / Rb)
    sel = (Ra > 6) & (Rb > 6) & (cc[:, 0] < 30) & (np.abs(cc[:, 2]) < 30)
    assert np.linalg.norm(phi[sel] - ana[sel]) / np.linalg.norm(ana[sel]) < 0.03


def test_tdem_initial_state_and_decay():
    hx, hy, hz, z0 = small_casing_mesh(1)
    m = OracleCylMesh([hx, hy, hz], [0, 0, z0])
    sig, mu = paint_casing_in_halfspace(m, **CASING)
    g = m.gridFz
    w = np.zeros(m.nFz)
    w[(g[:, 0] < m.hx.min()) & (g[:, 2] > -2.6) & (g[:, 2] < 0.1)] = 1.0
    se = np.r_[np.zeros(m.nFx), w] / m.area
    T = OracleTDEM(m, sig, mu, [(1e-5, 4), (1e-3, 4)], ground_first_cell=False)
    F = T.fields(se, refine=True)
    # galvanic DC state: D (j0 + s_e) = 0  
    div0 = T.Dt @ (F[:, 0] + se)
    assert np.linalg.norm(div0) < 1e-6 * np.linalg.norm(T.Dt @ se)      # raw 2-norm floor of the 1e13-contrast system
    en = [float(F[:, t] @ (T.MfRho @ F[:, t])) for t in range(F.shape[1])]
    assert all(en[t + 1] <= en[t] * (1 + 1e-9) for t in range(1, len(en) - 1))      # ohmic energy decays


def test_casing_currents_top_source():
    """I_z just below a source grounded on the casing top is ~1 A and leaks off monotonically"""
    hx, hy, hz, z0 = small_casing_mesh(1)
    m = OracleCylMesh([hx, hy, hz], [0, 0, z0])
    sig, mu = paint_casing_in_halfspace(m, **CASING)
    P = OracleDC(m, sig)
    a, b = np.array([[0.03, 0.0, -0.25]]), np.array([[4.0, 0.0, -0.25]])
    phi = P.fields(a, b, refine=True)
    j = P.j(phi)
    cur = casing_currents(j, m, CASING["casing_a"], CASING["casing_b"], CASING["casing_z"])
    iz = -cur["z"][1]                    # downward current
    assert 0.5 < iz[-1] <= 1.0 + 1e-9
    assert np.all(np.diff(iz) >= -1e-9)  # grows towards the top = leaks off downwards


def test_refined_solve_beats_plain_lu():
    hx, hy, hz, z0 = small_casing_mesh(1)
    m = OracleCylMesh([hx, hy, hz], [0, 0, z0])
    sig, mu = paint_casing_in_halfspace(m, **CASING)
    P = OracleDC(m, sig)
    A = P.getA()
    q = P.getRHS(np.array([[0.03, 0, -0.25]]), np.array([[2.0, 0, -0.25]]))[:, 0]
    x, x0 = solve_refined(A, q)
    d = np.sqrt(A.diagonal())
    r, r0 = (q - A @ x) / d, (q - A @ x0) / d
    assert np.linalg.norm(r) <= np.linalg.norm(r0) * 1.0001


def test_golden_fixture_matches_oracle():
    p = os.path.join(os.path.dirname(__file__), "golden", "small_casing.npz")
    g = np.load(p)
    for nth in (1, 4):
        hx, hy, hz, z0 = small_casing_mesh(nth)
        m = OracleCylMesh([hx, hy, hz], [0, 0, z0])
        sig, mu = paint_casing_in_halfspace(m, **CASING)
        tag = "n%d_" % nth
        assert np.array_equal(g[tag + "sigma"], sig) and np.array_equal(g[tag + "mu"], mu)
        assert np.allclose(g[tag + "MfRho"], m.getFaceInnerProduct(sig, invProp=True).diagonal(), rtol=1e-14)
        assert np.allclose(g[tag + "MeMu"], m.getEdgeInnerProduct(mu).diagonal(), rtol=1e-14)
        if nth == 1:
            P = OracleFDEM(m, sig, mu, "h")
            h = P.solve(100.0, g[tag + "s_e"], refine=True)
            assert np.linalg.norm(h - g[tag + "fdem_h_100Hz"]) / np.linalg.norm(h) < 1e-9


def ved_wholespace(rho, z, zs, sigma, freq):
    """E of a unit-moment (1 A m) vertical electric dipole at (0, zs) in a wholespace, e^{+i w t} convention,
    quasi-static: Ward & Hohmann (1988) eq. 2.40 rotated to z.  Returns (E_z, E_rho)."""
    k = np.sqrt(-1j * 2 * np.pi * freq * MU_0 * sigma)
    if k.imag > 0:
        k = -k                                       # e^{-ikr} must decay
    zz = z - zs
    r = np.sqrt(rho ** 2 + zz ** 2)
    pre = np.exp(-1j * k * r) / (4 * np.pi * sigma * r ** 3)
    A = -k ** 2 * r ** 2 + 3j * k * r + 3
    B = k ** 2 * r ** 2 - 1j * k * r - 1
    return pre * (zz ** 2 / r ** 2 * A + B), pre * (rho * zz / r ** 2 * A)


def ved_problem():
    hx = mesh_tensor([(1.0, 60), (1.0, 22, 1.35)])
    hz = mesh_tensor([(1.0, 22, -1.35), (1.0, 120), (1.0, 22, 1.35)])
    m = OracleCylMesh([hx, np.r_[2 * np.pi], hz], [0, 0, -hz.sum() / 2])
    g = m.gridFz
    sel = (g[:, 0] < hx[0]) & (g[:, 2] > -2.5) & (g[:, 2] < 2.5)          # 1 A along the axis cells, 5 faces of 1 m
    w = np.zeros(m.nFz)
    w[sel] = 1.0
    se = np.r_[np.zeros(m.nFx + m.nFy), w] / m.area
    return hx, hz, m, se, g[sel, 2]


def ved_check(m, e, zsrc, sigma, freq):
    gz, gx = m.gridFz, m.gridFx
    Ez = sum(ved_wholespace(gz[:, 0], gz[:, 2], zs, sigma, freq)[0] for zs in zsrc)
    Ex = sum(ved_wholespace(gx[:, 0], gx[:, 2], zs, sigma, freq)[1] for zs in zsrc)
    Rz, Rx = np.hypot(gz[:, 0], gz[:, 2]), np.hypot(gx[:, 0], gx[:, 2])
    sz, sx = (Rz > 12) & (Rz < 40), (Rx > 12) & (Rx < 40)
    ez, ex = e[m.nFx + m.nFy:], e[:m.nFx]
    return (np.linalg.norm(ez[sz] - Ez[sz]) / np.linalg.norm(Ez[sz]), np.linalg.norm(ex[sx] - Ex[sx]) / np.linalg.norm(Ex[sx]),
            np.linalg.norm(Ez[sz].imag) / np.linalg.norm(Ez[sz]))


def test_fdem_vertical_dipole_wholespace_analytic():
    """the FDEM-h oracle against the closed-form field of a vertical electric dipole in a
    wholespace -- a known answer for signs, the e^{+i w t} convention, C^T MfRho C + i w MeMu and j = C h - s_e.
    At 10 kHz (skin depth 16 m) 44 % of the field in the comparison shell is inductive (imaginary)."""
    hx, hz, m, se, zsrc = ved_problem()
    sigma = 0.1
    P = OracleFDEM(m, sigma * np.ones(m.nC), MU_0 * np.ones(m.nC), "h")
    for freq, imag_min in ((1e3, 0.08), (1e4, 0.35)):
        h = P.solve(freq, se)
        e = P.j_from(freq, h, se) / sigma
        rz, rx, imag_share = ved_check(m, e, zsrc, sigma, freq)
        assert rz < 0.015 and rx < 0.015, (freq, rz, rx)
        assert imag_share > imag_min


def ved_stepoff(rho, z, zs, sigma, t):
    """E of a unit-moment vertical electric dipole in a wholespace whose DC current is switched off at t = 0:
    e_DC - e_step-on(t), Ward & Hohmann (1988) eq. 2.50 rotated to z.  Returns (E_z, E_rho); t = 0 gives the DC field."""
    from scipy.special import erfc
    zz = z - zs
    r = np.sqrt(rho ** 2 + zz ** 2)
    pre = 1.0 / (4 * np.pi * sigma * r ** 3)
    ez, er = pre * (3 * zz ** 2 / r ** 2 - 1), pre * (3 * rho * zz / r ** 2)
    if t > 0:
        x = np.sqrt(MU_0 * sigma / (4 * t)) * r
        g1 = (4 / np.sqrt(np.pi) * x ** 3 + 6 / np.sqrt(np.pi) * x) * np.exp(-x ** 2) + 3 * erfc(x)
        g2 = (4 / np.sqrt(np.pi) * x ** 3 + 2 / np.sqrt(np.pi) * x) * np.exp(-x ** 2) + erfc(x)
        ez, er = ez - pre * (zz ** 2 / r ** 2 * g1 - g2), er - pre * (rho * zz / r ** 2 * g1)
    return ez, er


def ved_stepoff_errors(m, J, times, zsrc, sigma, steps):
    gz, gx = m.gridFz, m.gridFx
    Rz, Rx = np.hypot(gz[:, 0], gz[:, 2]), np.hypot(gx[:, 0], gx[:, 2])
    sz, sx = (Rz > 12) & (Rz < 40), (Rx > 12) & (Rx < 40)
    out = []
    for n in steps:
        e = J[:, n] / sigma
        Ez = sum(ved_stepoff(gz[:, 0], gz[:, 2], zs, sigma, times[n])[0] for zs in zsrc)
        Ex = sum(ved_stepoff(gx[:, 0], gx[:, 2], zs, sigma, times[n])[1] for zs in zsrc)
        ez, ex = e[m.nFx + m.nFy:], e[:m.nFx]
        out.append((np.linalg.norm(ez[sz] - Ez[sz]) / np.linalg.norm(Ez[sz]), np.linalg.norm(ex[sx] - Ex[sx]) / np.linalg.norm(Ex[sx])))
    return out


VED_DTS = [(2e-7, 25), (1e-6, 20), (5e-6, 16)]          # 61 steps to t = 1.05e-4 s


def test_tdem_vertical_dipole_stepoff_analytic():
    """the TDEM-j oracle (DC initial state + backward Euler) against the closed-form step-off
    response of a vertical electric dipole in a wholespace; halving the time steps halves the error (first order)."""
    hx, hz, m, se, zsrc = ved_problem()
    sigma = 0.1
    errs = {}
    for tag, scale in (("coarse", 1), ("fine", 2)):
        dts = mesh_tensor([(dt / scale, n * scale) for dt, n in VED_DTS])
        times = np.r_[0.0, np.cumsum(dts)]
        J = OracleTDEM(m, sigma * np.ones(m.nC), MU_0 * np.ones(m.nC), dts).fields(se)
        errs[tag] = ved_stepoff_errors(m, J, times, zsrc, sigma, [0, 10 * scale, 25 * scale, 35 * scale, 45 * scale])
    assert errs["coarse"][0][0] < 0.015 and errs["coarse"][0][1] < 0.015        # t = 0: the DC dipole
    for (cz, cx), (fz, fx) in zip(errs["coarse"][1:], errs["fine"][1:]):
        assert cz < 0.05 and cx < 0.09, errs
        assert fz < 0.7 * cz and fx < 0.7 * cx, errs                           # first-order convergence in dt


def test_dc_3d_off_axis_dipole_analytic():
    """3-D known answer for the DC oracle: a current dipole 8.5 m OFF the mesh axis in a wholespace against
    phi = I / (4 pi sigma) (1/Ra - 1/Rb).  The region around the axis is where the axis conventions act
    (SURVEY A.4 OQ1): axis_weight_face = 0.5 (default) is closer to the closed form than 1.0."""
    nth = 16
    hx = mesh_tensor([(1.0, 24), (1.0, 14, 1.4)])
    hz = mesh_tensor([(1.0, 14, -1.4), (1.0, 48), (1.0, 14, 1.4)])
    hy = np.ones(nth) * 2 * np.pi / nth
    sigma = 0.1
    th0 = (nth // 4 + 0.5) * 2 * np.pi / nth
    a, b = np.array([[8.5, th0, 6.5]]), np.array([[8.5, th0, -6.5]])
    err = {}
    for awf in (0.5, 1.0):
        m = OracleCylMesh([hx, hy, hz], [0, 0, -hz.sum() / 2], axis_weight_face=awf)
        P = OracleDC(m, sigma * np.ones(m.nC))
        phi = P.fields(a, b)[:, 0]
        q = P.getRHS(a, b)[:, 0]
        cc = m.gridCC
        X = np.c_[cc[:, 0] * np.cos(cc[:, 1]), cc[:, 0] * np.sin(cc[:, 1]), cc[:, 2]]
        Ra = np.maximum(np.linalg.norm(X - X[np.argmax(q)], axis=1), 1e-9)
        Rb = np.maximum(np.linalg.norm(X - X[np.argmin(q)], axis=1), 1e-9)
        ana = (1 / Ra - 1 / Rb) / (4 * np.pi * sigma)
        axis = (cc[:, 0] < 3) & (np.abs(cc[:, 2]) < 12) & (Ra > 4) & (Rb > 4)
        shell = (Ra > 4) & (Rb > 4) & (Ra < 18) & (Rb < 18)
        err[awf] = (np.linalg.norm(phi[axis] - ana[axis]) / np.linalg.norm(ana[axis]),
                    np.linalg.norm(phi[shell] - ana[shell]) / np.linalg.norm(ana[shell]))
    assert err[0.5][0] < 0.01 and err[0.5][1] < 0.06, err
    assert err[1.0][0] > 1.3 * err[0.5][0], err


def test_exact_assembly_and_data_sensitivity():
    """oracle.exact_mesh: the same operators assembled in long double.  (1) its matrices agree with the fp64 ones to
    rounding; (2) yet on a casing model the fp64-assembled DC system's own (refined) solution is ~1e-6 away from the
    exactly assembled one -- the entry roundings act as spurious leaks next to the 1e13 contrast; (3) in 3-D the plain
    L2 norm of j is at the mercy of ONE fp64 rounding of the right-hand side at low frequency (1/(w mu) amplification of
    a null-space component) while the energy norm and the casing currents are not: this fixes which quantities a
    1e-6 parity contract can be asserted on (tests/test_gpu_parity.py)."""
    from conftest import CASING, small_casing_mesh
    from oracle.cyl_oracle import (OracleCylMesh, OracleDC, OracleFDEM, paint_casing_in_halfspace, exact_mesh, solve_refined,
                                   energy_rel, casing_currents, rhs_rounding_sensitivity)
    import oracle.cyl_oracle as oc

    def rel(a, b):
        a, b = np.asarray(a, dtype=complex), np.asarray(b, dtype=complex)
        return float(np.linalg.norm(a - b) / np.linalg.norm(b))
    hx, hy, hz, z0 = small_casing_mesh(4)
    om = OracleCylMesh([hx, hy, hz], [0, 0, z0])
    ox = exact_mesh(om)
    sig, mu = paint_casing_in_halfspace(om, **CASING)
    Pd, Px = OracleDC(om, sig), OracleDC(ox, sig)
    A64, Ald = Pd.getA(), Px.getA()
    assert Ald.dtype == np.longdouble
    assert abs(A64 - Ald.astype(float)).max() <= 4e-16 * abs(A64).max()
    a, b = np.array([[0.03, np.pi / 4, -0.25]]), np.array([[2.0, 5 * np.pi / 4, -0.25]])
    x_asm, x_ex = Pd.fields(a, b, refine=True), Px.fields(a, b, refine=True)
    assert oc.LAST_REFINE_RESIDUAL < 1e-12
    assert 1e-8 < rel(x_asm, x_ex) < 5e-6                  # ~6e-7: what an fp64-assembled matrix costs
    g = om.gridFz
    w = np.zeros(om.nFz)
    w[(g[:, 0] < om.hx.min()) & (g[:, 2] > -2.6) & (g[:, 2] < 0.1) & (g[:, 1] < hy[0])] = 1.0
    se = np.r_[np.zeros(om.nFx + om.nFy), w] / om.area
    PX = OracleFDEM(ox, sig, mu, "h")
    rho = np.asarray(PX.MfRho.diagonal(), dtype=float)
    freq = 1.0
    A, rhs = PX.getA(freq), PX.getRHS(freq, se)
    x1 = solve_refined(A, rhs, iters=12)[0]
    assert oc.LAST_REFINE_RESIDUAL < 1e-12
    x2 = solve_refined(A, np.asarray(np.asarray(rhs, dtype=complex), dtype=np.clongdouble), iters=12)[0]
    j1, j2 = PX.j_from(freq, x1, se), PX.j_from(freq, x2, se)
    cas = (CASING["casing_a"], CASING["casing_b"], CASING["casing_z"])
    i1, i2 = casing_currents(np.asarray(j1, dtype=complex), om, *cas)["z"][1], casing_currents(np.asarray(j2, dtype=complex), om, *cas)["z"][1]
    assert rel(j2, j1) > 1e-6                               # plain L2: not determined by the fp64 data
    assert energy_rel(j2, j1, rho) < 1e-8 and rel(i2, i1) < 1e-12   # energy norm, casing current: determined
    assert abs(rhs_rounding_sensitivity(A, rhs, lambda x: PX.j_from(freq, x, se)) - rel(j2, j1)) < 1e-3 * rel(j2, j1)
