"""Numpy check of the theta-mode algebra used by the half-mode preconditioner:
K_m Hermitian; T_m^* K_m T_m real symmetric; K''_{-m} = S K''_m S."""
import sys, os
import numpy as np, scipy.sparse as sp
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle.cyl_oracle import OracleCylMesh, mesh_tensor

nx, nth, nz = 5, 8, 4
rng = np.random.default_rng(0)
hx = rng.uniform(0.5, 2, nx); hz = rng.uniform(0.5, 2, nz); hy = np.ones(nth) * 2 * np.pi / nth
M = OracleCylMesh([hx, hy, hz], [0, 0, 0])
# theta-invariant model
s2 = rng.uniform(0.1, 10, (nx, nz)); m2 = rng.uniform(1, 5, (nx, nz))
sig = np.repeat(s2[:, None, :], nth, axis=1).ravel(order="F"); mu = np.repeat(m2[:, None, :], nth, axis=1).ravel(order="F")
C = M.edgeCurl
MfRho = M.getFaceInnerProduct(sig, invProp=True); MeMu = M.getEdgeInnerProduct(mu)
MeMuI = M.getEdgeInnerProduct(mu, invMat=True)
Kedge = (C.T @ MfRho @ C).toarray()
Kface = (MfRho @ C @ MeMuI @ C.T @ MfRho).toarray()
nxy = nx * nth

def edge_plan():
    L = 3 * nx + 1; n2 = L * (nz + 1)
    mp = -np.ones(n2, dtype=int); js = np.full(n2, nx); comp = np.zeros(n2, dtype=int)
    for k in range(nz + 1):
        for i in range(nx):
            q = k * L + 3 * i
            mp[q] = i + nxy * k; comp[q] = 0
            mp[q + 1] = M.nEx + i + nxy * k; comp[q + 1] = 1
            if k < nz:
                mp[q + 2] = M.nEx + M.nEy + k * (nxy + 1) + 1 + i; comp[q + 2] = 2
        if k < nz:
            q = k * L + 3 * nx; mp[q] = M.nEx + M.nEy + k * (nxy + 1); js[q] = 0; comp[q] = 3
    return mp, js, comp

def face_plan():
    L = 3 * nx; n2 = L * (nz + 1)
    mp = -np.ones(n2, dtype=int); js = np.full(n2, nx); comp = np.zeros(n2, dtype=int)
    for k in range(nz + 1):
        for i in range(nx):
            q = k * L + 3 * i
            mp[q] = M.nFx + M.nFy + i + nxy * k; comp[q] = 2
            if k < nz:
                mp[q + 1] = i + nxy * k; comp[q + 1] = 0
                mp[q + 2] = M.nFx + i + nxy * k; comp[q + 2] = 1
    return mp, js, comp

def mode_mats(K, mp, js):
    n2 = len(mp); out = np.zeros((nth, n2, n2), dtype=complex)
    for q in range(n2):
        if mp[q] < 0: continue
        col = K[:, mp[q]]          # unit vector at (q, j = 0)
        for r in range(n2):
            if mp[r] < 0: continue
            if js[r] == 0:
                out[0, r, q] = col[mp[r]] * (1 if js[q] else 1)   # axis row: mode 0 only ... see below
                continue
            vals = col[mp[r] + np.arange(nth) * js[r]]
            out[:, r, q] = np.fft.fft(vals)       # sum_j v_j e^{-2 pi i j m / nth}
    return out

for name, K, plan, stag in (("edge", Kedge, edge_plan, 1), ("face", Kface, face_plan, 1)):
    mp, js, comp = plan()
    Km = mode_mats(K, mp, js)
    ok = (mp >= 0)
    for m in range(nth):
        A = Km[m][np.ix_(ok, ok)]
        cm = comp[ok]; jj = js[ok]
        if m > 0:      # axis unknown lives in mode 0 only
            keep = jj != 0; A = A[np.ix_(keep, keep)]; cm = cm[keep]
        herm = np.abs(A - A.conj().T).max() / np.abs(A).max()
        phi = 2 * np.pi * m / nth
        best = None
        for sgn in (+1, -1):
            for ii in (1j, -1j):
                c = ii * np.exp(sgn * 0.5j * phi)
                T = np.where(cm == stag, c, 1.0)
                App = (T.conj()[:, None] * A) * T[None, :]
                im = np.abs(App.imag).max() / np.abs(App).max()
                if best is None or im < best[0]: best = (im, sgn, ii, App)
        print(name, "m", m, "hermitian defect %.1e" % herm, "imag after T %.1e" % best[0], "sgn", best[1], "i", best[2])
    # -m relation with a fixed convention
    def Kpp(m, sgn, ii):
        A = Km[m][np.ix_(ok, ok)]; cm = comp[ok]; jj = js[ok]
        if m > 0: keep = jj != 0; A = A[np.ix_(keep, keep)]; cm = cm[keep]
        phi = 2 * np.pi * m / nth
        T = np.where(cm == stag, ii * np.exp(sgn * 0.5j * phi), 1.0)
        return (T.conj()[:, None] * A) * T[None, :], cm
    for sgn in (+1, -1):
        for m in range(1, nth // 2):
            A, cm = Kpp(m, sgn, 1j); B, _ = Kpp(nth - m, sgn, 1j)
            S = np.where(cm == stag, -1.0, 1.0)
            d = np.abs(B - (S[:, None] * A) * S[None, :]).max() / np.abs(A).max()
            print(name, "sgn", sgn, "m", m, "|K''(-m) - S K''(m) S| %.1e" % d, " |K''(-m)-K''(m)| %.1e" % (np.abs(B - A).max() / np.abs(A).max()))
