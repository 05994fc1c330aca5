"""Host restatement of the two-front ("twisted") elimination of cs_band.cu (band_plan_twist, k_form_twisted,
k_twist_combine, the three substitution phases of band_solve_t) in dense numpy, checked against a direct solve.
It pins the index conventions the CUDA kernels use: T = [dummies | levels below the separator | separator],
B = [dummies | levels above it, reversed | separator, reversed], B's separator block zero, Schur complements added
mirrored, separator right-hand side travelling with T."""
import numpy as np
import pytest


def twist_geometry(n2, lrow, nb=16, even=False):
    levels = n2 // lrow
    ks = levels // 2
    nt, w = lrow * ks, lrow
    nbot = n2 - nt - w
    M = (max(nt, nbot) + nb - 1) // nb * nb
    n2h = M + w
    if even and n2h % 2:
        n2h += 1
    return dict(n2=n2, nt=nt, w=w, M=M, padT=M - nt, padB=M - nbot, n2h=n2h)


def tw_orig(g, half, i):
    if half == 0:
        o = i - g["padT"]
        return -1 if (o < 0 or o >= g["nt"] + g["w"]) else o
    o = g["n2"] - 1 - (i - g["padB"])
    return -1 if (i < g["padB"] or o < g["nt"]) else o


def form_halves(A, g):
    n2h, nt, w = g["n2h"], g["nt"], g["w"]
    out = []
    for half in (0, 1):
        H = np.zeros((n2h, n2h), dtype=A.dtype)
        orig = [tw_orig(g, half, i) for i in range(n2h)]
        for j, oj in enumerate(orig):
            if oj < 0:
                H[j, j] = 1.0
                continue
            for i, oi in enumerate(orig):
                if oi < 0:
                    continue
                if half == 1 and nt <= oi < nt + w and nt <= oj < nt + w:
                    continue                                   # the separator block goes to T
                H[i, j] = A[oi, oj]
        out.append(H)
    return out, [[tw_orig(g, h, i) for i in range(n2h)] for h in (0, 1)]


def eliminate(H, stop):
    """LU without pivoting of the leading `stop` unknowns, in place: L below (unit), U on and above, Schur complement trailing"""
    H = H.copy()
    for k in range(stop):
        H[k + 1:, k] /= H[k, k]
        H[k + 1:, k + 1:] -= np.outer(H[k + 1:, k], H[k, k + 1:])
    return H


@pytest.mark.parametrize("lrow,levels,pad,cplx", [(7, 9, 0, False), (10, 12, 1, True), (5, 8, 0, True)])
def test_two_front_elimination_matches_direct_solve(lrow, levels, pad, cplx):
    rng = np.random.default_rng(lrow * 100 + levels)
    n2 = lrow * levels + pad
    A = np.zeros((n2, n2), dtype=complex if cplx else float)
    for k in range(levels):                                    # block tridiagonal in the levels, symmetric, diagonally dominant
        s = slice(k * lrow, (k + 1) * lrow)
        D = rng.standard_normal((lrow, lrow))
        A[s, s] = D + D.T + 6 * lrow * np.eye(lrow)
        if k + 1 < levels:
            s1 = slice((k + 1) * lrow, (k + 2) * lrow)
            C = rng.standard_normal((lrow, lrow))
            A[s1, s], A[s, s1] = C, C.T
    if cplx:
        A = A + 1j * np.diag(rng.uniform(0.1, 1.0, n2))        # complex symmetric shift
    for q in range(lrow * levels, n2):
        A[q, q] = 1.0                                          # the padding dummy of the plan
    b = rng.standard_normal(n2) + (1j * rng.standard_normal(n2) if cplx else 0)
    g = twist_geometry(n2, lrow, even=bool(pad))
    M, w, n2h = g["M"], g["w"], g["n2h"]
    assert M % 16 == 0 and g["padT"] >= 0 and g["padB"] >= 0
    (T, B), orig = form_halves(A, g)
    # both fronts up to the separator, Schur complements combined (B's mirrored), separator factored on T
    Tf, Bf = eliminate(T, M), eliminate(B, M)
    Tf[M:M + w, M:M + w] += Bf[M:M + w, M:M + w][::-1, ::-1]
    Tf = np.block([[Tf[:M, :M], Tf[:M, M:]], [Tf[M:, :M], eliminate(Tf[M:, M:], n2h - M)]])
    # right-hand sides: the separator travels with T
    bt = np.array([b[o] if o >= 0 else 0 for o in orig[0]], dtype=A.dtype)
    bb = np.array([b[o] if (o >= 0 and i < M) else 0 for i, o in enumerate(orig[1])], dtype=A.dtype)

    def forward(F, y, s0, s1):
        for k in range(s0, s1):
            y[k + 1:] -= F[k + 1:, k] * y[k]

    def backward(F, y, s1, s0, given=0):
        # the first `given` steps take the block solution as it is IN THE VECTOR (the kernel reads it from global memory:
        # its shared-memory window of those rows is overwritten by the updates of the steps before)
        xs = y.copy()
        for k in range(s1 - 1, s0 - 1, -1):
            if k < n2h - given:
                y[k] /= F[k, k]
                xk = y[k]
            else:
                xk = xs[k]
            y[:k] -= F[:k, k] * xk
        y[n2h - given:] = xs[n2h - given:]
    forward(Tf, bt, 0, M); forward(Bf, bb, 0, M)               # phase 1 (the partial sums of the separator rows stay in y)
    bt[M:M + w] += bb[M:M + w][::-1]
    sep = bt[M:].copy()                                        # phase 2: the separator on T
    forward(Tf[M:, M:], sep, 0, n2h - M)
    backward_sep = sep.copy()
    for k in range(n2h - M - 1, -1, -1):
        backward_sep[k] /= Tf[M + k, M + k]
        backward_sep[:k] -= Tf[M:M + k, M + k] * backward_sep[k]
    bt[M:] = backward_sep
    bb[M:M + w] = bt[M:M + w][::-1]
    backward(Tf, bt, n2h, 0, given=n2h - M)                    # phase 3: separator steps in "given" mode
    backward(Bf, bb, n2h, 0, given=n2h - M)
    x = np.zeros(n2, dtype=A.dtype)
    for i, o in enumerate(orig[0]):
        if o >= 0:
            x[o] = bt[i]
    for i, o in enumerate(orig[1]):
        if o >= 0 and i < M:
            x[o] = bb[i]
    xref = np.linalg.solve(A, b)
    assert np.linalg.norm(x - xref) / np.linalg.norm(xref) < 1e-11
