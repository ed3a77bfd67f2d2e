"""Known-answer vectors for the BR1-WG (17 x 17) and CGQ1-WG (13 x 13) local matrices on the rectangle
[0, hx] x [0, hy] in EXACT rational arithmetic (sympy), following PoroElasBR1WG_new.cc:811-925 and
PoroElasCGQ1_WG.cc:1194-1358 term by term -- independent of the oracle's quadrature, mapping and Gauss-Jordan code.
Parameters of SURVEY 8(c): hx = 1/2, hy = 1/4, lambda = 3, mu = 7/10, alpha = 93/100, c0 = 1/10, dt = 1/20, K = 1/1000.
Writes tests/golden/biot_exact_rectangle.npz.   Run:  python scripts/derive_biot_kat.py"""
import os

import numpy as np
import sympy as sp

R = sp.Rational
x, y = sp.symbols("x y")
hx, hy = R(1, 2), R(1, 4)
lam, mu, alpha, c0, dt, K = R(3), R(7, 10), R(93, 100), R(1, 10), R(1, 20), R(1, 1000)
xi, eta = x / hx, y / hy  # reference coordinates


def integ(f):
    return sp.integrate(sp.integrate(f, (x, 0, hx)), (y, 0, hy))


# Q1 nodal functions, deal.II vertex order (0,0) (1,0) (0,1) (1,1)
N = [(1 - xi) * (1 - eta), xi * (1 - eta), (1 - xi) * eta, xi * eta]
# Bernardi-Raugel face bubbles: reference value times the REFERENCE unit vector (mapping_none)
bub = [((1 - xi) * 4 * eta * (1 - eta), 0), (xi * 4 * eta * (1 - eta), 0), (0, 4 * xi * (1 - xi) * (1 - eta)), (0, 4 * xi * (1 - xi) * eta)]
# RT0 on the rectangle (Piola map of ((1-xi),0), (xi,0), (0,1-eta), (0,eta)):  w = J w_ref / det J
w_rt = [((1 - xi) / hy, 0), (xi / hy, 0), (0, (1 - eta) / hx), (0, eta / hx)]
area = hx * hy


def eps(u):
    gxx, gxy, gyx, gyy = sp.diff(u[0], x), sp.diff(u[0], y), sp.diff(u[1], x), sp.diff(u[1], y)
    return (gxx, (gxy + gyx) / 2, (gxy + gyx) / 2, gyy)


def div(u):
    return sp.diff(u[0], x) + sp.diff(u[1], y)


def local_matrix(element):
    if element == 0:  # (v0x v0y .. v3x v3y | b0 pf0 | b1 pf1 | b2 pf2 | b3 pf3 | pint)
        n = 17
        u = {2 * v + c: tuple(N[v] if d == c else 0 for d in range(2)) for v in range(4) for c in range(2)}
        for f in range(4):
            u[8 + 2 * f] = bub[f]
        pface = {9 + 2 * f: f for f in range(4)}
        pint = 16
    else:  # (v0x .. v3y | pf0 pf1 pf2 pf3 | pint)
        n = 13
        u = {2 * v + c: tuple(N[v] if d == c else 0 for d in range(2)) for v in range(4) for c in range(2)}
        pface = {8 + f: f for f in range(4)}
        pint = 12
    A = sp.zeros(n, n)
    # ---- weak gradients: M = RT0 Gram matrix, F = -int phi_int div w + sum_f int_f phi_face w.n, C = F M^-1
    M = sp.Matrix(4, 4, lambda k, l: integ(w_rt[k][0] * w_rt[l][0] + w_rt[k][1] * w_rt[l][1]))
    F = sp.zeros(n, 4)
    for k in range(4):
        F[pint, k] = -integ(div(w_rt[k]))
    # faces 0: x=0 (n=-x), 1: x=hx (+x), 2: y=0 (-y), 3: y=hy (+y); FE_FaceQ(0) = 1 on its own face
    for i, f in pface.items():
        for k in range(4):
            if f == 0:
                F[i, k] = sp.integrate((-w_rt[k][0]).subs(x, 0) if w_rt[k][0] != 0 else 0, (y, 0, hy))
            elif f == 1:
                F[i, k] = sp.integrate((w_rt[k][0]).subs(x, hx) if w_rt[k][0] != 0 else 0, (y, 0, hy))
            elif f == 2:
                F[i, k] = sp.integrate((-w_rt[k][1]).subs(y, 0) if w_rt[k][1] != 0 else 0, (x, 0, hx))
            else:
                F[i, k] = sp.integrate((w_rt[k][1]).subs(y, hy) if w_rt[k][1] != 0 else 0, (x, 0, hx))
    C = F * M.inv()
    A += dt * K * C * M * C.T  # BR1new:879-901 / CGQ1:1262-1284 with K = k I
    # ---- elasticity, coupling, storage
    if element == 0:
        dbar = {i: integ(div(ui)) / area for i, ui in u.items()}  # cell-averaged divergence, BR1new:867-876
        for i, ui in u.items():
            for j, uj in u.items():
                ei, ej = eps(ui), eps(uj)
                A[i, j] += integ(2 * mu * sum(a * b for a, b in zip(ei, ej))) + lam * dbar[i] * dbar[j] * area
            A[i, pint] += -alpha * dbar[i] * area
            A[pint, i] += alpha * dbar[i] * area
        A[pint, pint] += c0 * area
    else:
        dc = {i: div(ui).subs({x: hx / 2, y: hy / 2}) for i, ui in u.items()}  # QGauss(1): cell centre, CGQ1:1341-1358
        for i, ui in u.items():
            for j, uj in u.items():
                ei, ej = eps(ui), eps(uj)
                A[i, j] += integ(2 * mu * sum(a * b for a, b in zip(ei, ej))) + lam * dc[i] * dc[j] * area
            A[i, pint] += -alpha * dc[i] * area
            A[pint, i] += alpha * dc[i] * area
        A[pint, pint] += 2 * c0 * area  # the storage term is integrated twice (CGQ1:1335 and :1352)
    return A


if __name__ == "__main__":
    out = {}
    for el, name in ((0, "br1"), (1, "cgq1")):
        A = local_matrix(el)
        out[name] = np.array(A.evalf(30).tolist(), dtype=float)
        print(name, "trace", A.trace(), "=", float(A.trace()), " sum", sum(A), " A[pint,pint]", A[A.rows - 1, A.rows - 1])
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "biot_exact_rectangle.npz")
    np.savez_compressed(path, params=np.array([0.5, 0.25, 3.0, 0.7, 0.93, 0.1, 0.05, 1e-3]), **out)
