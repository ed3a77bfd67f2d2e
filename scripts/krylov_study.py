# CPU study (test infrastructure): MINRES + block-diagonal vs GMRES + block-triangular preconditioning of the
# three-field system (pe_k3.h), blocks exact or spectrally equivalent with a prescribed condition number.
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, scipy.linalg as sla, scipy.sparse as sp, scipy.sparse.linalg as spla
from oracle_lib import *
nref = int(sys.argv[1]) if len(sys.argv) > 1 else 4
lam = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
kap_u = float(sys.argv[3]) if len(sys.argv) > 3 else 3.0
kap_p = float(sys.argv[4]) if len(sys.argv) > 4 else 2.3
mu, alpha, c0, dt = 1.0, 1.0, 0.0, 0.1
N = 1 << nref
def mats(l, a):
    o = Oracle(BR1_WG, nref); o.set_material(l, mu, a, c0, dt); o.set_case(CASE_COSCOS)
    v, r, _ = o.assemble(0.1)
    return o, o.csr(v).toarray(), r
o, A, r = mats(lam, alpha)
_, A0, _ = mats(0.0, 1.0)
nu, npi, npf = o.n_u, o.n_pint, o.n_pface
Amu = A0[:nu, :nu]; B = A0[nu:nu + npi, :nu]; C = A0[nu:, nu:]
W = np.full(npi, 1.0 / N / N)
# Dirichlet rows of u: K3 keeps A0's diagonal, rhs rescaled
du, d0 = np.diag(A)[:nu], np.diag(Amu)
con = np.array([np.count_nonzero(A[i, :]) == 1 for i in range(nu)])
bu = r[:nu].copy(); bu[con] *= d0[con] / du[con]
n3 = nu + npi + npi + npf
K = np.zeros((n3, n3))
iu, ix, ip = slice(0, nu), slice(nu, nu + npi), slice(nu + npi, n3)
K[iu, iu] = Amu; K[iu, ix] = -B.T; K[ix, iu] = -B
Bc = B.copy(); 
K[ix, ix] = -np.diag(W / lam)
Wp = np.zeros((npi, npi + npf)); Wp[:, :npi] = np.diag(W)
K[ix, ip] = alpha / lam * Wp; K[ip, ix] = alpha / lam * Wp.T
Cp = C.copy(); Cp[:npi, :npi] += alpha * alpha / lam * np.diag(W)
K[ip, ip] = -Cp
# constrained u rows must not couple to xi
K[np.where(con)[0], nu:] = 0; K[nu:, np.where(con)[0]] = 0
b3 = np.concatenate([bu, np.zeros(npi), -r[nu:]])
x_ref = np.linalg.solve(K, b3)
x2 = np.linalg.solve(A, r)
print("three-field vs two-field solution:", np.linalg.norm(x_ref[:nu] - x2[:nu]) / np.linalg.norm(x2[:nu]),
      np.linalg.norm(x_ref[nu + npi:] - x2[nu:]) / np.linalg.norm(x2[nu:]))
rng = np.random.default_rng(0)
def equiv_inverse(M, kappa):
    """SPD P with spectrum(P^-1 M) spread over [1/kappa, 1] (random eigenbasis)"""
    w, Q = np.linalg.eigh(M)
    Mh = (Q * np.sqrt(w)) @ Q.T; Mih = (Q / np.sqrt(w)) @ Q.T
    n = M.shape[0]
    R, _ = np.linalg.qr(rng.normal(size=(n, n)))
    d = rng.uniform(1.0 / kappa, 1.0, size=n); d[0], d[1] = 1.0 / kappa, 1.0
    return Mih @ (R * d) @ R.T @ Mih
Pp_mat = Cp
for exact in (True, False):
    Pu_inv = np.linalg.inv(Amu) if exact else equiv_inverse(Amu, kap_u)
    Pp_inv = np.linalg.inv(Pp_mat) if exact else equiv_inverse(Pp_mat, kap_p)
    px = 1.0 / ((1 / lam + 1 / (2 * mu)) * W)
    def diag_prec(rv):
        return np.concatenate([Pu_inv @ rv[iu], px * rv[ix], Pp_inv @ rv[ip]])
    variant = "full"
    def tri_prec(rv):
        if variant == "full":       # upper triangular, all couplings
            zp = -(Pp_inv @ rv[ip])
            zx = -px * (rv[ix] - alpha / lam * (Wp @ zp))
            zu = Pu_inv @ (rv[iu] + B.T @ zx * (~con))
        elif variant == "u-xi":     # upper triangular, u-xi coupling only (p pipeline independent of the u pipeline)
            zp = -(Pp_inv @ rv[ip])
            zx = -px * rv[ix]
            zu = Pu_inv @ (rv[iu] + B.T @ zx * (~con))
        elif variant == "lower":    # lower triangular, all couplings
            zu = Pu_inv @ rv[iu]
            zx = -px * (rv[ix] + B @ zu)
            zp = -(Pp_inv @ (rv[ip] - alpha / lam * (Wp.T @ zx)))
        elif variant == "lower-u-xi":
            zu = Pu_inv @ rv[iu]
            zx = -px * (rv[ix] + B @ zu)
            zp = -(Pp_inv @ rv[ip])
        return np.concatenate([zu, zx, zp])
    its = [0]
    def cb(x): its[0] += 1
    M = spla.LinearOperator((n3, n3), matvec=diag_prec)
    x, info = spla.minres(sp.csr_matrix(K), b3, M=M, rtol=1e-10, maxiter=2000, callback=cb)
    print("exact" if exact else "kappa %.1f/%.1f" % (kap_u, kap_p), "MINRES block-diag its", its[0], "err", np.linalg.norm(x - x_ref) / np.linalg.norm(x_ref))
    res = []
    def cb2(rk): res.append(rk)
    # right-preconditioned GMRES: K P^-1 y = b, x = P^-1 y
    Kop = spla.LinearOperator((n3, n3), matvec=lambda y: K @ tri_prec(y))
    for variant in ("full", "u-xi", "lower", "lower-u-xi"):
        for m in (200, 10):
            res = []
            y, info = spla.gmres(Kop, b3, rtol=1e-10, restart=m, maxiter=300, callback=cb2, callback_type="pr_norm")
            x = tri_prec(y)
            print("   GMRES(%d) %-10s its" % (m, variant), len(res), "err %.1e" % (np.linalg.norm(x - x_ref) / np.linalg.norm(x_ref)))
