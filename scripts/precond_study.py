# CPU study (test infrastructure): spectra of candidate block preconditioners on the oracle's matrix.
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, scipy.linalg as sla
from oracle_lib import *
nref = int(sys.argv[1]) if len(sys.argv) > 1 else 4
N = 1 << nref; V = N + 1
o = Oracle(BR1_WG, nref); o.set_material(1, 1, 1, 0, 0.1); o.set_case(CASE_COSCOS)
v, r, _ = o.assemble(0.1)
A = o.csr(v).toarray()
nu, npi, npf = o.n_u, o.n_pint, o.n_pface
cd = o.cell_dofs(); vx, vy, cv, fk = o.mesh()
Auu = A[:nu, :nu]; C = A[nu:, nu:]; B = A[nu:, :nu]
# constrained dofs: rows with only diagonal
con = np.array([np.count_nonzero(A[i]) == 1 for i in range(A.shape[0])])
vdofs = np.zeros((V * V, 2), int)
for c in range(o.n_cells):
    for k in range(4):
        vdofs[cv[c, k]] = cd[c, 2 * k], cd[c, 2 * k + 1]
isvert = np.zeros(nu, bool); isvert[vdofs.ravel()] = True
free_u = ~con[:nu]
def cond(Pinv, M, free):
    Mf = M[np.ix_(free, free)]; Pf = Pinv[np.ix_(free, free)]
    ev = np.sort(np.real(sla.eigvals(Pf @ Mf)))
    return ev[0], ev[-1], ev[-1] / ev[0]
D = np.diag(1 / np.diag(Auu))
# exact aux solve on free vertex dofs
fv = np.where(isvert & free_u)[0]
Pi = np.zeros((nu, len(fv))); Pi[fv, np.arange(len(fv))] = 1
Avv_inv = np.linalg.inv(Pi.T @ Auu @ Pi)
P1 = D + Pi @ Avv_inv @ Pi.T
print("u: Jacobi(all)+exact vertex:", cond(P1, Auu, free_u))
Db = D.copy(); Db[isvert, isvert] = 0
P2 = Db + Pi @ Avv_inv @ Pi.T
print("u: Jacobi(bubbles only)+exact vertex:", cond(P2, Auu, free_u))
# pressure block
S = C  # c0 M + dt S (c0 = 0)
gamma = 1. / 3
pint_of = {}; 
Pp = np.zeros((npi + npf, V * V))
for c in range(o.n_cells):
    Pp[cd[c, 16] - nu, cv[c]] = 0.25
    for f, (a, b) in enumerate([(0, 2), (1, 3), (0, 1), (2, 3)]):
        Pp[cd[c, 9 + 2 * f] - nu, cv[c, a]] = 0.5; Pp[cd[c, 9 + 2 * f] - nu, cv[c, b]] = 0.5
# Schur with exact A
Sch = S + B @ np.linalg.inv(Auu) @ (-A[:nu, nu:])   # A[:nu,nu:] = -alpha B^T
free_p = ~con[nu:]
bnd_v = np.zeros(V * V, bool)
for c in range(o.n_cells):
    for f, (a, b) in enumerate([(0, 2), (1, 3), (0, 1), (2, 3)]):
        if fk[c, f]: bnd_v[cv[c, a]] = bnd_v[cv[c, b]] = True
Ppf = Pp[:, ~bnd_v]
Mint = np.zeros(npi + npf); Mint[:npi] = 1.0 / N / N
Shat = S + gamma * np.diag(Mint)
Aaux_inv = np.linalg.inv(Ppf.T @ Shat @ Ppf)
Dp = np.diag(1 / np.diag(Shat))
PS = Dp + Ppf @ Aaux_inv @ Ppf.T
print("p: kappa(Shat^-1 Schur) exact Shat:", cond(np.linalg.inv(Shat[np.ix_(free_p, free_p)]) if False else np.linalg.pinv(Shat * np.outer(free_p, free_p) + np.diag(~free_p * 1.0)), Sch, free_p))
print("p: Jacobi+exact aux vs Shat:", cond(PS, Shat, free_p))
print("p: Jacobi+exact aux vs Schur:", cond(PS, Sch, free_p))
# eliminate p_int exactly, then Jacobi + aux on faces
Sii = Shat[:npi, :npi]; Sif = Shat[:npi, npi:]; Sff = Shat[npi:, npi:]
F = Sff - Sif.T @ np.diag(1 / np.diag(Sii)) @ Sif
ff = free_p[npi:]
Pf = Pp[npi:][:, ~bnd_v]
PF = np.diag(1 / np.diag(F)) + Pf @ np.linalg.inv(Pf.T @ F @ Pf) @ Pf.T
print("faces: Jacobi+exact aux vs F:", cond(PF, F, ff))

# ---- variants on the u block
bub = np.where(~isvert & free_u)[0]
Abb = Auu[np.ix_(bub, bub)]; Abv = Auu[np.ix_(bub, fv)]; Avb = Abv.T
Dbi = np.diag(1 / np.diag(Abb))
def P_mult(r):
    rb, rv = r[bub], r[fv]
    y1 = Dbi @ rb
    xv = Avv_inv @ (rv - Avb @ y1)
    y2 = y1 + Dbi @ (rb - Abb @ y1 - Abv @ xv)
    out = np.zeros_like(r); out[bub] = y2; out[fv] = xv
    return out
n = nu
Pm = np.column_stack([P_mult(e) for e in np.eye(n)])
print("u: sym multiplicative bubble-Jacobi / exact vertex:", cond(Pm + np.diag(con[:nu] * 1.0), Auu, free_u), "sym err", np.abs(Pm - Pm.T).max())
# exact bubble block solve instead of Jacobi
Abbi = np.linalg.inv(Abb)
def P_mult2(r):
    rb, rv = r[bub], r[fv]
    y1 = Abbi @ rb
    xv = Avv_inv @ (rv - Avb @ y1)
    y2 = y1 + Abbi @ (rb - Abb @ y1 - Abv @ xv)
    out = np.zeros_like(r); out[bub] = y2; out[fv] = xv
    return out
Pm2 = np.column_stack([P_mult2(e) for e in np.eye(n)])
print("u: sym multiplicative exact-bubble / exact vertex:", cond(Pm2, Auu, free_u))
print("cond(Abb), cond(Dbi Abb):", np.linalg.cond(Abb), cond(Dbi, Abb, np.ones(len(bub), bool)))
# additive with exact bubble block
Pa = np.zeros((nu, nu)); Pa[np.ix_(bub, bub)] = Abbi; Pa += Pi @ Avv_inv @ Pi.T
print("u: additive exact-bubble + exact vertex:", cond(Pa, Auu, free_u))
# vertex space + Schur-corrected: A_vv - A_vb Abb^-1 A_bv  (static condensation of bubbles)

# ---- full coupled MINRES iteration counts with the candidate block preconditioner
import scipy.sparse.linalg as spla, scipy.sparse as sp
n = A.shape[0]
Dsign = np.ones(n); Dsign[nu:] = -1
K = Dsign[:, None] * A; rhs = Dsign * r
Dii = np.diag(Shat)[:npi]
Fd = np.diag(F)
Faux_inv = np.linalg.inv(Pf.T @ F @ Pf)
conu = np.where(con[:nu])[0]; conp = con[nu:]
def Pfull(rv, variant):
    out = np.zeros(n)
    ru, rp = rv[:nu], rv[nu:]
    if variant == 0:
        out[:nu] = P1 @ ru; out[nu:] = PS @ rp
    else:
        zu = P_mult(ru); zu[conu] = ru[conu] / np.diag(Auu)[conu]
        out[:nu] = zu
        ri, rf = rp[:npi], rp[npi:]
        t = rf - Sif.T @ (ri / Dii)
        yf = t / Fd + Pf @ (Faux_inv @ (Pf.T @ t))
        cf = conp[npi:]; yf[cf] = t[cf] / Fd[cf]
        yi = (ri - Sif @ yf) / Dii
        out[nu:nu + npi] = yi; out[nu + npi:] = yf
    return out
for variant in (0, 1):
    its = [0]
    def cb(x): its[0] += 1
    M = spla.LinearOperator((n, n), matvec=lambda x: Pfull(x, variant))
    x, info = spla.minres(sp.csr_matrix(K), rhs, M=M, rtol=1e-10, maxiter=2000, callback=cb)
    print("variant", variant, "minres its", its[0], "true rel res", np.linalg.norm(K @ x - rhs) / np.linalg.norm(rhs))
