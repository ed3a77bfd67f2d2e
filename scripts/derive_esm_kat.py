"""Known-answer vector for the WG(Q2,Q2;RT[2]) local matrix of EltStiffMat.cc on the unit square, K = I:
EXACT rational arithmetic (sympy) on a MONOMIAL basis of RT_[2] = Q_{3,2} e_x + Q_{2,3} e_y -- independent of the
oracle's Legendre basis, Gauss quadrature and Gauss-Jordan inverse.  A = F M^-1 F^T with
  M_kl = int w_k . w_l,   F_ik = - int phi_i^int div w_k + sum_f int_f phi_i^face w_k . n        (ESM:413-502)
local order: 3 FE_FaceQ(2) dofs per face (faces 0..3, Lagrange at 0, 1/2, 1 along the line), then the 9 FE_DGQ(2)
dofs (x fastest).  Writes tests/golden/esm_exact_unit_square.npy; exact values: trace = 496/5, A[0,0] = 32/15,
A[12,12] = 148/45.   Run:  python scripts/derive_esm_kat.py"""
import os
import sympy as sp, numpy as np, time
t0=time.time()
x,y=sp.symbols('x y')
# RT[2] = Q_{3,2} e_x + Q_{2,3} e_y, monomial basis
W=[(x**a*y**b,0) for b in range(3) for a in range(4)]+[(0,x**a*y**b) for b in range(4) for a in range(3)]
def lag(t):
    return [2*(t-sp.Rational(1,2))*(t-1), -4*t*(t-1), 2*t*(t-sp.Rational(1,2))]
Lx,Ly=lag(x),lag(y)
interior=[Lx[i]*Ly[j] for j in range(3) for i in range(3)]
# faces: 0: x=0, 1: x=1 (param y), 2: y=0, 3: y=1 (param x); normals -x,+x,-y,+y
def I2(f): return sp.integrate(sp.integrate(f,(x,0,1)),(y,0,1))
M=sp.zeros(24,24)
for i in range(24):
    for j in range(i,24):
        v=I2(W[i][0]*W[j][0]+W[i][1]*W[j][1]); M[i,j]=v; M[j,i]=v
F=sp.zeros(21,24)
for k in range(24):
    div=sp.diff(W[k][0],x)+sp.diff(W[k][1],y)
    for i in range(9):
        F[12+i,k]=-I2(interior[i]*div)
    for f in range(4):
        for i in range(3):
            if f==0: val=sp.integrate((lag(y)[i]*(-W[k][0])).subs(x,0),(y,0,1))
            if f==1: val=sp.integrate((lag(y)[i]*( W[k][0])).subs(x,1),(y,0,1))
            if f==2: val=sp.integrate((lag(x)[i]*(-W[k][1])).subs(y,0),(x,0,1))
            if f==3: val=sp.integrate((lag(x)[i]*( W[k][1])).subs(y,1),(x,0,1))
            F[3*f+i,k]=val
A=F*M.inv()*F.T
An=np.array(A.evalf(30).tolist(),dtype=float)
np.save(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden', 'esm_exact_unit_square.npy'), An)
print("exact done",time.time()-t0,"s; trace",float(A.trace()),"A[12,12]",A[12,12],"A[0,0]",A[0,0])
