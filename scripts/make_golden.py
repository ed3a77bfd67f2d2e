"""Generate the committed fixtures under tests/golden/ from the CPU oracle (oracle/liboracle.so).

The reference ships no golden vectors and cannot be built here (deal.II 9.1 + UMFPACK missing), so these files
pin the ORACLE (a regression in the restatement shows up as a diff) and give the GPU tests fixed numbers to
compare against; they are not outputs of deal.II.  Re-run:  python scripts/make_golden.py
"""
import os
import sys

import numpy as np
import scipy.sparse.linalg as spla

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle_lib import BR1_WG, CGQ1_WG, CASE_COSCOS, EsmOracle, Oracle  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def biot(element, name, nref=2, distort=0.15, lam=3.0, mu=0.7, alpha=0.93, c0=0.1, dt=0.05, steps=2):
    o = Oracle(element, nref, distort=distort)
    o.set_material(lam, mu, alpha, c0, dt)
    o.set_case(CASE_COSCOS)
    k = np.exp(0.5 * np.sin(1.0 + np.arange(o.n_cells)))  # deterministic, smooth in the cell index
    o.set_kcell(k)
    old = np.zeros(o.n_dofs)
    vals, rhss, sols = [], [], []
    for s in range(steps):
        v, r, _ = o.assemble(dt * (s + 1), old)
        x = spla.splu(o.csr(v).tocsc()).solve(r)
        vals.append(v)
        rhss.append(r)
        sols.append(x)
        old = x
    rp, ci = o.pattern()
    A0, b0, _, _ = o.local(0, dt)
    np.savez_compressed(os.path.join(OUT, name), nref=nref, distort=distort, material=np.array([lam, mu, alpha, c0, dt]),
                        kcell=k, cell_dofs=o.cell_dofs(), rowptr=rp, colidx=ci, values=np.array(vals), rhs=np.array(rhss),
                        solution=np.array(sols), local_matrix_cell0=A0, local_rhs_cell0=b0,
                        sizes=np.array([o.n_cells, o.n_u, o.n_pint, o.n_pface, o.n_dofs, o.nnz]))


def esm(name, nref=2, distort=0.15):
    o = EsmOracle(nref, distort=distort)
    k = np.exp(0.5 * np.sin(1.0 + np.arange(o.n_cells)))
    o.set_kcell(k)
    v, r, s0, _ = o.assemble()
    x = spla.spsolve(o.csr(v).tocsc(), r)
    rp, ci = o.pattern()
    A, b, _ = o.local_matrices(0, 2)
    np.savez_compressed(os.path.join(OUT, name), nref=nref, distort=distort, kcell=k, cell_dofs=o.cell_dofs(), rowptr=rp,
                        colidx=ci, values=v, rhs=r, solution0=s0, solution=x, local_matrices=A, local_rhs=b,
                        sizes=np.array([o.n_cells, o.n_dofs, o.nnz]))


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    biot(BR1_WG, "br1wg_n4.npz")
    biot(CGQ1_WG, "cgq1wg_n4.npz")
    esm("eltstiffmat_n4.npz")
    print(sorted(os.listdir(OUT)), sum(os.path.getsize(os.path.join(OUT, f)) for f in os.listdir(OUT)), "bytes")
