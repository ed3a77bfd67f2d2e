"""CPU tests of the EltStiffMat.cc oracle (oracle/esm_oracle.hpp) and of the host setup layer for the
WG(Q2,Q2;RT[2]) element.  The reference ships no golden vectors for this program (parity unpinned against
deal.II); what pins the restatement are the size-independent facts below."""
import numpy as np
import pytest
import scipy.sparse.linalg as spla

import poroelasticity_b200 as pb
from oracle_lib import EsmOracle


def test_shipped_mesh_sizes():
    """refine_global(1) (ESM:245-246): 4 cells, 72 dofs = 12 lines x 3 + 4 quads x 9."""
    o = EsmOracle(1)
    assert (o.n_cells, o.n_loc, o.n_dofs, o.nnz) == (4, 21, 72, 1728)


@pytest.mark.parametrize("distort", [0.0, 0.2])
def test_local_matrix_properties(distort):
    """C M_K C^T is symmetric positive semi-definite with exactly the constants in its kernel (the weak
    gradient of a constant (interior, face) pair vanishes), on Cartesian and on distorted cells."""
    o = EsmOracle(2, distort=distort)
    A, b, _ = o.local_matrices(5, 2)
    for M in A:
        assert np.abs(M - M.T).max() < 1e-13 * np.abs(M).max()
        assert np.abs(M.sum(1)).max() < 1e-12 * np.abs(M).max()
        ev = np.linalg.eigvalsh(0.5 * (M + M.T))
        assert ev[0] > -1e-12 * ev[-1] and ev[1] > 1e-6 * ev[-1]
    assert np.all(b[:, :12] == 0.0)  # the source only tests the interior functions (ESM:505-513)


def test_local_matrix_scales_with_k_and_not_with_h():
    """In 2-D the Darcy local matrix is invariant under uniform scaling of the cell and linear in K."""
    a, b2 = EsmOracle(1), EsmOracle(3)
    A1 = a.local_matrices(0, 1)[0][0]
    A3 = b2.local_matrices(0, 1)[0][0]
    assert np.abs(A1 - A3).max() < 1e-12 * np.abs(A1).max()
    b2.set_kcell(np.full(b2.n_cells, 2.5))
    assert np.abs(b2.local_matrices(0, 1)[0][0] - 2.5 * A1).max() < 1e-12 * np.abs(A1).max()


def test_convergence_of_the_manufactured_solution():
    """p = cos(pi x) cos(pi y) (ESM:121-187): nodal error of the interior Q2 function drops at third order."""
    errs = []
    for nref in (1, 2, 3):
        o = EsmOracle(nref)
        v, r, s0, _ = o.assemble()
        x = spla.spsolve(o.csr(v).tocsc(), r)
        vx, vy, cv = o.mesh()
        cd = o.cell_dofs()
        h = 1.0 / (1 << nref)
        e = 0.0
        for c in range(o.n_cells):
            for j in range(3):
                for i in range(3):
                    ex = np.cos(np.pi * (vx[cv[c, 0]] + i * h / 2)) * np.cos(np.pi * (vy[cv[c, 0]] + j * h / 2))
                    e = max(e, abs(x[cd[c, 12 + 3 * j + i]] - ex))
        errs.append(e)
    assert errs[1] < errs[0] / 3 and errs[2] < errs[1] / 6, errs


def test_apply_boundary_values_rule():
    """Boundary rows: diagonal kept, rest zero, rhs = diag * value, solution = value; columns eliminated."""
    o = EsmOracle(2)
    v0, r0, _, _ = o.assemble(with_boundary_values=False)
    v, r, s, _ = o.assemble()
    A0, A = o.csr(v0).toarray(), o.csr(v).toarray()
    bnd = np.nonzero(s)[0]
    assert bnd.size > 0
    free = np.setdiff1d(np.arange(o.n_dofs), np.nonzero(np.abs(A - np.diag(np.diag(A))).sum(1) == 0)[0])
    con = np.setdiff1d(np.arange(o.n_dofs), free)
    assert np.allclose(np.diag(A)[con], np.diag(A0)[con])
    assert np.allclose(r[con], np.diag(A0)[con] * s[con])
    assert np.abs(A[np.ix_(free, con)]).max() == 0.0
    assert np.allclose(r[free], r0[free] - A0[np.ix_(free, con)] @ s[con])
    assert np.allclose(A[np.ix_(free, free)], A0[np.ix_(free, free)])


@pytest.mark.parametrize("nref", [0, 1, 2, 4])
def test_host_numbering_and_pattern_bit_exact(nref):
    """distribute_dofs (first touch, lines then quad, no renumbering) and make_sparsity_pattern, ESM:297-324."""
    o = EsmOracle(nref)
    b = pb.WGDarcyEquation(host_setup_only=True)
    b.make_grid(nref)
    s = b.setup_system()
    assert (s.n_cells, s.n_loc, s.n_dofs, s.nnz) == (o.n_cells, 21, o.n_dofs, o.nnz)
    assert np.array_equal(b.cell_dofs(), o.cell_dofs())
    rp, ci = b.pattern()
    rp2, ci2 = o.pattern()
    assert np.array_equal(rp, rp2) and np.array_equal(ci, ci2)


def test_local_matrix_equals_exact_rational_derivation():
    """tests/golden/esm_exact_unit_square.npy: the 21 x 21 matrix derived with exact rational arithmetic on a
    monomial RT[2] basis (scripts/derive_esm_kat.py) -- independent of the oracle's basis, quadrature and inverse."""
    import os
    A = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "esm_exact_unit_square.npy"))
    assert abs(np.trace(A) - 496.0 / 5.0) < 1e-12 and abs(A[0, 0] - 32.0 / 15.0) < 1e-14 and abs(A[12, 12] - 148.0 / 45.0) < 1e-14
    o = EsmOracle(0)  # one cell: the unit square
    Ao = o.local_matrices(0, 1)[0][0]
    assert np.abs(Ao - A).max() < 1e-12 * np.abs(A).max()
