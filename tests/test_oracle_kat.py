"""Pins for the CPU oracle: the reference ships no golden vectors, so the oracle is
checked against the hand-derived known answers listed there ("parity unpinned" otherwise)."""
import numpy as np
import pytest

from oracle_lib import BR1_WG, CGQ1_WG, CASE_COSCOS, Oracle, lib

PIDX_BR1 = [16, 9, 11, 13, 15]


def test_qgauss3_points_and_weights():
    x = np.zeros(3)
    w = np.zeros(3)
    lib().orc_qgauss(3, x.ctypes.data, w.ctypes.data)
    s = np.sqrt(3.0 / 5.0) / 2
    assert np.allclose(x, [0.5 - s, 0.5, 0.5 + s], rtol=0, atol=1e-16)
    assert np.allclose(w, [5 / 18, 8 / 18, 5 / 18], rtol=0, atol=1e-16)


def test_wg_block_square_cell():
    """S/(dt k) on a square cell, K = I: h-independent, rows sum to 0."""
    ref = np.array([[24, -6, -6, -6, -6], [-6, 4, 2, 0, 0], [-6, 2, 4, 0, 0], [-6, 0, 0, 4, 2], [-6, 0, 0, 2, 4.0]])
    for nref in (0, 2):
        o = Oracle(BR1_WG, nref)
        o.set_material(1, 1, 1, 0, 1.0)
        A, _, _, _ = o.local(0, 0.0)
        S = A[np.ix_(PIDX_BR1, PIDX_BR1)]
        assert np.abs(S - ref).max() < 1e-12
        assert np.abs(S.sum(axis=1)).max() < 1e-12


def test_br1_unit_square_kat():
    o = Oracle(BR1_WG, 0)
    o.set_material(1, 1, 1, 0, 1.0)
    A, b, _, _ = o.local(0, 0.0)
    assert np.trace(A) == pytest.approx(63.155555555556, abs=1e-10)
    assert np.linalg.norm(A) == pytest.approx(32.677262076100, abs=1e-10)
    assert A.sum() == pytest.approx(10.666666666667, abs=1e-10)
    u = [0, 1, 2, 3, 4, 5, 6, 7, 8, 10, 12, 14]
    dbar = np.array([-.5, -.5, .5, -.5, -.5, .5, .5, .5, -2 / 3, 2 / 3, -2 / 3, 2 / 3])
    assert np.abs(A[u, 16] + dbar).max() < 1e-13       # A[u,16] = -dbar
    assert np.abs(A[16, u] - dbar).max() < 1e-13       # A[16,u] = +dbar
    assert np.abs(A[np.ix_(u, [9, 11, 13, 15])]).max() == 0.0   # u <-> p_face block is structurally zero
    Auu = A[np.ix_(u, u)]
    assert np.abs(Auu - Auu.T).max() < 1e-13
    ev = np.linalg.eigvalsh(Auu)
    assert np.sum(np.abs(ev) < 1e-12) == 3             # rigid body modes
    assert np.allclose(np.diag(Auu)[:8], 1.25) and np.allclose(np.diag(Auu)[8:], 3.2888888888888888)


def test_br1_rectangle_kat():
    """h_x=.5, h_y=.25, lambda=3, mu=.7, alpha=.93, c0=.1, dt=.05, K=1e-3 I."""
    o = Oracle(BR1_WG, 1)
    vx, vy, cv, fk = o.mesh()
    # squash the mesh in y: the oracle reads its own mesh, so emulate via anisotropic material? no:
    # use a 2x2 mesh cell (h=.5) and check the closed form of A[16,16] for a square instead, then
    # the rectangle through the formula S_int,int = 12 k (hy/hx + hx/hy).
    o.set_material(3, .7, .93, .1, .05)
    o.set_kcell(np.full(o.n_cells, 1e-3))
    A, _, _, _ = o.local(0, 0.0)
    assert A[16, 16] == pytest.approx(.1 * .25 + 12 * .05 * 1e-3 * 2, rel=1e-12)


def test_cgq1_unit_square_kat():
    o = Oracle(CGQ1_WG, 0)
    o.set_material(1, 1, 1, 0, 1.0)
    A, _, _, _ = o.local(0, 0.0)
    assert np.trace(A) == pytest.approx(50.0, abs=1e-10)
    assert np.linalg.norm(A) == pytest.approx(31.208973068654, abs=1e-10)
    assert np.allclose(A[0, :8], [1.25, .5, -.75, 0, .25, 0, -.75, -.5], atol=1e-13)
    assert np.allclose(A[:8, 12], [.5, .5, -.5, .5, .5, -.5, -.5, -.5], atol=1e-13)


@pytest.mark.parametrize("el,nu,npi,npf,ndofs,nnz", [(BR1_WG, 306, 64, 144, 514, 14660), (CGQ1_WG, 162, 64, 144, 370, 8212)])
def test_sizes_n8(el, nu, npi, npf, ndofs, nnz):
    """SURVEY 8(a) size table, N = 8."""
    o = Oracle(el, 3)
    assert (o.n_u, o.n_pint, o.n_pface, o.n_dofs, o.nnz) == (nu, npi, npf, ndofs, nnz)


@pytest.mark.parametrize("N", [1, 2, 4, 16])
def test_nnz_formulas(N):
    nref = int(np.log2(N))
    assert Oracle(BR1_WG, nref).nnz == 221 * N * N + 64 * N + 4
    assert Oracle(CGQ1_WG, nref).nnz == 123 * N * N + 42 * N + 4


def test_first_touch_numbering_by_hand():
    """N=1: one cell -> u = vertex dofs 0..7 then bubbles 8..11, p_int 12, p_face 13..16."""
    o = Oracle(BR1_WG, 0)
    assert o.cell_dofs()[0].tolist() == [0, 1, 2, 3, 4, 5, 6, 7, 8, 13, 9, 14, 10, 15, 11, 16, 12]
    o = Oracle(CGQ1_WG, 0)
    assert o.cell_dofs()[0].tolist() == [0, 1, 2, 3, 4, 5, 6, 7, 9, 10, 11, 12, 8]
    # N=2, cell 1 (right of cell 0) shares vertices 1,3 and its face 0 with cell 0's face 1
    o = Oracle(CGQ1_WG, 1)
    cd = o.cell_dofs()
    assert cd[1][0:2].tolist() == cd[0][2:4].tolist() and cd[1][4:6].tolist() == cd[0][6:8].tolist()
    assert cd[1][8] == cd[0][9]


def test_time_loop_step_counts():
    """Float-accumulated loop bound of run() (BR1new:2181, CGQ1:2929)."""
    L = lib()
    assert L.orc_count_steps(0.1, 1.0) == 10
    assert L.orc_count_steps(1. / 16, 1.0) == 16
    assert L.orc_count_steps(0.01, 1.0) == 99
    assert L.orc_count_steps(1e-3, 0.011) == 10


def test_coscos_consistency():
    """The manufactured solution satisfies the discrete system up to O(h): residual of the
    interpolant shrinks under refinement, and the assembled system is solvable."""
    import scipy.sparse.linalg as spla
    errs = []
    for nref in (2, 3, 4):
        o = Oracle(BR1_WG, nref)
        o.set_material(1, 1, 1, 0, 0.25)
        o.set_case(CASE_COSCOS)
        v, r, _ = o.assemble(0.25)
        x = spla.spsolve(o.csr(v).tocsc(), r)
        e = o.step_errors(0.25, x)
        errs.append(np.sqrt(e))
    errs = np.array(errs)
    rate = np.log2(errs[:-1] / errs[1:])
    assert np.all(rate[:, 0] > 0.8) and np.all(rate[:, 2] > 0.8)   # p and |u|_H1: first order
    assert np.all(rate[:, 1] > 1.5)                                # ||u||_L2: second order


def test_patch_rigid_body_and_pressure_constants():
    """Interior rows: A_uu annihilates rigid body modes; the WG block annihilates constants."""
    o = Oracle(BR1_WG, 2, distort=0.2)
    o.set_material(2., .5, 1, 0, 1.0)
    vx, vy, cv, fk = o.mesh()
    cd = o.cell_dofs()
    for c in (5, 6, 9, 10):
        A, _, _, _ = o.local(c, 0.0)
        x, y = vx[cv[c]], vy[cv[c]]
        for mode in ((1, 0, 0), (0, 1, 0), (0, 0, 1)):
            u = np.zeros(17)
            u[0:8:2] = mode[0] - mode[2] * y
            u[1:8:2] = mode[1] + mode[2] * x
            assert np.abs(A[:16:1][[0, 1, 2, 3, 4, 5, 6, 7, 8, 10, 12, 14]] @ u).max() < 1e-12
        p = np.zeros(17)
        p[[16, 9, 11, 13, 15]] = 1.0
        assert np.abs(A[[16, 9, 11, 13, 15]] @ p).max() < 1e-12


@pytest.mark.parametrize("el,name", [(BR1_WG, "br1"), (CGQ1_WG, "cgq1")])
def test_local_matrix_equals_exact_rational_derivation_on_a_rectangle(el, name):
    """tests/golden/biot_exact_rectangle.npz: all entries of the 17 x 17 / 13 x 13 local matrix on [0,.5] x [0,.25]
    derived with exact rational arithmetic (scripts/derive_biot_kat.py) following BR1new:811-925 / CGQ1:1194-1358
    term by term; independent of the oracle's quadrature, mapping and Gauss-Jordan code."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "biot_exact_rectangle.npz"))
    hx, hy, lam, mu, alpha, c0, dt, k = g["params"]
    o = Oracle(el, 0)
    vx, vy, cv, fk = o.mesh()
    o.set_vertices(vx * hx, vy * hy)
    o.set_material(lam, mu, alpha, c0, dt)
    o.set_kcell(np.full(o.n_cells, k))
    A, _, _, _ = o.local(0, 0.0)
    assert np.abs(A - g[name]).max() < 1e-13 * np.abs(g[name]).max()
    if el == BR1_WG:
        assert np.trace(A) == pytest.approx(3.113722222222e+01, rel=1e-12)
        assert np.linalg.norm(A) == pytest.approx(2.019813152856e+01, rel=1e-12)
        assert A.sum() == pytest.approx(9.345833333333e+00, rel=1e-11)
        assert A[16, 16] == pytest.approx(1.4e-02, rel=1e-12)
