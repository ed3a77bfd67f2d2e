"""Pins of the 3-D CGQ1-WG cell-loop restatement (oracle/poro_oracle_3d.hpp, PoroElasCGQ1_WG.cc:1165-1453 with dim = 3):
size-independent facts and closed forms (the reference ships no golden vectors)."""
import numpy as np

from oracle_lib import Oracle3D


def test_mesh_is_morton_ordered_and_conforming():
    o = Oracle3D(2)
    x, y, z, cv = o.mesh()
    assert o.n_cells == 64 and o.n_vertices == 125
    # cell 0 at the origin, its children-of-children order: cell 1 is the +x neighbour, 2 the +y, 4 the +z neighbour
    c0 = np.array([x[cv[0, 0]], y[cv[0, 0]], z[cv[0, 0]]])
    assert np.allclose(c0, 0)
    for c, d in ((1, 0), (2, 1), (4, 2)):
        p = np.array([x[cv[c, 0]], y[cv[c, 0]], z[cv[c, 0]]])
        e = np.zeros(3)
        e[d] = 0.25
        assert np.allclose(p, e)
    # vertex v of a cell sits at offset (v & 1, (v >> 1) & 1, v >> 2)
    for v in range(8):
        p = np.array([x[cv[5, v]], y[cv[5, v]], z[cv[5, v]]]) - np.array([x[cv[5, 0]], y[cv[5, 0]], z[cv[5, 0]]])
        assert np.allclose(p, 0.25 * np.array([v & 1, (v >> 1) & 1, v >> 2]))


def test_darcy_block_of_a_cube_closed_form():
    """K = k I on a cube: the 7 x 7 block (p_int, 6 faces) is dt k h [[36, -6 ...], [-6, 4, 2 (opposite face), 0 ...]]
    (exact rational derivation: RT0 Gram matrix per axis h^3/h^4 [[1/3, 1/6], [1/6, 1/3]], F = +-1)."""
    o = Oracle3D(1)
    dt, k = 0.3, 2.5
    o.set_material(0.0, 0.0, 0.0, 0.0, dt)
    o.set_kcell(np.full(o.n_cells, k))
    A, _ = o.local(3)
    h = 0.5
    P = [30] + list(range(24, 30))
    S = A[np.ix_(P, P)] / (dt * k * h)
    ref = np.zeros((7, 7))
    ref[0, 0] = 36
    ref[0, 1:] = ref[1:, 0] = -6
    for f in range(6):
        ref[1 + f, 1 + f] = 4
        ref[1 + f, 1 + (f ^ 1)] = 2
    assert np.abs(S - ref).max() < 1e-12
    assert np.abs(A[:24, :]).max() == 0 and np.abs(A[:, :24]).max() == 0


def test_elastic_block_has_the_six_rigid_body_modes():
    rng = np.random.default_rng(0)
    o = Oracle3D(1, distort=0.2)
    o.set_material(3.0, 0.7, 0.0, 0.0, 1.0)
    o.set_kcell(np.zeros(o.n_cells))
    x, y, z, cv = o.mesh()
    for c in (0, 7):
        A, _ = o.local(c)
        Au = A[:24, :24]
        assert np.abs(Au - Au.T).max() < 1e-13
        ev = np.linalg.eigvalsh(Au)
        assert np.sum(np.abs(ev) < 1e-12) == 6 and ev.min() > -1e-12
        # an explicit rotation about z: u = (-y, x, 0) at the vertices
        X = np.stack([x[cv[c]], y[cv[c]], z[cv[c]]], axis=1)
        rot = np.stack([-X[:, 1], X[:, 0], 0 * X[:, 0]], axis=1).ravel()
        assert np.abs(Au @ rot).max() < 1e-12


def test_coupling_and_storage_entries():
    """A[u, p_int] = -alpha div(phi)(centre) |J_c|, A[p_int, u] = +alpha ..., A[30, 30] = c0 (|E| + |J_c|) + Darcy,
    rhs_30 = c0 p_old |E| + alpha div(u_old)(centre) |J_c|  (CGQ1:1341-1358, 1421-1453: c0 counted twice on the left)."""
    o = Oracle3D(1)
    lam, mu, alpha, c0, dt = 2.0, 0.5, 0.9, 0.3, 0.1
    o.set_material(lam, mu, alpha, c0, dt)
    o.set_kcell(np.zeros(o.n_cells))
    h, vol = 0.5, 0.125
    old = np.random.default_rng(2).normal(size=31)
    A, b = o.local(0, old)
    # div of the vertex shape functions at the centre of a cube of side h: +-1/(4h) per component
    div_c = np.zeros(24)
    for v in range(8):
        for c in range(3):
            div_c[3 * v + c] = (1 if (v >> c) & 1 else -1) / (4 * h)
    assert np.abs(A[:24, 30] + alpha * div_c * vol).max() < 1e-14
    assert np.abs(A[30, :24] - alpha * div_c * vol).max() < 1e-14
    assert abs(A[30, 30] - 2 * c0 * vol) < 1e-14
    assert abs(b[30] - (c0 * old[30] * vol + alpha * (div_c @ old[:24]) * vol)) < 1e-13
    # lambda div div is rank one on the cube: subtracting it leaves the pure 2 mu eps:eps block
    o.set_material(0.0, mu, alpha, c0, dt)
    A0, _ = o.local(0, old)
    assert np.abs((A - A0)[:24, :24] - lam * vol * np.outer(div_c, div_c)).max() < 1e-13


def test_dof_numbering_3d_counts_and_first_cell():
    """distribute_dofs + component_wise on hexahedra: block sizes, and the first cell's dofs written out by hand
    (vertices 0..7 x 3 -> u 0..23; after the block sort the hex dof is the first p_int, the six quads the first
    six p_face numbers in face order)."""
    for n in (0, 1, 2):
        o = Oracle3D(n)
        N = 1 << n
        (n_u, n_pint, n_pface), cd, rp, ci = o.dofs_pattern()
        assert n_u == 3 * (N + 1) ** 3 and n_pint == N ** 3 and n_pface == 3 * N * N * (N + 1)
        assert list(cd[0, :24]) == list(range(24))
        assert cd[0, 30] == n_u and list(cd[0, 24:30]) == [n_u + n_pint + f for f in range(6)]
        # every dof appears, the pattern is symmetric with sorted rows and holds the diagonal
        n = n_u + n_pint + n_pface
        assert set(cd.ravel()) == set(range(n))
        import scipy.sparse as sp
        P = sp.csr_matrix((np.ones(ci.size), ci, rp), shape=(n, n))
        assert (P != P.T).nnz == 0 and np.all(P.diagonal() == 1)
        assert all(np.all(np.diff(ci[rp[r]:rp[r + 1]]) > 0) for r in range(n))


def test_dof_numbering_3d_shared_entities():
    """cell 1 (the +x neighbour of cell 0) reuses the four vertices and the quad of the common face"""
    o = Oracle3D(1)
    _, cd, _, _ = o.dofs_pattern()
    for a, b in ((1, 0), (3, 2), (5, 4), (7, 6)):  # vertex a of cell 0 == vertex b of cell 1
        assert list(cd[0, 3 * a:3 * a + 3]) == list(cd[1, 3 * b:3 * b + 3])
    assert cd[0, 25] == cd[1, 24]  # face x=1 of cell 0 is face x=0 of cell 1
    assert cd[0, 30] + 1 == cd[1, 30]
    # a row of an interior-face dof couples exactly the two cells' dofs: 31 + 31 - (12 + 1) shared
    (n_u, n_pint, n_pface), cd, rp, ci = o.dofs_pattern()
    r = cd[0, 25]
    assert rp[r + 1] - rp[r] == 31 + 31 - 13


def test_assemble_system_3d_sandwich_boundary_by_hand():
    """The shipped boundary data (CGQ1:972-981, 1563-1582, 1663-1679) on 4^3 cells: face kinds, total traction, and the
    form distribute_local_to_global leaves constrained dofs in (a positive diagonal, nothing else in row and column)."""
    import scipy.sparse as sp
    o = Oracle3D(2)
    o.set_material(3.0, 0.7, 0.93, 0.1, 1e-3)
    o.set_boundary()
    fk, tr = o.boundary()
    assert list(tr) == [0.0, 0.0, -1.0]
    assert (fk == 6).sum() == 16 and (fk == 1).sum() == 5 * 16 and (fk != 0).sum() == 6 * 16   # top: DIR_P | NEU_U
    (n_u, n_pi, n_pf), cd, rp, ci = o.dofs_pattern()
    v, r = o.assemble()
    n = n_u + n_pi + n_pf
    A = sp.csr_matrix((v, ci, rp), shape=(n, n)).toarray()
    # traction (0, 0, -1) on z = 1: every free top vertex (the 3 x 3 interior ones) receives -h^2 in its z row
    ru = r[:n_u].reshape(-1, 3)
    assert np.allclose(ru[:, :2], 0) and np.isclose(ru[:, 2].sum(), -9 / 16) and np.count_nonzero(ru[:, 2]) == 9
    assert np.allclose(r[n_u:], 0)          # f = s = 0 and old_solution = 0
    # constrained dofs: all displacement dofs of boundary vertices except the 9 free top ones, and the 16 top face pressures
    con = [i for i in range(n) if np.count_nonzero(A[i]) == 1 and np.count_nonzero(A[:, i]) == 1]
    assert len(con) == 3 * (125 - 27 - 9) + 16 and all(A[i, i] > 0 for i in con)
    # the reduced system is regular and the load pushes the free top vertices down
    free = np.setdiff1d(np.arange(n), con)
    x = np.linalg.solve(A[np.ix_(free, free)], r[free])
    uz = np.zeros(n)
    uz[free] = x
    assert uz[:n_u].reshape(-1, 3)[:, 2].min() < 0 and uz[:n_u].reshape(-1, 3)[:, 2].max() <= 1e-14


def test_assemble_system_3d_unconstrained_equals_sum_of_local_blocks():
    """no boundary data: the global matrix is the plain sum of the local blocks, rhs carries the old-solution terms"""
    o = Oracle3D(1, distort=0.2)
    o.set_material(2.0, 0.9, 0.8, 0.05, 0.01)
    o.set_boundary(np.zeros((o.n_cells, 6), np.uint8))
    (n_u, n_pi, n_pf), cd, rp, ci = o.dofs_pattern()
    n = n_u + n_pi + n_pf
    rng = np.random.default_rng(5)
    old = rng.normal(size=n)
    v, r = o.assemble(old)
    import scipy.sparse as sp
    A = sp.csr_matrix((v, ci, rp), shape=(n, n)).toarray()
    S, b = np.zeros((n, n)), np.zeros(n)
    for c in range(o.n_cells):
        Al, bl = o.local(c, old[cd[c]])
        S[np.ix_(cd[c], cd[c])] += Al
        b[cd[c]] += bl
    assert np.abs(A - S).max() <= 1e-13 * np.abs(S).max() and np.abs(r - b).max() <= 1e-13 * np.abs(b).max()
