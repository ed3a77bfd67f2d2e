"""3-D (first step): the CGQ1-WG cell loop on hexahedra (PoroElasCGQ1_WG.cc:1165-1453 with dim = 3, the
shipped main() of that file) through the C ABI against the literal 3-D oracle: 31 x 31 local matrices and local
right-hand sides to 1e-12, on the shipped 8^3 mesh with the sandwich permeability (CGQ1:176-189) and on distorted
hexahedra."""
import numpy as np
import pytest

import poroelasticity_b200 as pb
from oracle_lib import Oracle3D

pytestmark = pytest.mark.gpu


def rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def sandwich_k(o):
    """Coefficient::value_list of the shipped program: K = 1e-8 for 1/4 <= z <= 3/4, else 1 (evaluated at the cell centre)"""
    x, y, z, cv = o.mesh()
    zc = z[cv].mean(axis=1)
    return np.where((zc >= 0.25) & (zc <= 0.75), 1e-8, 1.0)


def test_unit_cube_mesh_matches_the_oracle():
    o = Oracle3D(3, distort=0.15, seed=7)
    b = pb.PoroElas_CGQ1_WG_3D()
    b.make_grid(3, distort=0.15, seed=7)
    x, y, z, cv = b.mesh3()
    xo, yo, zo, cvo = o.mesh()
    assert np.array_equal(cv, cvo)
    assert np.array_equal(x, xo) and np.array_equal(y, yo) and np.array_equal(z, zo)
    b.close()


@pytest.mark.parametrize("distort", [0.0, 0.2])
def test_local_matrices_3d(distort):
    rng = np.random.default_rng(3)
    nref = 3   # refine_global(3): the shipped 8 x 8 x 8 mesh (CGQ1:887)
    o = Oracle3D(nref, distort=distort)
    lam, mu, alpha, c0, dt = 1.0, 1.0, 1.0, 0.0, 1e-3   # CGQ1:863-866
    if distort > 0:
        lam, mu, alpha, c0, dt = 3.0, 0.7, 0.93, 0.1, 0.05
    k = sandwich_k(o) if distort == 0 else np.exp(rng.normal(size=o.n_cells))
    o.set_material(lam, mu, alpha, c0, dt)
    o.set_kcell(k)
    b = pb.PoroElas_CGQ1_WG_3D()
    b.set_mesh3(*o.mesh())
    b.time_step = dt
    b.set_material(lam, mu, alpha, c0, 1.0, k)
    old = rng.normal(size=(o.n_cells, 31))
    A, r = b.local_matrices3(0, o.n_cells, cell_old=old)
    for c in range(0, o.n_cells, 7):
        A_ref, b_ref = o.local(c, old[c])
        assert np.abs(A[c] - A_ref).max() <= 1e-12 * np.abs(A_ref).max(), c
        assert np.abs(r[c] - b_ref).max() <= 1e-12 * max(np.abs(b_ref).max(), 1e-300), c
    # a sub-range gives the same matrices
    A2 = b.local_matrices3(100, 50, want_rhs=False)
    assert np.array_equal(A2, A[100:150])
    b.close()


def test_calls_out_of_order_or_not_built_in_3d_are_refused():
    b = pb.PoroElas_CGQ1_WG_3D()
    b.make_grid(1)
    with pytest.raises(pb.PoroElasError):
        b.assemble_system(0.1)            # before distribute_dofs / make_sparsity_pattern
    with pytest.raises(pb.PoroElasError):
        b.solve_time_step()               # nothing assembled
    with pytest.raises(pb.PoroElasError):
        b.postprocess()                   # not built for dim = 3
    b.close()


@pytest.mark.parametrize("n_refine,distort,sandwich", [(3, 0.0, True), (2, 0.2, True), (2, 0.15, False)])
def test_assemble_system_3d(n_refine, distort, sandwich):
    """assemble_system on hexahedra through the C ABI against the oracle (CGQ1:1165-1706 with dim = 3): the shipped 8^3
    mesh with the sandwich permeability and boundary data, a distorted mesh, and a mesh with mixed user boundary kinds
    and a general traction; old_solution random; matrix and rhs to 1e-12."""
    o = Oracle3D(n_refine, distort)
    g = pb.PoroElas_CGQ1_WG_3D()
    g.make_grid(n_refine, distort)
    x, y, z, cv = o.mesh()
    zc = z[cv].mean(axis=1)
    k = np.where((zc >= 0.25) & (zc <= 0.75), 1e-8, 1.0)       # CGQ1:176-189
    lam, mu, al, c0, dt = 3.0, 0.7, 0.93, 0.1, 1e-3
    o.set_material(lam, mu, al, c0, dt)
    o.set_kcell(k)
    g.set_material(lam, mu, al, c0, k_cell=k)
    g.time_step = dt
    if sandwich:
        o.set_boundary()
    else:
        rng = np.random.default_rng(2)
        o.set_boundary()
        fk0, _ = o.boundary()                               # which faces are boundary faces
        fk = np.where(fk0 != 0, rng.integers(0, 8, size=fk0.shape).astype(np.uint8), 0).astype(np.uint8)
        tr = np.array([0.3, -0.2, 1.1])
        o.set_boundary(fk, tr)
        g.set_boundary3(fk, tr)
    s = g.setup_system()
    (n_u, n_pi, n_pf), cd, rp, ci = o.dofs_pattern()
    assert (s.n_u, s.n_pint, s.n_pface) == (n_u, n_pi, n_pf)
    rp2, ci2 = g.pattern()
    assert np.array_equal(rp2, rp) and np.array_equal(ci2, ci)
    old = np.random.default_rng(9).normal(size=n_u + n_pi + n_pf)
    v_ref, r_ref = o.assemble(old)
    g.assemble_system(0.0, old)
    v, r = g.system_matrix(), g.system_rhs()
    assert np.abs(v - v_ref).max() <= 1e-12 * np.abs(v_ref).max()
    assert np.abs(r - r_ref).max() <= 1e-12 * np.abs(r_ref).max()
    # a second assembly with the zero state reproduces the oracle as well (buffers are re-zeroed)
    v_ref0, r_ref0 = o.assemble()
    g.assemble_system(0.0, np.zeros_like(old))
    assert np.abs(g.system_matrix() - v_ref0).max() <= 1e-12 * np.abs(v_ref0).max()
    assert np.abs(g.system_rhs() - r_ref0).max() <= 1e-12 * max(np.abs(r_ref0).max(), 1e-300)
    g.close()


def test_time_steps_3d_shipped_program():
    """Three steps of the shipped 3-D program (CGQ1:859-866 data: lambda = mu = alpha = 1, c0 = 0, dt = 1/1000, 8^3 cells,
    sandwich K and boundary): assemble + solve through the C ABI, old_solution resident on the device, against the sparse
    LU solution of the oracle's system of the same step -- u and p within 1e-8 relative L2.
    The system has condition number 5e11 (K = 1e-8 in the layer, dt = 1e-3): a plain SuperLU solve is only good to 1.6e-6
    in p, so the comparison solution is computed on the symmetrically equilibrated matrix (condition 3e5) with one step
    of iterative refinement (UMFPACK, which the reference uses, scales rows by default as well)."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla

    def lu_equilibrated(A, b):
        sc = 1.0 / np.sqrt(np.abs(A.diagonal()))
        S = sp.diags(sc)
        Ah = (S @ A @ S).tocsc()
        lu = spla.splu(Ah)
        y = lu.solve(sc * b)
        y += lu.solve(sc * b - Ah @ y)
        return sc * y

    o = Oracle3D(3)
    g = pb.PoroElas_CGQ1_WG_3D()
    g.make_grid(3)
    k = sandwich_k(o)
    dt = 1e-3
    o.set_material(1.0, 1.0, 1.0, 0.0, dt)
    o.set_kcell(k)
    o.set_boundary()
    g.set_material(1.0, 1.0, 1.0, 0.0, k_cell=k)
    g.time_step = dt
    g.setup_system()
    g.rel_tol, g.max_iterations = 1e-11, 20000
    (n_u, n_pi, n_pf), cd, rp, ci = o.dofs_pattern()
    n = n_u + n_pi + n_pf
    old = np.zeros(n)
    for step in range(3):
        v_ref, r_ref = o.assemble(old)
        x_ref = lu_equilibrated(sp.csr_matrix((v_ref, ci, rp), shape=(n, n)), r_ref)
        g.assemble_system((step + 1) * dt, old if step == 0 else None)
        if step > 0:   # the resident state is the previous device solution: the rhs agrees to solver accuracy only
            assert rel(g.system_rhs(), r_ref) < 1e-7
        else:
            assert rel(g.system_matrix(), v_ref) < 1e-12 and rel(g.system_rhs(), r_ref) < 1e-12
        x = g.solve_time_step()
        assert g.stats.converged >= 1 and g.stats.rel_residual <= 1e-11 * (1 + 1e-9) or g.stats.converged == 2
        eu = np.linalg.norm(x[:n_u] - x_ref[:n_u]) / np.linalg.norm(x_ref[:n_u])
        ep = np.linalg.norm(x[n_u:] - x_ref[n_u:]) / np.linalg.norm(x_ref[n_u:])
        assert eu < 1e-8 and ep < 1e-8, (step, eu, ep, g.stats.iterations)
        old = x_ref
    # the load pushes the top down
    assert x[:n_u].reshape(-1, 3)[:, 2].min() < 0
    g.close()


def test_run_3d():
    g = pb.PoroElas_CGQ1_WG_3D()
    g.rel_tol, g.max_iterations = 1e-10, 20000
    hist = g.run(n_refine=2)
    assert len(hist) in (10, 11) and all(h[2] <= 1e-10 * (1 + 1e-9) for h in hist)
    g.close()


def test_cpp_host_driver_runs_the_shipped_3d_program():
    """host/poroelas_cgq1wg 3d: PoroElas_CGQ1_WG<3>(1/100, 1/1000, 1, 1, 1, 0).run() of the C++ mirror (CGQ1:859-866,
    2929, 3029) on 4^3 cells: the reference's stdout lines, every step converged, the top settles."""
    import os
    import re
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "host", "poroelas_cgq1wg")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-s", "-C", os.path.join(root, "host")])
    out = subprocess.run([exe, "3d", "2"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "Solving problem in 3 space dimensions." in out.stdout
    assert "Number of degrees of freedom: 679 (375+64+240)" in out.stdout
    assert out.stdout.count("Time step") in (10, 11)
    m = re.search(r"largest settlement: u_z = (-[\d.e+-]+)", out.stdout)
    assert m and float(m.group(1)) < 0
