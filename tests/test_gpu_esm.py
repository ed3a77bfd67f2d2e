"""GPU parity of the WG(Q2,Q2;RT[2]) Darcy path (EltStiffMat.cc) against oracle/esm_oracle.hpp:
local matrices and right-hand sides, assembled system after apply_boundary_values, CG solve."""
import numpy as np
import pytest

import poroelasticity_b200 as pb
from oracle_lib import EsmOracle

pytestmark = pytest.mark.gpu


def rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def cg_identity(A, b, x0, tol, max_it):
    """deal.II SolverCG + PreconditionIdentity (ESM:583-590): stop when ||r||_2 <= tol."""
    x = x0.copy()
    r = b - A @ x
    p = r.copy()
    rr = r @ r
    it = 0
    while np.sqrt(rr) > tol and it < max_it:
        Ap = A @ p
        a = rr / (p @ Ap)
        x += a * p
        r -= a * Ap
        rr_new = r @ r
        p = r + (rr_new / rr) * p
        rr = rr_new
        it += 1
    return x, it


def make(nref, distort=0.0, dfma=0):
    """dfma = 0: the FP64 tensor-core (DMMA) cell kernel, the default; 1: the first, DFMA-only version"""
    o = EsmOracle(nref, distort=distort)
    b = pb.WGDarcyEquation()
    b.set_tuning("esm_dfma", dfma)
    b.make_grid(nref, distort=distort)
    b.setup_system()
    b.set_case(pb.PE_CASE_DARCY_COSCOS)
    return o, b


@pytest.mark.parametrize("dfma", [0, 1])
@pytest.mark.parametrize("distort", [0.0, 0.2])
def test_local_matrices(distort, dfma):
    o, b = make(3, distort, dfma)
    k = np.exp(np.random.default_rng(1).normal(size=o.n_cells))
    o.set_kcell(k)
    b.set_material(1, 1, 1, 0, 1.0, k)
    A, rhs = b.local_matrices(0, o.n_cells, want_rhs=True)
    A_ref, b_ref, _ = o.local_matrices(0, o.n_cells)
    assert rel(A, A_ref) < 1e-12
    assert rel(rhs, b_ref) < 1e-12
    # per cell, not only against the largest entry of the whole batch; exact symmetry of the tile-mirrored block
    scale = np.abs(A_ref).reshape(o.n_cells, -1).max(axis=1)
    assert (np.abs(A - A_ref).reshape(o.n_cells, -1).max(axis=1) / scale).max() < 1e-11
    if dfma == 0:
        assert np.array_equal(A, A.transpose(0, 2, 1))


@pytest.mark.parametrize("dfma", [0, 1])
@pytest.mark.parametrize("nref,distort", [(0, 0.0), (1, 0.0), (3, 0.0), (4, 0.2)])
def test_assembled_system(nref, distort, dfma):
    """system_matrix, system_rhs and solution after assemble_system (incl. apply_boundary_values, ESM:565-576)."""
    o, b = make(nref, distort, dfma)
    v_ref, r_ref, s_ref, _ = o.assemble()
    b.assemble_system()
    assert rel(b.system_matrix(), v_ref) < 1e-12
    assert rel(b.system_rhs(), r_ref) < 1e-12
    assert rel(b.solution(), s_ref) < 1e-14 or np.abs(s_ref).max() == 0.0


def test_shipped_run_matches_cg_reference():
    """run() as shipped (refine_global(1), 72 dofs): same CG iterates as SolverCG to rounding."""
    o, b = make(1)
    x = b.run(1)
    v_ref, r_ref, s_ref, _ = o.assemble()
    A = o.csr(v_ref)
    x_ref, it_ref = cg_identity(A, r_ref, s_ref, 1e-8 * np.linalg.norm(r_ref), 1000)
    assert b.stats.converged and abs(b.stats.iterations - it_ref) <= 1
    assert np.linalg.norm(x - x_ref) <= 1e-8 * np.linalg.norm(x_ref)


def test_solve_on_a_finer_mesh():
    import scipy.sparse.linalg as spla
    o, b = make(4)
    v_ref, r_ref, s_ref, _ = o.assemble()
    b.assemble_system()
    b.rel_tol, b.max_iterations = 1e-12, 5000
    x = b.solve()
    x_ref = spla.spsolve(o.csr(v_ref).tocsc(), r_ref)
    assert np.linalg.norm(x - x_ref) <= 1e-8 * np.linalg.norm(x_ref)


def test_cfg5_secondary_checksum():
    """BASELINE cfg 5 (secondary): the dense 21 x 21 kernel over all cells in chunks; on the uniform mesh
    every local matrix is the same, so the checksum is n_cells * sum(A_cell) (= 0 up to rounding: rows sum
    to zero) and the sum of |entries| pins the values."""
    o, b = make(6)
    cs, sec = b.local_matrices_bench(1 << 10, repeats=1)
    one = b.local_matrices(0, 1)[0]
    assert abs(cs) <= 1e-9 * np.abs(one).sum() * o.n_cells
    assert sec > 0.0


def test_cpp_host_driver_runs():
    """host/eltstiffmat: the C++ twin of EltStiffMat.cc's main() on top of the C ABI, shipped refinement."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "host", "eltstiffmat")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-s", "-C", os.path.join(root, "host")])
    out = subprocess.run([exe, "1"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "Number of active cells: 4" in out.stdout
    assert "Number of rt degrees of freedom: 84" in out.stdout
    assert "Number of dof: 72" in out.stdout


def test_local_matrix_equals_exact_rational_derivation():
    """The CUDA kernel against the exact (sympy) unit-square matrix of scripts/derive_esm_kat.py."""
    import os
    A = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "esm_exact_unit_square.npy"))
    b = pb.WGDarcyEquation()
    b.make_grid(0)
    assert rel(b.local_matrices(0, 1)[0], A) < 1e-12
    # the local matrix is invariant under uniform refinement (2-D Darcy) -- any cell of a finer Cartesian mesh
    b2 = pb.WGDarcyEquation()
    b2.make_grid(5)
    assert rel(b2.local_matrices(777, 1)[0], A) < 1e-12


@pytest.mark.parametrize("nref,distort", [(1, 0.0), (3, 0.0), (3, 0.2)])
def test_postprocess_velocity_and_flux_errors(nref, distort):
    """postprocess() of EltStiffMat.cc (ESM:662-919): L2_error_vel and L2_error_flux of the Darcy velocity recovered
    from the discrete weak gradient, CUDA vs the literal oracle on the SAME solution vector (1e-10 relative: sums of
    squares over all cells), and on the solved system 3 significant digits of what the program prints."""
    import scipy.sparse.linalg as spla
    o, b = make(nref, distort)
    k = np.exp(0.5 * np.random.default_rng(3).normal(size=o.n_cells))
    o.set_kcell(k)
    b.set_material(1, 1, 1, 0, 1.0, k)
    rng = np.random.default_rng(4)
    x = rng.normal(size=o.n_dofs)
    b.assemble_system()
    b.set_solution(x)
    e_ref = np.sqrt(o.postprocess(x))
    e = b.postprocess()
    assert np.all(np.abs(e - e_ref) <= 1e-10 * e_ref), (e, e_ref)
    # the program as shipped: K = I, solve, print
    o2, b2 = make(nref, distort)
    v, r, s, _ = o2.assemble()
    x_ref = spla.splu(o2.csr(v).tocsc()).solve(r)
    b2.assemble_system()
    b2.solve()
    e2 = b2.postprocess()
    e2_ref = np.sqrt(o2.postprocess(x_ref))
    assert np.all(np.abs(e2 - e2_ref) <= 1e-3 * e2_ref), (e2, e2_ref)
