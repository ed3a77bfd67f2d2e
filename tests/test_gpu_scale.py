"""Parity at production scale: the code paths that only exist above one assembly chunk (chunk-range zeroing,
chunk-boundary REDs) and above the fused multigrid levels, the tolerance bench.py times with, the cfg 4 data at
128^2, and the behaviour of the stopping rule.  All through the C ABI against the CPU oracle (+ sparse LU)."""
import os
import sys

import numpy as np
import pytest

import poroelasticity_b200 as pb
from oracle_lib import BR1_WG, CGQ1_WG, CASE_COSCOS, Oracle, lib

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

TOL_MAT = 1e-12
TOL_SOL = 1e-8


def rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def lu(o, v, r):
    import scipy.sparse.linalg as spla
    return spla.splu(o.csr(v).tocsc(), permc_spec="MMD_AT_PLUS_A").solve(r)


def rel_l2_blocks(x, x_ref, n_u):
    eu = np.linalg.norm(x[:n_u] - x_ref[:n_u]) / np.linalg.norm(x_ref[:n_u])
    ep = np.linalg.norm(x[n_u:] - x_ref[n_u:]) / np.linalg.norm(x_ref[n_u:])
    return eu, ep


@pytest.fixture(scope="module")
def oracle128():
    out = {}
    for el in (BR1_WG, CGQ1_WG):
        o = Oracle(el, 7, distort=0.15)
        o.set_material(2.0, 0.8, 0.9, 0.05, 0.1)
        o.set_case(CASE_COSCOS)
        rng = np.random.default_rng(11)
        old = rng.normal(size=o.n_dofs)
        v, r, _ = o.assemble(0.3, old)
        out[el] = (o, old, v, r)
    return out


@pytest.mark.parametrize("el", [BR1_WG, CGQ1_WG])
@pytest.mark.parametrize("mode", ["atomics", "colour", "gather"])
def test_assembly_over_several_chunks(el, mode, oracle128):
    """128^2 = 16384 cells in chunks of 4096: four chunks, k_zero_ranges between them, REDs across chunk
    boundaries -- in all three scatter modes, matrix and rhs to 1e-12 of the oracle."""
    o, old, v_ref, r_ref = oracle128[el]
    b = pb.BiotSystem(element=el)
    b.set_tuning("asm_gather", 1 if mode == "gather" else 0)
    b.set_solver_options(1 | (32 if mode == "colour" else 0) | (1 << 8))   # chunk = 1 x 4096 cells
    b.make_grid_and_dofs(7, distort=0.15)
    b.time_step = 0.1
    b.set_material(2.0, 0.8, 0.9, 0.05)
    b.set_case(pb.PE_CASE_COSCOS)
    for rep in range(2):     # the second pass runs on values left by the first (zeroing / first-writer stores)
        b.assemble_system(0.3, old)
        v, r = b.system_matrix(), b.system_rhs()
        assert rel(v, v_ref) < TOL_MAT, (mode, rep)
        assert rel(r, r_ref) < TOL_MAT, (mode, rep)
    b.close()


def test_changing_options_after_setup_cannot_corrupt_the_matrix(oracle128):
    """flipping the scatter-mode bit (or only tuning omega / nu) after the pattern exists must not
    change what is assembled; a new chunk size rebuilds the chunk tables and keeps the resident solution."""
    o, old, v_ref, r_ref = oracle128[BR1_WG]
    b = pb.BiotSystem(element=BR1_WG)
    b.make_grid_and_dofs(7, distort=0.15)
    b.time_step = 0.1
    b.set_material(2.0, 0.8, 0.9, 0.05)
    b.set_case(pb.PE_CASE_COSCOS)
    b.assemble_system(0.3, old)
    assert rel(b.system_matrix(), v_ref) < TOL_MAT
    for opts in (33, 1, 1 | (2 << 8), 33 | (1 << 8), 1):
        b.set_solver_options(opts, omega=0.7, nu=2)
        b.assemble_system(0.3)          # old_solution resident: must have survived the table rebuild
        assert rel(b.system_matrix(), v_ref) < TOL_MAT, opts
        assert rel(b.system_rhs(), r_ref) < TOL_MAT, opts
    # a boundary / permeability change rebuilds the device tables but keeps old_solution
    x, y, cv, fk = b.mesh()
    b.set_face_kinds(fk)
    b.assemble_system(0.3)
    assert rel(b.system_rhs(), r_ref) < TOL_MAT
    assert np.array_equal(b.solution(), old)
    b.close()


@pytest.mark.parametrize("krylov", [1, 0])
def test_bench_tolerance_gives_1e8_solution_parity(krylov):
    """Both Krylov methods (1 = FGMRES + block-triangular preconditioner, the default; 0 = MINRES).  The tolerance bench.py times with (DEFAULT_REL_TOL), warm starts and extrapolated initial guesses included,
    on 256^2 (three un-fused multigrid levels, 16 chunks of 4096 cells): per-step u and p within 1e-8 relative L2
    of the sparse LU solution of the oracle's system, over four steps of the reference loop."""
    import bench
    dt, nref = 0.1, 8
    o = Oracle(BR1_WG, nref)
    o.set_material(1, 1, 1, 0, dt)
    o.set_case(CASE_COSCOS)
    b = pb.PoroElasBR1WG()
    b.set_solver_options(1 | (1 << 8))
    b.make_grid_and_dofs(nref)
    b.time_step = dt
    b.set_material(1, 1, 1, 0)
    b.set_case(pb.PE_CASE_COSCOS)
    b.rel_tol = bench.DEFAULT_REL_TOL
    b.set_krylov(krylov)
    old = np.zeros(o.n_dofs)
    its = []
    for k, t in enumerate(bench.run_times(dt)[:4]):
        v_ref, r_ref, _ = o.assemble(t, old)
        x_ref = lu(o, v_ref, r_ref)
        b.assemble_system(t)
        if k == 0:
            assert rel(b.system_matrix(), v_ref) < TOL_MAT and rel(b.system_rhs(), r_ref) < TOL_MAT
        x = b.solve_time_step()
        assert b.stats.converged == 1
        assert b.stats.rel_residual <= bench.DEFAULT_REL_TOL * (1 + 1e-9)
        eu, ep = rel_l2_blocks(x, x_ref, o.n_u)
        assert eu < TOL_SOL and ep < TOL_SOL, (k, eu, ep)
        its.append(b.stats.iterations)
        old = x_ref
    # warm starts pay: later steps need fewer iterations than the cold first step
    assert min(its[1:]) < its[0], its
    b.close()


def test_cfg4_at_128():
    """BASELINE cfg 4 data (lambda = 1e6, c0 = 1e-6, K = exp(2 Z)) on 128^2: matrix / rhs to 1e-12, residual to the
    requested tolerance; the solution agrees with sparse LU to what cond(A) ~ 1e11 leaves of LU itself."""
    L = lib()
    nref, n = 7, 1 << 14
    u1 = np.array([L.orc_u01(12345, c, 0) for c in range(n)])
    u2 = np.array([L.orc_u01(12345, c, 1) for c in range(n)])
    k = np.exp(2.0 * np.sqrt(-2 * np.log(u1)) * np.cos(2 * np.pi * u2))
    o = Oracle(BR1_WG, nref)
    o.set_material(1e6, 1, 1, 1e-6, 0.1)
    o.set_kcell(k)
    o.set_case(CASE_COSCOS)
    b = pb.PoroElasBR1WG()
    b.set_solver_options(1 | (1 << 8))
    b.make_grid_and_dofs(nref)
    b.time_step = 0.1
    b.set_material(1e6, 1, 1, 1e-6)
    b.set_lognormal_k(2.0, 12345)
    b.set_case(pb.PE_CASE_COSCOS)
    v_ref, r_ref, _ = o.assemble(0.1)
    b.assemble_system(0.1)
    assert rel(b.system_matrix(), v_ref) < TOL_MAT
    assert rel(b.system_rhs(), r_ref) < TOL_MAT
    b.rel_tol = 1e-9
    x = b.solve_time_step()
    assert b.stats.converged >= 1
    x_ref = lu(o, v_ref, r_ref)
    A = o.csr(v_ref)
    # both are solutions to working accuracy of an ill-conditioned system: compare residuals, and the solutions loosely
    assert np.linalg.norm(r_ref - A @ x) <= 2e-9 * np.linalg.norm(r_ref)
    eu, ep = rel_l2_blocks(x, x_ref, o.n_u)
    assert eu < 1e-5 and ep < 1e-5, (eu, ep)
    b.close()


def test_vanishing_rhs_does_not_stall_the_solver():
    """at t = 2 the coscos data vanish (sin(pi t / 2) = 0) while old_solution does not; a purely
    ||b||-relative test then asks for less than the rounding error of A x_old.  The solve must end converged
    (1: tolerance met, 2: backward stable at the rounding floor), as the LU solve of the reference would."""
    dt = 0.1
    o = Oracle(BR1_WG, 5)
    o.set_material(1, 1, 1, 0, dt)
    o.set_case(CASE_COSCOS)
    b = pb.PoroElasBR1WG()
    b.make_grid_and_dofs(5)
    b.time_step = dt
    b.set_material(1, 1, 1, 0)
    b.set_case(pb.PE_CASE_COSCOS)
    b.rel_tol = 1e-10
    # old_solution = LU solution at t = 1.9
    v, r, _ = o.assemble(1.9)
    old = lu(o, v, r)
    v2, r2, _ = o.assemble(2.0, old)
    x_ref = lu(o, v2, r2)
    b.assemble_system(2.0, old)
    x = b.solve_time_step()
    assert b.stats.converged in (1, 2)
    A = o.csr(v2)
    anorm = np.abs(A).sum(axis=1).max()
    res = np.linalg.norm(r2 - A @ x)
    assert res <= max(1e-10 * np.linalg.norm(r2), 64 * 2.3e-16 * (anorm * np.linalg.norm(x) + np.linalg.norm(r2)))
    # and it is the LU solution to the accuracy either of them has
    assert np.linalg.norm(x - x_ref) <= 1e-8 * max(np.linalg.norm(x_ref), np.linalg.norm(old) * 1e-6)
    b.close()


def test_full_reference_loop_cfg1_quick_variant():
    """cfg 1 quick variant: CGQ1-WG, 8 x 8, dt = 1/16, the complete run() to T = 1 (16 steps) with the
    l2-in-time error norms the reference prints (CGQ1:3016-3020) -- against the oracle loop, 3 significant digits."""
    dt, nref = 1.0 / 16, 3
    o = Oracle(CGQ1_WG, nref)
    o.set_material(1, 1, 1, 0, dt)
    o.set_case(CASE_COSCOS)
    b = pb.PoroElas_CGQ1_WG()
    b.make_grid_and_dofs(nref)
    b.time_step = dt
    b.set_material(1, 1, 1, 0)
    b.set_case(pb.PE_CASE_COSCOS)
    b.rel_tol = 1e-12
    old = np.zeros(o.n_dofs)
    acc_ref, acc = np.zeros(3), np.zeros(3)
    t, steps = dt, 0
    while t <= 1.0:
        v, r, _ = o.assemble(t, old)
        old = lu(o, v, r)
        acc_ref += dt * o.step_errors(t, old)
        b.assemble_system(t)
        x = b.solve_time_step()
        assert np.linalg.norm(x - old) <= TOL_SOL * np.linalg.norm(old)
        acc += dt * b.step_errors(t)
        t += dt
        steps += 1
    assert steps == 16
    n_ref = np.sqrt(np.append(acc_ref, acc_ref[1] + acc_ref[2]))
    n_gpu = np.sqrt(np.append(acc, acc[1] + acc[2]))
    assert np.all(np.abs(n_gpu - n_ref) <= 5e-4 * n_ref), (n_gpu, n_ref)
    b.close()


@pytest.mark.parametrize("el", [BR1_WG, CGQ1_WG])
def test_postprocess_dilation_magnitude_darcy_velocity(el):
    """postprocess() (CGQ1:1744-1966, BR1new:1189-1246, 1352-1606): cell-averaged dilation, |u|, RT0 coefficients of
    the Darcy velocity and its cell-centre value, on a distorted mesh with per-cell K, against the literal oracle:
    1e-12 of the largest entry."""
    rng = np.random.default_rng(5)
    nref = 4
    k = np.exp(rng.normal(size=1 << (2 * nref)))
    o = Oracle(el, nref, distort=0.2)
    o.set_material(1, 1, 1, 0, 0.1)
    o.set_kcell(k)
    o.set_case(CASE_COSCOS)
    b = pb.BiotSystem(element=el)
    b.make_grid_and_dofs(nref, distort=0.2)
    b.time_step = 0.1
    b.set_material(1, 1, 1, 0, 1.0, k)
    b.set_case(pb.PE_CASE_COSCOS)
    sol = rng.normal(size=o.n_dofs)
    b.set_solution(sol)
    d_ref, m_ref, beta_ref, v_ref = o.postprocess(sol)
    d, m, beta, v = b.postprocess()
    assert rel(d, d_ref) < TOL_MAT
    assert rel(m, m_ref) < TOL_MAT
    assert rel(beta, beta_ref) < TOL_MAT
    assert rel(v, v_ref) < TOL_MAT
    # and on a solved step: the recovered velocity approximates -K grad p of the manufactured solution
    o2 = Oracle(el, 5)
    o2.set_material(1, 1, 1, 0, 0.1)
    o2.set_case(CASE_COSCOS)
    b2 = pb.BiotSystem(element=el)
    b2.make_grid_and_dofs(5)
    b2.time_step = 0.1
    b2.set_material(1, 1, 1, 0)
    b2.set_case(pb.PE_CASE_COSCOS)
    b2.rel_tol = 1e-12
    b2.assemble_system(0.1)
    x = b2.solve_time_step()
    _, _, _, vc = b2.postprocess()
    _, _, _, vc_ref = o2.postprocess(x)
    assert rel(vc, vc_ref) < 1e-10
    vx, vy, cv, _ = o2.mesh()
    cx, cy = vx[cv].mean(axis=1), vy[cv].mean(axis=1)
    st = np.sin(np.pi / 2 * 0.1)
    exact = np.pi * st * np.stack([np.sin(np.pi * cx) * np.cos(np.pi * cy), np.cos(np.pi * cx) * np.sin(np.pi * cy)], axis=1)
    assert np.abs(vc - exact).max() < 0.05 * np.abs(exact).max()
    b.close()
    b2.close()


@pytest.mark.parametrize("reuse", [1, 0])
def test_operator_is_rebuilt_when_the_matrix_inputs_change(reuse):
    """The solver operator + preconditioner are kept across steps while material, time step, permeability, boundary
    and mesh stay the same (tuning key reuse_operator, default 1).  Two steps on one matrix, then a new time step
    length, then new material data, then a new permeability field: every step within 1e-8 of the sparse LU solution
    of the oracle's system -- a stale operator would converge to the wrong system and fail the true-residual test."""
    nref = 7
    o = Oracle(BR1_WG, nref)
    b = pb.PoroElasBR1WG()
    b.set_tuning("reuse_operator", reuse)
    b.make_grid_and_dofs(nref)
    b.set_case(pb.PE_CASE_COSCOS)
    o.set_case(CASE_COSCOS)
    b.rel_tol = 1e-10
    rng = np.random.default_rng(3)
    kcell = np.exp(rng.normal(size=o.n_cells))
    old = np.zeros(o.n_dofs)
    t = 0.0
    #        lambda mu  alpha c0    dt    k field
    plan = [(1.0, 1.0, 1.0, 0.0, 0.1, None), (1.0, 1.0, 1.0, 0.0, 0.1, None), (1.0, 1.0, 1.0, 0.0, 0.05, None),
            (50.0, 2.0, 0.7, 0.01, 0.05, None), (50.0, 2.0, 0.7, 0.01, 0.05, kcell), (50.0, 2.0, 0.7, 0.01, 0.05, kcell)]
    last = None
    for k, (lam, mu, al, c0, dt, kc) in enumerate(plan):
        t += dt
        if (lam, mu, al, c0, kc is None) != last:
            o.set_material(lam, mu, al, c0, dt)
            b.set_material(lam, mu, al, c0, k_cell=kc)
            if kc is not None:
                o.set_kcell(kc)
            last = (lam, mu, al, c0, kc is None)
        else:
            o.set_material(lam, mu, al, c0, dt)
        b.time_step = dt
        v_ref, r_ref, _ = o.assemble(t, old)
        x_ref = lu(o, v_ref, r_ref)
        b.assemble_system(t)
        assert rel(b.system_matrix(), v_ref) < TOL_MAT
        x = b.solve_time_step()
        assert b.stats.converged == 1, (k, b.stats.rel_residual)
        eu, ep = rel_l2_blocks(x, x_ref, o.n_u)
        assert eu < TOL_SOL and ep < TOL_SOL, (k, eu, ep)
        old = x_ref
    b.close()
