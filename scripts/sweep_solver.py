"""Solver parameter sweep on cfg 2 (N from argv): iterations and solve time per option set."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import poroelasticity_b200 as pb
nref = int(sys.argv[1]) if len(sys.argv) > 1 else 10
b = pb.PoroElasBR1WG()
s = b.make_grid_and_dofs(nref)
b.time_step = 0.1; b.set_material(1, 1, 1, 0); b.set_case(pb.PE_CASE_COSCOS)
for (mgopt, theta, omega, nu) in [(1, 1.0, 0.6, 2), (1, 1.0, 0.6, 1), (1, 1.0, 0.7, 1), (1, 1.0, 0.7, 2), (1, 1.0, 0.8, 2), (1, 1.0, 0.6, 3), (65, 1.0, 0.6, 2)]:
    b.set_solver_options(mgopt, theta, omega, nu)
    best = None
    for rep in range(2):
        b.set_solution(np.zeros(s.n_dofs))
        b.assemble_system(0.1)
        b.rel_tol = 1e-10; b.max_iterations = 600
        b.solve_time_step(want_solution=False, check=False)
        ms = b.stats.seconds * 1e3
        best = ms if best is None else min(best, ms)
    print("opts mg=%d omega=%.2f nu=%d: it=%d conv=%d rel=%.2e solve %.1f ms (%.3f ms/it)" % (mgopt, omega, nu,
          b.stats.iterations, b.stats.converged, b.stats.rel_residual, best, best / max(1, b.stats.iterations)), flush=True)
