import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import poroelasticity_b200 as pb
nref = int(sys.argv[1]) if len(sys.argv) > 1 else 9
b = pb.PoroElasBR1WG(); s = b.make_grid_and_dofs(nref)
b.time_step = 0.1; b.set_material(1, 1, 1, 0); b.set_case(pb.PE_CASE_COSCOS)
xs = {}
for tol in (1e-10, 1e-13):
    b.set_solution(np.zeros(s.n_dofs)); b.rel_tol = tol
    xs[tol] = []
    for k in range(5):
        print("---- tol %g step %d" % (tol, k), file=sys.stderr, flush=True)
        b.assemble_system(0.1 * (k + 1)); xs[tol].append(b.solve_time_step(check=False).copy())
        print("its", b.stats.iterations, "rel", b.stats.rel_residual, file=sys.stderr, flush=True)
for k in range(5):
    a, c = xs[1e-10][k], xs[1e-13][k]
    print("step", k, "true rel error of the 1e-10 solve: %.3e" % (np.linalg.norm(a - c) / np.linalg.norm(c)), file=sys.stderr)
