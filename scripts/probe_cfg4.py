"""BASELINE cfg 4 probe: MINRES iterations vs lambda / log-normal K on the GPU solver."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import poroelasticity_b200 as pb
nrefs = [int(a) for a in sys.argv[1].split(",")] if len(sys.argv) > 1 else [6, 8]
maxit = int(sys.argv[2]) if len(sys.argv) > 2 else 1500
for nref in nrefs:
    for lam, c0, sigma in [(1, 0, 0), (1, 0, 2), (1e2, 1e-6, 0), (1e4, 1e-6, 0), (1e6, 1e-6, 0), (1e6, 1e-6, 2)]:
        b = pb.PoroElasBR1WG()
        s = b.make_grid_and_dofs(nref)
        b.time_step = 0.1; b.set_material(lam, 1, 1, c0); b.set_case(pb.PE_CASE_COSCOS)
        if sigma > 0:
            b.set_lognormal_k(sigma, 12345)
        b.assemble_system(0.1)
        b.rel_tol = 1e-10; b.max_iterations = maxit
        b.solve_time_step(want_solution=False, check=False)
        print("N=%d lam=%g c0=%g sigma=%g: it=%d conv=%d rel=%.2e %.1f ms" % (1 << nref, lam, c0, sigma,
              b.stats.iterations, b.stats.converged, b.stats.rel_residual, b.stats.seconds * 1e3), flush=True)
        b.close()
