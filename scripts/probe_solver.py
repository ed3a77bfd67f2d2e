import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import poroelasticity_b200 as pb
nrefs = [int(a) for a in sys.argv[1].split(",")] if len(sys.argv) > 1 else [6, 8, 10]
opts = [(1, 1.0, 0.6, 1), (1, 1.0, 0.6, 1), (1, 1.0, 0.6, 2), (65, 1.0, 0.6, 1)]
for nref in nrefs:
    b = pb.PoroElasBR1WG()
    t0 = time.time(); s = b.make_grid_and_dofs(nref); ts = time.time() - t0
    b.time_step = 0.1; b.set_material(1, 1, 1, 0); b.set_case(pb.PE_CASE_COSCOS)
    b.assemble_system(0.1); b.assemble_system(0.1)
    tm = b.timings()
    sp = b.bench_spmv(20)
    b.set_solver_options(17, 1.0, 0.6, 1); sp_old = b.bench_spmv(20); b.set_solver_options(1, 1.0, 0.6, 1)
    print('   spmv warp-per-row %.3f ms, stream %.3f ms' % (sp_old, sp))
    print("N=%d setup %.1fs asm %.3f ms (kernel %.3f) -> %.3g cells/s | spmv %.3f ms %.0f GB/s" % (
        1 << nref, ts, tm["assemble_ms"], tm["assemble_kernel_ms"], s.n_cells / tm["assemble_kernel_ms"] * 1e3, sp,
        (12. * s.nnz + 20 * s.n_dofs) / sp / 1e6), flush=True)
    for o in opts:
        b.set_solver_options(*o)
        b.set_solution(np.zeros(s.n_dofs))
        b.rel_tol = 1e-10; b.max_iterations = 600
        b.solve_time_step(want_solution=False, check=False)
        print("   opts %s: minres it=%d conv=%d rel=%.2e %.1f ms (%.3f ms/it)" % (o, b.stats.iterations, b.stats.converged,
              b.stats.rel_residual, b.stats.seconds * 1e3, b.stats.seconds * 1e3 / max(1, b.stats.iterations)), flush=True)
    b.close()
