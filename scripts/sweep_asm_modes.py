"""Assembly kernel time per scatter mode on cfg 2 (N from argv): RED atomics, colour-ordered, tiled gather."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import poroelasticity_b200 as pb
nref = int(sys.argv[1]) if len(sys.argv) > 1 else 10
ref = None
for name, opts, gather in (("RED atomics", 1, 0), ("tiled gather", 1, 1), ("colour-ordered", 33, 0)):
    b = pb.PoroElasBR1WG()
    b.set_tuning("asm_gather", gather)
    b.set_solver_options(opts)
    s = b.make_grid_and_dofs(nref)
    b.time_step = 0.1; b.set_material(1, 1, 1, 0); b.set_case(pb.PE_CASE_COSCOS)
    best = 1e9
    for rep in range(4):
        b.assemble_system(0.1)
        best = min(best, b.timings()["assemble_kernel_ms"])
    v = b.system_matrix() if nref <= 10 else None
    if v is not None:
        if ref is None:
            ref = v
        err = np.abs(v - ref).max() / np.abs(ref).max()
    else:
        err = float("nan")
    print("%-16s assemble kernels %.3f ms  (%.3g cells/s, %.1f%% of HBM roof at 1904 B/cell)  max rel diff to RED %.1e"
          % (name, best, s.n_cells / best * 1e3, 1904 * s.n_cells / best * 1e3 / 1e9 / 6547.2 * 100, err), flush=True)
    b.close()
