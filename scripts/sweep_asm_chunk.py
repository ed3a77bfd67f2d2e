"""Assembly chunk size sweep (atomics mode: rows first touched by a chunk are zeroed right before it runs)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import poroelasticity_b200 as pb
nref = int(sys.argv[1]) if len(sys.argv) > 1 else 10
for units in (2, 4, 8, 12, 16, 24, 32):
    b = pb.PoroElasBR1WG()
    b.set_solver_options(1 | (units << 8))
    s = b.make_grid_and_dofs(nref)
    b.time_step = 0.1; b.set_material(1, 1, 1, 0); b.set_case(pb.PE_CASE_COSCOS)
    best = 1e9
    for rep in range(4):
        b.assemble_system(0.1)
        best = min(best, b.timings()["assemble_kernel_ms"])
    print("chunk %6d cells: assembly kernels %.3f ms -> %.3g cells/s" % (units * 4096, best, s.n_cells / best * 1e3), flush=True)
    b.close()
