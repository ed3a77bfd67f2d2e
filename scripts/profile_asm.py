import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import poroelasticity_b200 as pb
b = pb.PoroElasBR1WG()
if len(sys.argv) > 1: b.set_solver_options(int(sys.argv[1]), 1.0, 0.6, 1)
s = b.make_grid_and_dofs(10)
b.time_step = 0.1; b.set_material(1, 1, 1, 0); b.set_case(pb.PE_CASE_COSCOS)
b.assemble_system(0.1); b.assemble_system(0.2)
print(b.timings())
print("spmv", b.bench_spmv(3))
