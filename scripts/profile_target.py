"""Short single-GPU program for ncu: one assembly, a few solver SpMVs, 6 MINRES iterations on cfg 2, and one
assembly of the WG(Q2,Q2;RT[2]) Darcy problem (EltStiffMat.cc) at 256 x 256."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import poroelasticity_b200 as pb
nref = int(sys.argv[1]) if len(sys.argv) > 1 else 10
b = pb.PoroElasBR1WG()
s = b.make_grid_and_dofs(nref)
b.time_step = 0.1; b.set_material(1, 1, 1, 0); b.set_case(pb.PE_CASE_COSCOS)
b.assemble_system(0.1)
print("spmv ms", b.bench_spmv(3))
b.max_iterations = 6; b.rel_tol = 1e-10
b.solve_time_step(want_solution=False, check=False)
print("ok", b.stats.iterations, b.timings())
b.close()
e = pb.WGDarcyEquation()
e.make_grid(8); e.setup_system(); e.set_case(pb.PE_CASE_DARCY_COSCOS)
e.assemble_system()
print("esm", e.timings())
