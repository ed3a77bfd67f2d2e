import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import poroelasticity_b200 as pb
b = pb.PoroElasBR1WG()
b.set_tuning("asm_gather", int(sys.argv[1]) if len(sys.argv) > 1 else 1)
s = b.make_grid_and_dofs(10)
b.time_step = 0.1; b.set_material(1, 1, 1, 0); b.set_case(pb.PE_CASE_COSCOS)
b.assemble_system(0.1)
b.assemble_system(0.1)
print(b.timings())
