"""Smoother-weight sweep (pe_set_tuning keys mg_omega2 / cmg_omega / cmg_omega2 / face_omega):
iterations and solve time on cfg 2 (N from argv) and on a cfg-4 like case."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import poroelasticity_b200 as pb
nref = int(sys.argv[1]) if len(sys.argv) > 1 else 10
cases = [("cfg2", 1.0, 0.0, 0.0), ("cfg4", 1e6, 1e-6, 2.0)]
C = dict(PE_CMG_OMEGA="0.6", PE_CMG_OMEGA2="1.2")
variants = [dict(C, PE_MG_OMEGA2="1.0"), dict(C, PE_MG_OMEGA2="0.8"), dict(C, PE_MG_OMEGA2="0.9"),
            dict(C, PE_MG_OMEGA2="1.0", PE_FACE_OMEGA="0.6"), dict(C, PE_MG_OMEGA2="1.0", PE_FACE_OMEGA="0.65"),
            dict(PE_CMG_OMEGA="0.55", PE_CMG_OMEGA2="1.1", PE_MG_OMEGA2="1.0"),
            dict(PE_CMG_OMEGA="0.65", PE_CMG_OMEGA2="1.3", PE_MG_OMEGA2="1.0"),
            dict(PE_CMG_OMEGA="0.6", PE_CMG_OMEGA2="1.4", PE_MG_OMEGA2="1.0")]
for name, lam, c0, sigma in cases:
    b = pb.PoroElasBR1WG()
    s = b.make_grid_and_dofs(nref if name == "cfg2" else min(nref, 9))
    b.time_step = 0.1; b.set_material(lam, 1, 1, c0); b.set_case(pb.PE_CASE_COSCOS)
    if sigma > 0:
        b.set_lognormal_k(sigma, 12345)
    for v in variants:
        for k in ("PE_MG_OMEGA2", "PE_CMG_OMEGA", "PE_CMG_OMEGA2", "PE_FACE_OMEGA"):
            b.set_tuning(k[3:].lower(), float(v.get(k, -1.0)))
        b.set_solution(np.zeros(s.n_dofs))
        b.assemble_system(0.1)
        b.rel_tol = 1e-10; b.max_iterations = 2000
        b.solve_time_step(want_solution=False, check=False)
        print("%s %-90s it=%4d conv=%d solve %.1f ms" % (name, str(v), b.stats.iterations, b.stats.converged, b.stats.seconds * 1e3), flush=True)
    b.close()
