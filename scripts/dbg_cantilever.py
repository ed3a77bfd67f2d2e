import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, scipy.sparse.linalg as spla
import poroelasticity_b200 as pb
from oracle_lib import *
lam, mu, alpha, c0, dt, kperm = 1e6 / 7, 1e6 / 28, 0.93, 0.0, 0.01, 1e-7
nref = 4
o = Oracle(BR1_WG, nref, bc_mode=1); o.set_material(lam, mu, alpha, c0, dt); o.set_case(CASE_ZERO); o.set_kcell(np.full(o.n_cells, kperm))
for rec, kry, ig in ((4, 1, 2), (0, 1, 0), (0, 0, 0)):
    b = pb.BiotSystem(element=BR1_WG); b.make_grid_and_dofs(nref); b.set_face_kinds(o.mesh()[3], traction=(0.0, -1.0))
    b.time_step = dt; b.set_material(lam, mu, alpha, c0, kperm); b.set_case(pb.PE_CASE_ZERO)
    b.set_recycling(rec); b.set_krylov(kry); b.set_initial_guess(ig)
    b.rel_tol, b.max_iterations = 1e-12, 4000
    old = np.zeros(o.n_dofs)
    for step in range(2):
        t = dt * (step + 1)
        v_ref, r_ref, _ = o.assemble(t, old)
        A = o.csr(v_ref)
        x_ref = spla.splu(A.tocsc()).solve(r_ref)
        b.assemble_system(t)
        x = b.solve_time_step(check=False)
        nu = o.n_u
        print(rec, kry, step, "conv", b.stats.converged, "its", b.stats.iterations, "restarts", b.stats.restarts, "rel", b.stats.rel_residual,
              "eu %.2e ep %.2e" % (np.linalg.norm(x[:nu] - x_ref[:nu]) / np.linalg.norm(x_ref[:nu]), np.linalg.norm(x[nu:] - x_ref[nu:]) / np.linalg.norm(x_ref[nu:])),
              "anorm %.2e xnorm %.2e bnorm %.2e" % (np.abs(A).sum(axis=1).max(), np.linalg.norm(x_ref), np.linalg.norm(r_ref)), flush=True)
        old = x_ref
    b.close()
