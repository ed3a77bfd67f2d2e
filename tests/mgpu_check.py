"""Multi-GPU parity check, run under torchrun (one process per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/mgpu_check.py
Every rank assembles its owned rows (ghost cells recomputed, no communication), the distributed MINRES
exchanges halo dofs with NCCL and all-reduces dot products.  Owned rows of matrix / rhs / solution are
compared with the CPU oracle."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    import scipy.sparse.linalg as spla
    import poroelasticity_b200 as pb
    from oracle_lib import BR1_WG, CGQ1_WG, CASE_COSCOS, Oracle

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    nref = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    ok = True
    for el in (BR1_WG, CGQ1_WG):
        dt = 0.1
        o = Oracle(el, nref, distort=0.1)
        o.set_material(1.0, 1.0, 1.0, 0.01, dt)
        o.set_case(CASE_COSCOS)
        b = pb.BiotSystem(element=el, device=local, rank=rank, world_size=world)
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt = torch.frombuffer(bytearray(pb.comm_unique_id()), dtype=torch.uint8).cuda()
        dist.broadcast(idt, 0)
        b.comm_init(idt.cpu().numpy().tobytes())
        s = b.make_grid_and_dofs(nref, distort=0.1)
        b.time_step = dt
        b.set_material(1.0, 1.0, 1.0, 0.01)
        b.set_case(pb.PE_CASE_COSCOS)
        # owned global rows of this rank: rows first touched by owned cells
        cd = o.cell_dofs()
        first = np.full(o.n_dofs, o.n_cells, np.int64)
        for c in range(o.n_cells - 1, -1, -1):
            first[cd[c]] = c
        owned = np.nonzero((first >= s.cell_begin) & (first < s.cell_end))[0]
        assert owned.size == s.n_owned_rows
        rpo, cio = o.pattern()
        old = np.zeros(o.n_dofs)
        t = dt
        for step in range(2):
            v_ref, r_ref, _ = o.assemble(t, old)
            x_ref = spla.splu(o.csr(v_ref).tocsc()).solve(r_ref)
            b.assemble_system(t, old if step == 0 else None)
            v, r = b.system_matrix(), b.system_rhs()
            vref_owned = np.concatenate([v_ref[rpo[g]:rpo[g + 1]] for g in owned])
            em = np.abs(v - vref_owned).max() / np.abs(v_ref).max()
            er = np.abs(r - r_ref[owned]).max() / np.abs(r_ref).max()
            b.rel_tol = 1e-12
            x = b.solve_time_step()
            es = np.linalg.norm(x - x_ref[owned]) / np.linalg.norm(x_ref[owned])
            good = em < 1e-12 and er < 1e-12 and es < 1e-8 and b.stats.converged
            ok &= bool(good)
            print("rank %d el %d step %d: matrix %.1e rhs %.1e solution %.1e its %d ghost_cells %d %s"
                  % (rank, el, step, em, er, es, b.stats.iterations, s.n_ghost_cells, "OK" if good else "FAIL"), flush=True)
            old = x_ref
            t += dt
        b.close()
    # ---- distributed grid levels (row slabs) against the replicated V-cycles: the two are the same operator, so
    # MINRES must take the same number of iterations and return the same solution (N = 256: two distributed levels)
    nref2 = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    sols, its = [], []
    for replicated in (True, False):
        b = pb.BiotSystem(element=BR1_WG, device=local, rank=rank, world_size=world)
        b.set_tuning("mg_dist_min_n", 128)   # distribute the 256^2 and 128^2 levels of this small mesh
        b.set_tuning("mg_replicated", 1 if replicated else 0)
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt = torch.frombuffer(bytearray(pb.comm_unique_id()), dtype=torch.uint8).cuda()
        dist.broadcast(idt, 0)
        b.comm_init(idt.cpu().numpy().tobytes())
        b.make_grid_and_dofs(nref2, distort=0.1)
        b.time_step = 0.1
        b.set_material(100.0, 1.0, 1.0, 0.01)
        b.set_lognormal_k(1.0, 7)
        b.set_case(pb.PE_CASE_COSCOS)
        b.assemble_system(0.1)
        b.rel_tol = 1e-10
        sols.append(b.solve_time_step())
        its.append((b.stats.iterations, bool(b.stats.converged), b.stats.rel_residual))
        b.close()
    ed = np.linalg.norm(sols[0] - sols[1]) / np.linalg.norm(sols[0])
    good = its[0][1] and its[1][1] and abs(its[0][0] - its[1][0]) <= 4 and ed < 1e-8
    ok &= bool(good)
    print("rank %d distributed vs replicated V-cycles (N=%d): iterations %d / %d, solution difference %.1e %s"
          % (rank, 1 << nref2, its[0][0], its[1][0], ed, "OK" if good else "FAIL"), flush=True)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    if rank == 0:
        print("MGPU_CHECK", "PASS" if int(flag) == 1 else "FAIL")
    sys.exit(0 if int(flag) == 1 else 1)


if __name__ == "__main__":
    main()
