#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path (BASELINE.json metric: cells assembled/s and seconds
per time step (assemble + solve)).

A "step" is one backward-Euler time step of the reference's run() loop on BASELINE cfg 2
(PoroElasBR1WG, 2-D unit square, 1024 x 1024 quads, coscos data, dt = 0.1):
    assemble_system(t)   -> K1+K2+K3 CUDA kernels (matrix AND rhs rebuilt, as the reference does)
    solve_time_step()    -> preconditioned MINRES on the CUDA CSR SpMV
`value`  = cells * steps / device time, old_solution resident in HBM (old_solution = solution).
`e2e`    = the same through the C ABI with HOST buffers: old_solution copied host->device and the
           solution copied device->host inside the timed region, every step.
`--impl reference` times the CPU oracle (literal restatement of the reference loops + sparse LU:
the reference itself needs deal.II 9.1 + UMFPACK, which cannot be installed here) on a bounded
sample of the same workload (a smaller mesh of the same family) with the same metric.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "cells assembled/s per time step (assemble+solve), BR1-WG"
UNIT = "cells/s"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag, self.proc = index, [], False, None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(",")])
                if self.stop_flag:
                    break
        except Exception:
            pass

    def finish(self):
        self.stop_flag = True
        if self.proc:
            try:
                self.proc.terminate()
            except Exception:
                pass
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for k, nm in enumerate(names):
                if len(r) > 3 + k and r[3 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


def cpu_baseline(n_refine_cpu, steps, dt):
    """Oracle (CPU port of the reference loops) + SuperLU per step, 1 thread, on a smaller mesh of the
    same family.  Returns (cells/s, seconds, description)."""
    import numpy as np
    import scipy.sparse.linalg as spla
    from oracle_lib import BR1_WG, CASE_COSCOS, Oracle
    o = Oracle(BR1_WG, n_refine_cpu)
    o.set_material(1, 1, 1, 0, dt)
    o.set_case(CASE_COSCOS)
    old = np.zeros(o.n_dofs)
    t0 = time.perf_counter()
    t_asm = 0.0
    for k in range(steps):
        v, r, sec = o.assemble(dt * (k + 1), old)
        t_asm += sec
        old = spla.splu(o.csr(v).tocsc()).solve(r)   # factor + solve every step (BR1new:1181-1183)
    total = time.perf_counter() - t0
    N = 1 << n_refine_cpu
    return o.n_cells * steps / total, total, ("oracle port, %dx%d BR1-WG, %d steps: assemble %.2fs + LU %.2fs"
                                              % (N, N, steps, t_asm, total - t_asm))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np
    ncpu = args.cpu_refine
    steps = max(1, args.steps)
    for _ in range(max(0, min(args.warmup, 1))):
        cpu_baseline(max(ncpu - 2, 2), 1, 0.1)
    val, sec, desc = cpu_baseline(ncpu, steps, 0.1)
    N = 1 << args.n_refine
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": args.warmup, "ms_per_step": sec / steps * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "PoroElasBR1WG 2D unit square %dx%d quads, coscos, dt=0.1" % (N, N),
                       "sample": desc},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": 1, "kind": "port", "sample": desc},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--n-refine", type=int, default=0,
                    help="mesh = 2^n x 2^n cells; default: BASELINE configs[1] (1024^2, n = 10) on one GPU, configs[2] "
                         "(4096^2 partitioned across 2/4/8 B200, n = 12) on several")
    ap.add_argument("--cpu-refine", type=int, default=7, help="mesh of the bounded CPU sample")
    ap.add_argument("--rel-tol", type=float, default=1e-10)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--solver-options", type=int, default=1, help="pe_set_solver_options bit field")
    args = ap.parse_args()
    if args.n_refine <= 0:
        args.n_refine = 10 if args.gpus <= 1 and int(os.environ.get("WORLD_SIZE", "1")) <= 1 else 12
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import poroelasticity_b200 as pb

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    verbose = os.environ.get("PE_BENCH_VERBOSE") == "1"

    def say(msg):
        if verbose:
            print("[rank %d %.1fs] %s" % (rank, time.perf_counter() - t_start, msg), file=sys.stderr, flush=True)

    t_start = time.perf_counter()
    dt = 0.1
    b = pb.PoroElasBR1WG(device=local_rank, rank=rank, world_size=world)
    if world > 1:
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt = torch.frombuffer(bytearray(pb.comm_unique_id()), dtype=torch.uint8).cuda()
        dist.broadcast(idt, 0)
        b.comm_init(idt.cpu().numpy().tobytes())
    t_setup = time.perf_counter()
    s = b.make_grid_and_dofs(args.n_refine)
    t_setup = time.perf_counter() - t_setup
    say("setup done: owned rows %d nnz %d ghost cells %d" % (s.n_owned_rows, s.nnz_owned, s.n_ghost_cells))
    b.time_step = dt
    b.set_material(1.0, 1.0, 1.0, 0.0)
    b.set_case(pb.PE_CASE_COSCOS)
    b.rel_tol = args.rel_tol
    b.set_solver_options(args.solver_options)
    n_cells, n_dofs = s.n_cells, s.n_dofs
    N = 1 << args.n_refine

    # end-to-end arm: the host program's buffers are pinned (as a host code that drives a GPU library would
    # allocate them), so the H2D / D2H copies inside the C ABI calls run at PCIe speed
    pinned_sol = torch.zeros(s.n_owned_rows, dtype=torch.float64).pin_memory().numpy()

    def step(k, host_io, old_host):
        t = dt * (k + 1)
        if host_io:
            b.assemble_system(t, old_host)       # H2D of old_solution inside
            x = b.solve_time_step(want_solution=True, check=False, out=pinned_sol)   # D2H of the solution inside
            return x
        b.assemble_system(t)                      # old_solution resident on the device
        b.solve_time_step(want_solution=False, check=False)
        return None

    # ---- device-resident arm ------------------------------------------------------------------
    b.set_solution(np.zeros(s.n_owned_rows if world == 1 else n_dofs))
    say("device state ready")
    for k in range(args.warmup):
        step(k, False, None)
        say("warmup step %d: %d iterations" % (k, b.stats.iterations))
    b.set_solution(np.zeros(s.n_owned_rows if world == 1 else n_dofs))
    launches0 = b.timings()["launches"]
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    asm_ms, asm_k_ms, sol_ms, iters, spmvs = [], [], [], [], []
    for k in range(args.steps):
        step(k, False, None)
        tm = b.timings()
        asm_ms.append(tm["assemble_ms"])
        asm_k_ms.append(tm["assemble_kernel_ms"])
        sol_ms.append(tm["solve_ms"])
        iters.append(b.stats.iterations)
        spmvs.append(b.stats.spmv_count)
    barrier()
    say("timed steps done")
    wall = time.perf_counter() - t0
    # device time of the steps = sum of the CUDA-event timed phases on the handle's stream
    dev_s = (sum(asm_ms) + sum(sol_ms)) * 1e-3
    tt = torch.tensor([dev_s, wall], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dev_s, wall = float(tt[0]), float(tt[1])
    launches = b.timings()["launches"] - launches0
    clocks = sampler.finish() if rank == 0 else None
    converged = bool(b.stats.converged)
    rel_res = b.stats.rel_residual

    # ---- dominant kernel (SpMV) in isolation, CUDA events on the library's stream -----------------
    hbm, peak_src = peaks()
    # the solver's SpMV runs on its own operator: the three-field matrix K3 in a pruned CSR (pe_k3.h);
    # algorithmic bytes = 12 B per stored entry + 20 B per row of THAT matrix
    spmv_ms = None
    try:
        spmv_ms = b.bench_spmv(20)
    except Exception:
        spmv_ms = None
    fp64_tflops = None
    try:
        fp64_tflops = b.bench_fp64()
    except Exception:
        fp64_tflops = None
    s3 = b.sizes()
    nnz_owned, rows_owned = (s3.solver_nnz, s3.solver_rows) if s3.solver_nnz > 0 else (s.nnz_owned, s.n_owned_rows)
    spmv_bytes = 12.0 * nnz_owned + 20.0 * rows_owned

    # ---- end-to-end arm: host buffers through the C ABI -------------------------------------------
    # every step: old_solution host -> device inside pe_assemble_system, solution device -> host inside
    # pe_solve_time_step.
    # Host side of the N-GPU program: one process per GPU on ONE node, so old_solution (n_dofs doubles, global
    # order) lives in a POSIX shared-memory segment that every rank maps and registers as pinned memory.  Per
    # step every rank copies its owned part D2H into its pinned buffer, writes it into its three contiguous
    # global ranges of the shared vector, and after a barrier reads its owned + halo entries from it (H2D inside
    # pe_assemble_system).  No rank can leave the solve (NCCL collectives) before all ranks have finished reading.
    shm = None
    if world == 1:
        old = torch.zeros(n_dofs, dtype=torch.float64).pin_memory().numpy()
        pinned_sol = torch.zeros(n_dofs, dtype=torch.float64).pin_memory().numpy()
    else:
        from multiprocessing import shared_memory
        names = [None]
        if rank == 0:
            shm = shared_memory.SharedMemory(create=True, size=n_dofs * 8)
            names[0] = shm.name
        dist.broadcast_object_list(names, src=0)
        if rank != 0:
            shm = shared_memory.SharedMemory(name=names[0])
            try:  # Python < 3.13 also registers ATTACHED segments with the resource tracker: rank 0 owns the unlink
                from multiprocessing import resource_tracker
                resource_tracker.unregister(shm._name, "shared_memory")
            except Exception:
                pass
        old = np.ndarray((n_dofs,), dtype=np.float64, buffer=shm.buf)
        if rank == 0:
            old[:] = 0.0
        rc = torch.cuda.cudart().cudaHostRegister(old.ctypes.data, n_dofs * 8, 0)
        say("cudaHostRegister of the shared old_solution: %s" % (rc,))
        rb, re_ = b.owned_ranges()
        dist.barrier()

    def e2e_step(k):
        nonlocal old, pinned_sol
        x = step(k, True, old)
        if world == 1:
            old, pinned_sol = pinned_sol, old      # old_solution = solution (BR1new:2241): swap the host buffers
        else:
            o = 0
            for q in range(3):
                n = int(re_[q] - rb[q])
                old[int(rb[q]):int(re_[q])] = x[o:o + n]
                o += n
            dist.barrier()

    e2e_step(0)
    barrier()
    t0 = time.perf_counter()
    for k in range(args.steps):
        e2e_step(k)
    barrier()
    e2e_s = time.perf_counter() - t0
    tt = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    e2e_s = float(tt[0])
    e2e = {"value": n_cells * args.steps / e2e_s, "unit": UNIT,
           "h2d_bytes_per_step": int(n_dofs * 8),  # owned parts of old_solution summed over the ranks (halo entries not counted)
           "d2h_bytes_per_step": int(n_dofs * 8), "ms_per_step": e2e_s / args.steps * 1e3}

    # orderly shutdown on every rank (the library's NCCL communicator, then torch's)
    stats_launches = int(launches)
    b.close()
    if world > 1:
        dist.barrier()
        torch.cuda.cudart().cudaHostUnregister(old.ctypes.data)
        del old
        shm.close()
        if rank == 0:
            shm.unlink()
        dist.destroy_process_group()
    if rank != 0:
        return
    asm_kernel_s = (sum(asm_k_ms) / len(asm_k_ms)) * 1e-3
    # per GPU: every rank assembles n_cells / world cells (+ its ghost layer) in the same time
    roof_asm = {"bound": "hbm", "kernel": "k_assemble_interior<BR1>", "achieved": 1904.0 * n_cells / world / asm_kernel_s / 1e9,
                "peak": hbm, "unit": "GB/s", "frac": 1904.0 * n_cells / world / asm_kernel_s / 1e9 / hbm, "traffic": None,
                "bytes_per_cell": 1904, "cells_per_s_per_gpu": n_cells / world / asm_kernel_s, "peak_source": peak_src}
    roof = None
    if spmv_ms:
        ach = spmv_bytes / (spmv_ms * 1e-3) / 1e9
        # DRAM bytes of one launch from the ncu --set full capture of this kernel on this workload
        # (profiles/r01g_ncu_full_summary.md: 2.082 GB read + 65.4 MB written); other workloads: not captured
        traffic = 2.147e9 if (world == 1 and args.n_refine == 10) else None
        roof = {"bound": "hbm", "kernel": "k_spmv_stream (three-field solver matrix, %d rows, %d entries)" % (rows_owned, nnz_owned), "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm,
                "traffic": traffic, "traffic_source": "ncu dram__bytes_read.sum + dram__bytes_write.sum, profiles/r01g_ncu_full_summary.md",
                "bytes_per_launch": spmv_bytes, "ms_per_launch": spmv_ms, "peak_source": peak_src}
    line = {"metric": METRIC, "value": n_cells * args.steps / dev_s, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_s / args.steps * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "PoroElasBR1WG 2D unit square %dx%d quads, coscos, dt=0.1, backward Euler" % (N, N),
                       "n_dofs": n_dofs, "nnz": s.nnz if world == 1 else None, "rel_tol": args.rel_tol,
                       "l2": "inputs (GB-sized CSR arrays) exceed L2; no flush needed", "setup_s": t_setup,
                       "scaling_note": "BASELINE.json configs: [1] 1024^2 on 1 B200, [2] 4096^2 partitioned across 2/4/8 B200; "
                                       "strong scaling holds among N = 2, 4, 8 (same mesh)"},
            "s_per_time_step": dev_s / args.steps, "assemble_ms": sum(asm_ms) / len(asm_ms),
            "assembly_cells_per_s": n_cells / asm_kernel_s, "solve_ms": sum(sol_ms) / len(sol_ms),
            "solver": {"method": "MINRES on the three-field (u, total pressure, p) form + block preconditioner (bubble Jacobi x Q1 V-cycle; cell mass; p_int elimination + face V-cycle)", "iterations": iters, "spmv": spmvs,
                       "converged": converged, "rel_residual": rel_res},
            "wall_s": wall, "gpu_launches": int(launches), "clocks": clocks, "e2e": e2e,
            "roofline": roof if roof else roof_asm, "roofline_assembly": roof_asm,
            "fp64_peak_measured_tflops": fp64_tflops}
    if not args.no_cpu:
        val, sec, desc = cpu_baseline(args.cpu_refine, 2, dt)
        line["cpu_baseline"] = {"value": val, "unit": UNIT, "cores": 1, "kind": "port", "sample": desc}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
