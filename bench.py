#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path (BASELINE.json metric: cells assembled/s and seconds
per time step (assemble + solve)).

A "step" is one backward-Euler time step of the reference's run() loop (PoroElasBR1WG_new.cc:2156-2243)
on BASELINE cfg 2 (PoroElasBR1WG, 2-D unit square, 1024 x 1024 quads, coscos data, dt = 0.1):
    assemble_system(t)   -> K1+K2+K3 CUDA kernels (matrix AND rhs rebuilt, as the reference does)
    solve_time_step()    -> preconditioned MINRES on the CUDA CSR SpMV
The time loop is the reference's: `for (time = dt; time <= 1; time += dt)` with a float-accumulated bound
(BR1new:2181: 10 steps at dt = 0.1); when --steps exceeds one run() the loop starts again from
old_solution = 0 (BR1new:2163).  Every timed step must converge: the line carries per-step iterations,
residuals and flags, and the program exits non-zero if any timed step did not converge.

`value`  = cells * steps / device time, old_solution resident in HBM (old_solution = solution).
`e2e`    = the same through the C ABI with HOST buffers: old_solution copied host->device and the
           solution copied device->host inside the timed region, every step.
`--impl reference` times the CPU oracle (literal restatement of the reference loops + sparse LU with the
fill-reducing ordering on A^T + A that UMFPACK's symmetric strategy uses; the reference itself needs deal.II 9.1 +
UMFPACK, which cannot be installed here) on bounded samples of the same workload: meshes 128^2, 256^2, 512^2 of
the same family, 1-2 steps each (direct LU at 1024^2 takes hours).  The GPU arm reports `value_at_reference_n` on
those same meshes so that a same-N ratio exists.
`--config 4` / `--config 5`: BASELINE configs[3] (near-incompressible, log-normal K: iterations, residual
history, s/step) and configs[4] (local-matrix microbenchmark, cells/s + rooflines); lines kept in profiles/.
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
DEFAULT_REL_TOL = 1e-10   # validated: tests/test_gpu_parity.py::test_bench_tolerance_gives_1e8_solution_parity + solution_check below
REF_MESHES = (7, 8, 9)    # 128^2, 256^2, 512^2: what the CPU arm can factorise in minutes


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel, n_refine, world):
    """DRAM bytes per launch of `kernel` from the committed ncu --set full capture (profiles/ncu_traffic.json:
    dram__bytes_read.sum + dram__bytes_write.sum), or None when that workload was not captured."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(p):
        return None, None
    for e in json.load(open(p)).get("captures", []):
        if e["kernel"] == kernel and e["n_refine"] == n_refine and e["n_gpus"] == world:
            return float(e["dram_bytes_per_launch"]), e["source"]
    return None, None


def cpu_info():
    model = "unknown"
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                model = line.split(":", 1)[1].strip()
                break
    except Exception:
        pass
    return {"nproc": os.cpu_count(), "model": model}


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


def run_times(dt, t_end=1.0):
    """the reference's time loop: for (time = dt; time <= t_end; time += dt), float accumulated (BR1new:2181)"""
    out, t = [], dt
    while t <= t_end:
        out.append(t)
        t += dt
    return out


def workload_config(n_refine, rel_tol):
    """the same dict in both arms (the reference arm times bounded samples of this workload)"""
    N = 1 << n_refine
    n_u, n_pint, n_pface = 2 * (N + 1) ** 2 + 2 * N * (N + 1), N * N, 2 * N * (N + 1)
    return {"workload": "PoroElasBR1WG 2D unit square %dx%d quads, coscos, K=I, lambda=mu=alpha=1, c0=0, dt=0.1, backward Euler, "
                        "reference run() loop (time <= 1, old_solution = 0 at each start)" % (N, N),
            "n_cells": N * N, "n_dofs": n_u + n_pint + n_pface, "nnz": 221 * N * N + 64 * N + 4, "rel_tol": rel_tol,
            "l2": "inputs (GB-sized CSR arrays) exceed L2; no flush needed",
            "reference_arm_sample": "CPU arm: oracle port + sparse LU, 1 core, on meshes 128^2 / 256^2 / 512^2 of the same "
                                    "family (LU at 1024^2 takes hours); GPU arm reports value_at_reference_n on the same meshes"}


# ---------------------------------------------------------------------------------------------------------
# CPU arm (oracle port + sparse LU); the only place besides cpu_baseline where bench.py executes oracle/
# ---------------------------------------------------------------------------------------------------------
def lu_solve(A_csr, rhs):
    """SparseDirectUMFPACK stand-in (BR1new:1181-1183): SuperLU, factor + solve every step.  permc_spec
    MMD_AT_PLUS_A = minimum degree on the pattern of A^T + A, the ordering family UMFPACK's symmetric strategy
    picks for these structurally symmetric matrices (3x less fill and time than scipy's COLAMD default)."""
    import scipy.sparse.linalg as spla
    return spla.splu(A_csr.tocsc(), permc_spec="MMD_AT_PLUS_A").solve(rhs)


def cpu_run(n_refine, steps, dt):
    """`steps` steps of the reference loop on the CPU oracle.  Returns a dict with cells/s and the phase times."""
    import numpy as np
    from oracle_lib import BR1_WG, CASE_COSCOS, Oracle
    o = Oracle(BR1_WG, n_refine)
    o.set_material(1, 1, 1, 0, dt)
    o.set_case(CASE_COSCOS)
    old = np.zeros(o.n_dofs)
    times = run_times(dt)
    t0 = time.perf_counter()
    t_asm = 0.0
    for k in range(steps):
        v, r, sec = o.assemble(times[k % len(times)], old)
        t_asm += sec
        old = lu_solve(o.csr(v), r)
    total = time.perf_counter() - t0
    N = 1 << n_refine
    return {"n": N, "steps": steps, "value": o.n_cells * steps / total, "ms_per_step": total / steps * 1e3,
            "assemble_s_per_step": t_asm / steps, "lu_s_per_step": (total - t_asm) / steps,
            "desc": "oracle port + SuperLU(MMD_AT_PLUS_A), %dx%d BR1-WG, %d step(s): assemble %.2fs + LU %.2fs"
                    % (N, N, steps, t_asm, total - t_asm)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    budget = args.ref_budget_s
    t_begin = time.perf_counter()
    for _ in range(max(0, min(args.warmup, 1))):
        cpu_run(5, 1, 0.1)
    per_n, est = [], None
    for nref in REF_MESHES:
        left = budget - (time.perf_counter() - t_begin)
        if est is not None and est > left:
            break
        steps = 1 if est is None else max(1, min(2, int(0.25 * left / est)))
        r = cpu_run(nref, min(steps, max(1, args.steps)), 0.1)
        per_n.append(r)
        # next mesh: 4x the cells; the LU part grows ~8.3x per refinement (measured 64^2 -> 128^2), assembly 4x
        est = 4.0 * r["assemble_s_per_step"] + 8.3 * r["lu_s_per_step"]
    top = per_n[-1]
    ci = cpu_info()
    cfg = workload_config(args.n_refine, args.rel_tol)
    line = {"impl": "reference", "metric": METRIC, "value": top["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": top["ms_per_step"], "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
            "sample_n": top["n"], "steps_timed_at_sample_n": top["steps"],
            "per_n": [{k: r[k] for k in ("n", "steps", "value", "ms_per_step", "assemble_s_per_step", "lu_s_per_step")}
                      for r in per_n],
            "cpu_baseline": {"value": top["value"], "unit": UNIT, "cores": 1, "kind": "port", "sample": top["desc"],
                             "nproc": ci["nproc"], "cpu_model": ci["model"],
                             "note": "the reference is a serial program (no threads, no MPI); value = the largest mesh timed"},
            "e2e": {"value": top["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------
# BASELINE configs[3] and configs[4]
# ---------------------------------------------------------------------------------------------------------
def run_cfg4(args):
    """near-incompressible (lambda = 1e6, c0 = 1e-6), log-normal K = exp(2 Z): solver-iteration stress test"""
    import numpy as np
    import poroelasticity_b200 as pb
    nref = args.n_refine
    N = 1 << nref
    b = pb.PoroElasBR1WG(device=0)
    s = b.make_grid_and_dofs(nref)
    b.time_step = 0.1
    b.set_material(1e6, 1.0, 1.0, 1e-6)
    b.set_lognormal_k(2.0, 12345)
    b.set_case(pb.PE_CASE_COSCOS)
    b.rel_tol = args.rel_tol
    b.set_solver_options(args.solver_options)
    b.set_krylov(args.krylov, args.gmres_m)
    b.set_tuning("reuse_operator", 1 if args.reuse_operator else 0)
    b.set_solution(np.zeros(s.n_owned_rows))
    times = run_times(0.1)
    hist = []
    for k in range(args.steps):
        t = times[k % len(times)]
        b.assemble_system(t)
        b.solve_time_step(want_solution=False, check=False)
        tm = b.timings()
        hist.append({"t": t, "iterations": b.stats.iterations, "converged": int(b.stats.converged),
                     "rel_residual": b.stats.rel_residual, "restarts": b.stats.restarts,
                     "assemble_ms": tm["assemble_ms"], "solve_ms": tm["solve_ms"]})
    ms = [h["assemble_ms"] + h["solve_ms"] for h in hist]
    line = {"config": 4, "metric": "s per time step (assemble+solve) and MINRES iterations, BR1-WG cfg 4",
            "workload": "PoroElasBR1WG 2D %dx%d, lambda=1e6, mu=1, alpha=1, c0=1e-6, K=exp(2 Z) per cell (seed 12345), coscos "
                        "forcing, dt=0.1" % (N, N),
            "n_gpus": 1, "rel_tol": args.rel_tol, "krylov": "FGMRES(%d)" % args.gmres_m if args.krylov == 1 else "MINRES", "steps": hist, "s_per_step_mean": sum(ms) / len(ms) * 1e-3,
            "iterations_mean": sum(h["iterations"] for h in hist) / len(hist),
            "all_converged": all(h["converged"] for h in hist),
            "cells_per_s": s.n_cells / (sum(ms) / len(ms) * 1e-3)}
    b.close()
    print(json.dumps(line))
    return 0 if line["all_converged"] else 1


def run_cfg5(args):
    """EltStiffMat-style batched local matrices only: BR1-WG, CGQ1-WG, WG(Q2,Q2;RT[2]) on a distorted mesh with
    log-normal K; cells/s and both roofline fractions.  The dense matrices are written to HBM in chunks
    (8192^2 x 289 x 8 B = 155 GB does not fit beside the mesh) and each chunk buffer is overwritten by
    the next chunk; the checksum pass sums every entry once.
    flop_per_cell_estimate is the LEAN count of the algorithm, not the executed (padded) count: Q2/RT2 = symmetric Gram
    300 x 32 MAC + Cholesky 24^3/3 + substitution 21 x 24^2 + symmetric Y^T Y 231 x 24 MAC = 47 kflop."""
    import poroelasticity_b200 as pb
    hbm, peak_src = peaks()
    nref = args.n_refine
    N = 1 << nref
    out = []
    fp64 = None
    for el, name, n_loc, flop in ((pb.PE_BR1_WG, "BR1-WG 17x17", 17, 9.0e3), (pb.PE_CGQ1_WG, "CGQ1-WG 13x13", 13, 5.0e3),
                                  (pb.PE_WG_Q2RT2, "WG(Q2,Q2;RT[2]) 21x21", 21, 47e3)):
        b = pb.BiotSystem(element=el, device=0)
        b.make_grid(nref, distort=0.2, seed=12345)
        b.set_material(1.0, 1.0, 1.0, 0.0)
        b.set_lognormal_k(2.0, 12345)
        b.time_step = 0.1
        chunk = args.chunk_cells
        cs, sec = b.local_matrices_bench(chunk, repeats=args.steps)
        if fp64 is None:
            fp64 = b.bench_fp64()
        bytes_cell = 64 + 8 + 8 * n_loc * n_loc
        cps = N * N / sec
        out.append({"element": name, "cells_per_s": cps, "seconds": sec, "checksum": cs, "bytes_per_cell": bytes_cell,
                    "hbm_gbs": cps * bytes_cell / 1e9, "hbm_frac": cps * bytes_cell / 1e9 / hbm,
                    "flop_per_cell_estimate": flop, "fp64_tflops": cps * flop / 1e12,
                    "fp64_frac": cps * flop / 1e12 / fp64 if fp64 else None})
        b.close()
    line = {"config": 5, "metric": "cells/s, local matrices only (EltStiffMat microbenchmark)",
            "workload": "%dx%d cells (%d), vertices displaced <= 0.2 h (seed 12345), K = exp(2 Z) per cell; dense local matrices "
                        "written entry-major in chunks of %d cells" % (N, N, N * N, args.chunk_cells),
            "n_gpus": 1, "repeats": args.steps, "hbm_peak_gbs": hbm, "peak_source": peak_src,
            "fp64_peak_measured_tflops": fp64, "elements": out}
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------------------
def parity_check(pb, torch, dist, rank, world, local_rank):
    """Small-mesh comparison with the CPU oracle on THIS rank layout, before the timed region (the driver's GPU test
    box has one GPU, so this is where the N-GPU answer is checked): matrix / rhs of the owned rows to 1e-12,
    per-step solution to 1e-8 against sparse LU.  The oracle is the checker here, never the thing measured."""
    import numpy as np
    from oracle_lib import BR1_WG, CASE_COSCOS, Oracle
    nref, dt = 5, 0.1
    o = Oracle(BR1_WG, nref, distort=0.1)
    o.set_material(1.0, 1.0, 1.0, 0.01, dt)
    o.set_case(CASE_COSCOS)
    b = pb.PoroElasBR1WG(device=local_rank, rank=rank, world_size=world)
    if world > 1:
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt = torch.frombuffer(bytearray(pb.comm_unique_id()), dtype=torch.uint8).cuda()
        dist.broadcast(idt, 0)
        b.comm_init(idt.cpu().numpy().tobytes())
    s = b.make_grid_and_dofs(nref, distort=0.1)
    b.time_step = dt
    b.set_material(1.0, 1.0, 1.0, 0.01)
    b.set_case(pb.PE_CASE_COSCOS)
    owned = b.owned_rows()
    rpo, _ = o.pattern()
    old = np.zeros(o.n_dofs)
    em = er = es = 0.0
    conv = True
    its = []
    for step in range(2):
        t = dt * (step + 1)
        v_ref, r_ref, _ = o.assemble(t, old)
        x_ref = lu_solve(o.csr(v_ref), r_ref)
        b.assemble_system(t, old if step == 0 else None)
        v, r = b.system_matrix(), b.system_rhs()
        vref_owned = np.concatenate([v_ref[rpo[g]:rpo[g + 1]] for g in owned])
        em = max(em, float(np.abs(v - vref_owned).max() / np.abs(v_ref).max()))
        er = max(er, float(np.abs(r - r_ref[owned]).max() / np.abs(r_ref).max()))
        b.rel_tol = 1e-12
        x = b.solve_time_step(check=False)
        conv = conv and bool(b.stats.converged)
        its.append(int(b.stats.iterations))
        num = torch.tensor([float(np.sum((x - x_ref[owned]) ** 2)), float(np.sum(x_ref[owned] ** 2))],
                           dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(num)
        es = max(es, float(torch.sqrt(num[0] / num[1])))
        old = x_ref
    b.close()
    tt = torch.tensor([em, er, 0.0 if conv else 1.0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    em, er, bad = float(tt[0]), float(tt[1]), float(tt[2])
    ok = em < 1e-12 and er < 1e-12 and es < 1e-8 and bad == 0.0
    return {"mesh": "32x32 BR1-WG, vertices displaced 0.1 h, c0=0.01, 2 steps, partitioned over %d rank(s)" % world,
            "oracle": "CPU restatement + sparse LU", "matrix_max_rel": em, "rhs_max_rel": er, "solution_rel_l2": es,
            "iterations": its, "converged": bad == 0.0, "tolerances": {"matrix": 1e-12, "rhs": 1e-12, "solution": 1e-8}, "ok": ok}


def gpu_at_reference_n(pb, torch, n_refine, rel_tol, options, local_rank, krylov=(1, 16)):
    """the GPU arm on a mesh the CPU arm can also run: one run() of 10 steps, device-resident and end to end"""
    import numpy as np
    dt = 0.1
    b = pb.PoroElasBR1WG(device=local_rank)
    s = b.make_grid_and_dofs(n_refine)
    b.time_step = dt
    b.set_material(1.0, 1.0, 1.0, 0.0)
    b.set_case(pb.PE_CASE_COSCOS)
    b.rel_tol = rel_tol
    b.set_solver_options(options)
    b.set_krylov(*krylov)
    times = run_times(dt)
    zero = np.zeros(s.n_owned_rows)
    for rep in range(2):          # rep 0 = warm-up
        b.set_solution(zero)
        dev_ms, conv = 0.0, True
        for t in times:
            b.assemble_system(t)
            b.solve_time_step(want_solution=False)
            tm = b.timings()
            dev_ms += tm["assemble_ms"] + tm["solve_ms"]
            conv = conv and bool(b.stats.converged)
    old = torch.zeros(s.n_dofs, dtype=torch.float64).pin_memory().numpy()
    new = torch.zeros(s.n_dofs, dtype=torch.float64).pin_memory().numpy()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for t in times:
        b.assemble_system(t, old)
        b.solve_time_step(want_solution=True, out=new)
        old, new = new, old
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    b.close()
    N = 1 << n_refine
    return {"n": N, "steps": len(times), "value": s.n_cells * len(times) / (dev_ms * 1e-3), "ms_per_step": dev_ms / len(times),
            "e2e_value": s.n_cells * len(times) / e2e_s, "e2e_ms_per_step": e2e_s / len(times) * 1e3, "converged": conv}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--config", type=int, default=0,
                    help="0: headline (BASELINE configs[1] on one GPU, configs[2] on several); 4, 5: BASELINE configs[3], [4]")
    ap.add_argument("--n-refine", type=int, default=0,
                    help="mesh = 2^n x 2^n cells; default: BASELINE configs[1] (1024^2, n = 10) on one GPU, configs[2] "
                         "(4096^2 partitioned across 2/4/8 B200, n = 12) on several; config 5: 13 (8192^2)")
    ap.add_argument("--cpu-refine", type=int, default=7, help="mesh of the bounded CPU sample of the GPU arm's cpu_baseline leg")
    ap.add_argument("--ref-budget-s", type=float, default=270.0, help="--impl reference: wall-clock budget for the CPU meshes")
    ap.add_argument("--rel-tol", type=float, default=DEFAULT_REL_TOL)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-checks", action="store_true", help="skip parity_check / solution_check / value_at_reference_n")
    ap.add_argument("--chunk-cells", type=int, default=1 << 20, help="config 5: cells per output chunk")
    ap.add_argument("--solver-options", type=int, default=1, help="pe_set_solver_options bit field")
    ap.add_argument("--initial-guess", type=int, default=2, help="pe_set_initial_guess order")
    ap.add_argument("--recycle", type=int, default=4, help="pe_set_recycling: previous corrections kept (0 = off)")
    ap.add_argument("--krylov", type=int, default=1, help="pe_set_krylov method: 1 = FGMRES + block-triangular P, 0 = MINRES")
    ap.add_argument("--gmres-m", type=int, default=16, help="pe_set_krylov restart length")
    ap.add_argument("--reuse-operator", action="store_true",
                    help="run the HEADLINE in cached mode (pe_set_tuning reuse_operator=1: solver operator + preconditioner kept "
                         "while the matrix inputs do not change).  Default: everything is rebuilt in every step, as the "
                         "reference refactorises in every step, and the cached mode is timed in a second pass and reported "
                         "under `cached_mode` (SURVEY 8(f).1: never the headline)")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.n_refine <= 0:
        args.n_refine = 13 if args.config == 5 else (10 if args.gpus <= 1 and world <= 1 else 12)
    if args.impl == "reference":
        return run_reference(args)
    if args.config == 4:
        sys.exit(run_cfg4(args))
    if args.config == 5:
        sys.exit(run_cfg5(args))

    import numpy as np
    import torch
    import poroelasticity_b200 as pb

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dist = None
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
    times = run_times(dt)

    # ---- correctness of this rank layout against the oracle, before anything is timed ---------------------
    parity = None
    if not args.no_checks:
        parity = parity_check(pb, torch, dist, rank, world, local_rank)
        say("parity_check: %s" % (parity,))

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
    b.set_initial_guess(args.initial_guess)
    b.set_krylov(args.krylov, args.gmres_m)
    b.set_recycling(args.recycle)
    b.set_tuning("reuse_operator", 1 if args.reuse_operator else 0)
    n_cells, n_dofs = s.n_cells, s.n_dofs
    N = 1 << args.n_refine
    zero_state = torch.zeros(s.n_owned_rows if world == 1 else n_dofs, dtype=torch.float64).pin_memory().numpy()

    # ---- solution_check: the bench tolerance against a solve to the rounding floor, on THIS mesh --------------
    # three steps of run() at rel_tol and again at 1e-13 (each trajectory feeds on its own old_solution); relative
    # L2 difference of u and of p per step.  north_star: per-step u and p within 1e-8 of the direct solve.
    sol_check = None
    n_u_owned = None
    if not args.no_checks:
        rb, re_ = b.owned_ranges()
        n_u_owned = int(re_[0] - rb[0])
        traj = {}
        for tol in (args.rel_tol, 1e-13):
            b.set_solution(zero_state)
            b.rel_tol = tol
            xs, info = [], []
            for k in range(3):
                b.assemble_system(times[k])
                xs.append(b.solve_time_step(want_solution=True, check=False).copy())
                info.append({"iterations": int(b.stats.iterations), "converged": int(b.stats.converged),
                             "rel_residual": float(b.stats.rel_residual)})
            traj[tol] = (xs, info)
        b.rel_tol = args.rel_tol
        du, dp = [], []
        for k in range(3):
            xa, xb = traj[args.rel_tol][0][k], traj[1e-13][0][k]
            q = torch.tensor([np.sum((xa[:n_u_owned] - xb[:n_u_owned]) ** 2), np.sum(xb[:n_u_owned] ** 2),
                              np.sum((xa[n_u_owned:] - xb[n_u_owned:]) ** 2), np.sum(xb[n_u_owned:] ** 2)],
                             dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(q)
            du.append(float(torch.sqrt(q[0] / q[1])))
            dp.append(float(torch.sqrt(q[2] / q[3])))
        sol_check = {"what": "3 steps of run() at rel_tol vs the same at rel_tol 1e-13 on this mesh: relative L2 difference per step",
                     "u_rel_l2": du, "p_rel_l2": dp, "bench": traj[args.rel_tol][1], "tight": traj[1e-13][1],
                     "bound": 1e-8, "ok": max(du + dp) < 1e-8}
        del traj
        say("solution_check: %s" % (sol_check,))

    def step(t, host_io, old_host, out_host):
        if host_io:
            b.assemble_system(t, old_host)       # H2D of old_solution inside
            return b.solve_time_step(want_solution=True, out=out_host)   # D2H of the solution inside
        b.assemble_system(t)                      # old_solution resident on the device
        b.solve_time_step(want_solution=False)    # raises if the solve does not converge
        return None

    # ---- device-resident arm ------------------------------------------------------------------
    # warm-up: the first steps of a run(); the timed region then starts its own run() at t = dt
    b.set_solution(zero_state)
    for k in range(args.warmup):
        step(times[k % len(times)], False, None, None)
        say("warmup step %d: %d iterations" % (k, b.stats.iterations))
    launches0 = b.timings()["launches"]
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    t0 = time.perf_counter()
    asm_ms, asm_k_ms, sol_ms, per_step = [], [], [], []
    failed = None
    for k in range(args.steps):
        if k % len(times) == 0:
            b.set_solution(zero_state)            # old_solution = 0 (BR1new:2163): a new run()
        try:
            step(times[k % len(times)], False, None, None)
        except pb.PoroElasError as e:
            if e.code != pb._capi.PE_ERR_NOT_CONVERGED:
                raise
            failed = str(e)
        tm = b.timings()
        asm_ms.append(tm["assemble_ms"])
        asm_k_ms.append(tm["assemble_kernel_ms"])
        sol_ms.append(tm["solve_ms"])
        per_step.append({"t": round(times[k % len(times)], 10), "iterations": int(b.stats.iterations),
                         "spmv": int(b.stats.spmv_count), "converged": int(b.stats.converged),
                         "rel_residual": float(b.stats.rel_residual), "restarts": int(b.stats.restarts)})
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
    all_converged = all(p["converged"] for p in per_step)

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
        pinned_sol = None
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

    e2e_conv = True

    def e2e_step(k):
        nonlocal old, pinned_sol, e2e_conv
        if k % len(times) == 0:                   # a new run(): old_solution = 0
            if world == 1:
                old[:] = 0.0
            else:
                for q in range(3):
                    old[int(rb[q]):int(re_[q])] = 0.0
                dist.barrier()
        if world == 1:
            step(times[k % len(times)], True, old, pinned_sol)
            old, pinned_sol = pinned_sol, old      # old_solution = solution (BR1new:2241): swap the host buffers
        else:
            b.assemble_system(times[k % len(times)], old)   # H2D of the owned + halo entries of old_solution inside
            dist.barrier()                                  # every rank has read the shared vector
            b.solve_time_step(want_solution=False)
            b.solution_global(old)                          # D2H of the owned ranges straight into the shared vector
            dist.barrier()
        e2e_conv = e2e_conv and bool(b.stats.converged)

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
           "d2h_bytes_per_step": int(n_dofs * 8), "ms_per_step": e2e_s / args.steps * 1e3, "all_converged": e2e_conv}

    # ---- cached mode, reported separately and never as the headline: the solver operator and its
    # preconditioner are kept while the matrix inputs do not change; matrix and rhs are still assembled in every step
    cached = None
    if not args.reuse_operator:
        try:
            b.set_tuning("reuse_operator", 1)
            b.set_solution(zero_state)
            for k in range(args.warmup):
                step(times[k % len(times)], False, None, None)
            barrier()
            c_ms, c_it, c_ok = 0.0, [], True
            for k in range(args.steps):
                if k % len(times) == 0:
                    b.set_solution(zero_state)
                try:
                    step(times[k % len(times)], False, None, None)
                except pb.PoroElasError as e:
                    if e.code != pb._capi.PE_ERR_NOT_CONVERGED:
                        raise
                tm = b.timings()
                c_ms += tm["assemble_ms"] + tm["solve_ms"]
                c_it.append(int(b.stats.iterations))
                c_ok = c_ok and bool(b.stats.converged)
            barrier()
            tt = torch.tensor([c_ms * 1e-3], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            c_s = float(tt[0])
            cached = {"what": "same K steps with pe_set_tuning(reuse_operator = 1): solver operator + preconditioner rebuilt only when "
                              "the matrix inputs change; system matrix and rhs assembled in every step as in the headline",
                      "ms_per_step": c_s / args.steps * 1e3, "value": n_cells * args.steps / c_s, "unit": UNIT,
                      "iterations": c_it, "all_converged": c_ok}
            say("cached mode: %s" % (cached,))
        except Exception as e:   # the headline above is complete: report the failure instead of losing the line
            cached = {"error": repr(e)}
            say("cached mode failed: %r" % (e,))

    # orderly shutdown on every rank (the library's NCCL communicator, then torch's)
    b.close()
    if world > 1:
        dist.barrier()
        torch.cuda.cudart().cudaHostUnregister(old.ctypes.data)
        del old
        shm.close()
        if rank == 0:
            shm.unlink()

    # ---- the same arm on the meshes the CPU arm can run (same physics, same tolerance) -----------------------
    at_ref = None
    if world == 1 and not args.no_checks:
        at_ref = [gpu_at_reference_n(pb, torch, n, args.rel_tol, args.solver_options, local_rank, (args.krylov, args.gmres_m))
                  for n in REF_MESHES]
    if world > 1:
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
        traffic, traffic_src = ncu_traffic("k_spmv_stream", args.n_refine, world)
        roof = {"bound": "hbm", "kernel": "k_spmv_stream (three-field solver matrix, %d rows, %d entries)" % (rows_owned, nnz_owned),
                "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm, "traffic": traffic, "traffic_source": traffic_src,
                "bytes_per_launch": spmv_bytes, "ms_per_launch": spmv_ms, "peak_source": peak_src}
    iters = [p["iterations"] for p in per_step]
    step_bytes = (sum(p["spmv"] for p in per_step) / len(per_step)) * spmv_bytes + 1904.0 * n_cells / world
    cfg = workload_config(args.n_refine, args.rel_tol)
    line = {"metric": METRIC, "value": n_cells * args.steps / dev_s, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_s / args.steps * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
            "setup_s": t_setup,
            "scaling_note": "BASELINE.json configs: [1] 1024^2 on 1 B200, [2] 4096^2 partitioned across 2/4/8 B200; strong scaling "
                            "holds among N = 2, 4, 8 (same mesh); the 1-GPU line of the 4096^2 mesh is kept in profiles/",
            "mode": "cached (--reuse-operator)" if args.reuse_operator else "everything rebuilt in every step",
            "cached_mode": cached,
            "s_per_time_step": dev_s / args.steps, "assemble_ms": sum(asm_ms) / len(asm_ms),
            "assembly_cells_per_s": n_cells / asm_kernel_s, "solve_ms": sum(sol_ms) / len(sol_ms),
            "solver": {"method": ("FGMRES(%d), block upper-triangular preconditioner" % args.gmres_m if args.krylov == 1 else
                                  "MINRES, block-diagonal preconditioner") +
                                 " on the three-field (u, total pressure, p) form; blocks: bubble Jacobi x Q1 V-cycle; cell mass; "
                                 "p_int elimination + face V-cycle; preconditioner coefficients stored in FP32, Krylov vectors / "
                                 "SpMV / true residual in FP64; initial guess = order-%d extrapolation in time" % args.initial_guess +
                                 ("; CACHED MODE: solver operator + preconditioner rebuilt only when the matrix inputs (material, "
                                  "dt, K, boundary, mesh) change -- the system matrix and rhs are still assembled in every step"
                                  if args.reuse_operator else
                                  "; system matrix, rhs, solver operator and preconditioner all rebuilt in every step (the "
                                  "reference refactorises in every step)"),
                       "stopping": "true residual of the reference system ||b - A x|| <= rel_tol ||b||",
                       "iterations": iters, "per_step": per_step, "converged": all_converged,
                       "rel_residual_max": max(p["rel_residual"] for p in per_step)},
            "whole_step_hbm_frac": step_bytes / (dev_s / args.steps) / 1e9 / hbm,
            "wall_s": wall, "gpu_launches": int(launches), "clocks": clocks, "e2e": e2e,
            "roofline": roof if roof else roof_asm, "roofline_assembly": roof_asm,
            "fp64_peak_measured_tflops": fp64_tflops, "parity_check": parity, "solution_check": sol_check,
            "value_at_reference_n": at_ref}
    if not args.no_cpu:
        r = cpu_run(args.cpu_refine, 2, dt)
        ci = cpu_info()
        line["cpu_baseline"] = {"value": r["value"], "unit": UNIT, "cores": 1, "kind": "port", "sample": r["desc"],
                                "nproc": ci["nproc"], "cpu_model": ci["model"]}
    print(json.dumps(line))
    bad = []
    if failed or not all_converged:
        bad.append("a timed step did not converge: %s" % (failed,))
    if not e2e_conv:
        bad.append("an end-to-end step did not converge")
    if parity is not None and not parity["ok"]:
        bad.append("parity_check failed")
    if bad:
        print("bench.py: " + "; ".join(bad), file=sys.stderr)
        sys.exit(1)


if __name__ == "__main__":
    main()
