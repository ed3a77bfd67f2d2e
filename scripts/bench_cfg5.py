"""BASELINE cfg 5: batched local-matrix assembly only (no scatter, no solve) on a synthetic mesh."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import poroelasticity_b200 as pb
nref = int(sys.argv[1]) if len(sys.argv) > 1 else 12
el = int(sys.argv[2]) if len(sys.argv) > 2 else 0
distort = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
chunk = int(sys.argv[4]) if len(sys.argv) > 4 else (1 << 20)
dfma = int(sys.argv[5]) if len(sys.argv) > 5 else 0   # element 2: 1 = the first (DFMA-only) cell kernel
b = pb.BiotSystem(element=el)
b.set_tuning("esm_dfma", dfma)
t0 = time.time(); b.make_grid(nref, distort=distort); ts = time.time() - t0
b.time_step = 0.1; b.set_material(1, 1, 1, 0); b.set_lognormal_k(2.0, 12345)
b.set_case(pb.PE_CASE_DARCY_COSCOS if el == 2 else pb.PE_CASE_COSCOS)
cs, sec = b.local_matrices_bench(chunk, repeats=3, t=0.1)
n = 1 << (2 * nref)
nl = {0: 17, 1: 13, 2: 21}[el]
bytes_per_cell = 64 + 8 + 8 * nl * nl
fp64 = b.bench_fp64()
flop_per_cell = {0: 9.0e3, 1: 5.0e3, 2: 47.0e3}[el]   # lean FP64 work per cell
print(json.dumps({"fp64_peak_measured_tflops": fp64, "tflops": n * flop_per_cell / sec * 1e-12,
                  "frac_of_fp64_peak": n * flop_per_cell / sec * 1e-12 / fp64, "cfg": 5, "element": el, "esm_dfma": dfma, "cells": n, "distort": distort, "chunk_cells": chunk, "seconds": sec, "cells_per_s": n / sec,
                  "GBps_algorithmic": n * bytes_per_cell / sec / 1e9, "frac_of_6547": n * bytes_per_cell / sec / 1e9 / 6547.2, "checksum": cs, "mesh_s": ts}))
