"""Host-side mirror of the reference's classes for the hot path, on top of the C ABI.

The reference programs are one class per file with the private members make_grid_and_dofs() /
setup_system(), assemble_system(t), solve_time_step() / solve() and run()
(PoroElasBR1WG_new.cc:62-137, PoroElasCGQ1_WG.cc, EltStiffMat.cc).  The same names are kept here
so that the parity tests read like the reference; the C++ twin lives in host/.  Nothing in this
module computes: every method forwards to libporoelas_b200.so.
"""
import ctypes as C
import math

import numpy as np

from . import _capi as capi
from ._capi import (PE_BR1_WG, PE_CGQ1_WG, PE_WG_Q2RT2, PE_CASE_ZERO, PE_CASE_COSCOS, PE_CASE_DARCY_COSCOS, PE_FACE_DIR_U,
                    PE_FACE_DIR_P, PE_FACE_NEU_U, PoroElasError, pe_config, pe_sizes, pe_solve_stats, ptr)


class BiotSystem:
    """One reference program instance bound to one GPU."""

    element = PE_BR1_WG

    def __init__(self, element=None, device=-1, rank=0, world_size=1, host_setup_only=False, dim=2):
        self.L = capi.load()
        if element is not None:
            self.element = element
        cfg = pe_config(self.element, dim, device, rank, world_size, 1 if host_setup_only else 0)
        h = C.c_void_p()
        rc = self.L.pe_create(C.byref(h), C.byref(cfg))
        if rc != 0:
            raise PoroElasError(rc, "pe_create failed (no CUDA device? this library has no CPU path)")
        self.h = h
        self.rank, self.world_size = rank, world_size
        # constructor init-list of the reference's "general case" (BR1new:532-539)
        self.lam, self.mu, self.alpha, self.capacity = 1.0, 1.0, 1.0, 0.0
        self.time_step = 1.0 / 16
        self.total_time = 1.0
        self.stats = pe_solve_stats()
        self.rel_tol, self.max_iterations = 1e-10, 200000

    def close(self):
        if getattr(self, "h", None):
            self.L.pe_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise PoroElasError(rc, self.L.pe_last_error(self.h).decode())

    # ---- make_grid_and_dofs()  BR1new:580-681 / CGQ1:883-1027
    def make_grid_and_dofs(self, n_refine, distort=0.0, seed=12345,
                           boundary_kinds=PE_FACE_DIR_U | PE_FACE_DIR_P, pattern=True):
        self._ck(self.L.pe_make_unit_square(self.h, n_refine, distort, seed, boundary_kinds))
        self._ck(self.L.pe_distribute_dofs(self.h))
        if pattern:
            self._ck(self.L.pe_make_sparsity_pattern(self.h))
        return self.sizes()

    def make_grid(self, n_refine, distort=0.0, seed=12345, boundary_kinds=0):
        self._ck(self.L.pe_make_unit_square(self.h, n_refine, distort, seed, boundary_kinds))

    def set_mesh(self, x, y, cell_vertices):
        x = np.ascontiguousarray(x, np.float64)
        y = np.ascontiguousarray(y, np.float64)
        cv = np.ascontiguousarray(cell_vertices, np.uint32)
        self._ck(self.L.pe_set_mesh(self.h, x.size, ptr(x), ptr(y), 1, cv.shape[0], ptr(cv)))

    def distribute_dofs(self):
        self._ck(self.L.pe_distribute_dofs(self.h))

    def set_dofs(self, n_u, n_pint, n_pface, cell_dofs):
        cd = np.ascontiguousarray(cell_dofs, np.uint32)
        self._ck(self.L.pe_set_dofs(self.h, n_u, n_pint, n_pface, ptr(cd)))

    def make_sparsity_pattern(self):
        self._ck(self.L.pe_make_sparsity_pattern(self.h))

    def set_pattern(self, rowptr, colidx):
        rp = np.ascontiguousarray(rowptr, np.int64)
        ci = np.ascontiguousarray(colidx, np.int32)
        self._ck(self.L.pe_set_pattern(self.h, ptr(rp), ptr(ci)))

    def sizes(self):
        s = pe_sizes()
        self._ck(self.L.pe_get_sizes(self.h, C.byref(s)))
        return s

    def owned_ranges(self):
        """the three contiguous global row ranges [begin, end) this rank owns (u | p_int | p_face)"""
        b = np.zeros(3, np.int64)
        e = np.zeros(3, np.int64)
        self._ck(self.L.pe_get_owned_ranges(self.h, ptr(b), ptr(e)))
        return b, e

    def owned_rows(self):
        """global indices of the rows this rank owns, in local order"""
        b = np.zeros(3, np.int64)
        e = np.zeros(3, np.int64)
        self._ck(self.L.pe_get_owned_ranges(self.h, ptr(b), ptr(e)))
        return np.concatenate([np.arange(b[k], e[k]) for k in range(3)])

    def mesh(self):
        s = self.sizes()
        x = np.zeros(s.n_vertices)
        y = np.zeros(s.n_vertices)
        cv = np.zeros((s.n_cells, 4), np.uint32)
        fk = np.zeros((s.n_cells, 4), np.uint8)
        self._ck(self.L.pe_get_mesh(self.h, ptr(x), ptr(y), ptr(cv), ptr(fk)))
        return x, y, cv, fk

    def cell_dofs(self):
        s = self.sizes()
        cd = np.zeros((s.n_cells, s.n_loc), np.uint32)
        self._ck(self.L.pe_get_dofs(self.h, ptr(cd)))
        return cd

    def pattern(self):
        s = self.sizes()
        rp = np.zeros(s.n_owned_rows + 1, np.int64)
        ci = np.zeros(s.nnz_owned, np.int32)
        self._ck(self.L.pe_get_pattern(self.h, ptr(rp), ptr(ci)))
        return rp, ci

    def dealii_block_order(self):
        s = self.sizes()
        p = np.zeros(s.nnz, np.int64)
        self._ck(self.L.pe_get_dealii_block_order(self.h, ptr(p)))
        return p

    # ---- problem data
    def set_material(self, lam=None, mu=None, alpha=None, capacity=None, k_scalar=1.0, k_cell=None):
        self.lam = self.lam if lam is None else lam
        self.mu = self.mu if mu is None else mu
        self.alpha = self.alpha if alpha is None else alpha
        self.capacity = self.capacity if capacity is None else capacity
        kc = None if k_cell is None else np.ascontiguousarray(k_cell, np.float64)
        self._ck(self.L.pe_set_material(self.h, self.lam, self.mu, self.alpha, self.capacity, k_scalar, ptr(kc)))

    def set_lognormal_k(self, sigma, seed):
        self._ck(self.L.pe_set_lognormal_k(self.h, sigma, seed))

    def set_case(self, case_id, params=None):
        p = None if params is None else np.ascontiguousarray(params, np.float64)
        self._ck(self.L.pe_set_case(self.h, case_id, ptr(p)))

    def set_boundary(self, cells, faces, kinds, traction=(0.0, 0.0)):
        c = np.ascontiguousarray(cells, np.uint32)
        f = np.ascontiguousarray(faces, np.uint8)
        k = np.ascontiguousarray(kinds, np.uint8)
        t = np.ascontiguousarray(traction, np.float64)
        self._ck(self.L.pe_set_boundary(self.h, c.size, ptr(c), ptr(f), ptr(k), ptr(t)))

    def set_face_kinds(self, face_kinds, traction=(0.0, 0.0)):
        fk = np.asarray(face_kinds, np.uint8).reshape(-1, 4)
        cells, faces = np.nonzero(fk)
        self.set_boundary(cells, faces, fk[cells, faces], traction)

    # ---- hot path
    def assemble_system(self, t, old_solution=None, time_step=None):
        dt = self.time_step if time_step is None else time_step
        o = None if old_solution is None else np.ascontiguousarray(old_solution, np.float64)
        self._ck(self.L.pe_assemble_system(self.h, t, dt, ptr(o)))

    def assemble_rhs_only(self, t, old_solution=None, time_step=None):
        dt = self.time_step if time_step is None else time_step
        o = None if old_solution is None else np.ascontiguousarray(old_solution, np.float64)
        self._ck(self.L.pe_assemble_rhs_only(self.h, t, dt, ptr(o)))

    def system_matrix(self):
        v = np.zeros(self.sizes().nnz_owned)
        self._ck(self.L.pe_get_matrix(self.h, ptr(v)))
        return v

    def system_rhs(self):
        r = np.zeros(self.sizes().n_owned_rows)
        self._ck(self.L.pe_get_rhs(self.h, ptr(r)))
        return r

    def solve_time_step(self, want_solution=True, check=True, out=None):
        """out: optional caller-owned buffer for the solution (e.g. pinned host memory)"""
        n = self.sizes().n_owned_rows
        sol = (out if out is not None else np.zeros(n)) if want_solution else None
        rc = self.L.pe_solve_time_step(self.h, self.rel_tol, self.max_iterations, ptr(sol), C.byref(self.stats))
        if rc != 0 and (check or rc != capi.PE_ERR_NOT_CONVERGED):
            self._ck(rc)
        return sol

    def set_solver_options(self, use_multigrid=1, theta=1.0, omega=0.6, nu=2):
        self._ck(self.L.pe_set_solver_options(self.h, use_multigrid, theta, omega, nu))

    def set_tuning(self, key, value):
        self._ck(self.L.pe_set_tuning(self.h, key.encode(), float(value)))

    def set_krylov(self, method=1, restart=0):
        """1 = FGMRES(restart) + block-triangular preconditioner (default), 0 = MINRES + block-diagonal"""
        self._ck(self.L.pe_set_krylov(self.h, method, restart))

    def set_recycling(self, pairs):
        """pairs (0..4) of previous corrections the next solves project onto (same matrix only); 0 = off"""
        self._ck(self.L.pe_set_recycling(self.h, pairs))

    def set_initial_guess(self, order):
        """0 = old_solution, 1 / 2 = linear / quadratic extrapolation in time (default 2)"""
        self._ck(self.L.pe_set_initial_guess(self.h, order))

    def solution(self):
        s = np.zeros(self.sizes().n_owned_rows)
        self._ck(self.L.pe_get_solution(self.h, ptr(s)))
        return s

    def solution_global(self, out):
        """owned part of the device solution into `out` (n_dofs doubles, global deal.II order)"""
        self._ck(self.L.pe_get_solution_global(self.h, ptr(out)))
        return out

    def set_solution(self, sol):
        s = np.ascontiguousarray(sol, np.float64)
        self._ck(self.L.pe_set_solution(self.h, ptr(s)))

    def spmv(self, x):
        x = np.ascontiguousarray(x, np.float64)
        y = np.zeros(self.sizes().n_owned_rows)
        self._ck(self.L.pe_spmv(self.h, ptr(x), ptr(y)))
        return y

    def bench_spmv(self, repeats=20):
        ms = C.c_double()
        self._ck(self.L.pe_bench_spmv(self.h, repeats, C.byref(ms)))
        return ms.value

    def bench_fp64(self):
        """measured DFMA throughput of the device in TFLOP/s (FP64 roofline denominator)"""
        t = C.c_double()
        self._ck(self.L.pe_bench_fp64(self.h, C.byref(t)))
        return t.value

    def local_matrices(self, first, n, t=0.0, old_solution=None, time_step=None, want_rhs=False):
        dt = self.time_step if time_step is None else time_step
        nl = self.sizes().n_loc
        out = np.zeros((n, nl, nl))
        rhs = np.zeros((n, nl)) if want_rhs else None
        o = None if old_solution is None else np.ascontiguousarray(old_solution, np.float64)
        self._ck(self.L.pe_local_matrices(self.h, first, n, t, dt, ptr(o), ptr(out), ptr(rhs)))
        return (out, rhs) if want_rhs else out

    def local_matrices_bench(self, chunk_cells, repeats=1, t=0.0, time_step=None):
        dt = self.time_step if time_step is None else time_step
        cs, sec = C.c_double(), C.c_double()
        self._ck(self.L.pe_local_matrices_bench(self.h, t, dt, chunk_cells, repeats, C.byref(cs), C.byref(sec)))
        return cs.value, sec.value

    def step_errors(self, t):
        e = np.zeros(3)
        self._ck(self.L.pe_step_errors(self.h, t, ptr(e)))
        return e

    def postprocess(self):
        """(div_displacement, displacement_magnitude, darcy beta [n][4], cell-centre Darcy velocity [n][2]) of the owned cells"""
        s = self.sizes()
        n = s.cell_end - s.cell_begin
        d, m, b, v = np.zeros(n), np.zeros(n), np.zeros((n, 4)), np.zeros((n, 2))
        self._ck(self.L.pe_postprocess(self.h, ptr(d), ptr(m), ptr(b), ptr(v)))
        return d, m, b, v

    def timings(self):
        a, k, s, m = C.c_double(), C.c_double(), C.c_double(), C.c_double()
        n = C.c_int64()
        self._ck(self.L.pe_get_timings(self.h, C.byref(a), C.byref(k), C.byref(s), C.byref(m), C.byref(n)))
        return {"assemble_ms": a.value, "assemble_kernel_ms": k.value, "solve_ms": s.value,
                "spmv_ms_avg": m.value, "launches": n.value}

    def comm_init(self, unique_id_bytes):
        b = np.frombuffer(bytes(unique_id_bytes), np.uint8).copy()
        self._ck(self.L.pe_comm_init(self.h, ptr(b)))

    # ---- run(): the time loop  BR1new:2156-2243 / CGQ1:2908-2992
    def run(self, n_refine, t_end=1.0, verbose=False, errors=False):
        self.make_grid_and_dofs(n_refine)
        self.set_material()
        step, time = 1, self.time_step
        hist = []
        while time <= t_end:  # float-accumulated loop bound, as in the reference (BR1new:2181)
            if verbose:
                print("Time step %d at t=%g time_step %g" % (step, time, self.time_step))
            self.assemble_system(time)  # old_solution = previous solution, on the device
            self.solve_time_step(want_solution=False)
            hist.append((time, self.stats.iterations, self.stats.rel_residual))
            time += self.time_step
            step += 1
        return hist


def comm_unique_id():
    L = capi.load()
    b = np.zeros(128, np.uint8)
    rc = L.pe_comm_unique_id(ptr(b))
    if rc != 0:
        raise PoroElasError(rc, "ncclGetUniqueId failed")
    return b.tobytes()


class PoroElasBR1WG(BiotSystem):
    """PoroElasBR1WG<2> of PoroElasBR1WG_new.cc."""
    element = PE_BR1_WG


class PoroElas_CGQ1_WG(BiotSystem):
    """PoroElas_CGQ1_WG<2> of PoroElasCGQ1_WG.cc (2-D general case)."""
    element = PE_CGQ1_WG

    def __init__(self, **kw):
        super().__init__(**kw)
        self.time_step = 1.0 / 1024  # CGQ1:825


class PoroElas_CGQ1_WG_3D(BiotSystem):
    """PoroElas_CGQ1_WG<3> (the shipped main(), CGQ1:3029): first step -- the cell loop on hexahedra."""
    element = PE_CGQ1_WG

    def __init__(self, **kw):
        super().__init__(dim=3, **kw)
        self.time_step = 1e-3  # CGQ1:863

    def make_grid(self, n_refine=3, distort=0.0, seed=12345):
        self._ck(self.L.pe_make_unit_cube(self.h, n_refine, distort, seed))

    def set_mesh3(self, x, y, z, cell_vertices):
        x, y, z = (np.ascontiguousarray(a, np.float64) for a in (x, y, z))
        cv = np.ascontiguousarray(cell_vertices, np.uint32)
        self._ck(self.L.pe_set_mesh3(self.h, x.size, ptr(x), ptr(y), ptr(z), cv.shape[0], ptr(cv)))

    def mesh3(self):
        nv, nc = C.c_int64(), C.c_int64()
        self._ck(self.L.pe_get_mesh3(self.h, C.byref(nv), C.byref(nc), None, None, None, None))
        x, y, z = np.zeros(nv.value), np.zeros(nv.value), np.zeros(nv.value)
        cv = np.zeros((nc.value, 8), np.uint32)
        self._ck(self.L.pe_get_mesh3(self.h, None, None, ptr(x), ptr(y), ptr(z), ptr(cv)))
        return x, y, z, cv

    def setup_system(self):
        """make_grid_and_dofs part 2 (CGQ1:902-943): distribute_dofs + component_wise, make_sparsity_pattern"""
        self._ck(self.L.pe_distribute_dofs(self.h))
        self._ck(self.L.pe_make_sparsity_pattern(self.h))
        return self.sizes()

    def run(self, n_refine=3, t_end=0.011, verbose=False):
        """run() of the shipped program (CGQ1:2962-3011): hyper_cube + refine_global(3), sandwich permeability and
        boundary data, `for (time = dt; time <= 0.011; time += dt)` (CGQ1:2929), old_solution = solution on the device.
        Returns [(time, iterations, rel_residual)]."""
        self.make_grid(n_refine)
        x, y, z, cv = self.mesh3()
        zc = z[cv].mean(axis=1)
        self.set_material(k_cell=np.where((zc >= 0.25) & (zc <= 0.75), 1e-8, 1.0))   # Coefficient, CGQ1:176-189
        self.setup_system()
        hist, time, first = [], self.time_step, True
        while time <= t_end:
            if verbose:
                print("Time step %d at t=%g" % (len(hist) + 1, time))
            self.assemble_system(time, np.zeros(self.sizes().n_dofs) if first else None)
            self.solve_time_step(want_solution=False)
            hist.append((time, self.stats.iterations, self.stats.rel_residual))
            time += self.time_step
            first = False
        return hist

    def set_boundary3(self, face_kinds, traction=None):
        """face_kinds [n_cells][6] of PE_FACE_* bits; make_grid installs the shipped sandwich data"""
        fk = np.ascontiguousarray(face_kinds, np.uint8)
        tr = None if traction is None else np.ascontiguousarray(traction, np.float64)
        self._ck(self.L.pe_set_boundary3(self.h, ptr(fk), ptr(tr)))

    def local_matrices3(self, first, n, cell_old=None, want_rhs=True):
        out = np.zeros((n, 31, 31))
        rhs = np.zeros((n, 31)) if want_rhs else None
        o = None if cell_old is None else np.ascontiguousarray(cell_old, np.float64)
        self._ck(self.L.pe_local_matrices3(self.h, first, n, self.time_step, ptr(o), ptr(out), ptr(rhs)))
        return (out, rhs) if want_rhs else out


class WGDarcyEquation(BiotSystem):
    """WGDarcyEquation<2> of EltStiffMat.cc: WG(Q2,Q2;RT[2]) Darcy problem, stationary.
    make_grid (ESM:242-289), setup_system (294-326), assemble_system (333-578), solve (583-590), run (985-996)."""
    element = PE_WG_Q2RT2

    def __init__(self, **kw):
        super().__init__(**kw)
        self.rel_tol, self.max_iterations = 1e-8, 1000  # SolverControl(1000, 1e-8 * ||rhs||), ESM:585

    def make_grid(self, n_refine=1, distort=0.0, seed=12345):
        # hyper_cube(0,1) + refine_global(1); every boundary face carries the pressure Dirichlet value
        self._ck(self.L.pe_make_unit_square(self.h, n_refine, distort, seed, PE_FACE_DIR_P))

    def setup_system(self):
        self._ck(self.L.pe_distribute_dofs(self.h))
        self._ck(self.L.pe_make_sparsity_pattern(self.h))
        return self.sizes()

    def assemble_system(self, t=0.0, old_solution=None, time_step=1.0):
        self._ck(self.L.pe_assemble_system(self.h, 0.0, 1.0, None))

    def solve(self, want_solution=True, check=True):
        return self.solve_time_step(want_solution=want_solution, check=check)

    def postprocess(self):
        """(L2_error_vel, L2_error_flux) as EltStiffMat.cc prints them (ESM:915-917)"""
        e = np.zeros(2)
        self._ck(self.L.pe_darcy_errors(self.h, ptr(e)))
        return np.sqrt(e)

    def run(self, n_refine=1):
        self.make_grid(n_refine)
        self.setup_system()
        self.set_case(PE_CASE_DARCY_COSCOS)
        self.assemble_system()
        return self.solve()
