/* poroelas_b200.h -- C ABI of the B200 (sm_100a) replacement for the hot path of
 * sophy1029/PoroElasticity: per-time-step assembly of the coupled Biot system and the sparse solve.
 *
 * The reference has no plugin / FFI interface; its seam is the private member functions of one
 * class per program.  Each entry point below names the reference code it replaces
 * ("BR1new" = PoroElasBR1WG_new.cc, "CGQ1" = PoroElasCGQ1_WG.cc, "ESM" = EltStiffMat.cc,
 * "BR1old" = PoroElasBR1WG.cc).  INTEGRATION.md shows the deal.II-side binding.
 *
 * Conventions: every function returns 0 on success or a negative pe_status; it never throws.
 * pe_last_error(h) returns the text of the last failure on that handle.  All pointers are HOST
 * pointers unless the parameter name ends in _dev; the caller owns every buffer it passes and may
 * free it when the call returns; the library owns all device memory.  One handle is used from one
 * host thread and is bound to one CUDA device.  There is no CPU fallback: without a CUDA device
 * pe_create fails with PE_ERR_NO_DEVICE.
 */
#ifndef POROELAS_B200_H
#define POROELAS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pe_handle pe_handle;

typedef enum pe_status {
  PE_OK             = 0,
  PE_ERR_ARG        = -1, /* bad argument / call order */
  PE_ERR_NO_DEVICE  = -2, /* no CUDA device: this library has no CPU path */
  PE_ERR_CUDA       = -3, /* a CUDA runtime call or kernel failed */
  PE_ERR_PATTERN    = -4, /* an entry of a local matrix is not in the sparsity pattern */
  PE_ERR_NOT_CONVERGED = -5,
  PE_ERR_NCCL       = -6,
  PE_ERR_ALLOC      = -7
} pe_status;

/* FESystem of the program.  BR1new:517-519, CGQ1:808-810, ESM:229-233 */
typedef enum pe_element {
  PE_BR1_WG   = 0, /* (FE_BernardiRaugel(1), FE_DGQ(0), FE_FaceQ(0))  17 dofs/cell in 2-D */
  PE_CGQ1_WG  = 1, /* (FE_Q(1)^dim,          FE_DGQ(0), FE_FaceQ(0))  13 dofs/cell in 2-D */
  PE_WG_Q2RT2 = 2  /* (FE_DGQ(2), FE_FaceQ(2)) + RT[2]: Darcy only, 21 dofs/cell (EltStiffMat) */
} pe_element;

/* Problem data selected in the reference by commenting blocks in/out.
 * PE_CASE_ZERO   : shipped state of BR1new (u = p = f = s = 0)            BR1new:220-223,384,424-425
 * PE_CASE_COSCOS : the "coscos" manufactured solution                      BR1new:207-210,376-378,409-412
 * PE_CASE_DARCY_COSCOS : p = cos(pi x) cos(pi y), f = 2 pi^2 p, K = I     ESM:147,167,187        */
typedef enum pe_case { PE_CASE_ZERO = 0, PE_CASE_COSCOS = 1, PE_CASE_DARCY_COSCOS = 2 } pe_case;

/* What the reference does on a boundary face (it selects these by boundary_id and by blocks). */
enum {
  PE_FACE_DIR_U = 1, /* displacement Dirichlet   BR1new:1109-1159 (boundary_id()==0), CGQ1:1673-1679 */
  PE_FACE_DIR_P = 2, /* face-pressure Dirichlet  BR1new:1063-1070, CGQ1:1663-1670                   */
  PE_FACE_NEU_U = 4  /* traction t_N on the face BR1new:963-990,   CGQ1:1563-1584                   */
};

typedef struct pe_config {
  int32_t element;    /* pe_element */
  int32_t dim;        /* 2; 3 = CGQ1-WG on hexahedra: cell loop, numbering and sparsity (pe_make_unit_cube, pe_set_mesh3, pe_local_matrices3) */
  int32_t device;     /* CUDA device ordinal; -1 = current device */
  int32_t rank;       /* this process's rank in the mesh partition (0 if world_size == 1) */
  int32_t world_size; /* number of GPUs / processes the mesh is partitioned over */
  int32_t flags;      /* PE_CFG_* */
  int32_t reserved[2];
} pe_config;

/* pe_config.flags: setup layer only (mesh, DoF numbering, sparsity: host code by design, SURVEY 8(a)
 * rows S1-S3).  No CUDA context is created; every hot-path call fails with PE_ERR_NO_DEVICE. */
enum { PE_CFG_HOST_SETUP_ONLY = 1 };

typedef struct pe_sizes {
  int64_t n_cells, n_vertices, n_loc, n_u, n_pint, n_pface, n_dofs, nnz;
  /* partition view of this rank (== the global numbers when world_size == 1) */
  int64_t cell_begin, cell_end;         /* owned cells (contiguous in deal.II active-cell order) */
  int64_t n_owned_rows, nnz_owned;      /* rows owned = dofs first touched by an owned cell */
  int64_t n_ghost_cells;                /* cells assembled redundantly for owned rows */
  /* the Krylov solver's own operator (three-field form, pruned CSR; 0 before the first solve) */
  int64_t solver_rows, solver_nnz;
  int64_t reserved[1];
} pe_sizes;

typedef struct pe_solve_stats {
  int32_t iterations;
  int32_t converged;        /* 0 = no; 1 = ||b - A x|| <= rel_tol ||b||; 2 = rounding floor of the data reached:
                             * ||b - A x|| <= 64 eps (||A||_inf ||x|| + ||b||), backward stable like the LU solve */
  double  rhs_norm;         /* ||b||_2 (preconditioned norm for MINRES is reported separately) */
  double  residual_norm;    /* true ||b - A x||_2 at exit */
  double  rel_residual;     /* residual_norm / rhs_norm */
  double  seconds;          /* device time of the solve (CUDA events) */
  int32_t spmv_count;
  int32_t restarts;         /* MINRES restarts from the current iterate (recurrence lost the true residual) */
} pe_solve_stats;

/* ---- lifetime ------------------------------------------------------------------------------ */
/* class constructor: BR1new:511-577, CGQ1:802-880, ESM:226-236 */
int pe_create(pe_handle **h, const pe_config *cfg);
int pe_destroy(pe_handle *h);
const char *pe_last_error(const pe_handle *h);
const char *pe_version(void);

/* ---- S1: mesh.  GridGenerator::hyper_cube(0,1) + refine_global(n)  BR1new:583-584 CGQ1:886-887
 * ESM:245-246.  Active cells in deal.II (Morton) order, vertices of a cell in deal.II order.
 * distort > 0 displaces interior vertices by <= distort*h (counter-based generator on seed,
 * BASELINE cfg 5 / CGQ1:897 distort_random).  All boundary faces get kinds `boundary_kinds`. */
int pe_make_unit_square(pe_handle *h, int n_refine, double distort, uint64_t seed, int boundary_kinds);
/* Any quadrilateral mesh handed over by a deal.II host: triangulation.get_vertices() as SoA or
 * AoS (xy_stride = 1 or 2) and cell->vertex_index(v), 4 per cell, active-cell order. */
int pe_set_mesh(pe_handle *h, int64_t n_vertices, const double *x, const double *y, int64_t xy_stride,
                int64_t n_cells, const uint32_t *cell_vertices);

/* ---- three dimensions (first step): the shipped main() of PoroElasCGQ1_WG.cc is the 3-D program
 * (CGQ1:3029; hyper_cube + refine_global(3), CGQ1:886-887).  A dim = 3 handle holds a hexahedral mesh (active cells in
 * deal.II's 3-D Morton order, cell vertex v at offset (v & 1, (v >> 1) & 1, v >> 2)) and runs the CELL LOOP of
 * assemble_system (CGQ1:1165-1453): dense 31 x 31 local matrices out[n][31][31] and right-hand sides rhs_out[n][31]
 * for f = s = 0 (the sandwich problem), local dof order 3 v + c | 24 + face | 30 = p_int; cell_old = the 31 old-solution
 * values of every cell (may be NULL).  pe_distribute_dofs, pe_make_sparsity_pattern, pe_get_sizes, pe_get_dofs and
 * pe_get_pattern work on a dim = 3 handle (host side: first-touch numbering of vertices x 3 | FE_FaceQ(0) quads | FE_DGQ(0)
 * hex per cell + component_wise, full-coupling CSR; one rank), and so do pe_assemble_system and pe_solve_time_step
 * (below; first version of the 3-D solve).  Postprocessing, error norms and several GPUs in 3-D are not built. */
int pe_make_unit_cube(pe_handle *h, int n_refine, double distort, uint64_t seed);
int pe_set_mesh3(pe_handle *h, int64_t n_vertices, const double *x, const double *y, const double *z, int64_t n_cells,
                 const uint32_t *cell_vertices);
int pe_get_mesh3(const pe_handle *h, int64_t *n_vertices, int64_t *n_cells, double *x, double *y, double *z,
                 uint32_t *cell_vertices);
int pe_local_matrices3(pe_handle *h, int64_t first_cell, int64_t n, double time_step, const double *cell_old, double *out,
                       double *rhs_out);
/* Boundary data of assemble_system in 3-D: face_kinds[n_cells][6] with PE_FACE_* bits (0 on interior faces), one constant
 * traction vector for the PE_FACE_NEU_U faces (NULL = 0).  Dirichlet values are zero (BoundaryValues, CGQ1:697-725).
 * pe_make_unit_cube installs the data of the shipped program: faces with centre z = 1 = boundary id 1 (CGQ1:972-981) carry
 * PE_FACE_DIR_P | PE_FACE_NEU_U with traction (0, 0, -1) (CGQ1:1563-1582, 1663-1670), all other boundary faces
 * PE_FACE_DIR_U (CGQ1:1673-1679).  pe_set_mesh3 clears the boundary data.
 * pe_assemble_system on a dim = 3 handle (after pe_distribute_dofs + pe_make_sparsity_pattern; one GPU): cell loop,
 * traction term, constraints of the whole mesh, distribute_local_to_global (CGQ1:1165-1706); t is ignored (no
 * time-dependent data), old_solution = host vector of all dofs or NULL = the state resident on the device (zero after
 * setup, the solution of the previous pe_solve_time_step afterwards); pe_get_matrix / pe_get_rhs return the result in
 * the CSR order of pe_get_pattern.
 * pe_solve_time_step on a dim = 3 handle (CGQ1:1712-1722): first version -- MINRES on the row-signed two-field system
 * with a diagonal / Schur-diagonal preconditioner, warm start from old_solution, true residual <= rel_tol ||b||; no
 * multigrid in 3-D yet (adequate for the shipped 8^3 mesh).  Constrained dofs come out as their boundary value 0. */
int pe_set_boundary3(pe_handle *h, const uint8_t *face_kinds, const double *traction);

/* ---- S2: DoFs.  distribute_dofs + DoFRenumbering::component_wise  BR1new:586-597 CGQ1:902-919 */
int pe_distribute_dofs(pe_handle *h);
/* ... or the host's own numbering: cell->get_dof_indices(), n_loc per cell, FESystem order */
int pe_set_dofs(pe_handle *h, int64_t n_u, int64_t n_pint, int64_t n_pface, const uint32_t *cell_dofs);

/* ---- S3: sparsity.  make_sparsity_pattern(dof_handler, dsp, constraints, false) + copy_from
 * BR1new:612-621 CGQ1:934-943 ESM:320-324.  Canonical form: one global CSR, columns ascending. */
int pe_make_sparsity_pattern(pe_handle *h);
int pe_set_pattern(pe_handle *h, const int64_t *rowptr, const int32_t *colidx);

int pe_get_sizes(const pe_handle *h, pe_sizes *s);
/* the three contiguous GLOBAL row ranges [u | p_int | p_face] this rank owns (local row l of pe_get_rhs /
 * pe_get_solution runs through them in this order) */
int pe_get_owned_ranges(const pe_handle *h, int64_t begin[3], int64_t end[3]);
int pe_get_mesh(const pe_handle *h, double *x, double *y, uint32_t *cell_vertices, uint8_t *face_kinds);
int pe_get_dofs(const pe_handle *h, uint32_t *cell_dofs);
int pe_get_pattern(const pe_handle *h, int64_t *rowptr, int32_t *colidx);
/* position of every CSR entry in deal.II's own BlockSparsityPattern storage order (block-major,
 * diagonal entry first in each row of the square diagonal blocks): perm[k_dealii] = k_csr */
int pe_get_dealii_block_order(const pe_handle *h, int64_t *perm);

/* ---- problem data ---------------------------------------------------------------------------- */
/* constructor init-list lambda, mu, alpha, capacity  BR1new:555-558 CGQ1:863-866;
 * Coefficient::value_list BR1new:153-181: K = k_scalar I, or per-cell scalars k_cell[n_cells] */
int pe_set_material(pe_handle *h, double lambda, double mu, double alpha, double c0, double k_scalar,
                    const double *k_cell);
/* per-cell K = exp(sigma Z), Z ~ N(0,1) from the counter generator keyed on (seed, cell id)  (cfg 4) */
int pe_set_lognormal_k(pe_handle *h, double sigma, uint64_t seed);
/* ExactSolution / BoundaryValues / BodyRightHandSide / FluidRightHandSide  BR1new:183-471.
 * params[5] = {f_lambda, f_mu, f_alpha, s_capacity, s_alpha}: the constants those classes hard-code
 * (BR1new:369-370,400-402); NULL = {1,1,1,0,1}. */
int pe_set_case(pe_handle *h, int case_id, const double *params);
/* set_boundary_id loops BR1new:652-666 CGQ1:990-1000 + the Dirichlet/Neumann blocks.
 * The resident old_solution survives this call and pe_set_lognormal_k / pe_set_material (only the device tables are
 * rebuilt).  Multi-GPU: the rebuild re-runs the halo setup, which is COLLECTIVE -- every rank must make the same
 * sequence of pe_set_boundary / pe_set_lognormal_k calls before its next pe_assemble_system. */
int pe_set_boundary(pe_handle *h, int64_t n_bfaces, const uint32_t *cell, const uint8_t *face,
                    const uint8_t *kinds, const double traction[2]);

/* ---- hot path -------------------------------------------------------------------------------- */
/* assemble_system(t)  BR1new:684-1174  CGQ1:1030-1729  ESM:333-578.
 * old_solution: n_dofs doubles (deal.II block order [u | p_int | p_face]); NULL = the solution of
 * the previous pe_solve_time_step, already on the device (old_solution = solution, BR1new:2241),
 * or zero before the first solve (BR1new:2163). */
int pe_assemble_system(pe_handle *h, double t, double time_step, const double *old_solution);
/* the same, matrix kept from the previous call, right-hand side only (SURVEY 8(f).1 "cached") */
int pe_assemble_rhs_only(pe_handle *h, double t, double time_step, const double *old_solution);
/* system_matrix values (global CSR order of pe_get_pattern) and system_rhs */
int pe_get_matrix(const pe_handle *h, double *values);
int pe_get_rhs(const pe_handle *h, double *rhs);

/* solve_time_step()  BR1new:1177-1187 CGQ1:1732-1742 BR1old:859-870: UMFPACK LU in the reference,
 * here preconditioned MINRES on the pressure-row-negated (symmetric) form to ||r|| <= rel_tol ||b||.
 * solve()  ESM:583-590: SolverCG + PreconditionIdentity, tol = rel_tol*||b|| (1e-8), <= 1000 its.
 * solution may be NULL (result stays on the device as the next old_solution).
 * Stopping rule of MINRES: the TRUE residual of the reference system decides, ||b - A x||_2 <= rel_tol ||b||_2
 * (stats.converged = 1).  If the iterate has reached the rounding floor of the data first (e.g. a right-hand side
 * that vanishes against A x_old), it is accepted when it is backward stable like an LU solution,
 * ||b - A x||_2 <= 64 eps (||A||_inf ||x||_2 + ||b||_2)  (stats.converged = 2).  Otherwise PE_ERR_NOT_CONVERGED.
 * Multi-GPU: collective; a failed NCCL call inside the solve is reported as PE_ERR_NCCL. */
int pe_solve_time_step(pe_handle *h, double rel_tol, int max_iterations, double *solution,
                       pe_solve_stats *stats);
/* preconditioner of the Krylov solve.  use_multigrid bit 0: 1 = block preconditioner of the three-field form with
 * the vertex-grid and cell-grid V-cycles (meshes made by pe_make_unit_square), 0 = inverse block diagonal only.
 * omega = weight of the damped block-Jacobi sweeps on the vertex grids (every second sweep uses 1.5 omega), nu =
 * sweeps before and after each coarse correction; theta is reserved.  Defaults 1, 1.0, 0.6, 2.  Higher bits of
 * use_multigrid: 16 = warp-per-row SpMV, 32 = colour-ordered deterministic scatter (STRUCTURAL: recorded here, applied
 * by the next pe_make_sparsity_pattern / pe_set_pattern; an existing pattern keeps the mode its tables were built
 * for), 64 = no CUDA graphs, 128 = CUDA graphs also with NCCL operations (multi-GPU), bits 8.. = assembly chunk
 * size in units of 4096 cells (0 = keep; the chunk tables are rebuilt before the next assembly). */
int pe_set_solver_options(pe_handle *h, int use_multigrid, double theta, double omega, int nu);
/* Tuning / A-B switches (the hot path reads no environment variables; PE_TIMING=1 only switches the phase tracer on).
 * Keys: "mg_omega2", "cmg_omega", "cmg_omega2", "face_omega" (smoother weights); "mg_fused_n" (largest level of the
 * single-CTA multigrid kernel); "asm_gather" (1: gather assembly instead of the RED scatter); multi-GPU: "mg_replicated",
 * "mg_whole_slabs", "one_lane" (1: the round-1 communication schemes, kept for A/B runs), "mg_dist_min_n" (smallest grid
 * level distributed over the ranks); "reuse_operator" (CACHED MODE, default 0 = the solver operator and its
 * preconditioner are rebuilt in every pe_solve_time_step, as the reference refactorises in every step; 1 = they are
 * rebuilt only when something that enters the matrix values changes -- material data, time step, permeability,
 * boundary kinds, mesh / numbering.  The matrix itself is assembled by every pe_assemble_system and the true residual is
 * taken on it either way); "esm_dfma" (1: the first, DFMA-only
 * WG(Q2,Q2;RT[2]) cell kernel instead of the FP64 tensor-core one).  Structural keys rebuild the device tables before
 * the next assembly. */
int pe_set_tuning(pe_handle *h, const char *key, double value);
/* Krylov method of pe_solve_time_step for the Biot elements on meshes with a refinement hierarchy:
 * method 1 (default) = restarted flexible GMRES(restart) on the row-signed three-field system with the block
 * upper-triangular preconditioner [P_u -B^T 0; 0 P_xi 0; 0 0 P_p] (2.6x fewer iterations than method 0 at
 * lambda ~ mu, scripts/krylov_study.py); method 0 = MINRES with the block-diagonal SPD preconditioner.
 * restart: basis vectors per cycle, 2..24 (0 = keep, default 16).  Same stopping rule for both. */
int pe_set_krylov(pe_handle *h, int method, int restart);
/* Initial guess of the Krylov solve (a direct solver has none; the result is the same up to rel_tol):
 * order 0 = old_solution (BR1new:2241), 1 / 2 = linear / quadratic extrapolation in time from the solutions of the
 * previous steps kept on the device (default 2; falls back to a lower order until enough steps exist, when the
 * time does not advance, or after pe_set_solution). */
int pe_set_initial_guess(pe_handle *h, int order);
/* Projection onto the corrections of the previous solves with the same matrix (Fischer, successive right-hand sides):
 * `pairs` (0 .. 8, default 4; 0 = off) pairs (d_i, K d_i) are kept on the device (when full, the pair that contributed
 * least to the latest projection is replaced); the space is dropped whenever a quantity that enters the matrix changes
 * (mesh, material, time step, permeability, boundary kinds) and when a new trajectory starts (pe_set_solution, or the
 * time does not advance).  GMRES path. */
int pe_set_recycling(pe_handle *h, int pairs);
int pe_get_solution(const pe_handle *h, double *solution);
int pe_set_solution(pe_handle *h, const double *solution);
/* the owned part of the device solution written into a vector of n_dofs doubles in GLOBAL deal.II order (the three
 * owned row ranges of pe_get_owned_ranges): what a host program with one old_solution vector shared by all ranks of
 * the node wants (one process per GPU; every rank fills its own ranges). */
int pe_get_solution_global(const pe_handle *h, double *solution_global);

/* y = system_matrix * x (the SpMV the solver is built on), host in / host out: parity hook */
int pe_spmv(pe_handle *h, const double *x, double *y);

/* The cell loop alone, BR1new:791-961 / CGQ1:1165-1453 / ESM:397-513: dense local matrices (and
 * local right-hand sides) of cells [first, first+n), before constraints, out[n][n_loc][n_loc]
 * row-major, rhs_out[n][n_loc] (may be NULL). */
int pe_local_matrices(pe_handle *h, int64_t first_cell, int64_t n, double t, double time_step,
                      const double *old_solution, double *out, double *rhs_out);
/* BASELINE cfg 5: the same kernel over all cells in chunks, matrices written to HBM only
 * (entry-major [n_loc*n_loc][chunk]); checksum = sum of all entries; seconds = device time. */
int pe_local_matrices_bench(pe_handle *h, double t, double time_step, int64_t chunk_cells, int repeats,
                            double *checksum, double *seconds);

/* postprocess_errors BR1new:1251-1349 / CGQ1:2229-2357: squared errors of the device solution at
 * time t: e[0] = ||p_int - p||^2, e[1] = ||u_h - u||^2, e[2] = |u_h - u|_H1^2  (QGauss(3)). */
int pe_step_errors(pe_handle *h, double t, double e[3]);
/* run() accumulates dt*e over the steps and prints the square roots: L2L2Error_intp, L2L2Error_u,
 * L2h1semiError_u, L2h1Error_u = sqrt(sum dt (e[1]+e[2]))  (BR1new:2251-2266). */

/* postprocess()  CGQ1:1744-1966 (dilation 1849-1860, |u| 1862-1877, Darcy velocity in DGRT 1881-1964) and
 * BR1new:1189-1246 + postprocess_darcy_velocity BR1new:1352-1606 / CGQ1:1973-2225, from the device solution, for the
 * cells this rank owns (deal.II active-cell order, n = cell_end - cell_begin):
 *   div_displacement[n]        cell average of div u_h              (QGauss(3))
 *   displacement_magnitude[n]  sqrt of the cell average of |u_h|^2
 *   darcy_beta[n][4]           RT0 coefficients of the Darcy velocity -K grad_w p_h  (= darcy_velocity of CGQ1)
 *   velocity_center[n][2]      that velocity at the cell midpoint   (= velocity.txt of BR1new:1598-1602)
 * Any pointer may be NULL. */
int pe_postprocess(pe_handle *h, double *div_displacement, double *displacement_magnitude, double *darcy_beta,
                   double *velocity_center);

/* postprocess() of EltStiffMat.cc (ESM:662-919) for the WG(Q2,Q2;RT[2]) Darcy element, from the device solution:
 * e[0] = L2_error_vel^2  = sum_cells int |u_h - u|^2,  u_h = -K grad_w p_h in RT[2], u = Velocity::value (ESM:200-211)
 * e[1] = L2_error_flux^2 = sum_cells sum_faces (|E| / |f|) int_f ((u_h - u).n)^2     (the program prints the square roots) */
int pe_darcy_errors(pe_handle *h, double e[2]);

/* the solver's SpMV kernel alone, `repeats` back-to-back launches on the handle's stream timed with
 * CUDA events (roofline measurement; x = current solution vector): average ms per launch */
int pe_bench_spmv(pe_handle *h, int repeats, double *ms_per_launch);
/* FP64 roofline denominator: a dependent-chain-free DFMA loop on every SM of the handle's device, timed with CUDA
 * events (the driver's MEASURED_PEAKS.json holds no FP64 figure): achieved TFLOP/s */
int pe_bench_fp64(pe_handle *h, double *tflops);

/* ---- multi-GPU (one process per GPU) --------------------------------------------------------- */
/* 128-byte NCCL unique id made on rank 0 and broadcast by the host program (torch.distributed,
 * MPI, a file ...) */
int pe_comm_unique_id(uint8_t id[128]);
int pe_comm_init(pe_handle *h, const uint8_t id[128]);

/* ---- timing ---------------------------------------------------------------------------------- */
/* device milliseconds (CUDA events on the handle's stream) of the last pe_assemble_system and
 * of its dominant kernel; launches = kernels launched by this handle since pe_create */
int pe_get_timings(const pe_handle *h, double *assemble_ms, double *assemble_kernel_ms,
                   double *solve_ms, double *spmv_ms_avg, int64_t *launches);

#ifdef __cplusplus
}
#endif
#endif /* POROELAS_B200_H */
