// Solver interface between pe_api.cu and pe_solver.cu.
#pragma once
#include "../../include/poroelas_b200.h"

#include <cuda_runtime.h>

#include <cstdint>
#include <vector>

namespace pe
{
  struct MinresState;
  // logical index i of an owned unknown -> physical vector index i + (i >= at ? len : 0)   (pe_k3.h)
  struct Gap
  {
    int64_t at = INT64_MAX, len = 0;
  };
  // the two executable CUDA graphs of one MINRES iteration (one per parity of the buffer rotation), kept on the
  // handle across solves; key = fingerprint of everything baked into them (buffers, by-value kernel arguments)
  struct MinresGraphs
  {
    cudaGraphExec_t exec[2]     = {nullptr, nullptr};
    int64_t         launches[2] = {0, 0};
    uint64_t        key         = 0;
    MinresGraphs()              = default;
    MinresGraphs(const MinresGraphs &) = delete;
    MinresGraphs &operator=(const MinresGraphs &) = delete;
    ~MinresGraphs();
    void clear();
  };
  constexpr int kGmresMax = 24; // longest GMRES cycle (basis vectors kept)
  struct GmresGraphs
  {
    cudaGraphExec_t exec[kGmresMax]     = {};
    int64_t         launches[kGmresMax] = {};
    uint64_t        key                 = 0;
    int             m                   = 0;
    GmresGraphs()                       = default;
    GmresGraphs(const GmresGraphs &)    = delete;
    GmresGraphs &operator=(const GmresGraphs &) = delete;
    ~GmresGraphs();
    void clear();
  };
  // Projection onto the corrections of the previous solves with the SAME matrix (Fischer 1998: successive right-hand
  // sides): pairs (d_i, q_i = K d_i) with orthonormal q_i.  Before the first GMRES cycle the initial residual is
  // projected, x += sum <r0, q_i> d_i; after the solve the new correction is appended (oldest pair dropped).
  struct Recycle
  {
    int      k_max = 0, k = 0, next = 0; // capacity, pairs held, slot of the next pair
    double  *D = nullptr, *Q = nullptr;  // k_max vectors each, `stride` apart
    double  *x0 = nullptr, *r0 = nullptr; // iterate / residual after the projection (for the new pair)
    int64_t  stride = 0;
    uint64_t matrix_key = 0;             // fingerprint of the operator the pairs belong to
    double   alpha[8]   = {};            // |<r_0, q_i>| of the latest projection: the least used pair is replaced
    bool     have_alpha = false;
  };
  struct SolverCtx
  {
    int64_t        L = 0, H = 0;    // owned rows, halo columns
    int64_t        n_pos = 0;       // rows [0,n_pos) keep their sign, rows >= n_pos are negated (p rows)
    int64_t        pint_begin = 0, pint_end = 0; // local row range of the interior pressures
    const int64_t *rowptr = nullptr;
    const int32_t *colidx = nullptr;
    const double  *values = nullptr;
    const int64_t *diag_pos = nullptr; // position of the diagonal entry of every owned row (null: search the row)
    const int32_t *rowblk = nullptr; // row runs of the stream SpMV (n_rowblk + 1 entries), null = warp-per-row
    int64_t        n_rowblk = 0;
    const uint8_t *col_is_u = nullptr; // per local column (L+H): 1 if displacement dof (multi-GPU only)
    cudaStream_t   stream = nullptr;
    // multi-GPU hooks (null on one GPU): fill x[L..L+H) from the owners; sum n doubles over ranks
    void (*halo)(void *ctx, double *x)               = nullptr;
    void (*allreduce)(void *ctx, double *v, int n)   = nullptr;
    void *halo_ctx                                   = nullptr;
    void *precond_ctx                                          = nullptr;
    int      use_graphs     = 1;       // capture the MINRES iteration into CUDA graphs (single GPU)
    int64_t *launch_counter = nullptr; // kernels launched (graph replays counted node by node)
    // full replacement of the diagonal preconditioner: z = P^{-1} r  (pe_precond.cu)
    void (*precond_apply)(void *ctx, const double *r, double *z) = nullptr;
    Gap    gap;                 // vector layout of the owned unknowns (three-field system)
    double bnorm2 = -1.;        // >= 0: ||b||^2 the stopping test is relative to (else computed from b)
    // optional: *nrm2_dev += ||true residual(x)||^2 of this rank (else b - A x with the ctx matrix)
    void (*true_residual)(void *ctx, const double *x, double *nrm2_dev) = nullptr;
    Recycle *recycle = nullptr;       // GMRES: projection space carried from solve to solve (null: none)
    void (*on_restart)(void *ctx, double *x) = nullptr; // GMRES: called on the iterate before every cycle
    bool          x0_is_zero = false;  // initial guess known to be zero: ||b||_{P^-1} = the initial eta
    MinresGraphs *graphs     = nullptr; // cache of the captured iteration (null: capture per solve)
    GmresGraphs  *gmres_graphs = nullptr;
    uint64_t      graph_key  = 0;
    double       *calib      = nullptr; // calibration factor between the P^-1 norm and the 2-norm, kept across solves
    double (*anorm)(void *ctx) = nullptr; // lazy ||A||_inf of the reference matrix (rounding-floor test only)
    int64_t       floor_rows = 0;       // leading logical rows whose ||x|| enters the floor test (0 = all; three-field
                                        // vectors: the (u, p) part only -- xi = O(lambda) is not an unknown of the reference)
    cudaEvent_t  *ev_ring    = nullptr; // 16 events (disable-timing) for the look-ahead convergence test
  };

  void   build_row_blocks(int64_t n_rows, const int64_t *rowptr, std::vector<int32_t> &blk);
  void   spmv_signed(const SolverCtx &c, const double *x, double *y, double *dot);
  void   scale_vector(const SolverCtx &c, double *x, double a);
  // *nrm2_dev += || b - A x ||^2 over the owned rows (plain A: no sign flip); y = A x is scratch (L long)
  void   residual_norm2(const SolverCtx &c, const double *b, const double *x, double *y, double *nrm2_dev);
  void   norm2(const SolverCtx &c, const double *x, double *nrm2_dev);
  // pos[r] = index of A(r, r) in the CSR arrays of the ctx matrix (once per pattern)
  void   diagonal_positions(const SolverCtx &c, int64_t *pos);
  // *out_dev = max row sum of |A| over the owned rows of the ctx matrix (this rank)
  void   matrix_inf_norm(const SolverCtx &c, double *out_dev);
  void   lincomb3(int64_t n, double w0, const double *a, double w1, const double *b, double w2, const double *c, double *out,
                  cudaStream_t s);
  // diag[r] = A(r, r) for the ctx matrix (owned rows)
  void   extract_diagonal(const SolverCtx &c, double *diag);
  int    build_preconditioner(const SolverCtx &c, double *diag_full, double *pinv);
  int    minres_solve(const SolverCtx &c, const double *b, double *x, const double *pinv, double *work[7],
                      MinresState *st, double *h_pinned, double rel_tol, int max_it, pe_solve_stats *stats);
  // restarted flexible GMRES on the row-signed system (c.n_pos = first negated row); c.precond_apply = P^-1
  int    fgmres_solve(const SolverCtx &c, const double *b, double *x, double *V, double *Z, int64_t stride, double *w, int m,
                      void *st_dev, double *h_pinned, double rel_tol, int max_it, pe_solve_stats *stats);
  size_t gmres_state_bytes();
  int    cg_solve(const SolverCtx &c, const double *b, double *x, double *work[3], void *st_dev, double *h_pinned,
                  double rel_tol, int max_it, pe_solve_stats *stats);
  size_t minres_state_bytes();
} // namespace pe
