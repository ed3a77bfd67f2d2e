// Internal state behind pe_handle.  Device memory layout (HBM), all structure-of-arrays so that a
// warp of 32 consecutive cells reads/writes 32 consecutive words:
//   vx, vy          double [n_vertices]
//   cell_vertices   uint32 [4][n_asm]          asm cells = owned cells then ghost cells
//   cell_dofs       int32  [n_loc][n_asm]      LOCAL numbering: < L owned row, >= L halo column
//   kcell           double [n_asm]
//   cell_flags      uint32 [n_asm]             3 bits of PE_FACE_* per face + CGQ1 vertex mask
//   rowptr          int64  [L+1]   colidx int32 [nnz] (local numbering)   values double [nnz]
//   scatter_off     uint8  [n_loc*n_loc][n_asm]  position of entry (i,j) inside CSR row of dof i
//   rhs, sol, ...   double [L (+H)]
#pragma once
#include "../../include/poroelas_b200.h"
#include "pe_cell.cuh"
#include "pe_host.h"

#include <cuda_runtime.h>

#include <cstdlib>
#include <string>
#include <vector>

struct ncclComm;
namespace pe { struct MgHierarchy; struct BlockPrecond; struct K3System; struct CellMg; struct MinresGraphs; struct GmresGraphs; struct Recycle; }

namespace pe
{
  template <class T>
  struct DevBuf
  {
    T     *p = nullptr;
    size_t n = 0;
    ~DevBuf() { release(); }
    DevBuf()                = default;
    DevBuf(const DevBuf &)  = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    DevBuf(DevBuf &&o) noexcept : p(o.p), n(o.n)
    {
      o.p = nullptr;
      o.n = 0;
    }
    DevBuf &
    operator=(DevBuf &&o) noexcept
    {
      if (this != &o)
        {
          release();
          p   = o.p;
          n   = o.n;
          o.p = nullptr;
          o.n = 0;
        }
      return *this;
    }
    void
    release()
    {
      if (p)
        cudaFree(p);
      p = nullptr;
      n = 0;
    }
    cudaError_t
    alloc(const size_t count)
    {
      release();
      n = count;
      if (count == 0)
        return cudaSuccess;
      return cudaMalloc((void **)&p, count * sizeof(T));
    }
  };

  // Row-slab distribution of the multigrid grids (lexicographic, rows contiguous): rank r owns rows
  // [bounds[r], bounds[r+1]) of every distributed level and keeps one halo row on each side, exchanged in
  // place with the two neighbouring ranks.  Hooks are filled by pe_comm.cu (NCCL); world == 1: unused.
  struct GridComm
  {
    int   rank = 0, world = 1;
    void *ctx  = nullptr;
    int   dist_min_n = 1024; // smallest level (cells per side) distributed over the ranks
    // send my first / last owned row (a, b-1) down / up, receive rows a-1 and b; n_comp arrays comp_stride apart
    void (*exchange_rows)(void *ctx, double *p, int64_t row_len, int n_comp, int64_t comp_stride, int64_t a, int64_t b) = nullptr;
    // every rank receives the slabs [bounds[q], bounds[q+1]) of all ranks q (in place)
    void (*allgather_rows)(void *ctx, double *p, int64_t row_len, int n_comp, int64_t comp_stride, const int64_t *bounds) = nullptr;
    // out[slab of rank q] = sum over the ranks of in[slab of rank q], delivered to rank q only (half the traffic of
    // an all-reduce: a distributed level only needs its own rows of the right-hand side)
    void (*reduce_rows)(void *ctx, const double *in, double *out, int64_t row_len, int n_comp, int64_t comp_stride,
                        const int64_t *bounds) = nullptr;
    // box <-> slab transfers (pe_comm.cu): only the overlapping rectangles travel.  kind 0 = vertex grid, 1 = cell grid.
    //   boxes_to_slabs: out[my slab] = sum over ranks of in[their box];  slabs_to_boxes: p[my box + 1] from the owners
    void (*boxes_to_slabs)(void *ctx, const double *in, double *out, int64_t row_len, int n_comp, int64_t comp_stride,
                           const int64_t *bounds, int kind) = nullptr;
    void (*slabs_to_boxes)(void *ctx, double *p, int64_t row_len, int n_comp, int64_t comp_stride, const int64_t *bounds,
                           int kind) = nullptr;
  };

  // bounding box (cell indices, half open) of the cells a rank assembles on a structured mesh: the grid <-> Biot
  // vector transfer kernels of the preconditioner only visit this part of the grids
  struct CellBox
  {
    int i0 = 0, i1 = 0, j0 = 0, j1 = 0;
  };

  // Tuning / A-B switches of the library, set through pe_set_tuning (no environment variables are read by the hot path)
  struct Tuning
  {
    int    asm_gather     = 0;    // gather assembly instead of RED scatter (measured slower, kept as a tested alternative)
    double mg_omega2      = -1.;  // weight of every second vertex-grid sweep (< 0: 1.5 omega)
    double cmg_omega      = -1.;  // weights of the cell-grid sweeps (< 0: built-in 0.6 / 1.2)
    double cmg_omega2     = -1.;
    double face_omega     = -1.;  // damped Jacobi on the face Schur complement (< 0: 0.65)
    int    mg_replicated  = 0;    // multi-GPU: every rank runs the complete V-cycles (first multi-GPU version)
    int    mg_whole_slabs = 0;    // multi-GPU: reduce / all-gather whole slabs instead of box x slab rectangles
    int    one_lane       = 0;    // multi-GPU: no second stream + communicator for the pressure pipeline
    int    mg_fused_n     = -1;   // largest level handled by the single-CTA multigrid kernel (< 0: built-in)
    int    mg_dist_min_n  = 1024; // smallest grid level (cells per side) that is distributed over the ranks
    int    esm_dfma       = 0;    // WG(Q2,Q2;RT[2]): the first, DFMA-only cell kernel instead of the DMMA one (A/B)
    int    reuse_operator = 0;    // 1 = cached mode: keep the solver operator + preconditioner while the inputs of the
                                  // matrix are unchanged (default 0: rebuilt by every solve, as the reference refactorises)
  };

  struct HaloPeer
  {
    int                  rank = -1;
    std::vector<int32_t> send_rows; // local owned rows this peer needs from us
    int64_t              recv_off = 0, recv_cnt = 0; // slice of our halo [L + recv_off, ...)
    DevBuf<int32_t>      d_send_rows;
    DevBuf<double>       d_send_buf, d_send_buf_xi;
  };
} // namespace pe

struct pe_handle
{
  pe_config    cfg{};
  std::string  err;
  int          device = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t  ev[8]  = {};

  // ---- host side
  // dim == 3 (CGQ1-WG on hexahedra, first step: local matrices): vertices and 8 vertex ids per cell, deal.II order
  struct Mesh3
  {
    std::vector<double>   x, y, z;
    std::vector<uint32_t> cell_vertices; // [n_cells][8]
    int64_t               n_cells = 0;
    std::vector<uint8_t>  face_kinds;    // [n_cells][6], PE_FACE_* bits; empty = no boundary data
    double                traction[3] = {0., 0., 0.};
  } mesh3;
  // device side of assemble_system in 3-D (pe_api.cu: assemble3)
  struct Dev3
  {
    pe::DevBuf<double>   x, y, z, kc, A, b, old, cell_old;
    pe::DevBuf<uint32_t> cv, con_mask;
    pe::DevBuf<uint8_t>  neu_mask;
    pe::DevBuf<int32_t>  cdofs;
    pe::DevBuf<int>      err;
    int64_t          chunk = 0;
    bool             ready = false;
  } dev3;
  pe::HostMesh    mesh;
  pe::HostDofs    dofs;
  pe::Partition   part;
  pe::HostPattern pat;
  bool            have_mesh = false, have_dofs = false, have_pattern = false, device_ready = false;
  bool            matrix_valid = false;
  uint64_t        data_epoch = 1; // bumped when permeability / boundary data that enter the matrix change
  uint64_t        dofs_epoch = 1, sol_dofs_epoch = 0; // numbering generation / the one d_sol belongs to
  double          lambda = 1., mu = 1., alpha = 1., c0 = 0., k_scalar = 1.;
  std::vector<double> k_cell; // empty = scalar
  bool            k_lognormal = false;
  double          k_sigma = 0.;
  uint64_t        k_seed = 0;
  int             case_id = 0;
  double          case_params[5] = {1., 1., 1., 0., 1.};
  double          last_dt = 0.;

  // ---- device side
  int64_t n_asm = 0, n_bnd = 0, L = 0, H = 0, nnz = 0;
  pe::DevBuf<double>   d_vx, d_vy, d_k, d_values, d_rhs, d_sol;
  pe::DevBuf<uint32_t> d_cv, d_flags, d_rhs_first;
  pe::DevBuf<double>   d_k_morton; // per-cell K in cell (Morton) order for the multigrid levels
  std::vector<int64_t> bnd_color_begin;
  // atomics mode: chunks of asm cells and the value ranges (3 per chunk) zeroed before each chunk
  std::vector<int64_t> chunk_cell_begin, chunk_zero, chunk_rows;
  // gather assembly (pe_assemble.cu): four slot bytes per CSR entry, the <= 4 cells adjacent to every owned row,
  // the staging buffer T[NZ][chunk_cells]; opt-in (PE_ASM_GATHER=1), measured slower than the default RED scatter
  bool                 gather_ok = false;
  pe::DevBuf<uint8_t>  d_gsrc;
  pe::DevBuf<int32_t>  d_rowadj, d_tile_ghost_ptr, d_tile_ghost;
  pe::DevBuf<int64_t>  d_tile_rows;
  pe::DevBuf<double>   d_T;
  uint64_t             op_key       = 0;     // fingerprint of the matrix inputs the solver operator was built from
  bool                 op_key_valid = false; // false: the next solve rebuilds operator + preconditioner unconditionally
  int64_t              op_builds    = 0;     // update_solver_operator() calls so far (tests / tracing)
  int64_t              chunk_cells = 32768;
  // requested by pe_set_solver_options; scatter mode takes effect at the next pattern build, the chunk size at the
  // next (re)build of the chunk tables -- never on tables that were built for another mode
  int                  pending_scatter_mode = 0;
  int64_t              pending_chunk_cells  = 32768;
  pe::DevBuf<int32_t>  d_cdofs, d_colidx, d_bnd_list, d_rowblk, d_out_slot;
  int64_t              n_rowblk = 0;
  int                  spmv_kernel = 1; // 1 = stream, 0 = warp per row
  pe::DevBuf<int64_t>  d_rowptr, d_diag_pos;
  pe::DevBuf<uint8_t>  d_off, d_is_ghost; // d_is_ghost: 1 for ghost cells (multi-GPU only, else unallocated)
  pe::DevBuf<int>      d_err;
  // solver work vectors (each L+H long so that any of them can be the SpMV input)
  pe::DevBuf<double> d_w[12];
  pe::DevBuf<double> d_pinv;   // block-Jacobi preconditioner (inverse diagonal)
  pe::DevBuf<double> d_scal;   // small scalar scratch (dots)
  double            *h_scal = nullptr; // pinned
  double            *h_stage = nullptr; // pinned staging buffer for the halo part of a host vector (multi-GPU)
  size_t             h_stage_n = 0;

  // auxiliary-space multigrid preconditioner (pe_mg.cu); null = diagonal preconditioner only
  pe::MgHierarchy *mg = nullptr;
  pe::BlockPrecond *bp = nullptr;
  pe::CellMg       *cmg = nullptr; // cell-centred multigrid of the pressure block (pe_cmg.h)
  // three-field solver matrix (pe_k3.h); rebuilt values after every matrix assembly
  pe::K3System *k3 = nullptr;
  double        last_t = 0.;
  int              use_mg = 1, use_graphs = 1;
  double           mg_theta = 1.0, mg_omega = 0.6;
  int              mg_nu = 2;
  double           precond_ms = 0.;
  // Krylov state kept across solves: captured iteration graphs, calibration of the stopping estimate, ||A||_inf of
  // the assembled reference matrix (lazy, -1 = not computed), events of the look-ahead convergence test
  pe::MinresGraphs *graphs = nullptr;
  pe::GmresGraphs  *gmres_graphs = nullptr;
  int               krylov = 1, gmres_m = 16; // 0 = MINRES + block-diagonal P, 1 = FGMRES(m) + block-triangular P
  pe::DevBuf<double>  d_gm;                    // GMRES basis: (m + 1) v's, m z's
  pe::DevBuf<double>  d_recycle;               // projection space carried between solves: 2 k_max + 2 vectors
  pe::Recycle        *recycle = nullptr;
  int                 recycle_k = 4;           // pairs kept (0 = off)
  pe::DevBuf<uint8_t> d_gmres_state;
  uint64_t          graph_epoch = 1; // bumped whenever buffers or by-value kernel arguments of the iteration change
  double            graph_fp[8] = {};
  double            minres_calib = 1.;
  double            anorm = -1.;
  cudaEvent_t       ev_ring[16] = {};
  // initial guess by extrapolation in time from the previous solutions (uniform steps only): x_n, x_{n-1}, x_{n-2}
  pe::DevBuf<double> d_hist[2];
  double             hist_t[3] = {0., 0., 0.}; // times of d_sol (after a solve), d_hist[0], d_hist[1]
  int                hist_n = 0;               // number of valid solutions in (d_sol, d_hist[0], d_hist[1])
  int                extrapolate = 2;          // max order of the extrapolation (0 = plain warm start)

  // WG_Q2RT2 (EltStiffMat.cc): boundary dofs of interpolate_boundary_values (ESM:565-571), physical support points
  std::vector<int32_t> esm_bdofs;
  std::vector<double>  esm_bx, esm_by;
  pe::DevBuf<int32_t>  d_esm_bdofs;
  pe::DevBuf<double>   d_esm_bvals;
  pe::DevBuf<uint8_t>  d_esm_isbnd;

  // multi-GPU
  ncclComm                 *comm = nullptr;
  std::vector<pe::HaloPeer> peers;
  cudaStream_t              comm_stream = nullptr;
  cudaStream_t              stream2 = nullptr; // lane 1: pressure pipeline of the preconditioner (own NCCL communicator)
  int                       cur_lane = 0;
  pe::Tuning                tune;      // lane the host is enqueuing on (pe_comm.cu: lane_stream / lane_comm)

  // timings
  double  assemble_ms = 0., assemble_kernel_ms = 0., solve_ms = 0., spmv_ms_avg = 0.;
  int64_t launches = 0;
};
