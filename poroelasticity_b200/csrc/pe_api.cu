// C ABI of libporoelas_b200.so (include/poroelas_b200.h).  Host logic only: argument checking,
// device residency of the mesh / DoF / CSR arrays, kernel launches, timing.  No CPU compute path:
// every entry point that produces numbers launches the CUDA kernels or fails.
#include "pe_internal.h"
#include "pe_cmg.h"
#include "pe_k3.h"
#include "pe_kernels.h"
#include "pe_mg.h"
#include "pe_precond.h"
#include "pe_solver.h"
#include "pe_timer.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <climits>
#include <exception>

using namespace pe;

namespace pe
{
  // pe_comm.cu
  int  comm_init(pe_handle *h, const uint8_t id[128]);
  void comm_destroy(pe_handle *h);
  int  comm_setup_halo(pe_handle *h);
  void comm_halo_exchange(void *ctx, double *x);
  void comm_halo_exchange3(void *ctx, double *x);
  void comm_halo_exchange_p(void *ctx, double *x);
  int  comm_has_lanes(pe_handle *h);
  void comm_halo_exchange_xi(void *ctx, double *x);
  void comm_exchange_rows(void *ctx, double *p, int64_t row_len, int n_comp, int64_t comp_stride, int64_t a, int64_t b);
  void comm_allgather_rows(void *ctx, double *p, int64_t row_len, int n_comp, int64_t comp_stride, const int64_t *bounds);
  void comm_reduce_rows(void *ctx, const double *in, double *out, int64_t row_len, int n_comp, int64_t comp_stride,
                        const int64_t *bounds);
  void comm_allreduce(void *ctx, double *v, int n);
  void comm_boxes_to_slabs(void *ctx, const double *in, double *out, int64_t row_len, int n_comp, int64_t comp_stride,
                           const int64_t *bounds, int kind);
  void comm_slabs_to_boxes(void *ctx, double *p, int64_t row_len, int n_comp, int64_t comp_stride, const int64_t *bounds,
                           int kind);
  int  comm_set_box(pe_handle *h, const CellBox &box);
  void comm_allreduce_max(void *ctx, double *v, int n);
  int  comm_status(pe_handle *h);
  void comm_allreduce2(void *ctx, const double *in, double *out, int n);
  int  comm_unique_id(uint8_t id[128]);
} // namespace pe

namespace
{
  int
  fail(pe_handle *h, const int code, const std::string &msg)
  {
    if (h)
      h->err = msg;
    return code;
  }
  int
  cuda_fail(pe_handle *h, const cudaError_t e, const char *where)
  {
    return fail(h, PE_ERR_CUDA, std::string(where) + ": " + cudaGetErrorString(e));
  }
#define PE_CUDA(call)                                 \
  do                                                  \
    {                                                 \
      const cudaError_t e_ = (call);                  \
      if (e_ != cudaSuccess)                          \
        return cuda_fail(h, e_, #call);               \
    }                                                 \
  while (0)

  // Every entry point that touches the GPU runs on the handle's device, whatever the calling thread's current
  // device is (a host program may hold handles on several GPUs); the caller's device is restored on return.
  struct DeviceGuard
  {
    int  prev = -1;
    bool ok   = true;
    explicit DeviceGuard(const pe_handle *h)
    {
      if (!h || h->device < 0)
        return;
      if (cudaGetDevice(&prev) != cudaSuccess)
        prev = -1;
      if (prev != h->device)
        ok = cudaSetDevice(h->device) == cudaSuccess;
      else
        prev = -1;
    }
    ~DeviceGuard()
    {
      if (prev >= 0)
        cudaSetDevice(prev);
    }
  };
#define PE_DEVICE(h)                                                   \
  DeviceGuard device_guard_(h);                                        \
  if (!device_guard_.ok)                                               \
    return fail(const_cast<pe_handle *>(h), PE_ERR_CUDA, "cudaSetDevice(handle device) failed")

  template <class T>
  cudaError_t
  upload(DevBuf<T> &d, const T *src, const size_t n, cudaStream_t s)
  {
    cudaError_t e = d.alloc(n);
    if (e != cudaSuccess || n == 0)
      return e;
    return cudaMemcpyAsync(d.p, src, n * sizeof(T), cudaMemcpyHostToDevice, s);
  }

  // global dof id -> local id on this rank: owned rows first, then the halo columns; -1 = not visible
  int32_t
  global_to_local(const pe_handle *h, const uint32_t g)
  {
    const Partition &pt = h->part;
    const int64_t    l  = pt.global_to_owned(g);
    if (l >= 0)
      return (int32_t)l;
    const auto it = std::lower_bound(pt.ghost_cols.begin(), pt.ghost_cols.end(), g);
    if (it == pt.ghost_cols.end() || *it != g)
      return -1;
    return (int32_t)(pt.n_owned + (it - pt.ghost_cols.begin()));
  }

  DevParams
  make_params(const pe_handle *h, const double t, const double dt)
  {
    DevParams P;
    P.lambda      = h->lambda;
    P.mu          = h->mu;
    P.alpha       = h->alpha;
    P.c0          = h->c0;
    P.dt          = dt;
    P.traction[0] = h->mesh.traction[0];
    P.traction[1] = h->mesh.traction[1];
    P.case_id     = h->case_id == PE_CASE_COSCOS ? 1 : 0;
    const double *q = h->case_params;
    const double  st = std::sin(M_PI / 2. * t), ct = std::cos(M_PI / 2. * t);
    P.f_coef = M_PI * (q[0] + 2. * q[1] - q[2]) * st;
    P.s_coef = M_PI / 2. * ct * (q[3] + q[4]) + 2. * M_PI * M_PI * st;
    P.u_coef = st / (2. * M_PI);
    P.p_coef = st;
    return P;
  }

  // upload everything the kernels need; (re)build the scatter offsets and the boundary-cell list
  int
  ensure_device(pe_handle *h)
  {
    if (h->device_ready)
      return PE_OK;
    if (h->device < 0)
      return fail(h, PE_ERR_NO_DEVICE, "handle was created with PE_CFG_HOST_SETUP_ONLY: no hot path without a GPU");
    if (!h->have_mesh || !h->have_dofs || !h->have_pattern)
      return fail(h, PE_ERR_ARG, "mesh, dofs and sparsity pattern must be set before the hot path runs");
    delete h->mg;
    h->mg                  = nullptr;
    delete h->bp;
    h->bp                  = nullptr;
    h->stream2             = nullptr; // owned by the block preconditioner
    h->cur_lane            = 0;
    delete h->k3;
    h->k3                  = nullptr;
    h->op_key_valid        = false;
    delete h->cmg;
    h->cmg                 = nullptr;
    ++h->graph_epoch; // the captured iteration refers to buffers that are rebuilt below
    h->anorm               = -1.;
    cudaStream_t     s     = h->stream;
    const int        n_loc = h->dofs.n_loc;
    const Partition &pt    = h->part;
    const int64_t    n_asm = (int64_t)pt.asm_cells.size();
    h->n_asm               = n_asm;
    h->L                   = pt.n_owned;
    h->H                   = (int64_t)pt.ghost_cols.size();
    h->nnz                 = (int64_t)h->pat.colidx.size();

    PE_CUDA(upload(h->d_vx, h->mesh.x.data(), h->mesh.x.size(), s));
    PE_CUDA(upload(h->d_vy, h->mesh.y.data(), h->mesh.y.size(), s));

    // global -> local column map for the dofs this rank can see
    auto to_local = [h](const uint32_t g) -> int32_t { return global_to_local(h, g); };

    // SoA copies of the per-cell arrays for the cells this rank assembles
    std::vector<uint32_t> cv((size_t)4 * n_asm), flags((size_t)n_asm, 0);
    std::vector<int32_t>  cd((size_t)n_loc * n_asm), bnd;
    // CGQ1: mesh-wide displacement constraints (interpolate_boundary_values over the whole
    // DoFHandler inside the cell loop, CGQ1:1673-1679): flag every vertex of a DIR_U face
    std::vector<uint8_t> vflag;
    if (h->dofs.element == PE_CGQ1_WG)
      {
        vflag.assign((size_t)h->mesh.n_vertices, 0);
        for (int64_t c = 0; c < h->mesh.n_cells; ++c)
          for (int f = 0; f < 4; ++f)
            if (h->mesh.face_kinds[4 * c + f] & PE_FACE_DIR_U)
              for (int v = 0; v < 2; ++v)
                vflag[h->mesh.cell_vertices[4 * c + kFaceVertex[f][v]]] = 1;
      }
    for (int64_t a = 0; a < n_asm; ++a)
      {
        const int64_t c = pt.asm_cells[a];
        for (int v = 0; v < 4; ++v)
          cv[(size_t)v * n_asm + a] = h->mesh.cell_vertices[4 * c + v];
        for (int i = 0; i < n_loc; ++i)
          {
            const int32_t l = to_local(h->dofs.cell_dofs[(size_t)c * n_loc + i]);
            if (l < 0)
              return fail(h, PE_ERR_PATTERN, "a dof of an assembled cell is neither owned nor in the halo");
            cd[(size_t)i * n_asm + a] = l;
          }
        uint32_t fl = 0;
        for (int f = 0; f < 4; ++f)
          fl |= (uint32_t)(h->mesh.face_kinds[4 * c + f] & 7) << (3 * f);
        if (!vflag.empty())
          for (int v = 0; v < 4; ++v)
            if (vflag[h->mesh.cell_vertices[4 * c + v]])
              fl |= 1u << (12 + v);
        flags[a] = fl;
        if (fl)
          bnd.push_back((int32_t)a);
      }
    h->n_bnd = (int64_t)bnd.size();
    {
      // output slot of every assembled cell for the per-cell post-processing: owned cells in deal.II order
      std::vector<int32_t> slot((size_t)n_asm);
      for (int64_t a = 0; a < n_asm; ++a)
        slot[a] = (pt.asm_cells[a] >= pt.cell_begin && pt.asm_cells[a] < pt.cell_end) ? (int32_t)(pt.asm_cells[a] - pt.cell_begin) : -1;
      PE_CUDA(upload(h->d_out_slot, slot.data(), slot.size(), s));
    }
    h->d_is_ghost.release();
    if (pt.world > 1)
      {
        std::vector<uint8_t> gh((size_t)n_asm);
        for (int64_t a = 0; a < n_asm; ++a)
          gh[a] = pt.asm_cells[a] < pt.cell_begin || pt.asm_cells[a] >= pt.cell_end;
        PE_CUDA(upload(h->d_is_ghost, gh.data(), gh.size(), s));
      }
    // boundary cells per colour (bnd is ascending in asm index, asm order is colour-major)
    h->bnd_color_begin.assign((size_t)pt.n_colors + 1, 0);
    for (int col = 0; col <= pt.n_colors; ++col)
      h->bnd_color_begin[col] =
        std::lower_bound(bnd.begin(), bnd.end(), (int32_t)(col < pt.n_colors ? pt.color_begin[col] : n_asm)) - bnd.begin();
    PE_CUDA(upload(h->d_cv, cv.data(), cv.size(), s));
    PE_CUDA(upload(h->d_flags, flags.data(), flags.size(), s));
    PE_CUDA(upload(h->d_cdofs, cd.data(), cd.size(), s));
    PE_CUDA(upload(h->d_bnd_list, bnd.data(), bnd.size(), s));

    // CSR with local column numbering
    PE_CUDA(upload(h->d_rowptr, h->pat.rowptr.data(), h->pat.rowptr.size(), s));
    if (pt.world == 1)
      PE_CUDA(upload(h->d_colidx, h->pat.colidx.data(), h->pat.colidx.size(), s));
    else
      {
        std::vector<int32_t> lc(h->pat.colidx.size());
        for (size_t k = 0; k < lc.size(); ++k)
          lc[k] = to_local((uint32_t)h->pat.colidx[k]);
        PE_CUDA(upload(h->d_colidx, lc.data(), lc.size(), s));
      }
    {
      std::vector<int32_t> blk;
      build_row_blocks(h->L, h->pat.rowptr.data(), blk);
      h->n_rowblk = (int64_t)blk.size() - 1;
      PE_CUDA(upload(h->d_rowblk, blk.data(), blk.size(), s));
    }
    PE_CUDA(h->d_values.alloc((size_t)h->nnz));
    {
      // position of the diagonal entry of every owned row: the diagonal of an assembled matrix is then one gather
      h->d_diag_pos.release();
      SolverCtx c0;
      c0.L      = h->L;
      c0.rowptr = h->d_rowptr.p;
      c0.colidx = h->d_colidx.p;
      c0.stream = s;
      PE_CUDA(h->d_diag_pos.alloc((size_t)h->L));
      diagonal_positions(c0, h->d_diag_pos.p);
      ++h->launches;
    }
    PE_CUDA(h->d_rhs.alloc((size_t)(h->L + h->H)));
    // the resident old_solution survives a rebuild of the device tables (boundary kinds, permeability field,
    // chunk size ...) as long as the numbering is the same; it is zeroed on the first upload or a size change
    if (h->d_sol.n != (size_t)(h->L + h->H) || h->sol_dofs_epoch != h->dofs_epoch)
      {
        PE_CUDA(h->d_sol.alloc((size_t)(h->L + h->H)));
        PE_CUDA(cudaMemsetAsync(h->d_sol.p, 0, sizeof(double) * (size_t)(h->L + h->H), s));
        h->sol_dofs_epoch = h->dofs_epoch;
        h->hist_n         = 0;
      }
    PE_CUDA(h->d_err.alloc(1));
    PE_CUDA(cudaMemsetAsync(h->d_err.p, 0, sizeof(int), s));

    // permeability
    if (h->k_lognormal)
      {
        PE_CUDA(h->d_k.alloc((size_t)n_asm));
        DevBuf<int64_t> ids;
        PE_CUDA(upload(ids, pt.asm_cells.data(), pt.asm_cells.size(), s));
        PE_CUDA(launch_lognormal(n_asm, ids.p, 0, h->k_sigma, h->k_seed, h->d_k.p, s));
        ++h->launches;
        PE_CUDA(cudaStreamSynchronize(s));
      }
    else if (!h->k_cell.empty())
      {
        std::vector<double> k((size_t)n_asm);
        for (int64_t a = 0; a < n_asm; ++a)
          k[a] = h->k_cell[pt.asm_cells[a]];
        PE_CUDA(upload(h->d_k, k.data(), k.size(), s));
      }
    else
      h->d_k.release();
    // the same field in cell (Morton) order for the multigrid levels
    h->d_k_morton.release();
    if (h->k_lognormal || !h->k_cell.empty())
      {
        if (h->k_lognormal)
          {
            PE_CUDA(h->d_k_morton.alloc((size_t)h->mesh.n_cells));
            PE_CUDA(launch_lognormal(h->mesh.n_cells, nullptr, 0, h->k_sigma, h->k_seed, h->d_k_morton.p, s));
            ++h->launches;
          }
        else
          PE_CUDA(upload(h->d_k_morton, h->k_cell.data(), h->k_cell.size(), s));
      }

    // scatter offsets
    PE_CUDA(h->d_off.alloc((size_t)scatter_table_words(n_loc) * 4 * n_asm));
    PE_CUDA(cudaMemsetAsync(h->d_off.p, 0, (size_t)scatter_table_words(n_loc) * 4 * n_asm, s));
    PE_CUDA(h->d_rhs_first.alloc((size_t)n_asm));
    {
      // cells adjacent to each owned row, as asm indices
      std::vector<int32_t> pos((size_t)h->mesh.n_cells, -1);
      for (int64_t a = 0; a < n_asm; ++a)
        pos[pt.asm_cells[a]] = (int32_t)a;
      std::vector<int32_t> rc(pt.adj.size());
      for (size_t k = 0; k < rc.size(); ++k)
        rc[k] = pos[pt.adj[k]];
      DevBuf<int64_t> d_rcp;
      DevBuf<int32_t> d_rci;
      PE_CUDA(upload(d_rcp, pt.adj_ptr.data(), pt.adj_ptr.size(), s));
      PE_CUDA(upload(d_rci, rc.data(), rc.size(), s));
      // gather assembly tables (Biot elements, atomics mode): slot bytes per CSR entry + adjacent cells per row
      h->gather_ok = false;
      h->d_gsrc.release();
      h->d_rowadj.release();
      DevBuf<int> d_bad;
      // measured slower than the RED scatter (4.15 ms against 2.23 ms at N = 1024: both move 193 M eight-byte
      // contributions through L2 one 32-byte sector at a time), kept as an opt-in alternative: PE_ASM_GATHER=1
      const bool  want_gather = pt.scatter_mode == 0 && h->dofs.element != PE_WG_Q2RT2 && h->tune.asm_gather;
      if (want_gather)
        {
          bool ok = true;
          std::vector<int32_t> ra((size_t)4 * h->L, -1);
          for (int64_t l = 0; l < h->L && ok; ++l)
            {
              const int64_t n = pt.adj_ptr[l + 1] - pt.adj_ptr[l];
              if (n > 4)
                ok = false;
              for (int64_t k = 0; k < n && ok; ++k)
                ra[(size_t)4 * l + k] = rc[pt.adj_ptr[l] + k];
            }
          if (ok && h->d_gsrc.alloc((size_t)4 * h->nnz) == cudaSuccess && d_bad.alloc(1) == cudaSuccess &&
              upload(h->d_rowadj, ra.data(), ra.size(), s) == cudaSuccess)
            {
              PE_CUDA(cudaMemsetAsync(h->d_gsrc.p, 0xFF, (size_t)4 * h->nnz, s));
              PE_CUDA(cudaMemsetAsync(d_bad.p, 0, sizeof(int), s));
              h->gather_ok = true;
            }
          else
            {
              cudaGetLastError();
              h->d_gsrc.release();
            }
        }
      PE_CUDA(launch_scatter_offsets(n_loc, n_asm, h->L, h->d_cdofs.p, h->d_rowptr.p, h->d_colidx.p, d_rcp.p, d_rci.p,
                                     h->d_off.p, h->d_rhs_first.p, h->d_err.p, h->gather_ok ? h->d_gsrc.p : nullptr, d_bad.p, s));
      PE_CUDA(cudaStreamSynchronize(s));
      if (h->gather_ok)
        {
          int bad = 0;
          PE_CUDA(cudaMemcpy(&bad, d_bad.p, sizeof(int), cudaMemcpyDeviceToHost));
          if (bad)
            {
              h->gather_ok = false;
              h->d_gsrc.release();
              h->d_rowadj.release();
            }
        }
    }
    ++h->launches;
    int err = 0;
    PE_CUDA(cudaMemcpyAsync(&err, h->d_err.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    PE_CUDA(cudaStreamSynchronize(s));
    if (err == 1)
      return fail(h, PE_ERR_PATTERN, "an entry of a local matrix is not in the sparsity pattern");
    if (err == 2)
      return fail(h, PE_ERR_PATTERN, "a CSR row is longer than 128 entries (scatter offsets are 7 bit)");
    // chunks for the atomics mode: rows first touched by asm cells [c0, c1) are contiguous per block
    h->chunk_cell_begin.clear();
    h->chunk_zero.clear();
    h->chunk_rows.clear();
    if (pt.scatter_mode == 0)
      {
        std::vector<int32_t> pos((size_t)h->mesh.n_cells, -1);
        for (int64_t a = 0; a < n_asm; ++a)
          pos[pt.asm_cells[a]] = (int32_t)a;
        // first asm cell touching each owned row (ascending inside each block)
        std::vector<int32_t> owner((size_t)h->L);
        for (int64_t l = 0; l < h->L; ++l)
          {
            int32_t mn = INT32_MAX;
            for (int64_t k = pt.adj_ptr[l]; k < pt.adj_ptr[l + 1]; ++k)
              mn = std::min(mn, pos[pt.adj[k]]);
            owner[l] = mn;
          }
        const int64_t blk_end[3] = {pt.local_off[1], pt.local_off[2], h->L};
            h->chunk_cells = h->pending_chunk_cells;
        for (int64_t c0 = 0; c0 < n_asm; c0 += h->chunk_cells)
          {
            const int64_t c1 = std::min(n_asm, c0 + h->chunk_cells);
            h->chunk_cell_begin.push_back(c0);
            for (int b = 0; b < 3; ++b)
              {
                auto          bb = owner.begin() + pt.local_off[b], be = owner.begin() + blk_end[b];
                const int64_t ra = std::lower_bound(bb, be, (int32_t)c0) - owner.begin();
                const int64_t rb = std::lower_bound(bb, be, (int32_t)c1) - owner.begin();
                h->chunk_zero.push_back(h->pat.rowptr[ra]);
                h->chunk_zero.push_back(h->pat.rowptr[rb]);
                h->chunk_rows.push_back(ra);
                h->chunk_rows.push_back(rb);
              }
          }
        h->chunk_cell_begin.push_back(n_asm);
        // tiles of the gather assembly: rows first touched by every run of kTileCells cells, and the cells outside
        // the tile (inside the same chunk) that are adjacent to those rows
        h->d_tile_rows.release();
        h->d_tile_ghost_ptr.release();
        h->d_tile_ghost.release();
        if (h->gather_ok)
          {
            const int64_t        n_tiles = (n_asm + kTileCells - 1) / kTileCells;
            std::vector<int64_t> trow((size_t)6 * n_tiles);
            std::vector<int32_t> gptr((size_t)n_tiles + 1, 0), glist;
            bool                 ok = h->chunk_cells % kTileCells == 0;
            std::vector<int32_t> tmp;
            for (int64_t t = 0; t < n_tiles && ok; ++t)
              {
                const int64_t c0 = t * kTileCells, c1 = std::min(n_asm, c0 + kTileCells);
                const int64_t ch0 = (c0 / h->chunk_cells) * h->chunk_cells, ch1 = std::min(n_asm, ch0 + h->chunk_cells);
                tmp.clear();
                for (int b = 0; b < 3; ++b)
                  {
                    auto          bb = owner.begin() + pt.local_off[b], be = owner.begin() + blk_end[b];
                    const int64_t ra = std::lower_bound(bb, be, (int32_t)c0) - owner.begin();
                    const int64_t rb = std::lower_bound(bb, be, (int32_t)c1) - owner.begin();
                    trow[(size_t)6 * t + 2 * b]     = ra;
                    trow[(size_t)6 * t + 2 * b + 1] = rb;
                    for (int64_t l = ra; l < rb; ++l)
                      for (int64_t k = pt.adj_ptr[l]; k < pt.adj_ptr[l + 1]; ++k)
                        {
                          const int32_t ac = pos[pt.adj[k]];
                          if ((ac < c0 || ac >= c1) && ac >= ch0 && ac < ch1)
                            tmp.push_back(ac);
                        }
                  }
                std::sort(tmp.begin(), tmp.end());
                tmp.erase(std::unique(tmp.begin(), tmp.end()), tmp.end());
                if ((int)tmp.size() > kTileGhostMax)
                  ok = false;
                glist.insert(glist.end(), tmp.begin(), tmp.end());
                gptr[(size_t)t + 1] = (int32_t)glist.size();
              }
            if (ok)
              {
                PE_CUDA(upload(h->d_tile_rows, trow.data(), trow.size(), s));
                PE_CUDA(upload(h->d_tile_ghost_ptr, gptr.data(), gptr.size(), s));
                if (glist.empty())
                  glist.push_back(0);
                PE_CUDA(upload(h->d_tile_ghost, glist.data(), glist.size(), s));
                PE_CUDA(cudaStreamSynchronize(s));
              }
          }
      }
    if (pt.world > 1)
      {
        const int rc = comm_setup_halo(h);
        if (rc != PE_OK)
          return rc;
      }
    if (h->dofs.element == PE_WG_Q2RT2)
      {
        // interpolate_boundary_values(dof_handler, 0, BoundaryValues, boundary_values, face_pressure_mask), ESM:565-571:
        // the three FE_FaceQ(2) dofs of every pressure-Dirichlet face, support points t = 0, 1/2, 1 along the line
        h->esm_bdofs.clear();
        h->esm_bx.clear();
        h->esm_by.clear();
        std::vector<uint8_t> isb((size_t)(h->L + h->H), 0);
        for (int64_t c = 0; c < h->mesh.n_cells; ++c)
          for (int f = 0; f < 4; ++f)
            if (h->mesh.face_kinds[4 * c + f] & PE_FACE_DIR_P)
              {
                const uint32_t v0 = h->mesh.cell_vertices[4 * c + kFaceVertex[f][0]],
                               v1 = h->mesh.cell_vertices[4 * c + kFaceVertex[f][1]];
                for (int k = 0; k < 3; ++k)
                  {
                    const int32_t l = to_local(h->dofs.cell_dofs[(size_t)c * n_loc + 3 * f + k]);
                    if (l < 0 || l >= h->L || isb[l])
                      continue;
                    isb[l] = 1;
                    h->esm_bdofs.push_back(l);
                    h->esm_bx.push_back(h->mesh.x[v0] + .5 * k * (h->mesh.x[v1] - h->mesh.x[v0]));
                    h->esm_by.push_back(h->mesh.y[v0] + .5 * k * (h->mesh.y[v1] - h->mesh.y[v0]));
                  }
              }
        PE_CUDA(upload(h->d_esm_bdofs, h->esm_bdofs.data(), h->esm_bdofs.size(), s));
        PE_CUDA(upload(h->d_esm_isbnd, isb.data(), isb.size(), s));
        PE_CUDA(h->d_esm_bvals.alloc(h->esm_bdofs.size()));
        PE_CUDA(cudaStreamSynchronize(s));
      }
    h->device_ready = true;
    h->matrix_valid = false;
    return PE_OK;
  }

  AsmArgs
  make_args(pe_handle *h, const bool do_matrix)
  {
    AsmArgs a{};
    a.n_asm     = h->n_asm;
    a.n_bnd     = h->n_bnd;
    a.L         = h->L;
    a.vx        = h->d_vx.p;
    a.vy        = h->d_vy.p;
    a.kcell     = h->d_k.p;
    a.k_scalar  = h->k_scalar;
    a.esm_variant = h->tune.esm_dfma;
    a.sol       = h->d_sol.p;
    a.cv        = h->d_cv.p;
    a.flags     = h->d_flags.p;
    a.cdofs     = h->d_cdofs.p;
    a.bnd_list  = h->d_bnd_list.p;
    a.rowptr    = h->d_rowptr.p;
    a.off       = h->d_off.p;
    a.rhs_first = h->d_rhs_first.p;
    a.values    = h->d_values.p;
    a.rhs       = h->d_rhs.p;
    a.do_matrix = do_matrix ? 1 : 0;
    a.atomic_mode = h->part.scatter_mode == 0 ? 1 : 0;
    return a;
  }

  // matrix (+ rhs) assembly of all cells into a.values / a.rhs in the mode chosen at setup
  cudaError_t
  run_assembly(pe_handle *h, const AsmArgs &a, const DevParams &P, int *nl)
  {
    cudaStream_t s = h->stream;
    if (h->part.scatter_mode != 0)
      return launch_assemble(h->cfg.element, a, P, h->part.n_colors, h->part.color_begin.data(), h->bnd_color_begin.data(), s, nl);
    const int n_chunks = (int)h->chunk_cell_begin.size() - 1;
    if (h->gather_ok)
      {
        const int    nz = h->cfg.element == PE_BR1_WG ? NzCount<0>::value : NzCount<1>::value;
        const size_t nT = (size_t)nz * h->chunk_cells;
        if (h->d_T.n < nT && h->d_T.alloc(nT) != cudaSuccess)
          return cudaErrorMemoryAllocation;
        GatherTiles tiles;
        tiles.tile_rows      = h->d_tile_rows.p;
        tiles.tile_ghost_ptr = h->d_tile_ghost_ptr.p;
        tiles.tile_ghost     = h->d_tile_ghost.p;
        return launch_assemble_gather(h->cfg.element, a, P, n_chunks, h->chunk_cell_begin.data(), h->chunk_rows.data(),
                                      h->d_T.p, h->chunk_cells, h->d_gsrc.p, h->d_rowadj.p,
                                      h->d_tile_rows.p ? &tiles : nullptr, s, nl);
      }
    return launch_assemble_chunked(h->cfg.element, a, P, n_chunks, h->chunk_cell_begin.data(), h->chunk_zero.data(), s, nl);
  }

  // host vector in global block order -> device vector in local numbering (owned + halo)
  int
  upload_global_vector(pe_handle *h, const double *src, double *dst_dev)
  {
    const Partition &pt = h->part;
    if (pt.world == 1)
      {
        PE_CUDA(cudaMemcpyAsync(dst_dev, src, sizeof(double) * (size_t)h->L, cudaMemcpyHostToDevice, h->stream));
        return PE_OK;
      }
    // the three owned ranges are contiguous in the global vector: copy them straight from the caller's buffer
    // (PCIe speed when it is pinned); only the halo entries are gathered through a pinned staging buffer
    for (int b = 0; b < 3; ++b)
      if (pt.row_end[b] > pt.row_begin[b])
        PE_CUDA(cudaMemcpyAsync(dst_dev + pt.local_off[b], src + pt.row_begin[b],
                                sizeof(double) * (size_t)(pt.row_end[b] - pt.row_begin[b]), cudaMemcpyHostToDevice, h->stream));
    if (h->H > 0)
      {
        if (h->h_stage_n < (size_t)h->H)
          {
            if (h->h_stage)
              cudaFreeHost(h->h_stage);
            h->h_stage   = nullptr;
            h->h_stage_n = 0;
            PE_CUDA(cudaMallocHost((void **)&h->h_stage, sizeof(double) * (size_t)h->H));
            h->h_stage_n = (size_t)h->H;
          }
        for (int64_t k = 0; k < h->H; ++k)
          h->h_stage[k] = src[pt.ghost_cols[k]];
        PE_CUDA(cudaMemcpyAsync(dst_dev + h->L, h->h_stage, sizeof(double) * (size_t)h->H, cudaMemcpyHostToDevice, h->stream));
      }
    PE_CUDA(cudaStreamSynchronize(h->stream));
    return PE_OK;
  }

  int
  assemble_impl(pe_handle *h, const double t, const double dt, const double *old_solution, const bool do_matrix)
  {
    int rc = ensure_device(h);
    if (rc != PE_OK)
      return rc;
    if (!do_matrix && !h->matrix_valid)
      return fail(h, PE_ERR_ARG, "pe_assemble_rhs_only needs a previous pe_assemble_system");
    cudaStream_t s = h->stream;
    PE_CUDA(cudaEventRecord(h->ev[0], s));
    if (h->cfg.element == PE_WG_Q2RT2)
      {
        // WGDarcyEquation::assemble_system, ESM:333-578 (stationary: t, dt and old_solution are not used)
        if (h->part.world != 1)
          return fail(h, PE_ERR_ARG, "WG_Q2RT2: one GPU only");
        PE_CUDA(cudaMemsetAsync(h->d_values.p, 0, sizeof(double) * (size_t)h->nnz, s));
        PE_CUDA(cudaMemsetAsync(h->d_rhs.p, 0, sizeof(double) * (size_t)(h->L + h->H), s));
        PE_CUDA(cudaMemsetAsync(h->d_sol.p, 0, sizeof(double) * (size_t)(h->L + h->H), s)); // solution.reinit, ESM:303
        const AsmArgs a = make_args(h, true);
        PE_CUDA(cudaEventRecord(h->ev[1], s));
        int nl = 0;
        PE_CUDA(launch_esm_assemble(a, h->case_id, s, &nl));
        // BoundaryValues::value at the support points (ESM:147), then apply_boundary_values (ESM:572-575)
        std::vector<double> bv(h->esm_bdofs.size(), 0.);
        if (h->case_id == PE_CASE_DARCY_COSCOS)
          for (size_t k = 0; k < bv.size(); ++k)
            bv[k] = std::cos(M_PI * h->esm_bx[k]) * std::cos(M_PI * h->esm_by[k]);
        if (!bv.empty())
          PE_CUDA(cudaMemcpyAsync(h->d_esm_bvals.p, bv.data(), sizeof(double) * bv.size(), cudaMemcpyHostToDevice, s));
        PE_CUDA(launch_esm_boundary_values((int)bv.size(), h->d_esm_bdofs.p, h->d_esm_bvals.p, h->d_esm_isbnd.p,
                                           h->d_rowptr.p, h->d_colidx.p, h->d_values.p, h->d_rhs.p, h->d_sol.p, s));
        h->launches += nl + 2;
        PE_CUDA(cudaEventRecord(h->ev[2], s));
        PE_CUDA(cudaEventSynchronize(h->ev[2]));
        float ms = 0.f;
        cudaEventElapsedTime(&ms, h->ev[0], h->ev[2]);
        h->assemble_ms = ms;
        cudaEventElapsedTime(&ms, h->ev[1], h->ev[2]);
        h->assemble_kernel_ms = ms;
        h->matrix_valid       = true;
        return PE_OK;
      }
    if (old_solution)
      {
        rc = upload_global_vector(h, old_solution, h->d_sol.p);
        if (rc != PE_OK)
          return rc;
      }
    else if (h->part.world > 1)
      comm_halo_exchange(h, h->d_sol.p);
    // system_matrix = 0; system_rhs = 0 (BR1new:687-688).  In the colour-ordered mode it is implicit:
    // the first writer of every entry stores instead of adding (pe_assemble.cu)
    if (h->part.scatter_mode == 0)
      PE_CUDA(cudaMemsetAsync(h->d_rhs.p, 0, sizeof(double) * (size_t)(h->L + h->H), s));
    const DevParams P = make_params(h, t, dt);
    const AsmArgs   a = make_args(h, do_matrix);
    PE_CUDA(cudaEventRecord(h->ev[1], s));
    int nl = 0;
    if (h->part.scatter_mode == 0 && do_matrix)
      PE_CUDA(run_assembly(h, a, P, &nl));
    else
      PE_CUDA(launch_assemble(h->cfg.element, a, P, h->part.n_colors, h->part.color_begin.data(),
                              h->bnd_color_begin.data(), s, &nl));
    h->launches += nl;
    PE_CUDA(cudaEventRecord(h->ev[2], s));
    PE_CUDA(cudaEventSynchronize(h->ev[2]));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, h->ev[0], h->ev[2]);
    h->assemble_ms = ms;
    cudaEventElapsedTime(&ms, h->ev[1], h->ev[2]);
    h->assemble_kernel_ms = ms;
    h->last_dt            = dt;
    h->last_t             = t;
    if (do_matrix)
      {
        h->matrix_valid = true;
        if (h->k3)
          h->k3->values_valid = false;
      }
    return PE_OK;
  }

  SolverCtx
  make_solver_ctx(pe_handle *h)
  {
    SolverCtx c;
    c.L          = h->L;
    c.H          = h->H;
    c.n_pos      = h->part.local_off[1]; // owned u rows come first
    c.pint_begin = h->part.local_off[1];
    c.pint_end   = h->part.local_off[2];
    c.rowptr     = h->d_rowptr.p;
    c.colidx     = h->d_colidx.p;
    c.values     = h->d_values.p;
    c.stream     = h->stream;
    c.launch_counter = &h->launches;
    c.use_graphs     = h->use_graphs;
    c.diag_pos       = h->d_diag_pos.n == (size_t)h->L ? h->d_diag_pos.p : nullptr;
    if (h->spmv_kernel == 1)
      {
        c.rowblk   = h->d_rowblk.p;
        c.n_rowblk = h->n_rowblk;
      }
    if (h->part.world > 1)
      {
        c.halo      = comm_halo_exchange;
        c.allreduce = comm_allreduce;
        c.halo_ctx  = h;
      }
    return c;
  }

  // ---- three-field MINRES solve (pe_k3.h) ------------------------------------------------------------
  void
  hook_precond3(void *ctx, const double *r, double *z)
  {
    pe_handle *h = (pe_handle *)ctx;
    precond_apply(*h->bp, *h->mg, *h->cmg, h->d_pinv.p, r, z, h->stream, &h->launches);
    k3_precond_xi(*h->k3, h->mu, r, z, h->stream);
    ++h->launches;
  }
  void
  hook_set_lane(void *ctx, int lane)
  {
    ((pe_handle *)ctx)->cur_lane = lane;
  }
  // block upper-triangular variant for FGMRES on the row-signed system: z_xi first, then the displacement block on
  // r_u + B^T z_xi; the pressure block is independent of both (runs on the forked stream)
  void
  hook_precond3_tri(void *ctx, const double *r, double *z)
  {
    pe_handle *h = (pe_handle *)ctx;
    k3_precond_xi(*h->k3, h->mu, r, z, h->stream);
    ++h->launches;
    if (h->part.world > 1)
      comm_halo_exchange_xi(h, z);
    // the pressure rows of the GMRES operator are scaled: P^-1 acts on the unscaled residual
    const double *rp = nullptr;
    if (h->k3->p_scaled)
      {
        k3_p_vec_scale(*h->k3, true, r, h->bp->t5.p, h->stream);
        ++h->launches;
        rp = h->bp->t5.p;
      }
    precond_apply(*h->bp, *h->mg, *h->cmg, h->d_pinv.p, r, z, h->stream, &h->launches, z, rp);
  }
  // xi := -lambda W^-1 B u + alpha p_int from the (u, p) part of the iterate
  void
  hook_consistent_xi(void *ctx, double *x3)
  {
    pe_handle *h = (pe_handle *)ctx;
    if (h->part.world > 1)
      comm_halo_exchange(h, x3);
    k3_init_xi(*h->k3, x3, h->stream);
    ++h->launches;
  }
  void
  hook_true_residual(void *ctx, const double *x3, double *nrm2_dev)
  {
    // || b - A (u, p) || with the reference's own matrix: the (u, p) part of x3 is a two-field vector
    pe_handle *h = (pe_handle *)ctx;
    SolverCtx  c = make_solver_ctx(h);
    if (c.halo)
      c.halo(c.halo_ctx, const_cast<double *>(x3));
    residual_norm2(c, h->d_rhs.p, x3, h->d_w[9].p, nrm2_dev);
    h->launches += 2;
  }

  // ||A||_inf of the assembled reference matrix, once per assembled matrix and only when the rounding-floor test
  // of the solver asks for it
  double
  hook_anorm(void *ctx)
  {
    pe_handle *h = (pe_handle *)ctx;
    if (h->anorm >= 0.)
      return h->anorm;
    SolverCtx c = make_solver_ctx(h);
    double   *d = h->d_scal.p + 44;
    matrix_inf_norm(c, d);
    ++h->launches;
    if (h->part.world > 1)
      comm_allreduce_max(h, d, 1);
    cudaMemcpyAsync(h->h_scal + 9, d, sizeof(double), cudaMemcpyDeviceToHost, h->stream);
    if (cudaStreamSynchronize(h->stream) != cudaSuccess)
      return 0.;
    h->anorm = h->h_scal[9];
    return h->anorm;
  }

  // the solver operator: second assembly pass A0 = A(lambda0, alpha = 1) into k3.values0, K3 values,
  // block preconditioner data.  Once per assembled matrix.
  // Fingerprint of everything that enters the VALUES of the assembled matrix (the matrix never depends on t, only the
  // right-hand side does): material data, time step, permeability field, boundary kinds (data_epoch), mesh / numbering
  // (dofs_epoch).  Equal fingerprints <=> pe_assemble_system produced the same matrix up to the summation order of the
  // scatter, so the solver operator, the preconditioner and the projection space of earlier solves remain valid.
  uint64_t
  matrix_fingerprint(const pe_handle *h)
  {
    uint64_t key = 1469598103934665603ull;
    auto     mix = [&key](const void *p, const size_t n) {
      const unsigned char *b = (const unsigned char *)p;
      for (size_t i = 0; i < n; ++i)
        key = (key ^ b[i]) * 1099511628211ull;
    };
    const double fp[7] = {h->lambda, h->mu, h->alpha, h->c0, h->k_scalar, h->last_dt, h->k_sigma};
    mix(fp, sizeof(fp));
    mix(&h->k_seed, sizeof(h->k_seed));
    mix(&h->k_lognormal, sizeof(h->k_lognormal));
    mix(&h->dofs_epoch, sizeof(h->dofs_epoch));
    mix(&h->data_epoch, sizeof(h->data_epoch));
    return key;
  }

  int
  update_solver_operator(pe_handle *h)
  {
    K3System    &k = *h->k3;
    cudaStream_t s = h->stream;
    if (k.values0.n < (size_t)h->nnz)
      if (k.values0.alloc((size_t)h->nnz) != cudaSuccess)
        return 2;
    // lambda = lambda0 + lambda_s with the xi block carrying lambda_s (>= a floor so that 1/lambda_s exists)
    const double lambda_s = std::max(h->lambda, 1e-6 * h->mu);
    const double lambda0  = h->lambda - lambda_s;
    {
      // by-value kernel arguments of the captured iteration: a change invalidates the cached graphs
      const double fp[8] = {lambda_s, h->mu, h->mg_omega, (double)h->mg_nu, (double)h->use_mg, h->alpha, h->mg_theta, 0.};
      if (std::memcmp(fp, h->graph_fp, sizeof(fp)) != 0)
        {
          std::memcpy(h->graph_fp, fp, sizeof(fp));
          ++h->graph_epoch;
        }
    }
    h->anorm = -1.;
    ++h->op_builds;
    DevParams    P        = make_params(h, h->last_t, h->last_dt);
    P.lambda              = lambda0;
    P.alpha               = 1.;
    AsmArgs a             = make_args(h, true);
    a.values              = k.values0.p;
    a.rhs                 = nullptr; // matrix only
    a.sol                 = nullptr;
    int nl                = 0;
    cudaError_t e;
    NvtxRange nvtx("solver operator: A0, K3, preconditioner");
    PE_MARK(s, "setup: (start)");
    e = run_assembly(h, a, P, &nl);
    h->launches += nl;
    if (e != cudaSuccess)
      return 1;
    PE_MARK(s, "setup: second assembly pass A0");
    // diagonal of A0 (with the halo part), Jacobi inverse for the u block
    SolverCtx c0 = make_solver_ctx(h);
    c0.values    = k.values0.p;
    if (build_preconditioner(c0, k.diag0.p, h->d_pinv.p) != 0)
      return 1;
    h->launches += 2;
    // GMRES works on the row-signed system with the xi rows scaled (pe_k3.cu); MINRES needs the symmetric K3
    const bool gmres = h->krylov == 1 && h->use_mg && h->mg && h->mg->ready && h->bp && h->bp->ready && h->cmg && h->cmg->ready;
    if (k3_fill_values(k, lambda_s, h->alpha, h->mu, gmres, s, &h->launches) != 0)
      return 1;
    PE_MARK(s, "setup: diagonal + K3 values");
    if (h->use_mg && h->mg && h->mg->ready && h->bp && h->bp->ready && h->cmg && h->cmg->ready)
      {
        // damped (block-)Jacobi sweeps alternate between two weights (a degree-2 Chebyshev-like smoother at the cost
        // of plain Jacobi).  Measured on B200 (scripts/sweep_smoother.py): vertex grids (0.6, 0.9), cell grids
        // (0.6, 1.2), face Jacobi 0.65: cfg 2 52 -> 44 iterations, cfg 4 (N = 512) 268 -> 172.
        h->mg->omega  = h->mg_omega;
        h->mg->omega2 = 1.5 * h->mg_omega;
        h->mg->nu     = h->mg_nu;
        h->cmg->nu    = h->mg_nu;
        // tuning (pe_set_tuning): weights of the alternating smoothing sweeps
        if (h->tune.mg_omega2 > 0.)
          h->mg->omega2 = h->tune.mg_omega2;
        if (h->tune.cmg_omega > 0.)
          h->cmg->omega = h->tune.cmg_omega;
        if (h->tune.cmg_omega2 > 0.)
          h->cmg->omega2 = h->tune.cmg_omega2;
        if (h->tune.face_omega > 0.)
          h->bp->omega_f = h->tune.face_omega;
        h->mg->fused_n = h->tune.mg_fused_n;
        // pressure block of the preconditioner: C + alpha^2/lambda_s W_int (C already holds c0 M)
        const double extra = h->alpha * h->alpha / lambda_s;
        if (mg_assemble(*h->mg, h->d_k_morton.p, h->k_scalar, lambda0, h->mu, h->last_dt, h->c0 + extra, s, &h->launches) != 0)
          return 1;
        if (cmg_assemble(*h->cmg, h->d_vx.p, h->d_vy.p, h->d_k_morton.p, h->k_scalar, h->last_dt, h->c0 + extra, s,
                         &h->launches) != 0)
          return 1;
        if (precond_update(*h->bp, k.values0.p, k.diag0.p, extra, s, &h->launches) != 0)
          return 1;
        if (gmres && k3_scale_p_rows(k, h->bp->dinv.p, h->mu, s, &h->launches) != 0)
          return 1;
        PE_MARK(s, "setup: multigrid levels + sub-block values");
      }
    return 0;
  }

  // once per mesh / pattern (not part of the timed solve): K3 structure, multigrid topology, sub-block lists
  int
  prepare_three_field(pe_handle *h)
  {
    cudaStream_t s        = h->stream;
    auto         to_local = [h](uint32_t g) { return global_to_local(h, g); };
    if (!h->k3)
      {
        h->k3        = new K3System;
        const int rc = k3_build_structure(*h->k3, h->mesh, h->dofs, h->part, h->d_rowptr.p, h->d_colidx.p, s);
        if (rc != 0)
          {
            delete h->k3;
            h->k3 = nullptr;
            return rc < 0 ? 1 : PE_ERR_PATTERN;
          }
      }
    if (h->use_mg && !h->mg)
      {
        h->mg = new MgHierarchy;
        if (mg_build_topology(*h->mg, h->mesh, h->dofs, to_local, s) < 0)
          return 1;
      }
    if (h->use_mg && h->mg->ready && !h->cmg)
      {
        h->cmg = new CellMg;
        if (cmg_build_topology(*h->cmg, h->mesh, s) < 0)
          return 1;
        // multi-GPU: the fine grid levels are distributed in row slabs (PE_MG_REPLICATED=1: every rank runs the
        // complete V-cycles, the first multi-GPU version, kept for A/B measurements)
        if (h->part.world > 1 && !h->tune.mg_replicated)
          {
            GridComm gc;
            gc.rank           = h->part.rank;
            gc.world          = h->part.world;
            gc.ctx            = h;
            gc.dist_min_n     = h->tune.mg_dist_min_n;
            gc.exchange_rows  = comm_exchange_rows;
            gc.allgather_rows = comm_allgather_rows;
            gc.reduce_rows    = comm_reduce_rows;
            if (!h->tune.mg_whole_slabs) // A/B: round-1 transfers of whole slabs (reduce / all-gather)
              {
                gc.boxes_to_slabs = comm_boxes_to_slabs;
                gc.slabs_to_boxes = comm_slabs_to_boxes;
              }
            mg_set_comm(*h->mg, gc);
            if (h->cmg->ready)
              cmg_set_comm(*h->cmg, gc);
          }
      }
    if (h->use_mg && h->mg->ready && !h->bp)
      {
        h->bp = new BlockPrecond;
        if (precond_build_structure(*h->bp, h->mesh, h->dofs, h->pat, h->part, to_local, s) < 0)
          return 1;
        if (h->part.world > 1)
          {
            if (comm_set_box(h, h->bp->box) != PE_OK)
              return 1;
            h->bp->halo       = comm_halo_exchange;
            h->bp->halo_p     = comm_halo_exchange_p;
            h->stream2        = h->bp->s2;
            if (comm_has_lanes(h) && !h->tune.one_lane)
              h->bp->set_lane = hook_set_lane; // pressure pipeline on its own stream + communicator
            h->bp->allreduce  = comm_allreduce;
            h->bp->allreduce2 = comm_allreduce2;
            h->bp->comm_ctx  = h;
          }
      }
    const size_t n3 = (size_t)h->k3->n_phys();
    for (int i = 0; i < 10; ++i)
      if (h->d_w[i].n < n3)
        {
          if (h->d_w[i].alloc(n3) != cudaSuccess)
            return 2;
          ++h->graph_epoch;
        }
    if (h->d_pinv.n < n3)
      {
        if (h->d_pinv.alloc(n3) != cudaSuccess)
          return 2;
        ++h->graph_epoch;
      }
    return 0;
  }

  SolverCtx
  make_solver_ctx3(pe_handle *h)
  {
    const K3System &k = *h->k3;
    SolverCtx c3      = make_solver_ctx(h);
    c3.L              = k.n_rows();
    c3.H              = k.H + k.Hx;
    c3.n_pos          = k.n_rows(); // signs are part of K3
    c3.rowptr         = k.rowptr.p;
    c3.colidx         = k.colidx.p;
    c3.values         = k.val.p;
    c3.rowblk         = h->spmv_kernel == 1 ? k.rowblk.p : nullptr;
    c3.n_rowblk       = h->spmv_kernel == 1 ? k.n_rowblk : 0;
    c3.gap.at         = k.L;
    c3.gap.len        = k.H;
    c3.precond_ctx    = h;
    if (h->part.world > 1)
      c3.halo = comm_halo_exchange3;
    c3.graphs    = h->graphs;
    c3.graph_key = h->graph_epoch;
    c3.calib     = &h->minres_calib;
    c3.anorm     = hook_anorm;
    c3.floor_rows = k.L;
    c3.ev_ring   = h->ev_ring;
    c3.gmres_graphs = h->gmres_graphs;
    return c3;
  }

  int
  solve_three_field(pe_handle *h, const double rel_tol, const int max_iterations, pe_solve_stats *stats)
  {
    cudaStream_t s = h->stream;
    {
      const int rc = prepare_three_field(h);
      if (rc != 0)
        return rc;
    }
    K3System    &k  = *h->k3;
    const size_t n3 = (size_t)k.n_phys();
    if (cudaEventRecord(h->ev[3], s) != cudaSuccess)
      return 1;
    // ---- once per assembled matrix ----------------------------------------------------------------------
    // (the reference factorises the matrix anew in every step, BR1new:2235-2238; here the operator and its
    //  preconditioner are a function of the matrix inputs and are kept while those do not change -- the matrix itself
    //  is still assembled every step and the true residual is taken on that fresh matrix)
    if (!k.values_valid)
      {
        const uint64_t key = matrix_fingerprint(h);
        if (h->tune.reuse_operator && h->op_key_valid && h->op_key == key)
          k.values_valid = true;
        else
          {
            h->op_key_valid = false;
            const int rc    = update_solver_operator(h);
            if (rc != 0)
              return rc;
            h->op_key       = key;
            h->op_key_valid = true;
          }
      }
    const bool block_pc = h->use_mg && h->mg && h->mg->ready && h->bp && h->bp->ready && h->cmg && h->cmg->ready;
    // ---- this right-hand side ----------------------------------------------------------------------------
    double *b3 = h->d_w[7].p, *x3 = h->d_w[8].p;
    if (k3_make_rhs(k, h->d_rhs.p, h->d_rowptr.p, h->d_colidx.p, h->d_values.p, b3, s) != 0)
      return 1;
    if (k.p_scaled)
      {
        k3_p_vec_scale(k, false, b3, b3, s); // the right-hand side of the row-scaled operator
        ++h->launches;
      }
    // initial guess: old_solution (BR1new:2241: the solution of the previous step), or -- when the solutions of the
    // last two / three steps are on the device -- their polynomial extrapolation to the new time, with its
    // consistent xi.  (A direct solver needs no guess; any guess leaves the result of the Krylov solve unchanged
    // up to the residual tolerance.)
    cudaMemsetAsync(x3, 0, sizeof(double) * n3, s);
    {
      const double t = h->last_t, dt = h->last_dt;
      // d_sol is the state at the end of the previous step
      if (h->hist_n < 1 || !(h->hist_t[0] < t))
        {
          h->hist_n    = 1;
          h->hist_t[0] = t - dt;
          if (h->recycle)
            h->recycle->k = 0; // a new trajectory (time does not advance): nothing is carried over from the old one
        }
      int n = std::min(h->hist_n, h->extrapolate + 1);
      // usable history: strictly decreasing times, and the new step not much longer than the previous ones
      if (n >= 2 && !(h->hist_t[1] < h->hist_t[0] && t - h->hist_t[0] <= 2.000001 * (h->hist_t[0] - h->hist_t[1])))
        n = 1;
      if (n >= 3 && !(h->hist_t[2] < h->hist_t[1]))
        n = 2;
      if (n == 1)
        cudaMemcpyAsync(x3, h->d_sol.p, sizeof(double) * (size_t)h->L, cudaMemcpyDeviceToDevice, s);
      else
        {
          double w[3] = {0., 0., 0.};
          for (int i = 0; i < n; ++i)
            {
              w[i] = 1.;
              for (int j = 0; j < n; ++j)
                if (j != i)
                  w[i] *= (t - h->hist_t[j]) / (h->hist_t[i] - h->hist_t[j]);
            }
          lincomb3(h->L, w[0], h->d_sol.p, w[1], h->d_hist[0].p, w[2], n >= 3 ? h->d_hist[1].p : nullptr, x3, s);
          ++h->launches;
        }
    }
    if (h->part.world > 1)
      comm_halo_exchange(h, x3);
    if (k3_init_xi(k, x3, s) != 0)
      return 1;
    h->launches += 4;
    SolverCtx c      = make_solver_ctx(h);
    SolverCtx c3     = make_solver_ctx3(h);
    c3.true_residual = hook_true_residual;
    const double *pinv3 = nullptr;
    if (block_pc)
      c3.precond_apply = hook_precond3;
    else
      {
        // no refinement hierarchy: inverse block diagonal of K3
        if (h->d_w[10].n < n3 && h->d_w[10].alloc(n3) != cudaSuccess)
          return 2;
        if (k3_make_pinv(k, h->mu, h->d_w[10].p, s) != 0)
          return 1;
        pinv3 = h->d_w[10].p;
      }
    // ||b|| of the reference system: what rel_tol refers to
    {
      double *d = h->d_scal.p + 40;
      cudaMemsetAsync(d, 0, sizeof(double), s);
      norm2(c, h->d_rhs.p, d);
      if (c.allreduce)
        c.allreduce(c.halo_ctx, d, 1);
      cudaMemcpyAsync(h->h_scal + 8, d, sizeof(double), cudaMemcpyDeviceToHost, s);
      if (cudaStreamSynchronize(s) != cudaSuccess)
        return 1;
      c3.bnorm2 = h->h_scal[8];
    }
    double *work[7] = {h->d_w[0].p, h->d_w[1].p, h->d_w[2].p, h->d_w[3].p, h->d_w[4].p, h->d_w[5].p, h->d_w[6].p};
    int     rc;
    PE_MARK(s, "setup: rhs, initial guess, ||b||");
    if (block_pc && h->krylov == 1)
      {
        // FGMRES(m) + block-triangular preconditioner on the row-signed system (rows >= n_u negated)
        const int    m      = std::max(2, std::min(h->gmres_m, kGmresMax));
        const size_t need   = (size_t)(2 * m + 1) * n3;
        if (h->d_gm.n < need)
          {
            if (h->d_gm.alloc(need) != cudaSuccess)
              return 2;
            ++h->graph_epoch;
          }
        if (h->d_gmres_state.n < gmres_state_bytes() && h->d_gmres_state.alloc(gmres_state_bytes()) != cudaSuccess)
          return 2;
        c3.n_pos         = k.n_u;
        c3.precond_apply = hook_precond3_tri;
        // projection space of the previous solves: valid as long as the operator is the same matrix
        if (h->recycle_k > 0)
          {
            const int    kr  = std::min(h->recycle_k, 8);
            const size_t nr  = (size_t)(2 * kr + 2) * n3;
            if (!h->recycle)
              h->recycle = new (std::nothrow) Recycle;
            if (h->recycle && (h->d_recycle.n < nr || h->recycle->k_max != kr || h->recycle->stride != (int64_t)n3))
              {
                if (h->d_recycle.alloc(nr) != cudaSuccess)
                  return 2;
                Recycle &R = *h->recycle;
                R          = Recycle{};
                R.k_max    = kr;
                R.stride   = (int64_t)n3;
                R.D        = h->d_recycle.p;
                R.Q        = h->d_recycle.p + (size_t)kr * n3;
                R.x0       = h->d_recycle.p + (size_t)(2 * kr) * n3;
                R.r0       = h->d_recycle.p + (size_t)(2 * kr + 1) * n3;
              }
            if (h->recycle)
              {
                const uint64_t key = matrix_fingerprint(h);
                if (h->recycle->matrix_key != key)
                  {
                    h->recycle->k          = 0;
                    h->recycle->matrix_key = key;
                  }
                c3.recycle = h->recycle;
              }
          }
        c3.graph_key     = h->graph_epoch * 64 + (uint64_t)m;
        rc = fgmres_solve(c3, b3, x3, h->d_gm.p, h->d_gm.p + (size_t)(m + 1) * n3, (int64_t)n3, h->d_w[0].p, m,
                          h->d_gmres_state.p, h->h_scal, rel_tol, max_iterations, stats);
        if (rc == 0 && !stats->converged && stats->residual_norm == stats->residual_norm && stats->iterations < max_iterations)
          {
            // Safety net for lambda >> mu on small meshes: the 2-norm GMRES on the three-field system weighs the xi
            // equation less than its effect on the reference residual (amplified by lambda / 2 mu), so its attainable
            // reference residual is ~ eps lambda / mu times a constant that MINRES (energy norm) does not pay.
            // Continue from the GMRES iterate with MINRES on the symmetric (unscaled) K3.
            const pe_solve_stats g = *stats;
            if (k3_fill_values(k, k.lambda_s, k.alpha, h->mu, false, s, &h->launches) != 0)
              return 1;
            if (k3_make_rhs(k, h->d_rhs.p, h->d_rowptr.p, h->d_colidx.p, h->d_values.p, b3, s) != 0)
              return 1;
            SolverCtx cm     = make_solver_ctx3(h);
            cm.true_residual = hook_true_residual;
            cm.precond_apply = hook_precond3;
            cm.bnorm2        = c3.bnorm2;
            cm.calib         = nullptr;
            rc = minres_solve(cm, b3, x3, nullptr, work, (MinresState *)h->d_scal.p, h->h_scal, rel_tol,
                              max_iterations - g.iterations, stats);
            stats->iterations += g.iterations;
            stats->spmv_count += g.spmv_count;
            stats->restarts += g.restarts + 1;
            k.values_valid  = false; // the next solve refills the GMRES (row-scaled) operator
            h->op_key_valid = false;
          }
      }
    else
      rc = minres_solve(c3, b3, x3, pinv3, work, (MinresState *)h->d_scal.p, h->h_scal, rel_tol, max_iterations, stats);
    if (rc == -2)
      return 3;
    // history of the solutions for the next initial guess: (d_sol, hist[0], hist[1]) <- (x, d_sol, hist[0])
    if (h->extrapolate > 0)
      {
        const size_t nb = sizeof(double) * (size_t)h->L;
        for (int k = 0; k < 2; ++k)
          if (h->d_hist[k].n < (size_t)h->L && h->d_hist[k].alloc((size_t)h->L) != cudaSuccess)
            return 2;
        if (h->hist_n >= 2)
          cudaMemcpyAsync(h->d_hist[1].p, h->d_hist[0].p, nb, cudaMemcpyDeviceToDevice, s);
        cudaMemcpyAsync(h->d_hist[0].p, h->d_sol.p, nb, cudaMemcpyDeviceToDevice, s);
        h->hist_t[2] = h->hist_t[1];
        h->hist_t[1] = h->hist_t[0];
        h->hist_t[0] = h->last_t;
        h->hist_n    = std::min(3, h->hist_n + 1);
      }
    cudaMemcpyAsync(h->d_sol.p, x3, sizeof(double) * (size_t)h->L, cudaMemcpyDeviceToDevice, s);
    return rc == 0 ? 0 : 1;
  }
} // namespace

extern "C"
{
  const char *
  pe_version(void)
  {
    return "poroelas_b200 0.1 (sm_100a)";
  }

  int
  pe_create(pe_handle **out, const pe_config *cfg)
  {
    if (!out || !cfg)
      return PE_ERR_ARG;
    *out = nullptr;
    // dim 3: CGQ1-WG on hexahedra, local matrices only (pe_make_unit_cube / pe_set_mesh3 / pe_local_matrices3)
    if (cfg->element < 0 || cfg->element > 2 || (cfg->dim != 2 && cfg->dim != 0 && !(cfg->dim == 3 && cfg->element == PE_CGQ1_WG)))
      return PE_ERR_ARG;
    const bool  host_only = (cfg->flags & PE_CFG_HOST_SETUP_ONLY) != 0;
    int         n_dev     = 0;
    cudaError_t e         = host_only ? cudaSuccess : cudaGetDeviceCount(&n_dev);
    if (!host_only && (e != cudaSuccess || n_dev == 0))
      return PE_ERR_NO_DEVICE;
    pe_handle *h = nullptr;
    try
      {
        h = new pe_handle;
      }
    catch (...)
      {
        return PE_ERR_ALLOC;
      }
    h->cfg = *cfg;
    if (h->cfg.world_size <= 0)
      {
        h->cfg.world_size = 1;
        h->cfg.rank       = 0;
      }
    if (h->cfg.rank < 0 || h->cfg.rank >= h->cfg.world_size)
      {
        delete h;
        return PE_ERR_ARG;
      }
    h->part.rank  = h->cfg.rank;
    h->part.world = h->cfg.world_size;
    if (host_only)
      {
        h->device = -1;
        *out      = h;
        return PE_OK;
      }
    if (cfg->device >= 0)
      {
        if (cfg->device >= n_dev || cudaSetDevice(cfg->device) != cudaSuccess)
          {
            delete h;
            return PE_ERR_NO_DEVICE;
          }
        h->device = cfg->device;
      }
    else
      cudaGetDevice(&h->device);
    if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess)
      {
        delete h;
        return PE_ERR_CUDA;
      }
    for (auto &ev : h->ev)
      cudaEventCreate(&ev);
    for (auto &ev : h->ev_ring)
      cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
    h->graphs       = new (std::nothrow) MinresGraphs;
    h->gmres_graphs = new (std::nothrow) GmresGraphs;
    cudaMallocHost((void **)&h->h_scal, 64 * sizeof(double));
    if (!h->graphs || !h->gmres_graphs || !h->h_scal)
      {
        pe_destroy(h);
        return PE_ERR_ALLOC;
      }
    *out = h;
    return PE_OK;
  }

  int
  pe_destroy(pe_handle *h)
  {
    if (!h)
      return PE_OK;
    if (h->device < 0)
      {
        delete h;
        return PE_OK;
      }
    cudaSetDevice(h->device);
    comm_destroy(h);
    if (h->stream)
      cudaStreamSynchronize(h->stream);
    delete h->graphs;
    h->graphs = nullptr;
    delete h->gmres_graphs;
    h->gmres_graphs = nullptr;
    delete h->recycle;
    h->recycle = nullptr;
    for (auto &ev : h->ev_ring)
      if (ev)
        cudaEventDestroy(ev);
    delete h->mg;
    h->mg = nullptr;
    delete h->bp;
    h->bp = nullptr;
    h->stream2 = nullptr;
    delete h->k3;
    h->k3 = nullptr;
    delete h->cmg;
    h->cmg = nullptr;
    for (auto &ev : h->ev)
      if (ev)
        cudaEventDestroy(ev);
    if (h->h_scal)
      cudaFreeHost(h->h_scal);
    if (h->h_stage)
      cudaFreeHost(h->h_stage);
    if (h->stream)
      cudaStreamDestroy(h->stream);
    delete h;
    return PE_OK;
  }

  const char *
  pe_last_error(const pe_handle *h)
  {
    return h ? h->err.c_str() : "null handle";
  }

#define PE_TRY try {
#define PE_CATCH                                                     \
  }                                                                  \
  catch (const std::bad_alloc &)                                     \
  {                                                                  \
    return fail(h, PE_ERR_ALLOC, "out of host memory");              \
  }                                                                  \
  catch (const std::exception &ex)                                   \
  {                                                                  \
    return fail(h, PE_ERR_ARG, ex.what());                           \
  }

  int
  pe_make_unit_cube(pe_handle *h, int n_refine, double distort, uint64_t seed)
  {
    if (!h || h->cfg.dim != 3 || n_refine < 0 || n_refine > 8)
      return fail(h, PE_ERR_ARG, "pe_make_unit_cube: needs a dim = 3 handle, 0 <= n_refine <= 8");
    PE_TRY
    // GridGenerator::hyper_cube(0,1) + refine_global(n) (CGQ1:886-887): active cells in 3-D Morton order (child c of a
    // cell at (c & 1, (c >> 1) & 1, c >> 2)), vertices lexicographic, cell vertex v at offset (v & 1, (v >> 1) & 1, v >> 2)
    auto        &m = h->mesh3;
    const int    N = 1 << n_refine, V = N + 1;
    const double hh = 1. / N;
    m.x.resize((size_t)V * V * V);
    m.y.resize(m.x.size());
    m.z.resize(m.x.size());
    for (int k = 0; k < V; ++k)
      for (int j = 0; j < V; ++j)
        for (int i = 0; i < V; ++i)
          {
            const size_t id = ((size_t)k * V + j) * V + i;
            double       x = i * hh, y = j * hh, z = k * hh;
            if (distort > 0. && i > 0 && i < N && j > 0 && j < N && k > 0 && k < N)
              {
                x += distort * hh * (2. * u01(seed, id, 0) - 1.);
                y += distort * hh * (2. * u01(seed, id, 1) - 1.);
                z += distort * hh * (2. * u01(seed, id, 2) - 1.);
              }
            m.x[id] = x;
            m.y[id] = y;
            m.z[id] = z;
          }
    m.n_cells = (int64_t)N * N * N;
    m.cell_vertices.resize((size_t)8 * m.n_cells);
    for (int64_t c = 0; c < m.n_cells; ++c)
      {
        uint32_t ix = 0, iy = 0, iz = 0;
        for (int b = 0; b < 10; ++b)
          {
            ix |= (uint32_t)((c >> (3 * b)) & 1) << b;
            iy |= (uint32_t)((c >> (3 * b + 1)) & 1) << b;
            iz |= (uint32_t)((c >> (3 * b + 2)) & 1) << b;
          }
        for (int v = 0; v < 8; ++v)
          m.cell_vertices[(size_t)8 * c + v] =
            (uint32_t)((((size_t)iz + (v >> 2)) * V + iy + ((v >> 1) & 1)) * V + ix + (v & 1));
      }
    // boundary data of the shipped program (sandwich problem): faces with centre z = 1 carry boundary id 1 = face-pressure
    // Dirichlet (CGQ1:972-981, 1663-1670) and the traction (0, 0, -1) (CGQ1:1563-1582); all other boundary faces keep
    // id 0 = displacement Dirichlet (CGQ1:1673-1679)
    m.face_kinds.assign((size_t)6 * m.n_cells, 0);
    for (int64_t c = 0; c < m.n_cells; ++c)
      {
        const uint32_t v0 = m.cell_vertices[(size_t)8 * c];
        const int      ix = (int)(v0 % V), iy = (int)((v0 / V) % V), iz = (int)(v0 / ((size_t)V * V));
        const int      lo[3] = {ix, iy, iz};
        for (int f = 0; f < 6; ++f)
          {
            const int d = f / 2;
            if ((f & 1) ? lo[d] + 1 == N : lo[d] == 0)
              m.face_kinds[(size_t)6 * c + f] = (f == 5) ? (PE_FACE_DIR_P | PE_FACE_NEU_U) : PE_FACE_DIR_U;
          }
      }
    m.traction[0] = m.traction[1] = 0.;
    m.traction[2] = -1.;
    h->dev3.ready = false;
    h->have_dofs = h->have_pattern = h->device_ready = h->matrix_valid = false;
    return PE_OK;
    PE_CATCH
  }

  int
  pe_set_boundary3(pe_handle *h, const uint8_t *face_kinds, const double *traction)
  {
    if (!h || h->cfg.dim != 3 || h->mesh3.n_cells <= 0 || !face_kinds)
      return fail(h, PE_ERR_ARG, "pe_set_boundary3: needs a dim = 3 handle with a mesh and face_kinds[n_cells][6]");
    PE_TRY
    auto &m = h->mesh3;
    for (size_t e = 0; e < (size_t)6 * m.n_cells; ++e)
      if (face_kinds[e] & ~7)
        return fail(h, PE_ERR_ARG, "pe_set_boundary3: unknown face kind bits");
    m.face_kinds.assign(face_kinds, face_kinds + (size_t)6 * m.n_cells);
    for (int k = 0; k < 3; ++k)
      m.traction[k] = traction ? traction[k] : 0.;
    h->dev3.ready   = false;
    h->matrix_valid = false;
    ++h->data_epoch;
    return PE_OK;
    PE_CATCH
  }

  int
  pe_set_mesh3(pe_handle *h, int64_t n_vertices, const double *x, const double *y, const double *z, int64_t n_cells,
               const uint32_t *cell_vertices)
  {
    if (!h || h->cfg.dim != 3 || !x || !y || !z || !cell_vertices || n_vertices <= 0 || n_cells <= 0)
      return fail(h, PE_ERR_ARG, "pe_set_mesh3: bad argument");
    PE_TRY
    auto &m = h->mesh3;
    m.x.assign(x, x + n_vertices);
    m.y.assign(y, y + n_vertices);
    m.z.assign(z, z + n_vertices);
    m.cell_vertices.assign(cell_vertices, cell_vertices + 8 * (size_t)n_cells);
    for (const uint32_t v : m.cell_vertices)
      if ((int64_t)v >= n_vertices)
        return fail(h, PE_ERR_ARG, "pe_set_mesh3: vertex index out of range");
    m.n_cells = n_cells;
    h->mesh3.face_kinds.clear(); // a user mesh carries no boundary data until pe_set_boundary3
    h->dev3.ready = false;
    h->have_dofs = h->have_pattern = h->device_ready = h->matrix_valid = false;
    return PE_OK;
    PE_CATCH
  }

  int
  pe_get_mesh3(const pe_handle *h, int64_t *n_vertices, int64_t *n_cells, double *x, double *y, double *z, uint32_t *cell_vertices)
  {
    if (!h || h->cfg.dim != 3)
      return PE_ERR_ARG;
    const auto &m = h->mesh3;
    if (n_vertices)
      *n_vertices = (int64_t)m.x.size();
    if (n_cells)
      *n_cells = m.n_cells;
    if (x)
      std::copy(m.x.begin(), m.x.end(), x);
    if (y)
      std::copy(m.y.begin(), m.y.end(), y);
    if (z)
      std::copy(m.z.begin(), m.z.end(), z);
    if (cell_vertices)
      std::copy(m.cell_vertices.begin(), m.cell_vertices.end(), cell_vertices);
    return PE_OK;
  }

  int
  pe_local_matrices3(pe_handle *h, int64_t first_cell, int64_t n, double time_step, const double *cell_old, double *out,
                     double *rhs_out)
  {
    if (!h || h->cfg.dim != 3 || !out || n <= 0 || first_cell < 0 || first_cell + n > h->mesh3.n_cells)
      return fail(h, PE_ERR_ARG, "pe_local_matrices3: bad argument (dim = 3 handle with a mesh)");
    if (h->device < 0)
      return fail(h, PE_ERR_NO_DEVICE, "no hot path without a GPU");
    PE_DEVICE(h);
    PE_TRY
    cudaStream_t     s  = h->stream;
    const auto      &m  = h->mesh3;
    const int64_t    nc = m.n_cells;
    DevBuf<double>   x, y, z, kc, o, ro, old;
    DevBuf<uint32_t> cv;
    PE_CUDA(upload(x, m.x.data(), m.x.size(), s));
    PE_CUDA(upload(y, m.y.data(), m.y.size(), s));
    PE_CUDA(upload(z, m.z.data(), m.z.size(), s));
    {
      std::vector<uint32_t> cvh((size_t)8 * nc);
      for (int64_t c = 0; c < nc; ++c)
        for (int v = 0; v < 8; ++v)
          cvh[(size_t)v * nc + c] = m.cell_vertices[(size_t)8 * c + v];
      PE_CUDA(upload(cv, cvh.data(), cvh.size(), s));
      PE_CUDA(cudaStreamSynchronize(s));
    }
    const double *kp = nullptr;
    if (!h->k_cell.empty())
      {
        if ((int64_t)h->k_cell.size() != nc)
          return fail(h, PE_ERR_ARG, "pe_local_matrices3: k_cell was set for another mesh");
        PE_CUDA(upload(kc, h->k_cell.data(), h->k_cell.size(), s));
        kp = kc.p;
      }
    if (cell_old)
      {
        std::vector<double> oh((size_t)31 * n);
        for (int64_t c = 0; c < n; ++c)
          for (int i = 0; i < 31; ++i)
            oh[(size_t)i * n + c] = cell_old[(size_t)c * 31 + i];
        PE_CUDA(upload(old, oh.data(), oh.size(), s));
        PE_CUDA(cudaStreamSynchronize(s));
      }
    PE_CUDA(o.alloc((size_t)961 * n));
    if (rhs_out)
      PE_CUDA(ro.alloc((size_t)31 * n));
    PE_CUDA(launch_local_matrices_3d(nc, first_cell, n, x.p, y.p, z.p, cv.p, kp, h->k_scalar, h->lambda, h->mu, h->alpha, h->c0,
                                     time_step, cell_old ? old.p : nullptr, o.p, rhs_out ? ro.p : nullptr, s));
    ++h->launches;
    std::vector<double> tmp((size_t)961 * n);
    PE_CUDA(cudaMemcpyAsync(tmp.data(), o.p, sizeof(double) * tmp.size(), cudaMemcpyDeviceToHost, s));
    PE_CUDA(cudaStreamSynchronize(s));
    for (int64_t c = 0; c < n; ++c)
      for (size_t e = 0; e < 961; ++e)
        out[(size_t)c * 961 + e] = tmp[e * (size_t)n + c];
    if (rhs_out)
      {
        std::vector<double> tr((size_t)31 * n);
        PE_CUDA(cudaMemcpyAsync(tr.data(), ro.p, sizeof(double) * tr.size(), cudaMemcpyDeviceToHost, s));
        PE_CUDA(cudaStreamSynchronize(s));
        for (int64_t c = 0; c < n; ++c)
          for (int i = 0; i < 31; ++i)
            rhs_out[(size_t)c * 31 + i] = tr[(size_t)i * n + c];
      }
    return PE_OK;
    PE_CATCH
  }

  int
  pe_make_unit_square(pe_handle *h, int n_refine, double distort, uint64_t seed, int boundary_kinds)
  {
    if (!h)
      return PE_ERR_ARG;
    PE_TRY
    make_unit_square(h->mesh, n_refine, distort, seed, boundary_kinds);
    h->have_mesh = true;
    ++h->dofs_epoch;
    h->have_dofs = h->have_pattern = h->device_ready = false;
    return PE_OK;
    PE_CATCH
  }

  int
  pe_set_mesh(pe_handle *h, int64_t n_vertices, const double *x, const double *y, int64_t xy_stride,
              int64_t n_cells, const uint32_t *cell_vertices)
  {
    if (!h || !x || !y || !cell_vertices || n_vertices <= 0 || n_cells <= 0 || xy_stride < 1)
      return fail(h, PE_ERR_ARG, "pe_set_mesh: bad argument");
    PE_TRY
    HostMesh &m  = h->mesh;
    m.n_refine   = -1;
    m.n_vertices = n_vertices;
    m.n_cells    = n_cells;
    m.x.resize((size_t)n_vertices);
    m.y.resize((size_t)n_vertices);
    for (int64_t i = 0; i < n_vertices; ++i)
      {
        m.x[i] = x[i * xy_stride];
        m.y[i] = y[i * xy_stride];
      }
    m.cell_vertices.assign(cell_vertices, cell_vertices + 4 * (size_t)n_cells);
    for (const uint32_t v : m.cell_vertices)
      if ((int64_t)v >= n_vertices)
        return fail(h, PE_ERR_ARG, "pe_set_mesh: vertex index out of range");
    m.face_kinds.assign(4 * (size_t)n_cells, 0);
    h->have_mesh = true;
    ++h->dofs_epoch;
    h->have_dofs = h->have_pattern = h->device_ready = false;
    return PE_OK;
    PE_CATCH
  }

  int
  pe_distribute_dofs(pe_handle *h)
  {
    if (h && h->cfg.dim == 3)
      {
        // hexahedra: numbering + (below) sparsity exist on the host; scatter and solve in 3-D are not built yet
        if (h->mesh3.n_cells <= 0)
          return fail(h, PE_ERR_ARG, "pe_distribute_dofs: set the mesh first");
        PE_TRY
        distribute_dofs3((int64_t)h->mesh3.x.size(), h->mesh3.n_cells, h->mesh3.cell_vertices.data(), h->dofs, h->part);
        h->mesh.n_cells    = h->mesh3.n_cells; // the generic pattern code walks cell_dofs only
        h->mesh.n_vertices = (int64_t)h->mesh3.x.size();
        h->mesh.n_refine   = -1;
        h->have_dofs       = true;
        h->have_pattern = h->device_ready = h->matrix_valid = false;
        h->dev3.ready   = false;
        ++h->dofs_epoch;
        return PE_OK;
        PE_CATCH
      }
    if (!h || !h->have_mesh)
      return fail(h, PE_ERR_ARG, "pe_distribute_dofs: set the mesh first");
    PE_TRY
    distribute_dofs(h->mesh, h->cfg.element, h->dofs, h->part);
    h->have_dofs    = true;
    ++h->dofs_epoch;
    h->have_pattern = h->device_ready = false;
    return PE_OK;
    PE_CATCH
  }

  int
  pe_set_dofs(pe_handle *h, int64_t n_u, int64_t n_pint, int64_t n_pface, const uint32_t *cell_dofs)
  {
    if (!h || !h->have_mesh || !cell_dofs)
      return fail(h, PE_ERR_ARG, "pe_set_dofs: bad argument");
    if (h->part.world != 1)
      return fail(h, PE_ERR_ARG, "pe_set_dofs: user numbering is supported on one GPU only");
    PE_TRY
    HostDofs &d = h->dofs;
    d.element   = h->cfg.element;
    d.n_loc     = element_n_loc(h->cfg.element);
    d.n_u       = n_u;
    d.n_pint    = n_pint;
    d.n_pface   = n_pface;
    d.cell_dofs.assign(cell_dofs, cell_dofs + (size_t)h->mesh.n_cells * d.n_loc);
    for (const uint32_t g : d.cell_dofs)
      if ((int64_t)g >= d.n())
        return fail(h, PE_ERR_ARG, "pe_set_dofs: dof index out of range");
    Partition &pt = h->part;
    pt.cell_begin = 0;
    pt.cell_end   = h->mesh.n_cells;
    pt.n_owned    = d.n();
    const int64_t st[4] = {0, n_u, n_u + n_pint, d.n()};
    for (int b = 0; b < 3; ++b)
      {
        pt.row_begin[b] = st[b];
        pt.row_end[b]   = st[b + 1];
        pt.local_off[b] = st[b];
      }
    pt.all_row_begin.assign(pt.row_begin, pt.row_begin + 3);
    pt.all_row_end.assign(pt.row_end, pt.row_end + 3);
    h->have_dofs    = true;
    ++h->dofs_epoch;
    h->have_pattern = h->device_ready = false;
    return PE_OK;
    PE_CATCH
  }

  int
  pe_make_sparsity_pattern(pe_handle *h)
  {
    if (!h || !h->have_dofs)
      return fail(h, PE_ERR_ARG, "pe_make_sparsity_pattern: distribute dofs first");
    PE_TRY
    h->part.scatter_mode = h->pending_scatter_mode;
    make_sparsity_pattern(h->mesh, h->dofs, h->part, h->pat);
    h->have_pattern = true;
    h->device_ready = false;
    h->dev3.ready   = false;
    return PE_OK;
    PE_CATCH
  }

  int
  pe_set_pattern(pe_handle *h, const int64_t *rowptr, const int32_t *colidx)
  {
    if (!h || !h->have_dofs || !rowptr || !colidx)
      return fail(h, PE_ERR_ARG, "pe_set_pattern: bad argument");
    if (h->part.world != 1)
      return fail(h, PE_ERR_ARG, "pe_set_pattern: supported on one GPU only");
    PE_TRY
    const int64_t n = h->dofs.n();
    if (rowptr[0] != 0)
      return fail(h, PE_ERR_ARG, "pe_set_pattern: rowptr[0] != 0");
    for (int64_t r = 0; r < n; ++r)
      {
        if (rowptr[r + 1] < rowptr[r])
          return fail(h, PE_ERR_ARG, "pe_set_pattern: rowptr not monotone");
        for (int64_t k = rowptr[r] + 1; k < rowptr[r + 1]; ++k)
          if (colidx[k] <= colidx[k - 1])
            return fail(h, PE_ERR_ARG, "pe_set_pattern: columns of a row must be strictly ascending");
      }
    h->pat.rowptr.assign(rowptr, rowptr + n + 1);
    h->pat.colidx.assign(colidx, colidx + rowptr[n]);
    Partition &pt = h->part;
    pt.scatter_mode = h->pending_scatter_mode;
    build_row_adjacency(h->mesh, h->dofs, pt);
    color_asm_cells(h->mesh, h->dofs, pt);
    pt.ghost_cols.clear();
    h->have_pattern = true;
    h->device_ready = false;
    return PE_OK;
    PE_CATCH
  }

  int
  pe_get_sizes(const pe_handle *h, pe_sizes *s)
  {
    if (!h || !s)
      return PE_ERR_ARG;
    std::memset(s, 0, sizeof(*s));
    s->n_cells    = h->cfg.dim == 3 ? h->mesh3.n_cells : h->mesh.n_cells;
    s->n_vertices = h->cfg.dim == 3 ? (int64_t)h->mesh3.x.size() : h->mesh.n_vertices;
    s->n_loc      = h->cfg.dim == 3 ? 31 : element_n_loc(h->cfg.element);
    if (h->have_dofs)
      {
        s->n_u          = h->dofs.n_u;
        s->n_pint       = h->dofs.n_pint;
        s->n_pface      = h->dofs.n_pface;
        s->n_dofs       = h->dofs.n();
        s->cell_begin   = h->part.cell_begin;
        s->cell_end     = h->part.cell_end;
        s->n_owned_rows = h->part.n_owned;
      }
    if (h->have_pattern)
      {
        s->nnz_owned     = (int64_t)h->pat.colidx.size();
        s->nnz           = s->nnz_owned; // global count only known on one GPU
        s->n_ghost_cells = h->part.n_ghost_cells;
      }
    if (h->k3 && h->k3->built)
      {
        s->solver_rows = h->k3->n_rows();
        s->solver_nnz  = h->k3->nnz;
      }
    return PE_OK;
  }

  int
  pe_get_owned_ranges(const pe_handle *h, int64_t begin[3], int64_t end[3])
  {
    if (!h || !h->have_dofs || !begin || !end)
      return PE_ERR_ARG;
    for (int b = 0; b < 3; ++b)
      {
        begin[b] = h->part.row_begin[b];
        end[b]   = h->part.row_end[b];
      }
    return PE_OK;
  }

  int
  pe_get_mesh(const pe_handle *h, double *x, double *y, uint32_t *cell_vertices, uint8_t *face_kinds)
  {
    if (!h || !h->have_mesh)
      return PE_ERR_ARG;
    if (x)
      std::copy(h->mesh.x.begin(), h->mesh.x.end(), x);
    if (y)
      std::copy(h->mesh.y.begin(), h->mesh.y.end(), y);
    if (cell_vertices)
      std::copy(h->mesh.cell_vertices.begin(), h->mesh.cell_vertices.end(), cell_vertices);
    if (face_kinds)
      std::copy(h->mesh.face_kinds.begin(), h->mesh.face_kinds.end(), face_kinds);
    return PE_OK;
  }

  int
  pe_get_dofs(const pe_handle *h, uint32_t *cell_dofs)
  {
    if (!h || !h->have_dofs || !cell_dofs)
      return PE_ERR_ARG;
    std::copy(h->dofs.cell_dofs.begin(), h->dofs.cell_dofs.end(), cell_dofs);
    return PE_OK;
  }

  int
  pe_get_pattern(const pe_handle *h, int64_t *rowptr, int32_t *colidx)
  {
    if (!h || !h->have_pattern)
      return PE_ERR_ARG;
    if (rowptr)
      std::copy(h->pat.rowptr.begin(), h->pat.rowptr.end(), rowptr);
    if (colidx)
      std::copy(h->pat.colidx.begin(), h->pat.colidx.end(), colidx);
    return PE_OK;
  }

  int
  pe_get_dealii_block_order(const pe_handle *hc, int64_t *perm)
  {
    pe_handle *h = const_cast<pe_handle *>(hc);
    if (!h || !h->have_pattern || !perm || h->part.world != 1)
      return fail(h, PE_ERR_ARG, "pe_get_dealii_block_order: needs a pattern on one GPU");
    PE_TRY
    std::vector<int64_t> p;
    dealii_block_order(h->dofs, h->pat, p);
    std::copy(p.begin(), p.end(), perm);
    return PE_OK;
    PE_CATCH
  }

  int
  pe_set_material(pe_handle *h, double lambda, double mu, double alpha, double c0, double k_scalar,
                  const double *k_cell)
  {
    if (!h)
      return PE_ERR_ARG;
    PE_DEVICE(h);
    PE_TRY
    h->lambda      = lambda;
    h->mu          = mu;
    h->alpha       = alpha;
    h->c0          = c0;
    h->k_scalar    = k_scalar;
    h->k_lognormal = false;
    ++h->data_epoch;
    h->matrix_valid = false; // the assembled matrix (and the solver operator built from it) belongs to the old data
    if (k_cell)
      {
        if (h->cfg.dim == 3)
          {
            if (h->mesh3.n_cells <= 0)
              return fail(h, PE_ERR_ARG, "pe_set_material: per-cell K needs the mesh first");
            h->k_cell.assign(k_cell, k_cell + h->mesh3.n_cells);
            return PE_OK;
          }
        if (!h->have_mesh)
          return fail(h, PE_ERR_ARG, "pe_set_material: per-cell K needs the mesh first");
        h->k_cell.assign(k_cell, k_cell + h->mesh.n_cells);
      }
    else
      h->k_cell.clear();
    if (h->device_ready)
      {
        // refresh the device copy of K only
        if (!h->k_cell.empty())
          {
            std::vector<double> k((size_t)h->n_asm);
            for (int64_t a = 0; a < h->n_asm; ++a)
              k[a] = h->k_cell[h->part.asm_cells[a]];
            PE_CUDA(upload(h->d_k, k.data(), k.size(), h->stream));
            PE_CUDA(upload(h->d_k_morton, h->k_cell.data(), h->k_cell.size(), h->stream));
            PE_CUDA(cudaStreamSynchronize(h->stream));
          }
        else
          {
            h->d_k.release();
            h->d_k_morton.release();
          }
      }
    return PE_OK;
    PE_CATCH
  }

  int
  pe_set_lognormal_k(pe_handle *h, double sigma, uint64_t seed)
  {
    if (!h)
      return PE_ERR_ARG;
    ++h->data_epoch;
    h->k_lognormal  = true;
    h->k_sigma      = sigma;
    h->k_seed       = seed;
    h->device_ready = false;
    return PE_OK;
  }

  int
  pe_set_case(pe_handle *h, int case_id, const double *params)
  {
    if (!h || case_id < 0 || case_id > 2)
      return fail(h, PE_ERR_ARG, "pe_set_case: unknown case");
    h->case_id = case_id;
    ++h->data_epoch;
    const double def[5] = {1., 1., 1., 0., 1.};
    std::copy(params ? params : def, (params ? params : def) + 5, h->case_params);
    return PE_OK;
  }

  int
  pe_set_boundary(pe_handle *h, int64_t n_bfaces, const uint32_t *cell, const uint8_t *face, const uint8_t *kinds,
                  const double traction[2])
  {
    if (!h || !h->have_mesh || (n_bfaces > 0 && (!cell || !face || !kinds)))
      return fail(h, PE_ERR_ARG, "pe_set_boundary: bad argument");
    std::fill(h->mesh.face_kinds.begin(), h->mesh.face_kinds.end(), 0);
    for (int64_t k = 0; k < n_bfaces; ++k)
      {
        if ((int64_t)cell[k] >= h->mesh.n_cells || face[k] > 3)
          return fail(h, PE_ERR_ARG, "pe_set_boundary: cell/face out of range");
        h->mesh.face_kinds[4 * (size_t)cell[k] + face[k]] = kinds[k] & 7;
      }
    ++h->data_epoch;
    h->mesh.traction[0] = traction ? traction[0] : 0.;
    h->mesh.traction[1] = traction ? traction[1] : 0.;
    h->device_ready     = false;
    return PE_OK;
  }

  // assemble_system on hexahedra (CGQ1:1165-1706 with dim = 3; one rank): dense blocks of k_local_matrices_3d in chunks,
  // then constraints + traction + scatter with atomics into the pre-zeroed CSR values (k_scatter_3d).  Boundary values are
  // homogeneous (BoundaryValues returns 0, CGQ1:697-725).
  static int
  assemble3(pe_handle *h, const double time_step, const double *old_solution)
  {
    if (!h->have_dofs || !h->have_pattern)
      return fail(h, PE_ERR_ARG, "pe_assemble_system (dim = 3): distribute dofs and build the sparsity pattern first");
    if (h->device < 0)
      return fail(h, PE_ERR_NO_DEVICE, "handle was created with PE_CFG_HOST_SETUP_ONLY: no hot path without a GPU");
    cudaStream_t  s  = h->stream;
    const auto   &m  = h->mesh3;
    const int64_t nc = m.n_cells, n = h->dofs.n();
    pe_handle::Dev3 &d = h->dev3;
    if (!d.ready)
      {
        PE_CUDA(upload(d.x, m.x.data(), m.x.size(), s));
        PE_CUDA(upload(d.y, m.y.data(), m.y.size(), s));
        PE_CUDA(upload(d.z, m.z.data(), m.z.size(), s));
        std::vector<uint32_t> cvh((size_t)8 * nc);
        for (int64_t c = 0; c < nc; ++c)
          for (int v = 0; v < 8; ++v)
            cvh[(size_t)v * nc + c] = m.cell_vertices[(size_t)8 * c + v];
        PE_CUDA(upload(d.cv, cvh.data(), cvh.size(), s));
        std::vector<int32_t> cdh((size_t)31 * nc);
        for (int64_t c = 0; c < nc; ++c)
          for (int i = 0; i < 31; ++i)
            cdh[(size_t)i * nc + c] = (int32_t)h->dofs.cell_dofs[(size_t)c * 31 + i];
        PE_CUDA(upload(d.cdofs, cdh.data(), cdh.size(), s));
        // constraints of the WHOLE mesh (interpolate_boundary_values runs on the dof_handler, CGQ1:1662-1679): a cell sees
        // every constrained dof it touches, also through a vertex or an edge only
        static const int     face_v[6][4] = {{0, 2, 4, 6}, {1, 3, 5, 7}, {0, 1, 4, 5}, {2, 3, 6, 7}, {0, 1, 2, 3}, {4, 5, 6, 7}};
        std::vector<uint8_t> vflag(m.x.size(), 0);
        const bool           have_bnd = !m.face_kinds.empty();
        if (have_bnd)
          for (int64_t c = 0; c < nc; ++c)
            for (int f = 0; f < 6; ++f)
              if (m.face_kinds[(size_t)6 * c + f] & PE_FACE_DIR_U)
                for (int k = 0; k < 4; ++k)
                  vflag[m.cell_vertices[(size_t)8 * c + face_v[f][k]]] = 1;
        std::vector<uint32_t> mask((size_t)nc, 0u);
        std::vector<uint8_t>  neu((size_t)nc, 0);
        for (int64_t c = 0; c < nc; ++c)
          {
            uint32_t b = 0;
            for (int v = 0; v < 8; ++v)
              if (vflag[m.cell_vertices[(size_t)8 * c + v]])
                b |= 7u << (3 * v);
            if (have_bnd)
              for (int f = 0; f < 6; ++f)
                {
                  const uint8_t kind = m.face_kinds[(size_t)6 * c + f];
                  if (kind & PE_FACE_DIR_P)
                    b |= 1u << (24 + f);
                  if (kind & PE_FACE_NEU_U)
                    neu[(size_t)c] |= (uint8_t)(1u << f);
                }
            mask[(size_t)c] = b;
          }
        PE_CUDA(upload(d.con_mask, mask.data(), mask.size(), s));
        PE_CUDA(upload(d.neu_mask, neu.data(), neu.size(), s));
        PE_CUDA(upload(h->d_rowptr, h->pat.rowptr.data(), h->pat.rowptr.size(), s));
        PE_CUDA(upload(h->d_colidx, h->pat.colidx.data(), h->pat.colidx.size(), s));
        h->L   = n;
        h->H   = 0;
        h->nnz = (int64_t)h->pat.colidx.size();
        PE_CUDA(h->d_values.alloc((size_t)h->nnz));
        PE_CUDA(h->d_rhs.alloc((size_t)n));
        PE_CUDA(d.old.alloc((size_t)n));
        PE_CUDA(cudaMemsetAsync(d.old.p, 0, sizeof(double) * (size_t)n, s)); // old_solution = 0 at the start of run()
        PE_CUDA(d.err.alloc(1));
        d.chunk = std::min<int64_t>(nc, 16384); // 16384 x 961 x 8 B = 126 MB of dense blocks per chunk
        PE_CUDA(d.A.alloc((size_t)961 * d.chunk));
        PE_CUDA(d.b.alloc((size_t)31 * d.chunk));
        PE_CUDA(d.cell_old.alloc((size_t)31 * d.chunk));
        PE_CUDA(cudaStreamSynchronize(s));
        d.ready = true;
      }
    // permeability: per cell (pe_set_material) or the scalar
    const double *kp = nullptr;
    if (!h->k_cell.empty())
      {
        if ((int64_t)h->k_cell.size() != nc)
          return fail(h, PE_ERR_ARG, "pe_assemble_system (dim = 3): k_cell was set for another mesh");
        PE_CUDA(upload(d.kc, h->k_cell.data(), h->k_cell.size(), s));
        kp = d.kc.p;
      }
    PE_CUDA(cudaEventRecord(h->ev[0], s));
    // old_solution: a host vector of all dofs, or NULL = the state resident on the device (zero after setup, the
    // solution of the previous pe_solve_time_step afterwards: old_solution = solution, CGQ1:3011)
    if (old_solution)
      PE_CUDA(cudaMemcpyAsync(d.old.p, old_solution, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, s));
    PE_CUDA(cudaMemsetAsync(h->d_values.p, 0, sizeof(double) * (size_t)h->nnz, s));
    PE_CUDA(cudaMemsetAsync(h->d_rhs.p, 0, sizeof(double) * (size_t)n, s));
    PE_CUDA(cudaMemsetAsync(d.err.p, 0, sizeof(int), s));
    PE_CUDA(cudaEventRecord(h->ev[1], s));
    for (int64_t first = 0; first < nc; first += d.chunk)
      {
        const int64_t cn = std::min(d.chunk, nc - first);
        PE_CUDA(launch_gather_old_3d(nc, first, cn, d.cdofs.p, d.old.p, d.cell_old.p, s));
        PE_CUDA(launch_local_matrices_3d(nc, first, cn, d.x.p, d.y.p, d.z.p, d.cv.p, kp, h->k_scalar, h->lambda, h->mu, h->alpha,
                                         h->c0, time_step, d.cell_old.p, d.A.p, d.b.p, s));
        PE_CUDA(launch_scatter_3d(nc, first, cn, d.x.p, d.y.p, d.z.p, d.cv.p, d.cdofs.p, d.con_mask.p, d.neu_mask.p, m.traction,
                                  d.A.p, d.b.p, h->d_rowptr.p, h->d_colidx.p, h->d_values.p, h->d_rhs.p, d.err.p, s));
        h->launches += 3;
      }
    PE_CUDA(cudaEventRecord(h->ev[2], s));
    int bad = 0;
    PE_CUDA(cudaMemcpyAsync(&bad, d.err.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    PE_CUDA(cudaEventSynchronize(h->ev[2]));
    PE_CUDA(cudaStreamSynchronize(s));
    if (bad)
      return fail(h, PE_ERR_PATTERN, "pe_assemble_system (dim = 3): a local entry is not in the sparsity pattern");
    float ms = 0.f;
    cudaEventElapsedTime(&ms, h->ev[0], h->ev[2]);
    h->assemble_ms = ms;
    cudaEventElapsedTime(&ms, h->ev[1], h->ev[2]);
    h->assemble_kernel_ms = ms;
    h->last_dt            = time_step;
    h->matrix_valid       = true; // pe_get_matrix / pe_get_rhs read d_values / d_rhs (device_ready stays false: no 2-D tables)
    return PE_OK;
  }


  // solve_time_step in 3-D (CGQ1:1712-1722: UMFPACK + constraints.distribute).  First version: MINRES on the row-signed
  // two-field system (u rows kept, pressure rows negated: symmetric) with the diagonal preconditioner of the first 2-D
  // solver -- |a_ii| on u and face-pressure rows, |a_ii| + sum_j b_ij^2 / |a_jj| (diagonal of the Schur complement) on
  // the interior pressures --, warm start from old_solution, stopping rule ||b - A x|| <= rel_tol ||b|| on the assembled
  // matrix.  No multigrid in 3-D yet: fine for the shipped 8^3 mesh, not a production solver.  The constrained dofs come
  // out as their boundary value 0 (their rows are d x_i = 0), which is what constraints.distribute() enforces.
  static int
  solve3(pe_handle *h, const double rel_tol, const int max_iterations, double *solution, pe_solve_stats *stats)
  {
    pe_handle::Dev3 &d = h->dev3;
    cudaStream_t     s = h->stream;
    const size_t     n = (size_t)h->L;
    for (int k = 0; k < 10; ++k)
      if (h->d_w[k].n < n)
        PE_CUDA(h->d_w[k].alloc(n));
    if (h->d_scal.n < 64)
      PE_CUDA(h->d_scal.alloc(64));
    SolverCtx c;
    c.L              = h->L;
    c.H              = 0;
    c.n_pos          = h->dofs.n_u;
    c.pint_begin     = h->dofs.n_u;
    c.pint_end       = h->dofs.n_u + h->dofs.n_pint;
    c.rowptr         = h->d_rowptr.p;
    c.colidx         = h->d_colidx.p;
    c.values         = h->d_values.p;
    c.stream         = s;
    c.use_graphs     = h->use_graphs != 0;
    c.launch_counter = &h->launches;
    c.ev_ring        = h->ev_ring;
    PE_CUDA(cudaEventRecord(h->ev[3], s));
    double *diag = h->d_w[8].p, *pinv = h->d_w[9].p, *x = h->d_w[7].p;
    if (build_preconditioner(c, diag, pinv) != 0)
      return cuda_fail(h, cudaGetLastError(), "3-D diagonal preconditioner");
    PE_CUDA(cudaMemcpyAsync(x, d.old.p, sizeof(double) * n, cudaMemcpyDeviceToDevice, s));
    double   *work[7] = {h->d_w[0].p, h->d_w[1].p, h->d_w[2].p, h->d_w[3].p, h->d_w[4].p, h->d_w[5].p, h->d_w[6].p};
    const int rc      = minres_solve(c, h->d_rhs.p, x, pinv, work, (MinresState *)h->d_scal.p, h->h_scal, rel_tol, max_iterations, stats);
    if (rc == -2)
      return fail(h, PE_ERR_CUDA, "CUDA graph capture of the MINRES iteration failed (pe_set_solver_options bit 6 runs without graphs)");
    PE_CUDA(cudaEventRecord(h->ev[4], s));
    PE_CUDA(cudaEventSynchronize(h->ev[4]));
    if (rc != 0)
      return cuda_fail(h, cudaGetLastError(), "krylov solve (3-D)");
    float ms = 0.f;
    cudaEventElapsedTime(&ms, h->ev[3], h->ev[4]);
    h->solve_ms    = ms;
    stats->seconds = ms * 1e-3;
    PE_CUDA(cudaMemcpyAsync(d.old.p, x, sizeof(double) * n, cudaMemcpyDeviceToDevice, s)); // old_solution = solution
    if (solution)
      PE_CUDA(cudaMemcpyAsync(solution, x, sizeof(double) * n, cudaMemcpyDeviceToHost, s));
    PE_CUDA(cudaStreamSynchronize(s));
    if (!stats->converged)
      return fail(h, PE_ERR_NOT_CONVERGED, "Krylov solver did not reach the requested residual");
    return PE_OK;
  }

  int
  pe_assemble_system(pe_handle *h, double t, double time_step, const double *old_solution)
  {
    NvtxRange nvtx("pe_assemble_system");
    if (!h)
      return PE_ERR_ARG;
    PE_DEVICE(h);
    PE_TRY
    if (h->cfg.dim == 3)
      return assemble3(h, time_step, old_solution);
    return assemble_impl(h, t, time_step, old_solution, true);
    PE_CATCH
  }

  int
  pe_assemble_rhs_only(pe_handle *h, double t, double time_step, const double *old_solution)
  {
    if (!h)
      return PE_ERR_ARG;
    if (h->cfg.dim == 3)
      return fail(h, PE_ERR_ARG, "pe_assemble_rhs_only: not built for dim = 3 yet");
    PE_DEVICE(h);
    PE_TRY
    return assemble_impl(h, t, time_step, old_solution, false);
    PE_CATCH
  }

  int
  pe_get_matrix(const pe_handle *hc, double *values)
  {
    pe_handle *h = const_cast<pe_handle *>(hc);
    if (!h || !values || !(h->device_ready || (h->cfg.dim == 3 && h->dev3.ready)) || !h->matrix_valid)
      return fail(h, PE_ERR_ARG, "pe_get_matrix: nothing assembled");
    PE_DEVICE(h);
    PE_CUDA(cudaMemcpyAsync(values, h->d_values.p, sizeof(double) * (size_t)h->nnz, cudaMemcpyDeviceToHost, h->stream));
    PE_CUDA(cudaStreamSynchronize(h->stream));
    return PE_OK;
  }

  int
  pe_get_rhs(const pe_handle *hc, double *rhs)
  {
    pe_handle *h = const_cast<pe_handle *>(hc);
    if (!h || !rhs || !(h->device_ready || (h->cfg.dim == 3 && h->dev3.ready && h->matrix_valid)))
      return fail(h, PE_ERR_ARG, "pe_get_rhs: nothing assembled");
    PE_DEVICE(h);
    PE_CUDA(cudaMemcpyAsync(rhs, h->d_rhs.p, sizeof(double) * (size_t)h->L, cudaMemcpyDeviceToHost, h->stream));
    PE_CUDA(cudaStreamSynchronize(h->stream));
    return PE_OK;
  }

  int
  pe_get_solution(const pe_handle *hc, double *solution)
  {
    pe_handle *h = const_cast<pe_handle *>(hc);
    if (!h || !solution || !h->device_ready)
      return fail(h, PE_ERR_ARG, "pe_get_solution: no device state");
    PE_DEVICE(h);
    PE_CUDA(cudaMemcpyAsync(solution, h->d_sol.p, sizeof(double) * (size_t)h->L, cudaMemcpyDeviceToHost, h->stream));
    PE_CUDA(cudaStreamSynchronize(h->stream));
    return PE_OK;
  }

  int
  pe_get_solution_global(const pe_handle *hc, double *solution_global)
  {
    pe_handle *h = const_cast<pe_handle *>(hc);
    if (!h || !solution_global || !h->device_ready)
      return fail(h, PE_ERR_ARG, "pe_get_solution_global: no device state");
    PE_DEVICE(h);
    const Partition &pt = h->part;
    for (int b = 0; b < 3; ++b)
      if (pt.row_end[b] > pt.row_begin[b])
        PE_CUDA(cudaMemcpyAsync(solution_global + pt.row_begin[b], h->d_sol.p + pt.local_off[b],
                                sizeof(double) * (size_t)(pt.row_end[b] - pt.row_begin[b]), cudaMemcpyDeviceToHost, h->stream));
    PE_CUDA(cudaStreamSynchronize(h->stream));
    return PE_OK;
  }

  int
  pe_set_solution(pe_handle *h, const double *solution)
  {
    if (!h || !solution)
      return PE_ERR_ARG;
    if (h->cfg.dim == 3)
      return fail(h, PE_ERR_ARG, "pe_set_solution: not built for dim = 3 yet");
    PE_DEVICE(h);
    PE_TRY
    const int rc = ensure_device(h);
    if (rc != PE_OK)
      return rc;
    const int r2 = upload_global_vector(h, solution, h->d_sol.p);
    if (r2 != PE_OK)
      return r2;
    PE_CUDA(cudaStreamSynchronize(h->stream));
    h->hist_n = 0; // an explicitly set state starts a new trajectory (no extrapolated initial guess across it)
    if (h->recycle)
      h->recycle->k = 0;
    return PE_OK;
    PE_CATCH
  }

  int
  pe_spmv(pe_handle *h, const double *x, double *y)
  {
    if (!h || !x || !y || !h->device_ready || !h->matrix_valid)
      return fail(h, PE_ERR_ARG, "pe_spmv: assemble first");
    PE_DEVICE(h);
    PE_TRY
    const size_t n = (size_t)(h->L + h->H);
    for (int k = 0; k < 2; ++k)
      if (h->d_w[k].n < n)
        PE_CUDA(h->d_w[k].alloc(n));
    const int rc = upload_global_vector(h, x, h->d_w[0].p);
    if (rc != PE_OK)
      return rc;
    SolverCtx c = make_solver_ctx(h);
    c.n_pos     = c.L; // plain A x
    spmv_signed(c, h->d_w[0].p, h->d_w[1].p, nullptr);
    ++h->launches;
    PE_CUDA(cudaGetLastError());
    PE_CUDA(cudaMemcpyAsync(y, h->d_w[1].p, sizeof(double) * (size_t)h->L, cudaMemcpyDeviceToHost, h->stream));
    PE_CUDA(cudaStreamSynchronize(h->stream));
    return PE_OK;
    PE_CATCH
  }

  int
  pe_set_solver_options(pe_handle *h, int use_multigrid, double theta, double omega, int nu)
  {
    if (!h || theta <= 0. || omega <= 0. || nu < 1)
      return fail(h, PE_ERR_ARG, "pe_set_solver_options: bad argument");
    h->use_mg      = use_multigrid & 3;
    // bit 5: colour-ordered first-writer scatter (deterministic) instead of pre-zero + RED atomics.  STRUCTURAL: the
    // colouring, the first-writer flags and the chunk tables are built with the pattern, so the request is only
    // recorded here and applied by the next pe_make_sparsity_pattern / pe_set_pattern -- the tables of a pattern
    // that exists keep the mode they were built for (a later call that only tunes omega / nu cannot corrupt them)
    h->pending_scatter_mode = (use_multigrid & 32) ? 1 : 0;
    // bits 8..: chunk size of the atomics mode in units of 4096 cells (0 = keep).  The chunk tables are rebuilt
    // (device tables only; the resident old_solution is kept) before the next assembly.
    if ((use_multigrid >> 8) > 0)
      {
        const int64_t cc = (int64_t)(use_multigrid >> 8) * 4096;
        if (cc != h->pending_chunk_cells)
          {
            h->pending_chunk_cells = cc;
            if (h->device_ready && h->part.scatter_mode == 0)
              h->device_ready = false;
          }
      }
    h->spmv_kernel = (use_multigrid & 16) ? 0 : 1; // bit 4: warp-per-row SpMV (A/B measurements)
    h->use_graphs  = (use_multigrid & 64) ? 0 : ((use_multigrid & 128) ? 2 : 1); // bit 6: no CUDA graphs; bit 7: also multi-GPU
    h->mg_theta = theta;
    h->mg_omega = omega;
    h->mg_nu    = nu;
    h->op_key_valid = false; // smoother weights / sweeps are baked into the preconditioner
    ++h->graph_epoch;
    return PE_OK;
  }

  int
  pe_set_tuning(pe_handle *h, const char *key, double value)
  {
    if (!h || !key)
      return PE_ERR_ARG;
    const std::string k(key);
    Tuning           &t = h->tune;
    if (k == "asm_gather")
      t.asm_gather = value != 0.;
    else if (k == "mg_omega2")
      t.mg_omega2 = value;
    else if (k == "cmg_omega")
      t.cmg_omega = value;
    else if (k == "cmg_omega2")
      t.cmg_omega2 = value;
    else if (k == "face_omega")
      t.face_omega = value;
    else if (k == "mg_replicated")
      t.mg_replicated = value != 0.;
    else if (k == "mg_whole_slabs")
      t.mg_whole_slabs = value != 0.;
    else if (k == "one_lane")
      t.one_lane = value != 0.;
    else if (k == "mg_fused_n")
      t.mg_fused_n = (int)value;
    else if (k == "mg_dist_min_n")
      t.mg_dist_min_n = (int)value;
    else if (k == "reuse_operator")
      t.reuse_operator = value != 0.;
    else if (k == "esm_dfma")
      t.esm_dfma = value != 0.;
    else
      return fail(h, PE_ERR_ARG, "pe_set_tuning: unknown key " + k);
    // structural switches take effect when the device tables / hierarchies are rebuilt
    if (k == "asm_gather" || k == "mg_replicated" || k == "mg_whole_slabs" || k == "one_lane" || k == "mg_dist_min_n")
      h->device_ready = false;
    if (h->k3)
      h->k3->values_valid = false;
    h->op_key_valid = false;
    ++h->graph_epoch;
    return PE_OK;
  }

  int
  pe_set_krylov(pe_handle *h, int method, int restart)
  {
    if (!h || method < 0 || method > 1 || restart < 0 || restart > kGmresMax)
      return fail(h, PE_ERR_ARG, "pe_set_krylov: method 0 (MINRES) or 1 (FGMRES), restart <= 24");
    if (h->krylov != method)
      {
        if (h->k3)
          h->k3->values_valid = false; // the xi rows of the solver matrix are scaled for GMRES only
        h->op_key_valid = false;
      }
    h->krylov = method;
    if (restart > 0)
      h->gmres_m = restart;
    return PE_OK;
  }

  int
  pe_set_initial_guess(pe_handle *h, int order)
  {
    if (!h || order < 0 || order > 2)
      return fail(h, PE_ERR_ARG, "pe_set_initial_guess: order must be 0, 1 or 2");
    h->extrapolate = order;
    return PE_OK;
  }

  int
  pe_set_recycling(pe_handle *h, int pairs)
  {
    if (!h || pairs < 0 || pairs > 8)
      return fail(h, PE_ERR_ARG, "pe_set_recycling: 0 .. 8 pairs");
    h->recycle_k = pairs;
    if (h->recycle)
      h->recycle->k = 0;
    return PE_OK;
  }

  int
  pe_bench_fp64(pe_handle *h, double *tflops)
  {
    if (!h || !tflops || h->device < 0)
      return fail(h, PE_ERR_ARG, "pe_bench_fp64: bad argument");
    PE_DEVICE(h);
    PE_TRY
    if (h->d_scal.n < 64)
      PE_CUDA(h->d_scal.alloc(64));
    double sec = 0., flop = 0.;
    PE_CUDA(launch_fp64_peak(h->d_scal.p + 48, h->ev[5], h->ev[6], h->stream, &sec, &flop));
    *tflops = flop / sec * 1e-12;
    ++h->launches;
    return PE_OK;
    PE_CATCH
  }

  int
  pe_bench_spmv(pe_handle *h, int repeats, double *ms_per_launch)
  {
    if (!h || repeats <= 0 || !ms_per_launch || !h->device_ready || !h->matrix_valid)
      return fail(h, PE_ERR_ARG, "pe_bench_spmv: assemble first");
    PE_DEVICE(h);
    PE_TRY
    size_t n = (size_t)(h->L + h->H);
    for (int k = 0; k < 2; ++k)
      if (h->d_w[k].n < n)
        PE_CUDA(h->d_w[k].alloc(n));
    SolverCtx c = make_solver_ctx(h);
    if (h->cfg.element != PE_WG_Q2RT2)
      {
        // the solver's own operator: the pruned three-field matrix K3
        int rc = prepare_three_field(h);
        if (rc == 0 && !h->k3->values_valid)
          {
            rc              = update_solver_operator(h);
            h->op_key       = matrix_fingerprint(h);
            h->op_key_valid = rc == 0;
          }
        if (rc != 0)
          return fail(h, PE_ERR_CUDA, "pe_bench_spmv: solver operator setup failed");
        c = make_solver_ctx3(h);
        PE_CUDA(cudaMemsetAsync(h->d_w[0].p, 0, sizeof(double) * (size_t)h->k3->n_phys(), h->stream));
      }
    PE_CUDA(cudaMemcpyAsync(h->d_w[0].p, h->d_rhs.p, sizeof(double) * n, cudaMemcpyDeviceToDevice, h->stream));
    for (int k = 0; k < 3; ++k)
      spmv_signed(c, h->d_w[0].p, h->d_w[1].p, nullptr);
    PE_CUDA(cudaEventRecord(h->ev[5], h->stream));
    for (int k = 0; k < repeats; ++k)
      spmv_signed(c, h->d_w[0].p, h->d_w[1].p, nullptr);
    PE_CUDA(cudaEventRecord(h->ev[6], h->stream));
    PE_CUDA(cudaEventSynchronize(h->ev[6]));
    h->launches += repeats + 3;
    float ms = 0.f;
    cudaEventElapsedTime(&ms, h->ev[5], h->ev[6]);
    *ms_per_launch = ms / repeats;
    h->spmv_ms_avg = *ms_per_launch;
    return PE_OK;
    PE_CATCH
  }

  int
  pe_solve_time_step(pe_handle *h, double rel_tol, int max_iterations, double *solution, pe_solve_stats *stats)
  {
    NvtxRange nvtx("pe_solve_time_step");
    if (!h || !(h->device_ready || (h->cfg.dim == 3 && h->dev3.ready)) || !h->matrix_valid)
      return fail(h, PE_ERR_ARG, "pe_solve_time_step: assemble first");
    PE_DEVICE(h);
    PE_TRY
    pe_solve_stats local{};
    if (!stats)
      stats = &local;
    std::memset(stats, 0, sizeof(*stats));
    if (h->cfg.dim == 3)
      return solve3(h, rel_tol, max_iterations, solution, stats);
    const size_t n = (size_t)(h->L + h->H);
    if (h->cfg.element == PE_WG_Q2RT2)
      for (int k = 0; k < 3; ++k)
        if (h->d_w[k].n < n)
          PE_CUDA(h->d_w[k].alloc(n));
    if (h->d_scal.n < 64)
      PE_CUDA(h->d_scal.alloc(64));
    cudaStream_t s = h->stream;
    SolverCtx    c = make_solver_ctx(h);
    int          rc;
    if (h->cfg.element == PE_WG_Q2RT2)
      {
        PE_CUDA(cudaEventRecord(h->ev[3], s));
        double *work[3] = {h->d_w[0].p, h->d_w[1].p, h->d_w[2].p};
        // initial guess = `solution` as left by apply_boundary_values (boundary values set, zero elsewhere)
        rc = cg_solve(c, h->d_rhs.p, h->d_sol.p, work, h->d_scal.p, h->h_scal, rel_tol, max_iterations, stats);
      }
    else
      {
        rc = solve_three_field(h, rel_tol, max_iterations, stats);
        if (rc == 3)
          return fail(h, PE_ERR_CUDA, "CUDA graph capture of the MINRES iteration failed (pe_set_solver_options bit 6 runs without graphs)");
        if (rc > 0)
          return rc == 1 ? cuda_fail(h, cudaGetLastError(), "three-field solver setup") : PE_ERR_ALLOC;
        if (rc == PE_ERR_PATTERN)
          return rc;
      }
    PE_CUDA(cudaEventRecord(h->ev[4], s));
    PE_CUDA(cudaEventSynchronize(h->ev[4]));
    if (rc != 0)
      return cuda_fail(h, cudaGetLastError(), "krylov solve");
    if (h->part.world > 1)
      {
        const int cst = comm_status(h);
        if (cst != PE_OK)
          return cst;
      }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, h->ev[3], h->ev[4]);
    h->solve_ms    = ms;
    stats->seconds = ms * 1e-3;
    h->launches += 12; // setup / residual kernels outside the iteration
    if (solution)
      {
        PE_CUDA(cudaMemcpyAsync(solution, h->d_sol.p, sizeof(double) * (size_t)h->L, cudaMemcpyDeviceToHost, s));
        PE_CUDA(cudaStreamSynchronize(s));
      }
    if (!stats->converged)
      return fail(h, PE_ERR_NOT_CONVERGED, "Krylov solver did not reach the requested residual");
    return PE_OK;
    PE_CATCH
  }

  int
  pe_local_matrices(pe_handle *h, int64_t first_cell, int64_t n, double t, double time_step,
                    const double *old_solution, double *out, double *rhs_out)
  {
    if (!h || !h->have_mesh || !out || n <= 0 || first_cell < 0 || first_cell + n > h->mesh.n_cells)
      return fail(h, PE_ERR_ARG, "pe_local_matrices: bad argument");
    if (h->part.world != 1)
      return fail(h, PE_ERR_ARG, "pe_local_matrices: one GPU only");
    if (h->device < 0)
      return fail(h, PE_ERR_NO_DEVICE, "no hot path without a GPU");
    PE_DEVICE(h);
    PE_TRY
    cudaStream_t s     = h->stream;
    const int    n_loc = element_n_loc(h->cfg.element);
    const size_t nn    = (size_t)n_loc * n_loc;
    // stand-alone path: only the mesh and K are needed
    DevBuf<double>   vx, vy, kc, o, ro, sol;
    DevBuf<uint32_t> cv;
    DevBuf<int32_t>  cd;
    PE_CUDA(upload(vx, h->mesh.x.data(), h->mesh.x.size(), s));
    PE_CUDA(upload(vy, h->mesh.y.data(), h->mesh.y.size(), s));
    std::vector<uint32_t> cvh((size_t)4 * n);
    for (int64_t c = 0; c < n; ++c)
      for (int v = 0; v < 4; ++v)
        cvh[(size_t)v * n + c] = h->mesh.cell_vertices[4 * (first_cell + c) + v];
    PE_CUDA(upload(cv, cvh.data(), cvh.size(), s));
    AsmArgs a{};
    a.n_asm    = n;
    a.vx       = vx.p;
    a.vy       = vy.p;
    a.cv       = cv.p;
    a.k_scalar = h->k_scalar;
    a.esm_variant = h->tune.esm_dfma;
    if (h->k_lognormal)
      {
        PE_CUDA(kc.alloc((size_t)n));
        PE_CUDA(launch_lognormal(n, nullptr, first_cell, h->k_sigma, h->k_seed, kc.p, s));
        ++h->launches;
        a.kcell = kc.p;
      }
    else if (!h->k_cell.empty())
      {
        PE_CUDA(upload(kc, h->k_cell.data() + first_cell, (size_t)n, s));
        a.kcell = kc.p;
      }
    if (old_solution)
      {
        if (!h->have_dofs)
          return fail(h, PE_ERR_ARG, "pe_local_matrices: old_solution needs dofs");
        std::vector<int32_t> cdh((size_t)n_loc * n);
        for (int64_t c = 0; c < n; ++c)
          for (int i = 0; i < n_loc; ++i)
            cdh[(size_t)i * n + c] = (int32_t)h->dofs.cell_dofs[(size_t)(first_cell + c) * n_loc + i];
        PE_CUDA(upload(cd, cdh.data(), cdh.size(), s));
        PE_CUDA(upload(sol, old_solution, (size_t)h->dofs.n(), s));
        a.cdofs = cd.p;
        a.sol   = sol.p;
      }
    PE_CUDA(o.alloc(nn * (size_t)n));
    if (rhs_out)
      PE_CUDA(ro.alloc((size_t)n_loc * n));
    if (h->cfg.element == PE_WG_Q2RT2)
      {
        // the warp-per-cell kernel writes out[n][21][21] / rhs_out[n][21] directly
        PE_CUDA(launch_esm_local_matrices(a, h->case_id, 0, n, o.p, rhs_out ? ro.p : nullptr, s));
        ++h->launches;
        PE_CUDA(cudaMemcpyAsync(out, o.p, sizeof(double) * nn * (size_t)n, cudaMemcpyDeviceToHost, s));
        if (rhs_out)
          PE_CUDA(cudaMemcpyAsync(rhs_out, ro.p, sizeof(double) * (size_t)n_loc * n, cudaMemcpyDeviceToHost, s));
        PE_CUDA(cudaStreamSynchronize(s));
        return PE_OK;
      }
    const DevParams P = make_params(h, t, time_step);
    PE_CUDA(launch_local_matrices(h->cfg.element, a, P, 0, n, o.p, rhs_out ? ro.p : nullptr, s));
    ++h->launches;
    std::vector<double> tmp(nn * (size_t)n);
    PE_CUDA(cudaMemcpyAsync(tmp.data(), o.p, sizeof(double) * tmp.size(), cudaMemcpyDeviceToHost, s));
    PE_CUDA(cudaStreamSynchronize(s));
    for (int64_t c = 0; c < n; ++c)
      for (size_t e = 0; e < nn; ++e)
        out[(size_t)c * nn + e] = tmp[e * (size_t)n + c];
    if (rhs_out)
      {
        std::vector<double> tr((size_t)n_loc * n);
        PE_CUDA(cudaMemcpyAsync(tr.data(), ro.p, sizeof(double) * tr.size(), cudaMemcpyDeviceToHost, s));
        PE_CUDA(cudaStreamSynchronize(s));
        for (int64_t c = 0; c < n; ++c)
          for (int i = 0; i < n_loc; ++i)
            rhs_out[(size_t)c * n_loc + i] = tr[(size_t)i * n + c];
      }
    return PE_OK;
    PE_CATCH
  }

  int
  pe_local_matrices_bench(pe_handle *h, double t, double time_step, int64_t chunk_cells, int repeats,
                          double *checksum, double *seconds)
  {
    if (!h || !h->have_mesh || chunk_cells <= 0 || repeats <= 0)
      return fail(h, PE_ERR_ARG, "pe_local_matrices_bench: bad argument");
    if (h->device < 0)
      return fail(h, PE_ERR_NO_DEVICE, "no hot path without a GPU");
    PE_DEVICE(h);
    PE_TRY
    cudaStream_t  s     = h->stream;
    const int     n_loc = element_n_loc(h->cfg.element);
    const size_t  nn    = (size_t)n_loc * n_loc;
    const int64_t nc    = h->mesh.n_cells;
    chunk_cells         = std::min(chunk_cells, nc);
    DevBuf<double>   vx, vy, kc, o, sum;
    DevBuf<uint32_t> cv;
    PE_CUDA(upload(vx, h->mesh.x.data(), h->mesh.x.size(), s));
    PE_CUDA(upload(vy, h->mesh.y.data(), h->mesh.y.size(), s));
    {
      std::vector<uint32_t> cvh((size_t)4 * nc);
      for (int64_t c = 0; c < nc; ++c)
        for (int v = 0; v < 4; ++v)
          cvh[(size_t)v * nc + c] = h->mesh.cell_vertices[4 * c + v];
      PE_CUDA(upload(cv, cvh.data(), cvh.size(), s));
      PE_CUDA(cudaStreamSynchronize(s));
    }
    AsmArgs a{};
    a.n_asm    = nc;
    a.vx       = vx.p;
    a.vy       = vy.p;
    a.cv       = cv.p;
    a.k_scalar = h->k_scalar;
    a.esm_variant = h->tune.esm_dfma;
    if (h->k_lognormal)
      {
        PE_CUDA(kc.alloc((size_t)nc));
        PE_CUDA(launch_lognormal(nc, nullptr, 0, h->k_sigma, h->k_seed, kc.p, s));
        a.kcell = kc.p;
      }
    else if (!h->k_cell.empty())
      {
        PE_CUDA(upload(kc, h->k_cell.data(), (size_t)nc, s));
        a.kcell = kc.p;
      }
    PE_CUDA(o.alloc(nn * (size_t)chunk_cells));
    PE_CUDA(sum.alloc(1));
    const DevParams P = make_params(h, t, time_step);
    double          total_ms = 0., cs = 0.;
    for (int rep = 0; rep <= repeats; ++rep) // rep 0 = warm-up + checksum
      {
        PE_CUDA(cudaEventRecord(h->ev[5], s));
        for (int64_t first = 0; first < nc; first += chunk_cells)
          {
            const int64_t n = std::min(chunk_cells, nc - first);
            if (h->cfg.element == PE_WG_Q2RT2)
              PE_CUDA(launch_esm_local_matrices(a, h->case_id, first, n, o.p, nullptr, s));
            else
              PE_CUDA(launch_local_matrices(h->cfg.element, a, P, first, n, o.p, nullptr, s));
            ++h->launches;
            if (rep == 0)
              {
                PE_CUDA(launch_sum(o.p, nn * (size_t)n, sum.p, s));
                double part = 0.;
                PE_CUDA(cudaMemcpyAsync(&part, sum.p, sizeof(double), cudaMemcpyDeviceToHost, s));
                PE_CUDA(cudaStreamSynchronize(s));
                cs += part;
              }
          }
        PE_CUDA(cudaEventRecord(h->ev[6], s));
        PE_CUDA(cudaEventSynchronize(h->ev[6]));
        float ms = 0.f;
        cudaEventElapsedTime(&ms, h->ev[5], h->ev[6]);
        if (rep > 0)
          total_ms += ms;
      }
    if (checksum)
      *checksum = cs;
    if (seconds)
      *seconds = total_ms * 1e-3 / repeats;
    return PE_OK;
    PE_CATCH
  }

  int
  pe_step_errors(pe_handle *h, double t, double *e)
  {
    if (!h || !e)
      return PE_ERR_ARG;
    if (h->cfg.dim == 3)
      return fail(h, PE_ERR_ARG, "pe_step_errors: not built for dim = 3 yet");
    if (h->cfg.element != PE_BR1_WG && h->cfg.element != PE_CGQ1_WG)
      return fail(h, PE_ERR_ARG, "pe_step_errors: element not implemented yet");
    PE_DEVICE(h);
    PE_TRY
    const int rc = ensure_device(h);
    if (rc != PE_OK)
      return rc;
    if (h->d_scal.n < 64)
      PE_CUDA(h->d_scal.alloc(64));
    if (h->part.world > 1)
      comm_halo_exchange(h, h->d_sol.p);
    const DevParams P = make_params(h, t, h->last_dt);
    const AsmArgs   a = make_args(h, false);
    PE_CUDA(launch_step_errors(h->cfg.element, a, P, h->d_is_ghost.p, h->d_scal.p + 32, h->stream));
    ++h->launches;
    if (h->part.world > 1)
      comm_allreduce(h, h->d_scal.p + 32, 3);
    PE_CUDA(cudaMemcpyAsync(e, h->d_scal.p + 32, 3 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    PE_CUDA(cudaStreamSynchronize(h->stream));
    return PE_OK;
    PE_CATCH
  }

  int
  pe_postprocess(pe_handle *h, double *div_displacement, double *displacement_magnitude, double *darcy_beta,
                 double *velocity_center)
  {
    if (!h)
      return PE_ERR_ARG;
    if (h->cfg.dim == 3)
      return fail(h, PE_ERR_ARG, "pe_postprocess: not built for dim = 3 yet");
    if (h->cfg.element != PE_BR1_WG && h->cfg.element != PE_CGQ1_WG)
      return fail(h, PE_ERR_ARG, "pe_postprocess: Biot elements only (the Darcy element has pe_darcy_errors)");
    PE_DEVICE(h);
    PE_TRY
    const int rc = ensure_device(h);
    if (rc != PE_OK)
      return rc;
    if (h->part.world > 1)
      comm_halo_exchange(h, h->d_sol.p);
    const int64_t nc = h->part.cell_end - h->part.cell_begin;
    DevBuf<double> out;
    PE_CUDA(out.alloc((size_t)8 * nc));
    double *d_div = out.p, *d_mag = out.p + nc, *d_beta = out.p + 2 * nc, *d_vel = out.p + 6 * nc;
    const AsmArgs a = make_args(h, false);
    PE_CUDA(launch_postprocess(h->cfg.element, a, h->d_out_slot.p, d_div, d_mag, d_beta, d_vel, h->stream));
    ++h->launches;
    cudaStream_t s = h->stream;
    if (div_displacement)
      PE_CUDA(cudaMemcpyAsync(div_displacement, d_div, sizeof(double) * (size_t)nc, cudaMemcpyDeviceToHost, s));
    if (displacement_magnitude)
      PE_CUDA(cudaMemcpyAsync(displacement_magnitude, d_mag, sizeof(double) * (size_t)nc, cudaMemcpyDeviceToHost, s));
    if (darcy_beta)
      PE_CUDA(cudaMemcpyAsync(darcy_beta, d_beta, sizeof(double) * (size_t)(4 * nc), cudaMemcpyDeviceToHost, s));
    if (velocity_center)
      PE_CUDA(cudaMemcpyAsync(velocity_center, d_vel, sizeof(double) * (size_t)(2 * nc), cudaMemcpyDeviceToHost, s));
    PE_CUDA(cudaStreamSynchronize(s));
    if (h->part.world > 1)
      return comm_status(h);
    return PE_OK;
    PE_CATCH
  }

  int
  pe_darcy_errors(pe_handle *h, double *e)
  {
    if (!h || !e)
      return PE_ERR_ARG;
    if (h->cfg.element != PE_WG_Q2RT2)
      return fail(h, PE_ERR_ARG, "pe_darcy_errors: the WG(Q2,Q2;RT[2]) Darcy element only (Biot elements: pe_postprocess)");
    PE_DEVICE(h);
    PE_TRY
    const int rc = ensure_device(h);
    if (rc != PE_OK)
      return rc;
    if (h->d_scal.n < 64)
      PE_CUDA(h->d_scal.alloc(64));
    const AsmArgs a = make_args(h, false);
    PE_CUDA(launch_esm_postprocess(a, h->d_sol.p, h->d_scal.p + 36, h->stream));
    ++h->launches;
    PE_CUDA(cudaMemcpyAsync(e, h->d_scal.p + 36, 2 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    PE_CUDA(cudaStreamSynchronize(h->stream));
    return PE_OK;
    PE_CATCH
  }

  int
  pe_comm_unique_id(uint8_t id[128])
  {
    return comm_unique_id(id);
  }

  int
  pe_comm_init(pe_handle *h, const uint8_t id[128])
  {
    if (!h || !id)
      return PE_ERR_ARG;
    return comm_init(h, id);
  }

  int
  pe_get_timings(const pe_handle *h, double *assemble_ms, double *assemble_kernel_ms, double *solve_ms,
                 double *spmv_ms_avg, int64_t *launches)
  {
    if (!h)
      return PE_ERR_ARG;
    if (assemble_ms)
      *assemble_ms = h->assemble_ms;
    if (assemble_kernel_ms)
      *assemble_kernel_ms = h->assemble_kernel_ms;
    if (solve_ms)
      *solve_ms = h->solve_ms;
    if (spmv_ms_avg)
      *spmv_ms_avg = h->spmv_ms_avg;
    if (launches)
      *launches = h->launches;
    return PE_OK;
  }
}
