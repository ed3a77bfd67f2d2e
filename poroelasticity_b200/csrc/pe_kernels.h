// Kernel launchers shared between the translation units of libporoelas_b200.so.
#pragma once
#include "pe_cell.cuh"

#include <cuda_runtime.h>

#include <cstdint>

namespace pe
{
  struct AsmArgs
  {
    int64_t         n_asm, n_bnd, L;
    int64_t         a_begin, a_end, bnd_begin, bnd_end; // ranges of this launch (one colour)
    const uint32_t *rhs_first;                          // bit i: this cell is the first writer of rhs row i
    const double   *vx, *vy, *kcell, *sol;
    double          k_scalar;
    const uint32_t *cv, *flags;
    const int32_t  *cdofs, *bnd_list;
    const int64_t  *rowptr;
    const uint8_t  *off;
    double         *values, *rhs;
    int             do_matrix;
    int             atomic_mode; // 1: matrix pre-zeroed, every contribution is a RED (cells in Morton order)
    int             esm_variant; // WG(Q2,Q2;RT[2]) cell kernel: 0 = FP64 tensor-core (DMMA) version, 1 = first (DFMA) version
  };

  cudaError_t launch_scatter_offsets(int n_loc, int64_t n_asm, int64_t L, const int32_t *cdofs,
                                     const int64_t *rowptr, const int32_t *colidx, const int64_t *rc_ptr,
                                     const int32_t *rc_idx, uint8_t *off, uint32_t *rhs_first, int *err, uint8_t *gsrc,
                                     int *gather_bad, cudaStream_t s);
  int         scatter_table_words(int n_loc); // 32-bit words per cell of the scatter-offset table
  cudaError_t launch_assemble(int element, const AsmArgs &a, const DevParams &P, int n_colors, const int64_t *color_begin,
                              const int64_t *bnd_color_begin, cudaStream_t s, int *n_launches);
  cudaError_t launch_assemble_chunked(int element, const AsmArgs &a, const DevParams &P, int n_chunks,
                                      const int64_t *chunk_cell_begin, const int64_t *chunk_zero, cudaStream_t s,
                                      int *n_launches);
  // gather assembly (pe_assemble.cu): staged local entries, then one CTA per tile of kTileCells cells sums the CSR
  // entries of the rows first touched by the tile from shared memory; atomics only across chunk boundaries
  constexpr int kTileCells = 32, kTileGhostMax = 24, kTileThreads = 512;
  struct GatherTiles
  {
    const int64_t *tile_rows      = nullptr; // [n_tiles][6]
    const int32_t *tile_ghost_ptr = nullptr; // [n_tiles + 1]
    const int32_t *tile_ghost     = nullptr;
  };
  cudaError_t launch_assemble_gather(int element, const AsmArgs &a, const DevParams &P, int n_chunks,
                                     const int64_t *chunk_cell_begin, const int64_t *chunk_rows, double *T,
                                     int64_t chunk_stride, const uint8_t *gsrc, const int32_t *rowadj,
                                     const GatherTiles *tiles, cudaStream_t s, int *n_launches);
  cudaError_t launch_local_matrices(int element, const AsmArgs &a, const DevParams &P, int64_t first, int64_t n,
                                    double *out, double *rhs_out, cudaStream_t s);
  cudaError_t launch_step_errors(int element, const AsmArgs &a, const DevParams &P, const uint8_t *is_ghost, double *e,
                                 cudaStream_t s);
  // postprocess(): dilation, |u|, Darcy velocity per owned cell (out_slot[c] = output index, -1 for ghost cells)
  cudaError_t launch_postprocess(int element, const AsmArgs &a, const int32_t *out_slot, double *div_out, double *mag_out,
                                 double *beta_out, double *vel_out, cudaStream_t s);
  cudaError_t launch_lognormal(int64_t n, const int64_t *cell_ids, int64_t first, double sigma, uint64_t seed,
                               double *k, cudaStream_t s);
  cudaError_t launch_sum(const double *x, size_t n, double *out, cudaStream_t s);
  // DFMA peak of the device (16 independent chains per thread, all SMs): seconds and flop of the timed launch
  cudaError_t launch_fp64_peak(double *sink, cudaEvent_t e0, cudaEvent_t e1, cudaStream_t s, double *seconds, double *flop);

  // ---- WG(Q2,Q2;RT[2]) element of EltStiffMat.cc (pe_esm.cu)
  cudaError_t launch_esm_assemble(const AsmArgs &a, int case_id, cudaStream_t s, int *n_launches);
  cudaError_t launch_esm_local_matrices(const AsmArgs &a, int case_id, int64_t first, int64_t n, double *out, double *rhs_out,
                                        cudaStream_t s);
  // postprocess() ESM:662-919: e2[0] = L2_error_vel^2, e2[1] = L2_error_flux^2 of the solution `sol`
  cudaError_t launch_esm_postprocess(const AsmArgs &a, const double *sol, double *e2, cudaStream_t s);
  cudaError_t launch_esm_boundary_values(int n, const int32_t *dofs, const double *vals, const uint8_t *is_bnd,
                                         const int64_t *rowptr, const int32_t *colidx, double *values, double *rhs,
                                         double *sol, cudaStream_t s);

  // ---- CGQ1-WG on hexahedra (pe_cell3d.cu): dense local matrices entry-major [31*31][n], rhs [31][n]
  cudaError_t launch_local_matrices_3d(int64_t n_cells, int64_t first, int64_t n, const double *x, const double *y, const double *z,
                                       const uint32_t *cv, const double *kcell, double k_scalar, double lambda, double mu,
                                       double alpha, double c0, double dt, const double *cell_old, double *out, double *rhs_out,
                                       cudaStream_t s);

  // assemble_system in 3-D: gather of the old solution per cell, constraints + traction + scatter of the dense blocks
  cudaError_t launch_gather_old_3d(int64_t n_cells, int64_t first, int64_t n, const int32_t *cdofs, const double *old,
                                   double *cell_old, cudaStream_t s);
  cudaError_t launch_scatter_3d(int64_t n_cells, int64_t first, int64_t n, const double *x, const double *y, const double *z,
                                const uint32_t *cv, const int32_t *cdofs, const uint32_t *con_mask, const uint8_t *neu_mask,
                                const double traction[3], const double *A, const double *b, const int64_t *rowptr,
                                const int32_t *colidx, double *values, double *rhs, int *err, cudaStream_t s);

  // ---- solver kernels (pe_solver.cu)
  cudaError_t launch_spmv(int64_t n_rows, const int64_t *rowptr, const int32_t *colidx, const double *values,
                          const double *x, double *y, cudaStream_t s);
} // namespace pe
