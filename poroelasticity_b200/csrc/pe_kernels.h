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
    const double   *vx, *vy, *kcell, *sol;
    double          k_scalar;
    const uint32_t *cv, *flags;
    const int32_t  *cdofs, *bnd_list;
    const int64_t  *rowptr;
    const uint8_t  *off;
    double         *values, *rhs;
    int             do_matrix;
  };

  cudaError_t launch_scatter_offsets(int n_loc, int64_t n_asm, int64_t L, const int32_t *cdofs,
                                     const int64_t *rowptr, const int32_t *colidx, uint8_t *off, int *err,
                                     cudaStream_t s);
  cudaError_t launch_assemble(int element, const AsmArgs &a, const DevParams &P, cudaStream_t s, int *n_launches);
  cudaError_t launch_local_matrices(int element, const AsmArgs &a, const DevParams &P, int64_t first, int64_t n,
                                    double *out, double *rhs_out, cudaStream_t s);
  cudaError_t launch_lognormal(int64_t n, const int64_t *cell_ids, int64_t first, double sigma, uint64_t seed,
                               double *k, cudaStream_t s);
  cudaError_t launch_sum(const double *x, size_t n, double *out, cudaStream_t s);

  // ---- solver kernels (pe_solver.cu)
  cudaError_t launch_spmv(int64_t n_rows, const int64_t *rowptr, const int32_t *colidx, const double *values,
                          const double *x, double *y, cudaStream_t s);
} // namespace pe
