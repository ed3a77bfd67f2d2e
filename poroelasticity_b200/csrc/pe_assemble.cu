// CUDA kernels of the assembly hot path (sm_100a, FP64):
//   K1  cell kernel (pe_cell.cuh)  +  K2 scatter-add into the CSR values  +  K3 per-cell Dirichlet
//   elimination, fused.  Replaces the cell loop of assemble_system and
//   AffineConstraints::distribute_local_to_global (PoroElasBR1WG_new.cc:791-1170,
//   PoroElasCGQ1_WG.cc:1165-1725).
#include "pe_internal.h"
#include "pe_kernels.h"

namespace pe
{
  // ---------------------------------------------------------------------------------------------
  // setup kernel: position of entry (i,j) of every cell inside the CSR row of dof i
  // ---------------------------------------------------------------------------------------------
  __global__ void
  k_scatter_offsets(const int n_loc, const int64_t n_asm, const int64_t L, const int32_t *__restrict__ cdofs,
                    const int64_t *__restrict__ rowptr, const int32_t *__restrict__ colidx,
                    uint8_t *__restrict__ off, int *__restrict__ err)
  {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_asm * n_loc)
      return;
    const int64_t c = t % n_asm;
    const int     i = (int)(t / n_asm);
    const int32_t r = cdofs[(size_t)i * n_asm + c];
    if (r >= L)
      {
        for (int j = 0; j < n_loc; ++j)
          off[((size_t)i * n_loc + j) * n_asm + c] = 0;
        return;
      }
    const int64_t b = rowptr[r], e = rowptr[r + 1];
    if (e - b > 256)
      {
        atomicExch(err, 2);
        return;
      }
    for (int j = 0; j < n_loc; ++j)
      {
        const int32_t col = cdofs[(size_t)j * n_asm + c];
        int64_t       lo = b, hi = e;
        while (lo < hi)
          {
            const int64_t mid = (lo + hi) >> 1;
            if (colidx[mid] < col)
              lo = mid + 1;
            else
              hi = mid;
          }
        if (lo >= e || colidx[lo] != col)
          {
            atomicExch(err, 1);
            lo = b;
          }
        off[((size_t)i * n_loc + j) * n_asm + c] = (uint8_t)(lo - b);
      }
  }

  // ---------------------------------------------------------------------------------------------
  // emit functors
  // ---------------------------------------------------------------------------------------------
  template <int NL, bool MAT>
  struct ScatterEmit // straight into the CSR values with FP64 RED atomics
  {
    double        *values, *rhsp;
    const uint8_t *off; // + c
    size_t         n_asm;
    int64_t        rs[NL]; // CSR row start of local dof i, -1 if the row is not owned
    int32_t        row[NL];
    __device__ __forceinline__ void
    mat(const int i, const int j, const double v)
    {
      if (MAT && v != 0. && rs[i] >= 0)
        atomicAdd(values + rs[i] + off[(size_t)(i * NL + j) * n_asm], v);
    }
    __device__ __forceinline__ void
    rhs(const int i, const double v)
    {
      if (v != 0. && rs[i] >= 0)
        atomicAdd(rhsp + row[i], v);
    }
  };

  template <int NL>
  struct LocalEmit // into thread-local arrays (boundary cells)
  {
    double *A, *b;
    __device__ __forceinline__ void
    mat(const int i, const int j, const double v)
    {
      A[i * NL + j] = v;
    }
    __device__ __forceinline__ void
    rhs(const int i, const double v)
    {
      b[i] = v;
    }
  };

  template <int NL>
  struct SoaEmit // dense local matrices, entry-major [NL*NL][n] (BASELINE cfg 5)
  {
    double *out, *rhs_out;
    size_t  n;
    __device__ __forceinline__ void
    mat(const int i, const int j, const double v)
    {
      out[(size_t)(i * NL + j) * n] = v;
    }
    __device__ __forceinline__ void
    rhs(const int i, const double v)
    {
      if (rhs_out)
        rhs_out[(size_t)i * n] = v;
    }
  };

  template <int EL>
  __device__ __forceinline__ void
  load_cell(const AsmArgs &a, const int64_t c, double vx[4], double vy[4], int32_t idx[Elem<EL>::NL],
            double old[Elem<EL>::NL])
  {
#pragma unroll
    for (int v = 0; v < 4; ++v)
      {
        const uint32_t vi = a.cv[(size_t)v * a.n_asm + c];
        vx[v]             = __ldg(a.vx + vi);
        vy[v]             = __ldg(a.vy + vi);
      }
#pragma unroll
    for (int i = 0; i < Elem<EL>::NL; ++i)
      {
        idx[i] = a.cdofs[(size_t)i * a.n_asm + c];
        old[i] = a.sol ? __ldg(a.sol + idx[i]) : 0.;
      }
  }

  // ---------------------------------------------------------------------------------------------
  // interior cells (no boundary face, no constrained dof): K1 + K2
  // ---------------------------------------------------------------------------------------------
  template <int EL, bool MAT>
  __global__ void __launch_bounds__(128)
  k_assemble_interior(const AsmArgs a, const DevParams P)
  {
    constexpr int NL = Elem<EL>::NL;
    const int64_t c  = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= a.n_asm)
      return;
    if (a.flags[c] != 0u)
      return; // handled by k_assemble_boundary
    double  vx[4], vy[4], old[NL];
    int32_t idx[NL];
    load_cell<EL>(a, c, vx, vy, idx, old);
    Geom G;
    G.set(vx, vy);
    ScatterEmit<NL, MAT> E;
    E.values = a.values;
    E.rhsp   = a.rhs;
    E.off    = a.off + c;
    E.n_asm  = (size_t)a.n_asm;
#pragma unroll
    for (int i = 0; i < NL; ++i)
      {
        E.row[i] = idx[i];
        E.rs[i]  = idx[i] < a.L ? __ldg(a.rowptr + idx[i]) : -1;
      }
    cell_kernel<EL>(G, a.kcell ? a.kcell[c] : a.k_scalar, P, old, E);
  }

  // ---------------------------------------------------------------------------------------------
  // boundary cells: K1 into thread-local storage, then AffineConstraints::distribute_local_to_global
  // with use_inhomogeneities_for_rhs = true (deal.II 9.1 affine_constraints.templates.h; SURVEY
  // 8(a) row A8):
  //   unconstrained row r:  Mat(r,c) += A(r,c) for unconstrained c;  rhs(r) += b(r) - sum_c A(r,c) g_c
  //   constrained dof c  :  d = |A(c,c)| != 0 ? |A(c,c)| : mean_i |A(i,i)| ;  Mat(c,c) += d ; rhs(c) += d g_c
  // ---------------------------------------------------------------------------------------------
  template <int EL>
  __global__ void __launch_bounds__(64)
  k_assemble_boundary(const AsmArgs a, const DevParams P)
  {
    constexpr int NL = Elem<EL>::NL;
    const int64_t t  = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.n_bnd)
      return;
    const int64_t c = a.bnd_list[t];
    double        vx[4], vy[4], old[NL];
    int32_t       idx[NL];
    load_cell<EL>(a, c, vx, vy, idx, old);
    Geom G;
    G.set(vx, vy);
    double        A[NL * NL], b[NL], g[NL];
    bool          con[NL];
    LocalEmit<NL> E{A, b};
    cell_kernel<EL>(G, a.kcell ? a.kcell[c] : a.k_scalar, P, old, E);
    cell_boundary<EL>(G, vx, vy, a.flags[c], P, b, con, g);

    bool   any = false;
    double avg = 0.;
    for (int i = 0; i < NL; ++i)
      {
        any |= con[i];
        avg += fabs(A[i * NL + i]);
      }
    avg /= (double)NL;
    for (int i = 0; i < NL; ++i)
      {
        if (idx[i] >= a.L)
          continue; // row owned by another rank
        const int64_t  rs = a.rowptr[idx[i]];
        const uint8_t *o  = a.off + (size_t)(i * NL) * a.n_asm + c;
        if (con[i])
          {
            const double d = fabs(A[i * NL + i]) != 0. ? fabs(A[i * NL + i]) : avg;
            if (a.do_matrix)
              atomicAdd(a.values + rs + o[(size_t)i * a.n_asm], d);
            atomicAdd(a.rhs + idx[i], d * g[i]);
            continue;
          }
        double val = b[i];
        for (int j = 0; j < NL; ++j)
          {
            const double v = A[i * NL + j];
            if (con[j])
              {
                if (g[j] != 0.)
                  val -= v * g[j];
              }
            else if (a.do_matrix && v != 0.)
              atomicAdd(a.values + rs + o[(size_t)j * a.n_asm], v);
          }
        if (val != 0.)
          atomicAdd(a.rhs + idx[i], val);
      }
    (void)any;
  }

  // ---------------------------------------------------------------------------------------------
  // local matrices only (BASELINE cfg 5 and the parity hook pe_local_matrices)
  // ---------------------------------------------------------------------------------------------
  template <int EL>
  __global__ void __launch_bounds__(128)
  k_local_matrices(const AsmArgs a, const DevParams P, const int64_t first, const int64_t n, double *out,
                   double *rhs_out)
  {
    constexpr int NL = Elem<EL>::NL;
    const int64_t t  = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n)
      return;
    const int64_t c = first + t;
    double        vx[4], vy[4], old[NL];
#pragma unroll
    for (int v = 0; v < 4; ++v)
      {
        const uint32_t vi = a.cv[(size_t)v * a.n_asm + c];
        vx[v]             = __ldg(a.vx + vi);
        vy[v]             = __ldg(a.vy + vi);
      }
#pragma unroll
    for (int i = 0; i < NL; ++i)
      old[i] = (a.sol && a.cdofs) ? __ldg(a.sol + a.cdofs[(size_t)i * a.n_asm + c]) : 0.;
    Geom G;
    G.set(vx, vy);
    SoaEmit<NL> E{out + t, rhs_out ? rhs_out + t : nullptr, (size_t)n};
    cell_kernel<EL>(G, a.kcell ? a.kcell[c] : a.k_scalar, P, old, E);
  }

  // K = exp(sigma Z), Z by Box-Muller from the counter generator keyed on (seed, global cell id)
  __device__ __forceinline__ uint64_t
  d_splitmix64(uint64_t z)
  {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
  }
  __device__ __forceinline__ double
  d_u01(uint64_t seed, uint64_t id, uint64_t k)
  {
    const uint64_t r = d_splitmix64(d_splitmix64(seed ^ (id * 0xD1B54A32D192ED03ull)) + k);
    return ((double)(r >> 11) + 0.5) * (1.0 / 9007199254740992.0);
  }
  __global__ void
  k_lognormal(const int64_t n, const int64_t *__restrict__ cell_ids, const int64_t first, const double sigma,
              const uint64_t seed, double *__restrict__ k)
  {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n)
      return;
    const uint64_t id = (uint64_t)(cell_ids ? cell_ids[t] : first + t);
    const double   u1 = d_u01(seed, id, 0), u2 = d_u01(seed, id, 1);
    const double   z  = sqrt(-2. * log(u1)) * cospi(2. * u2);
    k[t]              = exp(sigma * z);
  }

  __global__ void
  k_sum(const double *__restrict__ x, const size_t n, double *__restrict__ out)
  {
    double s = 0.;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
      s += x[i];
    for (int o = 16; o > 0; o >>= 1)
      s += __shfl_xor_sync(0xffffffffu, s, o);
    __shared__ double sh[32];
    if ((threadIdx.x & 31) == 0)
      sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32)
      {
        s = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.;
        for (int o = 16; o > 0; o >>= 1)
          s += __shfl_xor_sync(0xffffffffu, s, o);
        if (threadIdx.x == 0)
          atomicAdd(out, s);
      }
  }

  // ---------------------------------------------------------------------------------------------
  // launchers
  // ---------------------------------------------------------------------------------------------
  static inline unsigned
  grid_for(const int64_t n, const int block)
  {
    return (unsigned)((n + block - 1) / block);
  }

  cudaError_t
  launch_scatter_offsets(const int n_loc, const int64_t n_asm, const int64_t L, const int32_t *cdofs,
                         const int64_t *rowptr, const int32_t *colidx, uint8_t *off, int *err, cudaStream_t s)
  {
    const int64_t n = n_asm * n_loc;
    if (n == 0)
      return cudaSuccess;
    k_scatter_offsets<<<grid_for(n, 256), 256, 0, s>>>(n_loc, n_asm, L, cdofs, rowptr, colidx, off, err);
    return cudaGetLastError();
  }

  cudaError_t
  launch_assemble(const int element, const AsmArgs &a, const DevParams &P, cudaStream_t s, int *n_launches)
  {
    if (a.n_asm > 0)
      {
        const unsigned g = grid_for(a.n_asm, 128);
        if (element == 0 && a.do_matrix)
          k_assemble_interior<0, true><<<g, 128, 0, s>>>(a, P);
        else if (element == 0)
          k_assemble_interior<0, false><<<g, 128, 0, s>>>(a, P);
        else if (a.do_matrix)
          k_assemble_interior<1, true><<<g, 128, 0, s>>>(a, P);
        else
          k_assemble_interior<1, false><<<g, 128, 0, s>>>(a, P);
        ++*n_launches;
      }
    if (a.n_bnd > 0)
      {
        if (element == 0)
          k_assemble_boundary<0><<<grid_for(a.n_bnd, 64), 64, 0, s>>>(a, P);
        else
          k_assemble_boundary<1><<<grid_for(a.n_bnd, 64), 64, 0, s>>>(a, P);
        ++*n_launches;
      }
    return cudaGetLastError();
  }

  cudaError_t
  launch_local_matrices(const int element, const AsmArgs &a, const DevParams &P, const int64_t first,
                        const int64_t n, double *out, double *rhs_out, cudaStream_t s)
  {
    if (n == 0)
      return cudaSuccess;
    if (element == 0)
      k_local_matrices<0><<<grid_for(n, 128), 128, 0, s>>>(a, P, first, n, out, rhs_out);
    else
      k_local_matrices<1><<<grid_for(n, 128), 128, 0, s>>>(a, P, first, n, out, rhs_out);
    return cudaGetLastError();
  }

  cudaError_t
  launch_lognormal(const int64_t n, const int64_t *cell_ids, const int64_t first, const double sigma,
                   const uint64_t seed, double *k, cudaStream_t s)
  {
    if (n == 0)
      return cudaSuccess;
    k_lognormal<<<grid_for(n, 256), 256, 0, s>>>(n, cell_ids, first, sigma, seed, k);
    return cudaGetLastError();
  }

  cudaError_t
  launch_sum(const double *x, const size_t n, double *out, cudaStream_t s)
  {
    cudaMemsetAsync(out, 0, sizeof(double), s);
    k_sum<<<148 * 4, 256, 0, s>>>(x, n, out);
    return cudaGetLastError();
  }
} // namespace pe
