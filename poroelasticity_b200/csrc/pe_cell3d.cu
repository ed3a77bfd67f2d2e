// CGQ1-WG on hexahedra: the cell loop of PoroElas_CGQ1_WG<3>::assemble_system (PoroElasCGQ1_WG.cc:1165-1453; the
// shipped main() is the 3-D program, PoroElasCGQ1_WG.cc:3029).  First step of SURVEY 8(f).4: the 31 x 31 local
// matrices and local right-hand sides, one thread per cell, dense output -- no numbering / scatter / solve in 3-D yet.
//
// Lean form of the reference's loops (checked against the literal oracle, oracle/poro_oracle_3d.hpp):
//  * FE_RaviartThomas(0) with the contravariant Piola map: both parts of the weak-gradient right-hand side are
//    geometry independent on ANY trilinear hexahedron (int_K div w_k = +-1, int_f w_k . n = +-1 on the own face, 0 on
//    the others), so F is constant and the Darcy block is  dt k G M^-1 G^T  with the 7 x 6 sign matrix
//    G = [F_int; F_face] and M the 6 x 6 RT0 Gram matrix (QGauss<3>(3), Cholesky);
//  * 2 mu eps(phi_i):eps(phi_j) = mu (d_d N_a d_c N_b + delta_cd grad N_a . grad N_b) for phi_i = N_a e_c, phi_j = N_b e_d;
//  * lambda div div, -+alpha coupling and the SECOND c0 mass term live at the cell midpoint (QGauss<3>(1)).
// Local dof order (FESystem(FE_Q(1)^3, FE_DGQ(0), FE_FaceQ(0))): 3 v + c for vertex v (lexicographic), 24 + f for the
// face pressures, 30 for the interior pressure.
#include "pe_internal.h"
#include "pe_kernels.h"

namespace pe
{
  namespace
  {
    constexpr int NL3 = 31;
    __constant__ double kG3x[3] = {0.11270166537925831148, 0.5, 0.88729833462074168852};
    __constant__ double kG3w[3] = {5. / 18., 8. / 18., 5. / 18.};

    struct Hex
    {
      double vx[8], vy[8], vz[8];
      // J = dx/dr at r, reference gradients of the 8 trilinear shapes
      __device__ __forceinline__ void
      jac(const double r0, const double r1, const double r2, double J[3][3], double dN[8][3]) const
      {
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
          for (int b = 0; b < 3; ++b)
            J[a][b] = 0.;
#pragma unroll
        for (int v = 0; v < 8; ++v)
          {
            const double f0 = (v & 1) ? r0 : 1. - r0, f1 = ((v >> 1) & 1) ? r1 : 1. - r1, f2 = (v >> 2) ? r2 : 1. - r2;
            const double d0 = (v & 1) ? 1. : -1., d1 = ((v >> 1) & 1) ? 1. : -1., d2 = (v >> 2) ? 1. : -1.;
            dN[v][0] = d0 * f1 * f2;
            dN[v][1] = f0 * d1 * f2;
            dN[v][2] = f0 * f1 * d2;
#pragma unroll
            for (int b = 0; b < 3; ++b)
              {
                J[0][b] += vx[v] * dN[v][b];
                J[1][b] += vy[v] * dN[v][b];
                J[2][b] += vz[v] * dN[v][b];
              }
          }
      }
    };
    __device__ __forceinline__ double
    det3(const double J[3][3])
    {
      return J[0][0] * (J[1][1] * J[2][2] - J[1][2] * J[2][1]) - J[0][1] * (J[1][0] * J[2][2] - J[1][2] * J[2][0]) +
             J[0][2] * (J[1][0] * J[2][1] - J[1][1] * J[2][0]);
    }
    // physical gradients g[v][d] = sum_b dN[v][b] Jinv[b][d]
    __device__ __forceinline__ void
    phys_grads(const double J[3][3], const double det, const double dN[8][3], double g[8][3])
    {
      double Ji[3][3];
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b)
          {
            const int a1 = (a + 1) % 3, a2 = (a + 2) % 3, b1 = (b + 1) % 3, b2 = (b + 2) % 3;
            Ji[b][a] = (J[a1][b1] * J[a2][b2] - J[a1][b2] * J[a2][b1]) / det;
          }
#pragma unroll
      for (int v = 0; v < 8; ++v)
#pragma unroll
        for (int d = 0; d < 3; ++d)
          g[v][d] = dN[v][0] * Ji[0][d] + dN[v][1] * Ji[1][d] + dN[v][2] * Ji[2][d];
    }

    __global__ void __launch_bounds__(64)
    k_local_matrices_3d(const int64_t n_cells, const int64_t first, const int64_t n, const double *__restrict__ x,
                        const double *__restrict__ y, const double *__restrict__ z, const uint32_t *__restrict__ cv /* [8][n_cells] */,
                        const double *__restrict__ kcell, const double k_scalar, const double lambda, const double mu,
                        const double alpha, const double c0, const double dt, const double *__restrict__ cell_old /* [31][n] or null */,
                        double *__restrict__ out /* [31*31][n] */, double *__restrict__ rhs_out /* [31][n] or null */)
    {
      const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
      if (t >= n)
        return;
      const int64_t c = first + t;
      Hex           H;
#pragma unroll
      for (int v = 0; v < 8; ++v)
        {
          const uint32_t id = cv[(size_t)v * n_cells + c];
          H.vx[v]           = x[id];
          H.vy[v]           = y[id];
          H.vz[v]           = z[id];
        }
      auto A = [&](const int i, const int j) -> double & { return out[(size_t)(i * NL3 + j) * n + t]; };
      // ---- pass 1 over QGauss<3>(3): RT0 Gram matrix (per-axis pairs), cell volume, gradient table
      double M[6][6], grad[27][8][3], jxw[27], vol = 0.;
#pragma unroll
      for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int j = 0; j < 6; ++j)
          M[i][j] = 0.;
      for (int q = 0; q < 27; ++q)
        {
          const int    qi = q % 3, qj = (q / 3) % 3, qk = q / 9;
          const double r[3] = {kG3x[qi], kG3x[qj], kG3x[qk]}, wq = kG3w[qi] * kG3w[qj] * kG3w[qk];
          double       J[3][3], dN[8][3];
          H.jac(r[0], r[1], r[2], J, dN);
          const double det = det3(J);
          jxw[q]           = wq * det;
          vol += jxw[q];
          phys_grads(J, det, dN, grad[q]);
          // (J w^_i) . (J w^_j) / det * wq,  w^_{2d+s} = f_s(r_d) e_d
          const double rr = wq / det;
#pragma unroll
          for (int d = 0; d < 3; ++d)
#pragma unroll
            for (int e = 0; e < 3; ++e)
              {
                const double gde = (J[0][d] * J[0][e] + J[1][d] * J[1][e] + J[2][d] * J[2][e]) * rr;
#pragma unroll
                for (int s = 0; s < 2; ++s)
#pragma unroll
                  for (int u = 0; u < 2; ++u)
                    M[2 * d + s][2 * e + u] += (s ? r[d] : 1. - r[d]) * (u ? r[e] : 1. - r[e]) * gde;
              }
        }
      // ---- Darcy block: S = dt k G M^-1 G^T, rows / columns (p_int, faces 0..5); M^-1 columns by Cholesky
      {
        double Lc[6][6];
#pragma unroll
        for (int j = 0; j < 6; ++j)
          {
            double d = M[j][j];
            for (int k = 0; k < j; ++k)
              d -= Lc[j][k] * Lc[j][k];
            d        = sqrt(d);
            Lc[j][j] = d;
            for (int i = j + 1; i < 6; ++i)
              {
                double s = M[i][j];
                for (int k = 0; k < j; ++k)
                  s -= Lc[i][k] * Lc[j][k];
                Lc[i][j] = s / d;
              }
          }
        // Y = L^-1 G^T (6 x 7): column 0 = F_int = (+1,-1,+1,-1,+1,-1), column 1+f = F_face e_f = (-1)^(f+1) e_f
        double Y[6][7];
#pragma unroll
        for (int col = 0; col < 7; ++col)
          {
#pragma unroll
            for (int i = 0; i < 6; ++i)
              {
                double s = col == 0 ? ((i & 1) ? -1. : 1.) : (i == col - 1 ? ((i & 1) ? 1. : -1.) : 0.);
                for (int k = 0; k < i; ++k)
                  s -= Lc[i][k] * Y[k][col];
                Y[i][col] = s / Lc[i][i];
              }
          }
        const double kk = dt * (kcell ? kcell[c] : k_scalar);
#pragma unroll
        for (int a = 0; a < 7; ++a)
#pragma unroll
          for (int b = 0; b < 7; ++b)
            {
              double s = 0.;
#pragma unroll
              for (int i = 0; i < 6; ++i)
                s += Y[i][a] * Y[i][b];
              const int ia = a == 0 ? 30 : 23 + a, ib = b == 0 ? 30 : 23 + b;
              A(ia, ib)    = kk * s;
            }
      }
      // ---- midpoint rule: divergences of the vertex shapes, |J_c|
      double divc[8][3], jc;
      {
        double J[3][3], dN[8][3];
        H.jac(.5, .5, .5, J, dN);
        jc = det3(J);
        phys_grads(J, jc, dN, divc);
      }
      // ---- displacement block: mu (d_d N_a d_c N_b + delta_cd grad N_a . grad N_b) + lambda |J_c| d_c N_a d_d N_b (midpoint)
      for (int a = 0; a < 8; ++a)
        for (int b = 0; b < 8; ++b)
          {
            double e[3][3];
#pragma unroll
            for (int cc = 0; cc < 3; ++cc)
#pragma unroll
              for (int dd = 0; dd < 3; ++dd)
                e[cc][dd] = 0.;
            for (int q = 0; q < 27; ++q)
              {
                const double *ga = grad[q][a], *gb = grad[q][b];
                const double  dot = (ga[0] * gb[0] + ga[1] * gb[1] + ga[2] * gb[2]) * jxw[q];
#pragma unroll
                for (int cc = 0; cc < 3; ++cc)
#pragma unroll
                  for (int dd = 0; dd < 3; ++dd)
                    e[cc][dd] += ga[dd] * gb[cc] * jxw[q];
                e[0][0] += dot;
                e[1][1] += dot;
                e[2][2] += dot;
              }
#pragma unroll
            for (int cc = 0; cc < 3; ++cc)
#pragma unroll
              for (int dd = 0; dd < 3; ++dd)
                A(3 * a + cc, 3 * b + dd) = mu * e[cc][dd] + lambda * jc * divc[a][cc] * divc[b][dd];
          }
      // ---- coupling, storage, structural zeros
      for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int cc = 0; cc < 3; ++cc)
          {
            const int i = 3 * a + cc;
            A(i, 30)    = -alpha * divc[a][cc] * jc;
            A(30, i)    = alpha * divc[a][cc] * jc;
            for (int f = 0; f < 6; ++f)
              {
                A(i, 24 + f) = 0.;
                A(24 + f, i) = 0.;
              }
          }
      A(30, 30) += c0 * (vol + jc); // c0 mass at QGauss(3) AND at QGauss(1), CGQ1:1337 and 1354
      if (rhs_out)
        {
          double pold = 0., dold = 0.;
          if (cell_old)
            {
              pold = cell_old[(size_t)30 * n + t];
              for (int a = 0; a < 8; ++a)
#pragma unroll
                for (int cc = 0; cc < 3; ++cc)
                  dold += cell_old[(size_t)(3 * a + cc) * n + t] * divc[a][cc];
            }
          for (int i = 0; i < 30; ++i)
            rhs_out[(size_t)i * n + t] = 0.;
          rhs_out[(size_t)30 * n + t] = c0 * pold * vol + alpha * dold * jc;
        }
    }

    // ---- assemble_system in 3-D: old-solution gather, traction term, constraints, scatter ----------------------------
    // cell_old[i][t] = old[cell_dofs[i][first + t]]
    __global__ void
    k_gather_old_3d(const int64_t n_cells, const int64_t first, const int64_t n, const int32_t *__restrict__ cdofs /* [31][n_cells] */,
                    const double *__restrict__ old, double *__restrict__ cell_old /* [31][n] */)
    {
      const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
      if (e >= n * NL3)
        return;
      const int64_t t = e % n;
      const int     i = (int)(e / n);
      cell_old[e]     = old[cdofs[(size_t)i * n_cells + first + t]];
    }

    // int_f N_v dS over face f of the hexahedron with QGauss<2>(3) (fe_face_values.JxW, CGQ1:1053, 1569-1580)
    __device__ double
    face_shape_integral(const Hex &H, const int f, const int v)
    {
      const int d = f >> 1, a1 = (d + 1) % 3, a2 = (d + 2) % 3;
      double    sum = 0.;
      for (int q2 = 0; q2 < 3; ++q2)
        for (int q1 = 0; q1 < 3; ++q1)
          {
            double r[3];
            r[d]  = (f & 1) ? 1. : 0.;
            r[a1] = kG3x[q1];
            r[a2] = kG3x[q2];
            double J[3][3], dN[8][3];
            H.jac(r[0], r[1], r[2], J, dN);
            const double c0 = J[1][a1] * J[2][a2] - J[2][a1] * J[1][a2], c1 = J[2][a1] * J[0][a2] - J[0][a1] * J[2][a2],
                         c2 = J[0][a1] * J[1][a2] - J[1][a1] * J[0][a2];
            const double Nv = ((v & 1) ? r[0] : 1. - r[0]) * (((v >> 1) & 1) ? r[1] : 1. - r[1]) * ((v >> 2) ? r[2] : 1. - r[2]);
            sum += Nv * sqrt(c0 * c0 + c1 * c1 + c2 * c2) * kG3w[q1] * kG3w[q2];
          }
      return sum;
    }

    // One thread per (cell, local row i), cells fastest (the dense blocks are entry-major: coalesced).
    // AffineConstraints::distribute_local_to_global with homogeneous constraints (CGQ1:1697-1699; deal.II
    // [deal.II-memory]): unconstrained row i receives its unconstrained columns and b_i; a constrained dof receives
    // |A_ii| on the diagonal (the average of the |diagonal| of the block when A_ii = 0) and nothing else.
    __global__ void
    k_scatter_3d(const int64_t n_cells, const int64_t first, const int64_t n, const double *__restrict__ x, const double *__restrict__ y,
                 const double *__restrict__ z, const uint32_t *__restrict__ cv, const int32_t *__restrict__ cdofs,
                 const uint32_t *__restrict__ con_mask, const uint8_t *__restrict__ neu_mask, const double t0, const double t1,
                 const double t2, const double *__restrict__ A /* [31*31][n] */, const double *__restrict__ b /* [31][n] */,
                 const int64_t *__restrict__ rowptr, const int32_t *__restrict__ colidx, double *__restrict__ values,
                 double *__restrict__ rhs, int *__restrict__ err)
    {
      const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
      if (e >= n * NL3)
        return;
      const int64_t  t = e % n, c = first + t;
      const int      i = (int)(e / n);
      const uint32_t m  = con_mask[c];
      const int32_t  gi = cdofs[(size_t)i * n_cells + c];
      const int64_t  r0 = rowptr[gi], r1 = rowptr[gi + 1];
      auto           find = [&](const int32_t col) -> int64_t {
        int64_t lo = r0, hi = r1;
        while (lo < hi)
          {
            const int64_t mid = (lo + hi) >> 1;
            if (colidx[mid] < col)
              lo = mid + 1;
            else
              hi = mid;
          }
        return (lo < r1 && colidx[lo] == col) ? lo : -1;
      };
      if ((m >> i) & 1u)
        {
          double d = fabs(A[(size_t)(i * NL3 + i) * n + t]);
          if (d == 0.)
            {
              for (int k = 0; k < NL3; ++k)
                d += fabs(A[(size_t)(k * NL3 + k) * n + t]);
              d /= NL3;
            }
          const int64_t pos = find(gi);
          if (pos < 0)
            atomicExch(err, 1);
          else
            atomicAdd(values + pos, d);
          return;
        }
      for (int j = 0; j < NL3; ++j)
        {
          if ((m >> j) & 1u)
            continue;
          const double v = A[(size_t)(i * NL3 + j) * n + t];
          const int64_t pos = find(cdofs[(size_t)j * n_cells + c]);
          if (pos < 0)
            atomicExch(err, 1);
          else
            atomicAdd(values + pos, v); // structural zeros are added as well: the reference adds every local entry
        }
      double bi = b[(size_t)i * n + t];
      const uint8_t nm = neu_mask[c];
      if (nm != 0 && i < 24)
        {
          // traction term: phi_i = N_v e_k, local_rhs(i) += t_k int_f N_v dS on the faces that contain vertex v
          const int    v = i / 3, k = i % 3;
          const double tk = k == 0 ? t0 : (k == 1 ? t1 : t2);
          if (tk != 0.)
            {
              Hex H;
              for (int w = 0; w < 8; ++w)
                {
                  const uint32_t id = cv[(size_t)w * n_cells + c];
                  H.vx[w]           = x[id];
                  H.vy[w]           = y[id];
                  H.vz[w]           = z[id];
                }
              for (int f = 0; f < 6; ++f)
                {
                  const int  d  = f >> 1;
                  const bool on = (((v >> d) & 1) == (f & 1));
                  if (((nm >> f) & 1) && on)
                    bi += tk * face_shape_integral(H, f, v);
                }
            }
        }
      if (bi != 0.)
        atomicAdd(rhs + gi, bi);
    }
  } // namespace

  cudaError_t
  launch_gather_old_3d(const int64_t n_cells, const int64_t first, const int64_t n, const int32_t *cdofs, const double *old,
                       double *cell_old, cudaStream_t s)
  {
    if (n <= 0)
      return cudaSuccess;
    k_gather_old_3d<<<(unsigned)((n * NL3 + 255) / 256), 256, 0, s>>>(n_cells, first, n, cdofs, old, cell_old);
    return cudaGetLastError();
  }

  cudaError_t
  launch_scatter_3d(const int64_t n_cells, const int64_t first, const int64_t n, const double *x, const double *y, const double *z,
                    const uint32_t *cv, const int32_t *cdofs, const uint32_t *con_mask, const uint8_t *neu_mask,
                    const double traction[3], const double *A, const double *b, const int64_t *rowptr, const int32_t *colidx,
                    double *values, double *rhs, int *err, cudaStream_t s)
  {
    if (n <= 0)
      return cudaSuccess;
    k_scatter_3d<<<(unsigned)((n * NL3 + 127) / 128), 128, 0, s>>>(n_cells, first, n, x, y, z, cv, cdofs, con_mask, neu_mask,
                                                                  traction[0], traction[1], traction[2], A, b, rowptr, colidx,
                                                                  values, rhs, err);
    return cudaGetLastError();
  }

  cudaError_t
  launch_local_matrices_3d(const int64_t n_cells, const int64_t first, const int64_t n, const double *x, const double *y,
                           const double *z, const uint32_t *cv, const double *kcell, const double k_scalar, const double lambda,
                           const double mu, const double alpha, const double c0, const double dt, const double *cell_old, double *out,
                           double *rhs_out, cudaStream_t s)
  {
    if (n <= 0)
      return cudaSuccess;
    k_local_matrices_3d<<<(unsigned)((n + 63) / 64), 64, 0, s>>>(n_cells, first, n, x, y, z, cv, kcell, k_scalar, lambda, mu, alpha, c0,
                                                                 dt, cell_old, out, rhs_out);
    return cudaGetLastError();
  }
} // namespace pe
