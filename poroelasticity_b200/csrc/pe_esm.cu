// WG(Q2,Q2;RT[2]) Darcy element of EltStiffMat.cc (WGDarcyEquation<2>): cell kernel, scatter, boundary
// values.  Replaces EltStiffMat.cc:397-576 (RT Gram matrix + gauss_jordan 413-429, F 437-477, C = F M^-1
// 478, the q,i,j,k,l loop 483-502, rhs 505-513, distribute_local_to_global 551-552,
// interpolate_boundary_values + apply_boundary_values 565-576).
//
// Lean form.  With the contravariant Piola map w = J w^/det J both parts of the weak-gradient right-hand
// side are independent of the cell geometry:  int phi div w = int^ phi^ div^ w^  and  int_f phi w.n =
// int_f^ phi^ w^.n^, so F (21 x 24) is ONE constant matrix for every quadrilateral.  For K = k I the
// K-weighted Gram matrix is k M, hence  A = k F M^-1 F^T = k Y^T Y  with  L L^T = M (Cholesky), L Y = F^T.
// The RT[2] space is spanned by shifted Legendre products (the local matrix depends on the space only,
// not on its basis; on a parallelogram M is block diagonal and close to diagonal).
// One warp builds one cell with M, L and Y in registers (rows / columns exchanged by warp shuffles); ~45 kflop
// per cell, the only compute-bound kernel of the path (AI ~ 42 flop/B against 3.6 KB written).
#include "pe_internal.h"
#include "pe_kernels.h"

#include <cmath>
#include <vector>

namespace pe
{
  namespace
  {
    constexpr int kRt = 24, kLoc = 21, kQ = 16;

    struct EsmTables
    {
      double F[kLoc * kRt];   // weak-gradient right-hand side, constant
      double px[kQ * 12];     // x-carriers P_a(xi) P_b(eta), a = k % 4, b = k / 4, at the 16 Gauss points
      double py[kQ * 12];     // y-carriers, a = k % 3, b = k / 3
      double pv[kQ * 9];      // FE_DGQ(2) values (Lagrange at 0, 1/2, 1; x fastest)
      double qx[kQ], qy[kQ], qw[kQ];
    };
    __constant__ EsmTables cT;
    // the same tables in global memory: lane-dependent addresses serialise in the constant cache (one request per
    // distinct address), which bounded BOTH cell kernels at 8.8e7 cells/s before the tensor-core version read them from here
    // (constant memory and allocations belong to one device: both copies are kept per device, a process may hold
    //  handles on several GPUs)
    constexpr int kMaxDev = 64;
    EsmTables    *dT[kMaxDev] = {};

    void
    gauss4(double x[4], double w[4])
    {
      // QGauss<1>(4) on [0,1]
      const double a = std::sqrt(3. / 7. - 2. / 7. * std::sqrt(6. / 5.)), b = std::sqrt(3. / 7. + 2. / 7. * std::sqrt(6. / 5.));
      const double wa = (18. + std::sqrt(30.)) / 36., wb = (18. - std::sqrt(30.)) / 36.;
      x[0] = .5 - .5 * b;
      x[1] = .5 - .5 * a;
      x[2] = .5 + .5 * a;
      x[3] = .5 + .5 * b;
      w[0] = w[3] = .5 * wb;
      w[1] = w[2] = .5 * wa;
    }
    void
    legendre01(const double x, double P[4], double dP[4])
    {
      const double s = 2. * x - 1.;
      P[0]  = 1.;
      P[1]  = s;
      P[2]  = .5 * (3. * s * s - 1.);
      P[3]  = .5 * (5. * s * s * s - 3. * s);
      dP[0] = 0.;
      dP[1] = 2.;
      dP[2] = 6. * s;
      dP[3] = 15. * s * s - 3.;
    }
    void
    lagrange3(const double x, double L[3])
    {
      L[0] = 2. * (x - .5) * (x - 1.);
      L[1] = -4. * x * (x - 1.);
      L[2] = 2. * x * (x - .5);
    }
    // reference RT[2] basis: k < 12: (P_a P_b, 0), a = k%4, b = k/4;  k >= 12: (0, P_a P_b), a = (k-12)%3, b = (k-12)/3
    void
    rt2_ref(const double xi, const double eta, double w[kRt][2], double dv[kRt])
    {
      double Px[4], dPx[4], Py[4], dPy[4];
      legendre01(xi, Px, dPx);
      legendre01(eta, Py, dPy);
      for (int k = 0; k < 12; ++k)
        {
          w[k][0] = Px[k % 4] * Py[k / 4];
          w[k][1] = 0.;
          dv[k]   = dPx[k % 4] * Py[k / 4];
          w[12 + k][0] = 0.;
          w[12 + k][1] = Px[k % 3] * Py[k / 3];
          dv[12 + k]   = Px[k % 3] * dPy[k / 3];
        }
    }
    bool tables_ready[kMaxDev] = {};
    int
    current_device()
    {
      int dev = 0;
      cudaGetDevice(&dev);
      return dev >= 0 && dev < kMaxDev ? dev : 0;
    }
    cudaError_t
    upload_tables()
    {
      const int dev = current_device();
      if (tables_ready[dev])
        return cudaSuccess;
      static EsmTables T;
      double           gx[4], gw[4];
      gauss4(gx, gw);
      for (int i = 0; i < kLoc * kRt; ++i)
        T.F[i] = 0.;
      for (int qj = 0; qj < 4; ++qj)
        for (int qi = 0; qi < 4; ++qi)
          {
            const int    q  = qj * 4 + qi;
            const double xi = gx[qi], eta = gx[qj], wq = gw[qi] * gw[qj];
            T.qx[q] = xi;
            T.qy[q] = eta;
            T.qw[q] = wq;
            double w[kRt][2], dv[kRt], Lx[3], Ly[3];
            rt2_ref(xi, eta, w, dv);
            lagrange3(xi, Lx);
            lagrange3(eta, Ly);
            for (int k = 0; k < 12; ++k)
              {
                T.px[q * 12 + k] = w[k][0];
                T.py[q * 12 + k] = w[12 + k][1];
              }
            for (int j = 0; j < 3; ++j)
              for (int i = 0; i < 3; ++i)
                {
                  T.pv[q * 9 + 3 * j + i] = Lx[i] * Ly[j];
                  // interior part of F: - int phi_i div w_k                          ESM:437-452
                  for (int k = 0; k < kRt; ++k)
                    T.F[(12 + 3 * j + i) * kRt + k] -= Lx[i] * Ly[j] * dv[k] * wq;
                }
          }
      // face part of F: + int_f phi_i (w_k . n)                                       ESM:454-476
      for (int f = 0; f < 4; ++f)
        for (int q = 0; q < 4; ++q)
          {
            const double t = gx[q], xi = f == 0 ? 0. : (f == 1 ? 1. : t), eta = f == 2 ? 0. : (f == 3 ? 1. : t);
            const double nx = f == 0 ? -1. : (f == 1 ? 1. : 0.), ny = f == 2 ? -1. : (f == 3 ? 1. : 0.);
            double       w[kRt][2], dv[kRt], L[3];
            rt2_ref(xi, eta, w, dv);
            lagrange3(t, L);
            for (int i = 0; i < 3; ++i)
              for (int k = 0; k < kRt; ++k)
                T.F[(3 * f + i) * kRt + k] += L[i] * (w[k][0] * nx + w[k][1] * ny) * gw[q];
          }
      cudaError_t e = cudaMemcpyToSymbol(cT, &T, sizeof(T));
      if (e == cudaSuccess && !dT[dev])
        e = cudaMalloc(&dT[dev], sizeof(T));
      if (e == cudaSuccess)
        e = cudaMemcpy(dT[dev], &T, sizeof(T), cudaMemcpyHostToDevice);
      tables_ready[dev] = e == cudaSuccess;
      return e;
    }

    constexpr int kWarps = 4;

    struct EsmArgs
    {
      int64_t         n_asm, first, n;
      const double   *vx, *vy, *kcell;
      double          k_scalar;
      const uint32_t *cv;
      const int32_t  *cdofs;
      const int64_t  *rowptr;
      const uint8_t  *off;
      double         *values, *rhs; // assembly mode
      double         *out, *rhs_out; // dense mode: out[n][21][21], rhs_out[n][21]
      int             case_id;
      const EsmTables *tab;          // tables in global memory (tensor-core version)
    };

    // One warp per cell, everything but the final 21 x 21 block in REGISTERS: lane i holds row i of the Gram
    // matrix M (later of its Cholesky factor L), lane c holds column c of Y = L^{-1} F^T, lane i accumulates row
    // i of A = k Y^T Y; rows / columns travel between lanes by warp shuffles with compile-time register
    // indices (all loops over matrix indices are fully unrolled).  The first version kept M and Y in shared
    // memory and was bound by the latency of ~3 dependent shared-memory round trips per update (ncu: FP64 pipe
    // 19 %, 3.98e7 cells/s).
    template <bool DENSE>
    __global__ void __launch_bounds__(kWarps * 32)
    k_esm_cell(const EsmArgs a)
    {
      __shared__ double  sG[kWarps][kQ * 4];       // per Gauss point: g00, g01, g11 (J^T J w / det)
      // per warp: the Cholesky factor L (24 x 24, later overwritten by the finished 21 x 21 block) and Y (24 x 21);
      // rows of L / columns of Y are read back as BROADCASTS (one address per warp), two doubles per LDS.128
      __shared__ __align__(16) double sL[kWarps][kRt * kRt];
      __shared__ __align__(16) double sYs[kWarps][kRt * (kLoc + 1)];
      __shared__ int64_t sRow[kWarps][kLoc];       // CSR row starts of the 21 local dofs
      // reference tables: constant memory serialises lane-dependent addresses, shared memory does not
      __shared__ __align__(16) double sPx[kQ * 12], sPy[kQ * 12];
      __shared__ double               sFt[kRt * kLoc];
      for (int e = threadIdx.x; e < kQ * 12; e += blockDim.x)
        {
          sPx[e] = cT.px[e];
          sPy[e] = cT.py[e];
        }
      for (int e = threadIdx.x; e < kRt * kLoc; e += blockDim.x)
        sFt[e] = cT.F[(e % kLoc) * kRt + e / kLoc]; // F^T (24 x 21)
      __syncthreads();
      const int      warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
      const int64_t  t    = (int64_t)blockIdx.x * kWarps + warp;
      if (t >= a.n)
        return;
      constexpr unsigned FULL = 0xffffffffu;
      const int64_t      c    = a.first + t;
      double            *G    = sG[warp];
      // ---- geometry at the Gauss points (lanes 0..15)
      double vx[4], vy[4];
#pragma unroll
      for (int v = 0; v < 4; ++v)
        {
          const uint32_t vi = a.cv[(size_t)v * a.n_asm + c];
          vx[v]             = __ldg(a.vx + vi);
          vy[v]             = __ldg(a.vy + vi);
        }
      Geom geo;
      geo.set(vx, vy);
      double fq = 0.; // f(x_q) JxW_q of this lane's Gauss point
      if (lane < kQ)
        {
          double j00, j01, j10, j11;
          geo.jac(cT.qx[lane], cT.qy[lane], j00, j01, j10, j11);
          const double det = j00 * j11 - j01 * j10, r = cT.qw[lane] / det;
          G[lane * 4 + 0]  = r * (j00 * j00 + j10 * j10);
          G[lane * 4 + 1]  = r * (j00 * j01 + j10 * j11);
          G[lane * 4 + 2]  = r * (j01 * j01 + j11 * j11);
          if (a.case_id == PE_CASE_DARCY_COSCOS)
            {
              double px, py;
              geo.point(cT.qx[lane], cT.qy[lane], px, py);
              fq = 2. * M_PI * M_PI * cospi(px) * cospi(py) * det * cT.qw[lane]; // RightHandSide, ESM:167
            }
        }
      if (!DENSE && lane < kLoc)
        sRow[warp][lane] = a.rowptr[a.cdofs[(size_t)lane * a.n_asm + c]];
      __syncwarp();
      // ---- row `lane` of the Gram matrix of the Piola-mapped basis                              ESM:413-426
      //      M_ij = sum_q (J w^_i).(J w^_j) w_q / det ; lanes >= 24 carry a harmless copy of row 23
      const int  li   = lane < kRt ? lane : kRt - 1;
      const bool is_x = li < 12;
      double     M[kRt];
#pragma unroll
      for (int j = 0; j < kRt; ++j)
        M[j] = 0.;
#pragma unroll 2
      for (int q = 0; q < kQ; ++q)
        {
          const double mine = is_x ? sPx[q * 12 + li] : sPy[q * 12 + li - 12];
          const double wx = mine * (is_x ? G[q * 4 + 0] : G[q * 4 + 1]), wy = mine * (is_x ? G[q * 4 + 1] : G[q * 4 + 2]);
#pragma unroll
          for (int j = 0; j < 12; j += 2)
            {
              const double2 ox = *reinterpret_cast<const double2 *>(&sPx[q * 12 + j]);
              const double2 oy = *reinterpret_cast<const double2 *>(&sPy[q * 12 + j]);
              M[j] += wx * ox.x;
              M[j + 1] += wx * ox.y;
              M[12 + j] += wy * oy.x;
              M[12 + j + 1] += wy * oy.y;
            }
        }
      // ---- Cholesky M = L L^T, right-looking; replaces gauss_jordan (ESM:429).  After step k: M[k] = L_ik
      double invd[kRt];
#pragma unroll
      for (int k = 0; k < kRt; ++k)
        {
          const double dkk = __shfl_sync(FULL, M[k], k);
          invd[k]          = rsqrt(dkk);
          const double lik = M[k] * invd[k]; // L_ik for lanes i > k, L_kk = sqrt(M_kk) on lane k
          M[k]             = lik;
#pragma unroll
          for (int j = k + 1; j < kRt; ++j)
            {
              const double ljk = __shfl_sync(FULL, lik, j);
              M[j] -= lik * ljk; // meaningful for i >= j; the upper triangle is never read
            }
        }
      // ---- L Y = F^T: lane = column of Y (one local dof); the rows of L come back from shared memory as
      //      broadcasts (the shuffle version of this loop and of Y^T Y cost ~1,600 of the 2,200 SHFL per cell)
      double *Ls = sL[warp], *Ys = sYs[warp];
      if (lane < kRt)
        {
#pragma unroll
          for (int q = 0; q < kRt; ++q)
            Ls[lane * kRt + q] = M[q];
        }
      __syncwarp();
      const int lc = lane < kLoc ? lane : kLoc - 1;
      double    Y[kRt];
#pragma unroll
      for (int r = 0; r < kRt; ++r)
        {
          double sacc = sFt[r * kLoc + lc];
#pragma unroll
          for (int q = 0; q + 1 < r; q += 2)
            {
              const double2 l2 = *reinterpret_cast<const double2 *>(&Ls[r * kRt + q]);
              sacc -= l2.x * Y[q] + l2.y * Y[q + 1];
            }
          if (r & 1)
            sacc -= Ls[r * kRt + r - 1] * Y[r - 1];
          Y[r] = sacc * invd[r];
          if (lane < kLoc)
            Ys[r * (kLoc + 1) + lane] = Y[r];
        }
      __syncwarp();
      // ---- A = k Y^T Y  (C M_K C^T of ESM:478-502): lane i builds row i, columns of Y broadcast from shared memory
      const double kperm = a.kcell ? a.kcell[c] : a.k_scalar;
      double       Arow[kLoc];
#pragma unroll
      for (int j = 0; j < kLoc; ++j)
        Arow[j] = 0.;
#pragma unroll
      for (int r = 0; r < kRt; ++r)
        {
          const double yr = Y[r];
#pragma unroll
          for (int j = 0; j + 1 < kLoc; j += 2)
            {
              const double2 y2 = *reinterpret_cast<const double2 *>(&Ys[r * (kLoc + 1) + j]);
              Arow[j] += yr * y2.x;
              Arow[j + 1] += yr * y2.y;
            }
          Arow[kLoc - 1] += yr * Ys[r * (kLoc + 1) + kLoc - 1];
        }
      __syncwarp();
      double *A = Ls; // L is dead: reuse its storage for the finished block (coalesced output below)
      if (lane < kLoc)
        {
#pragma unroll
          for (int j = 0; j < kLoc; ++j)
            A[lane * kLoc + j] = kperm * Arow[j];
        }
      // ---- right-hand side (ESM:505-513)
      double bi = 0.; // lanes 0..8: b of interior dof `lane`
      for (int q = 0; q < kQ; ++q)
        {
          const double f = __shfl_sync(FULL, fq, q);
          if (lane < 9)
            bi += f * cT.pv[q * 9 + lane];
        }
      __syncwarp();
      for (int e = lane; e < kLoc * kLoc; e += 32)
        {
          if (DENSE)
            a.out[(size_t)t * kLoc * kLoc + e] = A[e];
          else
            {
              const uint32_t o = a.off[((size_t)(e >> 2) * a.n_asm + c) * 4 + (e & 3)];
              atomicAdd(a.values + sRow[warp][e / kLoc] + (o & 127u), A[e]);
            }
        }
      const double bsh = __shfl_sync(FULL, bi, lane >= 12 ? lane - 12 : 0);
      if (DENSE)
        {
          if (a.rhs_out && lane < kLoc)
            a.rhs_out[(size_t)t * kLoc + lane] = lane >= 12 ? bsh : 0.;
        }
      else if (lane < 9)
        {
          const int32_t r2 = a.cdofs[(size_t)(12 + lane) * a.n_asm + c];
          if (bi != 0.)
            atomicAdd(a.rhs + r2, bi);
        }
    }

    // ---- FP64 tensor-core version ------------------------------------------------------------------------------
    // D (8 x 8) += A (8 x 4, row) B (4 x 8, col), mma.sync m8n8k4 f64: with g = lane / 4, k = lane % 4 a lane holds
    // A[g][k], B[k][g] and D[g][2 k], D[g][2 k + 1].
    __device__ __forceinline__ void
    dmma(double &d0, double &d1, const double a, const double b)
    {
      asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
    }

    constexpr int kMs = kRt + 1; // row stride of the Gram matrix in shared memory: lane-per-row reads hit 16 distinct banks
    constexpr int kYs = 28;      // row stride of Y: the 4 x 8 fragment loads of a half warp fall into distinct banks

    // One warp per cell, the two GEMM-shaped parts on the FP64 tensor cores:
    //   Gram matrix   M = U W^T, U_i(q,b) = w_i(q) G_{type(i) b}(q), W_j(q,b) = w_j(q) [type(j) = b]   (24 x 32 x 24; the
    //                 k-steps in which W vanishes for a whole column tile are skipped: 32 DMMA for the 6 lower tiles);
    //                 the B fragments and the carrier part of the A fragments are per-lane CONSTANTS (12 registers),
    //                 the geometry enters through 12 values G(q) per lane
    //   Cholesky      lane i = row i in registers, as in the first version (shuffle broadcasts)
    //   L Y = F^T     column-oriented: as soon as Y_q is known all later rows are updated (24 independent FMAs) --
    //                 the row-oriented loop of the first version was one dependent chain of 276 FMAs per cell
    //   A = k Y^T Y   fragments of Y serve as A and B operand alike: 18 loads, 36 DMMA for the 6 lower tiles
    // (ESM:413-429 Gram + inverse, 478 C = F M^-1, 483-502 the q,i,j,k,l loop, 505-513 rhs).
    template <bool DENSE>
    __global__ void __launch_bounds__(kWarps * 32, 4)
    k_esm_cell_mma(const EsmArgs a)
    {
      __shared__ double sG[kWarps][kQ * 3];
      __shared__ __align__(16) double sM[kWarps][kRt * kMs];  // Gram matrix (stride 25), then L^T (stride 24), then A (21 x 21)
      __shared__ __align__(16) double sYs[kWarps][kRt * kYs]; // Y, columns 21..23 zero
      __shared__ int64_t sRow[DENSE ? 1 : kWarps][kLoc];
      __shared__ double  sFt[kRt * kLoc], sPv[kQ * 9];
      const EsmTables *__restrict__ T = a.tab;
      for (int e = threadIdx.x; e < kRt * kLoc; e += blockDim.x)
        sFt[(e % kRt) * kLoc + e / kRt] = T->F[e]; // F^T (24 x 21); coalesced read
      for (int e = threadIdx.x; e < kQ * 9; e += blockDim.x)
        sPv[e] = T->pv[e];
      const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, k4 = lane & 3;
      // per-lane constants, loaded once per (persistent) warp:
      // carrier values of this lane's fragment entries cv[t][sq] = w_{8 t + g}(q = 4 sq + k4); its Gauss point
      double cv[3][4];
#pragma unroll
      for (int t = 0; t < 3; ++t)
#pragma unroll
        for (int sq = 0; sq < 4; ++sq)
          {
            const int i = 8 * t + g, q = 4 * sq + k4;
            cv[t][sq]   = i < 12 ? T->px[q * 12 + i] : T->py[q * 12 + i - 12];
          }
      const double qx = T->qx[lane & (kQ - 1)], qy = T->qy[lane & (kQ - 1)], qw = T->qw[lane & (kQ - 1)];
      __syncthreads();
      constexpr unsigned FULL = 0xffffffffu;
      double            *G    = sG[warp];
      // persistent warps: the grid is a few CTAs per SM, every warp walks over the cells with the grid stride
      for (int64_t t = (int64_t)blockIdx.x * kWarps + warp; t < a.n; t += (int64_t)gridDim.x * kWarps)
      {
      const int64_t c = a.first + t;
      double vx[4], vy[4];
#pragma unroll
      for (int v = 0; v < 4; ++v)
        {
          const uint32_t vi = a.cv[(size_t)v * a.n_asm + c];
          vx[v]             = __ldg(a.vx + vi);
          vy[v]             = __ldg(a.vy + vi);
        }
      Geom geo;
      geo.set(vx, vy);
      double fq = 0.;
      if (lane < kQ)
        {
          double j00, j01, j10, j11;
          geo.jac(qx, qy, j00, j01, j10, j11);
          const double det = j00 * j11 - j01 * j10, r = qw / det;
          G[lane * 3 + 0]  = r * (j00 * j00 + j10 * j10);
          G[lane * 3 + 1]  = r * (j00 * j01 + j10 * j11);
          G[lane * 3 + 2]  = r * (j01 * j01 + j11 * j11);
          if (a.case_id == PE_CASE_DARCY_COSCOS)
            {
              double px, py;
              geo.point(qx, qy, px, py);
              fq = 2. * M_PI * M_PI * cospi(px) * cospi(py) * det * qw; // RightHandSide, ESM:167
            }
        }
      if (!DENSE && lane < kLoc)
        sRow[warp][lane] = a.rowptr[a.cdofs[(size_t)lane * a.n_asm + c]];
      __syncwarp();
      // ---- Gram matrix on the tensor cores                                                       ESM:413-426
      double *Ms = sM[warp];
      {
        // metric at this lane's four Gauss points q = 4 sq + k4: (g00, g01, g11)
        double Gq[4][3];
#pragma unroll
        for (int sq = 0; sq < 4; ++sq)
#pragma unroll
          for (int e = 0; e < 3; ++e)
            Gq[sq][e] = G[(4 * sq + k4) * 3 + e];
        // basis functions 0..11 carry the x component, 12..23 the y component: tile 0 all x, tile 1 mixed, tile 2 all y
        const bool isx[3] = {true, g < 4, false};
        // A fragments: U_i(q, b) for b = x (k-steps 0..3) and b = y (k-steps 4..7)
        double af[3][8];
#pragma unroll
        for (int ti = 0; ti < 3; ++ti)
#pragma unroll
          for (int s = 0; s < 8; ++s)
            {
              const int b = s >> 2, sq = s & 3;
              af[ti][s]   = cv[ti][sq] * (isx[ti] ? Gq[sq][b] : Gq[sq][b + 1]);
            }
        // lower tiles; a column tile of pure x (y) functions only sees the k-steps of b = x (y)
#pragma unroll
        for (int tj = 0; tj < 3; ++tj)
#pragma unroll
          for (int ti = tj; ti < 3; ++ti)
            {
              double d0 = 0., d1 = 0.;
#pragma unroll
              for (int s = 0; s < 8; ++s)
                {
                  const int b = s >> 2, sq = s & 3;
                  if ((tj == 0 && b == 1) || (tj == 2 && b == 0))
                    continue;
                  const bool   match = b == 0 ? isx[tj] : !isx[tj];
                  const double bf    = match ? cv[tj][sq] : 0.;
                  dmma(d0, d1, af[ti][s], bf);
                }
              double *o = &Ms[(8 * ti + g) * kMs + 8 * tj + 2 * k4];
              o[0]      = d0;
              o[1]      = d1;
            }
      }
      __syncwarp();
      // ---- Cholesky M = L L^T: lane i holds row i (lanes >= 24 a harmless copy of row 23); the entries above the
      //      diagonal tiles were never written and are never used
      const int li = lane < kRt ? lane : kRt - 1;
      double    M[kRt];
#pragma unroll
      for (int j = 0; j < kRt; ++j)
        M[j] = Ms[li * kMs + j];
      __syncwarp();
      // Column k of L goes to shared memory as row k of L^T (the layout the substitution below wants) and comes back
      // as broadcast reads, two entries per LDS.128: 24 STS + 138 LDS instead of the 552 SHFL of the first version.
      // The pivot travels on its own: every lane keeps d = M_ii - sum_{m<k} L_im^2 of its row, so that the next pivot
      // only waits for rsqrt -> multiply -> one FMA and not for the shared-memory round trip of the column.
      double *Lt = Ms; // every lane holds its row of M in registers by now
      double  invd[kRt];
      double  d = Ms[li * kMs + li];
      __syncwarp();
#pragma unroll
      for (int k = 0; k < kRt; ++k)
        {
          const double dkk = __shfl_sync(FULL, d, k);
          invd[k]          = rsqrt(dkk);
          const double lik = M[k] * invd[k]; // L_ik for lanes i > k, L_kk = sqrt(M_kk) on lane k
          d -= lik * lik;
          if (lane < kRt)
            Lt[k * kRt + lane] = lik;
          __syncwarp();
          if ((k + 1) & 1) // k + 1 odd (and < 24): a single entry, the pairs start at the next even row
            M[k + 1] -= lik * Lt[k * kRt + k + 1];
#pragma unroll
          for (int j = k + 1 + ((k + 1) & 1); j + 1 < kRt; j += 2)
            {
              const double2 l2 = *reinterpret_cast<const double2 *>(&Lt[k * kRt + j]);
              M[j] -= lik * l2.x; // meaningful for i >= j; the upper triangle is never read
              M[j + 1] -= lik * l2.y;
            }
        }
      __syncwarp();
      // ---- L Y = F^T, column-oriented; lane = column of Y (one local dof)
      double *Ys = sYs[warp];
      {
        const int lc = lane < kLoc ? lane : kLoc - 1;
        double    acc[kRt];
#pragma unroll
        for (int r = 0; r < kRt; ++r)
          acc[r] = sFt[r * kLoc + lc];
#pragma unroll
        for (int q = 0; q < kRt; ++q)
          {
            const double yq = acc[q] * invd[q];
            if (lane < kRt)
              Ys[q * kYs + lane] = lane < kLoc ? yq : 0.;
            // rows r > q: pairs (r, r + 1) with r even come back as one LDS.128
            if ((q + 1) & 1)
              acc[q + 1] -= Lt[q * kRt + q + 1] * yq;
#pragma unroll
            for (int r = q + 1 + ((q + 1) & 1); r + 1 < kRt; r += 2)
              {
                const double2 l2 = *reinterpret_cast<const double2 *>(&Lt[q * kRt + r]);
                acc[r] -= l2.x * yq;
                acc[r + 1] -= l2.y * yq;
              }
          }
      }
      __syncwarp();
      // ---- A = k Y^T Y on the tensor cores                                                       ESM:478-502
      const double kperm = a.kcell ? a.kcell[c] : a.k_scalar;
      double       yf[3][6];
#pragma unroll
      for (int tt = 0; tt < 3; ++tt)
#pragma unroll
        for (int s = 0; s < 6; ++s)
          yf[tt][s] = Ys[(4 * s + k4) * kYs + 8 * tt + g];
      double *A = Ms; // L^T is dead
      __syncwarp();
#pragma unroll
      for (int ti = 0; ti < 3; ++ti)
#pragma unroll
        for (int tj = 0; tj <= ti; ++tj)
          {
            double d0 = 0., d1 = 0.;
#pragma unroll
            for (int s = 0; s < 6; ++s)
              dmma(d0, d1, yf[ti][s], yf[tj][s]);
            const int row = 8 * ti + g, col = 8 * tj + 2 * k4;
            d0 *= kperm;
            d1 *= kperm;
            if (row < kLoc)
              {
                if (col < kLoc)
                  {
                    A[row * kLoc + col] = d0;
                    if (ti != tj)
                      A[col * kLoc + row] = d0;
                  }
                if (col + 1 < kLoc)
                  {
                    A[row * kLoc + col + 1] = d1;
                    if (ti != tj)
                      A[(col + 1) * kLoc + row] = d1;
                  }
              }
          }
      // ---- right-hand side (ESM:505-513)
      double bi = 0.;
      for (int q = 0; q < kQ; ++q)
        {
          const double f = __shfl_sync(FULL, fq, q);
          if (lane < 9)
            bi += f * sPv[q * 9 + lane];
        }
      __syncwarp();
      for (int e = lane; e < kLoc * kLoc; e += 32)
        {
          if (DENSE)
            a.out[(size_t)t * kLoc * kLoc + e] = A[e];
          else
            {
              const uint32_t o = a.off[((size_t)(e >> 2) * a.n_asm + c) * 4 + (e & 3)];
              atomicAdd(a.values + sRow[warp][e / kLoc] + (o & 127u), A[e]);
            }
        }
      const double bsh = __shfl_sync(FULL, bi, lane >= 12 ? lane - 12 : 0);
      if (DENSE)
        {
          if (a.rhs_out && lane < kLoc)
            a.rhs_out[(size_t)t * kLoc + lane] = lane >= 12 ? bsh : 0.;
        }
      else if (lane < 9)
        {
          const int32_t r2 = a.cdofs[(size_t)(12 + lane) * a.n_asm + c];
          if (bi != 0.)
            atomicAdd(a.rhs + r2, bi);
        }
      __syncwarp(); // the next cell reuses this warp's shared-memory buffers
      }
    }

    // ---- MatrixTools::apply_boundary_values (ESM:572-575), two passes ------------------------------------
    // pass 1: eliminate the column of every boundary dof d:  rhs_r -= a_rd g_d (free rows r), a_rd = 0
    __global__ void
    k_esm_bv_columns(const int n, const int32_t *__restrict__ dofs, const double *__restrict__ vals,
                     const uint8_t *__restrict__ is_bnd, const int64_t *__restrict__ rowptr,
                     const int32_t *__restrict__ colidx, double *__restrict__ values, double *__restrict__ rhs)
    {
      const int t = blockIdx.x * blockDim.x + threadIdx.x;
      if (t >= n)
        return;
      const int32_t d = dofs[t];
      const double  g = vals[t];
      for (int64_t k = rowptr[d]; k < rowptr[d + 1]; ++k)
        {
          const int32_t r = colidx[k]; // symmetric pattern: the rows that couple to d
          if (r == d)
            continue;
          for (int64_t k2 = rowptr[r]; k2 < rowptr[r + 1]; ++k2)
            if (colidx[k2] == d)
              {
                const double v = values[k2];
                if (!is_bnd[r] && v != 0. && g != 0.)
                  atomicAdd(rhs + r, -v * g);
                values[k2] = 0.;
                break;
              }
        }
    }
    // pass 2: the rows of the boundary dofs: zero except the diagonal, rhs = diag g, solution = g
    __global__ void
    k_esm_bv_rows(const int n, const int32_t *__restrict__ dofs, const double *__restrict__ vals,
                  const int64_t *__restrict__ rowptr, const int32_t *__restrict__ colidx, double *__restrict__ values,
                  double *__restrict__ rhs, double *__restrict__ sol)
    {
      const int t = blockIdx.x * blockDim.x + threadIdx.x;
      if (t >= n)
        return;
      const int32_t d    = dofs[t];
      double        diag = 0.;
      for (int64_t k = rowptr[d]; k < rowptr[d + 1]; ++k)
        if (colidx[k] == d)
          diag = values[k];
        else
          values[k] = 0.;
      rhs[d] = diag * vals[t];
      sol[d] = vals[t];
    }

    // ---- postprocess()  EltStiffMat.cc:662-919 ------------------------------------------------------------
    // One thread per cell (a diagnostic, not a hot kernel): beta = -k M^-1 F^T p_cell are the RT[2] coefficients of the
    // Darcy velocity -K grad_w p_h (for K = k I the matrix D = M^-1 E of ESM:757-758 is k I), M by Cholesky in
    // thread-local memory.  e[0] += int |u_h - u|^2 (QGauss(4)),  e[1] += sum_f (|E|/|f|) int_f ((u_h - u).n)^2.
    __device__ __forceinline__ void
    d_legendre01(const double x, double P[4])
    {
      const double t = 2. * x - 1.;
      P[0] = 1.;
      P[1] = t;
      P[2] = .5 * (3. * t * t - 1.);
      P[3] = .5 * (5. * t * t * t - 3. * t);
    }
    // reference velocity sum_k beta_k w^_k at (xi, eta)
    __device__ __forceinline__ void
    d_rt2_combine(const double *beta, const double xi, const double eta, double &hx, double &hy)
    {
      double Px[4], Py[4];
      d_legendre01(xi, Px);
      d_legendre01(eta, Py);
      hx = hy = 0.;
      for (int k = 0; k < 12; ++k)
        {
          hx += beta[k] * Px[k % 4] * Py[k / 4];
          hy += beta[12 + k] * Px[k % 3] * Py[k / 3];
        }
    }
    __global__ void __launch_bounds__(64)
    k_esm_postprocess(const EsmArgs a, const double *__restrict__ sol, double *__restrict__ e2)
    {
      const int64_t c  = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
      double        ev = 0., ef = 0.;
      if (c < a.n_asm)
        {
          double vx[4], vy[4];
#pragma unroll
          for (int v = 0; v < 4; ++v)
            {
              const uint32_t vi = a.cv[(size_t)v * a.n_asm + c];
              vx[v]             = a.vx[vi];
              vy[v]             = a.vy[vi];
            }
          Geom G;
          G.set(vx, vy);
          const double kperm = a.kcell ? a.kcell[c] : a.k_scalar;
          double       g[kRt], M[kRt * kRt];
          for (int k = 0; k < kRt; ++k)
            g[k] = 0.;
          for (int i = 0; i < kLoc; ++i)
            {
              const double si = sol[a.cdofs[(size_t)i * a.n_asm + c]];
              for (int k = 0; k < kRt; ++k)
                g[k] += cT.F[i * kRt + k] * si;
            }
          for (int k = 0; k < kRt * kRt; ++k)
            M[k] = 0.;
          double area = 0.;
          for (int q = 0; q < kQ; ++q)
            {
              double j00, j01, j10, j11;
              G.jac(cT.qx[q], cT.qy[q], j00, j01, j10, j11);
              const double det = j00 * j11 - j01 * j10, r = cT.qw[q] / det;
              area += cT.qw[q] * det;
              // w_k = J w^_k / det: x-carriers use column 0 of J, y-carriers column 1
              const double a00 = (j00 * j00 + j10 * j10) * r, a01 = (j00 * j01 + j10 * j11) * r, a11 = (j01 * j01 + j11 * j11) * r;
              for (int i = 0; i < kRt; ++i)
                {
                  const double pi = i < 12 ? cT.px[q * 12 + i] : cT.py[q * 12 + i - 12];
                  for (int j = 0; j <= i; ++j)
                    {
                      const double pj = j < 12 ? cT.px[q * 12 + j] : cT.py[q * 12 + j - 12];
                      const double cc = i < 12 ? a00 : (j < 12 ? a01 : a11);
                      M[i * kRt + j] += pi * pj * cc;
                    }
                }
            }
          // Cholesky M = L L^T (lower triangle in place), then L L^T y = g
          for (int j = 0; j < kRt; ++j)
            {
              double d = M[j * kRt + j];
              for (int k = 0; k < j; ++k)
                d -= M[j * kRt + k] * M[j * kRt + k];
              d              = sqrt(d);
              M[j * kRt + j] = d;
              for (int i = j + 1; i < kRt; ++i)
                {
                  double t = M[i * kRt + j];
                  for (int k = 0; k < j; ++k)
                    t -= M[i * kRt + k] * M[j * kRt + k];
                  M[i * kRt + j] = t / d;
                }
            }
          for (int i = 0; i < kRt; ++i)
            {
              double t = g[i];
              for (int k = 0; k < i; ++k)
                t -= M[i * kRt + k] * g[k];
              g[i] = t / M[i * kRt + i];
            }
          for (int i = kRt - 1; i >= 0; --i)
            {
              double t = g[i];
              for (int k = i + 1; k < kRt; ++k)
                t -= M[k * kRt + i] * g[k];
              g[i] = t / M[i * kRt + i];
            }
          for (int k = 0; k < kRt; ++k)
            g[k] = -kperm * g[k]; // beta
          // velocity error, QGauss<2>(4)
          for (int q = 0; q < kQ; ++q)
            {
              double j00, j01, j10, j11, hx = 0., hy = 0., px, py, sx, cx, sy, cy;
              G.jac(cT.qx[q], cT.qy[q], j00, j01, j10, j11);
              const double det = j00 * j11 - j01 * j10;
              for (int k = 0; k < 12; ++k)
                {
                  hx += g[k] * cT.px[q * 12 + k];
                  hy += g[12 + k] * cT.py[q * 12 + k];
                }
              const double ux = (j00 * hx + j01 * hy) / det, uy = (j10 * hx + j11 * hy) / det;
              G.point(cT.qx[q], cT.qy[q], px, py);
              sincospi(px, &sx, &cx);
              sincospi(py, &sy, &cy);
              const double dx = ux - M_PI * sx * cy, dy = uy - M_PI * cx * sy; // Velocity::value, ESM:200-211
              ev += (dx * dx + dy * dy) * cT.qw[q] * det;
            }
          // normal flux error per face, QGauss<1>(4); face f: 0 x^=0, 1 x^=1, 2 y^=0, 3 y^=1
          const double gw0 = sqrt(cT.qw[0]);
          for (int f = 0; f < 4; ++f)
            {
              const int    v0 = f == 0 ? 0 : (f == 1 ? 1 : (f == 2 ? 0 : 2)), v1 = f == 0 ? 2 : (f == 1 ? 3 : (f == 2 ? 1 : 3));
              const double flen = hypot(vx[v1] - vx[v0], vy[v1] - vy[v0]);
              double       loc  = 0.;
              for (int q = 0; q < 4; ++q)
                {
                  const double t = cT.qx[q], w1 = cT.qw[q] / gw0; // 1-D Gauss point / weight
                  const double xi = f == 0 ? 0. : (f == 1 ? 1. : t), eta = f == 2 ? 0. : (f == 3 ? 1. : t);
                  double       j00, j01, j10, j11, hx, hy, px, py, sx, cx, sy, cy;
                  G.jac(xi, eta, j00, j01, j10, j11);
                  const double det = j00 * j11 - j01 * j10;
                  d_rt2_combine(g, xi, eta, hx, hy);
                  const double ux = (j00 * hx + j01 * hy) / det, uy = (j10 * hx + j11 * hy) / det;
                  // tangent = column of J along the face; outward normal = rotated tangent
                  const double tx = f < 2 ? j01 : j00, ty = f < 2 ? j11 : j10, tl = sqrt(tx * tx + ty * ty);
                  double       nx = ty / tl, ny = -tx / tl; // right of the tangent
                  if (f == 0 || f == 3)
                    {
                      nx = -nx;
                      ny = -ny;
                    }
                  G.point(xi, eta, px, py);
                  sincospi(px, &sx, &cx);
                  sincospi(py, &sy, &cy);
                  const double dn = (ux - M_PI * sx * cy) * nx + (uy - M_PI * cx * sy) * ny;
                  loc += dn * dn * tl * w1;
                }
              ef += loc / flen * area;
            }
        }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1)
        {
          ev += __shfl_xor_sync(0xffffffffu, ev, o);
          ef += __shfl_xor_sync(0xffffffffu, ef, o);
        }
      if ((threadIdx.x & 31) == 0)
        {
          atomicAdd(e2, ev);
          atomicAdd(e2 + 1, ef);
        }
    }
  } // namespace

  cudaError_t
  launch_esm_postprocess(const AsmArgs &a, const double *sol, double *e2, cudaStream_t s)
  {
    cudaError_t e = upload_tables();
    if (e != cudaSuccess)
      return e;
    cudaMemsetAsync(e2, 0, 2 * sizeof(double), s);
    if (a.n_asm == 0)
      return cudaSuccess;
    EsmArgs x{};
    x.n_asm    = a.n_asm;
    x.vx       = a.vx;
    x.vy       = a.vy;
    x.kcell    = a.kcell;
    x.k_scalar = a.k_scalar;
    x.cv       = a.cv;
    x.cdofs    = a.cdofs;
    k_esm_postprocess<<<(unsigned)((a.n_asm + 63) / 64), 64, 0, s>>>(x, sol, e2);
    return cudaGetLastError();
  }

  // persistent grid of the tensor-core kernel: 4 CTAs of 4 warps per SM (128 registers, 48 KB shared memory each)
  static unsigned
  mma_grid(const int64_t n_cells)
  {
    static int n_sm = 0;
    if (n_sm == 0)
      {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n_sm <= 0)
          n_sm = 148;
      }
    return (unsigned)std::min<int64_t>((n_cells + kWarps - 1) / kWarps, (int64_t)n_sm * 4);
  }

  cudaError_t
  launch_esm_assemble(const AsmArgs &a, const int case_id, cudaStream_t s, int *n_launches)
  {
    cudaError_t e = upload_tables();
    if (e != cudaSuccess || a.n_asm == 0)
      return e;
    EsmArgs x{};
    x.n_asm    = a.n_asm;
    x.first    = 0;
    x.n        = a.n_asm;
    x.vx       = a.vx;
    x.vy       = a.vy;
    x.kcell    = a.kcell;
    x.k_scalar = a.k_scalar;
    x.cv       = a.cv;
    x.cdofs    = a.cdofs;
    x.rowptr   = a.rowptr;
    x.off      = a.off;
    x.values   = a.values;
    x.rhs      = a.rhs;
    x.case_id  = case_id;
    x.tab = dT[current_device()];
    if (a.esm_variant == 0)
      k_esm_cell_mma<false><<<mma_grid(x.n), kWarps * 32, 0, s>>>(x);
    else
      k_esm_cell<false><<<(unsigned)((x.n + kWarps - 1) / kWarps), kWarps * 32, 0, s>>>(x);
    ++*n_launches;
    return cudaGetLastError();
  }

  cudaError_t
  launch_esm_local_matrices(const AsmArgs &a, const int case_id, const int64_t first, const int64_t n, double *out,
                            double *rhs_out, cudaStream_t s)
  {
    cudaError_t e = upload_tables();
    if (e != cudaSuccess || n == 0)
      return e;
    EsmArgs x{};
    x.n_asm    = a.n_asm;
    x.first    = first;
    x.n        = n;
    x.vx       = a.vx;
    x.vy       = a.vy;
    x.kcell    = a.kcell;
    x.k_scalar = a.k_scalar;
    x.cv       = a.cv;
    x.out      = out;
    x.rhs_out  = rhs_out;
    x.case_id  = case_id;
    x.tab = dT[current_device()];
    if (a.esm_variant == 0)
      k_esm_cell_mma<true><<<mma_grid(n), kWarps * 32, 0, s>>>(x);
    else
      k_esm_cell<true><<<(unsigned)((n + kWarps - 1) / kWarps), kWarps * 32, 0, s>>>(x);
    return cudaGetLastError();
  }

  cudaError_t
  launch_esm_boundary_values(const int n, const int32_t *dofs, const double *vals, const uint8_t *is_bnd,
                             const int64_t *rowptr, const int32_t *colidx, double *values, double *rhs, double *sol,
                             cudaStream_t s)
  {
    if (n == 0)
      return cudaSuccess;
    k_esm_bv_columns<<<(n + 127) / 128, 128, 0, s>>>(n, dofs, vals, is_bnd, rowptr, colidx, values, rhs);
    k_esm_bv_rows<<<(n + 127) / 128, 128, 0, s>>>(n, dofs, vals, rowptr, colidx, values, rhs, sol);
    return cudaGetLastError();
  }
} // namespace pe
