// Auxiliary-space geometric multigrid preconditioner for the Biot system (new work: the reference
// factorises with UMFPACK every step, PoroElasBR1WG_new.cc:1177-1187; BASELINE north_star asks for a
// preconditioned Krylov solver).
//
//   P^{-1} = theta D^{-1}                       (diagonal part: bubbles, p_int/p_face oscillations)
//          + Pi_u V_u Pi_u^T                    (vector Q1 elasticity V-cycle on the vertex grids)
//          + Pi_p V_p Pi_p^T                    (scalar Q1 reaction-diffusion V-cycle)
//
// Pi_u injects a Q1 displacement into the vertex dofs of the BR1 / CGQ1 space (Q1 is a subspace);
// Pi_p maps a Q1 pressure to (p_int = mean of the 4 cell vertices, p_face = mean of the 2 face
// vertices).  The Q1 operators live on the nested vertex grids of refine_global
// ((2^l+1)^2 vertices, lexicographic) as 9-point block stencils in SoA layout -- no index arrays,
// fully coalesced -- and are re-discretised on every level from the level's own vertex coordinates
// and the cell permeabilities averaged over children (Morton order makes children contiguous).
// Smoother: damped (block-)Jacobi; V(2,2); R = P^T (bilinear): V_u, V_p are symmetric positive
// definite, so MINRES stays applicable.
#include "pe_internal.h"
#include "pe_mg.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <functional>

namespace pe
{
  namespace
  {
    // Stencil coefficients are STORED in single precision and used in double-precision arithmetic: the
    // V-cycle is a preconditioner (any fixed symmetric positive definite operator serves MINRES), its
    // smoothers are bound by the HBM traffic of the coefficients (ncu: 6.0 TB/s), and rounding the
    // entries of a symmetric matrix keeps it symmetric.  288 -> 144 bytes per vertex and sweep.
    using stencil_t = MgLevel::stencil_t;

    __device__ __forceinline__ uint32_t
    part1by1(uint32_t x)
    {
      x &= 0x0000ffffu;
      x = (x ^ (x << 8)) & 0x00ff00ffu;
      x = (x ^ (x << 4)) & 0x0f0f0f0fu;
      x = (x ^ (x << 2)) & 0x33333333u;
      x = (x ^ (x << 1)) & 0x55555555u;
      return x;
    }
    __device__ __forceinline__ uint32_t
    morton(const int ix, const int iy)
    {
      return part1by1((uint32_t)ix) | (part1by1((uint32_t)iy) << 1);
    }

    // ---- stencil assembly: one thread per vertex, contributions of the <= 4 adjacent cells ---------
    // NC = 2: mu (d_cd grad N_a . grad N_b + d_d N_a d_c N_b) + lambda dbar dbar |E|   (vertex block
    //         of the BR1 operator, BR1new:904-925)
    // NC = 1: dtk grad N_a . grad N_b + gamma N_a N_b
    template <int NC>
    __global__ void
    k_mg_assemble(const int N, const double *__restrict__ vx, const double *__restrict__ vy,
                  const double *__restrict__ kc, const double k_scalar, const uint8_t *__restrict__ mask,
                  const double c_a /* mu | dt */, const double c_b /* lambda | gamma */, stencil_t *__restrict__ st,
                  double *__restrict__ dinv, const int v_begin, const int v_end)
    {
      const int V = N + 1, nv = V * V;
      const int v = v_begin + blockIdx.x * blockDim.x + threadIdx.x;
      if (v >= v_end)
        return;
      const int i = v % V, j = v / V;
      double    acc[9][NC * NC];
#pragma unroll
      for (int k = 0; k < 9; ++k)
#pragma unroll
        for (int e = 0; e < NC * NC; ++e)
          acc[k][e] = 0.;
      const bool self_masked = mask && mask[v];
      if (!self_masked)
        for (int cj = j - 1; cj <= j; ++cj)
          for (int ci = i - 1; ci <= i; ++ci)
            {
              if (ci < 0 || cj < 0 || ci >= N || cj >= N)
                continue;
              const int a  = (i - ci) + 2 * (j - cj);
              double    px[4], py[4];
#pragma unroll
              for (int b = 0; b < 4; ++b)
                {
                  const int w = (cj + (b >> 1)) * V + ci + (b & 1);
                  px[b]       = vx[w];
                  py[b]       = vy[w];
                }
              Geom G;
              G.set(px, py);
              const double kperm = kc ? kc[morton(ci, cj)] : k_scalar;
              double       gxx[4] = {0, 0, 0, 0}, gyy[4] = {0, 0, 0, 0}, gxy[4] = {0, 0, 0, 0}, gyx[4] = {0, 0, 0, 0};
              double       mass[4] = {0, 0, 0, 0}, dxs[4] = {0, 0, 0, 0}, dys[4] = {0, 0, 0, 0}, W = 0.;
              const double gp[2] = {0.21132486540518711775, 0.78867513459481288225};
#pragma unroll
              for (int qy = 0; qy < 2; ++qy)
#pragma unroll
                for (int qx = 0; qx < 2; ++qx)
                  {
                    const double xi = gp[qx], eta = gp[qy];
                    double       j00, j01, j10, j11, nx[4], ny[4];
                    G.jac(xi, eta, j00, j01, j10, j11);
                    const double det = j00 * j11 - j01 * j10;
                    const double r   = 0.25 / det;
                    shape_grads<4>(xi, eta, j00, j01, j10, j11, nx, ny);
                    W += 0.25 * det;
#pragma unroll
                    for (int b = 0; b < 4; ++b)
                      {
                        gxx[b] += r * nx[a] * nx[b];
                        gyy[b] += r * ny[a] * ny[b];
                        gxy[b] += r * nx[a] * ny[b];
                        gyx[b] += r * ny[a] * nx[b];
                        dxs[b] += 0.25 * nx[b];
                        dys[b] += 0.25 * ny[b];
                      }
                    if (NC == 1)
                      {
                        double s[4];
                        shape_values<4>(xi, eta, s);
#pragma unroll
                        for (int b = 0; b < 4; ++b)
                          mass[b] += 0.25 * det * s[a] * s[b];
                      }
                  }
#pragma unroll
              for (int b = 0; b < 4; ++b)
                {
                  const int wi = ci + (b & 1), wj = cj + (b >> 1);
                  const int w  = wj * V + wi;
                  if (b != a && mask && mask[w])
                    continue; // Dirichlet neighbour: column eliminated
                  const int k = (wj - j + 1) * 3 + (wi - i + 1);
                  if (NC == 2)
                    {
                      const double iw = 1. / W; // dbar = int d/dx N / |E|, times |E| once: dxs dxs / W
                      acc[k][0] += c_a * (2. * gxx[b] + gyy[b]) + c_b * dxs[a] * dxs[b] * iw;
                      acc[k][3 % (NC * NC)] += c_a * (gxx[b] + 2. * gyy[b]) + c_b * dys[a] * dys[b] * iw;
                      // (a,x),(b,y): mu d_y N_a d_x N_b ; (a,y),(b,x): mu d_x N_a d_y N_b
                      acc[k][1 % (NC * NC)] += c_a * gyx[b] + c_b * dxs[a] * dys[b] * iw;
                      acc[k][2 % (NC * NC)] += c_a * gxy[b] + c_b * dys[a] * dxs[b] * iw;
                    }
                  else
                    acc[k][0] += c_a * kperm * (gxx[b] + gyy[b]) + c_b * mass[b];
                }
            }
      if (self_masked)
        {
          acc[4][0] = 1.;
          if (NC == 2)
            acc[4][3 % (NC * NC)] = 1.;
        }
#pragma unroll
      for (int k = 0; k < 9; ++k)
#pragma unroll
        for (int e = 0; e < NC * NC; ++e)
          st[(size_t)(k * NC * NC + e) * nv + v] = (stencil_t)acc[k][e];
      if (NC == 1)
        dinv[v] = acc[4][0] != 0. ? 1. / (double)(stencil_t)acc[4][0] : 0.;
      else
        {
          // the diagonal block as it is stored (rounded), so that the smoother inverts the stored operator
          const double a00 = (stencil_t)acc[4][0], a01 = (stencil_t)acc[4][1 % (NC * NC)], a10 = (stencil_t)acc[4][2 % (NC * NC)],
                       a11 = (stencil_t)acc[4][3 % (NC * NC)];
          const double d   = a00 * a11 - a01 * a10;
          const double id  = d != 0. ? 1. / d : 0.;
          dinv[v]                  = a11 * id;
          dinv[(size_t)nv + v]     = -a01 * id;
          dinv[(size_t)2 * nv + v] = -a10 * id;
          dinv[(size_t)3 * nv + v] = a00 * id;
        }
    }

    // (A x)[v] for the 9-point block stencil; x is [NC][nv]
    template <int NC>
    __device__ __forceinline__ void
    stencil_row(const int N, const int v, const stencil_t *__restrict__ st, const double *__restrict__ x, double out[NC])
    {
      const int V = N + 1, nv = V * V;
      const int i = v % V, j = v / V;
#pragma unroll
      for (int c = 0; c < NC; ++c)
        out[c] = 0.;
#pragma unroll
      for (int dj = -1; dj <= 1; ++dj)
#pragma unroll
        for (int di = -1; di <= 1; ++di)
          {
            const int wi = i + di, wj = j + dj;
            if (wi < 0 || wj < 0 || wi > N || wj > N)
              continue;
            const int w = wj * V + wi, k = (dj + 1) * 3 + (di + 1);
            double    xv[NC];
#pragma unroll
            for (int d = 0; d < NC; ++d)
              xv[d] = x[(size_t)d * nv + w];
#pragma unroll
            for (int c = 0; c < NC; ++c)
#pragma unroll
              for (int d = 0; d < NC; ++d)
                out[c] += (double)st[(size_t)(k * NC * NC + c * NC + d) * nv + v] * xv[d];
          }
    }

    // xo = xi + omega Dinv (b - A xi)     (xi == nullptr: xo = omega Dinv b)
    template <int NC>
    __global__ void
    k_mg_jacobi(const int N, const stencil_t *__restrict__ st, const double *__restrict__ dinv,
                const double *__restrict__ b, const double *__restrict__ xi, double *__restrict__ xo, const double omega,
                const int v_begin, const int v_end)
    {
      const int nv = (N + 1) * (N + 1);
      const int v  = v_begin + blockIdx.x * blockDim.x + threadIdx.x;
      if (v >= v_end)
        return;
      double r[NC];
      if (xi)
        {
          stencil_row<NC>(N, v, st, xi, r);
#pragma unroll
          for (int c = 0; c < NC; ++c)
            r[c] = b[(size_t)c * nv + v] - r[c];
        }
      else
        {
#pragma unroll
          for (int c = 0; c < NC; ++c)
            r[c] = b[(size_t)c * nv + v];
        }
      if (NC == 1)
        xo[v] = (xi ? xi[v] : 0.) + omega * dinv[v] * r[0];
      else
        {
          const double d00 = dinv[v], d01 = dinv[(size_t)nv + v], d10 = dinv[(size_t)2 * nv + v], d11 = dinv[(size_t)3 * nv + v];
          xo[v]              = (xi ? xi[v] : 0.) + omega * (d00 * r[0] + d01 * r[NC - 1]);
          xo[(size_t)nv + v] = (xi ? xi[(size_t)nv + v] : 0.) + omega * (d10 * r[0] + d11 * r[NC - 1]);
        }
    }

    // coarsest level (<= 1024 vertices): `sweeps` damped Jacobi sweeps from x = 0 in ONE single-CTA launch
    // (ping-pong in shared memory); replaces 24 tiny launches per V-cycle.
    template <int NC>
    __global__ void __launch_bounds__(1024)
    k_mg_coarse(const int N, const stencil_t *__restrict__ st, const double *__restrict__ dinv,
                const double *__restrict__ b, double *__restrict__ xo, const double omega, const int sweeps)
    {
      extern __shared__ double sm[]; // 2 * NC * nv
      const int nv = (N + 1) * (N + 1);
      double   *x0 = sm, *x1 = sm + NC * nv;
      for (int v = threadIdx.x; v < NC * nv; v += blockDim.x)
        x0[v] = 0.;
      __syncthreads();
      for (int it = 0; it < sweeps; ++it)
        {
          for (int v = threadIdx.x; v < nv; v += blockDim.x)
            {
              double r[NC];
              stencil_row<NC>(N, v, st, x0, r);
#pragma unroll
              for (int c = 0; c < NC; ++c)
                r[c] = b[(size_t)c * nv + v] - r[c];
              if (NC == 1)
                x1[v] = x0[v] + omega * dinv[v] * r[0];
              else
                {
                  const double d00 = dinv[v], d01 = dinv[(size_t)nv + v], d10 = dinv[(size_t)2 * nv + v],
                               d11 = dinv[(size_t)3 * nv + v];
                  x1[v]      = x0[v] + omega * (d00 * r[0] + d01 * r[NC - 1]);
                  x1[nv + v] = x0[nv + v] + omega * (d10 * r[0] + d11 * r[NC - 1]);
                }
            }
          __syncthreads();
          double *t = x0;
          x0        = x1;
          x1        = t;
        }
      for (int v = threadIdx.x; v < NC * nv; v += blockDim.x)
        xo[v] = x0[v];
    }

    // ---- all levels with N <= kFusedN in ONE single-CTA launch: vectors in shared memory, stencils from L2 --------
    // (every small level costs ~7 launches of 3-9 us each in the graph; fused: one ~50 us kernel)
    constexpr int kFusedN   = 32;
    constexpr int kFusedLev = 6;
    struct FusedArgs
    {
      int            nlev;
      int            N[kFusedLev];
      const stencil_t *st[kFusedLev];
      const double    *dinv[kFusedLev];
      const uint8_t *mask[kFusedLev];
      const double  *b0;
      double        *out;
      const double  *coarse_inv; // dense inverse of the last level (n*n + flag), or null
    };
    template <int NC>
    __device__ __forceinline__ void
    jacobi_point(const int N, const int v, const stencil_t *__restrict__ st, const double *__restrict__ dinv, const double *b,
                 const double *xi, double *xo, const double omega)
    {
      const int nv = (N + 1) * (N + 1);
      double    r[NC];
      if (xi)
        {
          stencil_row<NC>(N, v, st, xi, r);
#pragma unroll
          for (int c = 0; c < NC; ++c)
            r[c] = b[c * nv + v] - r[c];
        }
      else
        {
#pragma unroll
          for (int c = 0; c < NC; ++c)
            r[c] = b[c * nv + v];
        }
      if (NC == 1)
        xo[v] = (xi ? xi[v] : 0.) + omega * dinv[v] * r[0];
      else
        {
          const double d00 = dinv[v], d01 = dinv[(size_t)nv + v], d10 = dinv[(size_t)2 * nv + v], d11 = dinv[(size_t)3 * nv + v];
          xo[v]      = (xi ? xi[v] : 0.) + omega * (d00 * r[0] + d01 * r[NC - 1]);
          xo[nv + v] = (xi ? xi[nv + v] : 0.) + omega * (d10 * r[0] + d11 * r[NC - 1]);
        }
    }
    template <int NC>
    __global__ void __launch_bounds__(1024)
    k_mg_fused(const FusedArgs a, const double omega, const double omega2, const int nu, const int coarse_sweeps)
    {
      extern __shared__ double sm[]; // per level: b, x0, x1 (NC * nv each)
      double *B[kFusedLev], *X[kFusedLev], *Y[kFusedLev];
      {
        double *p = sm;
        for (int m = 0; m < a.nlev; ++m)
          {
            const int n = NC * (a.N[m] + 1) * (a.N[m] + 1);
            B[m] = p;
            X[m] = p + n;
            Y[m] = p + 2 * n;
            p += 3 * n;
          }
      }
      const int t = threadIdx.x, nt = blockDim.x;
      {
        const int n = NC * (a.N[0] + 1) * (a.N[0] + 1);
        for (int i = t; i < n; i += nt)
          B[0][i] = a.b0[i];
      }
      __syncthreads();
      // ---- down
      for (int m = 0; m < a.nlev; ++m)
        {
          const int  N = a.N[m], nv = (N + 1) * (N + 1);
          const bool last = m + 1 == a.nlev;
          const int  sweeps = last ? coarse_sweeps : nu;
          if (last && a.coarse_inv && a.coarse_inv[NC * nv * NC * nv] == 1.)
            {
              // exact coarse solve: x = A^{-1} b with the precomputed dense inverse
              const int n = NC * nv;
              for (int r = t; r < n; r += nt)
                {
                  double sum = 0.;
                  for (int c = 0; c < n; ++c)
                    sum += a.coarse_inv[r * n + c] * B[m][c];
                  X[m][r] = sum;
                }
              __syncthreads();
              break;
            }
          for (int v = t; v < nv; v += nt)
            jacobi_point<NC>(N, v, a.st[m], a.dinv[m], B[m], nullptr, X[m], omega);
          __syncthreads();
          for (int k = 1; k < sweeps; ++k)
            {
              const double om = (k & 1) ? omega2 : omega; // alternating (Chebyshev-like) weights
              for (int v = t; v < nv; v += nt)
                jacobi_point<NC>(N, v, a.st[m], a.dinv[m], B[m], X[m], Y[m], om);
              __syncthreads();
              double *q = X[m];
              X[m]      = Y[m];
              Y[m]      = q;
            }
          if (last)
            break;
          // residual -> Y[m]; full-weighting restriction -> B[m+1]
          for (int v = t; v < nv; v += nt)
            {
              double ax[NC];
              stencil_row<NC>(N, v, a.st[m], X[m], ax);
#pragma unroll
              for (int c = 0; c < NC; ++c)
                Y[m][c * nv + v] = B[m][c * nv + v] - ax[c];
            }
          __syncthreads();
          const int Nc = N / 2, Vc = Nc + 1, nvc = Vc * Vc, Vf = N + 1;
          for (int vc = t; vc < nvc; vc += nt)
            {
              const int I = vc % Vc, J = vc / Vc;
              double    sres[NC];
#pragma unroll
              for (int c = 0; c < NC; ++c)
                sres[c] = 0.;
              if (!(a.mask[m + 1] && a.mask[m + 1][vc]))
                for (int dj = -1; dj <= 1; ++dj)
                  for (int di = -1; di <= 1; ++di)
                    {
                      const int fi = 2 * I + di, fj = 2 * J + dj;
                      if (fi < 0 || fj < 0 || fi > N || fj > N)
                        continue;
                      const double w = (di == 0 ? 1. : .5) * (dj == 0 ? 1. : .5);
#pragma unroll
                      for (int c = 0; c < NC; ++c)
                        sres[c] += w * Y[m][c * nv + fj * Vf + fi];
                    }
#pragma unroll
              for (int c = 0; c < NC; ++c)
                B[m + 1][c * nvc + vc] = sres[c];
            }
          __syncthreads();
        }
      // ---- up
      for (int m = a.nlev - 2; m >= 0; --m)
        {
          const int N = a.N[m], Vf = N + 1, nv = Vf * Vf, Nc = N / 2, Vc = Nc + 1, nvc = Vc * Vc;
          for (int v = t; v < nv; v += nt)
            {
              if (a.mask[m] && a.mask[m][v])
                continue;
              const int i = v % Vf, j = v / Vf;
              const int I0 = i >> 1, J0 = j >> 1, I1 = (i + 1) >> 1, J1 = (j + 1) >> 1;
#pragma unroll
              for (int c = 0; c < NC; ++c)
                {
                  const double *q = X[m + 1] + c * nvc;
                  X[m][c * nv + v] += 0.25 * (q[J0 * Vc + I0] + q[J0 * Vc + I1] + q[J1 * Vc + I0] + q[J1 * Vc + I1]);
                }
            }
          __syncthreads();
          for (int k = 0; k < nu; ++k)
            {
              const double om = ((nu - 1 - k) & 1) ? omega2 : omega;
              for (int v = t; v < nv; v += nt)
                jacobi_point<NC>(N, v, a.st[m], a.dinv[m], B[m], X[m], Y[m], om);
              __syncthreads();
              double *q = X[m];
              X[m]      = Y[m];
              Y[m]      = q;
            }
        }
      {
        const int n = NC * (a.N[0] + 1) * (a.N[0] + 1);
        for (int i = t; i < n; i += nt)
          a.out[i] = X[0][i];
      }
    }

    // dense inverse of the coarsest-level operator (N = 2: 9 vertices, NC * 9 unknowns), one thread, once per
    // assembled matrix.  inv[n*n] row-major, inv[n*n] = 1 if the factorisation succeeded (0: singular, e.g. pure
    // traction problems -> the fused kernel falls back to Jacobi sweeps).
    template <int NC>
    __global__ void
    k_mg_coarse_inverse(const int N, const stencil_t *__restrict__ st, double *__restrict__ inv)
    {
      constexpr int MAXN = NC * 25;
      const int     V = N + 1, nv = V * V, n = NC * nv;
      if (threadIdx.x != 0 || blockIdx.x != 0 || n > MAXN)
        return;
      __shared__ double A[MAXN * MAXN], B[MAXN * MAXN];
      for (int e = 0; e < n * n; ++e)
        {
          A[e] = 0.;
          B[e] = 0.;
        }
      for (int v = 0; v < nv; ++v)
        {
          const int i = v % V, j = v / V;
          for (int dj = -1; dj <= 1; ++dj)
            for (int di = -1; di <= 1; ++di)
              {
                const int wi = i + di, wj = j + dj;
                if (wi < 0 || wj < 0 || wi > N || wj > N)
                  continue;
                const int w = wj * V + wi, k = (dj + 1) * 3 + (di + 1);
                for (int c = 0; c < NC; ++c)
                  for (int d = 0; d < NC; ++d)
                    A[(c * nv + v) * n + d * nv + w] = (double)st[(size_t)(k * NC * NC + c * NC + d) * nv + v];
              }
        }
      for (int r = 0; r < n; ++r)
        B[r * n + r] = 1.;
      double amax = 0.;
      for (int e = 0; e < n * n; ++e)
        amax = fmax(amax, fabs(A[e]));
      bool ok = amax > 0.;
      for (int k = 0; k < n && ok; ++k)
        {
          int    p  = k;
          double pm = fabs(A[k * n + k]);
          for (int r = k + 1; r < n; ++r)
            if (fabs(A[r * n + k]) > pm)
              {
                pm = fabs(A[r * n + k]);
                p  = r;
              }
          if (pm < 1e-12 * amax)
            {
              ok = false;
              break;
            }
          if (p != k)
            for (int c = 0; c < n; ++c)
              {
                const double ta = A[k * n + c], tb = B[k * n + c];
                A[k * n + c]    = A[p * n + c];
                B[k * n + c]    = B[p * n + c];
                A[p * n + c]    = ta;
                B[p * n + c]    = tb;
              }
          const double ip = 1. / A[k * n + k];
          for (int c = 0; c < n; ++c)
            {
              A[k * n + c] *= ip;
              B[k * n + c] *= ip;
            }
          for (int r = 0; r < n; ++r)
            if (r != k)
              {
                const double f = A[r * n + k];
                if (f != 0.)
                  for (int c = 0; c < n; ++c)
                    {
                      A[r * n + c] -= f * A[k * n + c];
                      B[r * n + c] -= f * B[k * n + c];
                    }
              }
        }
      // symmetrise (the operator is symmetric; rounding of the elimination is not)
      for (int r = 0; r < n; ++r)
        for (int c = 0; c < n; ++c)
          inv[r * n + c] = ok ? .5 * (B[r * n + c] + B[c * n + r]) : 0.;
      inv[n * n] = ok ? 1. : 0.;
    }

    // r = b - A x
    template <int NC>
    __global__ void
    k_mg_residual(const int N, const stencil_t *__restrict__ st, const double *__restrict__ b,
                  const double *__restrict__ x, double *__restrict__ r, const int v_begin, const int v_end)
    {
      const int nv = (N + 1) * (N + 1);
      const int v  = v_begin + blockIdx.x * blockDim.x + threadIdx.x;
      if (v >= v_end)
        return;
      double ax[NC];
      stencil_row<NC>(N, v, st, x, ax);
#pragma unroll
      for (int c = 0; c < NC; ++c)
        r[(size_t)c * nv + v] = b[(size_t)c * nv + v] - ax[c];
    }

    // coarse b = P^T r   (full weighting), one thread per COARSE vertex
    template <int NC>
    __global__ void
    k_mg_restrict(const int Nf, const double *__restrict__ r, double *__restrict__ bc,
                  const uint8_t *__restrict__ mask_c, const int v_begin, const int v_end)
    {
      const int Nc = Nf / 2, Vc = Nc + 1, nvc = Vc * Vc, Vf = Nf + 1, nvf = Vf * Vf;
      const int vc = v_begin + blockIdx.x * blockDim.x + threadIdx.x;
      if (vc >= v_end)
        return;
      const int I = vc % Vc, J = vc / Vc;
      double    s[NC];
#pragma unroll
      for (int c = 0; c < NC; ++c)
        s[c] = 0.;
      if (!(mask_c && mask_c[vc]))
        for (int dj = -1; dj <= 1; ++dj)
          for (int di = -1; di <= 1; ++di)
            {
              const int fi = 2 * I + di, fj = 2 * J + dj;
              if (fi < 0 || fj < 0 || fi > Nf || fj > Nf)
                continue;
              const double w  = (di == 0 ? 1. : .5) * (dj == 0 ? 1. : .5);
              const int    vf = fj * Vf + fi;
#pragma unroll
              for (int c = 0; c < NC; ++c)
                s[c] += w * r[(size_t)c * nvf + vf];
            }
#pragma unroll
      for (int c = 0; c < NC; ++c)
        bc[(size_t)c * nvc + vc] = s[c];
    }

    // xf += P xc   (bilinear), one thread per FINE vertex
    template <int NC>
    __global__ void
    k_mg_prolong_add(const int Nf, const double *__restrict__ xc, double *__restrict__ xf,
                     const uint8_t *__restrict__ mask_f, const int v_begin, const int v_end)
    {
      const int Nc = Nf / 2, Vc = Nc + 1, nvc = Vc * Vc, Vf = Nf + 1, nvf = Vf * Vf;
      const int vf = v_begin + blockIdx.x * blockDim.x + threadIdx.x;
      if (vf >= v_end)
        return;
      if (mask_f && mask_f[vf])
        return;
      const int i = vf % Vf, j = vf / Vf;
      const int I0 = i >> 1, J0 = j >> 1, I1 = (i + 1) >> 1, J1 = (j + 1) >> 1;
#pragma unroll
      for (int c = 0; c < NC; ++c)
        {
          const double *q = xc + (size_t)c * nvc;
          xf[(size_t)c * nvf + vf] +=
            0.25 * (q[J0 * Vc + I0] + q[J0 * Vc + I1] + q[J1 * Vc + I0] + q[J1 * Vc + I1]);
        }
    }

    // coarse arrays by injection / averaging
    __global__ void
    k_mg_coarsen_geometry(const int Nc, const double *__restrict__ xf, const double *__restrict__ yf,
                          const uint8_t *__restrict__ mu_f, const uint8_t *__restrict__ mp_f, double *__restrict__ xc,
                          double *__restrict__ yc, uint8_t *__restrict__ mu_c, uint8_t *__restrict__ mp_c)
    {
      const int Vc = Nc + 1, Vf = 2 * Nc + 1;
      const int v  = blockIdx.x * blockDim.x + threadIdx.x;
      if (v >= Vc * Vc)
        return;
      const int I = v % Vc, J = v / Vc, f = (2 * J) * Vf + 2 * I;
      xc[v]   = xf[f];
      yc[v]   = yf[f];
      mu_c[v] = mu_f[f];
      mp_c[v] = mp_f[f];
    }
    inline unsigned
    g1(const int n)
    {
      return (unsigned)((n + 127) / 128);
    }
  } // namespace

  // -------------------------------------------------------------------------------------------------
  MgHierarchy::~MgHierarchy() = default;

  int
  mg_build_topology(MgHierarchy &mg, const HostMesh &mesh, const HostDofs &dofs,
                    const std::function<int32_t(uint32_t)> &to_local, cudaStream_t s)
  {
    mg.ready = false;
    if (mesh.n_refine < 1 || (dofs.element != 0 && dofs.element != 1))
      return 1; // no hierarchy available: caller falls back to the diagonal preconditioner
    const int N = 1 << mesh.n_refine, V = N + 1, n_loc = dofs.n_loc;
    mg.N0       = N;
    // level arrays
    int nl = 0;
    for (int n = N; n >= 2; n /= 2)
      ++nl;
    mg.lv.clear();
    mg.lv.resize((size_t)nl);
    // dof maps on the fine vertex grid
    std::vector<int32_t> vdof((size_t)V * V, -1), pint((size_t)N * N), hedge((size_t)V * N), vedge((size_t)N * V);
    std::vector<uint8_t> mask_u((size_t)V * V, 0), mask_p((size_t)V * V, 0);
    const int pf0 = dofs.element == 0 ? 9 : 8, pfs = dofs.element == 0 ? 2 : 1;
    for (int64_t c = 0; c < mesh.n_cells; ++c)
      {
        const uint32_t *cv = &mesh.cell_vertices[4 * (size_t)c];
        const uint32_t *cd = &dofs.cell_dofs[(size_t)c * n_loc];
        const int       ci = (int)(cv[0] % V), cj = (int)(cv[0] / V);
        for (int v = 0; v < 4; ++v)
          vdof[cv[v]] = to_local(cd[2 * v]);
        pint[(size_t)cj * N + ci]        = to_local(cd[n_loc - 1]);
        vedge[(size_t)cj * V + ci]       = to_local(cd[pf0 + 0 * pfs]); // face 0: x = left, vertical edge
        vedge[(size_t)cj * V + ci + 1]   = to_local(cd[pf0 + 1 * pfs]);
        hedge[(size_t)cj * N + ci]       = to_local(cd[pf0 + 2 * pfs]); // face 2: bottom, horizontal edge
        hedge[(size_t)(cj + 1) * N + ci] = to_local(cd[pf0 + 3 * pfs]);
        for (int f = 0; f < 4; ++f)
          {
            const uint8_t k = mesh.face_kinds[4 * (size_t)c + f];
            for (int e = 0; e < 2; ++e)
              {
                const uint32_t w = cv[kFaceVertex[f][e]];
                if (k & PE_FACE_DIR_U)
                  mask_u[w] = 1;
                if (k & PE_FACE_DIR_P)
                  mask_p[w] = 1;
              }
          }
      }
    auto up = [&](auto &dev, const auto &host) -> bool {
      if (dev.alloc(host.size()) != cudaSuccess)
        return false;
      return cudaMemcpyAsync(dev.p, host.data(), host.size() * sizeof(host[0]), cudaMemcpyHostToDevice, s) == cudaSuccess;
    };
    if (!up(mg.vdof, vdof) || !up(mg.pint, pint) || !up(mg.hedge, hedge) || !up(mg.vedge, vedge))
      return -1;
    MgLevel &l0 = mg.lv[0];
    l0.N        = N;
    if (!up(l0.x, mesh.x) || !up(l0.y, mesh.y) || !up(l0.mask_u, mask_u) || !up(l0.mask_p, mask_p))
      return -1;
    cudaStreamSynchronize(s);
    for (int l = 0; l < nl; ++l)
      {
        MgLevel &L = mg.lv[(size_t)l];
        L.N        = N >> l;
        const size_t nv = (size_t)(L.N + 1) * (L.N + 1);
        if (l > 0)
          {
            if (L.x.alloc(nv) || L.y.alloc(nv) || L.mask_u.alloc(nv) || L.mask_p.alloc(nv))
              return -1;
            MgLevel &F = mg.lv[(size_t)l - 1];
            k_mg_coarsen_geometry<<<g1((int)nv), 128, 0, s>>>(L.N, F.x.p, F.y.p, F.mask_u.p, F.mask_p.p, L.x.p, L.y.p,
                                                               L.mask_u.p, L.mask_p.p);
          }
        if (L.st_u.alloc(36 * nv) || L.dinv_u.alloc(4 * nv))
          return -1;
        for (int k = 0; k < 3; ++k)
          if (L.wu[k].alloc(2 * nv))
            return -1;
        if (L.bu.alloc(2 * nv))
          return -1;
      }
    if (cudaStreamSynchronize(s) != cudaSuccess)
      return -1;
    mg.ready = true;
    return 0;
  }

  // (re)discretise all levels from the current material data: called once per solve (the reference
  // refactorises every step; the "cached" mode of SURVEY 8(f).1 would skip this)
  int
  mg_assemble(MgHierarchy &mg, const double *kcell_fine /* Morton, may be null */, const double k_scalar,
              const double lambda, const double mu, const double dt, const double gamma, cudaStream_t s, int64_t *launches)
  {
    (void)kcell_fine; // the pressure block has its own cell-centred hierarchy (pe_cmg.cu)
    (void)k_scalar;
    (void)dt;
    (void)gamma;
    for (size_t l = 0; l < mg.lv.size(); ++l)
      {
        MgLevel  &L  = mg.lv[l];
        const int V  = L.N + 1;
        // a distributed level only smooths its own rows: stencils of the slab (round 1 built the whole level on every rank)
        const bool dist = (int)l < mg.n_dist;
        const int  v0   = dist ? (int)(mg.bounds[l][(size_t)mg.gc.rank] * V) : 0;
        const int  v1   = dist ? (int)(mg.bounds[l][(size_t)mg.gc.rank + 1] * V) : V * V;
        if (v1 > v0)
          k_mg_assemble<2><<<g1(v1 - v0), 128, 0, s>>>(L.N, L.x.p, L.y.p, nullptr, 1., L.mask_u.p, mu, lambda, L.st_u.p, L.dinv_u.p,
                                                       v0, v1);
        ++*launches;
      }
    {
      // exact coarse solve of the displacement hierarchy (replaces 24 Jacobi sweeps = 24 latency-bound phases of
      // the single-CTA kernel)
      MgLevel  &C = mg.lv.back();
      const int n = 2 * (C.N + 1) * (C.N + 1);
      if (C.N == 2)
        {
          if (mg.coarse_inv_u.n < (size_t)n * n + 1 && mg.coarse_inv_u.alloc((size_t)n * n + 1) != cudaSuccess)
            return -1;
          k_mg_coarse_inverse<2><<<1, 32, 0, s>>>(C.N, C.st_u.p, mg.coarse_inv_u.p);
          ++*launches;
        }
    }
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
  }

  template <int NC>
  static void
  vcycle(MgHierarchy &mg, const size_t l, const double omega, const int nu, cudaStream_t s, int64_t *launches)
  {
    MgLevel      &L    = mg.lv[l];
    const int     nv   = (L.N + 1) * (L.N + 1);
    const stencil_t *st = L.st_u.p;
    const double *dinv = L.dinv_u.p;
    double       *b    = L.bu.p;
    double       *x0   = L.wu[0].p;
    double       *x1   = L.wu[1].p;
    const bool    last = l + 1 == mg.lv.size();
    const int fused_n = mg.fused_n > 0 ? std::min(kFusedN, mg.fused_n) : kFusedN;
    if (L.N <= fused_n && !last)
      {
        FusedArgs a{};
        a.nlev = 0;
        for (size_t q = l; q < mg.lv.size() && a.nlev < kFusedLev; ++q)
          {
            MgLevel &Q     = mg.lv[q];
            a.N[a.nlev]    = Q.N;
            a.st[a.nlev]   = Q.st_u.p;
            a.dinv[a.nlev] = Q.dinv_u.p;
            a.mask[a.nlev] = Q.mask_u.p;
            ++a.nlev;
          }
        a.b0  = b;
        a.out = L.wu[2].p;
        a.coarse_inv = (NC == 2 && l + (size_t)a.nlev == mg.lv.size() && mg.coarse_inv_u.n > 0) ? mg.coarse_inv_u.p : nullptr;
        size_t smem = 0;
        for (int m = 0; m < a.nlev; ++m)
          smem += sizeof(double) * 3 * NC * (size_t)(a.N[m] + 1) * (a.N[m] + 1);
        static bool attr_set[3] = {false, false, false};
        if (!attr_set[NC])
          {
            cudaFuncSetAttribute(k_mg_fused<NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
            attr_set[NC] = true;
          }
        k_mg_fused<NC><<<1, 1024, smem, s>>>(a, omega, mg.omega2 > 0. ? mg.omega2 : omega, nu, 24);
        ++*launches;
        return;
      }
    if (last)
      {
        // coarse solve: 24 Jacobi sweeps in one launch, result straight into w[2]
        const int threads = std::min(1024, ((nv + 31) / 32) * 32);
        k_mg_coarse<NC><<<1, threads, sizeof(double) * 2 * NC * nv, s>>>(L.N, st, dinv, b, L.wu[2].p,
                                                                         omega, 24);
        ++*launches;
        return;
      }
    // ---- this level: distributed in row slabs (multi-GPU, fine levels) or complete
    const int     V    = L.N + 1;
    const bool    dist = (int)l < mg.n_dist;
    const int64_t ra = dist ? mg.bounds[l][(size_t)mg.gc.rank] : 0, rb = dist ? mg.bounds[l][(size_t)mg.gc.rank + 1] : V;
    const int     v0 = (int)(ra * V), v1 = (int)(rb * V), nloc = v1 - v0;
    auto exchange = [&](double *p) {
      if (dist)
        mg.gc.exchange_rows(mg.gc.ctx, p, V, NC, nv, ra, rb);
    };
    const int pre = nu;
    // pre-smoothing from x = 0
    k_mg_jacobi<NC><<<g1(nloc), 128, 0, s>>>(L.N, st, dinv, b, nullptr, x0, omega, v0, v1);
    ++*launches;
    const double omega2 = mg.omega2 > 0. ? mg.omega2 : omega; // sweeps alternate between the two weights
    for (int k = 1; k < pre; ++k)
      {
        exchange(x0);
        k_mg_jacobi<NC><<<g1(nloc), 128, 0, s>>>(L.N, st, dinv, b, x0, x1, (k & 1) ? omega2 : omega, v0, v1);
        ++*launches;
        std::swap(x0, x1);
      }
    {
      MgLevel      &C      = mg.lv[l + 1];
      const int     Vc     = C.N + 1, nvc = Vc * Vc;
      const bool    dist_c = (int)l + 1 < mg.n_dist;
      // coarse rows whose centre row 2J lies in my slab
      const int64_t ca = dist ? (ra + 1) / 2 : 0, cb = dist ? (rb + 1) / 2 : Vc;
      double       *bc = C.bu.p;
      exchange(x0);
      k_mg_residual<NC><<<g1(nloc), 128, 0, s>>>(L.N, st, b, x0, x1, v0, v1);
      exchange(x1);
      k_mg_restrict<NC><<<g1((int)((cb - ca) * Vc)), 128, 0, s>>>(L.N, x1, bc, C.mask_u.p,
                                                                  (int)(ca * Vc), (int)(cb * Vc));
      *launches += 2;
      if (dist && !dist_c)
        {
          // first replicated level: every rank needs the complete right-hand side
          std::vector<int64_t> cbnd((size_t)mg.gc.world + 1);
          for (int r = 0; r <= mg.gc.world; ++r)
            cbnd[(size_t)r] = (mg.bounds[l][(size_t)r] + 1) / 2;
          mg.gc.allgather_rows(mg.gc.ctx, bc, Vc, NC, nvc, cbnd.data());
        }
      vcycle<NC>(mg, l + 1, omega, nu, s, launches);
      k_mg_prolong_add<NC><<<g1(nloc), 128, 0, s>>>(L.N, C.wu[2].p, x0,
                                                    L.mask_u.p, v0, v1);
      ++*launches;
      for (int k = 0; k < nu; ++k)
        {
          exchange(x0);
          k_mg_jacobi<NC><<<g1(nloc), 128, 0, s>>>(L.N, st, dinv, b, x0, x1, ((nu - 1 - k) & 1) ? omega2 : omega, v0, v1);
          ++*launches;
          std::swap(x0, x1);
        }
    }
    // result of this level -> w[2]
    double *out = L.wu[2].p;
    for (int c = 0; c < NC; ++c)
      cudaMemcpyAsync(out + (size_t)c * nv + v0, x0 + (size_t)c * nv + v0, sizeof(double) * (size_t)nloc,
                      cudaMemcpyDeviceToDevice, s);
    if (dist)
      {
        if (l == 0 && mg.gc.slabs_to_boxes)
          mg.gc.slabs_to_boxes(mg.gc.ctx, out, V, NC, nv, mg.bounds[0].data(), 0); // every rank scatters the vertices of its box
        else if (l == 0)
          mg.gc.allgather_rows(mg.gc.ctx, out, V, NC, nv, mg.bounds[0].data());
        else
          exchange(out); // the finer level prolongates from my rows and one halo row on each side
      }
  }

  void
  mg_set_comm(MgHierarchy &mg, const GridComm &gc)
  {
    mg.gc     = gc;
    mg.n_dist = 0;
    mg.bounds.clear();
    if (gc.world <= 1 || !gc.exchange_rows || !gc.allgather_rows)
      return;
    const int W = gc.world;
    for (size_t l = 0; l + 1 < mg.lv.size(); ++l)
      {
        const int N = mg.lv[l].N;
        // distributed while every rank keeps >= 8 cell rows and the level is above the fused single-CTA levels
        // a distributed level pays ~6 NCCL row exchanges (~45 us each, measured on 8 GPUs); below ~1024^2 the
        // replicated sweep is cheaper than that latency
        if (N % (2 * W) != 0 || N / W < 8 || N < gc.dist_min_n)
          break;
        std::vector<int64_t> bd((size_t)W + 1);
        for (int r = 0; r < W; ++r)
          bd[(size_t)r] = (int64_t)N * r / W;
        bd[(size_t)W] = N + 1;
        mg.bounds.push_back(bd);
        ++mg.n_dist;
      }
  }

  void
  mg_vcycle(MgHierarchy &mg, const int nc, cudaStream_t s, int64_t *launches)
  {
    (void)nc; // only the vector (displacement) hierarchy lives here
    vcycle<2>(mg, 0, mg.omega, mg.nu, s, launches);
  }

} // namespace pe
