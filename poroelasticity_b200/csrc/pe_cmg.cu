// Cell-centred aggregation multigrid for the pressure block (see pe_cmg.h).  New work: the reference
// factorises the whole system with UMFPACK (PoroElasBR1WG_new.cc:1177-1187).
#include "pe_cmg.h"

#include <algorithm>
#include <cmath>

namespace pe
{
  namespace
  {
    constexpr int kCoarseN     = 32;   // levels with N <= kCoarseN run inside the single-CTA kernel
    constexpr int kCoarseCells = 1365; // 32^2 + 16^2 + ... + 1

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

    struct CellGeo
    {
      double cx, cy, area;
    };
    __device__ __forceinline__ CellGeo
    cell_geo(const int N, const int i, const int j, const double *__restrict__ vx, const double *__restrict__ vy)
    {
      const int    V  = N + 1, v0 = j * V + i;
      const double x0 = vx[v0], y0 = vy[v0], x1 = vx[v0 + 1], y1 = vy[v0 + 1];
      const double x2 = vx[v0 + V], y2 = vy[v0 + V], x3 = vx[v0 + V + 1], y3 = vy[v0 + V + 1];
      CellGeo      g;
      g.cx   = 0.25 * (x0 + x1 + x2 + x3);
      g.cy   = 0.25 * (y0 + y1 + y2 + y3);
      g.area = 0.5 * ((x3 - x0) * (y2 - y1) - (x2 - x1) * (y3 - y0));
      return g;
    }
    // face a-b seen from the cell centre c: length and distance of the centre to the face along its normal
    __device__ __forceinline__ void
    face_geo(const double ax, const double ay, const double bx, const double by, const double cx, const double cy,
             double &len, double &dist)
    {
      const double tx = bx - ax, ty = by - ay;
      len             = sqrt(tx * tx + ty * ty);
      const double mx = 0.5 * (ax + bx) - cx, my = 0.5 * (ay + by) - cy;
      dist            = fabs(mx * ty - my * tx) / len;
    }

    // ---- finest level: two-point-flux transmissibilities, absorption, face interpolation weights ----
    __global__ void
    k_cmg_assemble0(const int N, const double *__restrict__ vx, const double *__restrict__ vy,
                    const double *__restrict__ kc, const double k_scalar, const double dt, const double gamma,
                    const uint8_t *__restrict__ dir_v, const uint8_t *__restrict__ dir_h, double *__restrict__ tE,
                    double *__restrict__ tN, double *__restrict__ sig, double *__restrict__ wv_lo,
                    double *__restrict__ wv_hi, double *__restrict__ wh_lo, double *__restrict__ wh_hi)
    {
      const int t = blockIdx.x * blockDim.x + threadIdx.x;
      if (t >= N * N)
        return;
      const int     i = t % N, j = t / N, V = N + 1;
      const CellGeo g = cell_geo(N, i, j, vx, vy);
      const double  k = kc ? kc[morton(i, j)] : k_scalar;
      double        s = gamma * g.area;
      double        len, d1, d2;
      // east face: vertices (i+1, j) - (i+1, j+1); vertical edge index j*V + i+1
      {
        const int a = j * V + i + 1, b = a + V, e = j * V + i + 1;
        face_geo(vx[a], vy[a], vx[b], vy[b], g.cx, g.cy, len, d1);
        if (i < N - 1)
          {
            const CellGeo g2 = cell_geo(N, i + 1, j, vx, vy);
            const double  k2 = kc ? kc[morton(i + 1, j)] : k_scalar;
            face_geo(vx[a], vy[a], vx[b], vy[b], g2.cx, g2.cy, len, d2);
            const double c1 = k / d1, c2 = k2 / d2;
            tE[t]    = dt * len * c1 * c2 / (c1 + c2);
            wv_lo[e] = c1 / (c1 + c2);
            wv_hi[e] = c2 / (c1 + c2);
          }
        else
          {
            tE[t] = 0.;
            if (dir_v[e])
              s += dt * len * k / d1;
            wv_lo[e] = dir_v[e] ? 0. : 1.;
            wv_hi[e] = 0.;
          }
      }
      if (i == 0) // west boundary face: vertical edge j*V
        {
        const int a = j * V, b = a + V, e = j * V;
        face_geo(vx[a], vy[a], vx[b], vy[b], g.cx, g.cy, len, d1);
        if (dir_v[e])
          s += dt * len * k / d1;
        wv_lo[e] = 0.;
        wv_hi[e] = dir_v[e] ? 0. : 1.;
        }
      // north face: vertices (i, j+1) - (i+1, j+1); horizontal edge index (j+1)*N + i
      {
        const int a = (j + 1) * V + i, b = a + 1, e = (j + 1) * N + i;
        face_geo(vx[a], vy[a], vx[b], vy[b], g.cx, g.cy, len, d1);
        if (j < N - 1)
          {
            const CellGeo g2 = cell_geo(N, i, j + 1, vx, vy);
            const double  k2 = kc ? kc[morton(i, j + 1)] : k_scalar;
            face_geo(vx[a], vy[a], vx[b], vy[b], g2.cx, g2.cy, len, d2);
            const double c1 = k / d1, c2 = k2 / d2;
            tN[t]    = dt * len * c1 * c2 / (c1 + c2);
            wh_lo[e] = c1 / (c1 + c2);
            wh_hi[e] = c2 / (c1 + c2);
          }
        else
          {
            tN[t] = 0.;
            if (dir_h[e])
              s += dt * len * k / d1;
            wh_lo[e] = dir_h[e] ? 0. : 1.;
            wh_hi[e] = 0.;
          }
      }
      if (j == 0) // south boundary face: horizontal edge i
        {
          const int a = i, b = a + 1, e = i;
          face_geo(vx[a], vy[a], vx[b], vy[b], g.cx, g.cy, len, d1);
          if (dir_h[e])
            s += dt * len * k / d1;
          wh_lo[e] = 0.;
          wh_hi[e] = dir_h[e] ? 0. : 1.;
        }
      sig[t] = s;
    }

    // Galerkin coarse operator for piecewise-constant prolongation over 2x2 aggregates
    __global__ void
    k_cmg_coarsen(const int Nc, const double *__restrict__ tEf, const double *__restrict__ tNf,
                  const double *__restrict__ sigf, double *__restrict__ tEc, double *__restrict__ tNc,
                  double *__restrict__ sigc)
    {
      const int t = blockIdx.x * blockDim.x + threadIdx.x;
      if (t >= Nc * Nc)
        return;
      const int I = t % Nc, J = t / Nc, Nf = 2 * Nc;
      const int f00 = (2 * J) * Nf + 2 * I, f10 = f00 + 1, f01 = f00 + Nf, f11 = f01 + 1;
      tEc[t]  = tEf[f10] + tEf[f11];
      tNc[t]  = tNf[f01] + tNf[f11];
      sigc[t] = sigf[f00] + sigf[f10] + sigf[f01] + sigf[f11];
    }

    // (diag, A x) of the 5-point operator at cell (i, j); x may live in shared or global memory
    __device__ __forceinline__ void
    stencil(const int N, const int i, const int j, const double *__restrict__ tE, const double *__restrict__ tN,
            const double *__restrict__ sig, const double *x, double &diag, double &ax)
    {
      const int    c  = j * N + i;
      const double te = tE[c], tw = i > 0 ? tE[c - 1] : 0., tn = tN[c], ts = j > 0 ? tN[c - N] : 0.;
      diag            = sig[c] + te + tw + tn + ts;
      double a        = diag * x[c];
      if (i < N - 1)
        a -= te * x[c + 1];
      if (i > 0)
        a -= tw * x[c - 1];
      if (j < N - 1)
        a -= tn * x[c + N];
      if (j > 0)
        a -= ts * x[c - N];
      ax = a;
    }
    // the same with x' = x + over * xc(parent) evaluated on the fly (prolongation fused into the smoother)
    __device__ __forceinline__ void
    stencil_pc(const int N, const int i, const int j, const double *__restrict__ tE, const double *__restrict__ tN,
               const double *__restrict__ sig, const double *x, const double *xc, const double over, double &diag,
               double &ax, double &xself)
    {
      const int Nc = N >> 1;
      auto      xp = [&](const int ii, const int jj) { return x[jj * N + ii] + over * xc[(jj >> 1) * Nc + (ii >> 1)]; };
      const int    c  = j * N + i;
      const double te = tE[c], tw = i > 0 ? tE[c - 1] : 0., tn = tN[c], ts = j > 0 ? tN[c - N] : 0.;
      diag            = sig[c] + te + tw + tn + ts;
      xself           = xp(i, j);
      double a        = diag * xself;
      if (i < N - 1)
        a -= te * xp(i + 1, j);
      if (i > 0)
        a -= tw * xp(i - 1, j);
      if (j < N - 1)
        a -= tn * xp(i, j + 1);
      if (j > 0)
        a -= ts * xp(i, j - 1);
      ax = a;
    }

    // xo = xi + omega (b - A xi) / diag      (xi == nullptr: xo = omega b / diag)
    // xc != nullptr: xi is first corrected by the prolongated coarse solution (post-smoothing step 1)
    __global__ void
    k_cmg_jacobi(const int N, const double *__restrict__ tE, const double *__restrict__ tN,
                 const double *__restrict__ sig, const double *__restrict__ b, const double *__restrict__ xi,
                 const double *__restrict__ xc, const double over, double *__restrict__ xo, const double omega,
                 const int c_begin, const int c_end)
    {
      const int t = c_begin + blockIdx.x * blockDim.x + threadIdx.x;
      if (t >= c_end)
        return;
      const int i = t % N, j = t / N;
      double    diag, ax, xs;
      if (!xi)
        {
          const double te = tE[t], tw = i > 0 ? tE[t - 1] : 0., tn = tN[t], ts = j > 0 ? tN[t - N] : 0.;
          diag            = sig[t] + te + tw + tn + ts;
          xo[t]           = diag > 0. ? omega * b[t] / diag : 0.;
          return;
        }
      if (xc)
        stencil_pc(N, i, j, tE, tN, sig, xi, xc, over, diag, ax, xs);
      else
        {
          stencil(N, i, j, tE, tN, sig, xi, diag, ax);
          xs = xi[t];
        }
      xo[t] = diag > 0. ? xs + omega * (b[t] - ax) / diag : 0.;
    }

    // bc(I, J) = sum over the 4 children of (b - A x): residual and restriction in one pass
    __global__ void
    k_cmg_restrict(const int Nf, const double *__restrict__ tE, const double *__restrict__ tN,
                   const double *__restrict__ sig, const double *__restrict__ b, const double *__restrict__ x,
                   double *__restrict__ bc, const int c_begin, const int c_end)
    {
      const int Nc = Nf >> 1;
      const int t  = c_begin + blockIdx.x * blockDim.x + threadIdx.x;
      if (t >= c_end)
        return;
      const int I = t % Nc, J = t / Nc;
      double    s = 0.;
#pragma unroll
      for (int dj = 0; dj < 2; ++dj)
#pragma unroll
        for (int di = 0; di < 2; ++di)
          {
            const int i = 2 * I + di, j = 2 * J + dj;
            double    diag, ax;
            stencil(Nf, i, j, tE, tN, sig, x, diag, ax);
            s += b[j * Nf + i] - ax;
          }
      bc[t] = s;
    }

    // ---- all levels with N <= 32 in one CTA, vectors in shared memory ---------------------------------
    __global__ void __launch_bounds__(1024)
    k_cmg_coarse(const int N0, const double *__restrict__ tE, const double *__restrict__ tN,
                 const double *__restrict__ sig, const double *__restrict__ b_in, double *__restrict__ x_out,
                 const double omega, const double omega2, const double over, const int nu)
    {
      __shared__ double sb[kCoarseCells], sx[2][kCoarseCells];
      int               cur[6]; // which of sx[0], sx[1] holds the iterate of level m
      const int         t = threadIdx.x;
      for (int c = t; c < N0 * N0; c += blockDim.x)
        sb[c] = b_in[c];
      __syncthreads();
      int off = 0, m = 0, n = N0;
      // ---- down
      for (;; ++m)
        {
          const double *te = tE + off, *tn = tN + off, *sg = sig + off;
          const int     i = t % n, j = t / n;
          const bool    act = t < n * n;
          if (n == 1)
            {
              if (act)
                sx[0][off] = sg[0] > 0. ? sb[off] / sg[0] : 0.;
              cur[m] = 0;
              __syncthreads();
              break;
            }
          int w = 0;
          if (act)
            {
              const double tw = i > 0 ? te[t - 1] : 0., ts = j > 0 ? tn[t - n] : 0.;
              const double dg = sg[t] + te[t] + tw + tn[t] + ts;
              sx[0][off + t]  = dg > 0. ? omega * sb[off + t] / dg : 0.;
            }
          __syncthreads();
          for (int k = 1; k < nu; ++k)
            {
              if (act)
                {
                  double dg, ax;
                  stencil(n, i, j, te, tn, sg, sx[w] + off, dg, ax);
                  sx[w ^ 1][off + t] = dg > 0. ? sx[w][off + t] + ((k & 1) ? omega2 : omega) * (sb[off + t] - ax) / dg : 0.;
                }
              w ^= 1;
              __syncthreads();
            }
          cur[m] = w;
          const int nc = n >> 1;
          if (t < nc * nc)
            {
              const int I = t % nc, J = t / nc;
              double    s = 0.;
              for (int dj = 0; dj < 2; ++dj)
                for (int di = 0; di < 2; ++di)
                  {
                    double dg, ax;
                    stencil(n, 2 * I + di, 2 * J + dj, te, tn, sg, sx[w] + off, dg, ax);
                    s += sb[off + (2 * J + dj) * n + 2 * I + di] - ax;
                  }
              sb[off + n * n + t] = s;
            }
          __syncthreads();
          off += n * n;
          n = nc;
        }
      // ---- up: level m is solved; correct and post-smooth levels m-1 ... 0
      while (m > 0)
        {
          const int offc = off, wc = cur[m];
          --m;
          n   = n << 1;
          off -= n * n;
          const double *te = tE + off, *tn = tN + off, *sg = sig + off;
          const int     i = t % n, j = t / n;
          const bool    act = t < n * n;
          int           w   = cur[m];
          for (int k = 0; k < nu; ++k)
            {
              if (act)
                {
                  double dg, ax, xs;
                  if (k == 0)
                    stencil_pc(n, i, j, te, tn, sg, sx[w] + off, sx[wc] + offc, over, dg, ax, xs);
                  else
                    {
                      stencil(n, i, j, te, tn, sg, sx[w] + off, dg, ax);
                      xs = sx[w][off + t];
                    }
                  sx[w ^ 1][off + t] = dg > 0. ? xs + (((nu - 1 - k) & 1) ? omega2 : omega) * (sb[off + t] - ax) / dg : 0.;
                }
              w ^= 1;
              __syncthreads();
            }
          if (nu == 0)
            {
              if (act)
                sx[w][off + t] += over * sx[wc][offc + (j >> 1) * (n >> 1) + (i >> 1)];
              __syncthreads();
            }
          cur[m] = w;
        }
      for (int c = t; c < N0 * N0; c += blockDim.x)
        x_out[c] = sx[cur[0]][c];
    }

    // ---- face <-> cell transfer ----------------------------------------------------------------------
    __device__ __forceinline__ double
    owned_val(const double *__restrict__ x, const int d, const int64_t L)
    {
      return (d >= 0 && d < L) ? x[d] : 0.;
    }
    __global__ void
    k_cmg_gather(const int N, const CellBox bx, const int64_t L, const int32_t *__restrict__ hedge,
                 const int32_t *__restrict__ vedge, const double *__restrict__ wv_lo, const double *__restrict__ wv_hi,
                 const double *__restrict__ wh_lo, const double *__restrict__ wh_hi, const double *__restrict__ res,
                 double *__restrict__ rc)
    {
      const int bw = bx.i1 - bx.i0, q = blockIdx.x * blockDim.x + threadIdx.x;
      if (q >= bw * (bx.j1 - bx.j0))
        return;
      const int i = bx.i0 + q % bw, j = bx.j0 + q / bw, V = N + 1, t = j * N + i;
      const int ew = j * V + i, ee = ew + 1, es = j * N + i, en = es + N;
      rc[t] = wv_hi[ew] * owned_val(res, vedge[ew], L) + wv_lo[ee] * owned_val(res, vedge[ee], L) +
              wh_hi[es] * owned_val(res, hedge[es], L) + wh_lo[en] * owned_val(res, hedge[en], L);
    }
    __global__ void
    k_cmg_scatter(const int N, const CellBox bx, const int64_t L, const int32_t *__restrict__ hedge,
                  const int32_t *__restrict__ vedge, const double *__restrict__ wv_lo, const double *__restrict__ wv_hi,
                  const double *__restrict__ wh_lo, const double *__restrict__ wh_hi, const double *__restrict__ e,
                  const double *__restrict__ y1, double *__restrict__ y)
    {
      const int V = N + 1;
      const int t = blockIdx.x * blockDim.x + threadIdx.x;
      const int bw = bx.i1 - bx.i0, bh = bx.j1 - bx.j0;
      const int n_v = (bw + 1) * bh, n_h = bw * (bh + 1); // vertical / horizontal edges of the box
      int       d;
      double    c;
      if (t < n_v)
        {
          const int i = bx.i0 + t % (bw + 1), j = bx.j0 + t / (bw + 1), q = j * V + i;
          d           = vedge[q];
          c           = (i > 0 ? wv_lo[q] * e[j * N + i - 1] : 0.) + (i < N ? wv_hi[q] * e[j * N + i] : 0.);
        }
      else if (t < n_v + n_h)
        {
          const int u = t - n_v, i = bx.i0 + u % bw, j = bx.j0 + u / bw, q = j * N + i;
          d           = hedge[q];
          c           = (j > 0 ? wh_lo[q] * e[(j - 1) * N + i] : 0.) + (j < N ? wh_hi[q] * e[j * N + i] : 0.);
        }
      else
        return;
      if (d < 0 || d >= L)
        return;
      y[d] = y1[d] + c;
    }

    inline unsigned
    g1(const int n)
    {
      return (unsigned)((n + 127) / 128);
    }
  } // namespace

  int
  cmg_build_topology(CellMg &cm, const HostMesh &mesh, cudaStream_t s)
  {
    cm.ready = false;
    if (mesh.n_refine < 1)
      return 1;
    const int N = 1 << mesh.n_refine, V = N + 1;
    cm.N0       = N;
    int nl      = 0;
    for (int n = N; n >= 1; n /= 2)
      ++nl;
    cm.lv.clear();
    cm.lv.resize((size_t)nl);
    cm.n_fine = 0;
    for (int l = 0; l < nl; ++l)
      {
        cm.lv[(size_t)l].N = N >> l;
        if ((N >> l) > kCoarseN)
          cm.n_fine = l + 1;
      }
    cm.c_N = N >> cm.n_fine;
    // fine levels: own arrays; coarse levels: packed
    for (int l = 0; l < cm.n_fine; ++l)
      {
        CmgLevel    &L = cm.lv[(size_t)l];
        const size_t n = (size_t)L.N * L.N;
        if (L.tE.alloc(n) || L.tN.alloc(n) || L.sig.alloc(n) || L.b.alloc(n) || L.x0.alloc(n) || L.x1.alloc(n))
          return -1;
      }
    size_t packed = 0;
    for (int n = cm.c_N; n >= 1; n /= 2)
      packed += (size_t)n * n;
    if (cm.c_tE.alloc(packed) || cm.c_tN.alloc(packed) || cm.c_sig.alloc(packed))
      return -1;
    {
      // right-hand side / solution of the first coarse level (global memory interface of the coarse kernel)
      CmgLevel    &L = cm.lv[(size_t)cm.n_fine];
      const size_t n = (size_t)L.N * L.N;
      if (L.b.alloc(n) || L.x0.alloc(n))
        return -1;
    }
    // pressure-Dirichlet flags per edge (lexicographic cells: cell (ci, cj) has vertex 0 = cj*V + ci)
    std::vector<uint8_t> dv((size_t)V * N, 0), dh((size_t)N * V, 0);
    for (int64_t c = 0; c < mesh.n_cells; ++c)
      {
        const uint32_t v0 = mesh.cell_vertices[4 * (size_t)c];
        const int      ci = (int)(v0 % V), cj = (int)(v0 / V);
        const uint8_t *fk = &mesh.face_kinds[4 * (size_t)c];
        if (fk[0] & PE_FACE_DIR_P)
          dv[(size_t)cj * V + ci] = 1;
        if (fk[1] & PE_FACE_DIR_P)
          dv[(size_t)cj * V + ci + 1] = 1;
        if (fk[2] & PE_FACE_DIR_P)
          dh[(size_t)cj * N + ci] = 1;
        if (fk[3] & PE_FACE_DIR_P)
          dh[(size_t)(cj + 1) * N + ci] = 1;
      }
    if (cm.dir_v.alloc(dv.size()) || cm.dir_h.alloc(dh.size()) || cm.wv_lo.alloc(dv.size()) || cm.wv_hi.alloc(dv.size()) ||
        cm.wh_lo.alloc(dh.size()) || cm.wh_hi.alloc(dh.size()))
      return -1;
    cudaMemcpyAsync(cm.dir_v.p, dv.data(), dv.size(), cudaMemcpyHostToDevice, s);
    cudaMemcpyAsync(cm.dir_h.p, dh.data(), dh.size(), cudaMemcpyHostToDevice, s);
    if (cudaStreamSynchronize(s) != cudaSuccess)
      return -1;
    cm.ready = true;
    return 0;
  }

  int
  cmg_assemble(CellMg &cm, const double *vx, const double *vy, const double *kcell_morton, const double k_scalar,
               const double dt, const double gamma, cudaStream_t s, int64_t *launches)
  {
    // coefficient arrays of level l
    auto coef = [&](const int l, double *&tE, double *&tN, double *&sig) {
      if (l < cm.n_fine)
        {
          tE  = cm.lv[(size_t)l].tE.p;
          tN  = cm.lv[(size_t)l].tN.p;
          sig = cm.lv[(size_t)l].sig.p;
          return;
        }
      size_t off = 0;
      for (int q = cm.n_fine; q < l; ++q)
        off += (size_t)cm.lv[(size_t)q].N * cm.lv[(size_t)q].N;
      tE  = cm.c_tE.p + off;
      tN  = cm.c_tN.p + off;
      sig = cm.c_sig.p + off;
    };
    double *tE, *tN, *sig;
    coef(0, tE, tN, sig);
    const int N = cm.N0;
    k_cmg_assemble0<<<g1(N * N), 128, 0, s>>>(N, vx, vy, kcell_morton, k_scalar, dt, gamma, cm.dir_v.p, cm.dir_h.p, tE, tN, sig,
                                              cm.wv_lo.p, cm.wv_hi.p, cm.wh_lo.p, cm.wh_hi.p);
    ++*launches;
    for (size_t l = 1; l < cm.lv.size(); ++l)
      {
        double *tEc, *tNc, *sigc;
        coef((int)l, tEc, tNc, sigc);
        const int Nc = cm.lv[l].N;
        k_cmg_coarsen<<<g1(Nc * Nc), 128, 0, s>>>(Nc, tE, tN, sig, tEc, tNc, sigc);
        ++*launches;
        tE  = tEc;
        tN  = tNc;
        sig = sigc;
      }
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
  }

  void
  cmg_gather(CellMg &cm, const int64_t L, const int32_t *hedge, const int32_t *vedge, const double *res, const CellBox &box,
             double *out, cudaStream_t s)
  {
    const int N = cm.N0, n = (box.i1 - box.i0) * (box.j1 - box.j0);
    if (n > 0)
      k_cmg_gather<<<g1(n), 128, 0, s>>>(N, box, L, hedge, vedge, cm.wv_lo.p, cm.wv_hi.p, cm.wh_lo.p, cm.wh_hi.p, res, out);
  }

  void
  cmg_scatter(CellMg &cm, const int64_t L, const int32_t *hedge, const int32_t *vedge, const double *e, const double *y1,
              double *y, const CellBox &box, cudaStream_t s)
  {
    const int N = cm.N0, bw = box.i1 - box.i0, bh = box.j1 - box.j0, n = (bw + 1) * bh + bw * (bh + 1);
    if (n > 0)
      k_cmg_scatter<<<g1(n), 128, 0, s>>>(N, box, L, hedge, vedge, cm.wv_lo.p, cm.wv_hi.p, cm.wh_lo.p, cm.wh_hi.p, e, y1, y);
  }

  void
  cmg_set_comm(CellMg &cm, const GridComm &gc)
  {
    cm.gc     = gc;
    cm.n_dist = 0;
    cm.bounds.clear();
    if (gc.world <= 1 || !gc.exchange_rows || !gc.allgather_rows)
      return;
    const int W = gc.world;
    for (int l = 0; l < cm.n_fine; ++l)
      {
        const int N = cm.lv[(size_t)l].N;
        // a distributed level pays ~6 NCCL row exchanges (~45 us each, measured on 8 GPUs); below ~1024^2 the
        // replicated sweep is cheaper than that latency
        if (N % (2 * W) != 0 || N / W < 8 || N < gc.dist_min_n)
          break;
        std::vector<int64_t> bd((size_t)W + 1);
        for (int r = 0; r <= W; ++r)
          bd[(size_t)r] = (int64_t)N * r / W;
        cm.bounds.push_back(bd);
        ++cm.n_dist;
      }
  }

  const double *
  cmg_vcycle(CellMg &cm, cudaStream_t s, int64_t *launches)
  {
    const int     nf = cm.n_fine;
    const double  om = cm.omega, om2 = cm.omega2, ov = cm.over;
    const int     nu = std::max(1, cm.nu);
    std::vector<double *> cur((size_t)nf + 1, nullptr);
    // slab of this rank on level l (all rows on replicated levels)
    auto slab = [&](const int l, int64_t &ra, int64_t &rb) {
      const bool dist = l < cm.n_dist;
      ra              = dist ? cm.bounds[(size_t)l][(size_t)cm.gc.rank] : 0;
      rb              = dist ? cm.bounds[(size_t)l][(size_t)cm.gc.rank + 1] : cm.lv[(size_t)l].N;
      return dist;
    };
    // ---- down
    for (int l = 0; l < nf; ++l)
      {
        CmgLevel &L = cm.lv[(size_t)l];
        int64_t   ra, rb;
        const bool dist = slab(l, ra, rb);
        const int  c0 = (int)(ra * L.N), c1 = (int)(rb * L.N), n = c1 - c0;
        auto exchange = [&](double *p) {
          if (dist)
            cm.gc.exchange_rows(cm.gc.ctx, p, L.N, 1, 0, ra, rb);
        };
        double *x0 = L.x0.p, *x1 = L.x1.p;
        k_cmg_jacobi<<<g1(n), 128, 0, s>>>(L.N, L.tE.p, L.tN.p, L.sig.p, L.b.p, nullptr, nullptr, 0., x0, om, c0, c1);
        for (int k = 1; k < nu; ++k)
          {
            exchange(x0);
            k_cmg_jacobi<<<g1(n), 128, 0, s>>>(L.N, L.tE.p, L.tN.p, L.sig.p, L.b.p, x0, nullptr, 0., x1, (k & 1) ? om2 : om, c0, c1);
            std::swap(x0, x1);
          }
        exchange(x0);
        cur[(size_t)l] = x0;
        CmgLevel &C    = cm.lv[(size_t)l + 1];
        const int64_t ca = ra / 2, cb = rb / 2; // slab bounds are even
        k_cmg_restrict<<<g1((int)((cb - ca) * C.N)), 128, 0, s>>>(L.N, L.tE.p, L.tN.p, L.sig.p, L.b.p, x0, C.b.p,
                                                                  (int)(ca * C.N), (int)(cb * C.N));
        *launches += nu + 1;
        if (dist && l + 1 >= cm.n_dist)
          {
            // first replicated level: complete right-hand side on every rank
            std::vector<int64_t> cbnd((size_t)cm.gc.world + 1);
            for (int r = 0; r <= cm.gc.world; ++r)
              cbnd[(size_t)r] = cm.bounds[(size_t)l][(size_t)r] / 2;
            cm.gc.allgather_rows(cm.gc.ctx, C.b.p, C.N, 1, 0, cbnd.data());
          }
      }
    // ---- coarse levels: one launch
    {
      CmgLevel &C = cm.lv[(size_t)nf];
      k_cmg_coarse<<<1, 1024, 0, s>>>(C.N, cm.c_tE.p, cm.c_tN.p, cm.c_sig.p, C.b.p, C.x0.p, om, om2, ov, nu);
      ++*launches;
      cur[(size_t)nf] = C.x0.p;
    }
    // ---- up
    for (int l = nf - 1; l >= 0; --l)
      {
        CmgLevel &L = cm.lv[(size_t)l];
        int64_t   ra, rb;
        const bool dist = slab(l, ra, rb);
        const int  c0 = (int)(ra * L.N), c1 = (int)(rb * L.N), n = c1 - c0;
        auto exchange = [&](double *p) {
          if (dist)
            cm.gc.exchange_rows(cm.gc.ctx, p, L.N, 1, 0, ra, rb);
        };
        double *x0 = cur[(size_t)l], *x1 = x0 == L.x0.p ? L.x1.p : L.x0.p;
        // x0 still carries its halo rows from the way down; the coarse result carries its own (below)
        k_cmg_jacobi<<<g1(n), 128, 0, s>>>(L.N, L.tE.p, L.tN.p, L.sig.p, L.b.p, x0, cur[(size_t)l + 1], ov, x1,
                                           ((nu - 1) & 1) ? om2 : om, c0, c1);
        std::swap(x0, x1);
        for (int k = 1; k < nu; ++k)
          {
            exchange(x0);
            k_cmg_jacobi<<<g1(n), 128, 0, s>>>(L.N, L.tE.p, L.tN.p, L.sig.p, L.b.p, x0, nullptr, 0., x1,
                                               ((nu - 1 - k) & 1) ? om2 : om, c0, c1);
            std::swap(x0, x1);
          }
        cur[(size_t)l] = x0;
        *launches += nu;
        if (dist)
          {
            if (l == 0 && cm.gc.slabs_to_boxes)
              cm.gc.slabs_to_boxes(cm.gc.ctx, x0, L.N, 1, 0, cm.bounds[0].data(), 1);
            else if (l == 0)
              cm.gc.allgather_rows(cm.gc.ctx, x0, L.N, 1, 0, cm.bounds[0].data());
            else
              exchange(x0); // the finer level corrects from my rows and one halo row on each side
          }
      }
    return cur[0];
  }
} // namespace pe
