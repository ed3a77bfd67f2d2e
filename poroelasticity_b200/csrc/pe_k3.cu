// Three-field solver matrix K3 (see pe_k3.h): structure from the reference pattern on the device,
// values by a gather from the second assembly pass.
#include "pe_k3.h"
#include "pe_solver.h"

#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <vector>

namespace pe
{
  namespace
  {
    constexpr uint32_t kNoSrc = 0xffffffffu;

    struct Ranges
    {
      int64_t L, H, Lx, n_u, pint_begin, pint_end, hu_end, hpi_end;
    };
    // 0: displacement, 1: interior pressure, 2: face pressure   (local column id: owned < L <= halo)
    __device__ __forceinline__ int
    col_class(const int32_t c, const Ranges &R)
    {
      if (c < R.L)
        return c < R.n_u ? 0 : (c < R.pint_end ? 1 : 2);
      return c < R.hu_end ? 0 : (c < R.hpi_end ? 1 : 2);
    }
    // physical index of the xi unknown that belongs to interior-pressure column c
    __device__ __forceinline__ int32_t
    xi_of(const int32_t c, const Ranges &R)
    {
      return (int32_t)(c < R.L ? R.L + R.H + (c - R.pint_begin) : R.L + R.H + R.Lx + (c - R.hu_end));
    }

    // pass 0 (colidx3 == nullptr): row lengths.  pass 1: columns, sources and the special positions.
    __global__ void
    k_k3_structure(const Ranges R, const int64_t *__restrict__ rowptr, const int32_t *__restrict__ colidx,
                   int64_t *__restrict__ cnt, const int64_t *__restrict__ rowptr3, int32_t *__restrict__ colidx3,
                   uint32_t *__restrict__ src, int64_t *__restrict__ pos)
    {
      const int64_t r3 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
      if (r3 >= R.L + R.Lx)
        return;
      const bool    xi_row = r3 >= R.L;
      const int64_t r      = xi_row ? R.pint_begin + (r3 - R.L) : r3;
      const int     rc     = xi_row ? 3 : (r < R.n_u ? 0 : (r < R.pint_end ? 1 : 2));
      const bool    fill   = colidx3 != nullptr;
      int64_t       o      = fill ? rowptr3[r3] : 0;
      const int64_t o0     = o;
      for (int64_t k = rowptr[r]; k < rowptr[r + 1]; ++k)
        {
          const int32_t c  = colidx[k];
          const int     cc = col_class(c, R);
          bool          keep;
          int32_t       c3 = c;
          if (rc == 0)
            {
              keep = cc != 2; // u row: A_mu and -B^T (on xi); the u <-> p_face block is structurally zero
              if (cc == 1)
                c3 = xi_of(c, R);
            }
          else if (rc == 3)
            keep = cc == 0; // xi row: -B
          else
            keep = cc != 0; // pressure rows: -(C + ...)
          if (!keep)
            continue;
          if (fill)
            {
              colidx3[o] = c3;
              src[o]     = (uint32_t)k;
              if (rc == 1 && c == r)
                pos[r - R.pint_begin] = o; // (p_int, p_int) diagonal
            }
          ++o;
        }
      if (rc == 1)
        {
          if (fill)
            {
              colidx3[o]                      = xi_of((int32_t)r, R);
              src[o]                          = kNoSrc;
              pos[R.Lx + (r - R.pint_begin)] = o; // (p_int, xi)
            }
          ++o;
        }
      if (rc == 3)
        {
          const int64_t i = r3 - R.L;
          if (fill)
            {
              colidx3[o]         = (int32_t)(R.L + R.H + i);
              src[o]             = kNoSrc;
              pos[2 * R.Lx + i]  = o; // (xi, xi)
              colidx3[o + 1]     = (int32_t)r;
              src[o + 1]         = kNoSrc;
              pos[3 * R.Lx + i]  = o + 1; // (xi, p_int)
            }
          o += 2;
        }
      if (!fill)
        cnt[r3] = o - o0;
    }

    __global__ void
    k_k3_gather(const int64_t nnz, const int64_t nnz_u, const uint32_t *__restrict__ src,
                const double *__restrict__ values0, double *__restrict__ val)
    {
      const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
      if (k >= nnz)
        return;
      const uint32_t s = src[k];
      double         v = 0.;
      if (s != kNoSrc)
        {
          v = __ldg(values0 + s);
          if (k >= nnz_u)
            v = -v;
        }
      val[k] = v;
    }
    __global__ void
    k_k3_special(const int64_t Lx, const int64_t *__restrict__ pos, const double *__restrict__ w,
                 const double lambda_s, const double alpha, double *__restrict__ val)
    {
      const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
      if (i >= Lx)
        return;
      const double wl = w[i] / lambda_s;
      val[pos[i]] -= alpha * alpha * wl;
      val[pos[Lx + i]]     = alpha * wl;
      val[pos[2 * Lx + i]] = -wl;
      val[pos[3 * Lx + i]] = alpha * wl;
    }

    // GMRES path: the xi rows are scaled by T_i = tau / sqrt(w_i), tau = 1 / (1/lambda_s + 1/(2 mu)) (= the xi block of
    // the preconditioner times sqrt(w)), so that the 2-norm of the Krylov residual weighs the xi equation like its
    // effect on the residual of the reference system,
    //     r_u(two-field) = r_u - lambda B^T W^-1 r_xi = r_u - (1 + lambda / 2 mu) (B^T W^-1/2) (T r_xi),   ||B^T W^-1/2|| = O(1)
    // and the scaled row keeps O(mu) entries for every lambda (tau -> lambda for lambda << mu, -> 2 mu for lambda >> mu;
    // scaling by lambda itself makes the xi - u coupling of the preconditioned operator O(lambda): GMRES stalls).
    __global__ void
    k_k3_scale_xi_rows(const int64_t L, const int64_t Lx, const int64_t *__restrict__ rowptr3, const double *__restrict__ w,
                       const double tau, double *__restrict__ val)
    {
      const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
      if (i >= Lx)
        return;
      const double T = tau * rsqrt(w[i]);
      for (int64_t k = rowptr3[L + i]; k < rowptr3[L + i + 1]; ++k)
        val[k] *= T;
    }
    // GMRES path: pressure row r (interior and face) scaled by S_r = sqrt(2 mu * dinv_r), dinv_r = 1 / (diagonal of the
    // pressure block of the preconditioner).  With the xi rows scaled by tau / sqrt(w) every row of the Krylov operator
    // is then measured in the units of the displacement rows (the 2-norm GMRES minimises approximates the P^-1 norm of
    // MINRES).  Without it the pressure rows of a stiff, nearly impermeable medium (mu ~ 1e5, dt K ~ 1e-9) are 1e12
    // times smaller than the rows they couple to and GMRES(m) stagnates on residuals that live in them.
    __global__ void
    k_k3_p_scale(const int64_t n, const int64_t pint_begin, const double *__restrict__ dinv, const double two_mu,
                 double *__restrict__ ps)
    {
      const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
      if (i < n)
        {
          const double d = dinv[pint_begin + i];
          ps[i]          = d > 0. ? sqrt(two_mu * d) : 1.;
        }
    }
    __global__ void
    k_k3_scale_p_rows(const int64_t n, const int64_t pint_begin, const int64_t *__restrict__ rowptr3,
                      const double *__restrict__ ps, double *__restrict__ val)
    {
      const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
      if (i >= n)
        return;
      const double S = ps[i];
      for (int64_t k = rowptr3[pint_begin + i]; k < rowptr3[pint_begin + i + 1]; ++k)
        val[k] *= S;
    }
    // v[p rows] *= ps  (inverse = 1: v /= ps, written to out)
    __global__ void
    k_k3_p_vec_scale(const int64_t n, const int64_t pint_begin, const double *__restrict__ ps, const int inverse,
                     const double *v, double *out)
    {
      const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
      if (i < n)
        out[pint_begin + i] = inverse ? v[pint_begin + i] / ps[i] : v[pint_begin + i] * ps[i];
    }
    __global__ void
    k_k3_rhs(const Ranges R, const double *__restrict__ rhs, double *__restrict__ b3)
    {
      const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
      if (i >= R.L + R.H + R.Lx)
        return;
      b3[i] = i < R.n_u ? rhs[i] : (i < R.L ? -rhs[i] : 0.);
    }
    __global__ void
    k_k3_rhs_con(const int64_t n, const int32_t *__restrict__ rows, const double *__restrict__ rhs,
                 const int64_t *__restrict__ rowptr, const int32_t *__restrict__ colidx,
                 const double *__restrict__ values_a, const double *__restrict__ diag0, double *__restrict__ b3)
    {
      const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
      if (i >= n)
        return;
      const int32_t r = rows[i];
      double        d = 0.;
      for (int64_t k = rowptr[r]; k < rowptr[r + 1]; ++k)
        if (colidx[k] == r)
          d = values_a[k];
      // the row reads d u_r = d g_r in the reference matrix and d0 u_r = d0 g_r in K3
      if (d != 0.)
        b3[r] = rhs[r] * (diag0[r] / d);
    }
    // inverse of the SPD block diagonal of K3 (fallback preconditioner for meshes without a hierarchy)
    __global__ void
    k_k3_pinv(const Ranges R, const double *__restrict__ diag0, const double *__restrict__ w, const double lambda_s,
              const double alpha, const double mu, double *__restrict__ pinv3)
    {
      const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
      if (i >= R.L + R.H + R.Lx)
        return;
      double d;
      if (i < R.L)
        {
          d = fabs(diag0[i]);
          if (i >= R.pint_begin && i < R.pint_end)
            d += alpha * alpha * w[i - R.pint_begin] / lambda_s;
        }
      else if (i < R.L + R.H)
        d = 1.;
      else
        d = (1. / lambda_s + 1. / (2. * mu)) * w[i - R.L - R.H];
      pinv3[i] = d != 0. ? 1. / d : 1.;
    }
    // xi_i = (lambda_s / w_i) sum_{u cols of the xi row} K3(xi_i, c) x_c  +  alpha p_int,i
    __global__ void
    k_k3_init_xi(const Ranges R, const int64_t *__restrict__ rowptr3, const int32_t *__restrict__ colidx3,
                 const double *__restrict__ val, const double *__restrict__ w, const double lambda_s, const double alpha,
                 const double xi_tau /* 0: rows not scaled */, double *__restrict__ x3)
    {
      const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
      if (i >= R.Lx)
        return;
      const int64_t b = rowptr3[R.L + i], e = rowptr3[R.L + i + 1] - 2;
      double        s = 0.;
      for (int64_t k = b; k < e; ++k)
        s += val[k] * x3[colidx3[k]];
      // rows scaled by T = tau / sqrt(w): (lambda_s / w) / T = (lambda_s / tau) / sqrt(w)
      x3[R.L + R.H + i] = (xi_tau > 0. ? (lambda_s / xi_tau) * rsqrt(w[i]) : lambda_s / w[i]) * s + alpha * x3[R.pint_begin + i];
    }
    __global__ void
    k_k3_precond_xi(const int64_t Lx, const double *__restrict__ w, const double scale, const int xi_scaled,
                    const double *__restrict__ r, double *__restrict__ z)
    {
      const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
      if (i < Lx)
        // scaled rows: z = P_xi^-1 T^-1 r = r / sqrt(w)   (T = scale / sqrt(w), P_xi = w / scale)
        z[i] = xi_scaled ? r[i] * rsqrt(w[i]) : r[i] * (scale / w[i]);
    }

    inline unsigned
    gb(const int64_t n, const int b = 256)
    {
      return (unsigned)std::max<int64_t>(1, (n + b - 1) / b);
    }
    Ranges
    ranges_of(const K3System &k)
    {
      return Ranges{k.L, k.H, k.Lx, k.n_u, k.pint_begin, k.pint_end, k.hu_end, k.hpi_end};
    }
  } // namespace

  int
  k3_build_structure(K3System &k, const HostMesh &mesh, const HostDofs &dofs, const Partition &part,
                     const int64_t *d_rowptr, const int32_t *d_colidx, cudaStream_t s)
  {
    k.built = k.values_valid = false;
    k.L          = part.n_owned;
    k.H          = (int64_t)part.ghost_cols.size();
    k.n_u        = part.local_off[1];
    k.pint_begin = part.local_off[1];
    k.pint_end   = part.local_off[2];
    k.Lx         = k.pint_end - k.pint_begin;
    // halo columns are sorted by global id: [u | p_int | p_face]
    const uint32_t g_u = (uint32_t)dofs.n_u, g_pi = (uint32_t)(dofs.n_u + dofs.n_pint);
    k.hu_end  = k.L + (std::lower_bound(part.ghost_cols.begin(), part.ghost_cols.end(), g_u) - part.ghost_cols.begin());
    k.hpi_end = k.L + (std::lower_bound(part.ghost_cols.begin(), part.ghost_cols.end(), g_pi) - part.ghost_cols.begin());
    k.Hx      = k.hpi_end - k.hu_end;
    if (k.n_phys() > 0x7fffffffll)
      return 1;
    const Ranges  R  = ranges_of(k);
    const int64_t n3 = k.n_rows();
    DevBuf<int64_t> cnt;
    if (cnt.alloc((size_t)n3 + 1) != cudaSuccess || k.rowptr.alloc((size_t)n3 + 1) != cudaSuccess)
      return -1;
    cudaMemsetAsync(cnt.p, 0, sizeof(int64_t) * (size_t)(n3 + 1), s);
    k_k3_structure<<<gb(n3, 128), 128, 0, s>>>(R, d_rowptr, d_colidx, cnt.p, nullptr, nullptr, nullptr, nullptr);
    {
      size_t tmp_bytes = 0;
      cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, cnt.p, k.rowptr.p, (int64_t)(n3 + 1), s);
      DevBuf<uint8_t> tmp;
      if (tmp.alloc(tmp_bytes + 16) != cudaSuccess)
        return -1;
      if (cub::DeviceScan::ExclusiveSum(tmp.p, tmp_bytes, cnt.p, k.rowptr.p, (int64_t)(n3 + 1), s) != cudaSuccess)
        return -1;
      if (cudaStreamSynchronize(s) != cudaSuccess)
        return -1;
    }
    std::vector<int64_t> rp((size_t)n3 + 1);
    if (cudaMemcpy(rp.data(), k.rowptr.p, sizeof(int64_t) * rp.size(), cudaMemcpyDeviceToHost) != cudaSuccess)
      return -1;
    k.nnz   = rp[(size_t)n3];
    k.nnz_u = rp[(size_t)k.n_u];
    if (k.colidx.alloc((size_t)k.nnz) != cudaSuccess || k.src.alloc((size_t)k.nnz) != cudaSuccess ||
        k.val.alloc((size_t)k.nnz) != cudaSuccess || k.pos.alloc((size_t)4 * k.Lx) != cudaSuccess)
      return -1;
    k_k3_structure<<<gb(n3, 128), 128, 0, s>>>(R, d_rowptr, d_colidx, nullptr, k.rowptr.p, k.colidx.p, k.src.p, k.pos.p);
    {
      std::vector<int32_t> blk;
      build_row_blocks(n3, rp.data(), blk);
      k.n_rowblk = (int64_t)blk.size() - 1;
      if (k.rowblk.alloc(blk.size()) != cudaSuccess)
        return -1;
      cudaMemcpyAsync(k.rowblk.p, blk.data(), sizeof(int32_t) * blk.size(), cudaMemcpyHostToDevice, s);
      cudaStreamSynchronize(s);
    }
    // weights w_K of the owned cells (|K| = det J at the cell centre: BR1new:867, CGQ1:1341) and the
    // owned displacement rows that carry a Dirichlet constraint
    const int n_loc = dofs.n_loc;
    std::vector<double> w((size_t)k.Lx, 0.);
    for (int64_t c = part.cell_begin; c < part.cell_end; ++c)
      {
        const uint32_t *cv = &mesh.cell_vertices[4 * (size_t)c];
        const double    x0 = mesh.x[cv[0]], y0 = mesh.y[cv[0]], x1 = mesh.x[cv[1]], y1 = mesh.y[cv[1]];
        const double    x2 = mesh.x[cv[2]], y2 = mesh.y[cv[2]], x3 = mesh.x[cv[3]], y3 = mesh.y[cv[3]];
        const int64_t   l  = part.global_to_owned(dofs.cell_dofs[(size_t)c * n_loc + n_loc - 1]);
        if (l < k.pint_begin || l >= k.pint_end)
          return 2;
        w[(size_t)(l - k.pint_begin)] = 0.5 * ((x3 - x0) * (y2 - y1) - (x2 - x1) * (y3 - y0));
      }
    std::vector<uint8_t> vflag;
    if (dofs.element == PE_CGQ1_WG)
      {
        vflag.assign((size_t)mesh.n_vertices, 0);
        for (int64_t c = 0; c < mesh.n_cells; ++c)
          for (int f = 0; f < 4; ++f)
            if (mesh.face_kinds[4 * c + f] & PE_FACE_DIR_U)
              for (int v = 0; v < 2; ++v)
                vflag[mesh.cell_vertices[4 * c + kFaceVertex[f][v]]] = 1;
      }
    std::vector<int32_t> con;
    auto add = [&](const uint32_t g) {
      const int64_t l = part.global_to_owned(g);
      if (l >= 0 && l < k.n_u)
        con.push_back((int32_t)l);
    };
    for (int64_t c = 0; c < mesh.n_cells; ++c)
      {
        const uint32_t *cd = &dofs.cell_dofs[(size_t)c * n_loc];
        for (int f = 0; f < 4; ++f)
          if (mesh.face_kinds[4 * c + f] & PE_FACE_DIR_U)
            {
              for (int v = 0; v < 2; ++v)
                {
                  add(cd[2 * kFaceVertex[f][v]]);
                  add(cd[2 * kFaceVertex[f][v] + 1]);
                }
              if (dofs.element == PE_BR1_WG)
                add(cd[8 + 2 * f]);
            }
        if (!vflag.empty())
          for (int v = 0; v < 4; ++v)
            if (vflag[mesh.cell_vertices[4 * c + v]])
              {
                add(cd[2 * v]);
                add(cd[2 * v + 1]);
              }
      }
    std::sort(con.begin(), con.end());
    con.erase(std::unique(con.begin(), con.end()), con.end());
    k.n_con = (int64_t)con.size();
    if (k.w.alloc(w.size()) != cudaSuccess || k.con_rows.alloc(con.size()) != cudaSuccess ||
        k.diag0.alloc((size_t)(k.L + k.H)) != cudaSuccess)
      return -1;
    if (!w.empty())
      cudaMemcpyAsync(k.w.p, w.data(), sizeof(double) * w.size(), cudaMemcpyHostToDevice, s);
    if (!con.empty())
      cudaMemcpyAsync(k.con_rows.p, con.data(), sizeof(int32_t) * con.size(), cudaMemcpyHostToDevice, s);
    if (cudaStreamSynchronize(s) != cudaSuccess)
      return -1;
    k.built = true;
    return 0;
  }

  int
  k3_fill_values(K3System &k, const double lambda_s, const double alpha, const double mu, const bool scale_xi_rows,
                 cudaStream_t s, int64_t *launches)
  {
    k.lambda_s  = lambda_s;
    k.alpha     = alpha;
    k.xi_scaled = scale_xi_rows;
    k.p_scaled  = false;
    k.xi_tau    = scale_xi_rows ? 1. / (1. / lambda_s + 1. / (2. * mu)) : 0.;
    if (k.nnz > 0)
      k_k3_gather<<<gb(k.nnz), 256, 0, s>>>(k.nnz, k.nnz_u, k.src.p, k.values0.p, k.val.p);
    if (k.Lx > 0)
      k_k3_special<<<gb(k.Lx), 256, 0, s>>>(k.Lx, k.pos.p, k.w.p, lambda_s, alpha, k.val.p);
    *launches += 2;
    if (scale_xi_rows && k.Lx > 0)
      {
        k_k3_scale_xi_rows<<<gb(k.Lx), 256, 0, s>>>(k.L, k.Lx, k.rowptr.p, k.w.p, k.xi_tau, k.val.p);
        ++*launches;
      }
    k.values_valid = true;
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
  }

  int
  k3_make_rhs(const K3System &k, const double *rhs, const int64_t *rowptr, const int32_t *colidx, const double *values_a,
              double *b3, cudaStream_t s)
  {
    const Ranges R = ranges_of(k);
    cudaMemsetAsync(b3, 0, sizeof(double) * (size_t)k.n_phys(), s);
    k_k3_rhs<<<gb(k.L + k.H + k.Lx), 256, 0, s>>>(R, rhs, b3);
    if (k.n_con > 0)
      k_k3_rhs_con<<<gb(k.n_con), 256, 0, s>>>(k.n_con, k.con_rows.p, rhs, rowptr, colidx, values_a, k.diag0.p, b3);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
  }

  int
  k3_scale_p_rows(K3System &k, const double *dinv, const double mu, cudaStream_t s, int64_t *launches)
  {
    const int64_t n = k.L - k.pint_begin;
    k.p_scaled      = false;
    if (n <= 0)
      return 0;
    if (k.p_scale.n < (size_t)n && k.p_scale.alloc((size_t)n) != cudaSuccess)
      return -1;
    k_k3_p_scale<<<gb(n), 256, 0, s>>>(n, k.pint_begin, dinv, 2. * mu, k.p_scale.p);
    k_k3_scale_p_rows<<<gb(n), 256, 0, s>>>(n, k.pint_begin, k.rowptr.p, k.p_scale.p, k.val.p);
    *launches += 2;
    k.p_scaled = true;
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
  }
  // out[p rows] = v[p rows] * S (inverse: / S); other entries of out are not touched
  void
  k3_p_vec_scale(const K3System &k, const bool inverse, const double *v, double *out, cudaStream_t s)
  {
    const int64_t n = k.L - k.pint_begin;
    if (n > 0 && k.p_scaled)
      k_k3_p_vec_scale<<<gb(n), 256, 0, s>>>(n, k.pint_begin, k.p_scale.p, inverse ? 1 : 0, v, out);
  }

  int
  k3_make_pinv(const K3System &k, const double mu, double *pinv3, cudaStream_t s)
  {
    cudaMemsetAsync(pinv3, 0, sizeof(double) * (size_t)k.n_phys(), s);
    k_k3_pinv<<<gb(k.L + k.H + k.Lx), 256, 0, s>>>(ranges_of(k), k.diag0.p, k.w.p, k.lambda_s, k.alpha, mu, pinv3);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
  }

  int
  k3_init_xi(const K3System &k, double *x3, cudaStream_t s)
  {
    if (k.Lx > 0)
      k_k3_init_xi<<<gb(k.Lx, 128), 128, 0, s>>>(ranges_of(k), k.rowptr.p, k.colidx.p, k.val.p, k.w.p, k.lambda_s, k.alpha,
                                                 k.xi_tau, x3);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
  }

  int
  k3_precond_xi(const K3System &k, const double mu, const double *r3, double *z3, cudaStream_t s)
  {
    const double scale = 1. / (1. / k.lambda_s + 1. / (2. * mu));
    if (k.Lx > 0)
      k_k3_precond_xi<<<gb(k.Lx), 256, 0, s>>>(k.Lx, k.w.p, scale, k.xi_scaled ? 1 : 0, r3 + k.L + k.H, z3 + k.L + k.H);
    return 0;
  }
} // namespace pe
