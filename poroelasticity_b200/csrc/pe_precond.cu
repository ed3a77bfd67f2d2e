// Block preconditioner of the MINRES solve (new work, see pe_mg.cu for the multigrid part):
//
//   P = diag(P_u, P_p),  both symmetric positive definite.
//
//   P_u^{-1}: symmetric multiplicative two-space method on A_uu, spaces = {face bubbles} + {vertex
//             dofs = Q1}: bubble Jacobi -> Q1 V-cycle on the updated residual -> bubble Jacobi.
//             (CGQ1 has no bubbles: P_u^{-1} is the V-cycle alone.)
//   P_p^{-1}: exact block elimination of the interior pressures (their block is diagonal) from
//             S^ = c0 M + dt S + gamma M_int, gamma = alpha^2/(lambda+2 mu):
//               F = S^_ff - S^_fi D_i^{-1} S^_if  on the face pressures,
//               F^{-1} ~ D_F^{-1} + Pi_f V_p Pi_f^T   (scalar Q1 reaction-diffusion V-cycle),
//             then back-substitution for p_int.
// Measured on the oracle's matrices (scripts/precond_study.py): kappa(P_u^{-1} A_uu) ~ 3.2,
// kappa(P_F^{-1} F) ~ 3.0, S^ vs the exact Schur complement 1.12; MINRES needs ~3x fewer iterations
// than with the additive Jacobi + V-cycle preconditioner.
//
// The sub-blocks of the system matrix that the method multiplies with (A_bu, A_vb, S_if, S_fi) are
// index lists into the assembled CSR values, built once per mesh on the host.
#include "pe_internal.h"
#include "pe_mg.h"
#include "pe_precond.h"
#include "pe_timer.h"

#include <algorithm>
#include <climits>
#include <functional>

namespace pe
{
  namespace
  {
    enum
    {
      CLS_VERTEX = 0,
      CLS_BUBBLE = 1,
      CLS_PINT   = 2,
      CLS_PFACE  = 3
    };

    // compact copy of the sub-block values, once per solve: val[k] = values[src[k]]
    __global__ void
    k_sub_gather_values(const int64_t n, const uint32_t *__restrict__ src, const uint8_t *__restrict__ neg,
                        const double *__restrict__ values, SubCsr::value_t *__restrict__ val)
    {
      const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
      if (k < n)
        {
          const double v = __ldg(values + src[k]);
          val[k]         = (SubCsr::value_t)((neg && neg[k]) ? -v : v);
        }
    }
    // on the owned face rows: y1 = omega_F dinvF g  (first Jacobi sweep on F from zero)
    __global__ void
    k_face_scale(const int n, const int32_t *__restrict__ rows, const double *__restrict__ wdinv,
                 const double *__restrict__ g, double *__restrict__ y)
    {
      const int r = blockIdx.x * blockDim.x + threadIdx.x;
      if (r >= n)
        return;
      const int row = rows[r];
      y[row]        = wdinv[row] * g[row];
    }
    __global__ void
    k_face_wdinv(const int n, const int32_t *__restrict__ rows, const double *__restrict__ dinv, const double omega,
                 double *__restrict__ wdinv)
    {
      const int r = blockIdx.x * blockDim.x + threadIdx.x;
      if (r >= n)
        return;
      wdinv[rows[r]] = omega * dinv[rows[r]];
    }
    // streamed sub-block product (same scheme as k_spmv_stream): a CTA owns a run of sub-rows holding
    // <= kSubNnz entries; coalesced value/column streaming with all loads batched, row sums from smem,
    // epilogue  out[row] = a[row] + m[row] * (b[row] - sum)   (null a -> 0, null m -> 1, null b -> +sum)
    constexpr int kSubNnz = 2048;
    __global__ void __launch_bounds__(256)
    k_sub_stream(const int32_t *__restrict__ rowblk, const int32_t *__restrict__ rows, const int32_t *__restrict__ rp,
                 const int32_t *__restrict__ col, const SubCsr::value_t *__restrict__ val, const double *x, const double *a,
                 const double *__restrict__ m, const double *b, double *out)
    {
      __shared__ double prod[kSubNnz];
      const int r0 = rowblk[blockIdx.x], r1 = rowblk[blockIdx.x + 1];
      const int k0 = rp[r0], n = rp[r1] - k0;
      constexpr int PER = kSubNnz / 256;
      int32_t       cc[PER];
      double        vv[PER];
#pragma unroll
      for (int u = 0; u < PER; ++u)
        {
          const int i = threadIdx.x + u * 256;
          cc[u]       = i < n ? __ldg(col + k0 + i) : -1;
        }
#pragma unroll
      for (int u = 0; u < PER; ++u)
        {
          const int i = threadIdx.x + u * 256;
          vv[u]       = i < n ? (double)__ldcs(val + k0 + i) : 0.;
        }
#pragma unroll
      for (int u = 0; u < PER; ++u)
        if (cc[u] >= 0)
          prod[threadIdx.x + u * 256] = vv[u] * x[cc[u]];
      __syncthreads();
      for (int r = r0 + threadIdx.x; r < r1; r += 256)
        {
          const int bb = rp[r] - k0, ee = rp[r + 1] - k0;
          double    s  = 0.;
          for (int k = bb; k < ee; ++k)
            s += prod[k];
          const int row = rows[r];
          double    v   = b ? b[row] - s : s;
          if (m)
            v *= m[row];
          if (a)
            v += a[row];
          out[row] = v;
        }
    }

    // t = [bubble: pinv r, other u: 0]; z = pinv r on Dirichlet-masked / all u dofs as default
    __global__ void
    k_pre_u(const int64_t n_u, const uint8_t *__restrict__ cls, const double *__restrict__ pinv,
            const double *__restrict__ r, double *__restrict__ t, double *__restrict__ z)
    {
      const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
      if (i >= n_u)
        return;
      const double j = pinv[i] * r[i];
      z[i]           = j; // overwritten for free vertex dofs and for bubbles later
      t[i]           = cls[i] == CLS_BUBBLE ? j : 0.;
    }
    // b_u[v] = r[d] - t2[d]
    __global__ void
    k_gather_u2(const int nv, const int V, const CellBox bx, const int64_t L, const int32_t *__restrict__ vdof,
                const uint8_t *__restrict__ mask, const double *__restrict__ r, const double *__restrict__ t2,
                double *__restrict__ bu)
    {
      const int bw = bx.i1 - bx.i0 + 1, t = blockIdx.x * blockDim.x + threadIdx.x;
      if (t >= bw * (bx.j1 - bx.j0 + 1))
        return;
      const int v = (bx.j0 + t / bw) * V + bx.i0 + t % bw;
      const int  d       = vdof[v];
      const bool m       = (mask && mask[v]) || d < 0 || d >= L; // only the owner contributes (all-reduced)
      bu[v]              = m ? 0. : r[d] - (t2 ? t2[d] : 0.);
      bu[(size_t)nv + v] = m ? 0. : r[d + 1] - (t2 ? t2[d + 1] : 0.);
    }
    // z[d] = t[d] = x_v on free vertex dofs
    __global__ void
    k_scatter_u2(const int nv, const int V, const CellBox bx, const int32_t *__restrict__ vdof,
                 const uint8_t *__restrict__ mask, const double *__restrict__ xu, double *__restrict__ z,
                 double *__restrict__ t)
    {
      const int bw = bx.i1 - bx.i0 + 1, q = blockIdx.x * blockDim.x + threadIdx.x;
      if (q >= bw * (bx.j1 - bx.j0 + 1))
        return;
      const int v = (bx.j0 + q / bw) * V + bx.i0 + q % bw;
      if (mask && mask[v])
        return;
      const int d = vdof[v];
      if (d < 0)
        return; // vertex not visible from this rank
      const double a0 = xu[v], a1 = xu[(size_t)nv + v];
      z[d]     = a0;
      z[d + 1] = a1;
      t[d]     = a0;
      t[d + 1] = a1;
    }
    // pressure diagonals: dinv_i = 1/(|A_ii| + gamma area_i)
    __global__ void
    k_pint_dinv(const int n, const int32_t *__restrict__ rows, const double *__restrict__ diag,
                const double *__restrict__ area, const double gamma, double *__restrict__ dinv)
    {
      const int r = blockIdx.x * blockDim.x + threadIdx.x;
      if (r >= n)
        return;
      const int    row = rows[r];
      const double d   = fabs(diag[row]) + gamma * area[r];
      dinv[row]        = d != 0. ? 1. / d : 1.;
    }
    // dinvF[f] = 1/(A_ff - sum_i A_fi^2 dinv_i)
    __global__ void
    k_face_dinv(const int n, const int32_t *__restrict__ rows, const int32_t *__restrict__ rp,
                const int32_t *__restrict__ col, const uint32_t *__restrict__ src, const double *__restrict__ values,
                const double *__restrict__ diag, double *__restrict__ dinv)
    {
      const int r = blockIdx.x * blockDim.x + threadIdx.x;
      if (r >= n)
        return;
      const int row = rows[r];
      double    d   = fabs(diag[row]);
      for (int k = rp[r]; k < rp[r + 1]; ++k)
        {
          const double v = values[src[k]];
          d -= v * v * dinv[col[k]];
        }
      dinv[row] = d > 0. ? 1. / d : 1.;
    }
    // t_i = dinv_i r_i on interior pressures (others untouched)
    __global__ void
    k_pre_p(const int n, const int32_t *__restrict__ rows, const double *__restrict__ dinv,
            const double *__restrict__ r, double *__restrict__ t)
    {
      const int k = blockIdx.x * blockDim.x + threadIdx.x;
      if (k >= n)
        return;
      const int row = rows[k];
      t[row]        = dinv[row] * r[row];
    }
    inline unsigned
    gb(const int64_t n, const int b = 128)
    {
      return (unsigned)std::max<int64_t>(1, (n + b - 1) / b);
    }
  } // namespace

  BlockPrecond::~BlockPrecond()
  {
    if (ev_fork)
      cudaEventDestroy(ev_fork);
    if (ev_join)
      cudaEventDestroy(ev_join);
    if (s2)
      cudaStreamDestroy(s2);
  }

  // ---- once per mesh: dof classes and the sub-block index lists ------------------------------------
  // Local numbering: owned rows [0, L), halo columns [L, L+H).  pat holds the owned rows with GLOBAL
  // column ids; to_local maps them.
  int
  precond_build_structure(BlockPrecond &bp, const HostMesh &mesh, const HostDofs &d, const HostPattern &pat,
                          const Partition &part, const std::function<int32_t(uint32_t)> &to_local, cudaStream_t s)
  {
    bp.ready        = false;
    const int     n_loc = d.n_loc;
    const int64_t L = part.n_owned, LH = L + (int64_t)part.ghost_cols.size();
    if ((int64_t)pat.colidx.size() > 0xffffffffll)
      return 1;
    std::vector<uint8_t> cls((size_t)LH, CLS_VERTEX);
    std::vector<double>  area_of_pint;
    std::vector<int32_t> pint_rows;
    for (int64_t c = 0; c < mesh.n_cells; ++c)
      {
        const uint32_t *cd = &d.cell_dofs[(size_t)c * n_loc];
        for (int i = 8; i < n_loc; ++i)
          {
            const int32_t l = to_local(cd[i]);
            if (l < 0)
              continue;
            cls[l] = i == n_loc - 1 ? CLS_PINT : ((d.element == 0 && ((i - 8) % 2 == 0)) ? CLS_BUBBLE : CLS_PFACE);
          }
      }
    // owned p_int rows with their cell areas (shoelace on vertices 0,1,3,2)
    for (int64_t c = part.cell_begin; c < part.cell_end; ++c)
      {
        const uint32_t *cv = &mesh.cell_vertices[4 * (size_t)c];
        const double    x0 = mesh.x[cv[0]], y0 = mesh.y[cv[0]], x1 = mesh.x[cv[1]], y1 = mesh.y[cv[1]];
        const double    x2 = mesh.x[cv[2]], y2 = mesh.y[cv[2]], x3 = mesh.x[cv[3]], y3 = mesh.y[cv[3]];
        pint_rows.push_back(to_local(d.cell_dofs[(size_t)c * n_loc + n_loc - 1]));
        area_of_pint.push_back(0.5 * ((x3 - x0) * (y2 - y1) - (x2 - x1) * (y3 - y0)));
      }
    // physical index (pe_k3.h) of the xi unknown that belongs to interior-pressure column c (local numbering)
    const int64_t pint_begin = part.local_off[1], Lx = part.local_off[2] - part.local_off[1];
    const int64_t hu_end = L + (std::lower_bound(part.ghost_cols.begin(), part.ghost_cols.end(), (uint32_t)d.n_u) -
                                part.ghost_cols.begin());
    auto xi_of = [&](const int32_t c) -> int32_t {
      return (int32_t)(c < L ? LH + (c - pint_begin) : LH + Lx + (c - hu_end));
    };
    auto build = [&](const int row_mask, const int col_mask, SubCsr &out, const int neg_mask = 0,
                     const bool cols_to_xi = false) -> bool {
      std::vector<int32_t>  rows, rp(1, 0), col;
      std::vector<uint32_t> src;
      std::vector<uint8_t>  neg;
      for (int64_t r = 0; r < L; ++r)
        {
          if (!((row_mask >> cls[r]) & 1))
            continue;
          rows.push_back((int32_t)r);
          for (int64_t k = pat.rowptr[r]; k < pat.rowptr[r + 1]; ++k)
            {
              const int32_t cc = to_local((uint32_t)pat.colidx[k]);
              if (cc >= 0 && ((col_mask >> cls[cc]) & 1))
                {
                  col.push_back(cols_to_xi ? xi_of(cc) : cc);
                  src.push_back((uint32_t)k);
                  if (neg_mask)
                    neg.push_back((uint8_t)((neg_mask >> cls[cc]) & 1));
                }
            }
          rp.push_back((int32_t)col.size());
        }
      out.n_rows = (int)rows.size();
      auto up    = [&](auto &dev, const auto &host) -> bool {
        if (dev.alloc(host.size()) != cudaSuccess)
          return false;
        if (host.empty())
          return true;
        return cudaMemcpyAsync(dev.p, host.data(), host.size() * sizeof(host[0]), cudaMemcpyHostToDevice, s) ==
               cudaSuccess;
      };
      // row runs for the streamed product
      std::vector<int32_t> blk(1, 0);
      {
        int r = 0;
        const int nr = (int)rows.size();
        while (r < nr)
          {
            int e = r + 1;
            while (e < nr && rp[e + 1] - rp[r] <= kSubNnz && e - r < 256)
              ++e;
            blk.push_back(e);
            r = e;
          }
      }
      out.n_blk = (int)blk.size() - 1;
      out.nnz   = (int64_t)col.size();
      const bool ok = up(out.rows, rows) && up(out.rp, rp) && up(out.col, col) && up(out.src, src) && up(out.rowblk, blk) &&
                      (neg.empty() || up(out.neg, neg)) && out.val.alloc(col.size()) == cudaSuccess;
      cudaStreamSynchronize(s);
      return ok;
    };
    bp.has_bubbles = d.element == 0;
    if (bp.has_bubbles)
      if (!build(1 << CLS_BUBBLE, (1 << CLS_VERTEX) | (1 << CLS_BUBBLE), bp.Abu) ||
          !build(1 << CLS_VERTEX, 1 << CLS_BUBBLE, bp.Avb))
        return -1;
    if (!build(1 << CLS_PINT, 1 << CLS_PFACE, bp.Sif) || !build(1 << CLS_PFACE, 1 << CLS_PINT, bp.Sfi) ||
        !build(1 << CLS_PFACE, (1 << CLS_PINT) | (1 << CLS_PFACE), bp.Spf, 1 << CLS_PINT))
      return -1;
    // -B^T: displacement rows x interior-pressure columns of A0 (alpha = 1), columns renumbered to the xi unknowns
    // of the three-field vectors (block-triangular preconditioner of the GMRES solve)
    if (!build((1 << CLS_VERTEX) | (1 << CLS_BUBBLE), 1 << CLS_PINT, bp.Bt, 0, true))
      return -1;
    if (bp.cls.alloc(cls.size()) != cudaSuccess || bp.pint_rows.alloc(pint_rows.size()) != cudaSuccess ||
        bp.pint_area.alloc(area_of_pint.size()) != cudaSuccess)
      return -1;
    cudaMemcpyAsync(bp.cls.p, cls.data(), cls.size(), cudaMemcpyHostToDevice, s);
    cudaMemcpyAsync(bp.pint_rows.p, pint_rows.data(), pint_rows.size() * sizeof(int32_t), cudaMemcpyHostToDevice, s);
    cudaMemcpyAsync(bp.pint_area.p, area_of_pint.data(), area_of_pint.size() * sizeof(double), cudaMemcpyHostToDevice, s);
    if (bp.t.alloc((size_t)LH) != cudaSuccess || bp.t2.alloc((size_t)LH) != cudaSuccess || bp.dinv.alloc((size_t)LH) != cudaSuccess ||
        bp.t3.alloc((size_t)LH) != cudaSuccess || bp.wdinv.alloc((size_t)LH) != cudaSuccess || bp.t4.alloc((size_t)LH) != cudaSuccess || bp.t5.alloc((size_t)LH) != cudaSuccess)
      return -1;
    cudaMemsetAsync(bp.t4.p, 0, sizeof(double) * (size_t)LH, s);
    cudaMemsetAsync(bp.t3.p, 0, sizeof(double) * (size_t)LH, s);
    cudaMemsetAsync(bp.wdinv.p, 0, sizeof(double) * (size_t)LH, s);
    cudaMemsetAsync(bp.t.p, 0, sizeof(double) * (size_t)LH, s);
    cudaMemsetAsync(bp.t2.p, 0, sizeof(double) * (size_t)LH, s);
    cudaMemsetAsync(bp.dinv.p, 0, sizeof(double) * (size_t)LH, s);
    if (cudaStreamSynchronize(s) != cudaSuccess)
      return -1;
    {
      // bounding box of the assembled cells (lexicographic vertex ids: cell (ci, cj) has vertex 0 = cj * V + ci)
      const int V = (1 << mesh.n_refine) + 1;
      CellBox   bx{INT32_MAX, 0, INT32_MAX, 0};
      for (const int64_t c : part.asm_cells)
        {
          const uint32_t v0 = mesh.cell_vertices[4 * (size_t)c];
          const int      ci = (int)(v0 % V), cj = (int)(v0 / V);
          bx.i0 = std::min(bx.i0, ci);
          bx.i1 = std::max(bx.i1, ci + 1);
          bx.j0 = std::min(bx.j0, cj);
          bx.j1 = std::max(bx.j1, cj + 1);
        }
      if (part.asm_cells.empty())
        bx = CellBox{0, 0, 0, 0};
      bp.box = bx;
      if (part.world > 1)
        {
          const size_t nvv = (size_t)V * V, ncc = (size_t)(V - 1) * (V - 1);
          if (bp.bu_local.alloc(2 * nvv) != cudaSuccess || bp.rc_local.alloc(ncc) != cudaSuccess)
            return -1;
          cudaMemsetAsync(bp.bu_local.p, 0, sizeof(double) * 2 * nvv, s);
          cudaMemsetAsync(bp.rc_local.p, 0, sizeof(double) * ncc, s);
          cudaStreamSynchronize(s);
        }
    }
    if (!bp.s2)
      {
        cudaStreamCreateWithFlags(&bp.s2, cudaStreamNonBlocking);
        cudaEventCreateWithFlags(&bp.ev_fork, cudaEventDisableTiming);
        cudaEventCreateWithFlags(&bp.ev_join, cudaEventDisableTiming);
      }
    bp.n_u     = part.local_off[1]; // owned displacement rows come first
    bp.n       = L;
    bp.n_pint  = (int)pint_rows.size();
    bp.ready   = true;
    return 0;
  }

  // ---- once per solve: the pressure diagonals from the freshly assembled values ---------------------
  int
  precond_update(BlockPrecond &bp, const double *values, const double *diag, const double gamma, cudaStream_t s,
                 int64_t *launches)
  {
    bp.values = values;
    for (SubCsr *sc : {&bp.Abu, &bp.Avb, &bp.Sif, &bp.Sfi, &bp.Spf, &bp.Bt})
      if (sc->nnz > 0)
        {
          k_sub_gather_values<<<gb(sc->nnz, 256), 256, 0, s>>>(sc->nnz, sc->src.p, sc->neg.p, values, sc->val.p);
          ++*launches;
        }
    k_pint_dinv<<<gb(bp.n_pint), 128, 0, s>>>(bp.n_pint, bp.pint_rows.p, diag, bp.pint_area.p, gamma, bp.dinv.p);
    if (bp.halo)
      bp.halo(bp.comm_ctx, bp.dinv.p); // D_i of the neighbours' interior pressures
    k_face_dinv<<<gb(bp.Sfi.n_rows), 128, 0, s>>>(bp.Sfi.n_rows, bp.Sfi.rows.p, bp.Sfi.rp.p, bp.Sfi.col.p, bp.Sfi.src.p,
                                                   values, diag, bp.dinv.p);
    k_face_wdinv<<<gb(bp.Sfi.n_rows), 128, 0, s>>>(bp.Sfi.n_rows, bp.Sfi.rows.p, bp.dinv.p, bp.omega_f, bp.wdinv.p);
    *launches += 3;
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
  }

  // ---- z = P^{-1} r ---------------------------------------------------------------------------------
  // Multi-GPU: sub-block products need the neighbours' entries (halo exchange of t and z); the two
  // V-cycles run replicated on every rank on the all-reduced grid right-hand sides.
  // z_xi: when non-null (physical three-field vector holding the already preconditioned xi part of z), the
  // displacement block is applied to  r_u + B^T z_xi  instead of r_u: block upper-triangular variant
  //   P = [ P_u  -B^T  0 ; 0  P_xi  0 ; 0  0  P_p ]   of the row-signed system (pe_solver.cu, FGMRES)
  void
  precond_apply(BlockPrecond &bp, MgHierarchy &mg, CellMg &cm, const double *pinv, const double *r, double *z,
                cudaStream_t s, int64_t *launches, const double *z_xi, const double *rp_in)
  {
    const double *rp = rp_in ? rp_in : r; // pressure part of the input (row-scaled GMRES operator: unscaled copy)
    MgLevel      &L0 = mg.lv[0];
    const int     N  = L0.N, nv = (N + 1) * (N + 1);
    const int64_t L  = bp.n;
    auto sub = [&](cudaStream_t st, const SubCsr &S, const double *x, const double *a, const double *m, const double *b,
                   double *out) {
      if (S.n_blk > 0)
        k_sub_stream<<<(unsigned)S.n_blk, 256, 0, st>>>(S.rowblk.p, S.rows.p, S.rp.p, S.col.p, S.val.p, x, a, m, b, out);
      ++*launches;
    };
    // block-triangular variant: ru = r_u + B^T z_xi  (the Bt block holds -B^T: ru = r - Bt z_xi), kept in t4
    const double *ru = r;
    if (z_xi && bp.Bt.n_blk > 0)
      {
        sub(s, bp.Bt, z_xi, nullptr, nullptr, r, bp.t4.p);
        ru = bp.t4.p;
      }
    // local Jacobi parts:  u: z = pinv r, t = y1 on bubbles;  p_int: t_i = dinv_i r_i
    k_pre_u<<<gb(bp.n_u, 256), 256, 0, s>>>(bp.n_u, bp.cls.p, pinv, ru, bp.t.p, z);
    k_pre_p<<<gb(bp.n_pint), 128, 0, s>>>(bp.n_pint, bp.pint_rows.p, bp.dinv.p, rp, bp.t.p);
    *launches += 2;
    if (bp.halo)
      bp.halo(bp.comm_ctx, bp.t.p);
    PE_MARK(s, "precond: local Jacobi + halo");
    // The displacement and the pressure part touch disjoint entries of t, t2, z: on one GPU the pressure
    // pipeline is forked onto a second stream (inside the captured CUDA graph: a parallel branch).
    // one GPU: always; several GPUs: when the communication layer has a second lane (stream + NCCL communicator)
    // (PE_TIMING serialises the two pipelines so that the phase marks, all on one stream then, add up)
    const bool   fork = bp.s2 != nullptr && (!bp.halo || bp.set_lane) && !PhaseTimer::get().on;
    cudaStream_t sp   = fork ? bp.s2 : s;
    auto         halo_p = [&](double *x) {
      if (bp.halo)
        (fork && bp.halo_p ? bp.halo_p : bp.halo)(bp.comm_ctx, x);
    };
    if (fork)
      {
        cudaEventRecord(bp.ev_fork, s);
        cudaStreamWaitEvent(sp, bp.ev_fork, 0);
      }
    // ---- displacement block: bubble sweep, vertex V-cycle on the updated residual, bubble sweep
    if (bp.has_bubbles)
      sub(s, bp.Avb, bp.t.p, nullptr, nullptr, nullptr, bp.t2.p);
    const int nbox_v = (bp.box.i1 - bp.box.i0 + 1) * (bp.box.j1 - bp.box.j0 + 1);
    k_gather_u2<<<gb(nbox_v), 128, 0, s>>>(nv, N + 1, bp.box, L, mg.vdof.p, L0.mask_u.p, ru, bp.has_bubbles ? bp.t2.p : nullptr,
                                           bp.allreduce2 ? bp.bu_local.p : L0.bu.p);
    ++*launches;
    PE_MARK(s, "precond: A_vb + vertex gather");
    if (bp.allreduce2)
      {
        // a distributed finest level only needs its own rows: reduce to the slab owners (half the traffic)
        if (mg.n_dist > 0 && mg.gc.boxes_to_slabs)
          mg.gc.boxes_to_slabs(mg.gc.ctx, bp.bu_local.p, L0.bu.p, N + 1, 2, nv, mg.bounds[0].data(), 0);
        else if (mg.n_dist > 0 && mg.gc.reduce_rows)
          mg.gc.reduce_rows(mg.gc.ctx, bp.bu_local.p, L0.bu.p, N + 1, 2, nv, mg.bounds[0].data());
        else
          bp.allreduce2(bp.comm_ctx, bp.bu_local.p, L0.bu.p, 2 * nv);
      }
    PE_MARK(s, "precond: all-reduce vertex rhs");
    mg_vcycle(mg, 2, s, launches);
    PE_MARK(s, "precond: vertex V-cycle");
    k_scatter_u2<<<gb(nbox_v), 128, 0, s>>>(nv, N + 1, bp.box, mg.vdof.p, L0.mask_u.p, L0.wu[2].p, z, bp.t.p);
    ++*launches;
    if (bp.has_bubbles && fork)
      sub(s, bp.Abu, bp.t.p, bp.t.p, pinv, ru, z); // z_b = y1 + pinv_b (r_b - A_bb y1 - A_bv x_v),  t = (x_v, y1)
    // ---- pressure block: interior pressures eliminated exactly; on the face Schur complement
    //      F = S_ff - S_fi D_i^{-1} S_if:  Jacobi, cell-space correction (pe_cmg.h), Jacobi
    if (fork && bp.set_lane)
      bp.set_lane(bp.comm_ctx, 1); // communication of the pressure pipeline: lane 1
    // g = r_f - S_fi D_i^{-1} r_i   -> t2 (face entries)
    sub(sp, bp.Sfi, bp.t.p, nullptr, nullptr, rp, bp.t2.p);
    // y1 = omega dinvF g            -> t (face entries)
    k_face_scale<<<gb(bp.Sfi.n_rows), 128, 0, sp>>>(bp.Sfi.n_rows, bp.Sfi.rows.p, bp.wdinv.p, bp.t2.p, bp.t.p);
    ++*launches;
    // F y = S_ff y - S_fi q,  q = D_i^{-1} S_if y  (q on the p_int entries of t, y on its face entries)
    auto apply_q = [&]() {
      halo_p(bp.t.p);
      sub(sp, bp.Sif, bp.t.p, nullptr, bp.dinv.p, nullptr, bp.t.p);
      halo_p(bp.t.p);
    };
    apply_q();
    // res = g - F y1                -> t3 (face entries)
    sub(sp, bp.Spf, bp.t.p, nullptr, nullptr, bp.t2.p, bp.t3.p);
    cmg_gather(cm, L, mg.hedge.p, mg.vedge.p, bp.t3.p, bp.box, bp.allreduce2 ? bp.rc_local.p : cm.lv[0].b.p, sp);
    ++*launches;
    PE_MARK(sp, "precond: face Schur products");
    if (bp.allreduce2)
      {
        if (cm.n_dist > 0 && cm.gc.boxes_to_slabs)
          cm.gc.boxes_to_slabs(cm.gc.ctx, bp.rc_local.p, cm.lv[0].b.p, N, 1, 0, cm.bounds[0].data(), 1);
        else if (cm.n_dist > 0 && cm.gc.reduce_rows)
          cm.gc.reduce_rows(cm.gc.ctx, bp.rc_local.p, cm.lv[0].b.p, N, 1, 0, cm.bounds[0].data());
        else
          bp.allreduce2(bp.comm_ctx, bp.rc_local.p, cm.lv[0].b.p, N * N);
      }
    PE_MARK(sp, "precond: all-reduce cell rhs");
    const double *e = cmg_vcycle(cm, sp, launches);
    PE_MARK(sp, "precond: cell V-cycle");
    // y2 = y1 + Pi_c e              -> t (face entries)
    cmg_scatter(cm, L, mg.hedge.p, mg.vedge.p, e, bp.t.p, bp.t.p, bp.box, sp);
    ++*launches;
    apply_q();
    // z_f = y2 + omega dinvF (g - F y2)
    sub(sp, bp.Spf, bp.t.p, bp.t.p, bp.wdinv.p, bp.t2.p, z);
    if (bp.has_bubbles && !fork)
      sub(s, bp.Abu, bp.t.p, bp.t.p, pinv, ru, z);
    halo_p(z); // face pressures of the neighbours
    // z_i = dinv_i (r_i - S_if z_f)
    sub(sp, bp.Sif, z, nullptr, bp.dinv.p, rp, z);
    PE_MARK(sp, "precond: face Schur products");
    if (fork && bp.set_lane)
      bp.set_lane(bp.comm_ctx, 0);
    if (fork)
      {
        cudaEventRecord(bp.ev_join, sp);
        cudaStreamWaitEvent(s, bp.ev_join, 0);
      }
  }
} // namespace pe
