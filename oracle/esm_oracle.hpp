// TEST INFRASTRUCTURE ONLY -- never linked into, imported by or executed from the product path.
//
// Literal CPU restatement of EltStiffMat.cc (WGDarcyEquation<2>): WG(Q2,Q2;RT[2]) Darcy problem,
//   setup_system      EltStiffMat.cc:294-326   (distribute_dofs, make_sparsity_pattern; NO renumbering)
//   assemble_system   EltStiffMat.cc:333-578   (RT Gram + gauss_jordan 413-429, F 437-477, C = F M^-1 478,
//                                               q,i,j,k,l loop 483-502, rhs 505-513, scatter 551-552,
//                                               interpolate_boundary_values + apply_boundary_values 565-576)
//   solve             EltStiffMat.cc:583-590   (SolverCG + PreconditionIdentity: done by the tests with numpy)
// on top of restatements of the deal.II pieces it leans on, each tagged [deal.II-memory].
//
// PARITY UNPINNED (see dealii_restated.hpp): no golden vectors exist for this program.  What pins this
// file are size-independent facts checked in tests/test_esm_oracle.py: constants have zero weak
// gradient (row sums of the local matrix vanish), symmetry and positive semi-definiteness, patch test
// with quadratic pressures (the WG(Q2,Q2;RT[2]) scheme reproduces them exactly), convergence order.
#pragma once
#include "dealii_restated.hpp"

namespace orc
{
  namespace esm
  {
    constexpr int kNRt = 24, kNLoc = 21, kNFaceDofs = 12;

    // shifted Legendre polynomials on [0,1] and derivatives, degree <= 3
    inline void
    legendre01(const double x, double P[4], double dP[4])
    {
      const double s = 2. * x - 1.;
      P[0]  = 1.;
      P[1]  = s;
      P[2]  = .5 * (3. * s * s - 1.);
      P[3]  = .5 * (5. * s * s * s - 3. * s);
      dP[0] = 0.;
      dP[1] = 2.;
      dP[2] = 2. * 3. * s;
      dP[3] = 2. * .5 * (15. * s * s - 3.);
    }
    // Reference basis of RT_[2] = Q_{3,2} e_x + Q_{2,3} e_y (FE_RaviartThomas(2), 24 dofs).
    // [deal.II-memory] deal.II uses the nodal basis of face / interior moments; the weak-Galerkin
    // local matrix C M_K C^T only depends on the SPACE (it is the K-weighted L2 norm of the discrete weak
    // gradient, an L2-type projection onto the space), so any basis of the same space gives the same
    // matrix up to rounding.  Legendre products keep the Gram matrix well conditioned.
    //   k = 0..11 : (P_a(xi) P_b(eta), 0), a = k % 4, b = k / 4
    //   k = 12..23: (0, P_a(xi) P_b(eta)), a = (k-12) % 3, b = (k-12) / 3
    inline void
    rt2_ref(const double xi, const double eta, double w[kNRt][2], double divw[kNRt])
    {
      double Px[4], dPx[4], Py[4], dPy[4];
      legendre01(xi, Px, dPx);
      legendre01(eta, Py, dPy);
      for (int k = 0; k < 12; ++k)
        {
          const int a = k % 4, b = k / 4;
          w[k][0]     = Px[a] * Py[b];
          w[k][1]     = 0.;
          divw[k]     = dPx[a] * Py[b];
        }
      for (int k = 0; k < 12; ++k)
        {
          const int a = k % 3, b = k / 3;
          w[12 + k][0] = 0.;
          w[12 + k][1] = Px[a] * Py[b];
          divw[12 + k] = Px[a] * dPy[b];
        }
    }
    // 1-D Lagrange basis at the support points 0, 1/2, 1 [deal.II-memory: FE_DGQ(2) and FE_FaceQ(2) use
    // the Lagrange polynomials of QGaussLobatto<1>(3) = {0, .5, 1}, ascending, lexicographic tensor product]
    inline void
    lagrange3(const double x, double L[3])
    {
      L[0] = 2. * (x - .5) * (x - 1.);
      L[1] = -4. * x * (x - 1.);
      L[2] = 2. * x * (x - .5);
    }
    // FESystem(FE_DGQ(2), FE_FaceQ(2)) local layout [deal.II-memory]: line dofs first (lines 0..3, three
    // FE_FaceQ dofs each, along the line direction = +y for faces 0,1 and +x for faces 2,3), then the nine
    // FE_DGQ dofs of the quad (x fastest):   0..2 face 0 | 3..5 face 1 | 6..8 face 2 | 9..11 face 3 | 12..20 interior
    inline void
    interior_values(const double xi, const double eta, double v[kNLoc])
    {
      double Lx[3], Ly[3];
      lagrange3(xi, Lx);
      lagrange3(eta, Ly);
      for (int i = 0; i < kNFaceDofs; ++i)
        v[i] = 0.; // FE_FaceQ: zero inside the cell
      for (int j = 0; j < 3; ++j)
        for (int i = 0; i < 3; ++i)
          v[12 + 3 * j + i] = Lx[i] * Ly[j];
    }
    inline void
    face_values(const int f, const double t, double v[kNLoc])
    {
      double L[3];
      lagrange3(t, L);
      for (int i = 0; i < kNLoc; ++i)
        v[i] = 0.;
      for (int k = 0; k < 3; ++k)
        v[3 * f + k] = L[k];
    }

    struct Problem
    {
      Mesh                 mesh;
      int                  n_loc = kNLoc;
      std::vector<dof_t>   cell_dofs;
      dof_t                n_dofs = 0;
      Pattern              pattern;
      std::vector<double>  kcell;
    };

    // DoFHandler::distribute_dofs (ESM:297-298): first touch, per cell lines 0..3 then the quad; no renumbering
    inline void
    distribute_dofs(Problem &P)
    {
      const Mesh &m = P.mesh;
      const int   N = m.N, V = N + 1, nc = m.n_cells();
      const dof_t inval = 0xffffffffu;
      auto line_id = [&](std::uint32_t ix, std::uint32_t iy, int f) -> size_t {
        switch (f)
          {
            case 0:
              return (size_t)iy * V + ix;
            case 1:
              return (size_t)iy * V + ix + 1;
            case 2:
              return (size_t)V * N + (size_t)iy * N + ix;
            default:
              return (size_t)V * N + (size_t)(iy + 1) * N + ix;
          }
      };
      std::vector<dof_t> ldof((size_t)2 * V * N, inval);
      P.cell_dofs.resize((size_t)nc * kNLoc);
      dof_t next = 0;
      for (int c = 0; c < nc; ++c)
        {
          std::uint32_t ix, iy;
          morton_decode(c, ix, iy);
          for (int f = 0; f < 4; ++f)
            {
              dof_t &d = ldof[line_id(ix, iy, f)];
              if (d == inval)
                {
                  d = next;
                  next += 3;
                }
              for (int k = 0; k < 3; ++k)
                P.cell_dofs[(size_t)c * kNLoc + 3 * f + k] = d + k;
            }
          for (int k = 0; k < 9; ++k)
            P.cell_dofs[(size_t)c * kNLoc + 12 + k] = next + k;
          next += 9;
        }
      P.n_dofs = next;
    }

    inline Problem
    make_problem(const int n_refine, const double distort, const std::uint64_t seed)
    {
      Problem P;
      P.mesh = make_unit_square(n_refine, distort, seed);
      distribute_dofs(P);
      DoFs d;
      d.n_loc     = kNLoc;
      d.cell_dofs = P.cell_dofs;
      d.n_u       = P.n_dofs; // all dofs in one block (DoFs::n() = n_u + n_pint + n_pface)
      P.pattern   = make_sparsity_pattern(d, P.mesh.n_cells());
      P.kcell.assign(P.mesh.n_cells(), 1.);
      return P;
    }

    inline double
    rhs_value(const double x, const double y) // RightHandSide::value ESM:167
    {
      return 2. * M_PI * M_PI * std::cos(M_PI * x) * std::cos(M_PI * y);
    }
    inline double
    boundary_value(const double x, const double y) // BoundaryValues::value ESM:147
    {
      return std::cos(M_PI * x) * std::cos(M_PI * y);
    }

    // the cell loop body, ESM:400-513
    inline void
    assemble_cell(const Problem &P, const int c, std::vector<double> &local_matrix, std::vector<double> &cell_rhs)
    {
      static const Quad2         q2 = qgauss_2d(4); // QGauss<dim>(fe_rt.degree+1), FE_RaviartThomas(2).degree = 3
      static std::vector<double> fx, fw;
      if (fx.empty())
        qgauss_1d(4, fx, fw);
      const CellGeom geom = P.mesh.geom(c);
      const int      nq   = (int)q2.p.size();
      const double   K    = P.kcell[c]; // Coefficient: K = k I
      std::vector<double> M(kNRt * kNRt, 0.), F(kNLoc * kNRt, 0.), C(kNLoc * kNRt, 0.);
      local_matrix.assign(kNLoc * kNLoc, 0.);
      cell_rhs.assign(kNLoc, 0.);
      // tables at the cell quadrature points
      std::vector<double> rt((size_t)nq * kNRt * 2), rtdiv((size_t)nq * kNRt), pv((size_t)nq * kNLoc), JxW(nq), px(nq), py(nq);
      for (int q = 0; q < nq; ++q)
        {
          const MapPoint m = map_at(geom, q2.p[q].x, q2.p[q].y);
          JxW[q]           = m.det * q2.w[q];
          px[q]            = m.x[0];
          py[q]            = m.x[1];
          double w[kNRt][2], dv[kNRt];
          rt2_ref(q2.p[q].x, q2.p[q].y, w, dv);
          for (int k = 0; k < kNRt; ++k)
            {
              // contravariant Piola map: w = J w_ref / det J ; div w = div_ref w_ref / det J
              rt[((size_t)q * kNRt + k) * 2 + 0] = (m.J[0][0] * w[k][0] + m.J[0][1] * w[k][1]) / m.det;
              rt[((size_t)q * kNRt + k) * 2 + 1] = (m.J[1][0] * w[k][0] + m.J[1][1] * w[k][1]) / m.det;
              rtdiv[(size_t)q * kNRt + k]        = dv[k] / m.det;
            }
          interior_values(q2.p[q].x, q2.p[q].y, &pv[(size_t)q * kNLoc]);
        }
      // Gram matrix of the RT basis and its inverse                                  ESM:413-429
      for (int q = 0; q < nq; ++q)
        for (int i = 0; i < kNRt; ++i)
          for (int j = 0; j < kNRt; ++j)
            M[i * kNRt + j] += (rt[((size_t)q * kNRt + i) * 2] * rt[((size_t)q * kNRt + j) * 2] +
                                rt[((size_t)q * kNRt + i) * 2 + 1] * rt[((size_t)q * kNRt + j) * 2 + 1]) *
                               JxW[q];
      gauss_jordan(M, kNRt);
      // F: interior part                                                             ESM:437-452
      for (int q = 0; q < nq; ++q)
        for (int i = 0; i < kNLoc; ++i)
          for (int k = 0; k < kNRt; ++k)
            F[i * kNRt + k] -= pv[(size_t)q * kNLoc + i] * rtdiv[(size_t)q * kNRt + k] * JxW[q];
      // F: face part                                                                 ESM:454-476
      for (int f = 0; f < 4; ++f)
        for (size_t q = 0; q < fx.size(); ++q)
          {
            const FacePoint fp = face_point(geom, f, fx[q], fw[q]);
            const Point2    r  = project_to_face(f, fx[q]);
            double          w[kNRt][2], dv[kNRt], fv[kNLoc];
            rt2_ref(r.x, r.y, w, dv);
            face_values(f, fx[q], fv);
            for (int i = 0; i < kNLoc; ++i)
              for (int k = 0; k < kNRt; ++k)
                {
                  const double wx = (fp.m.J[0][0] * w[k][0] + fp.m.J[0][1] * w[k][1]) / fp.m.det;
                  const double wy = (fp.m.J[1][0] * w[k][0] + fp.m.J[1][1] * w[k][1]) / fp.m.det;
                  F[i * kNRt + k] += fv[i] * (wx * fp.n[0] + wy * fp.n[1]) * fp.JxW;
                }
          }
      // C = F M^-1                                                                    ESM:478
      for (int i = 0; i < kNLoc; ++i)
        for (int k = 0; k < kNRt; ++k)
          {
            double s = 0;
            for (int l = 0; l < kNRt; ++l)
              s += F[i * kNRt + l] * M[l * kNRt + k];
            C[i * kNRt + k] = s;
          }
      // local matrix: the q,i,j,k,l loop                                              ESM:483-502
      for (int q = 0; q < nq; ++q)
        for (int i = 0; i < kNLoc; ++i)
          for (int j = 0; j < kNLoc; ++j)
            for (int k = 0; k < kNRt; ++k)
              {
                const double *wk = &rt[((size_t)q * kNRt + k) * 2];
                for (int l = 0; l < kNRt; ++l)
                  {
                    const double *wl = &rt[((size_t)q * kNRt + l) * 2];
                    local_matrix[i * kNLoc + j] +=
                      K * C[i * kNRt + k] * C[j * kNRt + l] * (wk[0] * wl[0] + wk[1] * wl[1]) * JxW[q];
                  }
              }
      // right-hand side                                                               ESM:505-513
      for (int q = 0; q < nq; ++q)
        for (int i = 0; i < kNLoc; ++i)
          cell_rhs[i] += pv[(size_t)q * kNLoc + i] * rhs_value(px[q], py[q]) * JxW[q];
    }

    // interpolate_boundary_values with the face-pressure mask (ESM:565-571): FE_FaceQ(2) face support
    // points t = 0, 1/2, 1 on every boundary line [deal.II-memory]
    inline std::map<dof_t, double>
    boundary_values(const Problem &P)
    {
      std::map<dof_t, double> bv;
      for (int c = 0; c < P.mesh.n_cells(); ++c)
        {
          const CellGeom g = P.mesh.geom(c);
          for (int f = 0; f < 4; ++f)
            if (P.mesh.face_bid[4 * c + f] == 0)
              for (int k = 0; k < 3; ++k)
                {
                  const Point2   r = project_to_face(f, .5 * k);
                  const MapPoint m = map_at(g, r.x, r.y);
                  bv[P.cell_dofs[(size_t)c * kNLoc + 3 * f + k]] = boundary_value(m.x[0], m.x[1]);
                }
        }
      return bv;
    }

    // MatrixTools::apply_boundary_values(boundary_values, matrix, solution, rhs, eliminate_columns = true)
    // (ESM:572-575) [deal.II-memory: matrix_tools.cc]: for every boundary dof in ascending order: the row is
    // zeroed except for its diagonal d (if d == 0 the first non-zero diagonal of the matrix is used),
    // rhs = d * value, solution = value; the column is eliminated symmetrically: rhs_r -= a_r,dof * value.
    inline void
    apply_boundary_values(const std::map<dof_t, double> &bv, const Pattern &pat, std::vector<double> &val,
                          std::vector<double> &solution, std::vector<double> &rhs)
    {
      auto entry = [&](const dof_t r, const dof_t col) -> std::int64_t {
        const std::int32_t *b = &pat.colidx[pat.rowptr[r]], *e = &pat.colidx[pat.rowptr[r + 1]];
        const std::int32_t *it = std::lower_bound(b, e, (std::int32_t)col);
        return (it != e && *it == (std::int32_t)col) ? (it - &pat.colidx[0]) : -1;
      };
      const dof_t n = (dof_t)rhs.size();
      double first_nonzero_diagonal_entry = 1.;
      for (dof_t i = 0; i < n; ++i)
        {
          const std::int64_t k = entry(i, i);
          if (k >= 0 && val[k] != 0.)
            {
              first_nonzero_diagonal_entry = val[k];
              break;
            }
        }
      for (const auto &kv : bv)
        {
          const dof_t        dof = kv.first;
          const std::int64_t kd  = entry(dof, dof);
          for (std::int64_t k = pat.rowptr[dof]; k < pat.rowptr[dof + 1]; ++k)
            if (k != kd)
              val[k] = 0.;
          double new_rhs;
          if (val[kd] != 0.)
            new_rhs = kv.second * val[kd];
          else
            {
              val[kd] = first_nonzero_diagonal_entry;
              new_rhs = kv.second * first_nonzero_diagonal_entry;
            }
          rhs[dof]             = new_rhs;
          const double diag    = val[kd];
          for (std::int64_t k = pat.rowptr[dof]; k < pat.rowptr[dof + 1]; ++k)
            {
              const dof_t row = (dof_t)pat.colidx[k]; // symmetric pattern: rows that couple to dof
              if (row == dof)
                continue;
              const std::int64_t kr = entry(row, dof);
              rhs[row] -= val[kr] / diag * new_rhs;
              val[kr] = 0.;
            }
          solution[dof] = kv.second;
        }
    }

    // assemble_system, ESM:333-578: values in the CSR order of `pattern`
    inline void
    assemble(const Problem &P, std::vector<double> &val, std::vector<double> &rhs, std::vector<double> &solution,
             const bool with_boundary_values)
    {
      CsrMatrix mat;
      mat.p = &P.pattern;
      mat.val.assign(P.pattern.colidx.size(), 0.);
      rhs.assign(P.n_dofs, 0.);
      solution.assign(P.n_dofs, 0.);
      std::vector<double> A, b;
      for (int c = 0; c < P.mesh.n_cells(); ++c)
        {
          assemble_cell(P, c, A, b);
          const dof_t *idx = &P.cell_dofs[(size_t)c * kNLoc];
          for (int i = 0; i < kNLoc; ++i)
            {
              for (int j = 0; j < kNLoc; ++j)
                mat.add(idx[i], idx[j], A[i * kNLoc + j]);
              rhs[idx[i]] += b[i];
            }
        }
      if (with_boundary_values)
        apply_boundary_values(boundary_values(P), P.pattern, mat.val, solution, rhs);
      val = mat.val;
    }
    // Velocity::value  ESM:200-211: exact Darcy velocity -K grad p of p = cos(pi x) cos(pi y), K = I
    inline void
    exact_velocity(const double x, const double y, double v[2])
    {
      v[0] = M_PI * std::sin(M_PI * x) * std::cos(M_PI * y);
      v[1] = M_PI * std::cos(M_PI * x) * std::sin(M_PI * y);
    }

    // postprocess()  ESM:662-919: e[0] = L2_error_vel^2 = sum_cells int |u_h - u|^2,  u_h = sum_k beta_k w_k,
    //   beta_k = - sum_j sum_i sol_i C_ij D_kj,  D = M^-1 E,  E = int K w.w   (ESM:730-818)
    // e[1] = L2_error_flux^2 = sum_cells sum_faces (|E| / |f|) int_f ((u_h - u).n)^2   (ESM:840-905)
    inline void
    postprocess(const Problem &P, const std::vector<double> &solution, double e[2])
    {
      static const Quad2         q2 = qgauss_2d(4);
      static std::vector<double> fx, fw;
      if (fx.empty())
        qgauss_1d(4, fx, fw);
      const int nq = (int)q2.p.size();
      e[0] = e[1] = 0.;
      for (int c = 0; c < P.mesh.n_cells(); ++c)
        {
          const CellGeom geom = P.mesh.geom(c);
          const double   K    = P.kcell[c];
          const dof_t   *idx  = &P.cell_dofs[(size_t)c * kNLoc];
          std::vector<double> M(kNRt * kNRt, 0.), E(kNRt * kNRt, 0.), D(kNRt * kNRt, 0.), F(kNLoc * kNRt, 0.),
            C(kNLoc * kNRt, 0.), beta(kNRt, 0.);
          std::vector<double> rt((size_t)nq * kNRt * 2), rtdiv((size_t)nq * kNRt), pv((size_t)nq * kNLoc), JxW(nq), px(nq), py(nq);
          double cell_area = 0.;
          for (int q = 0; q < nq; ++q)
            {
              const MapPoint m = map_at(geom, q2.p[q].x, q2.p[q].y);
              JxW[q]           = m.det * q2.w[q];
              cell_area += JxW[q];
              px[q] = m.x[0];
              py[q] = m.x[1];
              double w[kNRt][2], dv[kNRt];
              rt2_ref(q2.p[q].x, q2.p[q].y, w, dv);
              for (int k = 0; k < kNRt; ++k)
                {
                  rt[((size_t)q * kNRt + k) * 2 + 0] = (m.J[0][0] * w[k][0] + m.J[0][1] * w[k][1]) / m.det;
                  rt[((size_t)q * kNRt + k) * 2 + 1] = (m.J[1][0] * w[k][0] + m.J[1][1] * w[k][1]) / m.det;
                  rtdiv[(size_t)q * kNRt + k]        = dv[k] / m.det;
                }
              interior_values(q2.p[q].x, q2.p[q].y, &pv[(size_t)q * kNLoc]);
            }
          for (int q = 0; q < nq; ++q)
            for (int i = 0; i < kNRt; ++i)
              for (int j = 0; j < kNRt; ++j)
                {
                  const double ww = rt[((size_t)q * kNRt + i) * 2] * rt[((size_t)q * kNRt + j) * 2] +
                                    rt[((size_t)q * kNRt + i) * 2 + 1] * rt[((size_t)q * kNRt + j) * 2 + 1];
                  E[i * kNRt + j] += K * ww * JxW[q];
                  M[i * kNRt + j] += ww * JxW[q];
                }
          gauss_jordan(M, kNRt);
          for (int i = 0; i < kNRt; ++i)
            for (int j = 0; j < kNRt; ++j)
              {
                double s = 0;
                for (int k = 0; k < kNRt; ++k)
                  s += M[i * kNRt + k] * E[k * kNRt + j];
                D[i * kNRt + j] = s;
              }
          for (int q = 0; q < nq; ++q)
            for (int i = 0; i < kNLoc; ++i)
              for (int k = 0; k < kNRt; ++k)
                F[i * kNRt + k] -= pv[(size_t)q * kNLoc + i] * rtdiv[(size_t)q * kNRt + k] * JxW[q];
          for (int f = 0; f < 4; ++f)
            for (size_t q = 0; q < fx.size(); ++q)
              {
                const FacePoint fp = face_point(geom, f, fx[q], fw[q]);
                const Point2    r  = project_to_face(f, fx[q]);
                double          w[kNRt][2], dv[kNRt], fv[kNLoc];
                rt2_ref(r.x, r.y, w, dv);
                face_values(f, fx[q], fv);
                for (int i = 0; i < kNLoc; ++i)
                  for (int k = 0; k < kNRt; ++k)
                    {
                      const double wx = (fp.m.J[0][0] * w[k][0] + fp.m.J[0][1] * w[k][1]) / fp.m.det;
                      const double wy = (fp.m.J[1][0] * w[k][0] + fp.m.J[1][1] * w[k][1]) / fp.m.det;
                      F[i * kNRt + k] += fv[i] * (wx * fp.n[0] + wy * fp.n[1]) * fp.JxW;
                    }
              }
          for (int i = 0; i < kNLoc; ++i)
            for (int k = 0; k < kNRt; ++k)
              {
                double s = 0;
                for (int l = 0; l < kNRt; ++l)
                  s += F[i * kNRt + l] * M[l * kNRt + k];
                C[i * kNRt + k] = s;
              }
          for (int k = 0; k < kNRt; ++k)
            for (int j = 0; j < kNRt; ++j)
              for (int i = 0; i < kNLoc; ++i)
                beta[k] += -(solution[idx[i]] * C[i * kNRt + j] * D[k * kNRt + j]);
          // velocity error on the cell                                                  ESM:822-838
          for (int q = 0; q < nq; ++q)
            {
              double vh[2] = {0, 0}, ve[2];
              for (int k = 0; k < kNRt; ++k)
                {
                  vh[0] += beta[k] * rt[((size_t)q * kNRt + k) * 2];
                  vh[1] += beta[k] * rt[((size_t)q * kNRt + k) * 2 + 1];
                }
              exact_velocity(px[q], py[q], ve);
              e[0] += ((vh[0] - ve[0]) * (vh[0] - ve[0]) + (vh[1] - ve[1]) * (vh[1] - ve[1])) * JxW[q];
            }
          // normal flux error on the faces                                              ESM:840-905
          for (int f = 0; f < 4; ++f)
            {
              const int    v0 = face_vertices[f][0], v1 = face_vertices[f][1];
              const double face_length = std::hypot(geom.vx[v1] - geom.vx[v0], geom.vy[v1] - geom.vy[v0]);
              double       local = 0.;
              for (size_t q = 0; q < fx.size(); ++q)
                {
                  const FacePoint fp = face_point(geom, f, fx[q], fw[q]);
                  const Point2    r  = project_to_face(f, fx[q]);
                  double          w[kNRt][2], dv[kNRt], vh[2] = {0, 0}, ve[2];
                  rt2_ref(r.x, r.y, w, dv);
                  for (int k = 0; k < kNRt; ++k)
                    {
                      vh[0] += beta[k] * (fp.m.J[0][0] * w[k][0] + fp.m.J[0][1] * w[k][1]) / fp.m.det;
                      vh[1] += beta[k] * (fp.m.J[1][0] * w[k][0] + fp.m.J[1][1] * w[k][1]) / fp.m.det;
                    }
                  exact_velocity(fp.m.x[0], fp.m.x[1], ve);
                  const double dn = (vh[0] - ve[0]) * fp.n[0] + (vh[1] - ve[1]) * fp.n[1];
                  local += dn * dn * fp.JxW;
                }
              e[1] += local / face_length * cell_area;
            }
        }
    }
  } // namespace esm
} // namespace orc
