// TEST INFRASTRUCTURE ONLY -- never linked into, imported by or executed from the product path.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it.
//
// Literal CPU restatement of the reference's hot path (same loop nests, same order of operations):
//   BR1-WG   assemble_system : PoroElasBR1WG_new.cc:684-1174
//   CGQ1-WG  assemble_system : PoroElasCGQ1_WG.cc:1030-1729
//   error norms               : PoroElasBR1WG_new.cc:1251-1349, 2251-2266
//   postprocess (dilation, |u|, Darcy velocity): PoroElasCGQ1_WG.cc:1744-1966, 1973-2225;
//                                PoroElasBR1WG_new.cc:1189-1246, 1352-1606
// on top of the deal.II behaviour restated in dealii_restated.hpp.
//
// PARITY UNPINNED: the reference ships no golden vectors and cannot be built here (deal.II 9.1 +
// UMFPACK absent).  Pins are the hand-derived known answers of SURVEY.md 8(c) (tests/test_oracle_kat.py).
#pragma once
#include "dealii_restated.hpp"

namespace orc
{
  // ------------------------------------------------------------------------------------------
  // Problem data ("case"): the analytic functions of the reference, selected by id instead of by
  // commenting blocks.
  //   CASE_ZERO   : the shipped state of BR1new (cantilever): u=p=f=s=0        BR1new:220-223,384,424
  //   CASE_COSCOS : the commented "coscos" blocks                               BR1new:207-210,247-248,
  //                 278-283,376-378,409-412,453-456 ; CGQ1:259-262,582-584,635-638,707-710
  // BodyRightHandSide / FluidRightHandSide hard-code THEIR OWN lambda, mu, alpha, capacity
  // (BR1new:369-370,400-402) -- kept separate from the material constants (f_lambda, ...).
  // ------------------------------------------------------------------------------------------
  enum CaseId
  {
    CASE_ZERO   = 0,
    CASE_COSCOS = 1
  };
  struct Case
  {
    int    id       = CASE_ZERO;
    double f_lambda = 1., f_mu = 1., f_alpha = 1., s_capacity = 0., s_alpha = 1.;
  };

  inline void
  exact_u(const Case &cs, const double t, const double x, const double y, double u[2])
  {
    if (cs.id == CASE_COSCOS)
      {
        u[0] = 1. / (2. * M_PI) * std::sin(M_PI * x) * std::cos(M_PI * y) * std::sin(M_PI / 2 * t);
        u[1] = 1. / (2. * M_PI) * std::cos(M_PI * x) * std::sin(M_PI * y) * std::sin(M_PI / 2 * t);
      }
    else
      u[0] = u[1] = 0;
  }
  inline double
  exact_p(const Case &cs, const double t, const double x, const double y)
  {
    if (cs.id == CASE_COSCOS)
      return std::sin(M_PI / 2. * t) * std::cos(M_PI * x) * std::cos(M_PI * y);
    return 0;
  }
  // DisplacementGradient::vector_value (BR1new:278-283)
  inline void
  exact_grad_u(const Case &cs, const double t, const double x, const double y, double g[2][2])
  {
    g[0][0] = g[0][1] = g[1][0] = g[1][1] = 0;
    if (cs.id == CASE_COSCOS)
      {
        g[0][0] = 1. / (2 * M_PI) * std::sin(M_PI / 2 * t) * M_PI * std::cos(M_PI * x) * std::cos(M_PI * y);
        g[0][1] = -1. / (2 * M_PI) * std::sin(M_PI / 2 * t) * M_PI * std::sin(M_PI * x) * std::sin(M_PI * y);
        g[1][0] = g[0][1];
        g[1][1] = g[0][0];
      }
  }
  // BodyRightHandSide::value (BR1new:409-412)
  inline void
  body_rhs(const Case &cs, const double t, const double x, const double y, double f[2])
  {
    if (cs.id == CASE_COSCOS)
      {
        f[0] = std::sin(M_PI / 2. * t) * std::sin(M_PI * x) * std::cos(M_PI * y) * M_PI *
               (cs.f_lambda + 2 * cs.f_mu - cs.f_alpha);
        f[1] = std::sin(M_PI / 2. * t) * std::cos(M_PI * x) * std::sin(M_PI * y) * M_PI *
               (cs.f_lambda + 2 * cs.f_mu - cs.f_alpha);
      }
    else
      f[0] = f[1] = 0;
  }
  // FluidRightHandSide::value (BR1new:376-378)
  inline double
  fluid_rhs(const Case &cs, const double t, const double x, const double y)
  {
    if (cs.id == CASE_COSCOS)
      return std::cos(M_PI * x) * std::cos(M_PI * y) *
             (M_PI / 2 * std::cos(M_PI / 2. * t) * (cs.s_capacity + cs.s_alpha) +
              2 * M_PI * M_PI * std::sin(M_PI / 2. * t));
    return 0;
  }

  // face "kind" bits: what the reference does on a boundary face (selected there by boundary ids
  // and by which blocks are commented in)
  enum FaceKind
  {
    FK_DIR_U = 1, // displacement Dirichlet            BR1new:1120 (boundary_id()==0) / CGQ1:1674-1679
    FK_DIR_P = 2, // face-pressure Dirichlet           BR1new:1063-1070 (general case) / CGQ1:1663-1670
    FK_NEU_U = 4  // traction t_N on this face          BR1new:964-990 / CGQ1:1564-1584
  };

  struct Problem
  {
    Mesh                      mesh;
    DoFs                      dofs;
    Pattern                   pattern;
    std::vector<std::uint8_t> face_kind; // 4 per cell (0 on interior faces)
    std::vector<double>       kcell;     // scalar permeability per cell (K = k I)
    double                    lambda = 1, mu = 1, alpha = 1, capacity = 0, time_step = 0.1;
    double                    traction[2] = {0, 0};
    Case                      cs;
  };

  // ------------------------------------------------------------------------------------------
  // FEValues / FEFaceValues of the system element and of FE_RaviartThomas(0) on one cell, at an
  // arbitrary list of reference points (BR1new:693-709,794-795,844-845).
  // ------------------------------------------------------------------------------------------
  struct ShapeTables
  {
    int                 nq = 0, n_loc = 0;
    std::vector<double> JxW, px, py;
    std::vector<double> uval;  // [q][i][c]
    std::vector<double> ugrad; // [q][i][c][d] = d(phi_i)_c / dx_d
    std::vector<double> pint;  // [q][i]   FE_DGQ(0)
    std::vector<double> pface; // [q][i]   FE_FaceQ(0): 0 in the cell interior, 1 on the own face
    std::vector<double> rtval; // [q][k][c]
    std::vector<double> rtdiv; // [q][k]
    std::vector<double> nrm;   // [q][2]   (faces only)
    const double *
    u(int q, int i) const
    {
      return &uval[((size_t)q * n_loc + i) * 2];
    }
    const double *
    g(int q, int i) const
    {
      return &ugrad[((size_t)q * n_loc + i) * 4];
    }
    double
    div(int q, int i) const
    {
      const double *G = g(q, i);
      return G[0] + G[3];
    }
    const double *
    rt(int q, int k) const
    {
      return &rtval[((size_t)q * 4 + k) * 2];
    }
  };

  inline void
  fill_point(ShapeTables &T, const int q, const std::vector<LocalDof> &ld, const MapPoint &m,
             const double xi, const double eta, const int on_face /* -1: cell interior */)
  {
    const int n_loc = (int)ld.size();
    double    S[8], dS[8][2];
    scalar_shapes(xi, eta, S, dS);
    T.px[q] = m.x[0];
    T.py[q] = m.x[1];
    for (int i = 0; i < n_loc; ++i)
      {
        double *uv = &T.uval[((size_t)q * n_loc + i) * 2];
        double *ug = &T.ugrad[((size_t)q * n_loc + i) * 4];
        uv[0] = uv[1] = 0;
        ug[0] = ug[1] = ug[2] = ug[3] = 0;
        T.pint[(size_t)q * n_loc + i]  = 0;
        T.pface[(size_t)q * n_loc + i] = 0;
        const LocalDof &d              = ld[i];
        if (d.kind == U_VERTEX || d.kind == U_BUBBLE)
          {
            const int s = (d.kind == U_VERTEX) ? d.obj : 4 + d.obj;
            const int c = d.comp;
            uv[c]       = S[s];
            for (int dd = 0; dd < 2; ++dd)
              ug[c * 2 + dd] = dS[s][0] * m.Jinv[0][dd] + dS[s][1] * m.Jinv[1][dd];
          }
        else if (d.kind == P_INT)
          T.pint[(size_t)q * n_loc + i] = 1.;
        else if (d.kind == P_FACE && on_face == d.obj)
          T.pface[(size_t)q * n_loc + i] = 1.;
      }
    double w[4][2], dv[4];
    rt0_ref(xi, eta, w, dv);
    for (int k = 0; k < 4; ++k)
      {
        for (int c = 0; c < 2; ++c)
          T.rtval[((size_t)q * 4 + k) * 2 + c] = (m.J[c][0] * w[k][0] + m.J[c][1] * w[k][1]) / m.det;
        T.rtdiv[(size_t)q * 4 + k] = dv[k] / m.det;
      }
  }

  inline void
  alloc_tables(ShapeTables &T, const int nq, const int n_loc)
  {
    T.nq    = nq;
    T.n_loc = n_loc;
    T.JxW.assign(nq, 0);
    T.px.assign(nq, 0);
    T.py.assign(nq, 0);
    T.uval.assign((size_t)nq * n_loc * 2, 0);
    T.ugrad.assign((size_t)nq * n_loc * 4, 0);
    T.pint.assign((size_t)nq * n_loc, 0);
    T.pface.assign((size_t)nq * n_loc, 0);
    T.rtval.assign((size_t)nq * 8, 0);
    T.rtdiv.assign((size_t)nq * 4, 0);
    T.nrm.assign((size_t)nq * 2, 0);
  }

  inline ShapeTables
  fe_values(const CellGeom &c, const Quad2 &quad, const std::vector<LocalDof> &ld)
  {
    ShapeTables T;
    alloc_tables(T, (int)quad.w.size(), (int)ld.size());
    for (int q = 0; q < T.nq; ++q)
      {
        const MapPoint m = map_at(c, quad.p[q].x, quad.p[q].y);
        T.JxW[q]         = m.det * quad.w[q];
        fill_point(T, q, ld, m, quad.p[q].x, quad.p[q].y, -1);
      }
    return T;
  }

  inline ShapeTables
  fe_face_values(const CellGeom &c, const int f, const std::vector<double> &x1, const std::vector<double> &w1,
                 const std::vector<LocalDof> &ld)
  {
    ShapeTables T;
    alloc_tables(T, (int)w1.size(), (int)ld.size());
    for (int q = 0; q < T.nq; ++q)
      {
        const FacePoint fp = face_point(c, f, x1[q], w1[q]);
        const Point2    r  = project_to_face(f, x1[q]);
        T.JxW[q]           = fp.JxW;
        T.nrm[2 * q]       = fp.n[0];
        T.nrm[2 * q + 1]   = fp.n[1];
        fill_point(T, q, ld, fp.m, r.x, r.y, f);
      }
    return T;
  }

  // local index of the displacement / face-pressure dofs living on face f, in deal.II's face-dof
  // order (vertex 0 dofs, vertex 1 dofs, then line dofs)  -- what BR1new:1080-1107 recovers by
  // comparing global indices.
  inline void
  face_u_local(const Element e, const int f, int out[5], int &n)
  {
    n = 0;
    for (int v = 0; v < 2; ++v)
      for (int c = 0; c < 2; ++c)
        out[n++] = 2 * face_vertices[f][v] + c;
    if (e == BR1_WG)
      out[n++] = 8 + 2 * f;
  }
  inline int
  pface_local(const Element e, const int f)
  {
    return e == BR1_WG ? 9 + 2 * f : 8 + f;
  }

  // ------------------------------------------------------------------------------------------
  // One cell of assemble_system: local matrix, local rhs and the per-cell constraint object.
  // ------------------------------------------------------------------------------------------
  struct CellResult
  {
    std::vector<double> A, b;
    AffineConstraints   con;
  };

  inline void
  assemble_cell(const Problem &P, const int c, const double t, const std::vector<double> &old_solution,
                CellResult &R)
  {
    const Element               el    = P.dofs.element;
    const std::vector<LocalDof> ld    = local_dofs(el);
    const int                   n     = (int)ld.size(); // dofs_per_cell
    const int                   n_rt  = 4;              // dofs_per_cell_rt
    const dof_t                *idx   = &P.dofs.cell_dofs[(size_t)c * n];
    const CellGeom              geom  = P.mesh.geom(c);
    const double                dt    = P.time_step;
    const double                Kc    = P.kcell[c];
    static const Quad2          quad3 = qgauss_2d(3); // QGauss<dim>(3)       BR1new:690 CGQ1:1037
    static const Quad2          quad1 = qgauss_2d(1); // QGauss<dim>(1)       CGQ1:1040
    static std::vector<double>  fx, fw;
    if (fx.empty())
      qgauss_1d(3, fx, fw); // QGauss<dim-1>(3)  BR1new:691
    const ShapeTables fv = fe_values(geom, quad3, ld);
    const int         nq = fv.nq;

    std::vector<double> M(n_rt * n_rt, 0.), F((size_t)n * n_rt, 0.), C((size_t)n * n_rt, 0.);
    R.A.assign((size_t)n * n, 0.);
    R.b.assign(n, 0.);
    std::vector<double> &A = R.A, &b = R.b;

    // old solution at q-points (fe_values.get_function_values)             BR1new:806
    std::vector<double> cell_old(n);
    for (int i = 0; i < n; ++i)
      cell_old[i] = old_solution[idx[i]];

    // RT Gram matrix and inverse                                            BR1new:811-826
    for (int q = 0; q < nq; ++q)
      for (int i = 0; i < n_rt; ++i)
        for (int j = 0; j < n_rt; ++j)
          M[i * n_rt + j] += (fv.rt(q, i)[0] * fv.rt(q, j)[0] + fv.rt(q, i)[1] * fv.rt(q, j)[1]) * fv.JxW[q];
    gauss_jordan(M, n_rt);

    // F: interior part                                                      BR1new:828-840
    for (int q = 0; q < nq; ++q)
      for (int i = 0; i < n; ++i)
        for (int k = 0; k < n_rt; ++k)
          F[i * n_rt + k] -= fv.pint[(size_t)q * n + i] * fv.rtdiv[(size_t)q * 4 + k] * fv.JxW[q];
    // F: face part                                                          BR1new:842-860
    for (int f = 0; f < 4; ++f)
      {
        const ShapeTables ff = fe_face_values(geom, f, fx, fw, ld);
        for (int q = 0; q < ff.nq; ++q)
          for (int i = 0; i < n; ++i)
            for (int k = 0; k < n_rt; ++k)
              F[i * n_rt + k] += ff.pface[(size_t)q * n + i] *
                                 (ff.rt(q, k)[0] * ff.nrm[2 * q] + ff.rt(q, k)[1] * ff.nrm[2 * q + 1]) * ff.JxW[q];
      }
    // C = F * M^{-1}  (mmult)                                               BR1new:863
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n_rt; ++j)
        {
          double s = 0;
          for (int k = 0; k < n_rt; ++k)
            s += F[i * n_rt + k] * M[k * n_rt + j];
          C[i * n_rt + j] = s;
        }

    // average divergence                                                    BR1new:867-876
    const double        cell_area = cell_measure(geom);
    std::vector<double> avg_div(n, 0.);
    for (int i = 0; i < n; ++i)
      for (int q = 0; q < nq; ++q)
        avg_div[i] += fv.div(q, i) * fv.JxW[q] / cell_area;

    // dt (K grad_w p, grad_w q): the quintuple loop                         BR1new:879-901
    for (int q = 0; q < nq; ++q)
      for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j)
          for (int k = 0; k < n_rt; ++k)
            for (int l = 0; l < n_rt; ++l)
              A[i * n + j] += (dt * Kc * C[i * n_rt + k] * C[j * n_rt + l] *
                               (fv.rt(q, k)[0] * fv.rt(q, l)[0] + fv.rt(q, k)[1] * fv.rt(q, l)[1])) *
                              fv.JxW[q];

    auto sym_grad_dot = [&](const int q, const int i, const int j) {
      const double *gi = fv.g(q, i), *gj = fv.g(q, j);
      const double  si[4] = {gi[0], .5 * (gi[1] + gi[2]), .5 * (gi[1] + gi[2]), gi[3]};
      const double  sj[4] = {gj[0], .5 * (gj[1] + gj[2]), .5 * (gj[1] + gj[2]), gj[3]};
      return si[0] * sj[0] + si[1] * sj[1] + si[2] * sj[2] + si[3] * sj[3];
    };

    double div_old_red = 0; // CGQ1 only
    if (el == BR1_WG)
      {
        // elasticity + coupling + storage                                   BR1new:904-925
        for (int q = 0; q < nq; ++q)
          for (int i = 0; i < n; ++i)
            for (int j = 0; j < n; ++j)
              A[i * n + j] += (2. * P.mu * sym_grad_dot(q, i, j) + P.lambda * avg_div[i] * avg_div[j] -
                               P.alpha * fv.pint[(size_t)q * n + j] * avg_div[i] +
                               P.capacity * (fv.pint[(size_t)q * n + i] * fv.pint[(size_t)q * n + j]) +
                               P.alpha * (avg_div[j] * fv.pint[(size_t)q * n + i])) *
                              fv.JxW[q];
      }
    else
      {
        // 2 mu eps:eps + c0 mass, full rule                                 CGQ1:1325-1339
        for (int q = 0; q < nq; ++q)
          for (int i = 0; i < n; ++i)
            for (int j = 0; j < n; ++j)
              A[i * n + j] += (2. * P.mu * sym_grad_dot(q, i, j) +
                               P.capacity * (fv.pint[(size_t)q * n + i] * fv.pint[(size_t)q * n + j])) *
                              fv.JxW[q];
        // lambda div div, coupling, c0 mass AGAIN at the 1-point rule        CGQ1:1341-1358
        const ShapeTables fr = fe_values(geom, quad1, ld);
        for (int q = 0; q < fr.nq; ++q)
          for (int i = 0; i < n; ++i)
            for (int j = 0; j < n; ++j)
              A[i * n + j] += (P.lambda * fr.div(q, i) * fr.div(q, j) -
                               P.alpha * fr.pint[(size_t)q * n + j] * fr.div(q, i) +
                               P.capacity * (fr.pint[(size_t)q * n + i] * fr.pint[(size_t)q * n + j]) +
                               P.alpha * (fr.div(q, j) * fr.pint[(size_t)q * n + i])) *
                              fr.JxW[q];
        // div(u_old) at the reduced rule                                    CGQ1:1390-1392
        for (int i = 0; i < n; ++i)
          div_old_red += cell_old[i] * fr.div(0, i);
        // rhs term with reduced integration                                 CGQ1:1445-1453
        for (int q = 0; q < fr.nq; ++q)
          for (int i = 0; i < n; ++i)
            b[i] += (P.alpha * div_old_red * fr.pint[(size_t)q * n + i]) * fr.JxW[q];
      }

    // divergence of old displacement and its cell average                   BR1new:928-940
    std::vector<double> div_old(nq, 0.), old_pint(nq, 0.);
    for (int q = 0; q < nq; ++q)
      for (int i = 0; i < n; ++i)
        {
          div_old[q] += cell_old[i] * fv.div(q, i);
          old_pint[q] += cell_old[i] * fv.pint[(size_t)q * n + i]; // BR1new:943-945
        }
    double avg_div_old = 0;
    for (int q = 0; q < nq; ++q)
      avg_div_old += div_old[q] * fv.JxW[q] / cell_area;

    // rhs                                                                   BR1new:947-961 / CGQ1:1430-1443
    for (int q = 0; q < nq; ++q)
      {
        const double s = fluid_rhs(P.cs, t, fv.px[q], fv.py[q]);
        double       f[2];
        body_rhs(P.cs, t, fv.px[q], fv.py[q], f);
        for (int i = 0; i < n; ++i)
          {
            const double *phi = fv.u(q, i);
            const double  piq = fv.pint[(size_t)q * n + i];
            if (el == BR1_WG)
              b[i] += (f[0] * phi[0] + f[1] * phi[1] + P.capacity * old_pint[q] * piq + dt * s * piq +
                       P.alpha * avg_div_old * piq) *
                      fv.JxW[q];
            else
              b[i] += (f[0] * phi[0] + f[1] * phi[1] + P.capacity * old_pint[q] * piq + dt * s * piq) * fv.JxW[q];
          }
      }

    // Neumann (traction)                                                    BR1new:963-1015 / CGQ1:1563-1584
    for (int f = 0; f < 4; ++f)
      if (P.face_kind[4 * c + f] & FK_NEU_U)
        {
          const ShapeTables ff = fe_face_values(geom, f, fx, fw, ld);
          for (int q = 0; q < ff.nq; ++q)
            for (int i = 0; i < n; ++i)
              b[i] += (P.traction[0] * ff.u(q, i)[0] + P.traction[1] * ff.u(q, i)[1]) * ff.JxW[q];
        }

    // Dirichlet constraints, rebuilt per cell                               BR1new:1060-1165 / CGQ1:1660-1720
    R.con.clear();
    for (int f = 0; f < 4; ++f)
      {
        const std::uint8_t kind = P.face_kind[4 * c + f];
        if (kind & FK_DIR_P)
          {
            // interpolate_boundary_values with the face-pressure mask: FE_FaceQ(0) support point
            // = face midpoint [deal.II-memory]
            const Point2   r = project_to_face(f, .5);
            const MapPoint m = map_at(geom, r.x, r.y);
            const dof_t    g = idx[pface_local(el, f)];
            R.con.add_line(g);
            R.con.set_inhomogeneity(g, exact_p(P.cs, t, m.x[0], m.x[1]));
          }
        if (kind & FK_DIR_U)
          {
            int fu[5], nfu;
            face_u_local(el, f, fu, nfu);
            double vertex_values[4];
            for (int dof = 0; dof < 4; ++dof) // BR1new:1113-1127
              {
                const int v = face_vertices[f][dof / 2];
                double    u[2];
                exact_u(P.cs, t, geom.vx[v], geom.vy[v], u);
                vertex_values[dof] = u[dof % 2];
                R.con.add_line(idx[fu[dof]]);
                R.con.set_inhomogeneity(idx[fu[dof]], vertex_values[dof]);
              }
            if (el == BR1_WG)
              {
                // bubble dof: match the normal flux of (u - I_h u)          BR1new:1129-1157
                const ShapeTables ff       = fe_face_values(geom, f, fx, fw, ld);
                double            res_flux = 0, bubble_flux = 0;
                const double      nx = ff.nrm[0], ny = ff.nrm[1]; // normal_vector(0)
                for (int q = 0; q < ff.nq; ++q)
                  {
                    double res[2];
                    exact_u(P.cs, t, ff.px[q], ff.py[q], res);
                    for (int d = 0; d < 2; ++d)
                      for (int dof = 0; dof < 4; ++dof)
                        res[d] -= vertex_values[dof] * ff.u(q, fu[dof])[d];
                    const double *phi = ff.u(q, fu[4]);
                    bubble_flux += (phi[0] * nx + phi[1] * ny) * ff.JxW[q];
                    res_flux += (res[0] * nx + res[1] * ny) * ff.JxW[q];
                  }
                R.con.add_line(idx[fu[4]]);
                R.con.set_inhomogeneity(idx[fu[4]], res_flux / bubble_flux);
              }
          }
      }
  }

  // CGQ1 calls interpolate_boundary_values over the WHOLE mesh inside the cell loop
  // (CGQ1:1662-1679), so a cell sees constraints on all of its boundary dofs, also those it owns
  // only through a vertex.  (On the unit square this equals the per-face rule above; the global
  // map is kept for fidelity.)  Built once per assemble instead of once per cell: same result.
  inline AffineConstraints
  global_boundary_constraints(const Problem &P, const double t)
  {
    AffineConstraints con;
    const int         n = P.dofs.n_loc;
    for (int c = 0; c < P.mesh.n_cells(); ++c)
      {
        const CellGeom geom = P.mesh.geom(c);
        const dof_t   *idx  = &P.dofs.cell_dofs[(size_t)c * n];
        for (int f = 0; f < 4; ++f)
          {
            const std::uint8_t kind = P.face_kind[4 * c + f];
            if (kind & FK_DIR_P)
              {
                const Point2   r = project_to_face(f, .5);
                const MapPoint m = map_at(geom, r.x, r.y);
                con.set_inhomogeneity(idx[pface_local(P.dofs.element, f)], exact_p(P.cs, t, m.x[0], m.x[1]));
              }
            if (kind & FK_DIR_U)
              for (int dof = 0; dof < 4; ++dof)
                {
                  const int v = face_vertices[f][dof / 2];
                  double    u[2];
                  exact_u(P.cs, t, geom.vx[v], geom.vy[v], u);
                  con.set_inhomogeneity(idx[2 * v + dof % 2], u[dof % 2]);
                }
          }
      }
    return con;
  }

  // assemble_system(t): system_matrix = 0; system_rhs = 0; cell loop; distribute_local_to_global
  inline void
  assemble_system(const Problem &P, const double t, const std::vector<double> &old_solution,
                  std::vector<double> &values, std::vector<double> &rhs)
  {
    CsrMatrix mat;
    mat.p = &P.pattern;
    mat.val.assign(P.pattern.colidx.size(), 0.);
    rhs.assign(P.dofs.n(), 0.);
    CellResult        R;
    AffineConstraints global_con;
    if (P.dofs.element == CGQ1_WG)
      global_con = global_boundary_constraints(P, t);
    for (int c = 0; c < P.mesh.n_cells(); ++c)
      {
        assemble_cell(P, c, t, old_solution, R);
        const AffineConstraints &con = (P.dofs.element == CGQ1_WG) ? global_con : R.con;
        distribute_local_to_global(con, R.A, R.b, &P.dofs.cell_dofs[(size_t)c * P.dofs.n_loc], P.dofs.n_loc, mat,
                                   rhs);
      }
    values.swap(mat.val);
  }

  // ------------------------------------------------------------------------------------------
  // postprocess_errors (BR1new:1251-1349): per-step squared errors
  //   e[0] = ||p_int - p||^2_L2, e[1] = ||u_h - u||^2_L2, e[2] = |u_h - u|^2_H1   (QGauss(3))
  // ------------------------------------------------------------------------------------------
  inline void
  step_errors(const Problem &P, const double t, const std::vector<double> &solution, double e[3])
  {
    const std::vector<LocalDof> ld    = local_dofs(P.dofs.element);
    const int                   n     = (int)ld.size();
    static const Quad2          quad3 = qgauss_2d(3);
    e[0] = e[1] = e[2] = 0;
    for (int c = 0; c < P.mesh.n_cells(); ++c)
      {
        const CellGeom    geom = P.mesh.geom(c);
        const ShapeTables fv   = fe_values(geom, quad3, ld);
        const dof_t      *idx  = &P.dofs.cell_dofs[(size_t)c * n];
        for (int q = 0; q < fv.nq; ++q)
          {
            double uh[2] = {0, 0}, gh[4] = {0, 0, 0, 0}, ph = 0;
            for (int i = 0; i < n; ++i)
              {
                const double s = solution[idx[i]];
                uh[0] += s * fv.u(q, i)[0];
                uh[1] += s * fv.u(q, i)[1];
                for (int k = 0; k < 4; ++k)
                  gh[k] += s * fv.g(q, i)[k];
                ph += s * fv.pint[(size_t)q * n + i];
              }
            double ue[2], ge[2][2];
            exact_u(P.cs, t, fv.px[q], fv.py[q], ue);
            exact_grad_u(P.cs, t, fv.px[q], fv.py[q], ge);
            const double pe = exact_p(P.cs, t, fv.px[q], fv.py[q]);
            e[0] += (ph - pe) * (ph - pe) * fv.JxW[q];
            e[1] += ((uh[0] - ue[0]) * (uh[0] - ue[0]) + (uh[1] - ue[1]) * (uh[1] - ue[1])) * fv.JxW[q];
            e[2] += ((gh[0] - ge[0][0]) * (gh[0] - ge[0][0]) + (gh[1] - ge[0][1]) * (gh[1] - ge[0][1]) +
                     (gh[2] - ge[1][0]) * (gh[2] - ge[1][0]) + (gh[3] - ge[1][1]) * (gh[3] - ge[1][1])) *
                    fv.JxW[q];
          }
      }
  }
  // ------------------------------------------------------------------------------------------
  // postprocess() of CGQ1 (PoroElasCGQ1_WG.cc:1744-1966) = postprocess() + postprocess_darcy_velocity() of BR1new
  // (PoroElasBR1WG_new.cc:1189-1246, 1352-1606), per cell:
  //   div_displacement[c]       = (1/|E|) sum_q sum_i u_i div(phi_i) JxW                      CGQ1:1849-1860
  //   displacement_magnitude[c] = sqrt( (1/|E|) sum_q |u_h(x_q)|^2 JxW )                       CGQ1:1862-1877
  //   darcy beta[c][k]          = - sum_j sum_i sol_i C_ij D_kj ,  D = M^-1 E,  E = int K w.w ,  C = F M^-1
  //                                                                                            CGQ1:1881-1964
  //   velocity_center[c]        = sum_k beta_k w_k(cell midpoint)   (QMidpoint)                BR1new:1580-1588
  // ------------------------------------------------------------------------------------------
  inline void
  postprocess(const Problem &P, const std::vector<double> &solution, std::vector<double> &div_displacement,
              std::vector<double> &displacement_magnitude, std::vector<double> &beta, std::vector<double> &velocity_center)
  {
    const std::vector<LocalDof> ld    = local_dofs(P.dofs.element);
    const int                   n     = (int)ld.size();
    const int                   n_rt  = 4;
    static const Quad2          quad3 = qgauss_2d(3);
    static const Quad2          quadm = qgauss_2d(1); // QMidpoint<dim> = the one-point Gauss rule
    static std::vector<double>  fx, fw;
    if (fx.empty())
      qgauss_1d(3, fx, fw);
    const int nc = P.mesh.n_cells();
    div_displacement.assign(nc, 0.);
    displacement_magnitude.assign(nc, 0.);
    beta.assign((size_t)nc * n_rt, 0.);
    velocity_center.assign((size_t)nc * 2, 0.);
    for (int c = 0; c < nc; ++c)
      {
        const CellGeom    geom = P.mesh.geom(c);
        const ShapeTables fv   = fe_values(geom, quad3, ld);
        const dof_t      *idx  = &P.dofs.cell_dofs[(size_t)c * n];
        const double      Kc   = P.kcell[c];
        double            cell_area = 0; // cell->measure()
        for (int q = 0; q < fv.nq; ++q)
          cell_area += fv.JxW[q];
        double d = 0;
        for (int q = 0; q < fv.nq; ++q)
          for (int i = 0; i < n; ++i)
            d += solution[idx[i]] * fv.div(q, i) * fv.JxW[q] / cell_area;
        div_displacement[c] = d;
        double m2 = 0;
        for (int q = 0; q < fv.nq; ++q)
          {
            double u[2] = {0, 0};
            for (int i = 0; i < n; ++i)
              {
                u[0] += solution[idx[i]] * fv.u(q, i)[0];
                u[1] += solution[idx[i]] * fv.u(q, i)[1];
              }
            m2 += (u[0] * u[0] + u[1] * u[1]) * fv.JxW[q] / cell_area;
          }
        displacement_magnitude[c] = std::sqrt(m2);
        std::vector<double> M(n_rt * n_rt, 0.), E(n_rt * n_rt, 0.), D(n_rt * n_rt, 0.), F((size_t)n * n_rt, 0.),
          C((size_t)n * n_rt, 0.);
        for (int q = 0; q < fv.nq; ++q)
          for (int i = 0; i < n_rt; ++i)
            for (int k = 0; k < n_rt; ++k)
              {
                const double vv = fv.rt(q, i)[0] * fv.rt(q, k)[0] + fv.rt(q, i)[1] * fv.rt(q, k)[1];
                E[i * n_rt + k] += Kc * vv * fv.JxW[q];
                M[i * n_rt + k] += vv * fv.JxW[q];
              }
        gauss_jordan(M, n_rt);
        for (int i = 0; i < n_rt; ++i) // D = M^-1 E (mmult)
          for (int j = 0; j < n_rt; ++j)
            {
              double s = 0;
              for (int k = 0; k < n_rt; ++k)
                s += M[i * n_rt + k] * E[k * n_rt + j];
              D[i * n_rt + j] = s;
            }
        for (int q = 0; q < fv.nq; ++q)
          for (int i = 0; i < n; ++i)
            for (int k = 0; k < n_rt; ++k)
              F[i * n_rt + k] -= fv.pint[(size_t)q * n + i] * fv.rtdiv[(size_t)q * 4 + k] * fv.JxW[q];
        for (int f = 0; f < 4; ++f)
          {
            const ShapeTables ff = fe_face_values(geom, f, fx, fw, ld);
            for (int q = 0; q < ff.nq; ++q)
              for (int i = 0; i < n; ++i)
                for (int k = 0; k < n_rt; ++k)
                  F[i * n_rt + k] += ff.pface[(size_t)q * n + i] *
                                     (ff.rt(q, k)[0] * ff.nrm[2 * q] + ff.rt(q, k)[1] * ff.nrm[2 * q + 1]) * ff.JxW[q];
          }
        for (int i = 0; i < n; ++i) // C = F M^-1
          for (int j = 0; j < n_rt; ++j)
            {
              double s = 0;
              for (int k = 0; k < n_rt; ++k)
                s += F[i * n_rt + k] * M[k * n_rt + j];
              C[i * n_rt + j] = s;
            }
        for (int k = 0; k < n_rt; ++k)
          for (int j = 0; j < n_rt; ++j)
            for (int i = 0; i < n; ++i)
              beta[(size_t)c * n_rt + k] += -(solution[idx[i]] * C[i * n_rt + j] * D[k * n_rt + j]);
        const ShapeTables fm = fe_values(geom, quadm, ld);
        for (int q = 0; q < fm.nq; ++q)
          for (int k = 0; k < n_rt; ++k)
            {
              velocity_center[(size_t)c * 2] += beta[(size_t)c * n_rt + k] * fm.rt(q, k)[0];
              velocity_center[(size_t)c * 2 + 1] += beta[(size_t)c * n_rt + k] * fm.rt(q, k)[1];
            }
      }
  }
} // namespace orc
