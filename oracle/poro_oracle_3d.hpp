// TEST INFRASTRUCTURE ONLY -- never linked into, imported by or executed from the product path.
//
// Literal CPU restatement of the CELL LOOP of PoroElasCGQ1_WG.cc in THREE dimensions (the shipped main() is
// PoroElas_CGQ1_WG<3>, PoroElasCGQ1_WG.cc:3029; sandwich data 176-189, 859-866): per hexahedron the 31 x 31 local
// matrix and the local right-hand side of assemble_system (PoroElasCGQ1_WG.cc:1165-1453), with the same loop nests:
//   RT0 Gram matrix + gauss_jordan 1182-1209, F 1211-1245, C = F M^-1 1246, cell-averaged divergence 1247-1257,
//   Darcy q,i,j,k,l loop 1262-1284, 2 mu eps:eps + c0 mass at QGauss<3>(3) 1325-1339, lambda div div / -+alpha
//   coupling / c0 mass AGAIN at QGauss<3>(1) 1341-1358, right-hand side 1381-1453.
// on restatements of the deal.II pieces it leans on, each tagged [deal.II-memory]:
//   hyper_cube(0,1) + refine_global(n): active cells in 3-D Morton order (child c at (c&1, (c>>1)&1, c>>2));
//   hexahedron vertices lexicographic (v = x + 2 y + 4 z), faces 0:x=0 1:x=1 2:y=0 3:y=1 4:z=0 5:z=1;
//   FESystem(FE_Q(1)^3, FE_DGQ(0), FE_FaceQ(0)) local order: vertex dofs (3 components per vertex, vertices
//   0..7) = 0..23, quad dofs (FE_FaceQ(0) of faces 0..5) = 24..29, hex dof (FE_DGQ(0)) = 30;
//   FE_RaviartThomas(0): one function per face, contravariant Piola map (the local matrix depends on the SPACE only);
//   QGauss<3>(3) = 27 points x fastest, QGauss<2>(3) on faces, trilinear mapping.
// PARITY UNPINNED: no golden vectors exist; pinned by size-independent facts in tests/test_oracle_3d.py (symmetry
// of the symmetrised blocks, rigid-body null space of the elastic block, zero row sums of the Darcy block, the
// closed-form Darcy block of a cube, consistency with the 2-D oracle on an extruded cell).
#pragma once
#include "dealii_restated.hpp"
#include "poro_oracle.hpp" // FK_* face kinds

#include <array>
#include <map>
#include <set>

namespace orc
{
  namespace d3
  {
    constexpr int kNLoc = 31, kNRt = 6, kNU = 24, kPFace0 = 24, kPInt = 30;

    struct Mesh3
    {
      int                        N = 0;
      std::vector<double>        x, y, z;
      std::vector<std::uint32_t> cell_vertices; // 8 per cell, deal.II order
      int
      n_cells() const
      {
        return (int)(cell_vertices.size() / 8);
      }
    };
    inline void
    morton3_decode(std::uint32_t c, std::uint32_t &ix, std::uint32_t &iy, std::uint32_t &iz)
    {
      ix = iy = iz = 0;
      for (int b = 0; b < 10; ++b)
        {
          ix |= ((c >> (3 * b)) & 1u) << b;
          iy |= ((c >> (3 * b + 1)) & 1u) << b;
          iz |= ((c >> (3 * b + 2)) & 1u) << b;
        }
    }
    // GridGenerator::hyper_cube(0,1) + refine_global(n) (CGQ1:886-887); interior vertices displaced by
    // <= distort * h per coordinate with the counter generator (id = vertex id, k = coordinate)
    inline Mesh3
    make_unit_cube(const int n_refine, const double distort, const std::uint64_t seed)
    {
      Mesh3 m;
      m.N         = 1 << n_refine;
      const int N = m.N, V = N + 1;
      const double h = 1. / N;
      m.x.resize((size_t)V * V * V);
      m.y.resize(m.x.size());
      m.z.resize(m.x.size());
      for (int k = 0; k < V; ++k)
        for (int j = 0; j < V; ++j)
          for (int i = 0; i < V; ++i)
            {
              const size_t id = ((size_t)k * V + j) * V + i;
              double       x = i * h, y = j * h, z = k * h;
              if (distort > 0. && i > 0 && i < N && j > 0 && j < N && k > 0 && k < N)
                {
                  x += distort * h * (2. * u01(seed, id, 0) - 1.);
                  y += distort * h * (2. * u01(seed, id, 1) - 1.);
                  z += distort * h * (2. * u01(seed, id, 2) - 1.);
                }
              m.x[id] = x;
              m.y[id] = y;
              m.z[id] = z;
            }
      const size_t nc = (size_t)N * N * N;
      m.cell_vertices.resize(8 * nc);
      for (std::uint32_t c = 0; c < nc; ++c)
        {
          std::uint32_t ix, iy, iz;
          morton3_decode(c, ix, iy, iz);
          for (int v = 0; v < 8; ++v)
            m.cell_vertices[8 * (size_t)c + v] =
              (std::uint32_t)((((size_t)iz + (v >> 2)) * V + iy + ((v >> 1) & 1)) * V + ix + (v & 1));
        }
      return m;
    }

    struct Problem3
    {
      Mesh3               mesh;
      std::vector<double> kcell;
      double              lambda = 1, mu = 1, alpha = 1, capacity = 0, time_step = 1e-3;
      // boundary data of assemble_system: per cell and face a bit set FK_DIR_U | FK_DIR_P | FK_NEU_U (0 on interior
      // faces), one constant traction vector for the FK_NEU_U faces.  Dirichlet values are zero (BoundaryValues,
      // CGQ1:697-725, returns 0 in all components).
      std::vector<std::uint8_t> face_kind; // [n_cells][6]
      double                    traction[3] = {0, 0, 0};
    };
    // The shipped 3-D program (sandwich problem): boundary id 1 on faces with centre z = 1 (CGQ1:972-981) carries the
    // face-pressure Dirichlet condition (interpolate_boundary_values(.., 1, .., face_pressure_mask), CGQ1:1663-1670) and the
    // traction (0, 0, -1) (CGQ1:1563-1582); every other boundary face keeps id 0 = displacement Dirichlet (CGQ1:1673-1679).
    inline void
    set_sandwich_boundary(Problem3 &P)
    {
      const int nc = P.mesh.n_cells();
      P.face_kind.assign((size_t)6 * nc, 0);
      static const int face_v[6][4] = {{0, 2, 4, 6}, {1, 3, 5, 7}, {0, 1, 4, 5}, {2, 3, 6, 7}, {0, 1, 2, 3}, {4, 5, 6, 7}};
      for (int c = 0; c < nc; ++c)
        for (int f = 0; f < 6; ++f)
          {
            double ctr[3] = {0, 0, 0};
            for (int k = 0; k < 4; ++k)
              {
                const std::uint32_t id = P.mesh.cell_vertices[8 * (size_t)c + face_v[f][k]];
                ctr[0] += .25 * P.mesh.x[id];
                ctr[1] += .25 * P.mesh.y[id];
                ctr[2] += .25 * P.mesh.z[id];
              }
            const int  d  = f / 2;
            const bool at = std::fabs(ctr[d] - ((f & 1) ? 1. : 0.)) < 1e-12; // unit cube: at_boundary()
            if (!at)
              continue;
            P.face_kind[6 * (size_t)c + f] = std::fabs(ctr[2] - 1.) < 1e-12 ? (FK_DIR_P | FK_NEU_U) : FK_DIR_U;
          }
      P.traction[0] = P.traction[1] = 0;
      P.traction[2] = -1;
    }

    // trilinear map at a reference point: x, J = dx/dxi (rows x,y,z; columns xi,eta,zeta), det, J^-1
    struct Map3
    {
      double x[3], J[3][3], Jinv[3][3], det;
    };
    inline void
    shape_q1(const double r[3], double N[8], double dN[8][3])
    {
      for (int v = 0; v < 8; ++v)
        {
          const double f[3] = {(v & 1) ? r[0] : 1. - r[0], ((v >> 1) & 1) ? r[1] : 1. - r[1], (v >> 2) ? r[2] : 1. - r[2]};
          const double d[3] = {(v & 1) ? 1. : -1., ((v >> 1) & 1) ? 1. : -1., (v >> 2) ? 1. : -1.};
          N[v]     = f[0] * f[1] * f[2];
          dN[v][0] = d[0] * f[1] * f[2];
          dN[v][1] = f[0] * d[1] * f[2];
          dN[v][2] = f[0] * f[1] * d[2];
        }
    }
    inline Map3
    map_at(const double vx[8], const double vy[8], const double vz[8], const double r[3])
    {
      Map3   m;
      double N[8], dN[8][3];
      shape_q1(r, N, dN);
      for (int a = 0; a < 3; ++a)
        {
          m.x[a] = 0;
          for (int b = 0; b < 3; ++b)
            m.J[a][b] = 0;
        }
      for (int v = 0; v < 8; ++v)
        {
          const double p[3] = {vx[v], vy[v], vz[v]};
          for (int a = 0; a < 3; ++a)
            {
              m.x[a] += N[v] * p[a];
              for (int b = 0; b < 3; ++b)
                m.J[a][b] += p[a] * dN[v][b];
            }
        }
      const double(*J)[3] = m.J;
      m.det = J[0][0] * (J[1][1] * J[2][2] - J[1][2] * J[2][1]) - J[0][1] * (J[1][0] * J[2][2] - J[1][2] * J[2][0]) +
              J[0][2] * (J[1][0] * J[2][1] - J[1][1] * J[2][0]);
      for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b)
          {
            const int a1 = (a + 1) % 3, a2 = (a + 2) % 3, b1 = (b + 1) % 3, b2 = (b + 2) % 3;
            m.Jinv[b][a] = (J[a1][b1] * J[a2][b2] - J[a1][b2] * J[a2][b1]) / m.det;
          }
      return m;
    }
    // reference RT0 basis on the cube and its divergence: w_{2d} = (1 - x_d) e_d, w_{2d+1} = x_d e_d
    inline void
    rt0_ref3(const double r[3], double w[kNRt][3], double dv[kNRt])
    {
      for (int k = 0; k < kNRt; ++k)
        {
          w[k][0] = w[k][1] = w[k][2] = 0;
          const int d = k / 2;
          w[k][d]     = (k & 1) ? r[d] : 1. - r[d];
          dv[k]       = (k & 1) ? 1. : -1.;
        }
    }

    // FEValues of the system element + FE_RaviartThomas(0) at one mapped point
    struct Point3Tables
    {
      double JxW, x[3];
      double uval[kNLoc][3], ugrad[kNLoc][3][3]; // ugrad[i][c][d] = d(phi_i)_c / dx_d
      double pint[kNLoc], pface[kNLoc];
      double rt[kNRt][3], rtdiv[kNRt];
      double nrm[3];
    };
    inline void
    fill_point(Point3Tables &T, const Map3 &m, const double r[3], const int on_face)
    {
      double N[8], dN[8][3];
      shape_q1(r, N, dN);
      for (int i = 0; i < kNLoc; ++i)
        {
          for (int c = 0; c < 3; ++c)
            {
              T.uval[i][c] = 0;
              for (int d = 0; d < 3; ++d)
                T.ugrad[i][c][d] = 0;
            }
          T.pint[i] = T.pface[i] = 0;
        }
      for (int v = 0; v < 8; ++v)
        for (int c = 0; c < 3; ++c)
          {
            const int i  = 3 * v + c;
            T.uval[i][c] = N[v];
            for (int d = 0; d < 3; ++d)
              T.ugrad[i][c][d] = dN[v][0] * m.Jinv[0][d] + dN[v][1] * m.Jinv[1][d] + dN[v][2] * m.Jinv[2][d];
          }
      T.pint[kPInt] = 1.;
      if (on_face >= 0)
        T.pface[kPFace0 + on_face] = 1.; // FE_FaceQ(0): 1 on its own face, 0 inside the cell
      double w[kNRt][3], dv[kNRt];
      rt0_ref3(r, w, dv);
      for (int k = 0; k < kNRt; ++k)
        {
          for (int c = 0; c < 3; ++c)
            T.rt[k][c] = (m.J[c][0] * w[k][0] + m.J[c][1] * w[k][1] + m.J[c][2] * w[k][2]) / m.det;
          T.rtdiv[k] = dv[k] / m.det;
        }
      for (int a = 0; a < 3; ++a)
        T.x[a] = m.x[a];
    }

    inline void
    assemble_cell(const Problem3 &P, const int c, const std::vector<double> &cell_old /* 31 */, std::vector<double> &A,
                  std::vector<double> &b)
    {
      const int n = kNLoc, n_rt = kNRt;
      double    vx[8], vy[8], vz[8];
      for (int v = 0; v < 8; ++v)
        {
          const std::uint32_t id = P.mesh.cell_vertices[8 * (size_t)c + v];
          vx[v]                  = P.mesh.x[id];
          vy[v]                  = P.mesh.y[id];
          vz[v]                  = P.mesh.z[id];
        }
      static std::vector<double> gx, gw;
      if (gx.empty())
        qgauss_1d(3, gx, gw);
      const double dt = P.time_step, Kc = P.kcell[c];
      // QGauss<3>(3)
      std::vector<Point3Tables> fv(27);
      for (int qk = 0; qk < 3; ++qk)
        for (int qj = 0; qj < 3; ++qj)
          for (int qi = 0; qi < 3; ++qi)
            {
              const int    q    = (qk * 3 + qj) * 3 + qi;
              const double r[3] = {gx[qi], gx[qj], gx[qk]};
              const Map3   m    = map_at(vx, vy, vz, r);
              fv[q].JxW         = m.det * gw[qi] * gw[qj] * gw[qk];
              fill_point(fv[q], m, r, -1);
            }
      // QGauss<3>(1): the cell midpoint, weight 1
      Point3Tables fr;
      {
        const double r[3] = {.5, .5, .5};
        const Map3   m    = map_at(vx, vy, vz, r);
        fr.JxW            = m.det;
        fill_point(fr, m, r, -1);
      }
      std::vector<double> M(n_rt * n_rt, 0.), F((size_t)n * n_rt, 0.), C((size_t)n * n_rt, 0.);
      A.assign((size_t)n * n, 0.);
      b.assign(n, 0.);
      // RT Gram matrix and inverse                                                      CGQ1:1182-1209
      for (int q = 0; q < 27; ++q)
        for (int i = 0; i < n_rt; ++i)
          for (int j = 0; j < n_rt; ++j)
            M[i * n_rt + j] +=
              (fv[q].rt[i][0] * fv[q].rt[j][0] + fv[q].rt[i][1] * fv[q].rt[j][1] + fv[q].rt[i][2] * fv[q].rt[j][2]) * fv[q].JxW;
      gauss_jordan(M, n_rt);
      // F: interior part                                                                CGQ1:1211-1223
      for (int q = 0; q < 27; ++q)
        for (int i = 0; i < n; ++i)
          for (int k = 0; k < n_rt; ++k)
            F[i * n_rt + k] -= fv[q].pint[i] * fv[q].rtdiv[k] * fv[q].JxW;
      // F: face part, QGauss<2>(3) on every face                                        CGQ1:1224-1245
      for (int f = 0; f < 6; ++f)
        {
          const int d = f / 2, a1 = (d + 1) % 3, a2 = (d + 2) % 3;
          for (int q2 = 0; q2 < 3; ++q2)
            for (int q1 = 0; q1 < 3; ++q1)
              {
                double r[3];
                r[d]  = (f & 1) ? 1. : 0.;
                r[a1] = gx[q1];
                r[a2] = gx[q2];
                const Map3 m = map_at(vx, vy, vz, r);
                // tangents t1 = J e_a1, t2 = J e_a2; n dS = +-(t1 x t2) dxi1 dxi2, outward
                const double t1[3] = {m.J[0][a1], m.J[1][a1], m.J[2][a1]}, t2[3] = {m.J[0][a2], m.J[1][a2], m.J[2][a2]};
                double       cr[3] = {t1[1] * t2[2] - t1[2] * t2[1], t1[2] * t2[0] - t1[0] * t2[2], t1[0] * t2[1] - t1[1] * t2[0]};
                const double len   = std::sqrt(cr[0] * cr[0] + cr[1] * cr[1] + cr[2] * cr[2]);
                // (a1, a2) = (d+1, d+2) is a right-handed pair with d: t1 x t2 points along +J e_d; outward on the low face is -
                const double sgn = (f & 1) ? 1. : -1.;
                Point3Tables ff;
                fill_point(ff, m, r, f);
                ff.JxW = len * gw[q1] * gw[q2];
                for (int a = 0; a < 3; ++a)
                  ff.nrm[a] = sgn * cr[a] / len;
                for (int i = 0; i < n; ++i)
                  for (int k = 0; k < n_rt; ++k)
                    F[i * n_rt + k] +=
                      ff.pface[i] * (ff.rt[k][0] * ff.nrm[0] + ff.rt[k][1] * ff.nrm[1] + ff.rt[k][2] * ff.nrm[2]) * ff.JxW;
              }
        }
      // C = F M^-1                                                                       CGQ1:1246
      for (int i = 0; i < n; ++i)
        for (int j = 0; j < n_rt; ++j)
          {
            double s = 0;
            for (int k = 0; k < n_rt; ++k)
              s += F[i * n_rt + k] * M[k * n_rt + j];
            C[i * n_rt + j] = s;
          }
      // Darcy block: the q,i,j,k,l loop                                                  CGQ1:1262-1284
      for (int q = 0; q < 27; ++q)
        for (int i = 0; i < n; ++i)
          for (int j = 0; j < n; ++j)
            for (int k = 0; k < n_rt; ++k)
              for (int l = 0; l < n_rt; ++l)
                A[i * n + j] += (dt * Kc * C[i * n_rt + k] * C[j * n_rt + l] *
                                 (fv[q].rt[k][0] * fv[q].rt[l][0] + fv[q].rt[k][1] * fv[q].rt[l][1] + fv[q].rt[k][2] * fv[q].rt[l][2])) *
                                fv[q].JxW;
      // 2 mu eps(phi_i):eps(phi_j) + c0 phi_i^int phi_j^int at QGauss<3>(3)               CGQ1:1325-1339
      for (int q = 0; q < 27; ++q)
        for (int i = 0; i < n; ++i)
          for (int j = 0; j < n; ++j)
            {
              double ee = 0;
              for (int cc = 0; cc < 3; ++cc)
                for (int dd = 0; dd < 3; ++dd)
                  ee += .5 * (fv[q].ugrad[i][cc][dd] + fv[q].ugrad[i][dd][cc]) * .5 * (fv[q].ugrad[j][cc][dd] + fv[q].ugrad[j][dd][cc]);
              A[i * n + j] += (2. * P.mu * ee + P.capacity * fv[q].pint[i] * fv[q].pint[j]) * fv[q].JxW;
            }
      // lambda div div, -+ alpha coupling and c0 mass AGAIN at QGauss<3>(1)               CGQ1:1341-1358
      for (int i = 0; i < n; ++i)
        {
          const double di = fr.ugrad[i][0][0] + fr.ugrad[i][1][1] + fr.ugrad[i][2][2];
          for (int j = 0; j < n; ++j)
            {
              const double dj = fr.ugrad[j][0][0] + fr.ugrad[j][1][1] + fr.ugrad[j][2][2];
              A[i * n + j] += (P.lambda * di * dj - P.alpha * fr.pint[j] * di + P.capacity * fr.pint[i] * fr.pint[j] +
                               P.alpha * dj * fr.pint[i]) *
                              fr.JxW;
            }
        }
      // right-hand side with f = s = 0 (sandwich problem, CGQ1:582-584, 635-638 zero blocks):
      //   c0 p_old phi_i^int at QGauss(3)  +  alpha div(u_old) phi_i^int at QGauss(1)      CGQ1:1421-1453
      for (int q = 0; q < 27; ++q)
        {
          double p_old = 0;
          for (int i = 0; i < n; ++i)
            p_old += cell_old[i] * fv[q].pint[i];
          for (int i = 0; i < n; ++i)
            b[i] += P.capacity * p_old * fv[q].pint[i] * fv[q].JxW;
        }
      {
        double div_old = 0;
        for (int i = 0; i < n; ++i)
          div_old += cell_old[i] * (fr.ugrad[i][0][0] + fr.ugrad[i][1][1] + fr.ugrad[i][2][2]);
        for (int i = 0; i < n; ++i)
          b[i] += P.alpha * div_old * fr.pint[i] * fr.JxW;
      }
    }
    // ------------------------------------------------------------------------------------------
    // DoFHandler::distribute_dofs(fe) + DoFRenumbering::component_wise(dof_handler, block_component) with
    // block_component = {0,0,0,1,2} (CGQ1:902-919) on hexahedra [deal.II-memory]: first pass walks the active cells in
    // order and numbers, per cell, the not yet numbered vertices 0..7 (3 displacement dofs each, consecutively), then
    // the lines (FE_Q(1), FE_DGQ(0), FE_FaceQ(0) have no line dofs in 3-D), then the quads = faces 0..5 (one FE_FaceQ(0)
    // dof each), then the hex (one FE_DGQ(0) dof); component_wise is a stable sort into [u | p_int | p_face].
    // Face vertex sets: 0 {0,2,4,6}, 1 {1,3,5,7}, 2 {0,1,4,5}, 3 {2,3,6,7}, 4 {0,1,2,3}, 5 {4,5,6,7}.
    // ------------------------------------------------------------------------------------------
    struct Dofs3
    {
      std::int64_t       n_u = 0, n_pint = 0, n_pface = 0;
      std::vector<dof_t> cell_dofs; // [n_cells][31]
      std::int64_t
      n() const
      {
        return n_u + n_pint + n_pface;
      }
    };
    inline Dofs3
    distribute_dofs3(const Mesh3 &m)
    {
      static const int face_v[6][4] = {{0, 2, 4, 6}, {1, 3, 5, 7}, {0, 1, 4, 5}, {2, 3, 6, 7}, {0, 1, 2, 3}, {4, 5, 6, 7}};
      const int        nc = m.n_cells();
      const dof_t      inval = 0xffffffffu;
      std::vector<dof_t> vdof(m.x.size(), inval);
      std::map<std::array<std::uint32_t, 4>, dof_t> fdof;
      std::vector<int>   block;
      Dofs3              d;
      d.cell_dofs.resize((size_t)nc * kNLoc);
      dof_t next = 0;
      for (int c = 0; c < nc; ++c)
        {
          const std::uint32_t *cv = &m.cell_vertices[8 * (size_t)c];
          dof_t               *cd = &d.cell_dofs[(size_t)c * kNLoc];
          for (int v = 0; v < 8; ++v)
            {
              if (vdof[cv[v]] == inval)
                {
                  vdof[cv[v]] = next;
                  next += 3;
                  block.insert(block.end(), {0, 0, 0});
                }
              for (int k = 0; k < 3; ++k)
                cd[3 * v + k] = vdof[cv[v]] + k;
            }
          for (int f = 0; f < 6; ++f)
            {
              std::array<std::uint32_t, 4> key = {cv[face_v[f][0]], cv[face_v[f][1]], cv[face_v[f][2]], cv[face_v[f][3]]};
              std::sort(key.begin(), key.end());
              auto it = fdof.find(key);
              if (it == fdof.end())
                {
                  it = fdof.emplace(key, next++).first;
                  block.push_back(2);
                }
              cd[kPFace0 + f] = it->second;
            }
          cd[kPInt] = next++;
          block.push_back(1);
        }
      std::int64_t cnt[3] = {0, 0, 0};
      for (const int b : block)
        ++cnt[b];
      d.n_u     = cnt[0];
      d.n_pint  = cnt[1];
      d.n_pface = cnt[2];
      std::int64_t       run[3] = {0, cnt[0], cnt[0] + cnt[1]};
      std::vector<dof_t> renum(next);
      for (dof_t i = 0; i < next; ++i)
        renum[i] = (dof_t)run[block[i]]++;
      for (auto &g : d.cell_dofs)
        g = renum[g];
      return d;
    }
    // make_sparsity_pattern(dof_handler, dsp, constraints(empty), false) (CGQ1:934-943): full coupling, one global CSR
    // with ascending columns
    inline Pattern
    make_sparsity_pattern3(const Dofs3 &d, const int n_cells)
    {
      const std::int64_t            n = d.n();
      std::vector<std::set<dof_t>>  rows((size_t)n);
      for (int c = 0; c < n_cells; ++c)
        for (int i = 0; i < kNLoc; ++i)
          for (int j = 0; j < kNLoc; ++j)
            rows[d.cell_dofs[(size_t)c * kNLoc + i]].insert(d.cell_dofs[(size_t)c * kNLoc + j]);
      Pattern p;
      p.rowptr.assign((size_t)n + 1, 0);
      for (std::int64_t r = 0; r < n; ++r)
        p.rowptr[r + 1] = p.rowptr[r] + (std::int64_t)rows[r].size();
      p.colidx.reserve((size_t)p.rowptr[n]);
      for (std::int64_t r = 0; r < n; ++r)
        for (const dof_t c : rows[r])
          p.colidx.push_back((std::int32_t)c);
      return p;
    }
    // ------------------------------------------------------------------------------------------
    // assemble_system (CGQ1:1165-1706) in 3-D: cell loop, traction term on the FK_NEU_U faces with QGauss<2>(3)
    // (CGQ1:1563-1582), constraints of the whole mesh (interpolate_boundary_values is called on the dof_handler inside
    // the cell loop, CGQ1:1662-1679: a cell sees every constrained dof it touches), distribute_local_to_global.
    // ------------------------------------------------------------------------------------------
    inline void
    add_traction(const Problem3 &P, const int c, std::vector<double> &b)
    {
      double vx[8], vy[8], vz[8];
      for (int v = 0; v < 8; ++v)
        {
          const std::uint32_t id = P.mesh.cell_vertices[8 * (size_t)c + v];
          vx[v]                  = P.mesh.x[id];
          vy[v]                  = P.mesh.y[id];
          vz[v]                  = P.mesh.z[id];
        }
      static std::vector<double> gx, gw;
      if (gx.empty())
        qgauss_1d(3, gx, gw);
      for (int f = 0; f < 6; ++f)
        {
          if (!(P.face_kind[6 * (size_t)c + f] & FK_NEU_U))
            continue;
          const int d = f / 2, a1 = (d + 1) % 3, a2 = (d + 2) % 3;
          for (int q2 = 0; q2 < 3; ++q2)
            for (int q1 = 0; q1 < 3; ++q1)
              {
                double r[3];
                r[d]  = (f & 1) ? 1. : 0.;
                r[a1] = gx[q1];
                r[a2] = gx[q2];
                const Map3   m     = map_at(vx, vy, vz, r);
                const double t1[3] = {m.J[0][a1], m.J[1][a1], m.J[2][a1]}, t2[3] = {m.J[0][a2], m.J[1][a2], m.J[2][a2]};
                const double cr[3] = {t1[1] * t2[2] - t1[2] * t2[1], t1[2] * t2[0] - t1[0] * t2[2], t1[0] * t2[1] - t1[1] * t2[0]};
                const double JxW   = std::sqrt(cr[0] * cr[0] + cr[1] * cr[1] + cr[2] * cr[2]) * gw[q1] * gw[q2];
                Point3Tables ff;
                fill_point(ff, m, r, f);
                for (int i = 0; i < kNLoc; ++i)
                  b[i] += (P.traction[0] * ff.uval[i][0] + P.traction[1] * ff.uval[i][1] + P.traction[2] * ff.uval[i][2]) * JxW;
              }
        }
    }
    inline AffineConstraints
    boundary_constraints3(const Problem3 &P, const Dofs3 &dofs)
    {
      static const int  face_v[6][4] = {{0, 2, 4, 6}, {1, 3, 5, 7}, {0, 1, 4, 5}, {2, 3, 6, 7}, {0, 1, 2, 3}, {4, 5, 6, 7}};
      AffineConstraints con;
      for (int c = 0; c < P.mesh.n_cells(); ++c)
        for (int f = 0; f < 6; ++f)
          {
            const std::uint8_t kind = P.face_kind.empty() ? 0 : P.face_kind[6 * (size_t)c + f];
            const dof_t       *idx  = &dofs.cell_dofs[(size_t)c * kNLoc];
            if (kind & FK_DIR_P)
              con.set_inhomogeneity(idx[kPFace0 + f], 0.);
            if (kind & FK_DIR_U)
              for (int k = 0; k < 4; ++k)
                for (int cc = 0; cc < 3; ++cc)
                  con.set_inhomogeneity(idx[3 * face_v[f][k] + cc], 0.);
          }
      return con;
    }
    inline void
    assemble_system3(const Problem3 &P, const Dofs3 &dofs, const Pattern &pat, const std::vector<double> &old_solution,
                     std::vector<double> &values, std::vector<double> &rhs)
    {
      CsrMatrix mat;
      mat.p = &pat;
      mat.val.assign(pat.colidx.size(), 0.);
      rhs.assign((size_t)dofs.n(), 0.);
      const AffineConstraints con = boundary_constraints3(P, dofs);
      std::vector<double>     A, b, cell_old(kNLoc);
      for (int c = 0; c < P.mesh.n_cells(); ++c)
        {
          const dof_t *idx = &dofs.cell_dofs[(size_t)c * kNLoc];
          for (int i = 0; i < kNLoc; ++i)
            cell_old[i] = old_solution.empty() ? 0. : old_solution[idx[i]];
          assemble_cell(P, c, cell_old, A, b);
          if (!P.face_kind.empty())
            add_traction(P, c, b);
          distribute_local_to_global(con, A, b, idx, kNLoc, mat, rhs);
        }
      values.swap(mat.val);
    }
  } // namespace d3
} // namespace orc
