// TEST INFRASTRUCTURE ONLY -- C entry points onto the oracle for ctypes (tests/, smoke(), bench.py
// cpu_baseline / --impl reference).  Never linked into or loaded by poroelasticity_b200.
#include "esm_oracle.hpp"
#include "poro_oracle_3d.hpp"
#include "poro_oracle.hpp"

#include <chrono>
#include <cstring>

using namespace orc;

extern "C"
{
  // Unit square, refine_global(n_refine), all four sides boundary id 0.  bc_mode:
  //   0 = "general case": u and p_face Dirichlet on every boundary face (BASELINE cfg 1-4)
  //   1 = cantilever as shipped in BR1new: u Dirichlet on x=0, traction (0,-1) on y=1, no p Dirichlet
  void *
  orc_create(int element, int n_refine, double distort, std::uint64_t seed, int bc_mode)
  {
    Problem *P = new Problem;
    P->mesh    = make_unit_square(n_refine, distort, seed);
    P->dofs    = distribute_dofs(P->mesh, (Element)element);
    P->pattern = make_sparsity_pattern(P->dofs, P->mesh.n_cells());
    P->kcell.assign(P->mesh.n_cells(), 1.);
    P->face_kind.assign(4 * (size_t)P->mesh.n_cells(), 0);
    const int N = P->mesh.N;
    for (int c = 0; c < P->mesh.n_cells(); ++c)
      {
        std::uint32_t ix, iy;
        morton_decode(c, ix, iy);
        for (int f = 0; f < 4; ++f)
          if (P->mesh.face_bid[4 * c + f] != 255)
            {
              if (bc_mode == 0)
                P->face_kind[4 * c + f] = FK_DIR_U | FK_DIR_P;
              else
                {
                  if (f == 0)
                    P->face_kind[4 * c + f] = FK_DIR_U;
                  else if (f == 3 && (int)iy == N - 1)
                    P->face_kind[4 * c + f] = FK_NEU_U;
                }
            }
      }
    if (bc_mode == 1)
      {
        P->traction[0] = 0;
        P->traction[1] = -1;
      }
    return P;
  }
  void
  orc_destroy(void *h)
  {
    delete (Problem *)h;
  }
  void
  orc_set_material(void *h, double lambda, double mu, double alpha, double c0, double dt)
  {
    Problem *P   = (Problem *)h;
    P->lambda    = lambda;
    P->mu        = mu;
    P->alpha     = alpha;
    P->capacity  = c0;
    P->time_step = dt;
  }
  void
  orc_set_case(void *h, int id, double f_lambda, double f_mu, double f_alpha, double s_capacity, double s_alpha)
  {
    Problem *P       = (Problem *)h;
    P->cs.id         = id;
    P->cs.f_lambda   = f_lambda;
    P->cs.f_mu       = f_mu;
    P->cs.f_alpha    = f_alpha;
    P->cs.s_capacity = s_capacity;
    P->cs.s_alpha    = s_alpha;
  }
  void
  orc_set_kcell(void *h, const double *k)
  {
    Problem *P = (Problem *)h;
    std::copy(k, k + P->mesh.n_cells(), P->kcell.begin());
  }
  // overwrite the vertex coordinates (same topology): rectangles and other affine / bilinear images of the mesh
  void
  orc_set_vertices(void *h, const double *vx, const double *vy)
  {
    Problem *P = (Problem *)h;
    std::copy(vx, vx + P->mesh.vx.size(), P->mesh.vx.begin());
    std::copy(vy, vy + P->mesh.vy.size(), P->mesh.vy.begin());
  }
  void
  orc_set_traction(void *h, double tx, double ty)
  {
    Problem *P     = (Problem *)h;
    P->traction[0] = tx;
    P->traction[1] = ty;
  }
  void
  orc_set_face_kind(void *h, const std::uint8_t *fk)
  {
    Problem *P = (Problem *)h;
    std::copy(fk, fk + P->face_kind.size(), P->face_kind.begin());
  }
  // sizes: [n_cells, n_vertices, n_loc, n_u, n_pint, n_pface, n_dofs, nnz]
  void
  orc_sizes(void *h, std::int64_t *s)
  {
    Problem *P = (Problem *)h;
    s[0]       = P->mesh.n_cells();
    s[1]       = (std::int64_t)P->mesh.vx.size();
    s[2]       = P->dofs.n_loc;
    s[3]       = P->dofs.n_u;
    s[4]       = P->dofs.n_pint;
    s[5]       = P->dofs.n_pface;
    s[6]       = P->dofs.n();
    s[7]       = (std::int64_t)P->pattern.colidx.size();
  }
  void
  orc_get_mesh(void *h, double *vx, double *vy, std::uint32_t *cell_vertices, std::uint8_t *face_kind)
  {
    Problem *P = (Problem *)h;
    std::copy(P->mesh.vx.begin(), P->mesh.vx.end(), vx);
    std::copy(P->mesh.vy.begin(), P->mesh.vy.end(), vy);
    std::copy(P->mesh.cell_vertices.begin(), P->mesh.cell_vertices.end(), cell_vertices);
    std::copy(P->face_kind.begin(), P->face_kind.end(), face_kind);
  }
  void
  orc_get_dofs(void *h, std::uint32_t *cell_dofs)
  {
    Problem *P = (Problem *)h;
    std::copy(P->dofs.cell_dofs.begin(), P->dofs.cell_dofs.end(), cell_dofs);
  }
  void
  orc_get_pattern(void *h, std::int64_t *rowptr, std::int32_t *colidx)
  {
    Problem *P = (Problem *)h;
    std::copy(P->pattern.rowptr.begin(), P->pattern.rowptr.end(), rowptr);
    std::copy(P->pattern.colidx.begin(), P->pattern.colidx.end(), colidx);
  }
  // one cell: local matrix (n_loc^2, row major), local rhs, constraint flags/values per local dof
  void
  orc_local(void *h, int cell, double t, const double *old_solution, double *A, double *b, std::uint8_t *is_con,
            double *g)
  {
    Problem            *P = (Problem *)h;
    std::vector<double> old(old_solution, old_solution + P->dofs.n());
    CellResult          R;
    assemble_cell(*P, cell, t, old, R);
    std::copy(R.A.begin(), R.A.end(), A);
    std::copy(R.b.begin(), R.b.end(), b);
    if (is_con)
      for (int i = 0; i < P->dofs.n_loc; ++i)
        {
          const dof_t d  = P->dofs.cell_dofs[(size_t)cell * P->dofs.n_loc + i];
          auto        it = R.con.lines.find(d);
          is_con[i]      = it != R.con.lines.end();
          g[i]           = is_con[i] ? it->second : 0.;
        }
  }
  // local matrices of cells [first, first+n): out[n][n_loc^2]; returns seconds spent
  double
  orc_local_matrices(void *h, int first, int n, double t, double *out)
  {
    Problem            *P = (Problem *)h;
    std::vector<double> old(P->dofs.n(), 0.);
    CellResult          R;
    const auto          t0 = std::chrono::steady_clock::now();
    const size_t        nn = (size_t)P->dofs.n_loc * P->dofs.n_loc;
    for (int c = first; c < first + n; ++c)
      {
        assemble_cell(*P, c, t, old, R);
        if (out)
          std::copy(R.A.begin(), R.A.end(), out + (size_t)(c - first) * nn);
      }
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  }
  // assemble_system(t); returns seconds spent
  double
  orc_assemble(void *h, double t, const double *old_solution, double *values, double *rhs)
  {
    Problem            *P = (Problem *)h;
    std::vector<double> old(old_solution, old_solution + P->dofs.n()), v, r;
    const auto          t0 = std::chrono::steady_clock::now();
    assemble_system(*P, t, old, v, r);
    const double dtm = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    std::copy(v.begin(), v.end(), values);
    std::copy(r.begin(), r.end(), rhs);
    return dtm;
  }
  void
  orc_step_errors(void *h, double t, const double *solution, double *e3)
  {
    Problem            *P = (Problem *)h;
    std::vector<double> s(solution, solution + P->dofs.n());
    step_errors(*P, t, s, e3);
  }
  // postprocess (CGQ1:1744-1966, BR1new:1189-1246,1352-1606): div[n_cells], mag[n_cells], beta[n_cells][4], vel[n_cells][2]
  void
  orc_postprocess(void *h, const double *solution, double *div, double *mag, double *beta, double *vel)
  {
    Problem            *P = (Problem *)h;
    std::vector<double> s(solution, solution + P->dofs.n()), d, m, b, v;
    postprocess(*P, s, d, m, b, v);
    std::copy(d.begin(), d.end(), div);
    std::copy(m.begin(), m.end(), mag);
    std::copy(b.begin(), b.end(), beta);
    std::copy(v.begin(), v.end(), vel);
  }
  // number of steps the reference's float-accumulated loop runs:
  //   for (n=1, time=dt; time <= t_end; time += dt, ++n)      BR1new:2181, CGQ1:2929
  int
  orc_count_steps(double dt, double t_end)
  {
    int n = 0;
    for (double time = dt; time <= t_end; time += dt)
      ++n;
    return n;
  }
  void
  orc_gauss_jordan(double *a, int n)
  {
    std::vector<double> v(a, a + (size_t)n * n);
    gauss_jordan(v, n);
    std::copy(v.begin(), v.end(), a);
  }
  void
  orc_qgauss(int n, double *x, double *w)
  {
    std::vector<double> xv, wv;
    qgauss_1d(n, xv, wv);
    std::copy(xv.begin(), xv.end(), x);
    std::copy(wv.begin(), wv.end(), w);
  }
  double
  orc_u01(std::uint64_t seed, std::uint64_t id, std::uint64_t k)
  {
    return u01(seed, id, k);
  }

  // ---- EltStiffMat.cc (WGDarcyEquation<2>, WG(Q2,Q2;RT[2])) -------------------------------------------
  void *
  orc_esm_create(int n_refine, double distort, std::uint64_t seed)
  {
    return new esm::Problem(esm::make_problem(n_refine, distort, seed));
  }
  void
  orc_esm_destroy(void *h)
  {
    delete (esm::Problem *)h;
  }
  // sizes: [n_cells, n_vertices, n_loc, n_dofs, nnz]
  void
  orc_esm_sizes(void *h, std::int64_t *s)
  {
    esm::Problem *P = (esm::Problem *)h;
    s[0]            = P->mesh.n_cells();
    s[1]            = (std::int64_t)P->mesh.vx.size();
    s[2]            = esm::kNLoc;
    s[3]            = P->n_dofs;
    s[4]            = (std::int64_t)P->pattern.colidx.size();
  }
  void
  orc_esm_get_mesh(void *h, double *vx, double *vy, std::uint32_t *cell_vertices)
  {
    esm::Problem *P = (esm::Problem *)h;
    std::copy(P->mesh.vx.begin(), P->mesh.vx.end(), vx);
    std::copy(P->mesh.vy.begin(), P->mesh.vy.end(), vy);
    std::copy(P->mesh.cell_vertices.begin(), P->mesh.cell_vertices.end(), cell_vertices);
  }
  void
  orc_esm_get_dofs(void *h, std::uint32_t *cell_dofs)
  {
    esm::Problem *P = (esm::Problem *)h;
    std::copy(P->cell_dofs.begin(), P->cell_dofs.end(), cell_dofs);
  }
  void
  orc_esm_get_pattern(void *h, std::int64_t *rowptr, std::int32_t *colidx)
  {
    esm::Problem *P = (esm::Problem *)h;
    std::copy(P->pattern.rowptr.begin(), P->pattern.rowptr.end(), rowptr);
    std::copy(P->pattern.colidx.begin(), P->pattern.colidx.end(), colidx);
  }
  void
  orc_esm_set_kcell(void *h, const double *k)
  {
    esm::Problem *P = (esm::Problem *)h;
    std::copy(k, k + P->mesh.n_cells(), P->kcell.begin());
  }
  // local matrices / right-hand sides of cells [first, first+n); returns seconds
  double
  orc_esm_local_matrices(void *h, int first, int n, double *A, double *b)
  {
    esm::Problem       *P  = (esm::Problem *)h;
    const auto          t0 = std::chrono::steady_clock::now();
    std::vector<double> Al, bl;
    for (int c = 0; c < n; ++c)
      {
        esm::assemble_cell(*P, first + c, Al, bl);
        if (A)
          std::copy(Al.begin(), Al.end(), A + (size_t)c * esm::kNLoc * esm::kNLoc);
        if (b)
          std::copy(bl.begin(), bl.end(), b + (size_t)c * esm::kNLoc);
      }
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  }
  // assemble_system: CSR values, rhs, solution (boundary values set); returns seconds
  double
  orc_esm_assemble(void *h, double *values, double *rhs, double *solution, int with_boundary_values)
  {
    esm::Problem       *P  = (esm::Problem *)h;
    const auto          t0 = std::chrono::steady_clock::now();
    std::vector<double> v, r, s;
    esm::assemble(*P, v, r, s, with_boundary_values != 0);
    std::copy(v.begin(), v.end(), values);
    std::copy(r.begin(), r.end(), rhs);
    std::copy(s.begin(), s.end(), solution);
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  }
  // postprocess() ESM:662-919: e[0] = L2_error_vel^2, e[1] = L2_error_flux^2
  void
  orc_esm_postprocess(void *h, const double *solution, double *e2)
  {
    esm::Problem       *P = (esm::Problem *)h;
    std::vector<double> s(solution, solution + P->n_dofs);
    esm::postprocess(*P, s, e2);
  }
  // ---- 3-D CGQ1-WG cell loop (poro_oracle_3d.hpp)
  void *
  orc3_create(int n_refine, double distort, std::uint64_t seed)
  {
    d3::Problem3 *P = new d3::Problem3;
    P->mesh         = d3::make_unit_cube(n_refine, distort, seed);
    P->kcell.assign(P->mesh.n_cells(), 1.);
    return P;
  }
  void
  orc3_destroy(void *h)
  {
    delete (d3::Problem3 *)h;
  }
  void
  orc3_sizes(void *h, std::int64_t *s)
  {
    d3::Problem3 *P = (d3::Problem3 *)h;
    s[0]            = P->mesh.n_cells();
    s[1]            = (std::int64_t)P->mesh.x.size();
  }
  void
  orc3_get_mesh(void *h, double *x, double *y, double *z, std::uint32_t *cv)
  {
    d3::Problem3 *P = (d3::Problem3 *)h;
    std::copy(P->mesh.x.begin(), P->mesh.x.end(), x);
    std::copy(P->mesh.y.begin(), P->mesh.y.end(), y);
    std::copy(P->mesh.z.begin(), P->mesh.z.end(), z);
    std::copy(P->mesh.cell_vertices.begin(), P->mesh.cell_vertices.end(), cv);
  }
  void
  orc3_set_material(void *h, double lambda, double mu, double alpha, double c0, double dt)
  {
    d3::Problem3 *P = (d3::Problem3 *)h;
    P->lambda       = lambda;
    P->mu           = mu;
    P->alpha        = alpha;
    P->capacity     = c0;
    P->time_step    = dt;
  }
  void
  orc3_set_kcell(void *h, const double *k)
  {
    d3::Problem3 *P = (d3::Problem3 *)h;
    P->kcell.assign(k, k + P->mesh.n_cells());
  }
  // local matrix [31][31] and rhs [31] of cell c; cell_old = the 31 old-solution values of the cell
  void
  orc3_local(void *h, int c, const double *cell_old, double *A, double *b)
  {
    d3::Problem3       *P = (d3::Problem3 *)h;
    std::vector<double> o(cell_old, cell_old + d3::kNLoc), Av, bv;
    d3::assemble_cell(*P, c, o, Av, bv);
    std::copy(Av.begin(), Av.end(), A);
    std::copy(bv.begin(), bv.end(), b);
  }
  // 3-D numbering and sparsity: sizes[4] = n_u, n_pint, n_pface, nnz; arrays may be NULL (sizes only)
  void
  orc3_dofs_pattern(void *h, std::int64_t *sizes, std::uint32_t *cell_dofs, std::int64_t *rowptr, std::int32_t *colidx)
  {
    d3::Problem3   *P = (d3::Problem3 *)h;
    const d3::Dofs3 d = d3::distribute_dofs3(P->mesh);
    const Pattern   p = d3::make_sparsity_pattern3(d, P->mesh.n_cells());
    sizes[0]          = d.n_u;
    sizes[1]          = d.n_pint;
    sizes[2]          = d.n_pface;
    sizes[3]          = (std::int64_t)p.colidx.size();
    if (cell_dofs)
      std::copy(d.cell_dofs.begin(), d.cell_dofs.end(), cell_dofs);
    if (rowptr)
      std::copy(p.rowptr.begin(), p.rowptr.end(), rowptr);
    if (colidx)
      std::copy(p.colidx.begin(), p.colidx.end(), colidx);
  }
  // boundary data: face_kind[6 n_cells] (NULL: the shipped sandwich boundary, CGQ1:972-981) and the traction vector
  void
  orc3_set_boundary(void *h, const std::uint8_t *face_kind, const double *traction)
  {
    d3::Problem3 *P = (d3::Problem3 *)h;
    if (!face_kind)
      {
        d3::set_sandwich_boundary(*P);
        return;
      }
    P->face_kind.assign(face_kind, face_kind + 6 * (size_t)P->mesh.n_cells());
    for (int k = 0; k < 3; ++k)
      P->traction[k] = traction ? traction[k] : 0.;
  }
  void
  orc3_get_boundary(void *h, std::uint8_t *face_kind, double *traction)
  {
    d3::Problem3 *P = (d3::Problem3 *)h;
    std::copy(P->face_kind.begin(), P->face_kind.end(), face_kind);
    std::copy(P->traction, P->traction + 3, traction);
  }
  // assemble_system in 3-D: values in the CSR order of orc3_dofs_pattern, rhs; old_solution may be NULL (zero)
  void
  orc3_assemble(void *h, const double *old_solution, double *values, double *rhs)
  {
    d3::Problem3       *P = (d3::Problem3 *)h;
    const d3::Dofs3     d = d3::distribute_dofs3(P->mesh);
    const Pattern       p = d3::make_sparsity_pattern3(d, P->mesh.n_cells());
    std::vector<double> old, v, r;
    if (old_solution)
      old.assign(old_solution, old_solution + d.n());
    d3::assemble_system3(*P, d, p, old, v, r);
    std::copy(v.begin(), v.end(), values);
    std::copy(r.begin(), r.end(), rhs);
  }
}
