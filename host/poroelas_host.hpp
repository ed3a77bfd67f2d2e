// Host-side C++ mirror of the reference's classes for the hot path, on top of the C ABI
// (include/poroelas_b200.h).  Same member names as PoroElasBR1WG / PoroElas_CGQ1_WG
// (PoroElasBR1WG_new.cc:62-137): make_grid_and_dofs(), assemble_system(t), solve_time_step(), postprocess(),
// postprocess_errors(t, n), postprocess_darcy_velocity(n), run(); same stdout lines (BR1new:599-610, 2183-2185,
// 2263-2266).  No numerics live here.
//
// Defaults: the default constructors carry the reference's commented "general case" blocks (manufactured coscos
// solution, BR1new:532-539 / CGQ1:824-831), which is what BASELINE configs 1-4 run; the values the files SHIP with
// (cantilever bracket BR1new:551-558, 3-D sandwich CGQ1:859-866) are selected with `cantilever = true` /
// the six-argument constructors.
#pragma once
#include "../include/poroelas_b200.h"

#include <cmath>
#include <cstdio>
#include <fstream>
#include <iostream>
#include <stdexcept>
#include <string>
#include <vector>

namespace poroelas
{
  template <int dim>
  class BiotProgram
  {
  public:
    BiotProgram(const pe_element element, const double total_time_, const double time_step_, const double lambda_,
                const double mu_, const double alpha_, const double capacity_)
      : total_time(total_time_)
      , time_step(time_step_)
      , number_timestep((unsigned int)(total_time_ / time_step_))
      , timestep_number(1)
      , lambda(lambda_)
      , mu(mu_)
      , alpha(alpha_)
      , capacity(capacity_)
    {
      static_assert(dim == 2 || dim == 3, "dim = 2 or 3");
      pe_config cfg{};
      cfg.element    = element;
      cfg.dim        = dim;
      cfg.device     = -1;
      cfg.world_size = 1;
      check(pe_create(&h, &cfg), "pe_create");
    }
    ~BiotProgram() { pe_destroy(h); }
    BiotProgram(const BiotProgram &) = delete;

    // problem variants the reference selects by commenting blocks in and out
    int    refinement    = 5;
    int    case_id       = PE_CASE_COSCOS;
    int    boundary_kinds = PE_FACE_DIR_U | PE_FACE_DIR_P;
    double k_scalar      = 1.0;
    double rel_tol       = 1e-10;
    int    max_iterations = 100000;
    bool   cantilever    = false; // BR1new as shipped: u = 0 on x = 0, traction (0,-1) on y = 1, K = 1e-7
    bool   with_errors   = false; // postprocess_errors every step + the final l2-in-time line (BR1new:2193, 2251-2266)
    int    velocity_step = -1;    // postprocess_darcy_velocity writes velocity.txt at this step (BR1new:1356: t == 10)

    void
    run()
    {
      std::cout << "Solving problem in " << dim << " space dimensions." << std::endl;
      make_grid_and_dofs();
      // old_solution = 0 (BR1new:2163): the device solution starts at zero.
      // The 3-D program hard-codes its loop bound: `time <= 0.011` (CGQ1:2929, total_time = 1/100).
      const double t_end = dim == 3 ? 0.011 : total_time;
      for (timestep_number = 1, time = time_step; time <= t_end; time += time_step, ++timestep_number)
        {
          std::cout << "Time step " << timestep_number << " at t=" << time << " time_step " << time_step << std::endl;
          assemble_system(time);
          solve_time_step();
          if (dim == 3)
            continue; // postprocess / error norms / velocity output are not built in 3-D
          postprocess();
          if (with_errors)
            postprocess_errors(time, timestep_number);
          postprocess_darcy_velocity(timestep_number);
          // old_solution = solution (BR1new:2241): stays on the device
        }
      if (with_errors)
        std::cout << "L2L2Error_intp " << std::sqrt(err_acc[0]) << ", "
                  << "L2L2Error_u " << std::sqrt(err_acc[1]) << ", "
                  << "L2h1semiError_u " << std::sqrt(err_acc[2]) << ", "
                  << "L2h1Error_u " << std::sqrt(err_acc[1] + err_acc[2]) << std::endl;
    }

    // numerical dilation and |u| per cell of the last step (BR1new:1189-1246, CGQ1:1849-1877)
    const std::vector<double> &dilation() const { return div_displacement; }
    const std::vector<double> &magnitude() const { return displacement_magnitude; }

    // 3-D: the solution of the last step (pe_solve_time_step hands it out; there is no pe_get_solution in 3-D yet)
    const std::vector<double> &solution3() const { return last_solution; }
    int64_t
    n_u() const
    {
      pe_sizes s;
      pe_get_sizes(h, &s);
      return s.n_u;
    }
    const pe_solve_stats &last_solve() const { return stats; }
    pe_handle            *handle() { return h; }
    std::vector<double>
    solution() const
    {
      pe_sizes s;
      pe_get_sizes(h, &s);
      std::vector<double> x((size_t)s.n_owned_rows);
      pe_get_solution(h, x.data());
      return x;
    }

  protected:
    // 3-D: hyper_cube + refine_global (CGQ1:886-887), boundary ids of the sandwich problem (CGQ1:972-981, installed by
    // pe_make_unit_cube), Coefficient = 1e-8 for 1/4 <= z <= 3/4 (CGQ1:176-189) evaluated at the cell centre
    void
    make_grid_and_dofs3()
    {
      check(pe_make_unit_cube(h, refinement, 0., 0), "pe_make_unit_cube");
      int64_t nv = 0, nc = 0;
      check(pe_get_mesh3(h, &nv, &nc, nullptr, nullptr, nullptr, nullptr), "pe_get_mesh3");
      std::vector<double>   x((size_t)nv), y((size_t)nv), z((size_t)nv), k((size_t)nc);
      std::vector<uint32_t> cv((size_t)8 * nc);
      check(pe_get_mesh3(h, nullptr, nullptr, x.data(), y.data(), z.data(), cv.data()), "pe_get_mesh3");
      for (int64_t c = 0; c < nc; ++c)
        {
          double zc = 0.;
          for (int v = 0; v < 8; ++v)
            zc += z[cv[(size_t)8 * c + v]] / 8.;
          k[(size_t)c] = (zc >= 0.25 && zc <= 0.75) ? 1e-8 : 1.;
        }
      check(pe_distribute_dofs(h), "pe_distribute_dofs");
      check(pe_make_sparsity_pattern(h), "pe_make_sparsity_pattern");
      pe_sizes s;
      pe_get_sizes(h, &s);
      std::cout << "Number of active cells: " << s.n_cells << std::endl
                << "Number of degrees of freedom: " << s.n_dofs << " (" << s.n_u << '+' << s.n_pint << '+' << s.n_pface
                << ')' << std::endl
                << std::endl;
      check(pe_set_material(h, lambda, mu, alpha, capacity, 1.0, k.data()), "pe_set_material");
    }
    void
    make_grid_and_dofs()
    {
      if (dim == 3)
        {
          make_grid_and_dofs3();
          return;
        }
      check(pe_make_unit_square(h, refinement, 0., 0, cantilever ? 0 : boundary_kinds), "pe_make_unit_square");
      check(pe_distribute_dofs(h), "pe_distribute_dofs");
      check(pe_make_sparsity_pattern(h), "pe_make_sparsity_pattern");
      pe_sizes s;
      pe_get_sizes(h, &s);
      std::cout << "Number of active cells: " << s.n_cells << std::endl
                << "Number of degrees of freedom: " << s.n_dofs << " (" << s.n_u << '+' << s.n_pint << '+' << s.n_pface
                << ')' << std::endl
                << std::endl;
      std::cout << "   Number of rt degrees of freedom: " << s.n_pface << std::endl;
      if (cantilever)
        {
          // set_boundary_id loop of BR1new:652-666 + the Neumann block BR1new:964-990
          std::vector<uint8_t> fk(4 * (size_t)s.n_cells);
          pe_get_mesh(h, nullptr, nullptr, nullptr, nullptr);
          std::vector<uint32_t> cells;
          std::vector<uint8_t>  faces, kinds;
          const int64_t         N = int64_t(1) << refinement;
          for (int64_t c = 0; c < s.n_cells; ++c)
            {
              int64_t ix = 0, iy = 0;
              for (int b = 0; b < 16; ++b)
                {
                  ix |= ((c >> (2 * b)) & 1) << b;
                  iy |= ((c >> (2 * b + 1)) & 1) << b;
                }
              if (ix == 0)
                {
                  cells.push_back((uint32_t)c);
                  faces.push_back(0);
                  kinds.push_back(PE_FACE_DIR_U);
                }
              if (iy == N - 1)
                {
                  cells.push_back((uint32_t)c);
                  faces.push_back(3);
                  kinds.push_back(PE_FACE_NEU_U);
                }
            }
          const double traction[2] = {0., -1.};
          check(pe_set_boundary(h, (int64_t)cells.size(), cells.data(), faces.data(), kinds.data(), traction),
                "pe_set_boundary");
        }
      check(pe_set_material(h, lambda, mu, alpha, capacity, k_scalar, nullptr), "pe_set_material");
      check(pe_set_case(h, case_id, nullptr), "pe_set_case");
    }
    void
    assemble_system(const double t)
    {
      check(pe_assemble_system(h, t, time_step, nullptr), "pe_assemble_system");
    }
    void
    solve_time_step()
    {
      if (dim == 3)
        {
          pe_sizes s;
          pe_get_sizes(h, &s);
          last_solution.resize((size_t)s.n_dofs);
        }
      check(pe_solve_time_step(h, rel_tol, max_iterations, dim == 3 ? last_solution.data() : nullptr, &stats),
            "pe_solve_time_step");
    }
    // postprocess()  BR1new:1189-1246 / CGQ1:1744-1879: cell averages of div u_h and |u_h|
    void
    postprocess()
    {
      pe_sizes s;
      pe_get_sizes(h, &s);
      div_displacement.resize((size_t)s.n_cells);
      displacement_magnitude.resize((size_t)s.n_cells);
      check(pe_postprocess(h, div_displacement.data(), displacement_magnitude.data(), nullptr, nullptr), "pe_postprocess");
    }
    // postprocess_errors(t, n)  BR1new:1251-1349: squared errors of this step times dt, summed over the steps
    void
    postprocess_errors(const double t, const unsigned int /*n*/)
    {
      double e[3];
      check(pe_step_errors(h, t, e), "pe_step_errors");
      for (int k = 0; k < 3; ++k)
        err_acc[k] += time_step * e[k];
    }
    // postprocess_darcy_velocity(n)  BR1new:1352-1606: cell-centre Darcy velocities -> velocity.txt
    void
    postprocess_darcy_velocity(const unsigned int n)
    {
      if ((int)n != velocity_step)
        return;
      std::cout << "run the final time " << std::endl;
      pe_sizes s;
      pe_get_sizes(h, &s);
      std::vector<double> vel((size_t)2 * s.n_cells);
      check(pe_postprocess(h, nullptr, nullptr, nullptr, vel.data()), "pe_postprocess");
      std::ofstream out("velocity.txt");
      for (int64_t i = 0; i < s.n_cells; ++i)
        out << vel[2 * i] << " " << vel[2 * i + 1] << " " << 0 << std::endl;
    }
    void
    check(const int rc, const char *what) const
    {
      if (rc != PE_OK)
        throw std::runtime_error(std::string(what) + ": " + (h ? pe_last_error(h) : "no handle"));
    }

    pe_handle     *h = nullptr;
    pe_solve_stats stats{};
    std::vector<double> div_displacement, displacement_magnitude, last_solution;
    double              err_acc[3] = {0., 0., 0.};
    double         total_time, time = 0., time_step;
    unsigned int   number_timestep, timestep_number;
    const double   lambda, mu, alpha, capacity;
  };

  // PoroElasBR1WG<2>: "general case" constructor block BR1new:532-539
  template <int dim>
  struct PoroElasBR1WG : BiotProgram<dim>
  {
    PoroElasBR1WG()
      : BiotProgram<dim>(PE_BR1_WG, 1.0, 1. / 16, 1., 1., 1.0, 0.0)
    {}
    PoroElasBR1WG(double T, double dt, double lambda, double mu, double alpha, double c0)
      : BiotProgram<dim>(PE_BR1_WG, T, dt, lambda, mu, alpha, c0)
    {}
  };
  // PoroElas_CGQ1_WG<2>: "general case" constructor block CGQ1:824-831; PoroElas_CGQ1_WG<3>(1./100, 1./1000, 1, 1, 1, 0) with
  // refinement = 3 is the program the file ships (CGQ1:859-866, 3029)
  template <int dim>
  struct PoroElas_CGQ1_WG : BiotProgram<dim>
  {
    PoroElas_CGQ1_WG()
      : BiotProgram<dim>(PE_CGQ1_WG, 1.0, 1. / 1024, 1., 1., 1.0, 0.0)
    {}
    PoroElas_CGQ1_WG(double T, double dt, double lambda, double mu, double alpha, double c0)
      : BiotProgram<dim>(PE_CGQ1_WG, T, dt, lambda, mu, alpha, c0)
    {}
  };
} // namespace poroelas
