// main() of EltStiffMat.cc:1001-1009 (WGDarcyEquation<2>().run(), ESM:985-996) on the B200 hot path:
// make_grid -> setup_system -> assemble_system -> solve, with the reference's stdout lines.
//   ./eltstiffmat [refinement=1]
#include "../include/poroelas_b200.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <string>
#include <vector>

namespace
{
  struct WGDarcyEquation
  {
    pe_handle *h = nullptr;
    WGDarcyEquation()
    {
      pe_config cfg{};
      cfg.element    = PE_WG_Q2RT2; // fe (FE_DGQ<dim>(2), 1, FE_FaceQ<dim>(2), 1), fe_rt (2): ESM:226-236
      cfg.dim        = 2;
      cfg.device     = -1;
      cfg.world_size = 1;
      check(pe_create(&h, &cfg), "pe_create");
    }
    ~WGDarcyEquation() { pe_destroy(h); }
    void
    check(const int rc, const char *what) const
    {
      if (rc != PE_OK)
        throw std::runtime_error(std::string(what) + ": " + (h ? pe_last_error(h) : "no CUDA device"));
    }
    void
    make_grid(const int refinement) // ESM:242-289
    {
      check(pe_make_unit_square(h, refinement, 0., 0, PE_FACE_DIR_P), "pe_make_unit_square");
      pe_sizes s;
      pe_get_sizes(h, &s);
      std::printf("   Number of active cells: %lld\n", (long long)s.n_cells);
    }
    void
    setup_system() // ESM:294-326
    {
      check(pe_distribute_dofs(h), "pe_distribute_dofs");
      check(pe_make_sparsity_pattern(h), "pe_make_sparsity_pattern");
      pe_sizes s;
      pe_get_sizes(h, &s);
      std::printf("   Number of rt degrees of freedom: %lld\n", (long long)(s.n_dofs + 3 * s.n_cells)); // FE_RaviartThomas(2): 3 per face + 12 per cell
      std::printf("   Number of dof: %lld\n", (long long)s.n_dofs);
    }
    void
    assemble_system() // ESM:333-578
    {
      check(pe_set_case(h, PE_CASE_DARCY_COSCOS, nullptr), "pe_set_case");
      check(pe_assemble_system(h, 0., 1., nullptr), "pe_assemble_system");
    }
    pe_solve_stats
    solve(std::vector<double> &solution) // ESM:583-590: SolverControl(1000, 1e-8 * ||rhs||), CG + identity
    {
      pe_sizes s;
      pe_get_sizes(h, &s);
      solution.resize((size_t)s.n_dofs);
      pe_solve_stats st{};
      check(pe_solve_time_step(h, 1e-8, 1000, solution.data(), &st), "pe_solve_time_step");
      return st;
    }
    void
    run(const int refinement) // ESM:985-996
    {
      make_grid(refinement);
      setup_system();
      assemble_system();
      std::vector<double> solution;
      const pe_solve_stats st = solve(solution);
      std::printf("   CG: %d iterations, residual %.3e (relative %.3e)\n", st.iterations, st.residual_norm, st.rel_residual);
      // the nine interior values of cell 0 against p = cos(pi x) cos(pi y) (ESM:187), a quick sanity print
      pe_sizes s;
      pe_get_sizes(h, &s);
      std::vector<uint32_t> cd((size_t)s.n_cells * 21);
      pe_get_dofs(h, cd.data());
      std::printf("   p_h(cell 0 centre) = %.6f\n", solution[cd[12 + 4]]);
    }
  };
} // namespace

int
main(int argc, char **argv)
{
  try
    {
      WGDarcyEquation().run(argc > 1 ? std::atoi(argv[1]) : 1);
    }
  catch (const std::exception &e)
    {
      std::fprintf(stderr, "error: %s\n", e.what());
      return 1;
    }
  return 0;
}
