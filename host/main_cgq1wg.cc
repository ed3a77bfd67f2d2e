// main() of PoroElasCGQ1_WG.cc:3025-3034 on the B200 hot path.
//   ./poroelas_cgq1wg [refinement=3] [dt=1/1024] [t_end=1]        2-D general case (CGQ1:824-831)
//   ./poroelas_cgq1wg 3d [refinement=3]                            the program as shipped: PoroElas_CGQ1_WG<3>, sandwich
//                                                                  problem, T = 1/100, dt = 1/1000 (CGQ1:859-866, 3029)
#include "poroelas_host.hpp"

#include <algorithm>
#include <cstdlib>
#include <string>

int
main(int argc, char **argv)
{
  try
    {
      if (argc > 1 && std::string(argv[1]) == "3d")
        {
          poroelas::PoroElas_CGQ1_WG<3> prog(1. / 100, 1. / 1000, 1., 1., 1., 0.);
          prog.refinement = argc > 2 ? std::atoi(argv[2]) : 3;
          prog.run();
          std::printf("last solve: %d Krylov iterations, relative residual %.3e\n", prog.last_solve().iterations,
                      prog.last_solve().rel_residual);
          const std::vector<double> x = prog.solution3();
          double uz_min = 0.;
          for (size_t i = 2; i < (size_t)prog.n_u(); i += 3)
            uz_min = std::min(uz_min, x[i]);
          std::printf("largest settlement: u_z = %.6e\n", uz_min);
          return 0;
        }
      const int    refinement = argc > 1 ? std::atoi(argv[1]) : 3;
      const double dt         = argc > 2 ? std::atof(argv[2]) : 1. / 1024;
      const double T          = argc > 3 ? std::atof(argv[3]) : 1.0;
      poroelas::PoroElas_CGQ1_WG<2> prog(T, dt, 1., 1., 1., 0.);
      prog.refinement = refinement;
      prog.run();
      std::printf("last solve: %d Krylov iterations, relative residual %.3e\n", prog.last_solve().iterations,
                  prog.last_solve().rel_residual);
    }
  catch (const std::exception &e)
    {
      std::fprintf(stderr, "error: %s\n", e.what());
      return 1;
    }
  return 0;
}
