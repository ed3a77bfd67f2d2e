// main() of PoroElasCGQ1_WG.cc:3025-3034 (2-D general case) on the B200 hot path.
//   ./poroelas_cgq1wg [refinement=3] [dt=1/1024] [t_end=1]
#include "poroelas_host.hpp"

#include <cstdlib>

int
main(int argc, char **argv)
{
  try
    {
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
