// main() of PoroElasBR1WG_new.cc:2271-2279 on the B200 hot path.
//   ./poroelas_br1wg [refinement=5] [dt=0.0625] [cantilever=0]
#include "poroelas_host.hpp"

#include <cstdlib>

int
main(int argc, char **argv)
{
  try
    {
      const int    refinement = argc > 1 ? std::atoi(argv[1]) : 5;
      const double dt         = argc > 2 ? std::atof(argv[2]) : 1. / 16;
      const bool   cantilever = argc > 3 && std::atoi(argv[3]) != 0;
      if (cantilever)
        {
          // the shipped state of the file: BR1new:551-558, K = 1e-7 (BR1new:163), zero data
          poroelas::PoroElasBR1WG<2> prog(1.0, 0.01, 1e6 / 7, 1e6 / 28, 0.93, 0.0);
          prog.refinement = refinement;
          prog.cantilever = true;
          prog.case_id    = PE_CASE_ZERO;
          prog.k_scalar   = 1e-7;
          prog.run();
        }
      else
        {
          poroelas::PoroElasBR1WG<2> prog(1.0, dt, 1., 1., 1., 0.);
          prog.refinement  = refinement;
          prog.with_errors = true; // the manufactured solution: final l2-in-time error line as BR1new:2263-2266
          prog.run();
          std::printf("last solve: %d Krylov iterations, relative residual %.3e\n", prog.last_solve().iterations,
                      prog.last_solve().rel_residual);
        }
    }
  catch (const std::exception &e)
    {
      std::fprintf(stderr, "error: %s\n", e.what());
      return 1;
    }
  return 0;
}
