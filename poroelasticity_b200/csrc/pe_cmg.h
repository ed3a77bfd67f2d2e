// Cell-centred aggregation multigrid for the pressure block of the preconditioner (pe_cmg.cu).
//
// After the exact elimination of the interior pressures the face-pressure Schur complement F of the
// weak-Galerkin Darcy block (BR1new:879-901) is preconditioned by a symmetric multiplicative
// two-level method: damped Jacobi on F, correction from the auxiliary space of CELL values,
// damped Jacobi.  The cell space is mapped to the faces by the two-point-flux interface value
//     p_f = (k_1/d_1 p_1 + k_2/d_2 p_2) / (k_1/d_1 + k_2/d_2)
// (harmonic weighting: independent of coefficient jumps across the face, and cells that touch only
// in a vertex do not interact -- the vertex-based Q1 auxiliary space breaks down there: measured
// kappa = 46 for log-normal K with sigma = 2 at N = 16 against 2.3 for this one) and carries the
// two-point-flux 5-point operator T = dt * transmissibilities + gamma |K|.  T is inverted
// approximately by a V-cycle on the nested N x N, N/2 x N/2, ... cell grids with 2x2 aggregates:
// piecewise-constant prolongation, Galerkin coarse operators (again 5-point: the coarse
// transmissibility is the sum of the two fine ones crossing the aggregate boundary), coarse
// correction scaled by 2 (Braess 1995), damped Jacobi smoothing.  Levels with <= 32 x 32 cells run
// in ONE single-CTA kernel in shared memory.  Storage: 3 doubles per cell and level.
#pragma once
#include "pe_host.h"
#include "pe_internal.h"

#include <functional>
#include <vector>

namespace pe
{
  struct CmgLevel
  {
    int            N = 0;
    DevBuf<double> tE, tN, sig; // transmissibility to the east / north neighbour, absorption (reaction + Dirichlet faces)
    DevBuf<double> b, x0, x1;
  };
  struct CellMg
  {
    bool                  ready = false;
    int                   N0    = 0;
    std::vector<CmgLevel> lv;   // lv[0] = finest; levels with N <= 32 are packed behind lv[n_fine] for the coarse kernel
    int                   n_fine = 0;
    // packed coarse hierarchy (levels N_c, N_c/2, ..., 1): coefficient and vector arrays, level after level
    DevBuf<double>  c_tE, c_tN, c_sig;
    int             c_N = 0;
    // face <-> cell transfer on the finest grid (lexicographic cells): weights of the low / high cell of every
    // vertical edge [j*(N+1)+i] and horizontal edge [j*N+i]; 0 for pressure-Dirichlet faces
    DevBuf<double>  wv_lo, wv_hi, wh_lo, wh_hi;
    DevBuf<uint8_t> dir_v, dir_h; // 1: the face carries a pressure Dirichlet value
    double          omega = 0.6, omega2 = 1.2, over = 2.0; // smoothing sweeps alternate between omega and omega2
    int             nu    = 2;
    // multi-GPU: the first n_dist levels are distributed in row slabs of cells, bounds[l][r] = first cell row of rank r
    GridComm                          gc;
    int                               n_dist = 0;
    std::vector<std::vector<int64_t>> bounds;
  };
  void cmg_set_comm(CellMg &cm, const GridComm &gc);

  // 0 = built, 1 = mesh has no refine_global hierarchy, -1 = CUDA error
  int  cmg_build_topology(CellMg &cm, const HostMesh &mesh, cudaStream_t s);
  // (re)discretise all levels: once per assembled matrix.  vx, vy: finest vertex grid, lexicographic
  int  cmg_assemble(CellMg &cm, const double *vx, const double *vy, const double *kcell_morton, double k_scalar, double dt,
                    double gamma, cudaStream_t s, int64_t *launches);
  // rc = Pi_c^T res (owned faces only: d in [0, L)), res indexed through the edge -> dof maps
  // (only the cells / edges of `box` are visited; out = the level-0 right-hand side or a zero-padded local copy)
  void cmg_gather(CellMg &cm, int64_t L, const int32_t *hedge, const int32_t *vedge, const double *res, const CellBox &box,
                  double *out, cudaStream_t s);
  // lv[0].b -> approximate T^{-1} b in lv[0].x0 (result pointer returned)
  const double *cmg_vcycle(CellMg &cm, cudaStream_t s, int64_t *launches);
  // y[d] = y1[d] + (Pi_c e)[d] on owned faces
  void cmg_scatter(CellMg &cm, int64_t L, const int32_t *hedge, const int32_t *vedge, const double *e, const double *y1,
                   double *y, const CellBox &box, cudaStream_t s);
} // namespace pe
