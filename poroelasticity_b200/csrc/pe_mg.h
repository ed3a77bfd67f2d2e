// Auxiliary-space geometric multigrid (pe_mg.cu).
#pragma once
#include "pe_host.h"
#include "pe_internal.h"

#include <functional>
#include <vector>

namespace pe
{
  struct MgLevel
  {
    int             N = 0; // cells per side on this level
    using stencil_t = float; // storage type of the 9-point block stencils (arithmetic stays FP64, pe_mg.cu)
    DevBuf<stencil_t> st_u;
    DevBuf<double>    x, y, dinv_u, wu[3], bu;
    DevBuf<uint8_t> mask_u, mask_p;
  };
  struct MgHierarchy
  {
    bool                 ready = false;
    int                  N0    = 0;
    std::vector<MgLevel> lv;
    DevBuf<int32_t>      vdof, pint, hedge, vedge;
    DevBuf<double>       coarse_inv_u; // dense inverse of the coarsest displacement operator (+ validity flag)
    double               omega = 0.6, theta = 1.0;
    double               omega2 = 0.; // weight of every second smoothing sweep (<= 0: same as omega)
    int                  nu    = 2;
    int                  fused_n = -1; // largest level of the single-CTA kernel (< 0: built-in)
    // multi-GPU: the first n_dist levels are distributed in row slabs (bounds[l][r] = first vertex row of
    // rank r on level l), the coarser ones run replicated on every rank
    GridComm                          gc;
    int                               n_dist = 0;
    std::vector<std::vector<int64_t>> bounds;
    ~MgHierarchy();
  };
  // 0 = built, 1 = no hierarchy for this mesh (fallback: diagonal preconditioner), -1 = CUDA error
  // to_local: global dof id -> local id on this rank (owned < L <= halo) or -1 if not visible
  int  mg_build_topology(MgHierarchy &mg, const HostMesh &mesh, const HostDofs &dofs,
                         const std::function<int32_t(uint32_t)> &to_local, cudaStream_t s);
  // displacement operators only (the pressure block has its own cell-centred hierarchy, pe_cmg.h)
  int  mg_assemble(MgHierarchy &mg, const double *kcell_fine, double k_scalar, double lambda, double mu, double dt,
                   double gamma, cudaStream_t s, int64_t *launches);
  // V-cycle on level-0 right-hand side lv[0].bu (nc = 2) / lv[0].bp (nc = 1); result in lv[0].wu[2] / wp[2]
  void mg_vcycle(MgHierarchy &mg, int nc, cudaStream_t s, int64_t *launches);
  // choose the distributed levels for gc.world ranks (call after mg_build_topology)
  void mg_set_comm(MgHierarchy &mg, const GridComm &gc);
} // namespace pe
