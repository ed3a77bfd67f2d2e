// Block preconditioner (pe_precond.cu).
#pragma once
#include "pe_host.h"
#include "pe_internal.h"
#include "pe_cmg.h"
#include "pe_mg.h"

#include <functional>

namespace pe
{
  struct SubCsr
  {
    int              n_rows = 0, n_blk = 0;
    int64_t          nnz = 0;
    DevBuf<int32_t>  rows, rp, col, rowblk;
    // compact copy of the block's values, refreshed once per matrix.  Stored in single precision, used in FP64
    // arithmetic: the products are HBM bound (ncu: 4.5-5.4 TB/s) and a preconditioner may be any fixed
    // symmetric positive definite operator (rounding the entries of the symmetric blocks keeps them symmetric)
    using value_t = float;
    DevBuf<value_t>  val;
    DevBuf<uint32_t> src; // index into the assembled CSR values
    DevBuf<uint8_t>  neg; // optional: 1 = the entry enters the compact copy with flipped sign
  };
  struct BlockPrecond
  {
    bool            ready = false, has_bubbles = false;
    int64_t         n_u = 0, n = 0; // owned displacement rows, owned rows
    int             n_pint = 0;
    // multi-GPU hooks (null on one GPU)
    void (*halo)(void *ctx, double *x)             = nullptr;
    void (*allreduce)(void *ctx, double *v, int n) = nullptr;
    void (*allreduce2)(void *ctx, const double *in, double *out, int n) = nullptr; // out of place
    void *comm_ctx                                 = nullptr;
    void (*halo_p)(void *ctx, double *x)           = nullptr; // pressure blocks only
    void (*set_lane)(void *ctx, int lane)          = nullptr; // non-null: the pressure pipeline runs on lane 1 (own stream + communicator)
    CellBox         box;                 // cells this rank assembles (all cells on one GPU)
    DevBuf<double>  bu_local, rc_local;  // multi-GPU: this rank's contribution to the grid right-hand sides (zero outside box)
    SubCsr          Abu, Avb, Sif, Sfi, Spf; // Spf: p_face rows x [-S_fi | S_ff]
    SubCsr          Bt;                      // -B^T: displacement rows x xi columns (three-field physical indices)
    DevBuf<uint8_t> cls;
    DevBuf<int32_t> pint_rows;
    DevBuf<double>  pint_area, t, t2, t3, t4, t5, dinv, wdinv; // wdinv = omega_F / diag(F) on the face rows
    double          omega_f = 0.65;
    const double   *values = nullptr;
    // single GPU: the pressure pipeline runs on a forked stream beside the displacement V-cycle
    cudaStream_t    s2 = nullptr;
    cudaEvent_t     ev_fork = nullptr, ev_join = nullptr;
    ~BlockPrecond();
  };
  int  precond_build_structure(BlockPrecond &bp, const HostMesh &mesh, const HostDofs &d, const HostPattern &pat,
                               const Partition &part, const std::function<int32_t(uint32_t)> &to_local, cudaStream_t s);
  int  precond_update(BlockPrecond &bp, const double *values, const double *diag, double gamma, cudaStream_t s,
                      int64_t *launches);
  void precond_apply(BlockPrecond &bp, MgHierarchy &mg, CellMg &cm, const double *pinv, const double *r, double *z,
                     cudaStream_t s, int64_t *launches, const double *z_xi = nullptr, const double *rp = nullptr);
} // namespace pe
