// Three-field ("total pressure") form of the Biot system that the Krylov solver works on (pe_k3.cu).
//
// The reference's matrix (BR1new:904-925, CGQ1:1325-1358) is
//     A = [ A_mu + lambda B^T W^-1 B    -alpha B^T ]        B_Ki = int_K div phi_i (cell average / 1-point rule),
//         [ alpha B                      C          ]        W = diag(w_K), C = c0 M + dt S
// and for lambda >> mu every preconditioner of the u block built from Jacobi / Q1 multigrid breaks
// down (volumetric locking of the smoother; measured: MINRES diverges for lambda >= 10).  With the
// auxiliary cell unknown  xi = -lambda W^-1 B u + alpha p_int  the SAME solution (u, p) solves
//     [ A_mu        -B^T              0               ] [u ]   [ b_u ]
//     [ -B          -W/lambda         (alpha/lambda) W ] [xi] = [ 0   ]
//     [ 0           (alpha/lambda) W  -(C + alpha^2/lambda W_int) ] [p ]   [-b_p ]
// which is symmetric and has the lambda-robust block-diagonal preconditioner
//     diag( A_mu , (1/lambda + 1/(2 mu)) W , C + alpha^2/lambda W_int )        (Lee, Mardal, Winther 2017)
// K3 is stored as its own pruned CSR (the structurally zero u <-> p_face block of the reference pattern
// is dropped: 148 instead of 221 entries per BR1 cell), filled from a second assembly pass A0 with
// (lambda, alpha) = (0, 1) by a gather.
//
// Vector layout ("physical" index): [0,L) owned (u | p_int | p_face), [L,L+H) their halo, then the
// owned xi (Lx = owned cells) and the halo xi (Hx).  Row r3 of K3 ("logical" index, owned unknowns
// only) lives at physical index r3 + (r3 >= L ? H : 0).
#pragma once
#include "pe_host.h"
#include "pe_internal.h"

namespace pe
{
  struct K3System
  {
    bool    built = false, values_valid = false;
    bool    xi_scaled = false; // xi rows of val multiplied by xi_tau / sqrt(w) (GMRES path; MINRES needs the symmetric K3)
    double  xi_tau = 0.;       // 1 / (1/lambda_s + 1/(2 mu)), 0 when the rows are not scaled
    bool    p_scaled = false;  // pressure rows of val multiplied by p_scale (GMRES path)
    DevBuf<double> p_scale;    // [L - pint_begin]
    int64_t L = 0, H = 0, Lx = 0, Hx = 0;
    int64_t n_u = 0, pint_begin = 0, pint_end = 0, hu_end = 0, hpi_end = 0;
    int64_t nnz = 0, nnz_u = 0, n_rowblk = 0, n_con = 0;
    double  lambda_s = 1., lambda0 = 0., alpha = 1.;
    DevBuf<int64_t>  rowptr, pos;
    DevBuf<int32_t>  colidx, rowblk, con_rows;
    DevBuf<uint32_t> src;
    DevBuf<double>   val, w, values0, diag0;
    int64_t n_rows() const { return L + Lx; }
    int64_t n_phys() const { return L + H + Lx + Hx; }
  };

  // structure: once per sparsity pattern.  d_rowptr / d_colidx: the reference pattern in local numbering.
  int k3_build_structure(K3System &k, const HostMesh &mesh, const HostDofs &dofs, const Partition &part,
                         const int64_t *d_rowptr, const int32_t *d_colidx, cudaStream_t s);
  // values: once per assembled matrix.  values0 = A0 (same CSR as the reference pattern).
  int k3_fill_values(K3System &k, double lambda_s, double alpha, double mu, bool scale_xi_rows, cudaStream_t s,
                     int64_t *launches);
  // GMRES path: scale the pressure rows of val by sqrt(2 mu dinv) (dinv: inverse diagonal of the pressure block of the
  // preconditioner, local row index); the right-hand side / preconditioner input are scaled with k3_p_vec_scale
  int  k3_scale_p_rows(K3System &k, const double *dinv, double mu, cudaStream_t s, int64_t *launches);
  void k3_p_vec_scale(const K3System &k, bool inverse, const double *v, double *out, cudaStream_t s);
  // b3 = (b_u [constrained rows rescaled to the A0 diagonal] | -b_p | 0)
  int k3_make_rhs(const K3System &k, const double *rhs, const int64_t *rowptr, const int32_t *colidx, const double *values_a,
                  double *b3, cudaStream_t s);
  // inverse block diagonal of K3 in the physical layout (fallback preconditioner)
  int k3_make_pinv(const K3System &k, double mu, double *pinv3, cudaStream_t s);
  // xi = (lambda_s / w) (-B u) + alpha p_int from the (u, p) part of x3
  int k3_init_xi(const K3System &k, double *x3, cudaStream_t s);
  // z_xi = r_xi / ((1/lambda_s + 1/(2 mu)) w)
  int k3_precond_xi(const K3System &k, double mu, const double *r3, double *z3, cudaStream_t s);
} // namespace pe
