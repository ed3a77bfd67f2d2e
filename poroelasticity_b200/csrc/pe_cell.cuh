// Device-side cell kernel math (FP64, sm_100a): one thread builds the whole local matrix and local
// right-hand side of one cell in registers and hands every entry to an "emit" functor.
//
// Replaces the cell loop body of assemble_system:
//   BR1-WG  : PoroElasBR1WG_new.cc:794-961   (A1-A6 of SURVEY.md 8(a))
//   CGQ1-WG : PoroElasCGQ1_WG.cc:1168-1453
// in lean form: for RT0 the weak-gradient right-hand side F is the
// constant matrix F_int = (+1,-1,+1,-1), F_face = diag(-1,+1,-1,+1) on ANY quadrilateral, so the
// Darcy block is dt k F M^{-1} F^T with M the RT0 Gram matrix; the strain block is assembled from
// the three gradient-moment matrices of the 8 scalar shape functions.
#pragma once
#include <cstdint>

namespace pe
{
  struct DevParams
  {
    double lambda, mu, alpha, c0, dt;
    double traction[2];
    int    case_id;
    // time-dependent factors of the "coscos" case, evaluated on the host once per assemble:
    double f_coef; // pi (f_lambda + 2 f_mu - f_alpha) sin(pi t/2)              BR1new:409-412
    double s_coef; // (pi/2) cos(pi t/2)(s_capacity + s_alpha) + 2 pi^2 sin(pi t/2)  BR1new:376-378
    double u_coef; // sin(pi t/2) / (2 pi)                                      BR1new:207-208
    double p_coef; // sin(pi t/2)                                               BR1new:209-210
  };

  // QGauss<1>(3) on [0,1] (deal.II quadrature_lib.cc): 0.5 -+ sqrt(3/5)/2, 0.5
  __device__ constexpr double kGx[3] = {0.11270166537925831148, 0.5, 0.88729833462074168852};
  __device__ constexpr double kGw[3] = {5. / 18., 8. / 18., 5. / 18.};

  template <int EL>
  struct Elem;
  template <>
  struct Elem<0> // BR1-WG: (v0x v0y .. v3x v3y | b0 pf0 | b1 pf1 | b2 pf2 | b3 pf3 | pint)
  {
    static constexpr int NL = 17, NS = 8, NU = 12, PINT = 16;
    __host__ __device__ static constexpr int
    pface(int f)
    {
      return 9 + 2 * f;
    }
    __host__ __device__ static constexpr int
    bubble(int f)
    {
      return 8 + 2 * f;
    }
  };
  template <>
  struct Elem<1> // CGQ1-WG: (v0x .. v3y | pf0 pf1 pf2 pf3 | pint)
  {
    static constexpr int NL = 13, NS = 4, NU = 8, PINT = 12;
    __host__ __device__ static constexpr int
    pface(int f)
    {
      return 8 + f;
    }
    __host__ __device__ static constexpr int
    bubble(int)
    {
      return -1;
    }
  };

  // ---------------------------------------------------------------------------------------------
  // Scatter-offset table layout.  Entry (i,j) of the local matrix owns one byte (position inside the
  // CSR row of dof i + first-writer flag).  Bytes are grouped by the sweep that emits the entry
  // (see cell_kernel) and packed four to a 32-bit word, words stored [word][cell]: a kernel stages
  // exactly the words of its sweep through shared memory with independent coalesced loads --
  // one memory round trip instead of one per entry.
  // ---------------------------------------------------------------------------------------------
  template <int EL>
  struct SlotMap
  {
    using EE                  = Elem<EL>;
    static constexpr int NL   = EE::NL;
    static constexpr int NCX  = EL == 0 ? 6 : 4;  // displacement dofs per component
    static constexpr int NU   = EE::NU;
    // 0: x-displacement dof, 1: y-displacement dof, 2: p_face, 3: p_int
    __host__ __device__ static constexpr int
    kind(int i)
    {
      return i < 8 ? (i & 1) : (i == EE::PINT ? 3 : ((EL == 0 && ((i - 8) & 1) == 0) ? ((i - 8) / 2 < 2 ? 0 : 1) : 2));
    }
    __host__ __device__ static constexpr int
    ci(int i) // index among the dofs of its own component (vertices 0..3, bubbles 4,5)
    {
      return i < 8 ? (i >> 1) : 4 + (((i - 8) / 2) & 1);
    }
    __host__ __device__ static constexpr int
    ui(int i) // index among all displacement dofs
    {
      return i < 8 ? i : 8 + (i - 8) / 2;
    }
    __host__ __device__ static constexpr int
    p5(int i) // index in the pressure block: p_int 0, p_face f -> 1 + f
    {
      return i == EE::PINT ? 0 : 1 + (EL == 0 ? (i - 9) / 2 : i - 8);
    }
    static constexpr int NX = NCX * NCX + 2 * NU + 8 * NU; // entries emitted by sweep X
    static constexpr int NQ = 2 * NCX * NCX;
    static constexpr int NY = NCX * NCX + 25;
    static constexpr int WX = (NX + 3) / 4, WQ = (NQ + 3) / 4, WY = (NY + 3) / 4;
    __host__ __device__ static constexpr int
    word_begin(int sw)
    {
      return sw == 0 ? 0 : sw == 1 ? WX : sw == 2 ? WX + WQ : WX + WQ + WY;
    }
    // byte slot of entry (i,j): sweep-major, word aligned per sweep
    __host__ __device__ static constexpr int
    slot(int i, int j)
    {
      const int ki = kind(i), kj = kind(j);
      if (ki == 0 && kj == 0)
        return ci(i) * NCX + ci(j);
      if (ki <= 1 && kj == 3)
        return NCX * NCX + ui(i);
      if (ki == 3 && kj <= 1)
        return NCX * NCX + NU + ui(j);
      if (ki <= 1 && kj == 2)
        return NCX * NCX + 2 * NU + ui(i) * 4 + (p5(j) - 1);
      if (ki == 2 && kj <= 1)
        return NCX * NCX + 6 * NU + (p5(i) - 1) * NU + ui(j);
      if (ki == 1 && kj == 0)
        return 4 * WX + ci(i) * NCX + ci(j);
      if (ki == 0 && kj == 1)
        return 4 * WX + NCX * NCX + ci(i) * NCX + ci(j);
      if (ki == 1 && kj == 1)
        return 4 * (WX + WQ) + ci(i) * NCX + ci(j);
      return 4 * (WX + WQ) + NCX * NCX + p5(i) * 5 + p5(j);
    }
  };

  // ---------------------------------------------------------------------------------------------
  // Numbering of the STRUCTURALLY NON-ZERO entries of the local matrix (the u <-> p_face block is zero,
  // SURVEY 8(a) row A5): row-major over (i, j), zeros skipped.  193 (BR1) / 105 (CGQ1) entries: one byte.
  // Used by the gather assembly (pe_assemble.cu): entry nz of cell c is staged at T[nz][c].
  // Closed-form arithmetic so that it folds to a constant for compile-time (i, j).  -1 = structural zero.
  // ---------------------------------------------------------------------------------------------
  template <int EL>
  __host__ __device__ constexpr int
  nzslot(const int i, const int j)
  {
    using SM     = SlotMap<EL>;
    const int ki = SM::kind(i), kj = SM::kind(j); // 0,1: displacement, 2: p_face, 3: p_int
    if ((ki <= 1 && kj == 2) || (ki == 2 && kj <= 1))
      return -1;
    constexpr int NU = Elem<EL>::NU, USZ = NU + 1, PSZ = 5; // non-zeros in a u row / a p_face row
    // first slot of row i
    int base = 0;
    if (EL == 0)
      base = i < 8 ? USZ * i : (i == 16 ? 8 * USZ + 4 * (USZ + PSZ) : 8 * USZ + ((i - 8) / 2) * (USZ + PSZ) + (((i - 8) & 1) ? USZ : 0));
    else
      base = i < 8 ? USZ * i : (i == 12 ? 8 * USZ + 4 * PSZ : 8 * USZ + (i - 8) * PSZ);
    // rank of column j among the non-zeros of the row
    int rank = 0;
    if (ki <= 1)
      rank = kj == 3 ? NU : SM::ui(j);
    else if (ki == 2)
      rank = kj == 3 ? 4 : SM::p5(j) - 1;
    else
      rank = j;
    return base + rank;
  }
  template <int EL>
  struct NzCount
  {
    static constexpr int value = EL == 0 ? 193 : 105;
  };

  struct Geom
  {
    double x0, y0;                     // vertex 0
    double a1x, a1y, a2x, a2y, a3x, a3y; // x(xi,eta) = x0 + a1 xi + a2 eta + a3 xi eta
    __device__ __forceinline__ void
    set(const double vx[4], const double vy[4])
    {
      x0  = vx[0];
      y0  = vy[0];
      a1x = vx[1] - vx[0];
      a1y = vy[1] - vy[0];
      a2x = vx[2] - vx[0];
      a2y = vy[2] - vy[0];
      a3x = (vx[0] - vx[1]) - (vx[2] - vx[3]);
      a3y = (vy[0] - vy[1]) - (vy[2] - vy[3]);
    }
    __device__ __forceinline__ void
    jac(const double xi, const double eta, double &j00, double &j01, double &j10, double &j11) const
    {
      j00 = a1x + a3x * eta;
      j01 = a2x + a3x * xi;
      j10 = a1y + a3y * eta;
      j11 = a2y + a3y * xi;
    }
    __device__ __forceinline__ void
    point(const double xi, const double eta, double &px, double &py) const
    {
      px = x0 + a1x * xi + (a2x + a3x * xi) * eta;
      py = y0 + a1y * xi + (a2y + a3y * xi) * eta;
    }
  };

  // reference gradients of the scalar shape functions (Q1 nodal 0..3, BR bubbles 4..7) times the
  // adjugate of J:  nx = s_xi J11 - s_eta J10,  ny = -s_xi J01 + s_eta J00   (grad = n / det J)
  template <int NS>
  __device__ __forceinline__ void
  shape_grads(const double xi, const double eta, const double j00, const double j01, const double j10,
              const double j11, double nx[NS], double ny[NS])
  {
    const double mxi = 1. - xi, meta = 1. - eta;
    double       sx[NS], se[NS];
    sx[0] = -meta;
    se[0] = -mxi;
    sx[1] = meta;
    se[1] = -xi;
    sx[2] = -eta;
    se[2] = mxi;
    sx[3] = eta;
    se[3] = xi;
    if (NS == 8)
      {
        const double bx = 4. * xi * mxi, dbx = 4. * (1. - 2. * xi);
        const double by = 4. * eta * meta, dby = 4. * (1. - 2. * eta);
        sx[4] = -by;
        se[4] = mxi * dby;
        sx[5] = by;
        se[5] = xi * dby;
        sx[6] = dbx * meta;
        se[6] = -bx;
        sx[7] = dbx * eta;
        se[7] = bx;
      }
#pragma unroll
    for (int a = 0; a < NS; ++a)
      {
        nx[a] = sx[a] * j11 - se[a] * j10;
        ny[a] = se[a] * j00 - sx[a] * j01;
      }
  }

  template <int NS>
  __device__ __forceinline__ void
  shape_values(const double xi, const double eta, double s[NS])
  {
    const double mxi = 1. - xi, meta = 1. - eta;
    s[0] = mxi * meta;
    s[1] = xi * meta;
    s[2] = mxi * eta;
    s[3] = xi * eta;
    if (NS == 8)
      {
        const double bx = 4. * xi * mxi, by = 4. * eta * meta;
        s[4] = mxi * by;
        s[5] = xi * by;
        s[6] = bx * meta;
        s[7] = bx * eta;
      }
  }

  // displacement component carried by scalar shape a: vertex shapes carry both, bubbles 4,5 -> x,
  // bubbles 6,7 -> y.  Local dof index of (shape a, component c):
  template <int EL>
  __host__ __device__ constexpr int
  udof(int a, int c)
  {
    return a < 4 ? 2 * a + c : Elem<EL>::bubble(a - 4);
  }

  // inverse of a symmetric positive definite 4x4 (RT0 Gram matrix; FullMatrix::gauss_jordan at
  // BR1new:826).  m: 10 unique entries in row-major upper order 00 01 02 03 11 12 13 22 23 33.
  __device__ __forceinline__ void
  spd4_inverse(const double m[10], double g[10])
  {
    // Cholesky-free: block 2x2 inversion  M = [A B; B^T D]
    const double a00 = m[0], a01 = m[1], a11 = m[4];
    const double b00 = m[2], b01 = m[3], b10 = m[5], b11 = m[6];
    const double d00 = m[7], d01 = m[8], d11 = m[9];
    const double ida = 1. / (a00 * a11 - a01 * a01);
    const double i00 = a11 * ida, i01 = -a01 * ida, i11 = a00 * ida; // A^{-1}
    // T = A^{-1} B  (2x2)
    const double t00 = i00 * b00 + i01 * b10, t01 = i00 * b01 + i01 * b11;
    const double t10 = i01 * b00 + i11 * b10, t11 = i01 * b01 + i11 * b11;
    // S = D - B^T T (symmetric)
    const double s00 = d00 - (b00 * t00 + b10 * t10);
    const double s01 = d01 - (b00 * t01 + b10 * t11);
    const double s11 = d11 - (b01 * t01 + b11 * t11);
    const double ids = 1. / (s00 * s11 - s01 * s01);
    const double q00 = s11 * ids, q01 = -s01 * ids, q11 = s00 * ids; // S^{-1}
    // lower-right = S^{-1}; upper-right = -T S^{-1}; upper-left = A^{-1} + T S^{-1} T^T
    const double u00 = t00 * q00 + t01 * q01, u01 = t00 * q01 + t01 * q11; // T S^{-1}
    const double u10 = t10 * q00 + t11 * q01, u11 = t10 * q01 + t11 * q11;
    g[7] = q00;
    g[8] = q01;
    g[9] = q11;
    g[2] = -u00;
    g[3] = -u01;
    g[5] = -u10;
    g[6] = -u11;
    g[0] = i00 + (u00 * t00 + u01 * t01);
    g[1] = i01 + (u00 * t10 + u01 * t11);
    g[4] = i11 + (u10 * t10 + u11 * t11);
  }
  __host__ __device__ constexpr int
  sym4(int i, int j)
  {
    // index into the 10-entry upper-triangular storage
    return i <= j ? (i == 0 ? j : i == 1 ? 3 + j : i == 2 ? 5 + j : 9) : sym4(j, i);
  }

  // ---------------------------------------------------------------------------------------------
  // The cell kernel.  `old` = old_solution restricted to the cell (FESystem order).  Emit must
  // provide  void mat(int i, int j, double v)  and  void rhs(int i, double v); every (i,j) pair
  // of the NL x NL matrix is emitted exactly once (structural zeros included) with compile-time
  // constant indices after unrolling.
  //
  // Three quadrature sweeps keep the live accumulators of one thread within the register file (a
  // single sweep needs ~150 FP64 accumulators and spills):
  //   sweep X: A[(a,x),(b,x)] = mu int (2 dx s_a dx s_b + dy s_a dy s_b)  + int d/dx, d/dy s_a, |E|
  //   sweep Y: A[(a,y),(b,y)] = mu int (dx s_a dx s_b + 2 dy s_a dy s_b)  + RT0 Gram + load vector
  //   sweep Q: A[(a,y),(b,x)] = mu int dx s_a dy s_b   (and its transpose)
  // ---------------------------------------------------------------------------------------------
  template <int NS>
  struct QPoint
  {
    double r, jxw, wh, xi, eta, j00, j01, j10, j11;
    double nx[NS], ny[NS];
  };
  // QGauss<2>(3) in deal.II's tensor order (x fastest), in constant memory: the quadrature loop is
  // a real loop (9 trips) so that its body stays resident in the instruction cache -- the fully
  // unrolled version is ~160 KB of straight-line SASS and is instruction-fetch bound.
  __constant__ double kQxi[9]  = {0.11270166537925831148, 0.5, 0.88729833462074168852, 0.11270166537925831148, 0.5,
                                  0.88729833462074168852, 0.11270166537925831148, 0.5, 0.88729833462074168852};
  __constant__ double kQeta[9] = {0.11270166537925831148, 0.11270166537925831148, 0.11270166537925831148, 0.5, 0.5, 0.5,
                                  0.88729833462074168852, 0.88729833462074168852, 0.88729833462074168852};
  __constant__ double kQw[9]   = {25. / 324., 40. / 324., 25. / 324., 40. / 324., 64. / 324., 40. / 324.,
                                  25. / 324., 40. / 324., 25. / 324.};
  template <int NS>
  __device__ __forceinline__ void
  eval_qpoint(const Geom &G, const int qi, QPoint<NS> &q)
  {
    q.xi  = kQxi[qi];
    q.eta = kQeta[qi];
    q.wh  = kQw[qi];
    G.jac(q.xi, q.eta, q.j00, q.j01, q.j10, q.j11);
    const double det = q.j00 * q.j11 - q.j01 * q.j10;
    q.r              = q.wh / det;
    q.jxw            = q.wh * det;
    shape_grads<NS>(q.xi, q.eta, q.j00, q.j01, q.j10, q.j11, q.nx, q.ny);
  }
  // scalar shapes that carry an x (y) displacement component: vertices 0..3 and bubbles 4,5 (6,7)
  template <int NS>
  __host__ __device__ constexpr int
  xcar(int k)
  {
    return k; // 0..5
  }
  template <int NS>
  __host__ __device__ constexpr int
  ycar(int k)
  {
    return k < 4 ? k : k + 2; // 0..3, 6, 7
  }

  // SW selects which sweeps this instantiation runs (bit 0: X, bit 1: Q, bit 2: Y + Darcy + rhs).
  // The assembly launches one kernel per sweep: the entry sets are disjoint, the geometry is
  // recomputed (cheap) and every kernel stays within ~130 registers.  SW = 7 is the all-in-one
  // version used for boundary cells.
  template <int EL, int SW, class Emit>
  __device__ __forceinline__ void
  cell_kernel(const Geom &G, const double kperm, const DevParams &P, const double *old, Emit &E)
  {
    using EE         = Elem<EL>;
    constexpr int NS = EE::NS, PI = EE::PINT;
    constexpr int NC = NS == 8 ? 6 : 4; // carriers per component
    constexpr bool SX = (SW & 1) != 0, SQ = (SW & 2) != 0, SY = (SW & 4) != 0;

    // ---- divergence moments int d/dx s_a, int d/dy s_a and |E| (needed by every sweep) -------------
    double dx[NS], dy[NS], W = 0.;
    double axx[SX ? NC : 1][SX ? NC : 1], q_yx[SQ ? NC : 1][SQ ? NC : 1], ayy[SY ? NC : 1][SY ? NC : 1];
    double m[10], bs = 0.;
    double bfx[SY ? NS : 1], bfy[SY ? NS : 1];
#pragma unroll
    for (int a = 0; a < NS; ++a)
      dx[a] = dy[a] = 0.;
    if (SX)
      {
#pragma unroll
        for (int a = 0; a < NC; ++a)
#pragma unroll
          for (int b = 0; b < NC; ++b)
            axx[a % (SX ? NC : 1)][b % (SX ? NC : 1)] = 0.;
      }
    if (SQ)
      {
#pragma unroll
        for (int a = 0; a < NC; ++a)
#pragma unroll
          for (int b = 0; b < NC; ++b)
            q_yx[a % (SQ ? NC : 1)][b % (SQ ? NC : 1)] = 0.;
      }
    if (SY)
      {
#pragma unroll
        for (int a = 0; a < NC; ++a)
#pragma unroll
          for (int b = 0; b < NC; ++b)
            ayy[a % (SY ? NC : 1)][b % (SY ? NC : 1)] = 0.;
#pragma unroll
        for (int a = 0; a < NS; ++a)
          bfx[a % (SY ? NS : 1)] = bfy[a % (SY ? NS : 1)] = 0.;
#pragma unroll
        for (int k = 0; k < 10; ++k)
          m[k] = 0.;
      }

#pragma unroll 1
    for (int qi = 0; qi < 9; ++qi)
        {
          QPoint<NS> q;
          eval_qpoint<NS>(G, qi, q);
          W += q.jxw;
#pragma unroll
          for (int a = 0; a < NS; ++a)
            {
              dx[a] += q.wh * q.nx[a];
              dy[a] += q.wh * q.ny[a];
            }
          if (SX)
            {
#pragma unroll
              for (int ka = 0; ka < NC; ++ka)
                {
                  const int    a  = xcar<NS>(ka);
                  const double rx = 2. * q.r * q.nx[a], ry = q.r * q.ny[a];
#pragma unroll
                  for (int kb = ka; kb < NC; ++kb)
                    axx[ka % (SX ? NC : 1)][kb % (SX ? NC : 1)] += rx * q.nx[xcar<NS>(kb)] + ry * q.ny[xcar<NS>(kb)];
                }
            }
          if (SQ)
            {
#pragma unroll
              for (int ka = 0; ka < NC; ++ka)
                {
                  const double rx = q.r * q.nx[ycar<NS>(ka)];
#pragma unroll
                  for (int kb = 0; kb < NC; ++kb)
                    q_yx[ka % (SQ ? NC : 1)][kb % (SQ ? NC : 1)] += rx * q.ny[xcar<NS>(kb)];
                }
            }
          if (SY)
            {
#pragma unroll
              for (int ka = 0; ka < NC; ++ka)
                {
                  const int    a  = ycar<NS>(ka);
                  const double rx = q.r * q.nx[a], ry = 2. * q.r * q.ny[a];
#pragma unroll
                  for (int kb = ka; kb < NC; ++kb)
                    ayy[ka % (SY ? NC : 1)][kb % (SY ? NC : 1)] += rx * q.nx[ycar<NS>(kb)] + ry * q.ny[ycar<NS>(kb)];
                }
              // RT0 Gram: w_k = J w_hat_k / det;  (w_k . w_l) JxW = r (J w_hat_k).(J w_hat_l)
              {
                const double c00 = q.j00 * q.j00 + q.j10 * q.j10, c01 = q.j00 * q.j01 + q.j10 * q.j11,
                             c11 = q.j01 * q.j01 + q.j11 * q.j11;
                const double mxi = 1. - q.xi, meta = 1. - q.eta;
                const double rc00 = q.r * c00, rc01 = q.r * c01, rc11 = q.r * c11;
                m[0] += rc00 * (mxi * mxi);
                m[1] += rc00 * (mxi * q.xi);
                m[4] += rc00 * (q.xi * q.xi);
                m[2] += rc01 * (mxi * meta);
                m[3] += rc01 * (mxi * q.eta);
                m[5] += rc01 * (q.xi * meta);
                m[6] += rc01 * (q.xi * q.eta);
                m[7] += rc11 * (meta * meta);
                m[8] += rc11 * (meta * q.eta);
                m[9] += rc11 * (q.eta * q.eta);
              }
              if (P.case_id == 1)
                {
                  double px, py, sx, cx, sy, cy;
                  G.point(q.xi, q.eta, px, py);
                  sincospi(px, &sx, &cx);
                  sincospi(py, &sy, &cy);
                  const double fx = P.f_coef * sx * cy * q.jxw, fy = P.f_coef * cx * sy * q.jxw;
                  bs += cx * cy * q.jxw;
                  double s[NS];
                  shape_values<NS>(q.xi, q.eta, s);
#pragma unroll
                  for (int a = 0; a < NS; ++a)
                    {
                      if (a < 6)
                        bfx[a % (SY ? NS : 1)] += fx * s[a];
                      if (a < 4 || a >= 6)
                        bfy[a % (SY ? NS : 1)] += fy * s[a];
                    }
                }
            }
        }

    // ---- divergence vector of the u-dofs.  BR1: cell-averaged divergence (BR1new:867-876);
    // CGQ1: divergence at the 1-point rule (CGQ1:1341-1358)
    double dv[EE::NU];
    double wq, c0_weight;
    if (EL == 0)
      {
        // cell->measure(): shoelace on vertices 0,1,3,2 = det J at the cell centre    BR1new:867
        const double area = 0.25 * ((G.a1x + G.a1x + G.a3x) * (G.a2y + G.a2y + G.a3y) -
                                    (G.a2x + G.a2x + G.a3x) * (G.a1y + G.a1y + G.a3y));
        const double ia   = 1. / area;
#pragma unroll
        for (int a = 0; a < 4; ++a)
          {
            dv[2 * a]     = dx[a] * ia;
            dv[2 * a + 1] = dy[a] * ia;
          }
        dv[8 % EE::NU]  = dx[4 % NS] * ia;
        dv[9 % EE::NU]  = dx[5 % NS] * ia;
        dv[10 % EE::NU] = dy[6 % NS] * ia;
        dv[11 % EE::NU] = dy[7 % NS] * ia;
        wq              = W;
        c0_weight       = W;
      }
    else
      {
        double j00, j01, j10, j11, nx[NS], ny[NS];
        G.jac(.5, .5, j00, j01, j10, j11);
        const double detc = j00 * j11 - j01 * j10;
        shape_grads<NS>(.5, .5, j00, j01, j10, j11, nx, ny);
        const double id = 1. / detc;
#pragma unroll
        for (int a = 0; a < 4; ++a)
          {
            dv[2 * a]     = nx[a] * id;
            dv[2 * a + 1] = ny[a] * id;
          }
        wq        = detc;     // JxW of QGauss(1)
        c0_weight = W + detc; // storage term integrated twice: CGQ1:1335 and :1352
      }
    const double lw = P.lambda * wq;
    auto dvx = [&](int a) -> double { return a < 4 ? dv[2 * a] : dv[(8 + a - 4) % EE::NU]; };
    auto dvy = [&](int a) -> double { return a < 4 ? dv[2 * a + 1] : dv[(8 + a - 4) % EE::NU]; };

    if (SX)
      {
#pragma unroll
        for (int ka = 0; ka < NC; ++ka)
#pragma unroll
          for (int kb = 0; kb < NC; ++kb)
            {
              const int    a = xcar<NS>(ka), b = xcar<NS>(kb);
              const double v = kb >= ka ? axx[ka % (SX ? NC : 1)][kb % (SX ? NC : 1)] : axx[kb % (SX ? NC : 1)][ka % (SX ? NC : 1)];
              E.mat(udof<EL>(a, 0), udof<EL>(b, 0), P.mu * v + lw * dvx(a) * dvx(b));
            }
        // coupling with p_int (BR1new:915,917 / CGQ1:1351,1354), structural zeros u <-> p_face
        const double aw = P.alpha * wq;
#pragma unroll
        for (int i = 0; i < EE::NU; ++i)
          {
            const int li = (EL == 0 && i >= 8) ? EE::bubble(i - 8) : i;
            E.mat(li, PI, -aw * dv[i]);
            E.mat(PI, li, aw * dv[i]);
#pragma unroll
            for (int f = 0; f < 4; ++f)
              {
                E.mat(li, EE::pface(f), 0.);
                E.mat(EE::pface(f), li, 0.);
              }
          }
      }
    if (SQ)
      {
#pragma unroll
        for (int ka = 0; ka < NC; ++ka)
#pragma unroll
          for (int kb = 0; kb < NC; ++kb)
            {
              const int    a = ycar<NS>(ka), b = xcar<NS>(kb);
              const int    i = udof<EL>(a, 1), j = udof<EL>(b, 0);
              const double v = P.mu * q_yx[ka % (SQ ? NC : 1)][kb % (SQ ? NC : 1)] + lw * dvy(a) * dvx(b);
              E.mat(i, j, v); // (a,y),(b,x)
              E.mat(j, i, v); // the u block is symmetric
            }
      }
    if (SY)
      {
#pragma unroll
        for (int ka = 0; ka < NC; ++ka)
#pragma unroll
          for (int kb = 0; kb < NC; ++kb)
            {
              const int    a = ycar<NS>(ka), b = ycar<NS>(kb);
              const double v = kb >= ka ? ayy[ka % (SY ? NC : 1)][kb % (SY ? NC : 1)] : ayy[kb % (SY ? NC : 1)][ka % (SY ? NC : 1)];
              E.mat(udof<EL>(a, 1), udof<EL>(b, 1), P.mu * v + lw * dvy(a) * dvy(b));
            }
        // ---- Darcy block  dt k F M^{-1} F^T on (p_int, p_face 0..3)           BR1new:811-863,879-901
        {
          double g[10];
          spd4_inverse(m, g);
          const double sc    = P.dt * kperm;
          const double sg[4] = {-1., 1., -1., 1.}; // F_face = diag(sg); F_int = -sg
          double       rowsum[4];
#pragma unroll
          for (int k = 0; k < 4; ++k)
            {
              double s = 0.;
#pragma unroll
              for (int l = 0; l < 4; ++l)
                s += -sg[l] * g[sym4(k, l)];
              rowsum[k] = s; // (M^{-1} F_int^T)_k
            }
          double sii = 0.;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            sii += -sg[k] * rowsum[k];
          E.mat(PI, PI, P.c0 * c0_weight + sc * sii);
#pragma unroll
          for (int f = 0; f < 4; ++f)
            {
              const double v = sc * sg[f] * rowsum[f];
              E.mat(PI, EE::pface(f), v);
              E.mat(EE::pface(f), PI, v);
#pragma unroll
              for (int f2 = 0; f2 < 4; ++f2)
                E.mat(EE::pface(f), EE::pface(f2), sc * sg[f] * sg[f2] * g[sym4(f, f2)]);
            }
        }
        // ---- right-hand side                                                BR1new:928-961 / CGQ1:1381-1453
        double div_old = 0.; // BR1: average divergence of u_old ; CGQ1: divergence at the cell centre
#pragma unroll
        for (int i = 0; i < EE::NU; ++i)
          {
            const int li = (EL == 0 && i >= 8) ? EE::bubble(i - 8) : i;
            div_old += old[li] * dv[i];
          }
#pragma unroll
        for (int a = 0; a < NS; ++a)
          {
            if (a < 6)
              E.rhs(udof<EL>(a, 0), bfx[a % (SY ? NS : 1)]);
            if (a < 4 || a >= 6)
              E.rhs(udof<EL>(a, 1), bfy[a % (SY ? NS : 1)]);
          }
        E.rhs(PI, P.c0 * old[PI] * W + P.dt * P.s_coef * bs + P.alpha * div_old * wq);
#pragma unroll
        for (int f = 0; f < 4; ++f)
          E.rhs(EE::pface(f), 0.);
      }
  }

  // ---------------------------------------------------------------------------------------------
  // Boundary data of one cell (BR1new:963-1165, CGQ1:1563-1720): traction contribution to the local
  // rhs and the per-cell Dirichlet constraint lines (is_con / g per local dof).
  // flags: bits 3f..3f+2 = PE_FACE_* kinds of face f; bits 12..15 = vertex v carries a mesh-wide
  // displacement constraint (CGQ1's interpolate_boundary_values over the whole DoFHandler).
  // ---------------------------------------------------------------------------------------------
  __device__ __forceinline__ void
  exact_u(const DevParams &P, const double x, const double y, double &ux, double &uy)
  {
    if (P.case_id == 1)
      {
        double sx, cx, sy, cy;
        sincospi(x, &sx, &cx);
        sincospi(y, &sy, &cy);
        ux = P.u_coef * sx * cy;
        uy = P.u_coef * cx * sy;
      }
    else
      ux = uy = 0.;
  }
  __device__ __forceinline__ double
  exact_p(const DevParams &P, const double x, const double y)
  {
    if (P.case_id == 1)
      {
        double sx, cx, sy, cy;
        sincospi(x, &sx, &cx);
        sincospi(y, &sy, &cy);
        return P.p_coef * cx * cy;
      }
    return 0.;
  }

  template <int EL>
  __device__ inline void
  cell_boundary(const Geom &G, const double vx[4], const double vy[4], const uint32_t flags, const DevParams &P,
                double *b, bool *is_con, double *g)
  {
    using EE = Elem<EL>;
    const int fv[4][2] = {{0, 2}, {1, 3}, {0, 1}, {2, 3}};
    for (int i = 0; i < EE::NL; ++i)
      {
        is_con[i] = false;
        g[i]      = 0.;
      }
    for (int f = 0; f < 4; ++f)
      {
        const uint32_t kind = (flags >> (3 * f)) & 7u;
        if (kind == 0)
          continue;
        const int    v0 = fv[f][0], v1 = fv[f][1];
        const double tx = vx[v1] - vx[v0], ty = vy[v1] - vy[v0];
        const double len = sqrt(tx * tx + ty * ty);
        if (kind & 4u) // traction: b_i += int_f t_N . phi_i                   BR1new:981-988
          {
            // Q1 traces integrate to len/2 each; the bubble 4t(1-t) e_{f/2} integrates to 2/3 len
            for (int c = 0; c < 2; ++c)
              {
                b[2 * v0 + c] += P.traction[c] * (0.5 * len);
                b[2 * v1 + c] += P.traction[c] * (0.5 * len);
              }
            if (EL == 0)
              b[EE::bubble(f)] += P.traction[f / 2] * (2. / 3. * len);
          }
        if (kind & 2u) // p_face <- p(face midpoint)                           BR1new:1063-1070
          {
            const double mx = 0.5 * (vx[v0] + vx[v1]), my = 0.5 * (vy[v0] + vy[v1]);
            is_con[EE::pface(f)] = true;
            g[EE::pface(f)]      = exact_p(P, mx, my);
          }
        if (kind & 1u) // u Dirichlet                                          BR1new:1109-1159
          {
            double vv[4];
            exact_u(P, vx[v0], vy[v0], vv[0], vv[1]);
            exact_u(P, vx[v1], vy[v1], vv[2], vv[3]);
            is_con[2 * v0] = is_con[2 * v0 + 1] = is_con[2 * v1] = is_con[2 * v1 + 1] = true;
            g[2 * v0]     = vv[0];
            g[2 * v0 + 1] = vv[1];
            g[2 * v1]     = vv[2];
            g[2 * v1 + 1] = vv[3];
            if (EL == 0)
              {
                // outward normal at face quadrature point 0: sgn * row nb of J^{-1}, normalised
                const int    nb = f < 2 ? 0 : 1;
                const double sg = (f % 2 == 0) ? -1. : 1.;
                const double t0 = kGx[0];
                const double xi = f == 0 ? 0. : f == 1 ? 1. : t0, eta = f == 2 ? 0. : f == 3 ? 1. : t0;
                double       j00, j01, j10, j11;
                G.jac(xi, eta, j00, j01, j10, j11);
                // rows of adj(J): row0 = (j11, -j01), row1 = (-j10, j00); det > 0 drops out
                double nx = sg * (nb == 0 ? j11 : -j10), ny = sg * (nb == 0 ? -j01 : j00);
                const double nl = sqrt(nx * nx + ny * ny);
                nx /= nl;
                ny /= nl;
                double res_flux = 0., bubble_flux = 0.;
                for (int q = 0; q < 3; ++q)
                  {
                    const double t  = kGx[q], w = kGw[q] * len;
                    const double qx = vx[v0] + t * tx, qy = vy[v0] + t * ty;
                    double       ux, uy;
                    exact_u(P, qx, qy, ux, uy);
                    ux -= vv[0] * (1. - t) + vv[2] * t;
                    uy -= vv[1] * (1. - t) + vv[3] * t;
                    const double bub = 4. * t * (1. - t);
                    bubble_flux += (f / 2 == 0 ? nx : ny) * bub * w;
                    res_flux += (ux * nx + uy * ny) * w;
                  }
                is_con[EE::bubble(f)] = true;
                g[EE::bubble(f)]      = res_flux / bubble_flux;
              }
          }
      }
    if (EL == 1) // mesh-wide vertex constraints                              CGQ1:1673-1679
      for (int v = 0; v < 4; ++v)
        if ((flags >> (12 + v)) & 1u)
          {
            double ux, uy;
            exact_u(P, vx[v], vy[v], ux, uy);
            is_con[2 * v] = is_con[2 * v + 1] = true;
            g[2 * v]     = ux;
            g[2 * v + 1] = uy;
          }
  }
} // namespace pe
