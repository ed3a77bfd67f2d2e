// TEST INFRASTRUCTURE ONLY -- never linked into, imported by or executed from the product path.
//
// Restatements of the deal.II 9.1 library behaviour that the reference programs lean on
// (quadrature, Q1 mapping, FE shape functions, DoF numbering, sparsity, constraint elimination).
// deal.II itself is not present in /root/reference nor installed in this image, so every item in
// this file is tagged [deal.II-memory]: it is written from knowledge of deal.II 9.1.x and is kept
// behind ONE function each so that a box with a real deal.II can correct it.
//
// PARITY UNPINNED: the reference ships no golden vectors, no tests and cannot be built here
// (needs deal.II >= 9.1 + UMFPACK).  The only pins are the hand-derived known answers of
// SURVEY.md section 8(c), checked in tests/test_oracle_kat.py.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <vector>

namespace orc
{
  using dof_t = std::uint32_t; // deal.II types::global_dof_index (32 bit build)

  // ------------------------------------------------------------------------------------------
  // QGauss<1>(n) on [0,1].  [deal.II-memory] quadrature_lib.cc: Newton iteration on the Legendre
  // polynomial in long double, points ascending, x -> (1+x)/2, w -> w/2.
  // Used by: BR1new:690-691, CGQ1:1037-1041, ESM:336-337.
  // ------------------------------------------------------------------------------------------
  inline void
  qgauss_1d(const int n, std::vector<double> &x, std::vector<double> &w)
  {
    x.assign(n, 0.);
    w.assign(n, 0.);
    const int         m  = (n + 1) / 2;
    const long double pi = 3.14159265358979323846264338327950288L;
    for (int i = 1; i <= m; ++i)
      {
        long double z = std::cos(pi * (i - .25L) / (n + .5L));
        long double pp, p1, p2, p3;
        do
          {
            p1 = 1.L;
            p2 = 0.L;
            for (int j = 0; j < n; ++j)
              {
                p3 = p2;
                p2 = p1;
                p1 = ((2.L * j + 1.L) * z * p2 - j * p3) / (j + 1);
              }
            pp = n * (z * p1 - p2) / (z * z - 1.L);
            z  = z - p1 / pp;
          }
        while (std::fabs((double)(p1 / pp)) > 1e-19);
        const double xx = .5 * (double)z;
        x[i - 1]        = .5 - xx;
        x[n - i]        = .5 + xx;
        const double ww = (double)(1.L / ((1.L - z * z) * pp * pp));
        w[i - 1]        = ww;
        w[n - i]        = ww;
      }
  }

  struct Point2
  {
    double x, y;
  };

  // Tensor-product QGauss<2>(n): x runs fastest.  [deal.II-memory] Quadrature<dim>(Quadrature<1>)
  struct Quad2
  {
    std::vector<Point2> p;
    std::vector<double> w;
  };
  inline Quad2
  qgauss_2d(const int n)
  {
    std::vector<double> x, w;
    qgauss_1d(n, x, w);
    Quad2 q;
    for (int j = 0; j < n; ++j)
      for (int i = 0; i < n; ++i)
        {
          q.p.push_back({x[i], x[j]});
          q.w.push_back(w[i] * w[j]);
        }
    return q;
  }

  // ------------------------------------------------------------------------------------------
  // Reference cell conventions of GeometryInfo<2> [deal.II-memory]:
  //   vertices 0:(0,0) 1:(1,0) 2:(0,1) 3:(1,1)
  //   faces    0: x=0 (v0,v2)  1: x=1 (v1,v3)  2: y=0 (v0,v1)  3: y=1 (v2,v3)
  //   a point t on the reference face f maps to (0,t) (1,t) (t,0) (t,1)   (QProjector::project_to_face)
  // ------------------------------------------------------------------------------------------
  static const int face_vertices[4][2] = {{0, 2}, {1, 3}, {0, 1}, {2, 3}};

  inline Point2
  project_to_face(const int f, const double t)
  {
    switch (f)
      {
        case 0:
          return {0., t};
        case 1:
          return {1., t};
        case 2:
          return {t, 0.};
        default:
          return {t, 1.};
      }
  }

  // ------------------------------------------------------------------------------------------
  // MappingQ1 (bilinear) on one quadrilateral [deal.II-memory: MappingQGeneric(1)].
  // J[a][b] = d x_a / d xi_b ; covariant transform of a gradient = grad_ref * J^{-1}.
  // ------------------------------------------------------------------------------------------
  struct CellGeom
  {
    double vx[4], vy[4];
  };

  inline void
  q1_shape(const double xi, const double eta, double N[4], double dN[4][2])
  {
    N[0]     = (1. - xi) * (1. - eta);
    N[1]     = xi * (1. - eta);
    N[2]     = (1. - xi) * eta;
    N[3]     = xi * eta;
    dN[0][0] = -(1. - eta);
    dN[0][1] = -(1. - xi);
    dN[1][0] = (1. - eta);
    dN[1][1] = -xi;
    dN[2][0] = -eta;
    dN[2][1] = (1. - xi);
    dN[3][0] = eta;
    dN[3][1] = xi;
  }

  struct MapPoint
  {
    double x[2];
    double J[2][2];
    double Jinv[2][2];
    double det;
  };

  inline MapPoint
  map_at(const CellGeom &c, const double xi, const double eta)
  {
    double N[4], dN[4][2];
    q1_shape(xi, eta, N, dN);
    MapPoint m;
    m.x[0] = m.x[1] = 0;
    m.J[0][0] = m.J[0][1] = m.J[1][0] = m.J[1][1] = 0;
    for (int v = 0; v < 4; ++v)
      {
        m.x[0] += N[v] * c.vx[v];
        m.x[1] += N[v] * c.vy[v];
        for (int b = 0; b < 2; ++b)
          {
            m.J[0][b] += dN[v][b] * c.vx[v];
            m.J[1][b] += dN[v][b] * c.vy[v];
          }
      }
    m.det        = m.J[0][0] * m.J[1][1] - m.J[0][1] * m.J[1][0];
    m.Jinv[0][0] = m.J[1][1] / m.det;
    m.Jinv[0][1] = -m.J[0][1] / m.det;
    m.Jinv[1][0] = -m.J[1][0] / m.det;
    m.Jinv[1][1] = m.J[0][0] / m.det;
    return m;
  }

  // cell->measure() for a 2-D quad [deal.II-memory: GridTools::cell_measure<2>, shoelace formula on
  // the vertices taken counter-clockwise 0,1,3,2].
  inline double
  cell_measure(const CellGeom &c)
  {
    const int o[4] = {0, 1, 3, 2};
    double    a    = 0;
    for (int k = 0; k < 4; ++k)
      {
        const int i = o[k], j = o[(k + 1) % 4];
        a += c.vx[i] * c.vy[j] - c.vx[j] * c.vy[i];
      }
    return .5 * a;
  }

  // Face data of FEFaceValues: JxW = |J t_ref| w, outward unit normal n ~ J^{-T} n_ref.
  struct FacePoint
  {
    MapPoint m;
    double   JxW;
    double   n[2];
  };
  inline FacePoint
  face_point(const CellGeom &c, const int f, const double t, const double w)
  {
    const Point2 r = project_to_face(f, t);
    FacePoint    fp;
    fp.m = map_at(c, r.x, r.y);
    // reference tangent: faces 0,1 run in eta, faces 2,3 run in xi
    const int    tb   = (f < 2) ? 1 : 0;
    const double tx   = fp.m.J[0][tb], ty = fp.m.J[1][tb];
    const double len  = std::sqrt(tx * tx + ty * ty);
    fp.JxW            = len * w;
    const double sgn  = (f % 2 == 0) ? -1. : 1.;
    const int    nb   = (f < 2) ? 0 : 1; // reference normal direction
    double       nx   = sgn * fp.m.Jinv[nb][0];
    double       ny   = sgn * fp.m.Jinv[nb][1];
    const double nlen = std::sqrt(nx * nx + ny * ny);
    fp.n[0]           = nx / nlen;
    fp.n[1]           = ny / nlen;
    return fp;
  }

  // ------------------------------------------------------------------------------------------
  // Scalar reference shape functions used by the displacement elements.
  //   s = 0..3 : Q1 nodal functions                                   (FE_Q(1), FE_BernardiRaugel)
  //   s = 4..7 : Bernardi-Raugel face bubbles [deal.II-memory: PolynomialsBernardiRaugel,
  //              anisotropic indices 6,7,2,5]:
  //                f0:(1-x)4y(1-y)  f1:x 4y(1-y)  f2:4x(1-x)(1-y)  f3:4x(1-x)y
  //              direction of bubble f is the REFERENCE unit vector e_{f/2}; FE_BernardiRaugel uses
  //              mapping_none, i.e. value = reference value, gradient = grad_ref * J^{-1}.
  // ------------------------------------------------------------------------------------------
  inline void
  scalar_shapes(const double xi, const double eta, double S[8], double dS[8][2])
  {
    q1_shape(xi, eta, S, dS);
    const double bx = 4. * xi * (1. - xi), dbx = 4. * (1. - 2. * xi);
    const double by = 4. * eta * (1. - eta), dby = 4. * (1. - 2. * eta);
    S[4]     = (1. - xi) * by;
    dS[4][0] = -by;
    dS[4][1] = (1. - xi) * dby;
    S[5]     = xi * by;
    dS[5][0] = by;
    dS[5][1] = xi * dby;
    S[6]     = bx * (1. - eta);
    dS[6][0] = dbx * (1. - eta);
    dS[6][1] = -bx;
    S[7]     = bx * eta;
    dS[7][0] = dbx * eta;
    dS[7][1] = bx;
  }

  // FE_RaviartThomas(0) reference basis [deal.II-memory] and contravariant Piola map
  //   w = J w_ref / det J ,  div w = div_ref w_ref / det J.
  // The WG local matrix C M_K C^T is invariant under any change of basis of the RT space, so
  // sign/orientation conventions of deal.II cannot change assembled values.
  inline void
  rt0_ref(const double xi, const double eta, double w[4][2], double divw[4])
  {
    w[0][0] = 1. - xi;
    w[0][1] = 0;
    divw[0] = -1;
    w[1][0] = xi;
    w[1][1] = 0;
    divw[1] = 1;
    w[2][0] = 0;
    w[2][1] = 1. - eta;
    divw[2] = -1;
    w[3][0] = 0;
    w[3][1] = eta;
    divw[3] = 1;
  }

  // ------------------------------------------------------------------------------------------
  // FESystem local dof layout [deal.II-memory]: vertex dofs, then line dofs, then quad dofs; on
  // each object the base elements in order.
  //   BR1-WG (FE_BernardiRaugel(1), FE_DGQ(0), FE_FaceQ(0)), 17 dofs  (BR1new:517-519):
  //      0..7  (v0x v0y v1x v1y v2x v2y v3x v3y) ; 8,9=(bubble f0,pface f0) ; 10,11=f1 ; 12,13=f2 ;
  //      14,15=f3 ; 16=p_int
  //   CGQ1-WG (FE_Q(1)^2, FE_DGQ(0), FE_FaceQ(0)), 13 dofs            (CGQ1:808-810):
  //      0..7 vertex u ; 8..11 = pface f0..f3 ; 12 = p_int
  // ------------------------------------------------------------------------------------------
  enum Element
  {
    BR1_WG  = 0,
    CGQ1_WG = 1
  };
  enum DofKind
  {
    U_VERTEX,
    U_BUBBLE,
    P_FACE,
    P_INT
  };
  struct LocalDof
  {
    DofKind kind;
    int     obj;  // vertex or face number
    int     comp; // displacement direction for u dofs
  };

  inline std::vector<LocalDof>
  local_dofs(const Element e)
  {
    std::vector<LocalDof> d;
    for (int v = 0; v < 4; ++v)
      for (int c = 0; c < 2; ++c)
        d.push_back({U_VERTEX, v, c});
    for (int f = 0; f < 4; ++f)
      {
        if (e == BR1_WG)
          d.push_back({U_BUBBLE, f, f / 2});
        d.push_back({P_FACE, f, 0});
      }
    d.push_back({P_INT, 0, 0});
    return d;
  }

  // ------------------------------------------------------------------------------------------
  // Mesh: GridGenerator::hyper_cube(0,1) + refine_global(n)  (BR1new:583-584, CGQ1:886-887).
  // [deal.II-memory] active cells come out in Z (Morton) order: child c of a cell sits at
  // (c&1, c>>1).  Vertex *ids* are ours (lexicographic) -- they never enter a result.
  // Optional interior-vertex displacement (SURVEY 8(d) cfg 5) via a counter-based generator.
  // ------------------------------------------------------------------------------------------
  inline std::uint32_t
  morton_encode(std::uint32_t ix, std::uint32_t iy)
  {
    std::uint32_t m = 0;
    for (int b = 0; b < 16; ++b)
      m |= ((ix >> b) & 1u) << (2 * b) | ((iy >> b) & 1u) << (2 * b + 1);
    return m;
  }
  inline void
  morton_decode(std::uint32_t m, std::uint32_t &ix, std::uint32_t &iy)
  {
    ix = iy = 0;
    for (int b = 0; b < 16; ++b)
      {
        ix |= ((m >> (2 * b)) & 1u) << b;
        iy |= ((m >> (2 * b + 1)) & 1u) << b;
      }
  }

  // splitmix64 counter-based generator shared (as an algorithm) with the CUDA path: u01(seed,id,k)
  inline std::uint64_t
  splitmix64(std::uint64_t z)
  {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
  }
  inline double
  u01(std::uint64_t seed, std::uint64_t id, std::uint64_t k)
  {
    const std::uint64_t r = splitmix64(splitmix64(seed ^ (id * 0xD1B54A32D192ED03ull)) + k);
    return ((r >> 11) + 0.5) * (1.0 / 9007199254740992.0);
  }

  struct Mesh
  {
    int                        N = 0;
    std::vector<double>        vx, vy;        // (N+1)^2, lexicographic id = iy*(N+1)+ix
    std::vector<std::uint32_t> cell_vertices; // 4 per cell, Morton cell order
    std::vector<std::uint8_t>  face_bid;      // 4 per cell: boundary id, 255 = interior face
    int
    n_cells() const
    {
      return N * N;
    }
    CellGeom
    geom(const int c) const
    {
      CellGeom g;
      for (int v = 0; v < 4; ++v)
        {
          g.vx[v] = vx[cell_vertices[4 * c + v]];
          g.vy[v] = vy[cell_vertices[4 * c + v]];
        }
      return g;
    }
  };

  inline Mesh
  make_unit_square(const int n_refine, const double distort, const std::uint64_t seed)
  {
    Mesh m;
    m.N          = 1 << n_refine;
    const int N  = m.N;
    const int V  = N + 1;
    const double h = 1.0 / N;
    m.vx.resize(V * V);
    m.vy.resize(V * V);
    for (int iy = 0; iy < V; ++iy)
      for (int ix = 0; ix < V; ++ix)
        {
          double x = (double)ix / N, y = (double)iy / N; // dyadic, exact (= repeated midpoints)
          if (distort > 0 && ix > 0 && ix < N && iy > 0 && iy < N)
            {
              const std::uint64_t id = (std::uint64_t)iy * V + ix;
              x += distort * h * (2. * u01(seed, id, 0) - 1.);
              y += distort * h * (2. * u01(seed, id, 1) - 1.);
            }
          m.vx[iy * V + ix] = x;
          m.vy[iy * V + ix] = y;
        }
    m.cell_vertices.resize(4 * (size_t)N * N);
    m.face_bid.assign(4 * (size_t)N * N, 255);
    for (std::uint32_t c = 0; c < (std::uint32_t)(N * N); ++c)
      {
        std::uint32_t ix, iy;
        morton_decode(c, ix, iy);
        m.cell_vertices[4 * c + 0] = iy * V + ix;
        m.cell_vertices[4 * c + 1] = iy * V + ix + 1;
        m.cell_vertices[4 * c + 2] = (iy + 1) * V + ix;
        m.cell_vertices[4 * c + 3] = (iy + 1) * V + ix + 1;
        if (ix == 0)
          m.face_bid[4 * c + 0] = 0;
        if ((int)ix == N - 1)
          m.face_bid[4 * c + 1] = 0;
        if (iy == 0)
          m.face_bid[4 * c + 2] = 0;
        if ((int)iy == N - 1)
          m.face_bid[4 * c + 3] = 0;
      }
    return m;
  }

  // ------------------------------------------------------------------------------------------
  // DoFHandler::distribute_dofs + DoFRenumbering::component_wise
  // (BR1new:586-591, CGQ1:902-911).  [deal.II-memory] dof_handler_policy.cc
  // distribute_dofs_on_cell<2>: walk active cells in order; per cell number the not-yet-numbered
  // vertices 0..3 (all dofs of the vertex consecutively), then lines 0..3, then the quad.
  // component_wise: stable sort into blocks; BR1 dofs are non-primitive -> all in component 0;
  // CGQ1 passes block_component={0,0,1,2}.  Either way the blocks are [u | p_int | p_face].
  // ------------------------------------------------------------------------------------------
  struct DoFs
  {
    Element            element;
    int                n_loc;
    std::vector<dof_t> cell_dofs; // n_loc per cell, FESystem order, FINAL (renumbered) indices
    dof_t              n_u = 0, n_pint = 0, n_pface = 0;
    dof_t
    n() const
    {
      return n_u + n_pint + n_pface;
    }
  };

  inline DoFs
  distribute_dofs(const Mesh &mesh, const Element element)
  {
    const std::vector<LocalDof> ld    = local_dofs(element);
    const int                   n_loc = (int)ld.size();
    const int                   nc    = mesh.n_cells();
    const dof_t                 inval = 0xffffffffu;
    const int                   dpv = 2, dpl = (element == BR1_WG ? 2 : 1), dpq = 1;
    // line ids: vertical line at x-index ix, row iy: ix + (N+1)*iy ; horizontal: offset + ix + N*iy
    const int N = mesh.N, V = N + 1;
    auto      line_id = [&](std::uint32_t ix, std::uint32_t iy, int f) -> size_t {
      switch (f)
        {
          case 0:
            return (size_t)iy * V + ix;
          case 1:
            return (size_t)iy * V + ix + 1;
          case 2:
            return (size_t)V * N + (size_t)iy * N + ix;
          default:
            return (size_t)V * N + (size_t)(iy + 1) * N + ix;
        }
    };
    std::vector<dof_t> vdof((size_t)V * V, inval), ldof((size_t)2 * V * N, inval);
    std::vector<dof_t> first(nc * (size_t)n_loc);
    std::vector<int>   block; // block of every first-pass index
    dof_t              next = 0;
    for (int c = 0; c < nc; ++c)
      {
        std::uint32_t ix, iy;
        morton_decode(c, ix, iy);
        for (int v = 0; v < 4; ++v)
          {
            dof_t &d = vdof[mesh.cell_vertices[4 * c + v]];
            if (d == inval)
              {
                d = next;
                next += dpv;
                for (int k = 0; k < dpv; ++k)
                  block.push_back(0);
              }
          }
        for (int f = 0; f < 4; ++f)
          {
            dof_t &d = ldof[line_id(ix, iy, f)];
            if (d == inval)
              {
                d = next;
                next += dpl;
                if (element == BR1_WG)
                  block.push_back(0);
                block.push_back(2);
              }
          }
        const dof_t qd = next;
        next += dpq;
        block.push_back(1);
        // local -> first-pass index
        for (int i = 0; i < n_loc; ++i)
          {
            dof_t g = 0;
            switch (ld[i].kind)
              {
                case U_VERTEX:
                  g = vdof[mesh.cell_vertices[4 * c + ld[i].obj]] + ld[i].comp;
                  break;
                case U_BUBBLE:
                  g = ldof[line_id(ix, iy, ld[i].obj)];
                  break;
                case P_FACE:
                  g = ldof[line_id(ix, iy, ld[i].obj)] + (element == BR1_WG ? 1 : 0);
                  break;
                case P_INT:
                  g = qd;
                  break;
              }
            first[(size_t)c * n_loc + i] = g;
          }
      }
    // component_wise: stable by old index inside each block
    DoFs d;
    d.element = element;
    d.n_loc   = n_loc;
    std::vector<dof_t> cnt(3, 0);
    for (int b : block)
      ++cnt[b];
    d.n_u     = cnt[0];
    d.n_pint  = cnt[1];
    d.n_pface = cnt[2];
    std::vector<dof_t> start = {0, cnt[0], cnt[0] + cnt[1]};
    std::vector<dof_t> renum(next);
    for (dof_t i = 0; i < next; ++i)
      renum[i] = start[block[i]]++;
    d.cell_dofs.resize(first.size());
    for (size_t k = 0; k < first.size(); ++k)
      d.cell_dofs[k] = renum[first[k]];
    return d;
  }

  // ------------------------------------------------------------------------------------------
  // DoFTools::make_sparsity_pattern(dof_handler, dsp, constraints(empty), false) + copy_from
  // (BR1new:612-621, CGQ1:934-943): every pair of dofs sharing a cell.  Canonical form used for
  // all comparisons: one global CSR, columns ascending.  [deal.II-memory] deal.II's own
  // SparsityPattern stores the diagonal first in each row of the square diagonal blocks;
  // to_dealii_block_order() below converts positions.
  // ------------------------------------------------------------------------------------------
  struct Pattern
  {
    std::vector<std::int64_t> rowptr;
    std::vector<std::int32_t> colidx;
  };

  inline Pattern
  make_sparsity_pattern(const DoFs &d, const int n_cells)
  {
    const dof_t                     n = d.n();
    std::vector<std::vector<dof_t>> rows(n);
    for (int c = 0; c < n_cells; ++c)
      {
        const dof_t *cd = &d.cell_dofs[(size_t)c * d.n_loc];
        for (int i = 0; i < d.n_loc; ++i)
          for (int j = 0; j < d.n_loc; ++j)
            rows[cd[i]].push_back(cd[j]);
      }
    Pattern p;
    p.rowptr.assign(n + 1, 0);
    for (dof_t r = 0; r < n; ++r)
      {
        std::sort(rows[r].begin(), rows[r].end());
        rows[r].erase(std::unique(rows[r].begin(), rows[r].end()), rows[r].end());
        p.rowptr[r + 1] = p.rowptr[r] + (std::int64_t)rows[r].size();
      }
    p.colidx.resize(p.rowptr[n]);
    for (dof_t r = 0; r < n; ++r)
      std::copy(rows[r].begin(), rows[r].end(), p.colidx.begin() + p.rowptr[r]);
    return p;
  }

  // ------------------------------------------------------------------------------------------
  // FullMatrix::gauss_jordan (BR1new:826) -- in-place inverse with partial pivoting
  // [deal.II-memory: full_matrix.templates.h; LAPACK getrf/getri when configured -- equal to
  // rounding].  FullMatrix::mmult (BR1new:863): C = A * B.
  // ------------------------------------------------------------------------------------------
  inline void
  gauss_jordan(std::vector<double> &a, const int n)
  {
    // augmented [A | I] elimination with partial (row) pivoting
    std::vector<double> w((size_t)n * 2 * n, 0.);
    for (int i = 0; i < n; ++i)
      {
        for (int j = 0; j < n; ++j)
          w[(size_t)i * 2 * n + j] = a[i * n + j];
        w[(size_t)i * 2 * n + n + i] = 1.;
      }
    for (int j = 0; j < n; ++j)
      {
        int    r   = j;
        double max = std::fabs(w[(size_t)j * 2 * n + j]);
        for (int i = j + 1; i < n; ++i)
          if (std::fabs(w[(size_t)i * 2 * n + j]) > max)
            {
              max = std::fabs(w[(size_t)i * 2 * n + j]);
              r   = i;
            }
        if (r != j)
          for (int k = 0; k < 2 * n; ++k)
            std::swap(w[(size_t)j * 2 * n + k], w[(size_t)r * 2 * n + k]);
        const double piv = 1. / w[(size_t)j * 2 * n + j];
        for (int k = 0; k < 2 * n; ++k)
          w[(size_t)j * 2 * n + k] *= piv;
        for (int i = 0; i < n; ++i)
          {
            if (i == j)
              continue;
            const double f = w[(size_t)i * 2 * n + j];
            if (f == 0.)
              continue;
            for (int k = 0; k < 2 * n; ++k)
              w[(size_t)i * 2 * n + k] -= f * w[(size_t)j * 2 * n + k];
          }
      }
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j)
        a[i * n + j] = w[(size_t)i * 2 * n + n + j];
  }

  // ------------------------------------------------------------------------------------------
  // AffineConstraints (only what the reference uses: add_line + set_inhomogeneity, no entries)
  // and distribute_local_to_global(A, b, idx, Mat, rhs, use_inhomogeneities_for_rhs = true)
  // (BR1new:1062-1170, CGQ1:1662-1725).  [deal.II-memory] affine_constraints.templates.h:
  //   * unconstrained row r: Mat(r,c) += A(r,c) for unconstrained c;
  //                          rhs(r) += b(r) - sum_{c constrained, g_c != 0} A(r,c) g_c
  //   * constrained rows and columns are dropped;
  //   * set_matrix_diagonals: avg = mean_i |A(i,i)| over ALL local dofs; for every constrained
  //     local dof c: d = |A(c,c)| != 0 ? |A(c,c)| : avg ; Mat(c,c) += d ; rhs(c) += d g_c.
  // ------------------------------------------------------------------------------------------
  struct AffineConstraints
  {
    std::map<dof_t, double> lines;
    void
    clear()
    {
      lines.clear();
    }
    void
    add_line(const dof_t i)
    {
      lines.insert({i, 0.});
    }
    void
    set_inhomogeneity(const dof_t i, const double v)
    {
      lines[i] = v;
    }
    bool
    is_constrained(const dof_t i) const
    {
      return lines.count(i) != 0;
    }
  };

  struct CsrMatrix
  {
    const Pattern      *p = nullptr;
    std::vector<double> val;
    void
    add(const dof_t r, const dof_t c, const double v)
    {
      const std::int32_t *b = &p->colidx[p->rowptr[r]], *e = &p->colidx[p->rowptr[r + 1]];
      const std::int32_t *it = std::lower_bound(b, e, (std::int32_t)c);
      if (it == e || *it != (std::int32_t)c)
        {
          std::fprintf(stderr, "oracle: entry (%u,%u) not in sparsity pattern\n", r, c);
          std::abort();
        }
      val[it - &p->colidx[0]] += v;
    }
  };

  inline void
  distribute_local_to_global(const AffineConstraints   &con,
                             const std::vector<double> &A,
                             const std::vector<double> &b,
                             const dof_t               *idx,
                             const int                  n,
                             CsrMatrix                 &mat,
                             std::vector<double>       &rhs)
  {
    std::vector<char>   constrained(n);
    std::vector<double> g(n, 0.);
    int                 n_con = 0;
    for (int i = 0; i < n; ++i)
      {
        auto it        = con.lines.find(idx[i]);
        constrained[i] = it != con.lines.end();
        if (constrained[i])
          {
            g[i] = it->second;
            ++n_con;
          }
      }
    for (int i = 0; i < n; ++i)
      {
        if (constrained[i])
          continue;
        for (int j = 0; j < n; ++j)
          if (!constrained[j])
            mat.add(idx[i], idx[j], A[i * n + j]);
        double val = b[i];
        for (int j = 0; j < n; ++j)
          if (constrained[j] && g[j] != 0.)
            val -= A[i * n + j] * g[j];
        if (val != 0.)
          rhs[idx[i]] += val;
      }
    if (n_con > 0)
      {
        double avg = 0;
        for (int i = 0; i < n; ++i)
          avg += std::fabs(A[i * n + i]);
        avg /= n;
        for (int i = 0; i < n; ++i)
          if (constrained[i])
            {
              const double d = (std::fabs(A[i * n + i]) != 0) ? std::fabs(A[i * n + i]) : avg;
              mat.add(idx[i], idx[i], d);
              rhs[idx[i]] += d * g[i];
            }
      }
  }
} // namespace orc
