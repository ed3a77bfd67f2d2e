// Host-side setup: mesh, DoF numbering, sparsity pattern, partition.  See pe_host.h.
// deal.II behaviour reproduced here (all tagged [deal.II-memory] there):
//   * refine_global: active cells in Z order, child c at (c&1, c>>1)
//   * distribute_dofs: walk active cells; per cell number not-yet-numbered vertices 0..3, then
//     lines 0..3, then the cell interior; component_wise: stable sort into [u | p_int | p_face]
//   * make_sparsity_pattern: all pairs of dofs sharing a cell
#include "pe_host.h"

#include <algorithm>
#include <cmath>
#include <stdexcept>
#include <thread>
#include <unordered_map>

namespace pe
{
  uint64_t
  splitmix64(uint64_t z)
  {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
  }
  double
  u01(uint64_t seed, uint64_t id, uint64_t k)
  {
    const uint64_t r = splitmix64(splitmix64(seed ^ (id * 0xD1B54A32D192ED03ull)) + k);
    return ((double)(r >> 11) + 0.5) * (1.0 / 9007199254740992.0);
  }

  int
  element_n_loc(const int element)
  {
    switch (element)
      {
        case 0:
          return 17;
        case 1:
          return 13;
        case 2:
          return 21;
        default:
          return 0;
      }
  }

  // compact the even bits of m into the low half (inverse of bit interleaving)
  static inline uint32_t
  compact_even_bits(uint64_t m)
  {
    m &= 0x5555555555555555ull;
    m = (m | (m >> 1)) & 0x3333333333333333ull;
    m = (m | (m >> 2)) & 0x0f0f0f0f0f0f0f0full;
    m = (m | (m >> 4)) & 0x00ff00ff00ff00ffull;
    m = (m | (m >> 8)) & 0x0000ffff0000ffffull;
    m = (m | (m >> 16)) & 0x00000000ffffffffull;
    return (uint32_t)m;
  }

  void
  make_unit_square(HostMesh &m, const int n_refine, const double distort, const uint64_t seed,
                   const int boundary_kinds)
  {
    if (n_refine < 0 || n_refine > 14)
      throw std::invalid_argument("n_refine must be in [0,14]");
    const int64_t N = int64_t(1) << n_refine, V = N + 1;
    m.n_refine      = n_refine;
    m.n_vertices    = V * V;
    m.n_cells       = N * N;
    m.x.resize(m.n_vertices);
    m.y.resize(m.n_vertices);
    const double h = 1.0 / (double)N;
    for (int64_t iy = 0; iy < V; ++iy)
      for (int64_t ix = 0; ix < V; ++ix)
        {
          // dyadic coordinates: exactly what repeated midpoint refinement produces
          double px = (double)ix / (double)N, py = (double)iy / (double)N;
          if (distort > 0. && ix > 0 && ix < N && iy > 0 && iy < N)
            {
              const uint64_t id = (uint64_t)(iy * V + ix);
              px += distort * h * (2. * u01(seed, id, 0) - 1.);
              py += distort * h * (2. * u01(seed, id, 1) - 1.);
            }
          m.x[iy * V + ix] = px;
          m.y[iy * V + ix] = py;
        }
    m.cell_vertices.resize(4 * (size_t)m.n_cells);
    m.face_kinds.assign(4 * (size_t)m.n_cells, 0);
    const uint8_t bk = (uint8_t)boundary_kinds;
    for (int64_t c = 0; c < m.n_cells; ++c)
      {
        const int64_t ix = compact_even_bits((uint64_t)c), iy = compact_even_bits((uint64_t)c >> 1);
        uint32_t     *cv = &m.cell_vertices[4 * (size_t)c];
        cv[0]            = (uint32_t)(iy * V + ix);
        cv[1]            = cv[0] + 1;
        cv[2]            = (uint32_t)((iy + 1) * V + ix);
        cv[3]            = cv[2] + 1;
        uint8_t *fk      = &m.face_kinds[4 * (size_t)c];
        if (ix == 0)
          fk[0] = bk;
        if (ix == N - 1)
          fk[1] = bk;
        if (iy == 0)
          fk[2] = bk;
        if (iy == N - 1)
          fk[3] = bk;
      }
  }

  namespace
  {
    // edge (line) lookup keyed by the sorted vertex pair: a few inline slots per lower vertex,
    // overflow into a hash map (valence > 4 from the lower end only on unstructured meshes)
    struct EdgeTable
    {
      static constexpr int  kSlots = 4;
      std::vector<uint32_t> other, value;
      std::unordered_map<uint64_t, uint32_t> overflow;
      explicit EdgeTable(const int64_t n_vertices)
        : other((size_t)n_vertices * kSlots, 0xffffffffu)
        , value((size_t)n_vertices * kSlots, 0xffffffffu)
      {}
      // returns a reference to the stored value for edge (a,b); inserted = true if it was absent
      uint32_t &
      find_or_insert(uint32_t a, uint32_t b, bool &inserted)
      {
        if (a > b)
          std::swap(a, b);
        uint32_t *o = &other[(size_t)a * kSlots], *v = &value[(size_t)a * kSlots];
        for (int s = 0; s < kSlots; ++s)
          {
            if (o[s] == b)
              {
                inserted = false;
                return v[s];
              }
            if (o[s] == 0xffffffffu)
              {
                o[s]     = b;
                inserted = true;
                return v[s];
              }
          }
        const uint64_t key = ((uint64_t)a << 32) | b;
        auto           it  = overflow.find(key);
        inserted           = it == overflow.end();
        if (inserted)
          it = overflow.emplace(key, 0xffffffffu).first;
        return it->second;
      }
    };
  } // namespace

  void
  distribute_dofs(const HostMesh &m, const int element, HostDofs &d, Partition &part)
  {
    if (element < 0 || element > 2)
      throw std::invalid_argument("distribute_dofs: element not supported");
    const int n_loc = element_n_loc(element);
    d.element       = element;
    d.n_loc         = n_loc;
    // dofs per line: (bubble, p_face) | (p_face) | three FE_FaceQ(2) dofs.  WG_Q2RT2 (EltStiffMat.cc:297-298)
    // has no vertex dofs, nine FE_DGQ(2) dofs per quad and is NOT renumbered: all its dofs are kept in
    // block 0 so that the stable block sort below is the identity.
    const int      dpl   = (element == 0) ? 2 : (element == 1 ? 1 : 3);
    const int64_t  nc    = m.n_cells;
    const uint32_t inval = 0xffffffffu;

    // ---- pass 1: deal.II first-touch numbering ("first-pass index"), block of every index
    std::vector<uint32_t> vdof((size_t)m.n_vertices, inval);
    EdgeTable             edges(m.n_vertices);
    std::vector<uint8_t>  block_of;
    block_of.reserve((size_t)m.n_vertices * 2 + (size_t)nc * (2 * dpl + 9) + 16);
    d.cell_dofs.resize((size_t)nc * n_loc);
    // first-pass counter at the partition boundaries (cells nc*r/world)
    std::vector<int64_t> fp_at((size_t)part.world + 1, 0);
    int                  nb = 0;
    uint32_t             next = 0;
    for (int64_t c = 0; c < nc; ++c)
      {
        while (nb <= part.world && c == nc * nb / part.world)
          fp_at[nb++] = next;
        const uint32_t *cv = &m.cell_vertices[4 * (size_t)c];
        uint32_t       *cd = &d.cell_dofs[(size_t)c * n_loc];
        for (int v = 0; v < 4 && element != 2; ++v)
          {
            uint32_t &vd = vdof[cv[v]];
            if (vd == inval)
              {
                vd = next;
                next += 2;
                block_of.push_back(0);
                block_of.push_back(0);
              }
            cd[2 * v]     = vd;
            cd[2 * v + 1] = vd + 1;
          }
        for (int f = 0; f < 4; ++f)
          {
            bool      ins;
            uint32_t &e = edges.find_or_insert(cv[kFaceVertex[f][0]], cv[kFaceVertex[f][1]], ins);
            if (ins)
              {
                e = next;
                next += dpl;
                if (element == 2)
                  for (int k = 0; k < 3; ++k)
                    block_of.push_back(0);
                else
                  {
                    if (element == 0)
                      block_of.push_back(0); // bubble: FE_BernardiRaugel is non-primitive -> component 0
                    block_of.push_back(2);   // FE_FaceQ(0)
                  }
              }
            if (element == 0)
              {
                cd[8 + 2 * f]     = e;
                cd[8 + 2 * f + 1] = e + 1;
              }
            else if (element == 1)
              cd[8 + f] = e;
            else
              for (int k = 0; k < 3; ++k)
                cd[3 * f + k] = e + k;
          }
        if (element == 2)
          {
            for (int k = 0; k < 9; ++k)
              {
                cd[12 + k] = next++;
                block_of.push_back(0);
              }
          }
        else
          {
            cd[n_loc - 1] = next++;
            block_of.push_back(1); // FE_DGQ(0)
          }
      }
    while (nb <= part.world)
      fp_at[nb++] = next;

    // ---- pass 2: DoFRenumbering::component_wise = stable sort by block
    int64_t cnt[3] = {0, 0, 0};
    for (uint32_t i = 0; i < next; ++i)
      ++cnt[block_of[i]];
    d.n_u                 = cnt[0];
    d.n_pint              = cnt[1];
    d.n_pface             = cnt[2];
    const int64_t start[3] = {0, cnt[0], cnt[0] + cnt[1]};
    int64_t       run[3]  = {start[0], start[1], start[2]};
    std::vector<uint32_t> renum(next);
    part.all_row_begin.assign((size_t)part.world * 3, 0);
    part.all_row_end.assign((size_t)part.world * 3, 0);
    nb = 0;
    for (uint32_t i = 0; i <= next; ++i)
      {
        while (nb <= part.world && (int64_t)i == fp_at[nb])
          {
            for (int b = 0; b < 3; ++b)
              {
                if (nb < part.world)
                  part.all_row_begin[(size_t)nb * 3 + b] = run[b];
                if (nb > 0)
                  part.all_row_end[(size_t)(nb - 1) * 3 + b] = run[b];
              }
            ++nb;
          }
        if (i < next)
          renum[i] = (uint32_t)run[block_of[i]]++;
      }
    for (size_t k = 0; k < d.cell_dofs.size(); ++k)
      d.cell_dofs[k] = renum[d.cell_dofs[k]];

    part.cell_begin = nc * part.rank / part.world;
    part.cell_end   = nc * (part.rank + 1) / part.world;
    part.n_owned    = 0;
    for (int b = 0; b < 3; ++b)
      {
        part.row_begin[b] = part.all_row_begin[(size_t)part.rank * 3 + b];
        part.row_end[b]   = part.all_row_end[(size_t)part.rank * 3 + b];
        part.local_off[b] = part.n_owned;
        part.n_owned += part.row_end[b] - part.row_begin[b];
      }
  }

  // ---------------------------------------------------------------------------------------------
  // Hexahedra: DoFHandler::distribute_dofs + DoFRenumbering::component_wise for FESystem(FE_Q(1)^3, FE_DGQ(0),
  // FE_FaceQ(0)) (CGQ1:902-919 with dim = 3).  Per cell, in active-cell order: vertices 0..7 not numbered yet get three
  // consecutive displacement dofs, then the six quads one FE_FaceQ(0) dof each (a quad is identified by its four
  // vertex ids, so any conforming hexahedral mesh works), then the hex its FE_DGQ(0) dof; stable block sort.
  // ---------------------------------------------------------------------------------------------
  void
  distribute_dofs3(const int64_t n_vertices, const int64_t n_cells, const uint32_t *cv8, HostDofs &d, Partition &part)
  {
    if (part.world != 1)
      throw std::invalid_argument("distribute_dofs3: one rank only");
    constexpr int  n_loc = 31;
    const uint32_t inval = 0xffffffffu;
    static const int face_v[6][4] = {{0, 2, 4, 6}, {1, 3, 5, 7}, {0, 1, 4, 5}, {2, 3, 6, 7}, {0, 1, 2, 3}, {4, 5, 6, 7}};
    d.element = 1;
    d.n_loc   = n_loc;
    d.cell_dofs.resize((size_t)n_cells * n_loc);
    std::vector<uint32_t> vdof((size_t)n_vertices, inval);
    std::vector<uint8_t>  block_of;
    block_of.reserve((size_t)n_vertices * 3 + (size_t)n_cells * 4);
    // quad table: key = the three smallest vertex ids of the face (unique for conforming hexahedra) packed into 96
    // bits -> hashed open addressing over (a, b, c)
    struct Key
    {
      uint32_t a, b, c, dof;
    };
    size_t cap = 16;
    while (cap < (size_t)n_cells * 8)
      cap <<= 1;
    std::vector<Key> table(cap, Key{inval, inval, inval, inval});
    auto find_or_insert = [&](uint32_t v[4], bool &ins) -> uint32_t & {
      std::sort(v, v + 4);
      uint64_t hsh = splitmix64(((uint64_t)v[0] << 32) ^ v[1]) ^ splitmix64(((uint64_t)v[2] << 32) ^ v[3]);
      size_t   i   = (size_t)hsh & (cap - 1);
      while (true)
        {
          Key &k = table[i];
          if (k.a == inval)
            {
              k   = Key{v[0], v[1], v[2], inval};
              ins = true;
              return k.dof;
            }
          if (k.a == v[0] && k.b == v[1] && k.c == v[2])
            {
              ins = false;
              return k.dof;
            }
          i = (i + 1) & (cap - 1);
        }
    };
    uint32_t next = 0;
    for (int64_t c = 0; c < n_cells; ++c)
      {
        const uint32_t *cv = cv8 + 8 * (size_t)c;
        uint32_t       *cd = &d.cell_dofs[(size_t)c * n_loc];
        for (int v = 0; v < 8; ++v)
          {
            uint32_t &vd = vdof[cv[v]];
            if (vd == inval)
              {
                vd = next;
                next += 3;
                block_of.insert(block_of.end(), 3, (uint8_t)0);
              }
            cd[3 * v] = vd, cd[3 * v + 1] = vd + 1, cd[3 * v + 2] = vd + 2;
          }
        for (int f = 0; f < 6; ++f)
          {
            uint32_t  fv[4] = {cv[face_v[f][0]], cv[face_v[f][1]], cv[face_v[f][2]], cv[face_v[f][3]]};
            bool      ins;
            uint32_t &e = find_or_insert(fv, ins);
            if (ins)
              {
                e = next++;
                block_of.push_back(2);
              }
            cd[24 + f] = e;
          }
        cd[30] = next++;
        block_of.push_back(1);
      }
    int64_t cnt[3] = {0, 0, 0};
    for (uint32_t i = 0; i < next; ++i)
      ++cnt[block_of[i]];
    d.n_u     = cnt[0];
    d.n_pint  = cnt[1];
    d.n_pface = cnt[2];
    int64_t               run[3] = {0, cnt[0], cnt[0] + cnt[1]};
    std::vector<uint32_t> renum(next);
    for (uint32_t i = 0; i < next; ++i)
      renum[i] = (uint32_t)run[block_of[i]]++;
    for (auto &g : d.cell_dofs)
      g = renum[g];
    part.cell_begin = 0;
    part.cell_end   = n_cells;
    part.n_owned    = d.n();
    const int64_t st[4] = {0, d.n_u, d.n_u + d.n_pint, d.n()};
    for (int b = 0; b < 3; ++b)
      {
        part.row_begin[b] = st[b];
        part.row_end[b]   = st[b + 1];
        part.local_off[b] = st[b];
      }
    part.all_row_begin.assign(part.row_begin, part.row_begin + 3);
    part.all_row_end.assign(part.row_end, part.row_end + 3);
  }

  // ---------------------------------------------------------------------------------------------
  // Sparsity: for every owned row, the sorted union of the dofs of all cells containing the row.
  // ---------------------------------------------------------------------------------------------
  namespace
  {
    template <class F>
    void
    parallel_for(const int64_t n, F f)
    {
      unsigned nt = std::thread::hardware_concurrency();
      nt          = std::max(1u, std::min(nt, 32u));
      if (n < 4096)
        nt = 1;
      std::vector<std::thread> th;
      for (unsigned t = 0; t < nt; ++t)
        th.emplace_back([=]() { f(n * t / nt, n * (t + 1) / nt); });
      for (auto &x : th)
        x.join();
    }
  } // namespace

  void
  build_row_adjacency(const HostMesh &m, const HostDofs &d, Partition &part)
  {
    const int     n_loc = d.n_loc;
    const int64_t nc = m.n_cells, L = part.n_owned;
    std::vector<int64_t> &adj_ptr = part.adj_ptr, &adj = part.adj;
    adj_ptr.assign((size_t)L + 1, 0);
    for (int64_t c = part.cell_begin; c < nc; ++c) // owners are first touchers: no cell < cell_begin
      for (int i = 0; i < n_loc; ++i)
        {
          const int64_t l = part.global_to_owned(d.cell_dofs[(size_t)c * n_loc + i]);
          if (l >= 0)
            ++adj_ptr[l + 1];
        }
    for (int64_t l = 0; l < L; ++l)
      adj_ptr[l + 1] += adj_ptr[l];
    adj.resize((size_t)adj_ptr[L]);
    {
      std::vector<int64_t> fill(adj_ptr.begin(), adj_ptr.end() - 1);
      for (int64_t c = part.cell_begin; c < nc; ++c)
        for (int i = 0; i < n_loc; ++i)
          {
            const int64_t l = part.global_to_owned(d.cell_dofs[(size_t)c * n_loc + i]);
            if (l >= 0)
              adj[fill[l]++] = c;
          }
    }
    // ghost cells: cells >= cell_end adjacent to an owned row
    std::vector<int64_t> g;
    for (int64_t k = 0; k < (int64_t)adj.size(); ++k)
      if (adj[k] >= part.cell_end)
        g.push_back(adj[k]);
    std::sort(g.begin(), g.end());
    g.erase(std::unique(g.begin(), g.end()), g.end());
    part.asm_cells.resize((size_t)(part.cell_end - part.cell_begin));
    for (int64_t c = part.cell_begin; c < part.cell_end; ++c)
      part.asm_cells[c - part.cell_begin] = c;
    part.n_ghost_cells = (int64_t)g.size();
    part.asm_cells.insert(part.asm_cells.end(), g.begin(), g.end());
  }

  void
  color_asm_cells(const HostMesh &m, const HostDofs &d, Partition &part)
  {
    const int64_t    na = (int64_t)part.asm_cells.size();
    std::vector<int> color((size_t)na, 0);
    int              n_colors = 1;
    if (part.scatter_mode == 0)
      {
        // atomics: one "colour", cells stay in deal.II (Morton) order
      }
    else if (m.n_refine >= 0)
      {
        // refine_global meshes: the two low bits of the Morton index are (ix & 1, iy & 1)
        n_colors = 4;
        for (int64_t a = 0; a < na; ++a)
          color[a] = (int)(part.asm_cells[a] & 3);
      }
    else
      {
        // greedy colouring of the "shares an owned row" graph
        std::vector<int32_t> pos((size_t)m.n_cells, -1);
        for (int64_t a = 0; a < na; ++a)
          pos[part.asm_cells[a]] = (int32_t)a;
        std::fill(color.begin(), color.end(), -1);
        for (int64_t a = 0; a < na; ++a)
          {
            uint64_t       used = 0;
            const int64_t  c    = part.asm_cells[a];
            for (int i = 0; i < d.n_loc; ++i)
              {
                const int64_t l = part.global_to_owned(d.cell_dofs[(size_t)c * d.n_loc + i]);
                if (l < 0)
                  continue;
                for (int64_t k = part.adj_ptr[l]; k < part.adj_ptr[l + 1]; ++k)
                  {
                    const int32_t o = pos[part.adj[k]];
                    if (o >= 0 && color[o] >= 0)
                      used |= uint64_t(1) << color[o];
                  }
              }
            int col = 0;
            while ((used >> col) & 1)
              ++col;
            color[a] = col;
            n_colors = std::max(n_colors, col + 1);
          }
      }
    std::vector<int64_t> order((size_t)na);
    for (int64_t a = 0; a < na; ++a)
      order[a] = a;
    std::stable_sort(order.begin(), order.end(), [&](int64_t x, int64_t y) { return color[x] < color[y]; });
    std::vector<int64_t> sorted((size_t)na);
    part.n_colors = n_colors;
    part.color_begin.assign((size_t)n_colors + 1, 0);
    for (int64_t k = 0; k < na; ++k)
      {
        sorted[k] = part.asm_cells[order[k]];
        ++part.color_begin[(size_t)color[order[k]] + 1];
      }
    for (int c = 0; c < n_colors; ++c)
      part.color_begin[c + 1] += part.color_begin[c];
    part.asm_cells.swap(sorted);
  }

  void
  make_sparsity_pattern(const HostMesh &m, const HostDofs &d, Partition &part, HostPattern &p)
  {
    const int     n_loc = d.n_loc;
    const int64_t L = part.n_owned;
    build_row_adjacency(m, d, part);
    color_asm_cells(m, d, part);
    const std::vector<int64_t> &adj_ptr = part.adj_ptr, &adj = part.adj;
    // row lengths, then fill
    p.rowptr.assign((size_t)L + 1, 0);
    auto row_union = [&](const int64_t l, uint32_t *buf) -> int {
      int n = 0;
      for (int64_t k = adj_ptr[l]; k < adj_ptr[l + 1]; ++k)
        {
          const uint32_t *cd = &d.cell_dofs[(size_t)adj[k] * n_loc];
          for (int j = 0; j < n_loc; ++j)
            buf[n++] = cd[j];
        }
      std::sort(buf, buf + n);
      return (int)(std::unique(buf, buf + n) - buf);
    };
    int64_t max_adj = 0;
    for (int64_t l = 0; l < L; ++l)
      max_adj = std::max(max_adj, adj_ptr[l + 1] - adj_ptr[l]);
    parallel_for(L, [&](int64_t a, int64_t b) {
      std::vector<uint32_t> buf((size_t)max_adj * n_loc + 1);
      for (int64_t l = a; l < b; ++l)
        p.rowptr[l + 1] = row_union(l, buf.data());
    });
    for (int64_t l = 0; l < L; ++l)
      p.rowptr[l + 1] += p.rowptr[l];
    p.colidx.resize((size_t)p.rowptr[L]);
    parallel_for(L, [&](int64_t a, int64_t b) {
      std::vector<uint32_t> buf((size_t)max_adj * n_loc + 1);
      for (int64_t l = a; l < b; ++l)
        {
          const int n = row_union(l, buf.data());
          int32_t  *o = &p.colidx[(size_t)p.rowptr[l]];
          for (int k = 0; k < n; ++k)
            o[k] = (int32_t)buf[k];
        }
    });
    finish_partition(d, part, p);
  }

  void
  finish_partition(const HostDofs &d, Partition &part, const HostPattern &p)
  {
    (void)d;
    part.ghost_cols.clear();
    if (part.world == 1)
      return;
    std::vector<uint32_t> g;
    for (const int32_t c : p.colidx)
      if (part.global_to_owned(c) < 0)
        g.push_back((uint32_t)c);
    std::sort(g.begin(), g.end());
    g.erase(std::unique(g.begin(), g.end()), g.end());
    part.ghost_cols.swap(g);
  }

  // deal.II stores a BlockSparseMatrix block by block; inside each block a row's entries are the
  // columns ascending, except that square (diagonal) blocks keep the diagonal entry first.
  // perm[k_dealii] = k_csr with k_dealii running over blocks (0,0),(0,1),(1,0),(1,1), rows, entries.
  void
  dealii_block_order(const HostDofs &d, const HostPattern &p, std::vector<int64_t> &perm)
  {
    const int64_t n = d.n(), n_u = d.n_u;
    perm.clear();
    perm.reserve(p.colidx.size());
    for (int br = 0; br < 2; ++br)
      for (int bc = 0; bc < 2; ++bc)
        {
          const int64_t r0 = br ? n_u : 0, r1 = br ? n : n_u, c0 = bc ? n_u : 0, c1 = bc ? n : n_u;
          for (int64_t r = r0; r < r1; ++r)
            {
              const int64_t b = p.rowptr[r], e = p.rowptr[r + 1];
              const int64_t lo =
                std::lower_bound(p.colidx.begin() + b, p.colidx.begin() + e, (int32_t)c0) - p.colidx.begin();
              const int64_t hi =
                std::lower_bound(p.colidx.begin() + b, p.colidx.begin() + e, (int32_t)c1) - p.colidx.begin();
              int64_t diag = -1;
              if (br == bc)
                for (int64_t k = lo; k < hi; ++k)
                  if (p.colidx[k] == r)
                    diag = k;
              if (diag >= 0)
                perm.push_back(diag);
              for (int64_t k = lo; k < hi; ++k)
                if (k != diag)
                  perm.push_back(k);
            }
        }
  }
} // namespace pe
