// Host-side setup layer of poroelas_b200 (no CUDA in this header): mesh, deal.II-compatible DoF
// numbering, sparsity pattern, partition.  Replaces make_grid_and_dofs / setup_system
// (PoroElasBR1WG_new.cc:580-681, PoroElasCGQ1_WG.cc:883-1027, EltStiffMat.cc:242-326).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace pe
{
  // GeometryInfo<2> conventions of deal.II: vertex v of face f
  static const int kFaceVertex[4][2] = {{0, 2}, {1, 3}, {0, 1}, {2, 3}};

  struct HostMesh
  {
    int64_t               n_vertices = 0, n_cells = 0;
    std::vector<double>   x, y;          // n_vertices
    std::vector<uint32_t> cell_vertices; // [cell][4]
    std::vector<uint8_t>  face_kinds;    // [cell][4]   PE_FACE_* bits, 0 = nothing / interior
    double                traction[2] = {0., 0.};
    int                   n_refine = -1; // >= 0 when made by make_unit_square
  };

  struct HostDofs
  {
    int                   element = 0, n_loc = 0;
    int64_t               n_u = 0, n_pint = 0, n_pface = 0;
    std::vector<uint32_t> cell_dofs; // [cell][n_loc], FESystem order, final (renumbered) indices
    // block starts after walking cell c, for c = 0..n_cells  (only kept at partition boundaries)
    int64_t n() const { return n_u + n_pint + n_pface; }
  };

  // contiguous cell block owned by a rank and the three contiguous global row ranges it owns
  // (rows first touched by an owned cell): [u | p_int | p_face]
  struct Partition
  {
    int     rank = 0, world = 1;
    int64_t cell_begin = 0, cell_end = 0;
    int64_t row_begin[3] = {0, 0, 0}, row_end[3] = {0, 0, 0};
    int64_t local_off[3] = {0, 0, 0}; // local row index of the first row of each range
    int64_t n_owned = 0;              // L
    std::vector<uint32_t> ghost_cols;  // sorted global dof ids of the halo columns (H of them)
    std::vector<int64_t>  asm_cells;   // owned cells then ghost cells (global cell ids)
    int64_t n_ghost_cells = 0;
    // cells adjacent to every owned row (global cell ids, ascending): CSR over the owned rows
    std::vector<int64_t> adj_ptr, adj;
    // asm_cells is ordered colour-major: cells of one colour share no owned row
    int                  scatter_mode = 0; // 0 = pre-zeroed matrix + RED atomics, 1 = colour-ordered first-writer stores
    int                  n_colors = 0;
    std::vector<int64_t> color_begin; // n_colors + 1 offsets into asm_cells
    // owned-row ranges of every rank (filled on every rank; used to route halo exchange)
    std::vector<int64_t> all_row_begin, all_row_end; // [world][3]

    inline int64_t
    global_to_owned(const int64_t g) const
    {
      for (int b = 0; b < 3; ++b)
        if (g >= row_begin[b] && g < row_end[b])
          return g - row_begin[b] + local_off[b];
      return -1;
    }
    inline int64_t
    owned_to_global(const int64_t l) const
    {
      for (int b = 2; b >= 0; --b)
        if (l >= local_off[b])
          return l - local_off[b] + row_begin[b];
      return -1;
    }
  };

  struct HostPattern
  {
    std::vector<int64_t> rowptr; // n_owned + 1
    std::vector<int32_t> colidx; // global column ids, ascending in each row
  };

  int  element_n_loc(int element);
  void make_unit_square(HostMesh &m, int n_refine, double distort, uint64_t seed, int boundary_kinds);
  // first-touch numbering + component-wise renumbering; also fills the partition's row ranges
  void distribute_dofs(const HostMesh &m, int element, HostDofs &d, Partition &part);
  // hexahedra (dim = 3, CGQ1-WG, one rank): first-touch numbering of 8 vertices x 3, 6 FE_FaceQ(0) quads, 1 FE_DGQ(0)
  // hex per cell + component_wise; cell_vertices8 = [n_cells][8] in deal.II order
  void distribute_dofs3(int64_t n_vertices, int64_t n_cells, const uint32_t *cell_vertices8, HostDofs &d, Partition &part);
  // rows owned by the partition; ghost cells; halo columns
  void make_sparsity_pattern(const HostMesh &m, const HostDofs &d, Partition &part, HostPattern &p);
  // after pe_set_pattern: ghost cells / halo columns from a user pattern
  void finish_partition(const HostDofs &d, Partition &part, const HostPattern &p);
  // owned-row -> cells adjacency (fills part.adj_ptr / part.adj / asm_cells / ghost cells)
  void build_row_adjacency(const HostMesh &m, const HostDofs &d, Partition &part);
  // colour the assembled cells (structured meshes: 4 colours from the Morton index; otherwise greedy)
  // and sort asm_cells colour-major
  void color_asm_cells(const HostMesh &m, const HostDofs &d, Partition &part);
  // permutation from deal.II BlockSparsityPattern storage order to CSR order
  void dealii_block_order(const HostDofs &d, const HostPattern &p, std::vector<int64_t> &perm);

  // counter-based uniform generator shared by host and device code (splitmix64 finaliser)
  uint64_t splitmix64(uint64_t z);
  double   u01(uint64_t seed, uint64_t id, uint64_t k);
} // namespace pe
