// Header-only adaptor for a box that HAS deal.II (>= 9.1): feeds the C ABI from the reference's own
// Triangulation / DoFHandler / SparsityPattern, so that (a) the real reference can be timed against
// the same handle and (b) every [deal.II-memory] restatement of oracle/dealii_restated.hpp can be
// diffed against deal.II itself.  NOT compiled in this repository's build (deal.II is not installed
// here).  See INTEGRATION.md for the five-line patch to the reference.
#pragma once
#ifdef POROELAS_WITH_DEALII
#  include <deal.II/dofs/dof_handler.h>
#  include <deal.II/grid/tria.h>
#  include <deal.II/lac/block_sparse_matrix.h>
#  include <deal.II/lac/block_sparsity_pattern.h>
#  include <deal.II/lac/block_vector.h>

#  include "../include/poroelas_b200.h"

#  include <vector>

namespace poroelas
{
  // make_grid_and_dofs(): hand mesh, numbering and sparsity of the reference over to the library
  template <int dim>
  inline int
  attach(pe_handle *h, const dealii::Triangulation<dim> &tria, const dealii::DoFHandler<dim> &dof_handler,
         const dealii::BlockSparsityPattern &sp, const unsigned int n_u, const unsigned int n_p_interior,
         const unsigned int n_p_face)
  {
    const auto         &v = tria.get_vertices(); // contiguous Point<dim> = AoS doubles
    std::vector<uint32_t> cv, cd;
    const unsigned int    n_loc = dof_handler.get_fe().dofs_per_cell;
    std::vector<dealii::types::global_dof_index> idx(n_loc);
    for (const auto &cell : dof_handler.active_cell_iterators())
      {
        for (unsigned int k = 0; k < dealii::GeometryInfo<dim>::vertices_per_cell; ++k)
          cv.push_back(cell->vertex_index(k));
        cell->get_dof_indices(idx);
        cd.insert(cd.end(), idx.begin(), idx.end());
      }
    int rc = pe_set_mesh(h, (int64_t)v.size(), &v[0][0], &v[0][1], dim, (int64_t)tria.n_active_cells(), cv.data());
    if (rc)
      return rc;
    rc = pe_set_dofs(h, n_u, n_p_interior, n_p_face, cd.data());
    if (rc)
      return rc;
    // global CSR with ascending columns from the 2x2 block pattern
    const dealii::types::global_dof_index n = dof_handler.n_dofs();
    std::vector<int64_t>                  rowptr(n + 1, 0);
    std::vector<int32_t>                  col;
    for (dealii::types::global_dof_index r = 0; r < n; ++r)
      {
        std::vector<int32_t> row;
        for (auto it = sp.begin(r); it != sp.end(r); ++it)
          row.push_back((int32_t)it->column());
        std::sort(row.begin(), row.end());
        col.insert(col.end(), row.begin(), row.end());
        rowptr[r + 1] = (int64_t)col.size();
      }
    return pe_set_pattern(h, rowptr.data(), col.data());
  }

  // after pe_assemble_system: copy the values into the reference's own system_matrix / system_rhs
  inline int
  download(pe_handle *h, dealii::BlockSparseMatrix<double> &system_matrix, dealii::BlockVector<double> &system_rhs)
  {
    pe_sizes s;
    pe_get_sizes(h, &s);
    std::vector<double>  val((size_t)s.nnz), rhs((size_t)s.n_dofs);
    std::vector<int64_t> perm((size_t)s.nnz), rowptr((size_t)s.n_dofs + 1);
    std::vector<int32_t> col((size_t)s.nnz);
    int                  rc = pe_get_matrix(h, val.data());
    rc                      = rc ? rc : pe_get_rhs(h, rhs.data());
    rc                      = rc ? rc : pe_get_pattern(h, rowptr.data(), col.data());
    if (rc)
      return rc;
    for (int64_t r = 0; r < s.n_dofs; ++r)
      for (int64_t k = rowptr[r]; k < rowptr[r + 1]; ++k)
        system_matrix.set(r, col[k], val[k]);
    for (int64_t r = 0; r < s.n_dofs; ++r)
      system_rhs(r) = rhs[r];
    return 0;
  }
} // namespace poroelas
#endif
