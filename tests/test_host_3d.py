"""Host side of the 3-D CGQ1-WG path (pe_distribute_dofs / pe_make_sparsity_pattern with dim = 3) against the oracle's
restatement of DoFHandler::distribute_dofs + component_wise + make_sparsity_pattern on hexahedra: bit-exact.
No GPU: the handle is created with PE_CFG_HOST_SETUP_ONLY."""
import numpy as np
import pytest

from oracle_lib import Oracle3D
from poroelasticity_b200 import PoroElas_CGQ1_WG_3D, PoroElasError


@pytest.mark.parametrize("n_refine,distort", [(0, 0.0), (1, 0.0), (2, 0.2), (3, 0.0)])
def test_numbering_and_pattern_match_the_oracle(n_refine, distort):
    o = Oracle3D(n_refine, distort)
    (n_u, n_pint, n_pface), cd, rp, ci = o.dofs_pattern()
    g = PoroElas_CGQ1_WG_3D(host_setup_only=True)
    g.make_grid(n_refine, distort)
    g.distribute_dofs()
    g.make_sparsity_pattern()
    s = g.sizes()
    assert (s.n_u, s.n_pint, s.n_pface, s.n_loc, s.n_cells) == (n_u, n_pint, n_pface, 31, o.n_cells)
    assert np.array_equal(g.cell_dofs(), cd)
    rp2, ci2 = g.pattern()
    assert np.array_equal(rp2, rp) and np.array_equal(ci2, ci)
    g.close()


def test_user_mesh_in_another_cell_order():
    """numbering follows the order of the cells handed in (first touch), not the geometry"""
    o = Oracle3D(1)
    x, y, z, cv = o.mesh()
    perm = np.array([5, 2, 7, 0, 3, 6, 1, 4])
    g = PoroElas_CGQ1_WG_3D(host_setup_only=True)
    g.set_mesh3(x, y, z, cv[perm])
    g.distribute_dofs()
    cd = g.cell_dofs()
    assert list(cd[0, :24]) == list(range(24))
    s = g.sizes()
    assert cd[0, 30] == s.n_u and (s.n_u, s.n_pint, s.n_pface) == (81, 8, 36)
    # same entities, other numbers: the multiset of (cell -> number of cells sharing each face dof) is unchanged
    _, cnt = np.unique(cd[:, 24:30], return_counts=True)
    assert sorted(cnt) == [1] * 24 + [2] * 12
    g.close()


def test_hot_path_in_3d_is_refused_not_faked():
    g = PoroElas_CGQ1_WG_3D(host_setup_only=True)
    g.make_grid(1)
    g.distribute_dofs()
    g.make_sparsity_pattern()
    with pytest.raises(PoroElasError):
        g.assemble_system(0.0)
    g.close()


def test_calls_out_of_order_or_not_built_in_3d_are_refused():
    """the same refusals tests/test_gpu_3d.py checks on the GPU box are decided before any device access"""
    b = PoroElas_CGQ1_WG_3D(host_setup_only=True)
    b.make_grid(1)
    with pytest.raises(PoroElasError):
        b.assemble_system(0.1)            # before distribute_dofs / make_sparsity_pattern
    with pytest.raises(PoroElasError):
        b.solve_time_step()               # nothing assembled
    with pytest.raises(PoroElasError):
        b.postprocess()                   # not built for dim = 3
    with pytest.raises(PoroElasError):
        b.set_solution(np.zeros(10))
    b.close()
