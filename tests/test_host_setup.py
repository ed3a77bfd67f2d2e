"""Host logic of the product (no GPU): the setup layer S1-S3 of SURVEY.md 8(a) must be bit-exact
against the oracle's restatement of deal.II's numbering and sparsity, and the C-ABI library must
load, export every declared symbol and refuse to compute without a device."""
import ctypes as C
import re
import os

import numpy as np
import pytest

import poroelasticity_b200 as pb
from oracle_lib import Oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "poroelas_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(pe_[a-z_0-9]+)\s*\(", hdr)))
    L = pb._capi.load()
    for name in declared:
        assert hasattr(L, name), name
    assert sorted(declared) == sorted(pb.EXPORTS)


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("box has a GPU")
    with pytest.raises(pb.PoroElasError):
        pb.PoroElasBR1WG()
    b = pb.BiotSystem(element=0, host_setup_only=True)
    b.make_grid_and_dofs(2)
    with pytest.raises(pb.PoroElasError) as ei:
        b.assemble_system(0.1)
    assert ei.value.code == pb._capi.PE_ERR_NO_DEVICE
    with pytest.raises(pb.PoroElasError):
        b.local_matrices(0, 4)


@pytest.mark.parametrize("el", [0, 1])
@pytest.mark.parametrize("nref", [0, 1, 2, 3, 5])
def test_numbering_and_pattern_bit_exact(el, nref):
    o = Oracle(el, nref)
    b = pb.BiotSystem(element=el, host_setup_only=True)
    s = b.make_grid_and_dofs(nref)
    assert (s.n_cells, s.n_u, s.n_pint, s.n_pface, s.n_dofs, s.nnz) == (o.n_cells, o.n_u, o.n_pint, o.n_pface,
                                                                         o.n_dofs, o.nnz)
    assert np.array_equal(b.cell_dofs(), o.cell_dofs())
    rp, ci = b.pattern()
    rpo, cio = o.pattern()
    assert np.array_equal(rp, rpo) and np.array_equal(ci, cio)
    x, y, cv, fk = b.mesh()
    xo, yo, cvo, fko = o.mesh()
    assert np.array_equal(x, xo) and np.array_equal(y, yo) and np.array_equal(cv, cvo)
    assert np.array_equal(fk != 0, fko != 0)


def test_distorted_mesh_identical_to_oracle():
    o = Oracle(0, 4, distort=0.2, seed=12345)
    b = pb.BiotSystem(element=0, host_setup_only=True)
    b.make_grid_and_dofs(4, distort=0.2, seed=12345)
    assert np.array_equal(b.mesh()[0], o.mesh()[0]) and np.array_equal(b.mesh()[1], o.mesh()[1])


def test_user_supplied_mesh_dofs_pattern_roundtrip():
    """pe_set_mesh / pe_set_dofs / pe_set_pattern: what a deal.II host hands over."""
    o = Oracle(0, 3)
    vx, vy, cv, _ = o.mesh()
    b = pb.BiotSystem(element=0, host_setup_only=True)
    b.set_mesh(vx, vy, cv)
    b.distribute_dofs()
    assert np.array_equal(b.cell_dofs(), o.cell_dofs())
    b.make_sparsity_pattern()
    assert np.array_equal(b.pattern()[1], o.pattern()[1])
    b2 = pb.BiotSystem(element=0, host_setup_only=True)
    b2.set_mesh(vx, vy, cv)
    b2.set_dofs(o.n_u, o.n_pint, o.n_pface, o.cell_dofs())
    b2.set_pattern(*o.pattern())
    assert np.array_equal(b2.pattern()[0], o.pattern()[0])
    with pytest.raises(pb.PoroElasError):
        bad = o.pattern()[1].copy()
        bad[0], bad[1] = bad[1], bad[0]
        b2.set_pattern(o.pattern()[0], bad)


def test_dealii_block_order_is_a_permutation_with_diagonal_first():
    b = pb.BiotSystem(element=1, host_setup_only=True)
    s = b.make_grid_and_dofs(2)
    perm = b.dealii_block_order()
    rp, ci = b.pattern()
    assert sorted(perm.tolist()) == list(range(s.nnz))
    # first stored entry of row 0 of block (0,0) is its diagonal
    assert ci[perm[0]] == 0
    rows = np.repeat(np.arange(s.n_dofs), np.diff(rp))
    # block (0,0) entries come first
    n00 = np.sum((rows < s.n_u) & (ci < s.n_u))
    assert np.all(rows[perm[:n00]] < s.n_u) and np.all(ci[perm[:n00]] < s.n_u)


@pytest.mark.parametrize("world", [2, 4])
def test_partition_rows_are_contiguous_and_cover(world):
    """Owner = first-touching cell: every rank owns three contiguous global row ranges; together
    they tile [0, n_dofs) and the owned patterns add up to the global one."""
    o = Oracle(0, 3)
    rpo, cio = o.pattern()
    total = 0
    seen = np.zeros(o.n_dofs, bool)
    for r in range(world):
        b = pb.BiotSystem(element=0, rank=r, world_size=world, host_setup_only=True)
        s = b.make_grid_and_dofs(3)
        assert s.cell_end - s.cell_begin == o.n_cells // world
        rp, ci = b.pattern()
        total += s.nnz_owned
        assert (s.n_ghost_cells > 0) == (r < world - 1)  # ghosts are always LATER cells
        # rows: compare every owned row with the oracle's global row
        cd = o.cell_dofs()
        first_touch = {}
        for c in range(o.n_cells):
            for g in cd[c]:
                first_touch.setdefault(int(g), c)
        owned = sorted(g for g, c in first_touch.items() if s.cell_begin <= c < s.cell_end)
        assert len(owned) == s.n_owned_rows
        for l, g in enumerate(owned):
            assert not seen[g]
            seen[g] = True
            assert np.array_equal(ci[rp[l]:rp[l + 1]], cio[rpo[g]:rpo[g + 1]])
    assert total == o.nnz and seen.all()
