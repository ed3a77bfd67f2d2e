"""N > 1 host logic on CPU: two processes (gloo, world_size 2) build their partition views with host-setup-only
handles and check, by exchanging them, that the distributed layout is consistent: the owned row ranges tile the
global numbering, every halo column a rank needs is owned by the other rank, and the ghost cells of rank 0 are
cells of rank 1."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, element, nref, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import poroelasticity_b200 as pb
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        b = pb.BiotSystem(element=element, rank=rank, world_size=world, host_setup_only=True)
        s = b.make_grid_and_dofs(nref)
        owned = b.owned_rows()
        rp, ci = b.pattern()
        assert owned.size == s.n_owned_rows and rp.size == owned.size + 1
        # every rank learns the owner of every global row
        owner = torch.full((s.n_dofs,), -1, dtype=torch.int64)
        owner[torch.from_numpy(owned)] = rank
        dist.all_reduce(owner, op=dist.ReduceOp.MAX)
        cover = torch.zeros(s.n_dofs, dtype=torch.int64)
        cover[torch.from_numpy(owned)] = 1
        dist.all_reduce(cover, op=dist.ReduceOp.SUM)
        assert bool((cover == 1).all()), "owned ranges must tile [0, n_dofs) exactly once"
        # halo columns: referenced by my rows, owned by somebody else -> the owner must exist and differ from me
        halo = np.setdiff1d(np.unique(ci), owned)
        assert np.all(owner.numpy()[halo] >= 0) and np.all(owner.numpy()[halo] != rank)
        # cells: contiguous blocks, ghosts are later cells only
        cells = torch.tensor([s.cell_begin, s.cell_end, s.n_ghost_cells], dtype=torch.int64)
        allc = [torch.zeros_like(cells) for _ in range(world)]
        dist.all_gather(allc, cells)
        assert int(allc[0][0]) == 0 and int(allc[-1][1]) == s.n_cells
        for r in range(world - 1):
            assert int(allc[r][1]) == int(allc[r + 1][0])
        assert int(allc[-1][2]) == 0 and all(int(allc[r][2]) > 0 for r in range(world - 1))
        # the owned patterns add up to the global number of entries
        nnz = torch.tensor([s.nnz_owned], dtype=torch.int64)
        dist.all_reduce(nnz)
        one = pb.BiotSystem(element=element, host_setup_only=True)
        assert int(nnz) == one.make_grid_and_dofs(nref).nnz
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover - reported to the parent
        q.put((rank, "FAIL: %r" % (e,)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("element,world", [(0, 2), (1, 2), (0, 4)])
def test_partition_views_are_consistent(element, world):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + element + 10 * world
    procs = [ctx.Process(target=_worker, args=(r, world, port, element, 4, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(msg == "ok" for _, msg in res), res
