"""Multi-GPU script (torchrun): Game of Life on a raster whose cells are handed out to the ranks (BASELINE config 2 on several GPUs).
Every rank runs the same initialisation code (add_raster! + connect_raster_neighbors!), finish_init! hands the cells of rank 0 out in
equal blocks and broadcasts the id grid (broadcastids, /root/reference/src/MPI.jl:59-73, src/Raster.jl:64-75); the stencil edges follow
their targets, the neighbours of a block's border cells live on other ranks and are read through the halo.  rastervalues /
calc_raster join the ranks (src/Raster.jl:227,318,378 -> src/MPI.jl:492-517), so every rank sees the whole grid: compared with a numpy
Game of Life after every generation, and with the single-rank oracle on rank 0."""
import os
import sys

import numpy as np
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import vahana_b200 as vh  # noqa: E402
from mgpu_common import setup, oracle_backend  # noqa: E402
from models import gol_model  # noqa: E402


def life(a):
    n = sum(np.roll(np.roll(a, dx, 0), dy, 1) for dx in (-1, 0, 1) for dy in (-1, 0, 1) if (dx, dy) != (0, 0))
    return (n == 3) | (a & (n == 2))


def build(be, init, **kw):
    sim = vh.create_simulation(gol_model(), backend=be, **kw)
    ids = sim.add_raster("grid", init.shape, "Cell", np.asarray(init, dtype="?").reshape(-1, order="F").view([("active", "?")]))
    sim.connect_raster_neighbors("grid", "Neighbor")
    idmap = sim.finish_init(return_idmapping=True)
    return sim, ids, idmap


def main():
    be, local, rank, world, _ = setup()
    init = np.random.default_rng(2).random((61, 47)) < 0.35
    sim, ids, idmap = build(be, init, device=local)
    n = init.size
    b = vh.equal_partition(n, world)
    assert sim.num_agents("Cell") == n and len(sim.all_agents("Cell", all_ranks=False)) == b[rank + 1] - b[rank]
    assert sim.num_edges("Neighbor") == 8 * n
    # the id grid every rank holds: cell k of the column-major order lives on the rank of its block, with the number inside the block
    flat = ids.reshape(-1, order="F")
    for k in (0, 1, b[1] - 1, b[1] % n, n - 1):
        assert vh.agent_nr(int(flat[k])) == k + 1                       # (this rank's own initialisation phase; rank 0's is the one handed out)
        new = idmap[vh.agent_id(1, 0, k + 1)]
        owner = int(np.searchsorted(np.array(b[1:]), k, side="right"))
        assert new == vh.agent_id(1, owner, k - b[owner] + 1)
        pos = tuple(int(x) + 1 for x in np.unravel_index(k, init.shape, order="F"))
        assert sim.cellid("grid", pos) == new
    o = None
    if rank == 0 and world > 1:
        o, _, _ = build(oracle_backend(), init)
    a = init.copy()
    assert np.array_equal(sim.rastervalues("grid", "active", "Cell"), a)          # joined: the whole grid on every rank
    for gen in range(8):
        sim.apply("gol_life", "Cell", ["Cell", "Neighbor"], "Cell")
        a = life(a)
        got = sim.rastervalues("grid", "active", "Cell")
        assert np.array_equal(got, a), (gen, int((got != a).sum()))
        if o is not None:
            o.apply("gol_life", "Cell", ["Cell", "Neighbor"], "Cell")
            assert np.array_equal(o.rastervalues("grid", "active", "Cell"), got)
    assert sim.mapreduce("active", "+", "Cell", datatype="i8") == int(a.sum())
    ne = sim.calc_raster_num_edges("grid", "Neighbor") if hasattr(sim, "calc_raster_num_edges") else None
    if ne is not None:
        assert np.array_equal(np.asarray(ne).reshape(-1), np.full(n, 8))
    print(f"rank {rank}/{world}: ok, alive {int(a.sum())}", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
