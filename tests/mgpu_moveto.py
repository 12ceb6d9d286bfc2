"""Multi-GPU script (torchrun): agents placed on a raster (test/raster.jl:242-303 as `mpiexec` runs it).  Every rank runs the same
initialisation code — add_raster!, add_agent!, move_to! — and finish_init!(distribute = true) hands rank 0's content out: the cells in
equal blocks, the movers in equal blocks, every move_to! edge to the rank of its target (/root/reference/src/Raster.jl:437-477,
src/MPI.jl:11-84).  A cell then sums the values of the movers standing on it (most of them live on another rank: read through the
halo), every mover reads its cell's sum and moves to (sum, sum) inside a transition (Ctx::move_to towards a cell of another rank:
both edges travel through transmit_edges!), and the cells sum again.  calc_rasterstate / calc_raster join the ranks.  Compared with
the expected grids and, on rank 0, with the single-rank oracle."""
import os
import sys

import numpy as np
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import vahana_b200 as vh  # noqa: E402
from mgpu_common import setup, oracle_backend  # noqa: E402
from models import raster_model  # noqa: E402

DIMS = (10, 10)
VALUES = [1, 2, 3, 4, 6, 7, 2, 2, 1, 3]
PLACES = [(1, 1), (2, 2), (2, 2), (3, 7), (10, 10), (5, 5), (6, 1), (6, 1), (9, 3), (10, 1)]


def build(be, **kw):
    sim = vh.create_simulation(raster_model(), backend=be, **kw)
    sim.add_raster("raster", DIMS, "Position", lambda p: (0,))
    movers = [sim.add_agent("MovingAgent", v) for v in VALUES]
    for a, p in zip(movers, PLACES):
        sim.move_to("raster", a, p, "OnPosition", "OnPosition")
    sim.finish_init()
    return sim


def sums(values, places):
    g = np.zeros(DIMS, dtype=np.int64)
    for v, p in zip(values, places):
        g[p[0] - 1, p[1] - 1] += v
    return g


def main():
    be, local, rank, world, _ = setup()
    sim = build(be, device=local)
    o = build(oracle_backend()) if rank == 0 and world > 1 else None
    n = DIMS[0] * DIMS[1]
    assert sim.num_agents("Position") == n and sim.num_agents("MovingAgent") == len(VALUES)
    assert sim.num_edges("OnPosition") == 2 * len(VALUES)
    b = vh.equal_partition(len(VALUES), world)
    assert len(sim.all_agents("MovingAgent", all_ranks=False)) == b[rank + 1] - b[rank]

    def step(s, name):
        if name == "sum":
            s.apply("sum_on_pos", ["Position"], ["MovingAgent", "OnPosition"], ["Position"])
        else:
            s.apply("value_on_pos", ["MovingAgent"], ["Position", "OnPosition"], ["OnPosition", "MovingAgent"])

    values, places = list(VALUES), list(PLACES)
    for rnd in range(3):
        step(sim, "sum")
        want = sums(values, places)
        got = np.asarray(sim.calc_rasterstate("raster", "ids_sum", "Position"))
        assert np.array_equal(got, want), (rnd, got.tolist())
        ne = np.asarray(sim.calc_raster_num_edges("raster", "OnPosition"))
        assert np.array_equal(ne, sums([1] * len(values), places)), rnd
        if o is not None:
            step(o, "sum")
            assert np.array_equal(np.asarray(o.calc_rasterstate("raster", "ids_sum", "Position")), got)
        nxt = [int(want[p[0] - 1, p[1] - 1]) for p in places]
        if max(nxt) > min(DIMS):      # the next move would leave the raster
            break
        # every mover takes the sum of its cell as its value and moves to (sum, sum)
        step(sim, "move")
        if o is not None:
            step(o, "move")
        values, places = nxt, [(v, v) for v in nxt]
        assert sim.num_edges("OnPosition") == 2 * len(values)
        assert sorted(int(x) for x in sim.all_agents("MovingAgent")["value"]) == sorted(values)
    assert rnd == 1      # two sums around one move
    print(f"rank {rank}/{world}: ok", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
