"""Multi-GPU script (torchrun): remove_edges! on targets that live on another rank (transmit_remove_edges! /
removeedges_alltoall!, src/MPI.jl:432-479) - the counts of /root/reference/test/mpi/test_edgetypes.jl:295-449 under mpiexec.
The 100 agents are spread over the ranks in contiguous equal blocks, every edge is stored on its target's rank; on the cycle graph
the predecessor of a block's first agent lives on the previous rank, on the complete graph most neighbours do."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import vahana_b200 as vh  # noqa: E402
from mgpu_common import setup  # noqa: E402
from models import edges_model, foos  # noqa: E402

REMOVE_TYPES = ["EdgeD", "EdgeS", "EdgeE", "EdgeI", "EdgeSE", "EdgeSI", "EdgeEI", "EdgeSEI", "EdgeSTI", "EdgeSETI", "EdgeT", "EdgeST"]
REMOVE_FROM_TYPES = ["EdgeD", "EdgeS", "EdgeE", "EdgeSE", "EdgeT", "EdgeST"]
N = 100


def graph_sim(be, local, rank, world, ET, kind):
    bounds = vh.equal_partition(N, world)
    owner = np.searchsorted(np.array(bounds[1:]), np.arange(N), side="right")
    gid = np.array([vh.agent_id(1, int(owner[g]), g - bounds[owner[g]] + 1) for g in range(N)], dtype=np.uint64)   # "Agent" is type 1
    sim = vh.create_simulation(edges_model(), backend=be, device=local)
    ids = sim.add_agents("Agent", foos(range(bounds[rank] + 1, bounds[rank + 1] + 1)))     # foo = global index + 1
    assert np.array_equal(ids, gid[bounds[rank]:bounds[rank + 1]])
    sim.set_uniform_offset("Agent", bounds[rank])
    if kind == "cycle":     # cycle_digraph: i -> i + 1
        fg, tg = np.arange(N), np.roll(np.arange(N), -1)
    else:                   # complete_graph via add_graph!: both directions per undirected edge, in edge order
        uv = np.array([(i, j) for i in range(N) for j in range(i + 1, N)])
        fg = np.stack([uv[:, 0], uv[:, 1]], axis=1).reshape(-1)
        tg = np.stack([uv[:, 1], uv[:, 0]], axis=1).reshape(-1)
    mine = owner[tg] == rank                                                               # an edge is stored on its target's rank
    fr, to = gid[fg], gid[tg]
    stateful = "S" not in ET[4:]
    sim.add_edges(fr[mine], to[mine], ET, foos(np.zeros(int(mine.sum()), dtype=int)) if stateful else None)
    sim.finish_init(distribute=False)   # SPMD initialisation: every rank added its own block
    return sim


def main():
    be, local, rank, world, _ = setup()
    fresh = lambda ET, kind: graph_sim(be, local, rank, world, ET, kind)    # noqa: E731  (a new simulation per apply: no copy_simulation)
    for ET in REMOVE_TYPES:                                                       # test_edgetypes.jl:295-352
        sim = fresh(ET, "cycle")
        assert sim.num_edges(ET) == 100
        sim.apply(f"remove_own_if_even_{ET}", "Agent", ["Agent", ET], ET, add_existing=ET)
        assert sim.num_edges(ET) == 50, (ET, sim.num_edges(ET))
        if "I" not in ET[4:]:
            # the neighbour (= source of the own edge) of a block's first agent lives on the previous rank: the request travels
            sim = fresh(ET, "cycle")
            sim.apply(f"remove_neighbors_if_even_{ET}", "Agent", ["Agent", ET], ET, add_existing=ET)
            assert sim.num_edges(ET) == 50, (ET, sim.num_edges(ET))
            sim = fresh(ET, "cycle")
            sim.apply(f"remove_and_readd_{ET}", "Agent", ET, ET, add_existing=ET)
            assert sim.num_edges(ET) == 100, (ET, sim.num_edges(ET))
    # the order rule (src/Simulation.jl:792-800): a travelling removal is applied before the new edges arrive.  Every agent clears its
    # neighbour's row (on the previous rank for a block's first agent) and points an edge back at it: 100 edges, all reversed.
    for ET in ["EdgeD", "EdgeS", "EdgeT"]:
        sim = fresh(ET, "cycle")
        sim.apply(f"clear_neighbor_row_and_point_back_{ET}", "Agent", ["Agent", ET], ET, add_existing=ET)
        assert sim.num_edges(ET) == 100, (ET, sim.num_edges(ET))
        bounds = vh.equal_partition(N, world)
        mine = sim.all_agentids("Agent", all_ranks=False)
        sim.disable_transition_checks(True)
        for k, aid in enumerate(mine):
            g = bounds[rank] + k                                                  # global index; the successor owns the only edge of this row
            succ = (g + 1) % N
            owner = int(np.searchsorted(np.array(bounds[1:]), succ, side="right"))
            nb = sim.neighborids(int(aid), ET)
            assert [int(x) for x in np.atleast_1d(nb)] == [vh.agent_id(1, owner, succ - bounds[owner] + 1)], (ET, g, nb)
        sim.disable_transition_checks(False)
    for ET in REMOVE_FROM_TYPES:                                                  # test_edgetypes.jl:354-449
        single = "E" in ET[4:]
        kind = "cycle" if single else "complete"
        sim = fresh(ET, kind)
        assert sim.num_edges(ET) == (100 if single else 9900)
        sim.apply(f"remove_first_two_from_{ET}", "Agent", ["Agent", ET], ET, add_existing=ET)
        assert sim.num_edges(ET) == (0 if single else 9700), (ET, sim.num_edges(ET))
        if not single:
            # remove_edges!(sim, id, nid, ET) with a random neighbour nid: most of them live on another rank
            sim = fresh(ET, kind)
            sim.apply(f"remove_to_random_neighbor_if_even_{ET}", "Agent", ["Agent", ET], ET, add_existing=ET, seed=11)
            assert sim.num_edges(ET) == 9850, (ET, sim.num_edges(ET))
        sim = fresh(ET, kind)
        sim.apply(f"remove_from_zero_{ET}", "Agent", [], ET, add_existing=ET)
        assert sim.num_edges(ET) == (100 if single else 9900)
    print(f"rank {rank}/{world}: ok", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
