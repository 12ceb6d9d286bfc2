"""Multi-GPU script (torchrun): finish_init!(distribute = true, partition_algo = :EqualAgentNumbers) (src/Simulation.jl:403-476,
distribute! src/MPI.jl:11-84).  Every rank runs the same initialisation code (the docs' Hegselmann-Krause model on a
Barabasi-Albert graph, BASELINE config 1 at a fifth of its size); finish_init hands rank 0's graph out, and the sharded run must
follow the single-rank oracle on the same graph.  Three hand-outs: equal blocks, equal blocks of an agent count that the ranks do not
divide (blocks of different length: a remote source number may exceed this rank's own block), an explicit, unequal `partition`, and
`partition_algo = :Metis` (the reference's default) through the graph-growing stand-in."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import vahana_b200 as vh  # noqa: E402
from mgpu_common import setup, oracle_backend  # noqa: E402
from models import hk_model, hk_sim, ba_graph  # noqa: E402


def run_case(be, local, rank, world, dev, n, partition_of):
    uv = ba_graph(n, 8, 1)
    op0 = np.random.default_rng(1).random(n)
    g = vh.create_simulation(hk_model(), params={"eps": 0.02}, backend=be, device=local)
    ids = vh.add_graph(g, uv, n, "HKAgent", op0.view([("opinion", "f8")]), "Knows")      # the same code on every rank
    g.add_edges(ids, ids, "Knows")
    old = np.array([vh.agent_id(1, 0, k) for k in range(1, n + 1)], dtype=np.uint64)     # rank 0's ids of the init phase
    if partition_of == "Metis":        # the reference's default; graph growing stands in for Metis (vh.graph_growing_partition)
        m = g.finish_init(return_idmapping=True, partition_algo="Metis")
        owner = np.array([vh.process_nr(m[int(old[k])]) for k in range(n)])
        sizes_ = np.bincount(owner, minlength=world)
        assert sizes_.max() - sizes_.min() <= 1
    elif partition_of is None:
        m = g.finish_init(return_idmapping=True, partition_algo="EqualAgentNumbers")
        b = vh.equal_partition(n, world)
        owner = np.searchsorted(np.array(b[1:]), np.arange(n), side="right")
    else:
        owner = partition_of(n, world)
        m = g.finish_init(return_idmapping=True, partition={int(old[k]): int(owner[k]) + 1 for k in range(n)})
    # new id = (type, owner, position among the owner's agents in the old order)
    pos = np.zeros(n, dtype=np.int64)
    for r in range(world):
        sel = np.nonzero(owner == r)[0]
        pos[sel] = np.arange(1, len(sel) + 1)
    assert len(m) == n and all(m[int(old[k])] == vh.agent_id(1, int(owner[k]), int(pos[k])) for k in range(0, n, 97))
    sizes = [int((owner == r).sum()) for r in range(world)]
    assert len(g.all_agents("HKAgent", all_ranks=False)) == sizes[rank]
    assert g.num_agents("HKAgent") == n and g.num_edges("Knows") == 2 * len(uv) + n
    o = None
    if rank == 0:
        o, _ = hk_sim(oracle_backend(), n, uv, op0)
    order = np.concatenate([np.nonzero(owner == r)[0] for r in range(world)])            # global index of every position of the joined vector
    for step in range(4):
        g.apply("hk_step", "HKAgent", ["HKAgent", "Knows"], "HKAgent")
        mine = torch.from_numpy(g.all_agents("HKAgent", all_ranks=False)["opinion"].copy()).to(dev)
        parts = []
        for r in range(world):
            t = mine if r == rank else torch.empty(sizes[r], dtype=torch.float64, device=dev)
            dist.broadcast(t, src=r)
            parts.append(t)
        if rank == 0:
            o.apply("hk_step", "HKAgent", ["HKAgent", "Knows"], "HKAgent")
            got = np.zeros(n)
            got[order] = torch.cat(parts).cpu().numpy()
            np.testing.assert_allclose(got, o.all_agents("HKAgent")["opinion"], rtol=1e-12, atol=0)
    g.finish_simulation()


def main():
    n = int(os.environ.get("MGPU_N", "20000"))
    be, local, rank, world, dev = setup()
    run_case(be, local, rank, world, dev, n, None)
    run_case(be, local, rank, world, dev, n + 1 if world > 1 and (n + 1) % world else n + world + 1, None)     # blocks of different length
    # an explicit partition: rank 0 gets three quarters of the agents, interleaved with the others'
    run_case(be, local, rank, world, dev, n // 2, lambda nn, w: np.where(np.arange(nn) % 4 != 3, 0, 1 + (np.arange(nn) // 4) % max(w - 1, 1)) % w)
    run_case(be, local, rank, world, dev, n // 4, "Metis")
    print(f"rank {rank}/{world}: ok", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
