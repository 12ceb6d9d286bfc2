"""Multi-GPU script (torchrun): dead-agent edge purge across ranks (C7, src/AgentMethods.jl:338-345).  Every rank owns some
DAgent and DAgentRemove agents; edges run from the *next* rank's DAgentRemove agents to local DAgent agents.  Killing all
DAgentRemove agents must purge those edges on the ranks that store them (test/remove_agents.jl:39-50 under mpiexec)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import vahana_b200 as vh  # noqa: E402
from models import remove_agents_model  # noqa: E402
from mgpu_common import setup  # noqa: E402


def main():
    be, local, rank, world, _ = setup()
    sim = vh.create_simulation(remove_agents_model(), backend=be, device=local)
    ids = sim.add_agents("DAgent", np.array([(i,) for i in range(1, 4)], dtype=[("idx", "i8")]))
    rids = sim.add_agents("DAgentRemove", None, 2)
    nxt = (rank + 1) % world
    remote_rids = [vh.agent_id(2, nxt, k) for k in (1, 2)]
    remote_ids = [vh.agent_id(1, nxt, k) for k in (1, 2, 3)]
    for r in remote_rids:                      # edges from the next rank's DAgentRemove agents
        sim.add_edge(r, int(ids[1]), "DEdgeState", 0)
    for r in remote_ids:                       # and from its DAgent agents (these survive)
        sim.add_edge(r, int(ids[1]), "DEdge")
    sim.add_edge(int(ids[1]), int(ids[2]), "DEdge")
    sim.finish_init(distribute=False)   # SPMD initialisation
    assert sim.num_edges("DEdgeState") == 2 * world and sim.num_edges("DEdge") == 4 * world
    assert sim.num_agents("DAgentRemove") == 2 * world
    sim.apply("kill_all", ["DAgentRemove"], [], ["DAgentRemove"])
    assert sim.num_agents("DAgentRemove") == 0
    assert sim.num_edges("DEdgeState") == 0, sim.num_edges("DEdgeState")      # purged although the sources died on another rank
    assert sim.num_edges("DEdge") == 4 * world
    sim.apply("die_if_no_edges_DEdge", ["DAgent"], ["DAgent", "DEdge"], ["DAgent"])   # ids[0] dies everywhere
    assert sim.num_agents("DAgent") == 2 * world
    assert sim.num_edges("DEdge") == 3 * world                                # the edge from the next rank's ids[0] is gone
    print(f"rank {rank}/{world}: ok", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
