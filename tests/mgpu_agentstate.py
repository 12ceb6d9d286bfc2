"""Multi-GPU script (torchrun): /root/reference/test/mpi/test_agentstate.jl - one agent per rank in a chain, the edge state repeats
the source's state; after every update of the agent states the state read through the edge (an agent of the previous rank) must be
the new one, for :Immortal and for mortal agents.  Every rank runs the same initialisation code and
finish_init!(partition_algo = :EqualAgentNumbers) hands the chain out.  Not run on GPUs yet (MGPU_DRY=1 checks the logic on the CPU)."""
import os
import sys

import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from mgpu_common import setup  # noqa: E402
from models import agentstate_scenario  # noqa: E402


def main():
    be, local, rank, world, _ = setup()
    n = max(world, 2) if os.environ.get("MGPU_DRY") != "1" else 5          # `mpi.size` agents: one per rank
    for immortal in (True, False):
        sim = agentstate_scenario(be, n, immortal, device=local)
        assert sim.num_agents("ASAgent") == n
        assert len(sim.all_agentids("ASAgent", all_ranks=False)) == (1 if world == n else n)
    print(f"rank {rank}/{world}: ok", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
