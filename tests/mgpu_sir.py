"""Multi-GPU parity script (torchrun, one rank per GPU): the SIR model with persons and locations spread over the ranks in
contiguous equal blocks.  Every Visit / Exposure edge whose target lives on another rank is redistributed over NCCL
(transmit_edges!), sources become ghosts on the receiver.  Person states and location tallies must equal the single-rank
oracle bit for bit (with 2 visits per person the Float32 risk sum of a row is order independent)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import vahana_b200 as vh  # noqa: E402
from models import sir_model, sir_step, sir_sim, PERSON  # noqa: E402
from mgpu_common import setup, oracle_backend  # noqa: E402


def gather(local, sizes, rank, world, dev="cuda"):
    parts = []
    for r in range(world):
        t = torch.from_numpy(local.copy()).to(dev) if r == rank else torch.empty(sizes[r], dtype=torch.from_numpy(local[:0].copy()).dtype, device=dev)
        dist.broadcast(t, src=r)
        parts.append(t.cpu().numpy())
    return np.concatenate(parts)


def main():
    npers, nloc = int(os.environ.get("MGPU_N", "60000")), int(os.environ.get("MGPU_L", "4001"))
    be, local, rank, world, dev = setup()
    pb, lb = vh.equal_partition(npers, world), vh.equal_partition(nloc, world)
    st = np.zeros(npers, dtype=np.dtype(PERSON, align=True))
    st["state"] = (np.random.default_rng(7).random(npers) < 0.01).astype("u1")
    g = vh.create_simulation(sir_model(), params={"n_locations": nloc, "beta": 0.3, "n_ranks": world}, backend=be, device=local)
    g.add_agents("Person", st[pb[rank]:pb[rank + 1]])
    g.add_agents("Location", np.zeros(lb[rank + 1] - lb[rank], dtype=[("n_inf", "i4")]))
    g.set_uniform_offset("Person", pb[rank])
    g.set_uniform_offset("Location", lb[rank])
    g.finish_init(distribute=False)   # SPMD initialisation: every rank added its own block
    o = None
    if rank == 0:
        o = sir_sim(oracle_backend(), npers, nloc, beta=0.3)
    psz = [pb[r + 1] - pb[r] for r in range(world)]
    lsz = [lb[r + 1] - lb[r] for r in range(world)]
    for step in range(8):
        sir_step(g, step)
        nv, ne = g.num_edges("Visit"), g.num_edges("Exposure")
        ps = gather(g.all_agents("Person", all_ranks=False).view("i2").astype("i4"), psz, rank, world, dev)
        ls = gather(g.all_agents("Location", all_ranks=False)["n_inf"].copy(), lsz, rank, world, dev)
        if rank == 0:
            sir_step(o, step)
            assert nv == o.num_edges("Visit") == 2 * npers and ne == o.num_edges("Exposure") == 2 * npers
            assert np.array_equal(ps, o.all_agents("Person").view("i2").astype("i4")), step
            assert np.array_equal(ls, o.all_agents("Location")["n_inf"]), step
    print(f"rank {rank}/{world}: ok", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
