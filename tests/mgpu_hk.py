"""Multi-GPU parity script (run under torchrun, one rank per GPU): Hegselmann–Krause on the power-law graph sharded by
contiguous equal blocks, halo exchange over NCCL, compared with the single-rank oracle on the same global graph.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_hk.py"""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import vahana_b200 as vh  # noqa: E402
from models import hk_model  # noqa: E402


def main():
    n = int(os.environ.get("MGPU_N", "200000"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    be = vh.default_backend()
    be.init(local)
    be.set_stream(torch.cuda.current_stream().cuda_stream)
    assert be.init_distributed() == (rank, world)
    g = vh.create_simulation(hk_model(), backend=be, device=local)
    ne = C.c_uint64()
    be.check(be.lib.vbw_hk_powerlaw_build_sharded(g.h, 1, 0, C.c_uint64(n), C.c_uint64(4), C.c_uint64(5), C.c_double(6.8333), C.c_uint32(1000000),
                                                   C.c_uint64(50000), C.c_uint32(rank), C.c_uint32(world), C.byref(ne)))
    g.finish_init(distribute=False)   # SPMD initialisation: every rank added its own block
    bounds = vh.equal_partition(n, world)
    assert len(g.all_agents("HKAgent", all_ranks=False)) == bounds[rank + 1] - bounds[rank]

    o = None
    if rank == 0:   # the oracle runs the whole graph on one rank
        import subprocess
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True)
        ob = vh.load_backend(os.path.join(ROOT, "oracle", "_build", "libvahana_oracle.so"))
        tot = C.c_uint64()
        ob.lib.vbw_hk_powerlaw_host(C.c_uint64(n), 1, C.c_uint64(4), C.c_uint64(5), C.c_double(6.8333), C.c_uint32(1000000), None, None, None, C.byref(tot))
        fr, to, op = np.zeros(tot.value, dtype=np.uint64), np.zeros(tot.value, dtype=np.uint64), np.zeros(n)
        ob.lib.vbw_hk_powerlaw_host(C.c_uint64(n), 1, C.c_uint64(4), C.c_uint64(5), C.c_double(6.8333), C.c_uint32(1000000),
                                    fr.ctypes.data_as(C.c_void_p), to.ctypes.data_as(C.c_void_p), op.ctypes.data_as(C.c_void_p), C.byref(tot))
        o = vh.create_simulation(hk_model(), backend=ob)
        o.add_agents("HKAgent", op.view([("opinion", "f8")]))
        o.add_edges(fr, to, "Knows")
        o.finish_init()
        assert g.num_edges("Knows") == o.num_edges("Knows") == tot.value     # collective: summed over ranks
    else:
        g.num_edges("Knows")
    assert g.num_agents("HKAgent") == n

    for step in range(4):
        g.apply("hk_step", "HKAgent", ["HKAgent", "Knows"], "HKAgent")
        mine = torch.from_numpy(g.all_agents("HKAgent", all_ranks=False)["opinion"].copy()).cuda()
        sizes = [bounds[r + 1] - bounds[r] for r in range(world)]
        parts = [torch.empty(s, dtype=torch.float64, device="cuda") for s in sizes]
        dist.all_gather(parts, mine) if len(set(sizes)) == 1 else [dist.broadcast(parts[r] if r != rank else mine, src=r) for r in range(world)]
        if len(set(sizes)) != 1:
            parts[rank] = mine
        total = g.mapreduce("opinion", "+", "HKAgent")
        if rank == 0:
            o.apply("hk_step", "HKAgent", ["HKAgent", "Knows"], "HKAgent")
            ref = o.all_agents("HKAgent")["opinion"]
            got = torch.cat(parts).cpu().numpy()
            np.testing.assert_allclose(got, ref, rtol=1e-12, atol=0)
            assert abs(total - o.mapreduce("opinion", "+", "HKAgent")) < 1e-9 * n
    hb = C.c_uint64()
    be.lib.vb_halo_bytes(g.h, C.byref(hb))
    st = g.last_apply_stats()
    if os.environ.get("MGPU_EXPECT_PREFILTER"):      # the swept [local | ghost] read phase must be the one that ran
        assert st["prefiltered"] and st["source_blocks"] >= 2, st
    print(f"rank {rank}/{world}: ok, ghosts bytes/step {hb.value}, prefiltered {st['prefiltered']}, key blocks {st['source_blocks']}", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
