"""Multi-GPU script (torchrun): /root/reference/test/core.jl as `mpiexec` runs it (test/mpi/test_core.jl just includes it): every rank
runs the same initialisation code, finish_init!(return_idmapping = true, partition_algo = :EqualAgentNumbers) hands rank 0's network
out, ids are renewed with updateids, per-agent checks run `@onrankof` the agent, counts and folds are collective.
Written after round 1's GPU budget was spent: not run on GPUs yet."""
import os
import sys
from functools import reduce
import operator

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import vahana_b200 as vh  # noqa: E402
from mgpu_common import setup  # noqa: E402
from models import core_model, add_example_network  # noqa: E402

ALLAGENTTYPES = ["AMortal", "AImm", "AImmFixed"]


def createsim(be, local, loops=()):
    """createsim (test/core.jl:108-121); `loops` names the extra self loops core.jl adds behind finish_init! with its @onrankof hack
    (:303-311) - here they are part of the initialisation phase, which builds the same network"""
    sim = vh.create_simulation(core_model(), backend=be, device=local)
    a1, a2, a3, avids, avfids = add_example_network(sim)
    if "a1" in loops:
        sim.add_edge(a1, a1, "ESLDict1")
    if "avids" in loops:
        sim.add_edge(avids[0], avids[0], "ESLDict2")
    if "avfids" in loops:
        sim.add_edge(avfids[0], avfids[0], "ESLDict1")
        sim.add_edge(avfids[1], avfids[0], "ESLDict1")
        sim.add_edge(avfids[1], avfids[1], "ESLDict1")
    idmap = sim.finish_init(return_idmapping=True, partition_algo="EqualAgentNumbers")
    up = lambda x: vh.updateids(idmap, x)   # noqa: E731  (core.jl:116-118)
    return sim, up(a1), up(a2), up(a3), [int(x) for x in up(avids)], [int(x) for x in up(avfids)]


def main():
    be, local, rank, world, _ = setup()
    on = lambda aid: vh.process_nr(aid) == rank   # noqa: E731  (@onrankof)

    sim, a1, a2, a3, avids, avfids = createsim(be, local)
    # ids after the hand-out: equal blocks per type in id order (3 AMortal agents, 10 of the others)
    b3, b10 = vh.equal_partition(3, world), vh.equal_partition(10, world)
    for k, aid in enumerate([a1, a2, a3]):
        owner = int(np.searchsorted(np.array(b3[1:]), k, side="right"))
        assert aid == vh.agent_id(1, owner, k - b3[owner] + 1)
    for k, aid in enumerate(avids):
        owner = int(np.searchsorted(np.array(b10[1:]), k, side="right"))
        assert vh.process_nr(aid) == owner and vh.agent_nr(aid) == k - b10[owner] + 1
    # agentstate (core.jl:161-175)
    sim.disable_transition_checks(True)
    if on(a1):
        assert sim.agentstate(a1, "AMortal")["foo"] == 1 and sim.agentstate_flexible(a1)["foo"] == 1
    if on(a2):
        assert sim.agentstate(a2, "AMortal")["foo"] == 2
    for k in (0, 9):
        if on(avids[k]):
            assert sim.agentstate(avids[k], "AImm")["foo"] == k + 1
        if on(avfids[k]):
            assert sim.agentstate(avfids[k], "AImmFixed")["foo"] == k + 1
    # edges & neighborids (core.jl:177-199): the sources are agents of other ranks now, the order is the insertion order
    if on(a1):
        es = sim.edges(a1, "ESDict")
        assert [int(f) for f, _ in es] == [a2, a3, avids[0], avfids[9]], ([hex(int(f)) for f, _ in es], [hex(x) for x in (a2, a3, avids[0], avfids[9])])
        assert [int(s["foo"]) for _, s in es] == [1, 2, 3, 4]
        assert [int(x) for x in sim.neighborids(a1, "ESLDict1")] == avids and sim.num_edges(a1, "ESDict") == 4
        # (core.jl:201-208 reads neighborstates outside of a transition; the states of agents of other ranks are only transferred by the
        #  halo of an apply! that reads their type, so that check is left to the neighbour sums below)
    if on(a2):
        assert sim.edges(a2, "ESDict") is None and sim.neighborids(a2, "ESLDict1") is None and sim.num_edges(a2, "ESDict") == 0
    if on(avids[9]):
        assert [int(x) for x in sim.neighborids(avids[9], "ESLDict2")] == [avfids[9]]
    sim.disable_transition_checks(False)
    # all_agents & num_agents (core.jl:210-241): collective, joined over the ranks
    assert sim.num_agents("AMortalFixed") == 10 and sim.num_agents("AImm") == 10 and sim.num_agents("AMortal") == 3
    assert sorted(sim.all_agents("AMortalFixed")["foo"].tolist()) == list(range(1, 11))
    assert len(sim.all_agents("AImm", all_ranks=False)) == b10[rank + 1] - b10[rank]
    assert sorted(sim.all_agentids("AImm").tolist()) == sorted(avids)
    assert sim.num_edges("ESDict") == 4 and sim.num_edges("ESLDict1") == 10 and sim.num_edges("ESLDict2") == 10
    sim.apply("keep_foo_lt6", "AMortalFixed", "AMortalFixed", "AMortalFixed")
    assert sim.num_agents("AMortalFixed") == 5 and sorted(sim.all_agents("AMortalFixed")["foo"].tolist()) == [1, 2, 3, 4, 5]
    # mapreduce (core.jl:87-94,396-440): folds of 1..10 over the ranks
    r = list(range(1, 11))
    for T in ["AImmFixed", "AImmFixedOversize", "ADefault"]:
        assert sim.mapreduce("foo", "+", T) == sum(r) and sim.mapreduce("foo", "*", T) == reduce(operator.mul, r)
        assert sim.mapreduce("foo", "&", T, datatype="i8") == reduce(operator.and_, r) and sim.mapreduce("foo", "|", T, datatype="i8") == reduce(operator.or_, r)
        assert sim.mapreduce("foo", "max", T) == 10 and sim.mapreduce("foo", "min", T) == 1
    assert sim.mapreduce("bool", "&", "ADefault") is True
    sim.apply("set_bool_id_odd", ["ADefault"], ["ADefault"], ["ADefault"])
    assert sim.mapreduce("bool", "&", "ADefault") is False and sim.mapreduce("bool", "|", "ADefault") is True

    # neighbour sums through the halo (core.jl:299-393): 56 then 111; 2, 3, 4; 3, 5, 7
    def sums(loops, tname, etype, target_of, expects):
        sim, a1, a2, a3, avids, avfids = createsim(be, local, loops)
        target = target_of(a1, avids, avfids)
        for expect in expects:
            sim.apply(f"sum_state_neighbors_{etype}", [tname], ALLAGENTTYPES + [etype], [tname])
            sim.disable_transition_checks(True)
            if on(target):
                assert sim.agentstate(target, tname)["foo"] == expect, (tname, expect, sim.agentstate(target, tname)["foo"])
            sim.disable_transition_checks(False)
    sums(("a1",), "AMortal", "ESLDict1", lambda a1, av, avf: a1, (sum(r) + 1, 2 * sum(r) + 1))
    if world > 1:
        # agentstate of an agent of another rank (the reference's `foreignstate`, src/AgentMethods.jl:103-111): answered from the ghost
        # slot that mirrors it, as of the last halo exchange of an apply! that read its type; an agent nobody here refers to is refused
        simf, a1f, _, _, avf_ids, avff_ids = createsim(be, local, ("a1",))
        simf.apply("sum_state_neighbors_ESLDict1", ["AMortal"], ALLAGENTTYPES + ["ESLDict1"], ["AMortal"])
        simf.disable_transition_checks(True)
        if on(a1f):
            far = avf_ids[9]                                     # a source of a1's row, owned by the last rank
            assert not on(far) and simf.agentstate(far, "AImm")["foo"] == 10
            lonely = avff_ids[5]                                 # its only edge points to avids[5], which does not live here either
            if not on(lonely) and not on(avf_ids[5]):
                try:
                    simf.agentstate(lonely, "AImmFixed")
                    raise SystemExit("expected an AssertionError for an agent of another rank that is not mirrored here")
                except AssertionError:
                    pass
        simf.disable_transition_checks(False)
    sums(("avids",), "AImm", "ESLDict2", lambda a1, av, avf: av[0], (2, 3, 4))
    sums(("avfids",), "AImmFixed", "ESLDict1", lambda a1, av, avf: avf[0], (3, 5, 7))
    # add_agent_per_process! (core.jl:442-466): one new agent on every rank
    sim = vh.create_simulation(core_model(), backend=be, device=local)
    for i in range(1, 11):
        sim.add_agent("AMortal", i)
    sim.finish_init(partition_algo="EqualAgentNumbers")
    sim.add_agent_per_process("AMortal", 100)
    assert sim.num_agents("AMortal") == 10 + world
    print(f"rank {rank}/{world}: ok", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
