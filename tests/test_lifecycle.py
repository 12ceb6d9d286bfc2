"""Replays /root/reference/test/remove_agents.jl, addexisting.jl, independent.jl, graphs.jl and the aggregate /
dead-agent parts of edgesiterator.jl against the oracle and (marked gpu) the CUDA engine."""
import numpy as np
import pytest

import vahana_b200 as vh
from models import (remove_agents_model, addexisting_model, independent_model, graph_model, edges_model, foos,
                    STATEFUL_EDGE_TYPES, STATELESS_EDGE_TYPES)

P = 1   # mpi.size of the single-process run


@pytest.mark.parametrize("E", ["DEdge", "DEdgeST"])
def test_dying_agents(backend, E):  # test/remove_agents.jl:23-119
    for twice in (False, True):
        sim = vh.create_simulation(remove_agents_model(), backend=backend)
        ids = sim.add_agents("DAgent", np.array([(i,) for i in range(1, P * 3 + 1)], dtype=[("idx", "i8")]))
        rids = sim.add_agents("DAgentRemove", None, P)
        for _ in range(2 if twice else 1):
            for i in ids:
                sim.add_edge(i, ids[1], E)
            for i in rids:
                sim.add_edge(i, ids[1], "DEdgeState", 0)
        for _ in range(2 if twice else 1):
            sim.add_edge(ids[1], ids[2], E)
        sim.finish_init()
        m = 2 if twice else 1
        assert sim.num_edges(E) == (P * 3 + 1) * m
        assert sim.num_edges("DEdgeState") == P * m
        sim.apply("kill_all", ["DAgentRemove"], [], ["DAgentRemove"])
        assert sim.num_edges(E) == (P * 3 + 1) * m
        assert sim.num_edges("DEdgeState") == 0             # edges from the removed agents are purged
        sim.apply(f"die_if_no_edges_{E}", ["DAgent"], ["DAgent", E], ["DAgent"])
        assert sim.num_edges(E) == 3 * m
        assert sim.num_agents("DAgent") == 2


def test_add_existing(backend):  # test/addexisting.jl:9-66
    for model in (addexisting_model(), addexisting_model(compute_hints=("Immortal",))):
        sim = vh.create_simulation(model, backend=backend)
        computeid = sim.add_agent("ComputeAgent")
        constructedid = sim.add_agent("ConstructedAgent")
        sim.add_edge(constructedid, computeid, "Connection")
        sim.finish_init()
        sim.apply("noop", "ComputeAgent", [], "Connection", add_existing="Connection")
        assert sim.num_edges("Connection") == 1
        sim.apply("noop", ["ComputeAgent"], [], ["ConstructedAgent", "Connection"])
        assert sim.num_agents("ComputeAgent") == 1
        assert sim.num_agents("ConstructedAgent") == 0
        assert sim.num_edges("Connection") == 0
        sim.apply("construct_and_connect", ["ComputeAgent"], [], ["ConstructedAgent", "Connection"])
        assert sim.num_agents("ComputeAgent") == 1
        assert sim.num_agents("ConstructedAgent") == 1
        assert sim.num_edges("Connection") == 1
        sim.disable_transition_checks(True)
        assert sim.neighborids(computeid, "Connection") == [vh.agent_id(2, 0, 1)]
        sim.disable_transition_checks(False)
        sim.apply("noop", ["ComputeAgent"], [], ["ConstructedAgent", "Connection"], add_existing=["ConstructedAgent", "Connection"])
        assert sim.num_agents("ConstructedAgent") == 1
        assert sim.num_agents("ComputeAgent") == 1
        assert sim.num_edges("Connection") == 1


def test_add_existing_immortal_assertion(backend):  # test/addexisting.jl:68-85
    sim = vh.create_simulation(addexisting_model(constructed_hints=("Immortal",)), backend=backend)
    computeid = sim.add_agent("ComputeAgent")
    constructedid = sim.add_agent("ConstructedAgent")
    sim.add_edge(constructedid, computeid, "Connection")
    sim.finish_init()
    with pytest.raises(AssertionError):
        sim.apply("noop", ["ComputeAgent"], [], ["ConstructedAgent", "Connection"])


def test_independent(backend):  # test/independent.jl: :Independent must be observationally equal to the default
    sims = {}
    for T in ["AIndependent", "ANotIndependent", "AIndependentImmortal"]:
        sim = vh.create_simulation(independent_model(), backend=backend)
        ids = [sim.add_agent(T, i) for i in range(1, 11)]
        for fr in ids:
            for to in ids:
                sim.add_edge(fr, to, "AEdge")
        sim.finish_init()
        assert sim.num_agents(T) == 10 and sim.num_edges("AEdge") == 100
        sims[T] = sim

    def snapshot(T):
        s = sims[T]
        to, fr, st = s.all_edges("AFooEdge")
        return (s.num_agents(T), s.num_edges("AEdge"), s.num_edges("AFooEdge"), sorted(s.all_agents(T)["foo"].tolist()),
                [vh.agent_nr(x) for x in s.all_agentids(T)], [vh.agent_nr(x) for x in to], [vh.agent_nr(x) for x in fr])

    for tr, kw in [("indep_step_1", {}), ("indep_step_2", {}), ("indep_spawn", {"add_existing": "AFooEdge"})]:
        for T, s in sims.items():
            s.apply(tr, T, [T, "AEdge"], [T, "AFooEdge"], **kw)
        ref = snapshot("ANotIndependent")
        assert snapshot("AIndependent") == ref
        assert snapshot("AIndependentImmortal") == ref
    n, ne, nf, foosl, nrs, *_ = ref
    assert n == 8 + 16 and ne == 64          # 2 died, each of the 8 survivors spawned 2; their AEdge rows/entries were purged
    assert nrs[:2] == [1, 2]                 # the two freed slots were reused first (LIFO: slot 2, then slot 1)
    assert foosl.count(13) == 1 and foosl.count(23) == 1


def test_graphs(backend):  # test/graphs.jl:17-45
    nagents = 4
    sim = vh.create_simulation(graph_model(), backend=backend)
    uv = np.array([(i, j) for i in range(nagents) for j in range(i + 1, nagents)])
    states = np.array([(i, 0) for i in range(1, nagents + 1)], dtype=[("id", "i8"), ("sum_ids_neighbors", "i8")])
    vh.add_graph(sim, uv, nagents, "GraphA", states, "GraphE")
    sim.finish_init()
    sim.apply("sumids", ["GraphA"], ["GraphA", "GraphE"], ["GraphA"])
    assert sim.mapreduce("sum_ids_neighbors", "+", "GraphA") == sum(range(1, nagents + 1)) * (nagents - 1)


@pytest.mark.gpu
@pytest.mark.parametrize("nagents,block_mb", [(4, 0.000016), (300, 0.0004), (1500, 0.003)])
def test_graphs_source_blocked(cuda, nagents, block_mb):
    """test/graphs.jl:37-43 (complete graph, sum of neighbour ids == sum(1:n) * (n - 1)) with the read phase forced into per-source-block
    sweeps: the reference's golden value, bit-exact (integer fold), with rows of up to n - 1 entries spread over many blocks."""
    sim = vh.create_simulation(graph_model(), backend=cuda)
    uv = np.array([(i, j) for i in range(nagents) for j in range(i + 1, nagents)])
    states = np.array([(i, 0) for i in range(1, nagents + 1)], dtype=[("id", "i8"), ("sum_ids_neighbors", "i8")])
    vh.add_graph(sim, uv, nagents, "GraphA", states, "GraphE")
    sim.finish_init()
    sim.set_read_blocking(block_mb, 0.0, 1)
    for _ in range(2):
        sim.apply("sumids", ["GraphA"], ["GraphA", "GraphE"], ["GraphA"])
        assert sim.last_apply_stats()["source_blocks"] >= 2
        assert sim.mapreduce("sum_ids_neighbors", "+", "GraphA") == sum(range(1, nagents + 1)) * (nagents - 1)
        got = sim.all_agents("GraphA")
        assert np.array_equal(got["sum_ids_neighbors"], sum(range(1, nagents + 1)) - got["id"])


@pytest.mark.parametrize("ET", [t for t in STATEFUL_EDGE_TYPES + STATELESS_EDGE_TYPES if not ("S" in t[4:] and "I" in t[4:])])
def test_edges_iterator_lengths(oracle, ET):  # test/edgesiterator.jl:7-53 ("Edges Iter": every stored edge is visited once, write and read side)
    # (oracle only: written after this round's GPU budget was spent; the engine's all_edges / num_edges run in the tests around it)
    S, E1 = ("S" in ET[4:]), ("E" in ET[4:])
    sim = vh.create_simulation(edges_model(), backend=oracle)
    assert sim.num_edges(ET) == 0 and sim.num_edges(ET, write=True) == 0 and len(sim.all_edges(ET)[0]) == 0
    aids = sim.add_agents("Agent", foos(range(1, 11)))
    for i in aids:
        sim.add_edge(aids[0], i, ET, None if S else vh.agent_nr(i))
        if not E1:
            sim.add_edge(i, aids[0], ET, None if S else vh.agent_nr(i))
    expected = 10 if E1 else 20
    assert sim.num_edges(ET, write=True) == expected
    sim.finish_init()
    to, fr, st = sim.all_edges(ET)
    assert sim.num_edges(ET) == expected and len(to) == expected
    if "I" not in ET[4:]:
        pairs = sorted(zip((int(x) for x in to), (int(x) for x in fr)))
        want = [(int(i), int(aids[0])) for i in aids] + ([] if E1 else [(int(aids[0]), int(i)) for i in aids])
        assert pairs == sorted(want)
    if not S:
        assert sorted(int(x) for x in st["foo"]) == sorted(list(range(1, 11)) * (1 if E1 else 2))


@pytest.mark.parametrize("ET", STATEFUL_EDGE_TYPES)
def test_edges_aggregate(backend, ET):  # test/edgesiterator.jl:59-98
    sim = vh.create_simulation(edges_model(), backend=backend)
    aids = sim.add_agents("Agent", foos(range(1, 11)))
    for i in aids:
        sim.add_edge(aids[0], i, ET, vh.agent_nr(i))
    sim.finish_init()
    assert sim.mapreduce("foo", "+", ET) == 55
    sim.apply("kill_all", ["Agent"], [], ["Agent"])
    assert sim.mapreduce("foo", "+", ET) == 0
    sim = vh.create_simulation(edges_model(), backend=backend)
    aids = sim.add_agents("Agent", foos(range(1, 11)))
    bids = sim.add_agents("AgentB", foos(range(1, 11)))
    for i in range(10):
        sim.add_edge(bids[i], aids[i], ET, i + 1)
    sim.finish_init()
    assert sim.mapreduce("foo", "+", ET) == 55
    sim.apply("kill_all", ["Agent"], [], ["Agent"])
    assert sim.mapreduce("foo", "+", ET) == 0


@pytest.mark.parametrize("ET", STATEFUL_EDGE_TYPES + STATELESS_EDGE_TYPES)
def test_remove_edges_of_dead_agents(backend, ET):  # test/edgesiterator.jl:100-215
    S, E1, T, I = ("S" in ET[4:]), ("E" in ET[4:]), ("T" in ET[4:]), ("I" in ET[4:])
    sim = vh.create_simulation(edges_model(), backend=backend)
    aids = sim.add_agents("Agent", foos(range(1, 11)))
    bids = sim.add_agents("AgentB", foos(range(1, 11)))
    for i in range(10):
        sim.add_edge(bids[i], aids[i], ET, None if S else i + 1)
    sim.finish_init()
    assert sim.num_edges(ET) == 10
    sim.apply("kill_all", ["Agent"], [], ["Agent"])
    assert sim.num_edges(ET) == 0
    if I:
        return
    sim = vh.create_simulation(edges_model(), backend=backend)
    aids = sim.add_agents("Agent", foos(range(1, 11)))
    bids = sim.add_agents("AgentB", foos(range(1, 11)))
    for i in range(10):
        st = None if S else i + 1
        if not E1:
            sim.add_edge(aids[i], aids[i], ET, st)
            if not T:
                sim.add_edge(aids[i], bids[i], ET, st)
        sim.add_edge(bids[i], aids[i], ET, st)
        if not T:
            sim.add_edge(bids[i], bids[i], ET, st)
    sim.finish_init()
    sim.apply("keep_even_foo", "AgentB", "AgentB", "AgentB")
    count = 5
    if not E1:
        count += 10
    if not T:
        count += 5
    if not E1 and not T:
        count += 5
    assert sim.num_edges(ET) == count
