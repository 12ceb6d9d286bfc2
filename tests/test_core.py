"""Replays /root/reference/test/core.jl (single process) against the oracle and, marked gpu, the CUDA engine."""
from functools import reduce
import operator

import numpy as np
import pytest

import vahana_b200 as vh
from models import core_model, createsim, add_example_network, foos

ALLAGENTTYPES = ["AMortal", "AImm", "AImmFixed"]


def test_model_immortal():  # test/core.jl:124-131
    assert core_model().immortal == [False, False, True, True, True, False]


def test_with_edge(backend):  # test/core.jl:133-157
    sim = vh.create_simulation(core_model(), backend=backend)
    aids = [sim.add_agent("AMortal", i) for i in range(1, 101)]
    sim.add_edge(aids[0], aids[1], "ESLDict1")
    sim.finish_init()
    assert sim.num_agents("AMortal") == 100
    sim2 = sim.copy_simulation()
    sim.apply("kill_all", "AMortal", "AMortal", "AMortal", with_edge="ESLDict1")
    assert sim.num_agents("AMortal") == 99
    sim2.apply("kill_all", "AMortal", [], "AMortal", with_edge="ESLDict1")
    assert sim2.num_agents("AMortal") == 99


def test_agentstate(backend):  # test/core.jl:161-175
    sim, a1, a2, a3, avids, avfids = createsim(backend)
    with pytest.raises(AssertionError):
        sim.agentstate(a2, "AImm")
    with pytest.raises(AssertionError):
        sim.agentstate(avids[0], "AMortal")
    sim.disable_transition_checks(True)
    assert sim.agentstate(a1, "AMortal")["foo"] == 1
    assert sim.agentstate_flexible(a1)["foo"] == 1
    assert sim.agentstate(a2, "AMortal")["foo"] == 2
    assert sim.agentstate(avids[0], "AImm")["foo"] == 1
    assert sim.agentstate(avfids[0], "AImmFixed")["foo"] == 1
    assert sim.agentstate(avids[9], "AImm")["foo"] == 10
    assert sim.agentstate(avfids[9], "AImmFixed")["foo"] == 10
    sim.disable_transition_checks(False)


def test_edges_and_neighborids(backend):  # test/core.jl:177-199
    sim, a1, a2, a3, avids, avfids = createsim(backend)
    sim.disable_transition_checks(True)
    assert len(sim.edges(a1, "ESDict")) == 4
    assert len(sim.neighborids(a1, "ESLDict1")) == 10
    assert sim.edges(a2, "ESDict") is None
    assert sim.neighborids(a2, "ESLDict1") is None
    assert len(sim.neighborids(avids[0], "ESLDict2")) == 1
    assert len(sim.neighborids(avids[9], "ESLDict2")) == 1
    assert len(sim.neighborids_iter(avids[0], "ESLDict2")) == 1
    es = sim.neighborids(a1, "ESLDict1")
    assert es[0] == avids[0] and es[9] == avids[9]
    assert sim.neighborids(avids[9], "ESLDict2")[0] == avfids[9]
    # insertion order of the stateful edges is pinned
    es = sim.edges(a1, "ESDict")
    assert [f for f, _ in es] == [a2, a3, avids[0], avfids[9]]
    assert [int(s["foo"]) for _, s in es] == [1, 2, 3, 4]
    sim.disable_transition_checks(False)


def test_neighborstates(backend):  # test/core.jl:201-208
    sim, a1, a2, a3, avids, avfids = createsim(backend)
    sim.disable_transition_checks(True)
    assert 1 in [int(s["foo"]) for s in sim.neighborstates(a1, "ESLDict1", "AImm")]
    assert 3 in [int(s["foo"]) for s in sim.neighborstates_flexible(a1, "ESDict")]
    sim.disable_transition_checks(False)


def test_all_agents_num_agents(backend):  # test/core.jl:210-241
    sim, *_ = createsim(backend)
    copy = sim.copy_simulation()
    assert len(copy.all_agents("AMortalFixed")) == 10
    assert len(copy.all_agents("AImm")) == 10
    assert copy.num_agents("AMortalFixed") == 10 and copy.num_agents("AImm") == 10
    assert sorted(copy.all_agents("AMortalFixed")["foo"].tolist()) == list(range(1, 11))
    assert sorted(copy.all_agents("AImm")["foo"].tolist()) == list(range(1, 11))
    copy.apply("keep_foo_lt6", "AMortalFixed", "AMortalFixed", "AMortalFixed")
    assert len(copy.all_agents("AMortalFixed")) == 5
    assert copy.num_agents("AMortalFixed") == 5
    assert sorted(copy.all_agents("AMortalFixed")["foo"].tolist()) == [1, 2, 3, 4, 5]
    # the original is untouched
    assert sim.num_agents("AMortalFixed") == 10


def test_all_agentids(backend):  # test/core.jl:243-277
    t = vh.ModelTypes().register_agenttype("AMortal", [("foo", "i8")])
    sim = vh.create_simulation(vh.create_model(t, "all_agentids_test"), backend=backend)
    for i in range(1, 11):
        sim.add_agent("AMortal", i)
    sim.finish_init()
    assert len(sim.all_agentids("AMortal")) == 10
    sim.apply("keep_even_foo", "AMortal", "AMortal", "AMortal")
    ids = sim.all_agentids("AMortal")
    assert len(ids) == 5
    assert [vh.agent_nr(i) for i in ids] == [2, 4, 6, 8, 10]
    t = vh.ModelTypes().register_agenttype("AImm", [("foo", "i8")], "Immortal")
    sim = vh.create_simulation(vh.create_model(t, "all_agentids_test_immortal"), backend=backend)
    for i in range(1, 11):
        sim.add_agent("AImm", i)
    sim.finish_init()
    assert len(sim.all_agentids("AImm")) == 10


def test_num_edges(backend):  # test/core.jl:279-285
    sim, a1, a2, a3, avids, avfids = createsim(backend)
    sim.disable_transition_checks(True)
    assert sim.num_edges(a1, "ESDict") == 4
    assert sim.num_edges(a2, "ESDict") == 0
    assert sim.num_edges(avids[0], "ESLDict2") == 1
    sim.disable_transition_checks(False)


def test_transition_missing_read_asserts(backend):  # test/core.jl:290-297
    sim, *_ = createsim(backend)
    with pytest.raises(AssertionError):   # AImmFixed is missing in `read`
        sim.apply("sum_state_neighbors_ESLDict2", "AImm", ["ESLDict2"], [])


def test_transition_sums(backend):  # test/core.jl:299-393: 56 then 111; 2,3,4; 3,5,7
    sim, a1, a2, a3, avids, avfids = createsim(backend)
    sim.disable_transition_checks(True)
    sim.add_edge(a1, a1, "ESLDict1")
    sim.add_edge(avids[0], avids[0], "ESLDict2")
    sim.add_edge(avfids[0], avfids[0], "ESLDict1")
    sim.add_edge(avfids[1], avfids[0], "ESLDict1")
    sim.add_edge(avfids[1], avfids[1], "ESLDict1")
    sim.disable_transition_checks(False)
    for expect in (sum(range(1, 11)) + 1, 2 * sum(range(1, 11)) + 1):
        sim.apply("sum_state_neighbors_ESLDict1", ["AMortal"], ALLAGENTTYPES + ["ESLDict1"], ["AMortal"])
        sim.disable_transition_checks(True)
        assert sim.agentstate(a1, "AMortal")["foo"] == expect
        sim.disable_transition_checks(False)

    sim, a1, a2, a3, avids, avfids = createsim(backend)
    sim.disable_transition_checks(True)
    sim.add_edge(avids[0], avids[0], "ESLDict2")
    sim.disable_transition_checks(False)
    for expect in (2, 3, 4):
        sim.apply("sum_state_neighbors_ESLDict2", ["AImm"], ALLAGENTTYPES + ["ESLDict2"], ["AImm"])
        sim.disable_transition_checks(True)
        assert sim.agentstate(avids[0], "AImm")["foo"] == expect
        sim.disable_transition_checks(False)

    sim, a1, a2, a3, avids, avfids = createsim(backend)
    sim.disable_transition_checks(True)
    sim.add_edge(avfids[0], avfids[0], "ESLDict1")
    sim.add_edge(avfids[1], avfids[0], "ESLDict1")
    sim.add_edge(avfids[1], avfids[1], "ESLDict1")
    sim.disable_transition_checks(False)
    for expect in (3, 5, 7):
        sim.apply("sum_state_neighbors_ESLDict1", ["AImmFixed"], ALLAGENTTYPES + ["ESLDict1"], ["AImmFixed"])
        sim.disable_transition_checks(True)
        assert sim.agentstate(avfids[0], "AImmFixed")["foo"] == expect
        sim.disable_transition_checks(False)


def _test_aggregate(sim, T):  # test/core.jl:87-94
    r = list(range(1, 11))
    assert sim.mapreduce("foo", "+", T) == reduce(operator.add, r)
    assert sim.mapreduce("foo", "*", T) == reduce(operator.mul, r)
    assert sim.mapreduce("foo", "&", T, datatype="i8") == reduce(operator.and_, r)
    assert sim.mapreduce("foo", "|", T, datatype="i8") == reduce(operator.or_, r)
    assert sim.mapreduce("foo", "max", T) == 10
    assert sim.mapreduce("foo", "min", T) == 1


def test_mapreduce(backend):  # test/core.jl:396-440
    sim = vh.create_simulation(core_model(), backend=backend)
    add_example_network(sim)
    sim.finish_init()
    for T in ["AImmFixed", "AImmFixedOversize"]:
        _test_aggregate(sim, T)
    for T in ["AMortalFixed", "ADefault"]:
        _test_aggregate(sim, T)
        sim.apply("identity", [T], [T], [T])
        sim.mapreduce("foo", "+", T)
    assert sim.mapreduce("bool", "&", "ADefault") is True
    assert sim.mapreduce("bool", "|", "ADefault") is True
    sim.apply("set_bool_false", ["ADefault"], ["ADefault"], ["ADefault"])
    assert sim.mapreduce("bool", "&", "ADefault") is False
    assert sim.mapreduce("bool", "|", "ADefault") is False
    sim.apply("set_bool_id_odd", ["ADefault"], ["ADefault"], ["ADefault"])
    assert sim.mapreduce("bool", "&", "ADefault") is False
    assert sim.mapreduce("bool", "|", "ADefault") is True


def test_add_agent_per_process(backend):  # test/core.jl:442-466
    sim = vh.create_simulation(core_model(), backend=backend)
    for i in range(1, 11):
        sim.add_agent("AMortal", i)
    sim.finish_init()
    new_id = sim.add_agent_per_process("AMortal", 100)
    assert sim.num_agents("AMortal") == 10 + 1          # + mpi.size
    assert vh.agent_nr(new_id) == 11
    sim.disable_transition_checks(True)
    assert sim.agentstate(new_id, "AMortal")["foo"] == 100
    sim.disable_transition_checks(False)
    sim.apply("keep_even_foo", "AMortal", "AMortal", "AMortal")      # 2,4,6,8,10,100 survive
    assert sim.num_agents("AMortal") == 6
    again = sim.add_agent_per_process("AMortal", 7)                   # reuses the most recently freed slot
    assert vh.agent_nr(again) == 9
    assert sorted(sim.all_agents("AMortal")["foo"].tolist()) == [2, 4, 6, 7, 8, 10, 100]
