"""Replays /root/reference/test/edges.jl: every accessor for every legal hint combination (13 + 5 with `size`)."""
import numpy as np
import pytest

import vahana_b200 as vh
from models import edges_model, hashint, EDGE_TYPES, STATELESS_EDGE_TYPES, STATEFUL_EDGE_TYPES, foos


def _build(backend):
    sim = vh.create_simulation(edges_model(), backend=backend)
    a1, a2, a3 = (int(x) for x in sim.add_agents("Agent", foos([1, 2, 3])))
    # 2 -> 3 (state 1), [1 -> 3 (state 2) where several edges are allowed], 3 -> 1 (state 3)   (test/edges.jl:75-117)
    for t in STATELESS_EDGE_TYPES + STATEFUL_EDGE_TYPES:
        st = t in STATEFUL_EDGE_TYPES
        sim.add_edge(a2, a3, t, 1 if st else None)
        if hashint(t, "E") and not hashint(t, "T") and not (hashint(t, "S") and hashint(t, "I")):
            with pytest.raises(AssertionError):
                sim.add_edge(a1, a3, t, 2 if st else None)
        elif not hashint(t, "E"):
            sim.add_edge(a1, a3, t, 2 if st else None)
        sim.add_edge(a3, a1, t, 3 if st else None)
    sim.finish_init()
    return sim, a1, a2, a3


def test_edges_accessor(backend):  # test/edges.jl:123-141
    sim, a1, a2, a3 = _build(backend)
    sim.disable_transition_checks(True)
    for t in ["EdgeD", "EdgeT", "EdgeTs"]:
        e = sim.edges(a3, t)
        assert e[0][0] == a2 and e[0][1]["foo"] == 1      # insertion order is pinned
        assert e[1][0] == a1 and e[1][1]["foo"] == 2
        assert sim.edges(a2, t) is None
    e = sim.edges(a3, "EdgeE")
    assert e[0] == a2 and e[1]["foo"] == 1
    sim.disable_transition_checks(False)
    for t in ["EdgeI", "EdgeEI", "EdgeTI", "EdgeTsI", "EdgeS", "EdgeSE", "EdgeST", "EdgeSI", "EdgeSEI", "EdgeSTI", "EdgeSETI", "EdgeSTs",
              "EdgeSTsI", "EdgeSETsI"]:
        with pytest.raises(AssertionError):
            sim.edges(a1, t)


def test_neighborids(backend):  # test/edges.jl:143-161
    sim, a1, a2, a3 = _build(backend)
    sim.disable_transition_checks(True)
    for t in ["EdgeD", "EdgeT", "EdgeTs", "EdgeS", "EdgeST", "EdgeSTs"]:
        assert sim.neighborids(a3, t) == [a2, a1]
        assert sim.neighborids(a2, t) is None
    for t in ["EdgeE", "EdgeSE"]:
        assert sim.neighborids(a3, t) == a2
    sim.disable_transition_checks(False)
    for t in ["EdgeI", "EdgeTI", "EdgeTsI", "EdgeSI", "EdgeSTI", "EdgeSTsI", "EdgeEI", "EdgeSEI", "EdgeSETI", "EdgeSETsI"]:
        with pytest.raises(AssertionError):
            sim.neighborids(a1, t)


def test_edgestates(backend):  # test/edges.jl:163-188
    sim, a1, a2, a3 = _build(backend)
    sim.disable_transition_checks(True)
    for t in ["EdgeD", "EdgeT", "EdgeTs", "EdgeI", "EdgeTI", "EdgeTsI"]:
        assert sim.edgestates(a3, t)["foo"].tolist() == [1, 2]
        assert sim.edgestates(a2, t) is None
        assert sim.edgestates_iter(a3, t)["foo"].tolist() == [1, 2]
        assert sim.edgestates_iter(a2, t) is None
    for t in ["EdgeE", "EdgeEI"]:
        assert sim.edgestates(a3, t)["foo"] == 1
    sim.disable_transition_checks(False)
    for t in ["EdgeS", "EdgeST", "EdgeSTs", "EdgeSI", "EdgeSTI", "EdgeSTsI", "EdgeSE", "EdgeSEI", "EdgeSETI", "EdgeSETsI"]:
        with pytest.raises(AssertionError):
            sim.edgestates(a1, t)
        with pytest.raises(AssertionError):
            sim.edgestates_iter(a1, t)


def test_num_edges_and_has_edge(backend):  # test/edges.jl:190-214
    sim, a1, a2, a3 = _build(backend)
    sim.disable_transition_checks(True)
    for t in ["EdgeD", "EdgeT", "EdgeTs", "EdgeI", "EdgeTI", "EdgeTsI", "EdgeS", "EdgeST", "EdgeSTs", "EdgeSI", "EdgeSTI", "EdgeSTsI"]:
        assert sim.num_edges(a1, t) == 1
        assert sim.num_edges(a2, t) == 0
        assert sim.num_edges(a3, t) == 2
    for t in EDGE_TYPES:
        assert sim.has_edge(a1, t) is True
        assert sim.has_edge(a2, t) is False
        assert sim.has_edge(a3, t) is True
    sim.disable_transition_checks(False)
    for t in ["EdgeE", "EdgeEI", "EdgeSE", "EdgeSETI", "EdgeSETsI"]:
        with pytest.raises(AssertionError):
            sim.num_edges(a1, t)


def test_edge_mapreduce(backend):  # test/edges.jl:216-227
    sim, *_ = _build(backend)
    for t in ["EdgeD", "EdgeT", "EdgeI", "EdgeTI", "EdgeTs", "EdgeTsI"]:
        assert sim.mapreduce("foo", "+", t) == 6
    for t in ["EdgeE", "EdgeEI"]:
        assert sim.mapreduce("foo", "+", t) == 4
    for t in STATELESS_EDGE_TYPES:
        with pytest.raises(AssertionError):
            sim.mapreduce("foo", "+", t)


def test_vahana_state_checks(backend):  # test/edges.jl:241-248
    sim, *_ = _build(backend)
    for t in STATELESS_EDGE_TYPES:
        with pytest.raises(AssertionError):
            sim.add_edge(0, 0, t)
    for t in STATEFUL_EDGE_TYPES:
        with pytest.raises(AssertionError):
            sim.add_edge(0, 0, t, 0)


def test_edgetype_in_read_and_write(backend):  # test/edges.jl:250-273
    sim, *_ = _build(backend)
    for i, t in enumerate(EDGE_TYPES):
        with pytest.raises(AssertionError):
            sim.copy_simulation().apply(f"add_self_loop_{i}", "Agent", [], [])
        sim.copy_simulation().apply(f"add_self_loop_{i}", "Agent", [], [t])
        with pytest.raises(AssertionError):
            sim.copy_simulation().apply(f"touch_edge_{i}", "Agent", [], [])
        sim.copy_simulation().apply(f"touch_edge_{i}", "Agent", [t], [])


def test_single_edge_conflict_with_add_existing(backend):  # _can_add, src/EdgeMethods.jl:267-293
    """With `add_existing` the write container already holds the target's edge: adding a different one asserts (Dict containers,
    i.e. :SingleEdge without :SingleType), adding the identical one is allowed."""
    for t in ["EdgeE", "EdgeSE", "EdgeEI"]:
        i = EDGE_TYPES.index(t)
        sim, a1, a2, a3 = _build(backend)          # a3 <- a2 and a1 <- a3 exist; the self loops differ from both
        if t == "EdgeEI":                          # :IgnoreFrom: the value is the state alone; 0 differs from the stored 1 / 3
            with pytest.raises(AssertionError):
                sim.apply(f"add_self_loop_{i}", "Agent", [], [t], add_existing=[t])
        else:
            with pytest.raises(AssertionError):
                sim.apply(f"add_self_loop_{i}", "Agent", [], [t], add_existing=[t])
        # without add_existing the write container starts empty: every agent simply gets its self loop
        sim2, *_ = _build(backend)
        sim2.apply(f"add_self_loop_{i}", "Agent", [], [t])
        assert sim2.num_edges(t) == 3


def test_num_edges_write_flag(backend):  # test/edges.jl:278-316
    sim = vh.create_simulation(edges_model(), backend=backend)
    for t in EDGE_TYPES:
        assert sim.num_edges(t, write=False) == 0
    assert sim.num_agents("Agent") == 0
    id1, id2, id3 = (sim.add_agent("Agent", 0) for _ in range(3))
    assert sim.num_agents("Agent") == 3
    for t in EDGE_TYPES:
        st = 0 if t in STATEFUL_EDGE_TYPES else None
        sim.add_edge(id1, id1, t, st)
        sim.add_edge(id3, id3, t, st)
    for t in EDGE_TYPES:
        assert sim.num_edges(t, write=True) == 2
        assert sim.num_edges(t, write=False) == 0
    sim.finish_init()
    for t in EDGE_TYPES:
        assert sim.num_edges(t, write=False) == 2


@pytest.mark.parametrize("ET", ["EdgeD", "EdgeT", "EdgeI", "EdgeTI"])
def test_transition(backend, ET):  # test/edges.jl:319-385
    nagents = 6
    sim = vh.create_simulation(edges_model(), backend=backend)
    uv = np.array([(i, j) for i in range(nagents) for j in range(i + 1, nagents)])   # complete_graph edge order
    ids = sim.add_agents("Agent", foos(range(1, nagents + 1)))
    fr = np.stack([ids[uv[:, 0]], ids[uv[:, 1]]], axis=1).reshape(-1)
    to = np.stack([ids[uv[:, 1]], ids[uv[:, 0]]], axis=1).reshape(-1)
    st = np.stack([uv[:, 1] + 1, uv[:, 1] + 1], axis=1).reshape(-1)                  # e -> ET(e.dst)
    sim.add_edges(fr, to, ET, foos(st))
    sim.finish_init()

    def check(expected):
        sim.apply(f"store_num_edges_{ET}", ["Agent"], [ET], ["Agent"])
        assert sim.all_agents("Agent")["foo"].tolist() == [expected] * nagents

    check(nagents - 1)
    check(nagents - 1)   # write was empty: same result again
    sim.apply("identity", ["Agent"], [], [ET], add_existing=[ET])          # copy the edges
    check(nagents - 1)
    if ET == "EdgeD":
        before = sim.export_csr(ET, "Agent", nagents)
        sim.apply("readd_edges_EdgeD", ["Agent"], [ET], [ET])              # rewrite by re-adding
        check(nagents - 1)
        after = sim.export_csr(ET, "Agent", nagents)
        for b, a in zip(before, after):                                    # same rows, same order, same states
            assert np.array_equal(b, a)
        sim.apply("identity", ["Agent"], [], [])
        sim.apply("readd_edges_EdgeD", ["Agent"], [ET], [ET], add_existing=[ET])   # copy and re-add: twice now
        check((nagents - 1) * 2)
        sim.apply("identity", ["Agent"], [], [ET])                         # nobody adds them: gone
        check(0)
