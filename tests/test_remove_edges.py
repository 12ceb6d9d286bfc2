"""remove_edges! inside transitions: restates the single-process part of /root/reference/test/mpi/test_edgetypes.jl:295-449
(cycle graph / complete graph, counts 50, 100, 0, 9700, 9850, 9900) and the assertion rules of _can_remove_edges."""
import numpy as np
import pytest

import vahana_b200 as vh
from models import edges_model, foos

REMOVE_TYPES = ["EdgeD", "EdgeS", "EdgeE", "EdgeI", "EdgeSE", "EdgeSI", "EdgeEI", "EdgeSEI", "EdgeSTI", "EdgeSETI", "EdgeT", "EdgeST"]
REMOVE_FROM_TYPES = ["EdgeD", "EdgeS", "EdgeE", "EdgeSE", "EdgeT", "EdgeST"]


def _graph_sim(backend, ET, kind, n=100):
    sim = vh.create_simulation(edges_model(), backend=backend)
    ids = sim.add_agents("Agent", foos(range(1, n + 1)))
    stateful = "S" not in ET[4:]
    if kind == "cycle":     # cycle_digraph: i -> i + 1
        fr, to = ids, np.roll(ids, -1)
    else:                   # complete_graph via add_graph!: both directions per undirected edge, in edge order
        uv = np.array([(i, j) for i in range(n) for j in range(i + 1, n)])
        fr = np.stack([ids[uv[:, 0]], ids[uv[:, 1]]], axis=1).reshape(-1)
        to = np.stack([ids[uv[:, 1]], ids[uv[:, 0]]], axis=1).reshape(-1)
    sim.add_edges(fr, to, ET, foos(np.zeros(len(to), dtype=int)) if stateful else None)
    sim.finish_init()
    return sim


@pytest.mark.parametrize("ET", REMOVE_TYPES)
def test_remove_edges_to(backend, ET):   # test_edgetypes.jl:295-352
    sim = _graph_sim(backend, ET, "cycle")
    assert sim.num_edges(ET) == 100
    removed = sim.copy_simulation().apply(f"remove_own_if_even_{ET}", "Agent", ["Agent", ET], ET, add_existing=ET)
    assert removed.num_edges(ET) == 50
    if "I" not in ET[4:]:
        removed = sim.copy_simulation().apply(f"remove_neighbors_if_even_{ET}", "Agent", ["Agent", ET], ET, add_existing=ET)
        assert removed.num_edges(ET) == 50
        copy = sim.copy_simulation()
        copy.apply(f"remove_and_readd_{ET}", "Agent", ET, ET, add_existing=ET)      # removes are applied before the new edges
        assert copy.num_edges(ET) == 100
        copy.disable_transition_checks(True)
        ids = sim.all_agentids("Agent")
        assert copy.neighborids(int(ids[5]), ET) in ([int(ids[4])], int(ids[4]))
        copy.disable_transition_checks(False)


@pytest.mark.parametrize("ET", REMOVE_FROM_TYPES)
def test_remove_edges_from_to(backend, ET):   # test_edgetypes.jl:354-449
    single = "E" in ET[4:]
    sim = _graph_sim(backend, ET, "cycle" if single else "complete")
    assert sim.num_edges(ET) == (100 if single else 9900)
    sim2, sim3 = sim.copy_simulation(), sim.copy_simulation()
    sim.apply(f"remove_first_two_from_{ET}", "Agent", ["Agent", ET], ET, add_existing=ET)
    assert sim.num_edges(ET) == (0 if single else 9700)
    if not single:
        sim2.apply(f"remove_to_random_neighbor_if_even_{ET}", "Agent", ["Agent", ET], ET, add_existing=ET, seed=11)
        assert sim2.num_edges(ET) == 9850
    sim3.apply(f"remove_from_zero_{ET}", "Agent", [], ET, add_existing=ET)
    assert sim3.num_edges(ET) == (100 if single else 9900)
    with pytest.raises(AssertionError):   # not in write
        sim3.apply(f"remove_from_zero_{ET}", "Agent", [], [])
    with pytest.raises(AssertionError):   # in write but not in add_existing
        sim3.apply(f"remove_from_zero_{ET}", "Agent", [], ET)


@pytest.mark.gpu
@pytest.mark.parametrize("ET", ["EdgeD", "EdgeS", "EdgeT"])
def test_remove_edges_gpu_matches_oracle_exactly(oracle, cuda, ET):
    """beyond the counts: the surviving rows (order and states) are identical"""
    g, o = _graph_sim(cuda, ET, "complete", 40), _graph_sim(oracle, ET, "complete", 40)
    for sim in (g, o):
        sim.apply(f"remove_to_random_neighbor_if_even_{ET}", "Agent", ["Agent", ET], ET, add_existing=ET, seed=3)
        sim.apply(f"remove_first_two_from_{ET}", "Agent", ["Agent", ET], ET, add_existing=ET)
    a, b = g.export_csr(ET, "Agent", 40), o.export_csr(ET, "Agent", 40)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


@pytest.mark.parametrize("ET", ["EdgeD", "EdgeS", "EdgeT"])
def test_remove_other_row_then_add(oracle, ET):
    """a removal that names another agent's row, followed by an add to that row (the order rule of src/Simulation.jl:792-800):
    on the cycle i-1 -> i every row ends up with exactly the reversed edge.  (GPU variant: tests/test_zzn_remove_order.py)"""
    _check_reversed_cycle(_graph_sim(oracle, ET, "cycle"), ET)


def _check_reversed_cycle(sim, ET, n=100):
    ids = sim.all_agentids("Agent")
    sim.apply(f"clear_neighbor_row_and_point_back_{ET}", "Agent", ["Agent", ET], ET, add_existing=ET)
    assert sim.num_edges(ET) == n
    off, fr, st = sim.export_csr(ET, "Agent", n)
    assert np.array_equal(off, np.arange(n + 1))
    assert np.array_equal(fr, np.roll(ids, -1))                      # row i holds the edge from i + 1
    if st is not None:
        assert np.all(st["foo"] == 7)


def test_remove_edges_init_phase(backend):   # remove_edges! is allowed until finish_init! (EdgeMethods.jl:101-106)
    sim = vh.create_simulation(edges_model(), backend=backend)
    a, b, c = (int(x) for x in sim.add_agents("Agent", foos([1, 2, 3])))
    for t, st in (("EdgeD", 1), ("EdgeS", None), ("EdgeSI", None)):
        sim.add_edge(a, c, t, st)
        sim.add_edge(b, c, t, st)
        sim.add_edge(a, b, t, st)
    sim.remove_edges(a, c, "EdgeD")          # (from, to)
    sim.remove_edges(c, "EdgeS")             # whole row
    sim.remove_edges(c, "EdgeSI")
    sim.finish_init()
    assert sim.num_edges("EdgeD") == 2 and sim.num_edges("EdgeS") == 1 and sim.num_edges("EdgeSI") == 1
    sim.disable_transition_checks(True)
    assert sim.neighborids(c, "EdgeD") == [b]
    sim.remove_edges(b, "EdgeD")             # the "hack" of the reference's tests: outside of a transition with checks disabled
    assert sim.neighborids(b, "EdgeD") is None
    sim.disable_transition_checks(False)
    assert sim.num_edges("EdgeD") == 1
