"""Model definitions of the reference's test-suite, restated with the Python mirror."""
import numpy as np
import vahana_b200 as vh

FOO = [("foo", "i8")]
FOOBOOL = [("foo", "i8"), ("bool", "?")]


def core_model():
    """test/core.jl:28-38"""
    t = vh.ModelTypes()
    t.register_agenttype("AMortal", FOO)
    t.register_agenttype("AMortalFixed", FOO)
    t.register_agenttype("AImm", FOO, "Immortal")
    t.register_agenttype("AImmFixed", FOO, "Immortal")
    t.register_agenttype("AImmFixedOversize", FOO, "Immortal")
    t.register_agenttype("ADefault", FOOBOOL)
    t.register_edgetype("ESDict", FOO)
    t.register_edgetype("ESLDict1")
    t.register_edgetype("ESLDict2", None, "SingleType", target="AImm")
    return vh.create_model(t, "Test Core")


def foos(vals):
    return np.array([(v,) for v in vals], dtype=np.dtype(FOO, align=True))


def add_example_network(sim):
    """test/core.jl:41-69"""
    a1 = sim.add_agent("AMortal", 1)
    a2, a3 = sim.add_agents("AMortal", foos([2, 3]))
    avids = sim.add_agents("AImm", foos(range(1, 11)))
    avfids = sim.add_agents("AImmFixed", foos(range(1, 11)))
    sim.add_agents("AImmFixedOversize", foos(range(1, 11)))
    sim.add_agents("AMortalFixed", foos(range(1, 11)))
    sim.add_agents("ADefault", np.array([(i, True) for i in range(1, 11)], dtype=np.dtype(FOOBOOL, align=True)))
    sim.add_edge(a2, a1, "ESDict", 1)
    sim.add_edge(a3, a1, "ESDict", 2)
    sim.add_edge(avids[0], a1, "ESDict", 3)
    sim.add_edge(avfids[9], a1, "ESDict", 4)
    sim.add_edges(avids, np.full(10, a1, dtype=np.uint64), "ESLDict1")
    for i in range(10):
        sim.add_edge(avfids[i], avids[i], "ESLDict2")
    return int(a1), int(a2), int(a3), [int(x) for x in avids], [int(x) for x in avfids]


def createsim(backend):
    sim = vh.create_simulation(core_model(), backend=backend)
    ids = add_example_network(sim)
    sim.finish_init()
    return (sim,) + ids


def hk_model():
    """docs/examples/hegselmann.jl:27-45"""
    t = vh.ModelTypes()
    t.register_agenttype("HKAgent", [("opinion", "f8")])
    t.register_edgetype("Knows")
    t.register_param("eps", 0.02)
    return vh.create_model(t, "Hegselmann-Krause")


def hk_sim(backend, n, edges_uv, opinions, eps=0.02):
    """add_graph! + one self loop per agent (hegselmann.jl:85-95)"""
    sim = vh.create_simulation(hk_model(), params={"eps": eps}, backend=backend)
    ids = vh.add_graph(sim, edges_uv, n, "HKAgent", np.asarray(opinions, dtype="f8").view([("opinion", "f8")]), "Knows")
    sim.add_edges(ids, ids, "Knows")
    sim.finish_init()
    return sim, ids


def ba_graph(n, m, seed):
    import networkx as nx
    g = nx.barabasi_albert_graph(n, m, seed=seed)
    return np.array(list(g.edges()), dtype=np.int64)
