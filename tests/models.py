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


# ---- test/edges.jl:13-63 ----
EDGE_TYPES = ["EdgeD", "EdgeS", "EdgeE", "EdgeT", "EdgeI", "EdgeSE", "EdgeST", "EdgeSI", "EdgeEI", "EdgeTI", "EdgeSEI", "EdgeSTI",
              "EdgeSETI", "EdgeTs", "EdgeTsI", "EdgeSTs", "EdgeSTsI", "EdgeSETsI"]
STATELESS_EDGE_TYPES = ["EdgeS", "EdgeSE", "EdgeST", "EdgeSI", "EdgeSEI", "EdgeSTI", "EdgeSETI", "EdgeSTs", "EdgeSTsI", "EdgeSETsI"]
STATEFUL_EDGE_TYPES = ["EdgeD", "EdgeE", "EdgeT", "EdgeI", "EdgeEI", "EdgeTI", "EdgeTs", "EdgeTsI"]


def hashint(name, hint):
    return hint in name[4:]


def edges_model():
    t = vh.ModelTypes()
    t.register_agenttype("Agent", FOO)
    t.register_agenttype("AgentB", FOO)
    hintmap = {"S": "Stateless", "E": "SingleEdge", "T": "SingleType", "I": "IgnoreFrom"}
    for name in EDGE_TYPES:
        code = name[4:]
        hints = [hintmap[c] for c in code if c in hintmap]
        kw = {}
        if "T" in code:
            kw["target"] = "Agent"
        if "s" in code:
            kw["size"] = 10
        t.register_edgetype(name, None if "S" in code else FOO, *hints, **kw)
    return vh.create_model(t, "Test Edges")


def remove_agents_model():
    """test/remove_agents.jl:13-20"""
    t = vh.ModelTypes()
    t.register_agenttype("DAgent", [("idx", "i8")])
    t.register_agenttype("DAgentRemove")
    t.register_edgetype("DEdge")
    t.register_edgetype("DEdgeState", [("state", "i8")])
    t.register_edgetype("DSingleEdge", None, "SingleEdge")
    t.register_edgetype("DEdgeST", None, "SingleType", target="DAgent")
    return vh.create_model(t, "remove_agents")


def addexisting_model(compute_hints=(), constructed_hints=()):
    """test/addexisting.jl:88-110 (detect_stateless(true) is active in that file)"""
    old = vh.config.detect_stateless
    vh.detect_stateless(True)
    try:
        t = vh.ModelTypes()
        t.register_agenttype("ComputeAgent", None, *compute_hints)
        t.register_agenttype("ConstructedAgent", None, *constructed_hints)
        t.register_edgetype("Connection")
    finally:
        vh.detect_stateless(old)
    return vh.create_model(t, "Test add_existing")


def independent_model():
    """test/independent.jl:24-30"""
    t = vh.ModelTypes()
    t.register_agenttype("AIndependent", FOO, "Independent")
    t.register_agenttype("ANotIndependent", FOO)
    t.register_agenttype("AIndependentImmortal", FOO, "Independent")
    t.register_edgetype("AFooEdge", FOO)
    t.register_edgetype("AEdge")
    return vh.create_model(t, "Test Independent")


def graph_model():
    """test/graphs.jl:11-14"""
    t = vh.ModelTypes()
    t.register_agenttype("GraphA", [("id", "i8"), ("sum_ids_neighbors", "i8")])
    t.register_edgetype("GraphE")
    return vh.create_model(t, "Test Graph")


GRIDA = [("pos", "i8", (2,)), ("active", "?")]
GRID3D = [("pos", "i8", (3,)), ("active", "?")]


def raster_model():
    """test/raster.jl:22-29"""
    t = vh.ModelTypes()
    t.register_agenttype("GridA", GRIDA)
    t.register_agenttype("Grid3D", GRID3D)
    t.register_edgetype("GridE")
    t.register_agenttype("Position", [("ids_sum", "i8")])
    t.register_agenttype("MovingAgent", [("value", "i8")])
    t.register_edgetype("OnPosition")
    return vh.create_model(t, "Raster_Test")


def gol_model():
    """SURVEY.md Appendix C (Game of Life in the reference's API)"""
    t = vh.ModelTypes()
    t.register_agenttype("Cell", [("active", "?")], "Immortal")
    t.register_edgetype("Neighbor", None, "Stateless", "SingleType", target="Cell")
    return vh.create_model(t, "GoL")


def gol_sim(backend, init):
    """init: bool array (nx, ny); cell (i, j) is created in column-major order (Raster.jl:44-49)"""
    sim = vh.create_simulation(gol_model(), backend=backend)
    sim.add_raster("grid", init.shape, "Cell", np.asarray(init, dtype="?").reshape(-1, order="F").view([("active", "?")]))
    sim.connect_raster_neighbors("grid", "Neighbor")
    sim.finish_init()
    return sim


PERSON = [("state", "u1"), ("days", "u1")]


def sir_model():
    """SURVEY.md Appendix C (Episim-style SIR stand-in)"""
    t = vh.ModelTypes()
    t.register_agenttype("Person", PERSON, "Immortal")
    t.register_agenttype("Location", [("n_inf", "i4")], "Immortal")
    t.register_edgetype("Visit", [("infectious", "?")], "SingleType", "IgnoreSourceState", target="Location")
    t.register_edgetype("Exposure", [("risk", "f4")], "IgnoreFrom", "SingleType", target="Person")
    t.register_param("beta", 0.05)
    t.register_param("n_locations", 1)
    t.register_param("visits_per_step", 2)
    t.register_param("infectious_days", 10)
    t.register_param("n_ranks", 1)
    return vh.create_model(t, "SIR")


def sir_sim(backend, n_persons, n_locations, seed=7, frac=0.01, beta=0.05):
    sim = vh.create_simulation(sir_model(), params={"n_locations": n_locations, "beta": beta}, backend=backend)
    st = np.zeros(n_persons, dtype=np.dtype(PERSON, align=True))
    st["state"] = (np.random.default_rng(seed).random(n_persons) < frac).astype("u1")
    sim.add_agents("Person", st)
    sim.add_agents("Location", np.zeros(n_locations, dtype=[("n_inf", "i4")]))
    sim.finish_init()
    return sim


def sir_step(sim, step):
    """the four applies of one SIR step; `seed` keys the per-agent uniform table of each apply"""
    sim.apply("sir_visit", "Person", ["Person"], ["Visit"], seed=4 * step)
    sim.apply("sir_tally", "Location", ["Visit"], ["Location"], seed=4 * step + 1)
    sim.apply("sir_expose", "Location", ["Location", "Visit"], ["Exposure"], seed=4 * step + 2)
    sim.apply("sir_infect", "Person", ["Person", "Exposure"], ["Person"], seed=4 * step + 3)


ANIMAL = [("energy", "i8"), ("pos", "i8", (2,))]
PPCELL = [("pos", "i8", (2,)), ("countdown", "i8")]
PP_EDGES = ["Position{Predator}", "Position{Prey}", "View{Predator}", "View{Prey}", "VisiblePrey", "Die", "Eat"]


def pp_model():
    """docs/examples/predator.jl:50-183 (detect_stateless(true): all seven edge types become :Stateless)"""
    old = vh.config.detect_stateless
    vh.detect_stateless(True)
    try:
        t = vh.ModelTypes()
        t.register_agenttype("Predator", ANIMAL)
        t.register_agenttype("Prey", ANIMAL)
        t.register_agenttype("Cell", PPCELL)
        for e in PP_EDGES:
            t.register_edgetype(e)
    finally:
        vh.detect_stateless(old)
    t.register_param("restart", 5)
    for sp in ("pred", "prey"):
        t.register_param(f"{sp}_gain", 5)
        t.register_param(f"{sp}_loss", 1)
        t.register_param(f"{sp}_thres", 5)
        t.register_param(f"{sp}_prob", 20)
    return vh.create_model(t, "Predator Prey")


def pp_sim(backend, dims=(100, 100), nprey=2000, npred=500, seed=3):
    """init as in predator.jl:185-229 with numpy's generator in place of Julia's"""
    rng = np.random.default_rng(seed)
    sim = vh.create_simulation(pp_model(), backend=backend)
    n = dims[0] * dims[1]
    cells = np.zeros(n, dtype=np.dtype(PPCELL, align=True))
    ii, jj = np.meshgrid(np.arange(1, dims[0] + 1), np.arange(1, dims[1] + 1), indexing="ij")
    cells["pos"][:, 0] = ii.reshape(-1, order="F")
    cells["pos"][:, 1] = jj.reshape(-1, order="F")
    cells["countdown"] = np.where(rng.random(n) < 0.5, 0, rng.integers(1, 6, n))
    sim.add_raster("raster", dims, "Cell", cells)
    for species, count in (("Prey", nprey), ("Predator", npred)):
        for _ in range(count):
            pos = (int(rng.integers(1, dims[0] + 1)), int(rng.integers(1, dims[1] + 1)))
            aid = sim.add_agent(species, (int(rng.integers(1, 11)), pos))
            sim.move_to("raster", aid, pos, None, f"Position{{{species}}}")
            sim.move_to("raster", aid, pos, f"View{{{species}}}", f"View{{{species}}}", distance=1, metric="manhatten")
    sim.finish_init()
    return sim



def pp_sim_bulk(backend, d=2048, nprey=838861, npred=209715, seed=3):
    """BASELINE config 3 (b): the docs' predator/prey model on a d x d raster, built with bulk adds (the per-animal move_to! calls of
    predator.jl:199-215 written out as edge arrays in the same order: position edge, then cell -> animal and animal -> cell for the
    position and its four von Neumann neighbours)."""
    rng = np.random.default_rng(seed)
    sim = vh.create_simulation(pp_model(), backend=backend)
    n = d * d
    cells = np.zeros(n, dtype=np.dtype(PPCELL, align=True))
    ii, jj = np.meshgrid(np.arange(1, d + 1), np.arange(1, d + 1), indexing="ij")
    cells["pos"][:, 0] = ii.reshape(-1, order="F")
    cells["pos"][:, 1] = jj.reshape(-1, order="F")
    cells["countdown"] = np.where(rng.random(n) < 0.5, 0, rng.integers(1, 6, n))
    cellids = sim.add_raster("raster", (d, d), "Cell", cells).reshape(-1, order="F")
    offs = [(0, 0), (0, -1), (-1, 0), (1, 0), (0, 1)]       # stencil(:manhatten, 2, 1) with the centre first (move_to!)
    for species, count in (("Prey", nprey), ("Predator", npred)):
        st = np.zeros(count, dtype=np.dtype(ANIMAL, align=True))
        st["energy"] = rng.integers(1, 11, count)
        st["pos"][:, 0] = rng.integers(1, d + 1, count)
        st["pos"][:, 1] = rng.integers(1, d + 1, count)
        ids = sim.add_agents(species, st)
        x, y = st["pos"][:, 0] - 1, st["pos"][:, 1] - 1
        sim.add_edges(ids, cellids[x + y * d], f"Position{{{species}}}")
        fr, to = [], []
        for dx, dy in offs:
            c = cellids[((x + dx) % d) + ((y + dy) % d) * d]
            fr += [c, ids]
            to += [ids, c]
        sim.add_edges(np.stack(fr, axis=1).reshape(-1), np.stack(to, axis=1).reshape(-1), f"View{{{species}}}")
    sim.finish_init()
    return sim


def pp_step(sim, step):
    """step!(sim), predator.jl:437-469: six applies; `seed` keys each apply's uniform table"""
    s = 6 * step
    sim.apply("pp_move", ["Prey"], ["Prey", "View{Prey}", "Cell"], ["Prey", "View{Prey}", "Position{Prey}"], seed=s)
    sim.apply("pp_find_prey", ["Cell"], ["Position{Prey}", "View{Predator}"], ["VisiblePrey"], seed=s + 1)
    sim.apply("pp_move", ["Predator"], ["Predator", "View{Predator}", "Cell", "Prey", "VisiblePrey"],
              ["Predator", "View{Predator}", "Position{Predator}"], seed=s + 2)
    sim.apply("pp_grow_food", "Cell", "Cell", "Cell", seed=s + 3)
    sim.apply("pp_try_eat", ["Cell"], ["Cell", "Position{Predator}", "Position{Prey}"], ["Cell", "Die", "Eat"], seed=s + 4)
    keep = ["Position{Predator}", "Position{Prey}", "View{Predator}", "View{Prey}"]
    sim.apply("pp_try_reproduce", ["Predator", "Prey"], ["Predator", "Prey", "Die", "Eat"], ["Predator", "Prey"] + keep, add_existing=keep, seed=s + 5)


def pp_globals(sim):
    """update_globals, predator.jl:402-417"""
    return {"predator_pop": sim.mapreduce(None, "+", "Predator", init=0), "prey_pop": sim.mapreduce(None, "+", "Prey", init=0),
            "cells_with_food": sim.mapreduce("countdown", "+", "Cell", equals=0),
            "predator_energy": sim.mapreduce("energy", "+", "Predator", init=0), "prey_energy": sim.mapreduce("energy", "+", "Prey", init=0)}


# ---- docs/examples/tutorial1.jl: the market model ("Excess Demand") ----
BUYER = [("alpha", "f8"), ("B", "f8")]
SELLER = [("p", "f8"), ("d_y", "f8")]
BOUGHT = [("x", "f8"), ("y", "f8")]


def market_model():
    """tutorial1.jl:143-216: Buyer, Seller, KnownSeller (stateless), Bought; params numBuyer / numSeller / knownSellers; two globals"""
    t = vh.ModelTypes()
    t.register_agenttype("Buyer", BUYER)
    t.register_agenttype("Seller", SELLER)
    t.register_edgetype("KnownSeller")
    t.register_edgetype("Bought", BOUGHT)
    t.register_global("x_minus_y", [])
    t.register_global("p", [])
    return vh.create_model(t, "Excess Demand")


def market_inputs(n_buyers, n_sellers, known, seed):
    """Buyer() = Buyer(rand(), rand(1:100)), Seller() = Seller(rand() + 0.5, 0), `known` random sellers per buyer (tutorial1.jl:111-112,343-349)"""
    rng = np.random.default_rng(seed)
    buyers = np.zeros(n_buyers, dtype=BUYER)
    buyers["alpha"], buyers["B"] = rng.random(n_buyers), rng.integers(1, 101, n_buyers)
    sellers = np.zeros(n_sellers, dtype=SELLER)
    sellers["p"] = rng.random(n_sellers) + 0.5
    picks = rng.integers(0, n_sellers, (n_buyers, known))        # rand(sellerids, k): with replacement
    return buyers, sellers, picks


def market_sim(backend, buyers, sellers, picks):
    sim = vh.create_simulation(market_model(), backend=backend)
    bids = sim.add_agents("Buyer", buyers)
    sids = sim.add_agents("Seller", sellers)
    sim.add_edges(sids[picks.reshape(-1)], np.repeat(bids, picks.shape[1]), "KnownSeller")     # for b, for s: add_edge!(sim, s, b, KnownSeller())
    sim.finish_init()
    return sim


def market_step(sim, step):
    """run_simulation's loop body (tutorial1.jl:574-580); the map closures that are not field selectors are registered map functors"""
    sim.apply("market_calc_demand", "Buyer", ["Buyer", "Seller", "KnownSeller"], "Bought", seed=step)
    sim.push_global("x_minus_y", sim.mapreduce_fn("market_x_minus_y", "+", "Bought"))          # mapreduce(sim, b -> b.x - b.y, +, Bought)
    sim.apply("market_calc_price", "Seller", ["Seller", "Bought"], "Seller")
    m = sim.mapreduce_fn("market_revenue", "+", "Seller")                                       # calc_average_price (:565-569)
    q = sim.mapreduce("d_y", "+", "Seller")
    sim.push_global("p", m / q)


# ---- test/mpi/test_agentstate.jl ----
def agentstate_model(immortal=True):
    """test_agentstate.jl:25-46 (:Immortal agents) and :107-113 (mortal agents)"""
    t = vh.ModelTypes()
    t.register_agenttype("ASAgent", FOO, *(["Immortal"] if immortal else []))
    t.register_edgetype("EdgeState", FOO, "SingleEdge")
    t.register_edgetype("NewEdge", FOO, "SingleEdge")
    return vh.create_model(t, "agentstatetest" if immortal else "agentstatetest-mortal")


def agentstate_scenario(backend, n, immortal, device=0):
    """test_agentstate.jl:46-105 / :107-185: a chain ids[to-1] -> ids[to] whose edge states repeat the source's state; every check is an
    assertion inside a transition (ctx.require = the closure's @test), so a stale source state raises an AssertionError in apply!"""
    sim = vh.create_simulation(agentstate_model(immortal), backend=backend, device=device)
    ids = sim.add_agents("ASAgent", foos(range(1, n + 1)))
    for to in range(1, n):
        sim.add_edge(int(ids[to - 1]), int(ids[to]), "EdgeState", to)
    sim.finish_init(partition_algo="EqualAgentNumbers")
    A, ES, NE = "ASAgent", "EdgeState", "NewEdge"
    sim.apply("as_check_x1_EdgeState", [A], [A, ES], [])                 # the state can be read after the initialisation
    if not immortal:
        sim.apply("as_check_x1_EdgeState", [A], [A, ES], [])             # a second time: already transferred states (:137-144)
    sim.apply("as_double", [A], [A], [A])
    sim.apply("as_check_x2_EdgeState", [A], [A, ES], [])                 # the accessible state follows the update
    sim.apply("as_halve", [A], [A], [A])
    sim.apply("as_require_no_NewEdge", [A], [A, NE], [])                 # reading another edge type transfers nothing new
    sim.apply("as_copy_edges", [A], [ES], [NE])
    sim.apply("as_check_x1_EdgeState", [A], [A, ES], [])
    sim.apply("as_check_x1_NewEdge", [A], [A, NE], [])                   # and through the new edges
    assert sim.num_edges(NE) == n - 1 == sim.num_edges(ES)
    return sim
