"""Replays /root/reference/test/globals.jl and /root/reference/test/parametric_types.jl (the two remaining single-process
test files of the reference's runtests.jl that touch the path: host-side globals, G1; type parameters of agent / edge structs)."""
import numpy as np

import vahana_b200 as vh


def test_globals(backend):  # test/globals.jl:1-17
    types = vh.ModelTypes().register_agenttype("Agent", [("foo", "i8")])
    model = vh.create_model(types, "globals")
    sim = vh.create_simulation(model, None, {"foo": 0.0, "bar": []}, backend=backend)
    sim.set_global("foo", 1.1)
    sim.push_global("bar", 1)
    sim.push_global("bar", 2)
    assert sim.get_global("foo") == 1.1
    assert sim.get_global("bar") == [1, 2]
    sim.modify_global("foo", lambda v: v * 2)          # modify_global! = set_global!(f(get_global)), src/Global.jl:52-54
    assert sim.get_global("foo") == 2.2
    sim.finish_simulation()


def test_parametric_types(backend):  # test/parametric_types.jl:1-38
    # PAgent{Float64}, PAgent2 and PEdge{Float64}: a type parameter only fixes the field type, i.e. the registered layout
    types = (vh.ModelTypes()
             .register_agenttype("PAgent{Float64}", [("t", "f8")])
             .register_agenttype("PAgent2", [("t", "f8")])
             .register_edgetype("PEdge{Float64}", [("t", "f8")]))
    sim = vh.create_simulation(vh.create_model(types, "parametric types"), backend=backend)
    id1 = sim.add_agent("PAgent{Float64}", (1.0,))
    id2 = sim.add_agent("PAgent2", (2.0,))
    sim.add_edge(id1, id2, "PEdge{Float64}", (3.0,))
    sim.finish_init()
    sim.disable_transition_checks(True)
    e = sim.edges(id2, "PEdge{Float64}")
    assert e[0][0] == id1 and e[0][1]["t"] == 3.0                                   # first(edges(...)).state.t == 3.0
    assert [s["t"] for s in sim.edgestates(id2, "PEdge{Float64}")] == [3.0]
    assert [s["t"] for s in sim.neighborstates(id2, "PEdge{Float64}", "PAgent{Float64}")] == [1.0]
    assert vh.type_nr(id1) == 1 and vh.type_nr(id2) == 2
    # the identity transition of the test (`agent` returned unchanged) leaves both types as they were
    sim.apply("identity", "PAgent2", ["PAgent{Float64}", "PAgent2", "PEdge{Float64}"], "PAgent2")
    assert sim.all_agents("PAgent2")["t"].tolist() == [2.0]
    assert sim.all_agents("PAgent{Float64}")["t"].tolist() == [1.0]
    sim.finish_simulation()


def test_julia_wrapper_binds_only_declared_symbols():
    """The Julia `ccall` wrapper cannot run here (no Julia in the image); at least every symbol it binds must be
    declared in include/vahana_b200.h and exported by the library, and the entry points of the hot path must all be bound."""
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = open(os.path.join(root, "vahana.jl_b200", "julia", "VahanaB200.jl")).read()
    header = open(os.path.join(root, "include", "vahana_b200.h")).read()
    bound = set(re.findall(r"\(:(vb_\w+), LIB\)", src))
    assert bound, "no ccall found"
    for sym in bound:
        assert re.search(r"\b%s\(" % sym, header), sym + " is not declared in include/vahana_b200.h"
        assert sym in vh.ABI_SYMBOLS, sym
    for sym in ["vb_sim_create", "vb_add_agents", "vb_add_edges", "vb_finish_init", "vb_apply", "vb_edges_of", "vb_mapreduce",
                "vb_add_raster", "vb_connect_raster_neighbors", "vb_move_to", "vb_rastervalues", "vb_num_agents", "vb_all_agents"]:
        assert sym in bound, sym


def _declared(header_path):
    import re
    txt = open(header_path).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return set(re.findall(r"^\s*(?:const\s+)?[A-Za-z_][\w\s\*]*?\b(vbw?_\w+)\s*\(", txt, flags=re.M))


def test_cabi_library_exports_every_declared_symbol():
    """The drop-in boundary: both libraries (the CUDA engine — loaded here without a GPU, no compute call — and the oracle behind the
    same ABI) export every function include/vahana_b200.h and include/vahana_workloads.h declare; the Python mirror's ABI_SYMBOLS
    list is exactly the header's set."""
    import ctypes as C
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    core = _declared(os.path.join(root, "include", "vahana_b200.h"))
    work = _declared(os.path.join(root, "include", "vahana_workloads.h"))
    assert len(core) > 40 and "vb_apply" in core and "vb_register_transition" not in core
    assert core == set(vh.ABI_SYMBOLS), (core ^ set(vh.ABI_SYMBOLS))
    subprocess.run(["make", "-s", "-C", os.path.join(root, "oracle")], check=True)
    libs = [os.path.join(root, "oracle", "_build", "libvahana_oracle.so")]
    cuda_lib = os.path.join(root, "vahana.jl_b200", "csrc", "build", "libvahana_b200.so")
    if not os.path.exists(cuda_lib):      # nvcc cross-compiles without a GPU (minutes); normally __graft_entry__.build() ran before
        import sys
        subprocess.run([sys.executable, os.path.join(root, "vahana.jl_b200", "build.py")], check=True)
    libs.append(cuda_lib)
    for path in libs:
        lib = C.CDLL(path, mode=C.RTLD_LOCAL)
        missing = [s for s in sorted(core | work) if not hasattr(lib, s)]
        assert not missing, (path, missing)
    lib.vb_backend.restype = C.c_char_p
    assert lib.vb_backend() == b"cuda-sm100a"


def test_nonmutating_apply(backend):  # src/Simulation.jl:847-856: apply = copy_simulation + apply! on the copy
    from models import createsim
    sim = createsim(backend)[0]
    before = {T: sim.all_agents(T).tobytes() for T in ("AMortal", "AImm", "AImmFixed")}
    n0 = sim.num_transitions()
    new = vh.apply(sim, "kill_all", "AMortal", [], "AMortal")
    assert new.num_agents("AMortal") == 0 and sim.num_agents("AMortal") > 0          # only the copy changed
    assert {T: sim.all_agents(T).tobytes() for T in before} == before
    assert sim.num_transitions() == n0 and new.num_transitions() == n0 + 1
    new.finish_simulation()
    sim.finish_simulation()


def test_duration_log_in_the_reference_format(backend, tmp_path):  # src/Logging.jl:30-103: <Begin>/<End> pairs -> "x |#| Duration (ms)"
    import re
    from models import core_model, add_example_network
    sim = vh.create_simulation(core_model(), backend=backend, logging=True, log_path=str(tmp_path))
    add_example_network(sim)
    sim.finish_init()
    sim.apply("identity", "AMortal", ["AMortal"], ["AMortal"])
    sim.apply("kill_all", "AMortal", [], "AMortal")
    path = sim.logger.path
    assert path == str(tmp_path / "Test_Core_0.log")           # <name>_<rank>.log
    sim.finish_simulation()
    txt = open(path).read()
    recs = re.findall(r"Start \(sec\): ([0-9.e+-]+)\n   (\S+) \|#\| Duration \(ms\): ([0-9.e+-]+)\n((?:    \w+ = .*\n)*)", txt)
    assert [r[1] for r in recs] == ["finish_init!", "apply!", "apply!", "finish_simulation!"]
    assert "func = identity" in recs[1][3] and "transition = 2" in recs[1][3]       # finish_init! set num_transitions to 1
    assert "func = kill_all" in recs[2][3] and "transition = 3" in recs[2][3]
    assert all(float(r[2]) >= 0 for r in recs) and float(recs[0][0]) <= float(recs[1][0]) <= float(recs[2][0])


def test_graph_bridges(backend):
    """add_graph! from a graph object and vahanagraph (src/GraphsSupport.jl:34-57,212-289; test/graphs.jl:37-43: K4 id sums)"""
    import networkx as nx
    from models import hk_model
    sim = vh.create_simulation(hk_model(), backend=backend)
    g = nx.complete_graph(4)
    ids = vh.add_graph(sim, g, None, "HKAgent", np.arange(4.0).view([("opinion", "f8")]), "Knows")
    sim.finish_init()
    assert sim.num_edges("Knows") == 12 and len(ids) == 4
    vg = vh.vahanagraph(sim)
    assert np.array_equal(vg["g2v"], ids) and len(vg["src"]) == 12 and set(vg["edgetype"].tolist()) == {0}
    assert sorted(zip(vg["src"].tolist(), vg["dst"].tolist())) == sorted((u, v) for u in range(4) for v in range(4) if u != v)
    back = vh.to_networkx(sim)
    assert back.number_of_nodes() == 4 and back.number_of_edges() == 12 and back.nodes[2]["id"] == int(ids[2])
    d = nx.DiGraph([(0, 1), (1, 2), (2, 0), (0, 1)])            # a directed graph: one Vahana edge per graph edge
    sim2 = vh.create_simulation(hk_model(), backend=backend)
    vh.add_graph(sim2, d, None, "HKAgent", np.zeros(3).view([("opinion", "f8")]), "Knows")
    sim2.add_edges(sim2.all_agentids("HKAgent")[:1], sim2.all_agentids("HKAgent")[1:2], "Knows")    # a parallel edge 0 -> 1
    sim2.finish_init()
    assert sim2.num_edges("Knows") == 4
    assert len(vh.vahanagraph(sim2, drop_multiedges=True)["src"]) == 3
    assert len(vh.vahanasimplegraph(sim2)["src"]) == 3                       # vahanasimplegraph (src/GraphsSupport.jl:107-160): structure only


def test_show_and_dataframes(backend):
    """show(sim) (src/REPL.jl:81-168), DataFrame(sim, T) and GlobalsDataFrame(sim) (src/optional/DataFrames.jl:54-185)"""
    from models import market_inputs, market_sim, market_step
    buyers, sellers, picks = market_inputs(20, 3, 2, seed=1)
    sim = market_sim(backend, buyers, sellers, picks)
    for step in range(3):
        market_step(sim, step)
    text = sim.show()
    assert "Model Name: Excess Demand" in text and "Type Buyer with 20 agent(s)" in text and "Type Seller with 3 agent(s)" in text
    assert "Type KnownSeller with 40 edge(s)" in text and "Type Bought with 20 edge(s)" in text
    assert ":x_minus_y |> last :" in text and "(length: 3)" in text and "initialization" not in text
    df = sim.dataframe("Seller")
    assert list(df.columns) == ["id", "p", "d_y"] and len(df) == 3 and np.array_equal(df["p"].to_numpy(), sim.all_agents("Seller")["p"])
    assert np.array_equal(df["id"].to_numpy(), sim.all_agentids("Seller"))
    de = sim.dataframe("Bought", types=True)
    assert list(de.columns) == ["from", "from_type", "to", "to_type", "x", "y"] and len(de) == 20
    assert set(de["from_type"]) == {"Buyer"} and set(de["to_type"]) == {"Seller"}
    assert sim.dataframe("Bought", localnr=True)["to"].max() <= 3
    g = sim.globals_dataframe()
    assert list(g.columns) == ["x_minus_y", "p"] and len(g) == 3
    # a simulation that is still being initialised says so, a raster is listed with its dimensions
    from models import hk_model
    s2 = vh.create_simulation(hk_model(), backend=backend)
    assert "Still in initialization process!." in s2.show() and ":eps : 0.02" in s2.show()


def test_julia_ccall_argument_counts_match_the_header():
    """no Julia in the image: at least the number of arguments of every `ccall` in the wrapper must equal the number of parameters of
    the C prototype it binds (include/vahana_b200.h), and the argument-type tuple must be as long as the argument list"""
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = open(os.path.join(root, "vahana.jl_b200", "julia", "VahanaB200.jl")).read()
    header = re.sub(r"/\*.*?\*/", "", open(os.path.join(root, "include", "vahana_b200.h")).read(), flags=re.S)

    def split_top(s):
        out, depth, cur = [], 0, ""
        for ch in s:
            if ch in "([{":
                depth += 1
            elif ch in ")]}":
                depth -= 1
            if ch == "," and depth == 0:
                out.append(cur.strip())
                cur = ""
            else:
                cur += ch
        if cur.strip():
            out.append(cur.strip())
        return out

    def balanced(s, start):          # text between the parenthesis at `start` and its partner
        depth = 0
        for i in range(start, len(s)):
            if s[i] == "(":
                depth += 1
            elif s[i] == ")":
                depth -= 1
                if depth == 0:
                    return s[start + 1:i]
        raise AssertionError("unbalanced")

    nparams = {}
    for m in re.finditer(r"\b(vb_\w+)\s*\(", header):
        params = balanced(header, m.end() - 1).strip()
        nparams[m.group(1)] = 0 if params in ("", "void") else len(split_top(params))
    checked = 0
    for m in re.finditer(r"ccall\(\(:(vb_\w+), LIB\)", src):
        args = split_top(balanced(src, m.start() + len("ccall")))
        # args = [(:sym, LIB), RetType, (ArgTypes...), actual arguments...]
        types = args[2].strip()
        assert types.startswith("(") and types.endswith(")"), (m.group(1), types)
        ntypes = len(split_top(types[1:-1].rstrip(",")))
        if types[1:-1].strip() == "":
            ntypes = 0
        assert ntypes == nparams[m.group(1)], f"{m.group(1)}: {ntypes} ccall argument types, {nparams[m.group(1)]} parameters in the header"
        assert len(args) - 3 == ntypes, f"{m.group(1)}: {len(args) - 3} arguments passed for {ntypes} argument types"
        checked += 1
    assert checked > 30
