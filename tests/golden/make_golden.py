"""Generates the committed fixtures of tests/golden/ (run from the repo root: python tests/golden/make_golden.py).

The reference is pure Julia and cannot run in this image, and none of its tests pins a model-level result of the
BASELINE configs, so these vectors are NOT outputs of the reference: they are outputs of the CPU oracle (the restatement of the
reference's algorithm, itself pinned on the reference's known-answer tests in tests/test_core.py ... test_raster.py), each
cross-checked here against an independent numpy restatement where the model allows it (HK, Game of Life).  They anchor both the
oracle (CPU suite) and the CUDA engine (GPU suite) against drift.  Inputs are regenerated from seeds by the tests, only the
expected outputs are stored."""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import vahana_b200 as vh  # noqa: E402
from models import ba_graph, gol_sim, hk_sim, market_inputs, market_sim, market_step, pp_globals, pp_sim, pp_step, sir_sim, sir_step  # noqa: E402

HK = dict(n=2000, m=8, graph_seed=1, opinion_seed=1, steps=10)
GOL = dict(shape=(48, 40), seed=2, density=0.35, generations=25)
SIR = dict(n=3000, nl=250, beta=0.3, steps=12)
PP = dict(dims=(30, 30), nprey=180, npred=45, steps=20)
MARKET = dict(nb=5000, ns=40, known=3, seed=8, steps=20)            # the tutorial's market model (docs/examples/tutorial1.jl), 100 x its size
PP_DOCS = dict(dims=(100, 100), nprey=2000, npred=500, steps=400, every=25)      # BASELINE config 3 (a): the docs model at the docs size


def hk_numpy(n, uv, op, eps, steps):
    nb = [[] for _ in range(n)]
    for u, v in uv:
        nb[v].append(u)
        nb[u].append(v)
    for i in range(n):
        nb[i].append(i)
    for _ in range(steps):
        op = np.array([np.mean([op[j] for j in nb[i] if abs(op[j] - op[i]) < eps]) for i in range(n)])
    return op


def life_numpy(a):
    n = sum(np.roll(np.roll(a, dx, 0), dy, 1) for dx in (-1, 0, 1) for dy in (-1, 0, 1) if (dx, dy) != (0, 0))
    return (n == 3) | (a & (n == 2))


def pp_digest(sim):
    h = hashlib.sha256()
    for T in ("Predator", "Prey", "Cell"):
        h.update(sim.all_agents(T).tobytes())
        h.update(np.asarray(sim.all_agentids(T), dtype=np.uint64).tobytes())
    return h.hexdigest()


def market(ob):
    """sellers' (p, d_y) after MARKET['steps'] steps and the two global series; the oracle is bit-exact with the independent numpy
    restatement of tests/test_zzm_market.py (own Philox4x32-10), which is asserted here before the fixture is written"""
    from test_zzm_market import numpy_market
    buyers, sellers, picks = market_inputs(MARKET["nb"], MARKET["ns"], MARKET["known"], MARKET["seed"])
    sim = market_sim(ob, buyers, sellers, picks)
    for step in range(MARKET["steps"]):
        market_step(sim, step)
    s = sim.all_agents("Seller")
    p, d_y, xmy, avg, _ = numpy_market(buyers, sellers, picks, MARKET["steps"])
    assert np.array_equal(s["p"], p) and np.array_equal(s["d_y"], d_y)                                # independent restatement
    np.testing.assert_allclose(sim.get_global("p"), avg, rtol=1e-12)
    np.savez_compressed(os.path.join(HERE, "market_5000x40.npz"), p=s["p"], d_y=s["d_y"], x_minus_y=np.array(sim.get_global("x_minus_y")),
                        avg_p=np.array(sim.get_global("p")))


def main():
    import subprocess
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True)
    ob = vh.load_backend(os.path.join(ROOT, "oracle", "_build", "libvahana_oracle.so"))
    market(ob)
    if "--only-market" in sys.argv:
        return

    uv = ba_graph(HK["n"], HK["m"], HK["graph_seed"])
    op0 = np.random.default_rng(HK["opinion_seed"]).random(HK["n"])
    out = {}
    for eps in (0.02, 0.25):
        sim, _ = hk_sim(ob, HK["n"], uv, op0, eps)
        for _ in range(HK["steps"]):
            sim.apply("hk_step", "HKAgent", ["HKAgent", "Knows"], "HKAgent")
        op = sim.all_agents("HKAgent")["opinion"].copy()
        np.testing.assert_allclose(op, hk_numpy(HK["n"], uv, op0, eps, HK["steps"]), rtol=1e-12)     # independent restatement
        out["eps_%g" % eps] = op
    np.savez_compressed(os.path.join(HERE, "hk_ba2000.npz"), **out)

    init = np.random.default_rng(GOL["seed"]).random(GOL["shape"]) < GOL["density"]
    sim = gol_sim(ob, init)
    a = init.copy()
    for _ in range(GOL["generations"]):
        sim.apply("gol_life", "Cell", ["Cell", "Neighbor"], "Cell")
        a = life_numpy(a)
    grid = sim.rastervalues("grid", "active", "Cell")
    assert np.array_equal(grid, a)                                                                    # independent restatement
    np.savez_compressed(os.path.join(HERE, "gol_48x40.npz"), grid=np.packbits(grid), alive=int(grid.sum()))

    sim = sir_sim(ob, SIR["n"], SIR["nl"], beta=SIR["beta"])
    counts = []
    for step in range(SIR["steps"]):
        sir_step(sim, step)
        s = sim.all_agents("Person")["state"]
        counts.append([int((s == k).sum()) for k in range(3)])
    p = sim.all_agents("Person")
    np.savez_compressed(os.path.join(HERE, "sir_3000.npz"), state=p["state"], days=p["days"], n_inf=sim.all_agents("Location")["n_inf"],
                        counts=np.array(counts))

    sim = pp_sim(ob, PP["dims"], PP["nprey"], PP["npred"])
    traj = []
    for step in range(PP["steps"]):
        pp_step(sim, step)
        g = pp_globals(sim)
        traj.append([g["prey_pop"], g["predator_pop"], g["cells_with_food"], g["prey_energy"], g["predator_energy"]])
    np.savez_compressed(os.path.join(HERE, "pp_30x30.npz"), trajectory=np.array(traj, dtype=np.int64), digest=np.array(pp_digest(sim)))
    sim = pp_sim(ob, PP_DOCS["dims"], PP_DOCS["nprey"], PP_DOCS["npred"])
    traj = []
    for step in range(PP_DOCS["steps"]):
        pp_step(sim, step)
        if step % PP_DOCS["every"] == PP_DOCS["every"] - 1:
            g = pp_globals(sim)
            traj.append([step, g["prey_pop"], g["predator_pop"], g["cells_with_food"], g["prey_energy"], g["predator_energy"]])
    np.savez_compressed(os.path.join(HERE, "pp_docs_100x100.npz"), trajectory=np.array(traj, dtype=np.int64), digest=np.array(pp_digest(sim)))
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
