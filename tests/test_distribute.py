"""finish_init!(distribute = true) (src/Simulation.jl:403-476, distribute! src/MPI.jl:11-84): the host-side plan as a pure function,
and the exchange + rebuild on two gloo ranks against a recording stand-in for the engine library (the CUDA engine needs GPUs; the
2-GPU variant is tests/mgpu_distribute.py)."""
import os
import subprocess
import sys

import numpy as np
import pytest

import vahana_b200 as vh

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ids(tid, n, rank=0):
    return np.array([vh.agent_id(tid, rank, k) for k in range(1, n + 1)], dtype=np.uint64)


def test_plan_equal_agent_numbers_matches_reference_partition():
    """_create_equal_partition (src/Simulation.jl:353-367): rank i (1-based) owns ids[(i-1)s+1+min(i-1,r) : is+min(i,r)]"""
    for n, world in [(10, 3), (7, 7), (100, 8), (5, 2), (3, 4), (0, 2)]:
        ids = _ids(1, n)
        states = np.arange(n, dtype=np.int64) * 10
        shards, old, new, bounds = vh.plan_distribution({1: (ids, states)}, {}, world)
        s, r = divmod(n, world)
        for i in range(1, world + 1):
            lo, hi = (i - 1) * s + 1 + min(i - 1, r), i * s + min(i, r)          # Julia's 1-based inclusive range
            cnt, st = shards[i - 1]["agents"][1]
            assert cnt == max(0, hi - lo + 1)
            assert np.array_equal(st, states[lo - 1:hi])
        assert np.array_equal(old, ids)
        # new ids: dense per rank, in the old order
        for k, (o, nw) in enumerate(zip(old, new)):
            owner = int(np.searchsorted(np.array(bounds[1][1:]), k, side="right"))
            assert vh.type_nr(nw) == 1 and vh.process_nr(nw) == owner and vh.agent_nr(nw) == k - bounds[1][owner] + 1


def test_plan_edges_follow_their_target_in_add_order():
    a, b = _ids(1, 6), _ids(2, 4)
    rng = np.random.default_rng(0)
    allids = np.concatenate([a, b])
    fr, to = rng.choice(allids, 200), rng.choice(allids, 200)
    st = np.arange(200, dtype=np.int64)
    shards, old, new, _ = vh.plan_distribution({1: (a, None), 2: (b, np.arange(4.0))}, {"E": (fr, to, st), "S": (fr[:50], to[:50], None)}, 3)
    m = dict(zip(old.tolist(), new.tolist()))
    seen = []
    for r in range(3):
        f2, t2, s2 = shards[r]["edges"]["E"]
        assert all(vh.process_nr(t) == r for t in t2)                              # stored on the target's rank (src/EdgeMethods.jl:396-398)
        assert np.array_equal(f2, [m[int(x)] for x in fr[s2]]) and np.array_equal(t2, [m[int(x)] for x in to[s2]])
        assert np.all(np.diff(s2) > 0)                                             # add order kept: every target keeps its push! order
        seen += s2.tolist()
        assert shards[r]["edges"]["S"][2] is None and len(shards[r]["edges"]["S"][0]) == int(np.isin(np.arange(50), s2).sum())
    assert sorted(seen) == list(range(200))                                        # every edge on exactly one rank
    assert sum(shards[r]["agents"][2][0] for r in range(3)) == 4 and shards[0]["agents"][1][1] is None


def test_plan_explicit_partition_and_errors():
    a = _ids(1, 5)
    part = {int(a[0]): 2, int(a[1]): 1, int(a[2]): 2, int(a[3]): 2, int(a[4]): 1}    # 1-based ProcessIDs as in the reference
    shards, old, new, bounds = vh.plan_distribution({1: (a, np.arange(5))}, {}, 2, part)
    assert bounds == {}
    assert np.array_equal(shards[0]["agents"][1][1], [1, 4]) and np.array_equal(shards[1]["agents"][1][1], [0, 2, 3])
    m = dict(zip(old.tolist(), new.tolist()))
    assert m[int(a[1])] == vh.agent_id(1, 0, 1) and m[int(a[4])] == vh.agent_id(1, 0, 2) and m[int(a[3])] == vh.agent_id(1, 1, 3)
    assert np.array_equal(vh.updateids(m, a[:2]), [vh.agent_id(1, 1, 1), vh.agent_id(1, 0, 1)])
    # every rank ran the initialisation code with ids of its own rank: the rank bits of the old id do not matter (remove_process, src/Simulation.jl:481)
    assert vh.updateids(m, vh.agent_id(1, 3, 2)) == vh.agent_id(1, 0, 1) and vh.remove_process(vh.agent_id(2, 5, 7)) == vh.agent_id(2, 0, 7)
    with pytest.raises(AssertionError):
        vh.plan_distribution({1: (a, None)}, {}, 2, {int(a[0]): 1})                  # an agent without a rank
    with pytest.raises(AssertionError):
        vh.plan_distribution({1: (a, None)}, {}, 2, {int(x): 3 for x in a})          # rank outside of 1..mpi.size
    with pytest.raises(AssertionError):
        vh.plan_distribution({1: (a, None)}, {"E": (a[:1], np.array([vh.agent_id(1, 0, 99)], dtype=np.uint64), None)}, 2)   # dangling id


def test_plan_remaps_dense_and_sparse_ids_alike():
    """the id rewrite of the edges goes through a per-type table when the ids are dense (what add_agents! hands out) and through a
    search otherwise: same result, same errors"""
    rng = np.random.default_rng(5)
    dense, dense2 = _ids(1, 5000), _ids(2, 300)
    sparse = _ids(1, 200000)[::97]                                        # holes: more than four numbers per agent
    for a_ids in (dense, sparse):
        allids = np.concatenate([a_ids, dense2])
        fr, to = allids[rng.integers(0, len(allids), 20000)], allids[rng.integers(0, len(allids), 20000)]
        shards, old, new, _ = vh.plan_distribution({1: (a_ids, None), 2: (dense2, None)}, {"E": (fr, to, None)}, 3)
        m = dict(zip(old.tolist(), new.tolist()))
        got_f = np.concatenate([shards[r]["edges"]["E"][0] for r in range(3)])
        got_t = np.concatenate([shards[r]["edges"]["E"][1] for r in range(3)])
        owner = np.array([vh.process_nr(m[int(x)]) for x in to])
        order = np.concatenate([np.nonzero(owner == r)[0] for r in range(3)])
        assert [int(x) for x in got_f] == [m[int(x)] for x in fr[order]] and [int(x) for x in got_t] == [m[int(x)] for x in to[order]]
        for bad in (np.uint64(int(a_ids[-1]) + 1), np.uint64(int(a_ids[0]) - 1) if vh.agent_nr(int(a_ids[0])) > 1 else np.uint64(vh.agent_id(3, 0, 1)),
                    np.uint64(vh.agent_id(2, 0, 301))):
            with pytest.raises(AssertionError):
                vh.plan_distribution({1: (a_ids, None), 2: (dense2, None)}, {"E": (fr[:1], np.array([bad], dtype=np.uint64), None)}, 3)


def test_finish_init_single_rank_idmapping_is_identity(oracle):
    backend = oracle   # host logic of the mirror; the CUDA engine takes the same path in every GPU test that calls finish_init()
    from models import edges_model, foos
    sim = vh.create_simulation(edges_model(), backend=backend)
    ids = sim.add_agents("Agent", foos([1, 2, 3]))
    m = sim.finish_init(return_idmapping=True, partition_algo="EqualAgentNumbers")
    assert m == {int(i): int(i) for i in ids}
    sim2 = vh.create_simulation(edges_model(), backend=backend)
    assert sim2.finish_init() is sim2 and sim2.finish_init.__doc__


def test_graph_growing_partition_stands_in_for_metis():
    """partition_algo = :Metis (the reference's default, src/Simulation.jl:420-446) without Metis: parts of equal size (up to one agent),
    every agent placed, fewer cut edges than blocks of a shuffled numbering, deterministic"""
    rng = np.random.default_rng(4)
    side = 24
    n = side * side
    perm = rng.permutation(n)                                   # a grid graph whose agent numbers say nothing about positions
    ids = np.array([vh.agent_id(1, 0, int(k) + 1) for k in range(n)], dtype=np.uint64)
    x, y = np.meshgrid(np.arange(side), np.arange(side), indexing="ij")
    lin = (x * side + y)
    pairs = np.concatenate([np.stack([lin[:-1, :].ravel(), lin[1:, :].ravel()], 1), np.stack([lin[:, :-1].ravel(), lin[:, 1:].ravel()], 1)])
    fr, to = ids[perm[pairs[:, 0]]], ids[perm[pairs[:, 1]]]
    agents, edges = {1: (ids, None)}, {"E": (fr, to, None)}
    for world in (2, 3, 4):
        part = vh.graph_growing_partition(agents, edges, world)
        assert part == vh.graph_growing_partition(agents, edges, world)
        owner = np.array([part[int(i)] for i in ids])
        counts = np.bincount(owner, minlength=world + 1)[1:]
        assert counts.sum() == n and counts.max() - counts.min() <= 1 and owner.min() == 1 and owner.max() == world
        cut = int((np.array([part[int(i)] for i in fr]) != np.array([part[int(i)] for i in to])).sum())
        b = vh.equal_partition(n, world)
        blocks = np.searchsorted(np.array(b[1:]), np.arange(n), side="right")
        cut_blocks = int((blocks[perm[pairs[:, 0]]] != blocks[perm[pairs[:, 1]]]).sum())
        assert cut < cut_blocks / 3, (world, cut, cut_blocks)
        shards, old, new, _ = vh.plan_distribution(agents, edges, world, part)         # the plan takes it like any explicit partition
        assert sum(s_["agents"][1][0] for s_ in shards) == n


def test_raster_is_staged_for_the_hand_out(oracle):
    """add_raster! / connect_raster_neighbors! of the initialisation phase are kept on the host like every other add, so that
    finish_init!(distribute = true) can hand the cells out and broadcast the id grid (broadcastids, src/MPI.jl:59-73): the staged agents
    and edges are exactly what the engine holds, row order (per-target push! order) included."""
    from models import gol_model
    init = np.random.default_rng(3).random((5, 4)) < 0.5
    sim = vh.create_simulation(gol_model(), backend=oracle)
    sim._stage = {"agents": {}, "edges": {}}                    # what a multi-rank backend switches on
    ids = sim.add_raster("grid", init.shape, "Cell", np.asarray(init, dtype="?").reshape(-1, order="F").view([("active", "?")]))
    sim.connect_raster_neighbors("grid", "Neighbor")
    sim.connect_raster_neighbors("grid", "Neighbor", distance=1, metric="manhatten", periodic=False)     # a second call appends to the rows
    assert sim._unstageable is None
    (sids, sstates), = sim._stage["agents"][sim._aid["Cell"]]
    assert np.array_equal(sids, ids.reshape(-1, order="F")) and np.array_equal(sstates["active"], init.reshape(-1, order="F"))
    dims, tid, rids = sim._stage["rasters"]["grid"]
    assert dims == (5, 4) and tid == sim._aid["Cell"] and np.array_equal(rids, sids)
    fr = np.concatenate([c[0] for c in sim._stage["edges"]["Neighbor"]])
    to = np.concatenate([c[1] for c in sim._stage["edges"]["Neighbor"]])
    sim._stage = None
    sim.finish_init()
    off, efrom, _ = sim.export_csr("Neighbor", "Cell", init.size)
    assert fr.shape[0] == int(off[-1]) == 8 * 20 + (2 * 4 * 4 + 2 * 5 * 3)       # Moore, periodic + von Neumann, clipped (62 directed edges on 5 x 4)
    # per target: the staged edges in staging order are the engine's row
    for k, cid in enumerate(sids):
        assert np.array_equal(fr[to == cid], efrom[int(off[k]):int(off[k + 1])]), k
    # the plan hands cells and grid out consistently
    shards, old, new, bounds = vh.plan_distribution({tid: (sids, sstates)}, {"Neighbor": (fr, to, None)}, 2)
    grid_new = new[np.searchsorted(old, rids)]
    b = bounds[tid]
    assert [vh.process_nr(int(x)) for x in grid_new] == [0] * b[1] + [1] * (20 - b[1])
    assert sum(s_["edges"]["Neighbor"][1].shape[0] for s_ in shards) == fr.shape[0]


def test_move_to_is_staged_for_the_hand_out(oracle):
    """move_to! of the initialisation phase (src/Raster.jl:437-477) is kept on the host as the edges it adds — the cell at `pos`, then
    the stencil cells, per cell the from-raster edge before the to-raster edge — so that finish_init!(distribute = true) can hand out a
    model whose agents are placed on a raster (predator/prey, docs): the staged edges are exactly the rows the engine holds."""
    from models import raster_model
    sim = vh.create_simulation(raster_model(), backend=oracle)
    sim._stage = {"agents": {}, "edges": {}}
    cells = sim.add_raster("raster", (7, 5), "Position", lambda p: (0,))
    movers = [sim.add_agent("MovingAgent", v) for v in range(1, 7)]
    sim.move_to("raster", movers[0], (1, 1), "OnPosition", "OnPosition")                                         # both directions in one edge type
    sim.move_to("raster", movers[1], (4, 3), "OnPosition", None, distance=2)
    sim.move_to("raster", movers[2], (7, 5), None, "OnPosition", distance=2, metric="manhatten", periodic=False)   # clipped at the corner
    sim.move_to("raster", movers[3], (2, 2), "OnPosition", "OnPosition", distance=1.5, metric="euclidean", only_surrounding=True)
    sim.move_to("raster", movers[4], (4, 3), "OnPosition", "GridE", distance=1)                                  # two edge types
    sim.move_to("raster", movers[5], (4, 3), "OnPosition", None)
    assert sim._unstageable is None
    st = {name: (np.concatenate([c[0] for c in ch]), np.concatenate([c[1] for c in ch])) for name, ch in sim._stage["edges"].items()}
    sim._stage = None
    sim.finish_init()
    assert len(st["OnPosition"][0]) == sim.num_edges("OnPosition") == 2 + 25 + 6 + 2 * 8 + 9 + 1
    assert len(st["GridE"][0]) == sim.num_edges("GridE") == 9
    sim.disable_transition_checks(True)
    for name, (fr, to) in st.items():
        for target in [int(x) for x in np.unique(to)]:
            got = sim.neighborids(target, name)
            assert np.array_equal(fr[to == np.uint64(target)], np.asarray(got, dtype=np.uint64)), (name, target)
    assert int(cells[3, 2]) in [int(x) for x in st["GridE"][1]]
    # only rasters added through add_raster can be staged
    assert list(vh.raster_move_cells((3, 3), np.arange(9, dtype=np.uint64), (1, 1), 1, "chebyshev", False)) == [0, 1, 3, 4]


_GLOO = r'''
import os, sys, ctypes as C
import numpy as np, torch.distributed as dist
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
import vahana_b200 as vh
from models import edges_model, foos
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()


class Recorder:
    """stands in for the engine library: hands out ids of this rank and records what the mirror adds"""
    def __init__(self):
        self.next = {{}}; self.agents = {{}}; self.edges = {{}}; self.offsets = {{}}; self.created = 0; self.finished = 0
    def vb_comm_rank(self, r, w):
        r._obj.value, w._obj.value = rank, world; return 0
    def vb_sim_create(self, md, params, h):
        self.created += 1; self.next = {{}}; self.agents = {{}}; self.edges = {{}}; h._obj.value = 1000 + self.created; return 0
    def vb_set_config(self, *a): return 0
    def vb_sim_destroy(self, h): return 0
    def vb_add_agents(self, h, tid, buf, n, ids):
        tid, n = tid.value, n.value
        first = self.next.get(tid, 1); self.next[tid] = first + n
        out = (C.c_uint64 * n).from_address(ids.value)
        for k in range(n): out[k] = vh.agent_id(tid, rank, first + k)
        st = None if buf is None else np.frombuffer((C.c_uint8 * (n * 8)).from_address(buf.value), dtype=np.int64).copy()
        self.agents.setdefault(tid, []).append((n, st)); return 0
    def vb_add_edges(self, h, e, fr, to, st, n):
        n = n.value
        f = np.frombuffer((C.c_uint64 * n).from_address(fr.value), dtype=np.uint64).copy()
        t = np.frombuffer((C.c_uint64 * n).from_address(to.value), dtype=np.uint64).copy()
        self.edges.setdefault(e.value, []).append((f, t)); return 0
    def vb_set_uniform_offset(self, h, tid, off): self.offsets[tid.value] = off.value; return 0
    def vb_add_raster(self, h, name, nd, dims, tid, buf, ids):
        n = int(np.prod([dims[k] for k in range(nd.value)]))
        return self.vb_add_agents(h, tid, buf, C.c_uint64(n), ids)
    def vb_connect_raster_neighbors(self, *a): return 0
    def vb_set_raster(self, h, name, nd, dims, tid, ids):
        n = int(np.prod([dims[k] for k in range(nd.value)]))
        self.rasters = getattr(self, "rasters", {{}})
        self.rasters[name.decode()] = (tuple(dims[k] for k in range(nd.value)), tid.value, np.frombuffer((C.c_uint64 * n).from_address(ids.value), dtype=np.uint64).copy())
        return 0
    def vb_finish_init(self, h): self.finished += 1; return 0
    def vb_last_error(self): return b""


class FakeBackend:
    name = "recorder"
    def __init__(self): self.lib = Recorder()
    def init(self, device=0): pass
    def check(self, rc): assert rc == 0


be = FakeBackend()
sim = vh.create_simulation(edges_model(), backend=be)
assert sim._stage is not None
# every rank runs the same initialisation code; only rank 0's content counts (finish_init! docs, src/Simulation.jl:393-398)
n = 11
ids = sim.add_agents("Agent", foos(range(1, n + 1) if rank == 0 else range(100, 100 + n)))
fr, to = ids, np.roll(ids, -1)
sim.add_edges(fr, to, "EdgeS")
rank0_ids = np.array([vh.agent_id(1, 0, k) for k in range(1, n + 1)], dtype=np.uint64)
m = sim.finish_init(return_idmapping=True, partition_algo="EqualAgentNumbers")
b = vh.equal_partition(n, world)
lib = be.lib
assert lib.created == 2 and lib.finished == 1                       # the initialisation phase was replaced by the shard
cnt, st = lib.agents[1][0]
assert cnt == b[rank + 1] - b[rank] and np.array_equal(st, np.arange(b[rank] + 1, b[rank + 1] + 1)), (cnt, st)   # rank 0's states
assert lib.offsets[1] == b[rank]
assert len(m) == n
for k, old in enumerate(rank0_ids):
    owner = int(np.searchsorted(np.array(b[1:]), k, side="right"))
    assert m[int(old)] == vh.agent_id(1, owner, k - b[owner] + 1)
f, t = lib.edges[sim._eid["EdgeS"]][0]
mine = [k for k in range(n) if vh.process_nr(m[int(rank0_ids[(k + 1) % n])]) == rank]
assert np.array_equal(t, [m[int(rank0_ids[(k + 1) % n])] for k in mine]) and np.array_equal(f, [m[int(rank0_ids[k])] for k in mine])
# join (src/MPI.jl:481-517) behind all_agents / all_agentids / all_edges(all_ranks = true): rank order, on every rank
j = sim._join(np.arange(rank + 2) + 10 * rank)
assert np.array_equal(j, np.concatenate([np.arange(r + 2) + 10 * r for r in range(world)]))
# a raster: its cells are handed out like every agent, the stencil edges follow their targets, every rank receives the id grid
# (broadcastids, src/MPI.jl:59-73)
from models import gol_model
be3 = FakeBackend()
sim3 = vh.create_simulation(gol_model(), backend=be3)
dims = (6, 5)
init = (np.arange(30) % 3 == 0)
sim3.add_raster("grid", dims, "Cell", init.view([("active", "?")]))
sim3.connect_raster_neighbors("grid", "Neighbor")
m3 = sim3.finish_init(return_idmapping=True)
b3 = vh.equal_partition(30, world)
rdims, rtid, rids = be3.lib.rasters["grid"]
assert rdims == dims and rtid == 1
want = [vh.agent_id(1, int(np.searchsorted(np.array(b3[1:]), k, side="right")), k - b3[int(np.searchsorted(np.array(b3[1:]), k, side="right"))] + 1) for k in range(30)]
assert [int(x) for x in rids] == want                                   # the whole grid with the new ids, on every rank
cnt3, _st3 = be3.lib.agents[1][0]
assert cnt3 == b3[rank + 1] - b3[rank]
f3, t3 = be3.lib.edges[sim3._eid["Neighbor"]][0]
assert len(t3) == 8 * cnt3 and all(vh.process_nr(int(x)) == rank for x in t3)      # 8 stencil edges per local cell, stored with their target
fr_all, to_all = vh.raster_neighbor_edges(dims, np.array([vh.agent_id(1, 0, k + 1) for k in range(30)], dtype=np.uint64))
sel = [i for i in range(len(to_all)) if vh.process_nr(m3[int(to_all[i])]) == rank]
assert [int(x) for x in f3] == [m3[int(fr_all[i])] for i in sel] and [int(x) for x in t3] == [m3[int(to_all[i])] for i in sel]
# agents placed on a raster in the initialisation phase (move_to!, src/Raster.jl:437-477): both edge directions follow their targets
from models import raster_model
be4 = FakeBackend(); be4.lib.vb_move_to = lambda *a: 0
sim4 = vh.create_simulation(raster_model(), backend=be4)
grid4 = sim4.add_raster("raster", (4, 3), "Position", lambda p: (0,))
movers = [sim4.add_agent("MovingAgent", v) for v in range(1, 6)]
for k, a in enumerate(movers):
    sim4.move_to("raster", a, (1 + k % 4, 1 + k % 3), "OnPosition", "OnPosition", distance=1 if k == 2 else 0)
stage_fr = np.concatenate([c[0] for c in sim4._stage["edges"]["OnPosition"]]); stage_to = np.concatenate([c[1] for c in sim4._stage["edges"]["OnPosition"]])
assert len(stage_fr) == 2 * (4 + 9)
m4 = sim4.finish_init(return_idmapping=True)
stage_fr, stage_to = [vh.remove_process(x) for x in stage_fr], [vh.remove_process(x) for x in stage_to]      # the mapping is keyed by rank 0's ids (its content is what is handed out)
f4, t4 = be4.lib.edges[sim4._eid["OnPosition"]][0]
sel = [i for i in range(len(stage_to)) if vh.process_nr(m4[int(stage_to[i])]) == rank]
assert len(sel) > 0 and all(vh.process_nr(int(x)) == rank for x in t4)
assert [int(x) for x in f4] == [m4[int(stage_fr[i])] for i in sel] and [int(x) for x in t4] == [m4[int(stage_to[i])] for i in sel]
assert [int(x) for x in be4.lib.rasters["raster"][2]] == [m4[vh.remove_process(x)] for x in grid4.reshape(-1, order="F")]
# device-side bulk adds cannot be handed out
sim2 = vh.create_simulation(edges_model(), backend=be)
sim2._unstageable = "add_agents_device"
try:
    sim2.finish_init()
    raise SystemExit("expected an AssertionError")
except AssertionError:
    pass
sim2.finish_init(distribute=False)
print("gloo rank", rank, "ok", flush=True)
dist.destroy_process_group()
'''


def test_gloo_world2_distribute_plumbing(tmp_path):
    script = tmp_path / "gloo_distribute.py"
    script.write_text(_GLOO.format(root=ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29531", str(script)], capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert r.stdout.count("ok") == 2
