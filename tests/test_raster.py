"""Replays /root/reference/test/raster.jl: add_raster!, connect_raster_neighbors! (metrics, periodic / clipped, 2-4 dims),
calc_rasterstate / rastervalues, move_to! with surroundings."""
import numpy as np
import pytest

import vahana_b200 as vh
from models import raster_model


def _grid(backend, dims, hot, T="GridA", **kw):
    sim = vh.create_simulation(raster_model(), backend=backend)
    sim.add_raster("grid", dims, T, lambda p: (p, p == hot))
    sim.connect_raster_neighbors("grid", "GridE", **kw)
    sim.finish_init()
    return sim


@pytest.mark.parametrize("dims,hot", [((10, 8), (10, 8)), ((7, 9), (3, 5)), ((17, 6), (9, 1)), ((6, 7), (1, 1))])
def test_diffuse_2d(backend, dims, hot):  # test/raster.jl:31-105: 1 -> 9 -> 25 on a torus
    sim = _grid(backend, dims, hot)
    assert sim.num_edges("GridE") == 8 * dims[0] * dims[1]
    assert sim.mapreduce("active", "+", "GridA", datatype="i8") == 1
    sim.apply("diffuse", ["GridA"], ["GridA", "GridE"], ["GridA"])
    assert sim.mapreduce("active", "+", "GridA", datatype="i8") == 9
    sim.apply("diffuse", ["GridA"], ["GridA", "GridE"], ["GridA"])
    assert sim.mapreduce("active", "+", "GridA", datatype="i8") == 25
    if dims == (10, 8):
        r = sim.calc_rasterstate("grid", "active", "GridA")
        assert r.shape == (10, 8)
        assert not r[0, 2] and r[0, 0] and not r[3, 3]
        assert r.sum() == 25
        pos = sim.rastervalues("grid", "pos", "GridA")
        assert pos.shape == (10, 8, 2) and pos[3, 5].tolist() == [4, 6]      # cells are created in CartesianIndices order


@pytest.mark.parametrize("kw,expect", [
    (dict(), [27, 125]),
    (dict(periodic=False), [8, 27]),
    (dict(periodic=False, distance=2), [27, 125]),
    (dict(distance=1.5, metric="euclidean"), [27 - 8]),
    (dict(distance=1.5, metric="manhatten"), [7]),
    (dict(periodic=False, distance=2, metric="manhatten"), [10]),
])
def test_diffuse_3d(backend, kw, expect):  # test/raster.jl:112-228
    sim = _grid(backend, (6, 7, 6), (1, 1, 1), "Grid3D", **kw)
    assert sim.mapreduce("active", "+", "Grid3D", datatype="i8") == 1
    for e in expect:
        sim.apply("diffuse", ["Grid3D"], ["Grid3D", "GridE"], ["Grid3D"])
        assert sim.mapreduce("active", "+", "Grid3D", datatype="i8") == e


def test_raster_nodeid(backend):  # test/raster.jl:242-303
    sim = vh.create_simulation(raster_model(), backend=backend)
    sim.add_raster("raster", (10, 10), "Position", lambda p: (0,))
    sim.connect_raster_neighbors("raster", "GridE")
    p1, p2, p3 = (sim.add_agent("MovingAgent", v) for v in (1, 2, 3))
    sim.move_to("raster", p1, (1, 1), "OnPosition", "OnPosition")
    sim.move_to("raster", p2, (2, 2), "OnPosition", "OnPosition")
    sim.move_to("raster", p3, (2, 2), "OnPosition", "OnPosition")
    sim.finish_init()
    sim.apply("sum_on_pos", ["Position"], ["MovingAgent", "OnPosition"], ["Position"])
    r = sim.calc_rasterstate("raster", "ids_sum", "Position")
    assert r[0, 0] == 1 and r[0, 1] == 0 and r[1, 1] == 5 and r[1, 2] == 0
    sim.apply("value_on_pos", ["MovingAgent"], ["Position", "OnPosition"], ["OnPosition", "MovingAgent"])
    sim.apply("sum_on_pos", ["Position"], ["MovingAgent", "OnPosition"], ["Position"])
    r = sim.calc_rasterstate("raster", "ids_sum", "Position")
    assert r[0, 0] == 1 and r[0, 1] == 0 and r[1, 1] == 0 and r[4, 4] == 10
    assert sim.cellid("raster", (5, 5)) == vh.agent_id(3, 0, 45)
    ne = sim.calc_raster_num_edges("raster", "OnPosition")
    assert ne[0, 0] == 1 and ne[4, 4] == 2 and ne.sum() == 3


def test_move_to_dist(backend):  # test/raster.jl:305-345: 25 / 13 / 21, 4-D 81 / 9
    sim = vh.create_simulation(raster_model(), backend=backend)
    sim.add_raster("raster", (10, 10), "Position", lambda p: (0,))
    p1, p2, p3 = (sim.add_agent("MovingAgent", 1) for _ in range(3))
    sim.move_to("raster", p1, (4, 4), "OnPosition", None, distance=2)
    sim.move_to("raster", p2, (4, 4), "OnPosition", None, distance=2, metric="manhatten")
    sim.move_to("raster", p3, (4, 4), "OnPosition", None, distance=2.5, metric="euclidean")
    sim.finish_init()
    sim.disable_transition_checks(True)
    assert sim.num_edges(p1, "OnPosition") == 25
    assert sim.num_edges(p2, "OnPosition") == 13
    assert sim.num_edges(p3, "OnPosition") == 21
    sim.disable_transition_checks(False)
    sim = vh.create_simulation(raster_model(), backend=backend)
    sim.add_raster("raster", (10, 10, 10, 10), "Position", np.zeros(10000, dtype=[("ids_sum", "i8")]))
    p1, p2 = (sim.add_agent("MovingAgent", 1) for _ in range(2))
    sim.move_to("raster", p1, (4, 4, 4, 4), "OnPosition", None, distance=1)
    sim.move_to("raster", p2, (4, 4, 4, 4), "OnPosition", None, distance=1, metric="manhatten")
    sim.finish_init()
    sim.disable_transition_checks(True)
    assert sim.num_edges(p1, "OnPosition") == 81
    assert sim.num_edges(p2, "OnPosition") == 9
    sim.disable_transition_checks(False)


@pytest.mark.gpu
@pytest.mark.parametrize("dims,kw", [((9, 7), dict()), ((9, 7), dict(periodic=False)), ((5, 4), dict(distance=2)),
                                      ((12, 10), dict(distance=2, metric="manhatten", periodic=False)), ((3, 3), dict())])
def test_implicit_stencil_row_order_matches_oracle(oracle, cuda, dims, kw):
    """The CUDA engine keeps connect_raster_neighbors! edges implicit (grid-stencil path); rows must still list the sources in
    the reference's insertion order, including wrapped borders, clipped borders and the duplicates of tiny periodic rasters."""
    sims = []
    for be in (cuda, oracle):
        sim = vh.create_simulation(raster_model(), backend=be)
        sim.add_raster("grid", dims, "GridA", lambda p: (p, False))
        sim.connect_raster_neighbors("grid", "GridE", **kw)
        sim.finish_init()
        sims.append(sim)
    g, o = sims
    n = dims[0] * dims[1]
    assert g.num_edges("GridE") == o.num_edges("GridE")
    a, b = g.export_csr("GridE", "GridA", n), o.export_csr("GridE", "GridA", n)          # host-side enumeration
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    for sim in sims:                                                                       # device-side enumeration
        sim.apply("grid_probe", "GridA", ["GridA", "GridE"], "GridA")
    assert np.array_equal(g.all_agents("GridA")["pos"], o.all_agents("GridA")["pos"])
    ids = g.all_agentids("GridA")
    g.disable_transition_checks(True); o.disable_transition_checks(True)
    for i in (0, n // 2, n - 1):
        assert g.neighborids(int(ids[i]), "GridE") == o.neighborids(int(ids[i]), "GridE")
    # an explicit add turns the implicit container into ordinary rows, in the same order
    g.add_edge(int(ids[0]), int(ids[1]), "GridE"); o.add_edge(int(ids[0]), int(ids[1]), "GridE")
    a, b = g.export_csr("GridE", "GridA", n), o.export_csr("GridE", "GridA", n)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
