"""calc_rasterstate / mapreduce with registered map functors on the docs' predator/prey raster (docs/examples/predator.jl:487-489,
520-523): out[i] = f(state of cell i) against the field read-out.  Oracle in the CPU suite; the GPU variant sorts behind the
established suite (written after round 1's GPU budget was spent)."""
import numpy as np
import pytest

import vahana_b200 as vh
from models import pp_sim, pp_step


def _checks(backend):
    sim = pp_sim(backend, (12, 9), 30, 8)
    for step in range(3):
        pp_step(sim, step)
        cd = sim.rastervalues("raster", "countdown", "Cell")
        food = sim.calc_rasterstate_fn("raster", "pp_has_food", "?")
        assert food.shape == (12, 9) and food.dtype == np.bool_ and np.array_equal(food, cd == 0)
        assert np.array_equal(sim.calc_rasterstate_fn("raster", "pp_has_food", "i8"), (cd == 0).astype("i8"))
        assert np.array_equal(sim.calc_rasterstate_fn("raster", "pp_growth_progress", "f8"), 1.0 / (1.0 + cd.astype("f8")))
        assert sim.mapreduce_fn("pp_has_food", "+", "Cell", datatype="i8") == int((cd == 0).sum()) == sim.mapreduce("countdown", "+", "Cell", equals=0)
        np.testing.assert_allclose(sim.mapreduce_fn("pp_growth_progress", "+", "Cell"), (1.0 / (1.0 + cd)).sum(), rtol=1e-12)
    with pytest.raises(ValueError):
        sim.calc_rasterstate_fn("raster", "pp_has_food", "f8")            # integral functor, float result
    with pytest.raises(ValueError):
        sim.calc_rasterstate_fn("raster", "market_revenue", "f8")         # registered for another type


def test_raster_maps_oracle(oracle):
    _checks(oracle)


@pytest.mark.gpu
def test_raster_maps_gpu(cuda):
    _checks(cuda)


def test_calc_raster_general_form_and_lazy_accessors(oracle):
    """calc_raster(sim, raster, f, f_returns, accessible) with a host closure (src/Raster.jl:199-236: the docs' Game of Life read-out and
    the test/raster.jl:296 edge count), neighborstates_iter / neighborstates_flexible_iter (src/EdgeMethods.jl:783-803) and
    checked (src/Helpers.jl:33-37) — host-side conveniences of the reference API that need no kernel"""
    from models import raster_model
    sim = vh.create_simulation(raster_model(), backend=oracle)
    sim.add_raster("raster", (6, 4), "Position", lambda p: (10 * p[0] + p[1],))
    movers = [sim.add_agent("MovingAgent", v) for v in (1, 2, 3)]
    for a, p in zip(movers, [(1, 1), (2, 2), (2, 2)]):
        sim.move_to("raster", a, p, "OnPosition", "OnPosition")
    with pytest.raises(AssertionError):
        sim.calc_raster("raster", lambda cid: 0, "i8")
    sim.finish_init()
    ids = sim.raster_ids("raster")
    assert ids.shape == (6, 4) and int(ids[1, 2]) == sim.cellid("raster", (2, 3))
    got = sim.calc_raster("raster", lambda cid: sim.agentstate(cid, "Position")["ids_sum"], "i8", ["Position"])
    assert np.array_equal(got, sim.calc_rasterstate("raster", "ids_sum", "Position")) and got[5, 3] == 64
    ne = sim.calc_raster("raster", lambda cid: sim.num_edges(cid, "OnPosition"), "i8", ["OnPosition"])
    assert np.array_equal(ne, sim.calc_raster_num_edges("raster", "OnPosition")) and ne[1, 1] == 2 and ne.sum() == 3
    # lazy accessors: the movers standing on a cell
    sim.disable_transition_checks(True)
    cell = int(ids[1, 1])
    it = sim.neighborstates_iter(cell, "OnPosition", "MovingAgent")
    assert not isinstance(it, list) and [int(s["value"]) for s in it] == [2, 3]
    assert [int(s["value"]) for s in sim.neighborstates_flexible_iter(cell, "OnPosition")] == [2, 3]
    assert sim.neighborstates_iter(int(ids[3, 3]), "OnPosition", "MovingAgent") is None
    # checked: nothing happens for `nothing`
    seen = []
    vh.checked(seen.append, lambda f, it_: [f(x) for x in it_], sim.neighborids(int(ids[3, 3]), "OnPosition"))
    assert seen == []
    vh.checked(seen.append, lambda f, it_: [f(x) for x in it_], sim.neighborids(cell, "OnPosition"))
    assert seen == [int(movers[1]), int(movers[2])]
    sim.disable_transition_checks(False)
    assert vh.rootonly(lambda: 7) == 7
