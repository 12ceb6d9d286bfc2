"""calc_rasterstate / mapreduce with registered map functors on the docs' predator/prey raster (docs/examples/predator.jl:487-489,
520-523): out[i] = f(state of cell i) against the field read-out.  Oracle in the CPU suite; the GPU variant sorts behind the
established suite (written after round 1's GPU budget was spent)."""
import numpy as np
import pytest

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
