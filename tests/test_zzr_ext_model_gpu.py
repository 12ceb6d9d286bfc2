"""GPU half of tests/test_ext_model.py: the externally built transition and map are applied through the engine (written after
round 1's GPU budget was spent; sorts behind the established suite)."""
import pytest

from models import createsim
from test_ext_model import build_ext_model


@pytest.mark.gpu
def test_external_model_library_applies(cuda):
    cuda.load_model_library(build_ext_model())
    sim, a1, a2, a3, avids, avfids = createsim(cuda)
    sim.apply("ext_add_one_plus_degree", "AMortal", ["AMortal", "ESLDict1"], "AMortal")
    assert sorted(sim.all_agents("AMortal")["foo"].tolist()) == [3, 4, 12]          # 1 + 1 + 10 neighbours, 2 + 1, 3 + 1
    assert sim.mapreduce_fn("ext_square", "+", "AMortal", datatype="i8") == 9 + 16 + 144
