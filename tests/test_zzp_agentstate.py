"""test/mpi/test_agentstate.jl on one rank (the 2-4 rank run is tests/mgpu_agentstate.py): the checks of the reference's closures are
assertions inside the transitions (ctx.require), and a deliberately stale expectation must raise.  GPU variants sort behind the
established suite (written after round 1's GPU budget was spent)."""
import pytest

from models import agentstate_scenario


def _scenario(backend):
    for immortal in (True, False):
        sim = agentstate_scenario(backend, 6, immortal)
        # the negative: after another doubling the x1 check must fail inside the transition -> AssertionError from apply!
        sim.apply("as_double", ["ASAgent"], ["ASAgent"], ["ASAgent"])
        with pytest.raises(AssertionError):
            sim.apply("as_check_x1_EdgeState", ["ASAgent"], ["ASAgent", "EdgeState"], [])
        sim.apply("as_check_x2_EdgeState", ["ASAgent"], ["ASAgent", "EdgeState"], [])
        with pytest.raises(AssertionError):
            sim.apply("as_require_no_NewEdge", ["ASAgent"], ["ASAgent", "NewEdge"], [])


def test_agentstate_scenario_oracle(oracle):
    _scenario(oracle)


@pytest.mark.gpu
def test_agentstate_scenario_gpu(cuda):
    _scenario(cuda)
