"""Long runs at the BASELINE configs' own sizes (run last: they take seconds, not milliseconds).
Config 3 (a): the docs predator/prey model at the docs size (100 x 100 cells, 2000 prey, 500 predators, 400 steps of six applies,
docs/examples/predator.jl:146-153,437-469) — all-integer, so every population, energy sum and the final agent tables are bit-exact."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from make_golden import PP_DOCS, pp_digest  # noqa: E402
from models import pp_globals, pp_sim, pp_step  # noqa: E402


def _row(step, g):
    return [step, g["prey_pop"], g["predator_pop"], g["cells_with_food"], g["prey_energy"], g["predator_energy"]]


def test_predator_prey_docs_size_400_steps_golden(backend):
    """Oracle (CPU suite): the restatement reproduces the committed trajectory (tests/golden/pp_docs_100x100.npz).  CUDA engine (GPU
    suite): 2400 applies with births, deaths, slot reuse and seven edge types rebuilt or merged every step land on the same numbers."""
    gold = np.load(os.path.join(HERE, "golden", "pp_docs_100x100.npz"))
    sim = pp_sim(backend, PP_DOCS["dims"], PP_DOCS["nprey"], PP_DOCS["npred"])
    k = 0
    for step in range(PP_DOCS["steps"]):
        pp_step(sim, step)
        if step % PP_DOCS["every"] == PP_DOCS["every"] - 1:
            assert _row(step, pp_globals(sim)) == gold["trajectory"][k].tolist(), step
            k += 1
    assert k == len(gold["trajectory"])
    assert pp_digest(sim) == str(gold["digest"])
    pops = gold["trajectory"][:, 1:3]
    assert pops.min() > 0 and len(np.unique(pops[:, 0])) > 10          # both species survive and oscillate
