"""The SIR contact model (BASELINE config 5; SURVEY.md Appendix C — Episim itself is an external repository, so no reference test
pins its numbers): the oracle is checked bit for bit against an independent vectorised numpy restatement (own Philox4x32-10, own
container logic: bincount / stable argsort instead of append logs), so the fixtures under tests/golden/ and the GPU parity tests do not
rest on the oracle alone."""
import math

import numpy as np
import pytest

from models import sir_sim, sir_step
from test_zzm_market import philox_uniform


def numpy_sir_step(state, days, nl, step, beta=0.3, infectious_days=10):
    """one step of vahana.jl_b200/csrc/transitions/sir.h; returns the new person columns and what the containers must hold"""
    n = state.shape[0]
    slot = np.arange(n)
    # visit: two Visit edges per person, location = floor(u_k * n_locations); append order = person, then k
    loc = np.stack([np.minimum((philox_uniform(4 * step, slot, k) * nl).astype(np.int64), nl - 1) for k in range(2)], axis=1)
    infectious = state == 1
    to, fr, inf = loc.reshape(-1), np.repeat(slot, 2), np.repeat(infectious, 2)
    order = np.argsort(to, kind="stable")                        # a location's row: its visits in append order
    nv = np.bincount(to, minlength=nl)
    # tally
    n_inf = np.bincount(to, weights=inf, minlength=nl).astype(np.int32)
    # expose: every visit of a location with visitors gets Exposure(n_inf / n_visitors) in Float32; a person's row: ascending location
    with np.errstate(invalid="ignore", divide="ignore"):
        risk_loc = n_inf.astype(np.float32) / nv.astype(np.float32)
    lo, hi = loc.min(axis=1), loc.max(axis=1)
    rows = np.stack([risk_loc[lo], risk_loc[hi]], axis=1)
    # infect
    risk = (np.float32(0) + rows[:, 0]) + rows[:, 1]
    pinf = np.array([1.0 - math.exp(-beta * float(r)) for r in risk])
    u = philox_uniform(4 * step + 3, slot, 0)
    s, d = state.copy(), days.copy()
    catch = (state == 0) & (u < pinf)
    s[catch], d[catch] = 1, 0
    ill = state == 1
    d[ill] = days[ill] + 1
    s[ill & (d >= infectious_days)] = 2
    return s, d, dict(visit_off=np.concatenate([[0], np.cumsum(nv)]), visit_from=fr[order], visit_inf=inf[order], n_inf=n_inf, exposure=rows)


@pytest.mark.parametrize("n,nl,frac", [(5000, 400, 0.01), (3000, 7, 0.2), (400, 3000, 0.05)])      # typical, crowded locations, mostly empty locations
def test_sir_oracle_vs_numpy(oracle, n, nl, frac):
    sim = sir_sim(oracle, n, nl, beta=0.3, frac=frac)
    p = sim.all_agents("Person")
    state, days = p["state"].copy(), p["days"].copy()
    seen = set()
    for step in range(14):
        sir_step(sim, step)
        state, days, c = numpy_sir_step(state, days, nl, step)
        p = sim.all_agents("Person")
        assert np.array_equal(p["state"], state) and np.array_equal(p["days"], days), step
        assert np.array_equal(sim.all_agents("Location")["n_inf"], c["n_inf"])
        off, fr, st = sim.export_csr("Visit", "Location", nl)
        assert np.array_equal(np.asarray(off, dtype=np.int64), c["visit_off"])
        assert np.array_equal((fr & np.uint64((1 << 36) - 1)).astype(np.int64) - 1, c["visit_from"])
        assert np.array_equal(st["infectious"], c["visit_inf"])
        off, _, st = sim.export_csr("Exposure", "Person", n)
        assert np.array_equal(np.asarray(off, dtype=np.int64), 2 * np.arange(n + 1))
        assert np.array_equal(st["risk"].view("u4"), c["exposure"].reshape(-1).view("u4"))          # Float32 risks bit for bit
        seen.update(int(x) for x in np.unique(state))
    assert seen == {0, 1, 2}


def test_sir_golden_fixture_vs_numpy():
    """tests/golden/sir_3000.npz (written from the oracle) reproduced by the numpy restatement alone: the fixture the GPU suite
    replays is pinned by two independent implementations"""
    import os
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.join(here, "golden"))
    from make_golden import SIR
    from models import PERSON
    g = np.load(os.path.join(here, "golden", "sir_3000.npz"))
    state = (np.random.default_rng(7).random(SIR["n"]) < 0.01).astype("u1")        # sir_sim's defaults (seed 7, 1 % infectious)
    days = np.zeros(SIR["n"], dtype="u1")
    assert PERSON[0][0] == "state"
    for step in range(SIR["steps"]):
        state, days, c = numpy_sir_step(state, days, SIR["nl"], step, beta=SIR["beta"])
        assert [int((state == k).sum()) for k in range(3)] == g["counts"][step].tolist(), step
    assert np.array_equal(state, g["state"]) and np.array_equal(days, g["days"]) and np.array_equal(c["n_inf"], g["n_inf"])
