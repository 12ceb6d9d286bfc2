"""The docs' predator/prey model (BASELINE config 3, /root/reference/docs/examples/predator.jl) on a second, independent
implementation: a small pure-Python engine that keeps the reference's own container shapes — a `dict` target id -> list of source ids
per edge type (`Dict{AgentID, Vector{AgentID}}`, src/EdgeMethods.jl:9-41), agent vectors with a died flag and a LIFO list of reusable
numbers (src/AgentMethods.jl:37-63, 300-439) — and replays step! apply by apply with the documented definitions of the random
choices (vahana.jl_b200/csrc/transitions/predator.h).  The C++ oracle (append logs, sort, CSR) must agree with it bit for bit after
every step: agent tables with their ids (slot reuse), all seven edge containers row by row (push! order, add_existing merges, the
purge of edges of dead agents), the globals.  No reference fixture exists for this model (its trajectories depend on Julia's RNG);
this pins the oracle — and through it the fixtures under tests/golden/ and the GPU parity tests — on two implementations."""
import numpy as np
import pytest

import vahana_b200 as vh
from models import PP_EDGES, pp_globals, pp_sim, pp_step

PRED, PREY, CELL = 1, 2, 3
M32 = 0xFFFFFFFF
PARAMS = dict(restart=5, pred_gain=5, pred_loss=1, pred_thres=5, pred_prob=20, prey_gain=5, prey_loss=1, prey_thres=5, prey_prob=20)


def uniform(seed, slot, k):
    """Philox4x32-10, counter (slot lo, slot hi, k lo, k hi), key (seed lo, seed hi); 53 bits of the first two words"""
    c = [slot & M32, slot >> 32, k & M32, k >> 32]
    k0, k1 = seed & M32, (seed >> 32) & M32
    for _ in range(10):
        p0, p1 = 0xD2511F53 * c[0], 0xCD9E8D57 * c[2]
        c = [(p1 >> 32) ^ c[1] ^ k0, p1 & M32, (p0 >> 32) ^ c[3] ^ k1, p0 & M32]
        k0, k1 = (k0 + 0x9E3779B9) & M32, (k1 + 0xBB67AE85) & M32
    return float(((c[0] << 32) | c[1]) >> 11) * (1.0 / 9007199254740992.0)


def pick(u, n):
    return min(int(u * n), n - 1)


class World:
    """agents[T]: nr -> state (living agents only), reuse[T]: LIFO list, nextnr[T]; edges[name]: target id -> [source ids]"""

    def __init__(self, dims):
        self.dims = dims
        self.agents = {PRED: {}, PREY: {}, CELL: {}}
        self.reuse = {PRED: [], PREY: [], CELL: []}
        self.nextnr = {PRED: 1, PREY: 1, CELL: 1}
        self.edges = {name: {} for name in PP_EDGES}

    def new_agent(self, T, state, into):
        nr = self.reuse[T].pop() if self.reuse[T] else self.nextnr[T]       # _get_next_id: reusable numbers first, last freed first
        if nr == self.nextnr[T]:
            self.nextnr[T] += 1
        into[nr] = state
        return vh.agent_id(T, 0, nr)

    def cell_at(self, pos):
        x, y = (pos[0] - 1) % self.dims[0], (pos[1] - 1) % self.dims[1]
        return vh.agent_id(CELL, 0, 1 + x + y * self.dims[0])

    def place(self, out, aid, pos, species):
        """move!(sim, id, pos, Species), predator.jl:205-209: one Position edge, then View edges both ways for pos and its four neighbours"""
        out[f"Position{{{species}}}"].setdefault(self.cell_at(pos), []).append(aid)
        view = out[f"View{{{species}}}"]
        for dx, dy in ((0, 0), (0, -1), (-1, 0), (1, 0), (0, 1)):
            c = self.cell_at((pos[0] + dx, pos[1] + dy))
            view.setdefault(aid, []).append(c)
            view.setdefault(c, []).append(aid)

    def finish(self, written_agents, new_agents, written_edges, new_edges, keep=()):
        """finish_write!: written containers replace the read ones (add_existing: old rows first, new entries behind), numbers of
        agents that died become reusable in ascending order, edges from or to an agent that died are dropped everywhere"""
        dead = set()
        for T in written_agents:
            gone = sorted(nr for nr in self.agents[T] if nr not in new_agents[T])
            self.reuse[T] += gone
            dead.update(vh.agent_id(T, 0, nr) for nr in gone)
            self.agents[T] = new_agents[T]
        for name in written_edges:
            if name in keep:
                for to, fr in new_edges[name].items():
                    self.edges[name].setdefault(to, []).extend(fr)
            else:
                self.edges[name] = new_edges[name]
        if dead:
            for name, rows in self.edges.items():
                self.edges[name] = {to: kept for to, fr in rows.items() if to not in dead for kept in [[f for f in fr if f not in dead]] if kept}

    # ---- step!(sim), predator.jl:437-469 ----
    def move(self, species, seed):
        T = PREY if species == "Prey" else PRED
        loss = PARAMS["prey_loss"] if T == PREY else PARAMS["pred_loss"]
        view, cells = self.edges[f"View{{{species}}}"], self.agents[CELL]
        new, out = {}, {f"View{{{species}}}": {}, f"Position{{{species}}}": {}}
        for nr in sorted(self.agents[T]):
            energy, pos = self.agents[T][nr]
            aid = vh.agent_id(T, 0, nr)
            e = energy - loss
            if e <= 0:
                continue
            u = uniform(seed, nr - 1, 0)
            row = view[aid]
            if T == PREY:       # predator.jl:294-310: a random visible cell with grass, any visible cell if there is none
                grass = [c for c in row if cells[vh.agent_nr(c)][1] == 0]
                nxt = grass[pick(u, len(grass))] if grass else row[pick(u, len(row))]
                pos = cells[vh.agent_nr(nxt)][0]
            else:               # predator.jl:268-285: towards a random visible prey, else a random visible cell
                prey = self.edges["VisiblePrey"].get(aid, [])
                pos = self.agents[PREY][vh.agent_nr(prey[pick(u, len(prey))])][1] if prey else cells[vh.agent_nr(row[pick(u, len(row))])][0]
            new[nr] = (e, pos)
            self.place(out, aid, pos, species)
        self.finish([T], {T: new}, list(out), out)

    def find_prey(self):       # predator.jl:252-260
        out = {}
        pos_prey, view_pred = self.edges["Position{Prey}"], self.edges["View{Predator}"]
        for nr in sorted(self.agents[CELL]):
            cid = vh.agent_id(CELL, 0, nr)
            if cid in pos_prey and cid in view_pred:
                for prey in pos_prey[cid]:
                    for pred in view_pred[cid]:
                        out.setdefault(pred, []).append(prey)
        self.finish([], {}, ["VisiblePrey"], {"VisiblePrey": out})

    def grow_food(self):       # predator.jl:318-320
        self.agents[CELL] = {nr: (pos, cd - 1 if cd > 1 else 0) for nr, (pos, cd) in self.agents[CELL].items()}

    def try_eat(self, seed):   # predator.jl:330-354
        die, eat, new = {}, {}, {}
        for nr in sorted(self.agents[CELL]):
            pos, countdown = self.agents[CELL][nr]
            cid = vh.agent_id(CELL, 0, nr)
            preds = self.edges["Position{Predator}"].get(cid, [])
            prey = self.edges["Position{Prey}"].get(cid, [])[:64]
            left = list(range(len(prey)))                                   # indices of the prey nobody has eaten yet, in row order
            if preds and prey:
                order = sorted(range(len(preds)), key=lambda i: (uniform(seed, nr - 1, 16 + i), i))        # shuffle(predators)
                for t, bi in enumerate(order):
                    if not left:
                        break
                    idx = left.pop(pick(uniform(seed, nr - 1, min(1 + t, 15)), len(left)))               # rand(prey)
                    die.setdefault(prey[idx], []).append(cid)
                    eat.setdefault(preds[bi], []).append(cid)
            if left and countdown == 0:                                      # a prey that is left eats the grass
                idx = left[pick(uniform(seed, nr - 1, 0), len(left))]
                eat.setdefault(prey[idx], []).append(cid)
                countdown = PARAMS["restart"]
            new[nr] = (pos, countdown)
        self.finish([CELL], {CELL: new}, ["Die", "Eat"], {"Die": die, "Eat": eat})

    def try_reproduce(self, seed):     # predator.jl:364-392, called for Predator, then Prey; the Position / View types are add_existing
        keep = ["Position{Predator}", "Position{Prey}", "View{Predator}", "View{Prey}"]
        out = {name: {} for name in keep}
        new = {PRED: {}, PREY: {}}
        for T, species, sp in ((PRED, "Predator", "pred"), (PREY, "Prey", "prey")):
            for nr in sorted(self.agents[T]):
                energy, pos = self.agents[T][nr]
                aid = vh.agent_id(T, 0, nr)
                if T == PREY and aid in self.edges["Die"]:
                    continue
                if aid in self.edges["Eat"]:
                    energy += PARAMS[f"{sp}_gain"]
                if energy > PARAMS[f"{sp}_thres"] and uniform(seed, nr - 1, 0) * 100.0 < PARAMS[f"{sp}_prob"]:
                    child = int(round(energy / 2))                           # Int64(round(energy / 2)): halves go to the even neighbour
                    cid = self.new_agent(T, (child, pos), new[T])
                    self.place(out, cid, pos, species)
                    energy -= child
                new[T][nr] = (energy, pos)
        self.finish([PRED, PREY], {PRED: new[PRED], PREY: new[PREY]}, keep, out, keep=keep)

    def step(self, step):
        s = 6 * step
        self.move("Prey", s)
        self.find_prey()
        self.move("Predator", s + 2)
        self.grow_food()
        self.try_eat(s + 4)
        self.try_reproduce(s + 5)


def build_world(dims, nprey, npred, seed=3):
    """the draws of models.pp_sim (predator.jl:185-229 with numpy's generator), in its order"""
    rng = np.random.default_rng(seed)
    n = dims[0] * dims[1]
    countdown = np.where(rng.random(n) < 0.5, 0, rng.integers(1, 6, n))
    w = World(dims)
    for k in range(n):
        w.new_agent(CELL, ((k % dims[0] + 1, k // dims[0] + 1), int(countdown[k])), w.agents[CELL])
    for species, T, count in (("Prey", PREY, nprey), ("Predator", PRED, npred)):
        for _ in range(count):
            pos = (int(rng.integers(1, dims[0] + 1)), int(rng.integers(1, dims[1] + 1)))
            aid = w.new_agent(T, (int(rng.integers(1, 11)), pos), w.agents[T])
            w.place(w.edges, aid, pos, species)
    return w


def compare(sim, w, where):
    names = {PRED: "Predator", PREY: "Prey", CELL: "Cell"}
    for T, name in names.items():
        ids = [int(x) for x in sim.all_agentids(name)]
        assert ids == [vh.agent_id(T, 0, nr) for nr in sorted(w.agents[T])], (where, name)
        a = sim.all_agents(name)
        if T == CELL:
            assert [(tuple(int(v) for v in r["pos"]), int(r["countdown"])) for r in a] == [w.agents[T][nr] for nr in sorted(w.agents[T])], (where, name)
        else:
            assert [(int(r["energy"]), tuple(int(v) for v in r["pos"])) for r in a] == [w.agents[T][nr] for nr in sorted(w.agents[T])], (where, name)
    sim.disable_transition_checks(True)
    for ename in PP_EDGES:
        rows = w.edges[ename]
        assert sim.num_edges(ename) == sum(len(v) for v in rows.values()), (where, ename)
        for T, name in names.items():
            nrows = w.nextnr[T] - 1
            off, fr, _ = sim.export_csr(ename, name, nrows)
            for nr in range(1, nrows + 1):
                got = [int(x) for x in fr[int(off[nr - 1]):int(off[nr])]]
                assert got == rows.get(vh.agent_id(T, 0, nr), []), (where, ename, name, nr)
    sim.disable_transition_checks(False)


@pytest.mark.parametrize("dims,nprey,npred,steps", [((30, 30), 180, 45, 20), ((6, 5), 60, 25, 30), ((17, 23), 40, 90, 25)])
def test_predator_prey_oracle_vs_python_restatement(oracle, dims, nprey, npred, steps):
    sim = pp_sim(oracle, dims, nprey, npred)
    w = build_world(dims, nprey, npred)
    compare(sim, w, "init")
    births = deaths = 0
    for step in range(steps):
        before = {T: set(w.agents[T]) for T in (PRED, PREY)}
        pp_step(sim, step)
        w.step(step)
        compare(sim, w, step)
        for T in (PRED, PREY):
            births += len(set(w.agents[T]) - before[T])
            deaths += len(before[T] - set(w.agents[T]))
        g = pp_globals(sim)
        assert g["prey_pop"] == len(w.agents[PREY]) and g["predator_pop"] == len(w.agents[PRED])
        assert g["cells_with_food"] == sum(1 for _, cd in w.agents[CELL].values() if cd == 0)
        assert g["prey_energy"] == sum(e for e, _ in w.agents[PREY].values())
    assert births > 10 and deaths > 10          # the run exercised reuse of freed numbers and the purge


def test_predator_prey_golden_fixture_vs_python_restatement():
    """tests/golden/pp_30x30.npz (written from the oracle) reproduced by the Python restatement alone"""
    import os
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.join(here, "golden"))
    from make_golden import PP
    g = np.load(os.path.join(here, "golden", "pp_30x30.npz"))
    w = build_world(PP["dims"], PP["nprey"], PP["npred"])
    for step in range(PP["steps"]):
        w.step(step)
        row = [len(w.agents[PREY]), len(w.agents[PRED]), sum(1 for _, cd in w.agents[CELL].values() if cd == 0),
               sum(e for e, _ in w.agents[PREY].values()), sum(e for e, _ in w.agents[PRED].values())]
        assert row == g["trajectory"][step].tolist(), step


def test_predator_prey_docs_size_fixture_vs_python_restatement():
    """BASELINE config 3 (a) — the docs' own size, 100 x 100 cells, 2000 prey, 500 predators — against tests/golden/pp_docs_100x100.npz:
    the first two samples (steps 25 and 50) of the 400-step trajectory the GPU suite replays, from the Python restatement alone"""
    import os
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.join(here, "golden"))
    from make_golden import PP_DOCS
    g = np.load(os.path.join(here, "golden", "pp_docs_100x100.npz"))
    w = build_world(PP_DOCS["dims"], PP_DOCS["nprey"], PP_DOCS["npred"])
    samples = 0
    for step in range(2 * PP_DOCS["every"]):
        w.step(step)
        if step % PP_DOCS["every"] == PP_DOCS["every"] - 1:
            row = [step, len(w.agents[PREY]), len(w.agents[PRED]), sum(1 for _, cd in w.agents[CELL].values() if cd == 0),
                   sum(e for e, _ in w.agents[PREY].values()), sum(e for e, _ in w.agents[PRED].values())]
            assert row == g["trajectory"][samples].tolist(), step
            samples += 1
    assert samples == 2
