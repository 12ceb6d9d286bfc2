"""Secondary benchmarks for BASELINE configs 2, 3 and 5 (bench.py carries the headline config 4).
    python profiles/bench_configs.py gol|sir|pp [scale]
Prints one JSON line per config: per-apply device times (read+write phase / finish_write!), edges per second and the
finish_write! bandwidth against SURVEY.md §8(d)'s B_fin byte model."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import vahana_b200 as vh  # noqa: E402
from models import gol_sim, sir_sim, sir_step, pp_model, pp_step, PPCELL, ANIMAL  # noqa: E402

PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0


def backend():
    be = vh.default_backend()
    be.init(0)
    be.set_stream(torch.cuda.current_stream().cuda_stream)
    return be


def timed(fn, reps):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(reps):
        fn(i)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


def gol(scale):
    n = int(4096 * scale)
    init = np.random.default_rng(2).random((n, n)) < 0.35
    t0 = time.perf_counter()
    sim = gol_sim(backend(), init)
    build = time.perf_counter() - t0
    stats = []
    def step(i):
        sim.apply("gol_life", "Cell", ["Cell", "Neighbor"], "Cell")
        stats.append(sim.last_apply_stats())
    for i in range(3):
        step(i)
    stats.clear()
    dt = timed(step, 20)
    k = np.mean([s["ms_kernel"] for s in stats])
    cells = n * n
    print(json.dumps({"config": "gol", "cells": cells, "edges": 8 * cells, "build_s": build, "ms_per_apply_wall": dt * 1e3, "ms_kernel": k,
                      "cell_updates_per_s": cells / (k * 1e-3), "edges_per_s": 8 * cells / (k * 1e-3),
                      "csr_bytes_model_GBs": (5 * 8 * cells + 6 * cells) / (k * 1e-3) / 1e9, "frac_of_peak_csr_model": (5 * 8 * cells + 6 * cells) / (k * 1e-3) / 1e9 / PEAK,
                      "alive": sim.mapreduce("active", "+", "Cell", datatype="i8")}))


def sir(scale):
    npers, nloc = int(5e7 * scale), int(5e6 * scale)
    sim = sir_sim(backend(), npers, nloc, beta=0.3)
    per = {k: {"rw": [], "fin": [], "app": [], "read": []} for k in ("visit", "tally", "expose", "infect")}
    def step(i):
        for name, args in (("visit", ("sir_visit", "Person", ["Person"], ["Visit"])), ("tally", ("sir_tally", "Location", ["Visit"], ["Location"])),
                           ("expose", ("sir_expose", "Location", ["Location", "Visit"], ["Exposure"])),
                           ("infect", ("sir_infect", "Person", ["Person", "Exposure"], ["Person"]))):
            sim.apply(*args, seed=4 * i + len(per[name]["rw"]) % 4)
            st = sim.last_apply_stats()
            per[name]["rw"].append(st["ms_read_write"]); per[name]["fin"].append(st["ms_finish"])
            per[name]["app"].append(st["edges_appended"]); per[name]["read"].append(st["edges_read"])
    for i in range(2):
        step(i)
    for v in per.values():
        for l in v.values():
            l.clear()
    dt = timed(step, 5)
    out = {"config": "sir", "persons": npers, "locations": nloc, "ms_per_step_wall": dt * 1e3}
    for name, v in per.items():
        rw, fin, app = np.mean(v["rw"]), np.mean(v["fin"]), np.mean(v["app"])
        out[name] = {"ms_read_write": rw, "ms_finish": fin, "edges_appended": app, "edges_read": float(np.mean(v["read"]))}
        if app:
            s_e = 1 if name == "visit" else 4
            rows = nloc if name == "visit" else npers
            p = int(np.ceil(np.log2(rows) / 8))
            bfin = app * (4 + p * 2 * (8 + s_e) + 4) + 4 * rows
            out[name]["finish_GBs_model"] = bfin / (fin * 1e-3) / 1e9
            out[name]["finish_frac_of_peak"] = bfin / (fin * 1e-3) / 1e9 / PEAK
            out[name]["append_GBs_model"] = app * (8 + s_e) / (rw * 1e-3) / 1e9
    out["edges_appended_and_sorted_per_s"] = 2 * npers * 2 / (sum(np.mean(v["rw"]) + np.mean(v["fin"]) for v in per.values()) * 1e-3)
    print(json.dumps(out))


def pp(scale):
    d = int(2048 * scale)
    nprey, npred = int(838861 * scale * scale), int(209715 * scale * scale)
    rng = np.random.default_rng(3)
    sim = vh.create_simulation(pp_model(), backend=backend())
    n = d * d
    cells = np.zeros(n, dtype=np.dtype(PPCELL, align=True))
    ii, jj = np.meshgrid(np.arange(1, d + 1), np.arange(1, d + 1), indexing="ij")
    cells["pos"][:, 0] = ii.reshape(-1, order="F"); cells["pos"][:, 1] = jj.reshape(-1, order="F")
    cells["countdown"] = np.where(rng.random(n) < 0.5, 0, rng.integers(1, 6, n))
    cellids = sim.add_raster("raster", (d, d), "Cell", cells).reshape(-1, order="F")
    t0 = time.perf_counter()
    offs = [(0, 0), (0, -1), (-1, 0), (1, 0), (0, 1)]       # stencil(:manhatten, 2, 1) in product order with the centre first (move_to!)
    for species, count in (("Prey", nprey), ("Predator", npred)):
        st = np.zeros(count, dtype=np.dtype(ANIMAL, align=True))
        st["energy"] = rng.integers(1, 11, count)
        st["pos"][:, 0] = rng.integers(1, d + 1, count); st["pos"][:, 1] = rng.integers(1, d + 1, count)
        ids = sim.add_agents(species, st)
        x, y = st["pos"][:, 0] - 1, st["pos"][:, 1] - 1
        sim.add_edges(ids, cellids[x + y * d], f"Position{{{species}}}")
        fr, to = [], []
        for dx, dy in offs:     # per animal: (cell -> id, id -> cell) for the position, then for the 4 neighbours
            c = cellids[((x + dx) % d) + ((y + dy) % d) * d]
            fr += [c, ids]; to += [ids, c]
        fr = np.stack(fr, axis=1).reshape(-1); to = np.stack(to, axis=1).reshape(-1)
        sim.add_edges(fr, to, f"View{{{species}}}")
    sim.finish_init()
    build = time.perf_counter() - t0
    names = ["move_prey", "find_prey", "move_pred", "grow_food", "try_eat", "try_reproduce"]
    per = {k: {"rw": [], "fin": [], "app": []} for k in names}
    orig_apply = sim.apply
    counter = {"i": 0}
    def apply_rec(*a, **kw):
        orig_apply(*a, **kw)
        st = sim.last_apply_stats()
        k = names[counter["i"] % 6]; counter["i"] += 1
        per[k]["rw"].append(st["ms_read_write"]); per[k]["fin"].append(st["ms_finish"]); per[k]["app"].append(st["edges_appended"])
    sim.apply = apply_rec
    for i in range(2):
        pp_step(sim, i)
    for v in per.values():
        for l in v.values():
            l.clear()
    dt = timed(lambda i: pp_step(sim, 2 + i), 5)
    out = {"config": "pp", "raster": d, "prey": sim.mapreduce(None, "+", "Prey", init=0), "predators": sim.mapreduce(None, "+", "Predator", init=0),
           "build_s": build, "ms_per_step_wall": dt * 1e3}
    for k, v in per.items():
        out[k] = {"ms_read_write": float(np.mean(v["rw"])), "ms_finish": float(np.mean(v["fin"])), "edges_appended": float(np.mean(v["app"]))}
    print(json.dumps(out))


if __name__ == "__main__":
    which = sys.argv[1]
    scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    {"gol": gol, "sir": sir, "pp": pp}[which](scale)
