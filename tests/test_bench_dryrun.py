"""bench.py's engine arm cannot run without a GPU, and a typo in it would cost the round's bench line.  This test walks
`bench.run_engine` end to end on the CPU with the device pieces replaced by stand-ins — `torch.cuda` by a stub, the CUDA backend by the
oracle library (same C-ABI), the device workload generator by the host generator — and checks the JSON line it prints against the
bench contract (keys, types, internal consistency).  Nothing measured here means anything; only the plumbing is under test."""
import argparse
import ctypes as C
import io
import json
import sys
import time
import types
from contextlib import redirect_stdout

import numpy as np

import vahana_b200 as vh


class _Event:
    def __init__(self, enable_timing=True):
        self.t = 0.0

    def record(self):
        self.t = time.perf_counter()

    def elapsed_time(self, other):
        return (other.t - self.t) * 1e3


def test_bench_engine_arm_plumbing(oracle, monkeypatch):
    import torch
    import bench

    cuda = types.SimpleNamespace(
        is_available=lambda: True, set_device=lambda d: None, synchronize=lambda: None, Event=_Event,
        current_stream=lambda: types.SimpleNamespace(cuda_stream=0), device_count=lambda: 1)
    monkeypatch.setattr(torch, "cuda", cuda)
    monkeypatch.setattr(oracle, "init", lambda device=0: None, raising=False)
    monkeypatch.setattr(oracle, "set_stream", lambda s: None, raising=False)
    monkeypatch.setattr(oracle, "init_distributed", lambda rank=None, world=None: (0, 1), raising=False)
    monkeypatch.setattr(vh, "default_backend", lambda: oracle)
    lib = oracle.lib

    def build_sharded(h, atype, etype, n, sg, so, c, dmax, chunk, rank, world, ne_out):
        n = int(n.value)
        ne = C.c_uint64()
        lib.vbw_hk_powerlaw_host(C.c_uint64(n), 1, sg, so, c, dmax, None, None, None, C.byref(ne))
        fr = np.zeros(ne.value, dtype=np.uint64)
        to = np.zeros(ne.value, dtype=np.uint64)
        op = np.zeros(n, dtype=np.float64)
        lib.vbw_hk_powerlaw_host(C.c_uint64(n), 1, sg, so, c, dmax, fr.ctypes.data_as(C.c_void_p), to.ctypes.data_as(C.c_void_p),
                                 op.ctypes.data_as(C.c_void_p), C.byref(ne))
        ids = np.zeros(n, dtype=np.uint64)
        assert lib.vb_add_agents(h, C.c_int(1), op.ctypes.data_as(C.c_void_p), C.c_uint64(n), ids.ctypes.data_as(C.c_void_p)) == 0
        assert lib.vb_add_edges(h, C.c_int(0), fr.ctypes.data_as(C.c_void_p), to.ctypes.data_as(C.c_void_p), None, C.c_uint64(ne.value)) == 0
        ne_out._obj.value = ne.value
        return 0

    monkeypatch.setattr(lib, "vbw_hk_powerlaw_build_sharded", build_sharded, raising=False)
    real_stats = vh.Simulation.last_apply_stats

    def stats(self):                       # the oracle has no kernels: give the roofline arithmetic a non-zero duration
        st = real_stats(self)
        st["ms_kernel"] = st["ms_kernel"] or 1.0
        return st

    monkeypatch.setattr(vh.Simulation, "last_apply_stats", stats)
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        monkeypatch.delenv(k, raising=False)
    args = argparse.Namespace(gpus=1, steps=2, warmup=1, impl="engine", agents=3000.0, cpu_agents=2000, cpu_procs=1, no_cpu=False, no_secondary=False,
                              secondary_scale=1e-4)
    buf = io.StringIO()
    with redirect_stdout(buf):
        bench.run_engine(args)
    lines = [x for x in buf.getvalue().splitlines() if x.startswith("{")]
    assert len(lines) == 1, buf.getvalue()
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
                "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert key in d, key
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 3 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["unit"] == "edges/s" and d["dtype"] == "f64" and d["data"] == "synthetic" and d["scaling"] == "strong"
    assert d["config"]["workload"].startswith("hk-powerlaw") and d["config"]["agents"] == 3000 and d["config"]["edges"] > 3000
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12 and "traffic" in r
    assert abs(r["algorithmic_bytes_per_launch"] - (12.0 * d["config"]["edges"] + 20.0 * 3000)) < 1e-6
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == 1 and cb["value"] > 0 and "sample" in cb
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] >= 0 and e["d2h_bytes_per_step"] > 0      # (the oracle uploads no device view)
    assert e["with_state_download"]["d2h_bytes_per_step"] == 8 * 3000
    assert d["value"] > 0 and d["ms_per_step"] > 0 and isinstance(d["gpu_launches"], int)
    sec = d["config"]["secondary"]          # BASELINE configs 1, 2, 3, 5 and the docs' second eps, at toy sizes here
    assert set(sec) == {"hk_eps025", "gol_4096", "sir_50M_x_5M", "predator_prey_2048", "hk_100k"}
    for name, v in sec.items():
        assert "error" not in v, (name, v)
    assert sec["gol_4096"]["ms_per_generation"] > 0 and sec["gol_4096"]["algorithmic_bytes"] > 0
    assert sec["sir_50M_x_5M"]["visit"]["edges_appended"] > 0 and sec["sir_50M_x_5M"]["visit"]["finish_write_frac"] > 0
    assert sec["predator_prey_2048"]["applies_per_step"] == 6 and sec["hk_100k"]["edges_per_s"] > 0 and sec["hk_eps025"]["ms_per_step"] > 0
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
