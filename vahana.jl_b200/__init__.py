"""vahana_b200 — host-side mirror of Vahana.jl's API for the transition hot path.

The reference is Julia (absent from this image), so the host side above the C-ABI
(include/vahana_b200.h) is Python + ctypes, mirroring the reference's names, argument meaning
and error behaviour (AssertionError where the reference asserts):

    ModelTypes / register_agenttype! / register_edgetype! / register_param!   src/ModelTypes.jl:32-289
    create_model / create_simulation / finish_init! / apply! / mapreduce      src/Simulation.jl:115-856
    add_agent(s)! / agentstate / all_agents / num_agents                      src/Agent.jl, src/AgentMethods.jl
    add_edge(s)! / edges / neighborids / neighborstates / edgestates / num_edges / has_edge
                                                                              src/Edge.jl, src/EdgeMethods.jl
    add_raster! / connect_raster_neighbors! / calc_raster / rastervalues / move_to! / cellid
                                                                              src/Raster.jl
    get_global / set_global! / push_global! / modify_global!                  src/Global.jl (host-side state)

Everything that touches simulation state goes through the C-ABI of the CUDA engine
(vahana.jl_b200/csrc/build/libvahana_b200.so).  There is no CPU fallback: if the library is
missing, creating a simulation fails.  Tests may inject the CPU oracle (oracle/) as `backend=` to use
the same host code as a checker; the product never does.

Transition functions are CUDA device functors compiled into a model library and referred to by name
(`apply!(sim, "hk_step", ...)`), see include/vahana_device.cuh.
"""
from __future__ import annotations

import ctypes as C
import os
import time
from typing import Any, Iterable, Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_LIB = os.path.join(_HERE, "csrc", "build", "libvahana_b200.so")

EDGE_REF = 256
_AGENT_HINTS = {"Immortal": 1, "Independent": 2}
_EDGE_HINTS = {"Stateless": 1, "IgnoreFrom": 2, "SingleEdge": 4, "SingleType": 8, "IgnoreSourceState": 16}
_METRICS = {"chebyshev": 0, "euclidean": 1, "manhatten": 2}
_OPS = {"+": 0, "*": 1, "min": 2, "max": 3, "&": 4, "|": 5}
_DT = {np.dtype("i8"): 0, np.dtype("f8"): 1, np.dtype("?"): 2, np.dtype("i4"): 3, np.dtype("f4"): 4, np.dtype("u1"): 5}
_DT_NP = {v: k for k, v in _DT.items()}
ACC_EDGES, ACC_NEIGHBORIDS, ACC_EDGESTATES, ACC_NUM_EDGES, ACC_HAS_EDGE, ACC_NEIGHBORIDS_ITER, ACC_EDGESTATES_ITER = range(7)

# --- AgentID helpers (src/Agent.jl:30-118) ---------------------------------------------------------
BITS_TYPE, BITS_PROCESS, BITS_AGENTNR = 8, 20, 36
SHIFT_TYPE, SHIFT_RANK = BITS_PROCESS + BITS_AGENTNR, BITS_AGENTNR


def agent_id(typeid: int, rank: int, nr: int) -> int:
    return (typeid << SHIFT_TYPE) + (rank << SHIFT_RANK) + nr


def type_nr(aid: int) -> int:
    return int(aid) >> SHIFT_TYPE


def process_nr(aid: int) -> int:
    return (int(aid) >> SHIFT_RANK) & ((1 << BITS_PROCESS) - 1)


def agent_nr(aid: int) -> int:
    return int(aid) & ((1 << BITS_AGENTNR) - 1)


# --- config (src/Vahana.jl:42-84) -------------------------------------------------------------------
class _Config:
    detect_stateless = False
    asserts_enabled = True
    quiet = True


config = _Config()


def detect_stateless(flag: bool = True) -> None:
    config.detect_stateless = bool(flag)


def enable_asserts(flag: bool = True) -> None:
    config.asserts_enabled = bool(flag)


# --- C-ABI binding ------------------------------------------------------------------------------------
class _AgentTypeDesc(C.Structure):
    _fields_ = [("name", C.c_char_p), ("size", C.c_uint32), ("hints", C.c_uint32)]


class _EdgeTypeDesc(C.Structure):
    _fields_ = [("name", C.c_char_p), ("size", C.c_uint32), ("hints", C.c_uint32), ("target_type", C.c_int32),
                ("size_hint", C.c_uint64)]


class _ModelDesc(C.Structure):
    _fields_ = [("name", C.c_char_p), ("n_agent_types", C.c_uint32), ("agent_types", C.POINTER(_AgentTypeDesc)),
                ("n_edge_types", C.c_uint32), ("edge_types", C.POINTER(_EdgeTypeDesc)), ("param_size", C.c_uint32)]


# every symbol include/vahana_b200.h declares (tests check that the library exports all of them)
ABI_SYMBOLS = [
    "vb_last_error", "vb_backend", "vb_init", "vb_shutdown", "vb_comm_unique_id", "vb_comm_init", "vb_comm_rank",
    "vb_sim_create", "vb_sim_copy", "vb_sim_destroy", "vb_set_param", "vb_set_config", "vb_disable_transition_checks",
    "vb_add_agents", "vb_add_edges", "vb_remove_edges", "vb_add_raster", "vb_connect_raster_neighbors", "vb_move_to",
    "vb_cellid", "vb_finish_init", "vb_apply", "vb_has_transition", "vb_load_model_library", "vb_num_agents",
    "vb_all_agents", "vb_agentstate", "vb_num_edges_total", "vb_edges_of", "vb_all_edges", "vb_mapreduce", "vb_mapreduce_fn",
    "vb_rastervalues", "vb_calc_raster_num_edges", "vb_calc_rasterstate_fn", "vb_raster_info", "vb_num_transitions", "vb_export_csr",
    "vb_last_apply_stats", "vb_set_stream", "vb_last_kernel_ms", "vb_device_view_bytes", "vb_halo_bytes", "vb_set_uniform_offset", "vb_add_agent_per_process",
    "vb_last_apply_blocks", "vb_set_read_blocking", "vb_set_read_prefilter", "vb_last_apply_prefiltered", "vb_last_pass_rate", "vb_set_raster", "vb_last_halo_ms",
]


class Backend:
    """A loaded implementation of include/vahana_b200.h."""

    def __init__(self, path: str):
        if not os.path.exists(path):
            raise RuntimeError(
                f"vahana_b200: engine library not found at {path}. Build it with `python __graft_entry__.py build` "
                "(nvcc, sm_100a). There is no CPU fallback.")
        self.path = path
        self.lib = C.CDLL(path, mode=C.RTLD_LOCAL)
        self.lib.vb_last_error.restype = C.c_char_p
        self.lib.vb_backend.restype = C.c_char_p
        self.name = self.lib.vb_backend().decode()
        self._initialized = False

    def check(self, rc: int) -> None:
        if rc == 0:
            return
        msg = self.lib.vb_last_error().decode(errors="replace")
        if rc == 1:
            raise AssertionError(msg)
        if rc == 2:
            raise ValueError(msg)
        raise RuntimeError(f"vahana_b200 [{rc}]: {msg}")

    def init(self, device: int = 0) -> None:
        if not self._initialized:
            self.check(self.lib.vb_init(C.c_int(device)))
            self._initialized = True

    def load_model_library(self, path: str) -> None:
        """dlopen a model library (a separately compiled .cu that registers its transitions / maps with VB_REGISTER_TRANSITION /
        VB_REGISTER_MAP): the counterpart of `include("model.jl")` defining the closures."""
        self.check(self.lib.vb_load_model_library(os.path.abspath(path).encode()))

    def has_transition(self, name: str, agent_type: str) -> bool:
        return bool(self.lib.vb_has_transition(name.encode(), agent_type.encode()))

    def set_stream(self, cuda_stream: int) -> None:
        """Run all engine work on the given cudaStream_t (e.g. torch.cuda.current_stream().cuda_stream)."""
        self.check(self.lib.vb_set_stream(C.c_void_p(cuda_stream)))


    # -- multi-GPU: one process per GPU (src/MPIinit.jl) --
    def init_distributed(self, rank: Optional[int] = None, world: Optional[int] = None) -> tuple:
        """Creates the engine's NCCL communicator.  The 128-byte unique id is created on rank 0 and broadcast with
        torch.distributed (any backend), which must already be initialised when world > 1."""
        import torch
        import torch.distributed as dist
        if world is None:
            world = dist.get_world_size() if dist.is_initialized() else 1
            rank = dist.get_rank() if dist.is_initialized() else 0
        if world == 1:
            self.check(self.lib.vb_comm_init(0, 1, None))
            return 0, 1
        buf = (C.c_uint8 * 128)()
        if rank == 0:
            self.check(self.lib.vb_comm_unique_id(buf))
        dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
        t = torch.tensor(list(buf), dtype=torch.uint8, device=dev)
        dist.broadcast(t, src=0)
        ident = (C.c_uint8 * 128)(*t.cpu().tolist())
        self.check(self.lib.vb_comm_init(C.c_int(rank), C.c_int(world), ident))
        return rank, world


_default_backend: Optional[Backend] = None


def checked(f, g, itr, **kwargs):
    """checked(f, g, itr) (src/Helpers.jl:33-37): g(f, itr) unless itr is nothing, e.g. checked(fn, map, sim.neighborids(id, "Contact"))"""
    if itr is not None:
        return g(f, itr, **kwargs)
    return None


def rootonly(fn, *args, **kwargs):
    """@rootonly (src/Helpers.jl:236-242): run fn on rank 0 only"""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_rank() == 0:
        return fn(*args, **kwargs)
    return None


def equal_partition(n: int, nranks: int) -> list:
    """_create_equal_partition (src/Simulation.jl:353-367): contiguous blocks, sizes differ by at most one, larger blocks first.
    Returns the nranks + 1 block boundaries."""
    q, r = divmod(int(n), int(nranks))
    b = [0]
    for p in range(nranks):
        b.append(b[-1] + q + (1 if p < r else 0))
    return b


def plan_distribution(agents: dict, edges: dict, world: int, partition: Optional[dict] = None) -> tuple:
    """The host-side plan of distribute! (src/MPI.jl:11-84) as a pure function: which rank gets which agents under which new ids,
    and which edges (with both ids rewritten) follow their target.

    agents: {type id: (old ids in add order, states or None)}, edges: {edge name: (from ids, to ids, states or None)} - what rank 0
    holds after the initialisation phase.  partition: {old id: rank, 1-based as the reference's ProcessID}; None = the reference's
    :EqualAgentNumbers (contiguous equal blocks per type in id order, _create_equal_partition, src/Simulation.jl:353-367).
    Within a rank the agents keep their old order (the reference's order is the iteration order of a Dict, i.e. unpinned); the new
    id is (type, rank, position + 1) as add_agent! hands it out on the receiver (sendagents!, src/MPI.jl:112-149).  An edge goes to
    the new owner of its target, in the order the edges were added, so every target keeps its push! order (sendedges!, :289-351).
    Returns (shards, old, new, bounds): shards[r] = {"agents": {type id: (count, states)}, "edges": {name: (from, to, states)}},
    old/new = the idmapping as two aligned uint64 arrays sorted by old id, bounds = {type id: block boundaries} for equal blocks."""
    world = int(world)
    olds, news = [], []
    shards = [{"agents": {}, "edges": {}} for _ in range(world)]
    bounds = {}
    pkeys = pvals = None
    for tid in sorted(agents):
        ids, states = agents[tid]
        ids = np.asarray(ids, dtype=np.uint64).reshape(-1)
        n = ids.shape[0]
        if partition is None:
            b = equal_partition(n, world)
            bounds[tid] = b
            owner = np.searchsorted(np.array(b[1:]), np.arange(n), side="right")
        else:
            if pkeys is None:      # the partition as two aligned arrays, sorted by id: one vectorised lookup per type
                pkeys = np.fromiter((int(k) for k in partition.keys()), dtype=np.uint64, count=len(partition))
                pvals = np.fromiter((int(v) for v in partition.values()), dtype=np.int64, count=len(partition))
                o = np.argsort(pkeys, kind="stable")
                pkeys, pvals = pkeys[o], pvals[o]
            k = np.minimum(np.searchsorted(pkeys, ids), max(pkeys.shape[0] - 1, 0))
            found = (pkeys[k] == ids) if pkeys.shape[0] else np.zeros(n, dtype=bool)
            if n and not found.all():
                raise AssertionError(f"the partition does not name a rank for agent {int(ids[np.argmin(found)]):#x}")
            owner = (pvals[k] - 1) if n else np.zeros(0, dtype=np.int64)
            if n and (owner.min() < 0 or owner.max() >= world):
                raise AssertionError("the partition names a rank outside of 1..mpi.size")
        new = np.zeros(n, dtype=np.uint64)
        for r in range(world):
            sel = np.nonzero(owner == r)[0]                                      # ascending = old order
            new[sel] = (np.uint64(tid) << np.uint64(SHIFT_TYPE)) | (np.uint64(r) << np.uint64(SHIFT_RANK)) | (np.arange(1, len(sel) + 1, dtype=np.uint64))
            shards[r]["agents"][tid] = (len(sel), None if states is None else np.ascontiguousarray(np.asarray(states)[sel]))
        olds.append(ids)
        news.append(new)
    old = np.concatenate(olds) if olds else np.zeros(0, dtype=np.uint64)
    new = np.concatenate(news) if news else np.zeros(0, dtype=np.uint64)
    order = np.argsort(old, kind="stable")
    old, new = old[order], new[order]
    if old.shape[0] > 1 and (old[1:] == old[:-1]).any():
        raise AssertionError("an agent id was handed out twice")

    # ids of the initialisation phase are dense per type (type, rank, 1..n): a table indexed by the low bits answers a lookup with one
    # load where a binary search over 1e6 ids costs twenty cache misses (1e7 edges: 1.3 s instead of 12 s); sparse types are searched
    tables = {}
    tkey = (old >> np.uint64(SHIFT_TYPE)).astype(np.int64)
    for tid in np.unique(tkey).tolist():
        lo, hi = np.searchsorted(tkey, tid, side="left"), np.searchsorted(tkey, tid, side="right")
        span = int(old[hi - 1] - old[lo]) + 1
        if span <= 4 * (hi - lo) + 1024:
            tab = np.zeros(span, dtype=np.uint64)                   # 0 = no such agent (no AgentID is 0)
            tab[(old[lo:hi] - old[lo]).astype(np.int64)] = new[lo:hi]
            tables[tid] = (old[lo], tab)

    def remap(x):
        x = np.asarray(x, dtype=np.uint64).reshape(-1)
        if not x.shape[0]:
            return x
        out = np.zeros(x.shape[0], dtype=np.uint64)
        xt = (x >> np.uint64(SHIFT_TYPE)).astype(np.int64)
        searched = np.ones(x.shape[0], dtype=bool)
        for tid, (first, tab) in tables.items():
            sel = np.nonzero(xt == tid)[0]
            searched[sel] = False
            off = x[sel] - first                                    # wraps to a huge value below `first`
            ok = off < np.uint64(tab.shape[0])
            out[sel[ok]] = tab[off[ok].astype(np.int64)]
        if searched.any():
            q = x[searched]
            k = np.minimum(np.searchsorted(old, q), max(old.shape[0] - 1, 0))
            out[searched] = np.where(old[k] == q, new[k], np.uint64(0)) if old.shape[0] else np.uint64(0)
        if (out == 0).any():
            raise AssertionError("an edge names an agent that was never added")      # the reference: KeyError in idmapping[id]
        return out

    for name in edges:
        fr, to, states = edges[name]
        nfr, nto = remap(fr), remap(to)
        owner = ((nto >> np.uint64(SHIFT_RANK)) & np.uint64((1 << BITS_PROCESS) - 1)).astype(np.int64)
        for r in range(world):
            sel = np.nonzero(owner == r)[0]
            shards[r]["edges"][name] = (nfr[sel], nto[sel], None if states is None else np.ascontiguousarray(np.asarray(states)[sel]))
    return shards, old, new, bounds


def remove_process(aid: int) -> int:
    """remove_process (src/Agent.jl:81-82): the id with its rank bits cleared"""
    return int(aid) & ~(((1 << BITS_PROCESS) - 1) << SHIFT_RANK)


def raster_stencil(metric: str, ndims: int, distance: float):
    """_stencil_core (src/Raster.jl:82-96): the offsets within `distance` under `metric`, the first dimension running fastest, the
    centre left out"""
    d = int(np.floor(distance))
    if d < 0:
        return np.zeros((0, ndims), dtype=np.int64)
    axes = [np.arange(-d, d + 1, dtype=np.int64)] * ndims
    grid = np.stack(np.meshgrid(*axes, indexing="ij"), axis=-1).reshape(-1, ndims, order="F")      # first dimension fastest
    keep = (grid != 0).any(axis=1)
    if metric == "euclidean":
        keep &= np.sqrt((grid * grid).sum(axis=1).astype(np.float64)) <= distance
    elif metric == "manhatten":
        keep &= np.abs(grid).sum(axis=1).astype(np.float64) <= distance
    return grid[keep]


def raster_neighbor_edges(dims, cell_ids, distance=1, metric: str = "chebyshev", periodic: bool = True):
    """The edges connect_raster_neighbors! adds (src/Raster.jl:139-167), in its order: for every cell `org` in CartesianIndices order
    and every stencil offset o, an edge FROM the cell at org TO the cell at org + o (wrapped on a periodic raster, dropped when it
    leaves a clipped one).  Returns (from ids, to ids)."""
    dims = tuple(int(x) for x in dims)
    nd, n = len(dims), int(np.prod(dims))
    cell_ids = np.asarray(cell_ids, dtype=np.uint64).reshape(-1)
    st = raster_stencil(metric, nd, distance)
    pos = np.stack(np.unravel_index(np.arange(n), dims, order="F"), axis=1).astype(np.int64)          # 0-based positions, column-major order
    sh = pos[:, None, :] + st[None, :, :]                                                              # [cell, offset, dim]
    dd = np.array(dims, dtype=np.int64)
    inside = ((sh >= 0) & (sh < dd)).all(axis=2)
    keep = np.ones(inside.shape, dtype=bool) if periodic else inside
    sh = np.mod(sh, dd)
    strides = np.concatenate([[1], np.cumprod(dd[:-1])]).astype(np.int64)
    lin = (sh * strides).sum(axis=2)
    fr = np.broadcast_to(cell_ids[:, None], lin.shape)[keep]
    to = cell_ids[lin[keep]]
    return np.ascontiguousarray(fr), np.ascontiguousarray(to)


def raster_move_cells(dims, cell_ids, pos, distance=0, metric: str = "chebyshev", periodic: bool = True, only_surrounding: bool = False):
    """The cells move_to! connects an agent with (src/Raster.jl:437-477), in its order: the cell at `pos` (1-based; left out with
    only_surrounding), then, for distance >= 1, the cell at pos + o for every stencil offset o (wrapped on a periodic raster, skipped
    when it leaves a clipped one).  Returns their ids."""
    dd = np.array([int(x) for x in dims], dtype=np.int64)
    cell_ids = np.asarray(cell_ids, dtype=np.uint64).reshape(-1)
    p0 = np.array([int(x) - 1 for x in pos], dtype=np.int64)
    assert p0.shape[0] == dd.shape[0] and ((p0 >= 0) & (p0 < dd)).all(), "position outside the raster"
    strides = np.concatenate([[1], np.cumprod(dd[:-1])]).astype(np.int64)
    cells = [] if only_surrounding else [int((p0 * strides).sum())]
    if distance >= 1:
        sh = p0[None, :] + raster_stencil(metric, dd.shape[0], distance)
        keep = np.ones(sh.shape[0], dtype=bool) if periodic else ((sh >= 0) & (sh < dd)).all(axis=1)
        cells += [int(x) for x in (np.mod(sh, dd) * strides).sum(axis=1)[keep]]
    return cell_ids[np.array(cells, dtype=np.int64)]


def graph_growing_partition(agents: dict, edges: dict, world: int) -> dict:
    """A stand-in for the reference's default `partition_algo = :Metis` (src/Simulation.jl:420-446: Metis.partition on the graph of all
    agents and edges): greedy graph growing — the initial-partitioning step of Metis' own multilevel scheme — on the undirected agent
    graph.  Part p grows breadth-first from the unassigned vertex of smallest degree until it holds its share of the agents, so parts
    are connected where the graph allows it and equal in size up to one agent.  Deterministic.  Returns {old id: rank, 1-based} like the
    `partition` argument of finish_init!.  (Metis itself is not installed here; the partition only decides where agents live.)"""
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import breadth_first_order
    ids = np.concatenate([np.asarray(agents[t][0], dtype=np.uint64).reshape(-1) for t in sorted(agents)]) if agents else np.zeros(0, dtype=np.uint64)
    n = ids.shape[0]
    order = np.argsort(ids, kind="stable")
    sid = ids[order]

    def index(x):
        k = np.searchsorted(sid, np.asarray(x, dtype=np.uint64).reshape(-1))
        k = np.minimum(k, max(n - 1, 0))
        ok = sid[k] == np.asarray(x, dtype=np.uint64).reshape(-1) if n else np.zeros(0, dtype=bool)
        return order[k], ok
    rows, cols = [], []
    for name in edges:
        (u, oku), (v, okv) = index(edges[name][0]), index(edges[name][1])
        keep = oku & okv & (u != v)
        rows += [u[keep], v[keep]]
        cols += [v[keep], u[keep]]
    r = np.concatenate(rows) if rows else np.zeros(0, dtype=np.int64)
    c = np.concatenate(cols) if cols else np.zeros(0, dtype=np.int64)
    adj = coo_matrix((np.ones(r.shape[0], dtype=np.int8), (r, c)), shape=(n, n)).tocsr()
    deg = np.diff(adj.indptr)
    part = np.full(n, -1, dtype=np.int64)
    remaining = n
    for p in range(world):
        want = remaining // (world - p)
        got = 0
        while got < want:
            free = np.nonzero(part < 0)[0]
            seed = free[np.argmin(deg[free])]
            sub = adj[free][:, free]                                  # the graph of the unassigned vertices
            bfs = breadth_first_order(sub, int(np.searchsorted(free, seed)), directed=False, return_predecessors=False)
            take = free[bfs[: want - got]]
            part[take] = p
            got += take.shape[0]
        remaining -= want
    return {int(ids[k]): int(part[k]) + 1 for k in range(n)}


def updateids(idmapping, oldids):
    """updateids(idmap, oldids) (src/Simulation.jl:479-483, a helper of the reference's tests): the ids of the initialisation phase
    -> the ids after finish_init!.  The rank bits of the old id are ignored (every rank ran the same initialisation code with ids
    of its own rank; the mapping is keyed by rank 0's)."""
    scalar = np.ndim(oldids) == 0
    out = np.array([idmapping[remove_process(i)] for i in np.asarray(oldids, dtype=np.uint64).reshape(-1)], dtype=np.uint64)
    return int(out[0]) if scalar else out


def load_backend(path: Optional[str] = None) -> Backend:
    return Backend(path or os.environ.get("VAHANA_B200_LIB", DEFAULT_LIB))


def default_backend() -> Backend:
    """The CUDA engine.  Raises if it has not been built or is not the CUDA implementation."""
    global _default_backend
    if _default_backend is None:
        b = load_backend()
        if not b.name.startswith("cuda"):
            raise RuntimeError(f"vahana_b200: {b.path} is not the CUDA engine (backend={b.name})")
        _default_backend = b
    return _default_backend


# --- ModelTypes (src/ModelTypes.jl) ------------------------------------------------------------------
def _as_dtype(fields) -> Optional[np.dtype]:
    if fields is None:
        return None
    dt = np.dtype(fields, align=True) if not isinstance(fields, np.dtype) else fields
    if dt.names is None:
        raise AssertionError("agent/edge types must be structs (isbitstype)")
    return dt if dt.itemsize > 0 and len(dt.names) > 0 else None


class ModelTypes:
    def __init__(self):
        self.agent_names: list[str] = []
        self.agent_dtypes: dict[str, Optional[np.dtype]] = {}
        self.agent_hints: dict[str, set] = {}
        self.edge_names: list[str] = []
        self.edge_dtypes: dict[str, Optional[np.dtype]] = {}
        self.edge_hints: dict[str, set] = {}
        self.edge_kw: dict[str, dict] = {}
        self.params: list[tuple[str, Any]] = []
        self.globals: dict[str, Any] = {}

    # register_agenttype!: ModelTypes.jl:81-106
    def register_agenttype(self, name: str, fields=None, *hints: str) -> "ModelTypes":
        assert name not in self.agent_dtypes, f"Type {name} is already registered"
        assert len(self.agent_names) + 1 < 256, "Can not add new type, maximal number of types already registered"
        for h in hints:
            assert h in _AGENT_HINTS, f"The agent type hint {h} is unknown for type {name}"
        self.agent_names.append(name)
        self.agent_dtypes[name] = _as_dtype(fields)
        self.agent_hints[name] = set(hints)
        return self

    # register_edgetype!: ModelTypes.jl:158-230
    def register_edgetype(self, name: str, fields=None, *hints: str, target: Optional[str] = None, size: int = 0) -> "ModelTypes":
        assert name not in self.edge_dtypes, f"Type {name} is already registered"
        hs = set(hints)
        for h in hs:
            assert h in list(_EDGE_HINTS) + ["NumEdgesOnly", "HasEdgeOnly"], f"The edge type hint {h} is unknown for type {name}"
        if "NumEdgesOnly" in hs:
            hs |= {"Stateless", "IgnoreFrom"}
        if "HasEdgeOnly" in hs:
            hs |= {"Stateless", "IgnoreFrom", "SingleEdge"}
        hs -= {"NumEdgesOnly", "HasEdgeOnly"}
        if "SingleType" in hs:
            assert target is not None, f"For type {name} the :SingleType hint is set, but the target keyword is missing"
        if target is not None:
            hs.add("SingleType")
        dt = _as_dtype(fields)
        if dt is None and "Stateless" not in hs and config.detect_stateless:
            hs.add("Stateless")
        assert not ("SingleType" in hs and "SingleEdge" in hs and not ("Stateless" in hs and "IgnoreFrom" in hs)), \
            "The hints :SingleEdge and :SingleType can be only combined when the type has also the hints :Stateless and :IgnoreFrom"
        self.edge_names.append(name)
        self.edge_dtypes[name] = dt
        self.edge_hints[name] = hs
        self.edge_kw[name] = {"target": target, "size": int(size)}
        return self

    def register_param(self, name: str, default) -> "ModelTypes":
        self.params.append((name, default))
        return self

    def register_global(self, name: str, default) -> "ModelTypes":
        self.globals[name] = default
        return self


class Model:
    def __init__(self, types: ModelTypes, name: str):
        self.types, self.name = types, name
        self.immortal = [("Immortal" in types.agent_hints[n]) for n in types.agent_names]
        pf = []
        for pname, default in types.params:
            if isinstance(default, (bool, np.bool_)):
                pf.append((pname, "?"))
            elif isinstance(default, (int, np.integer)):
                pf.append((pname, "i8"))
            elif isinstance(default, (float, np.floating)):
                pf.append((pname, "f8"))
            elif isinstance(default, np.ndarray) or isinstance(default, np.void):
                pf.append((pname, default.dtype, default.shape) if isinstance(default, np.ndarray) else (pname, default.dtype))
            else:
                raise TypeError(f"parameter {pname}: unsupported type {type(default)}")
        self.param_dtype = np.dtype(pf, align=True) if pf else None


def create_model(types: ModelTypes, name: str) -> Model:
    return Model(types, name)


class VahanaLogger:
    """The reference's duration log (src/Logging.jl:30-73, file log/<name>_<rank>.log, src/Logging.jl:75-103): a message tagged
    `<Begin> x` is remembered, the matching `<End> x` writes
        Start (sec): <begin - logger start>
           x |#| Duration (ms): <duration>
            key = value            (the keyword arguments of the <Begin> message)
    so the reference's log tooling reads runs of this engine.  Host side only; nothing is written unless logging is switched on."""

    def __init__(self, filename: str, log_path: Optional[str] = None, rank: int = 0):
        d = log_path or "log"
        os.makedirs(d, exist_ok=True)
        self.path = os.path.join(d, f"{filename}_{rank}.log")
        self.stream = open(self.path, "w")
        self.starttime = time.time()
        self.begintimes = {}
        self.kwargs = {}

    def begin(self, what: str, **kwargs) -> None:
        self.begintimes[what] = time.time()
        self.kwargs[what] = kwargs

    def end(self, what: str) -> None:
        now = time.time()
        t0 = self.begintimes.get(what, now)
        self.stream.write(f"Start (sec): {t0 - self.starttime}\n   {what} |#| Duration (ms): {(now - t0) * 1000}\n")
        for k, v in self.kwargs.get(what, {}).items():
            self.stream.write(f"    {k} = {v}\n")
        self.stream.flush()

    def info(self, message: str) -> None:
        self.stream.write(f"Time (sec): {time.time() - self.starttime}\n  {message}\n")
        self.stream.flush()

    def close(self) -> None:
        if not self.stream.closed:
            self.stream.close()


# --- Simulation -----------------------------------------------------------------------------------------
class Simulation:
    def __init__(self, model: Model, params: Optional[dict] = None, globals_: Optional[dict] = None,
                 backend: Optional[Backend] = None, device: int = 0, _handle=None, logging: bool = False,
                 log_path: Optional[str] = None):
        self.model = model
        self.logger = VahanaLogger(model.name.replace(" ", "_"), log_path) if logging else None
        self.backend = backend or default_backend()
        self.backend.init(device)
        self.lib = self.backend.lib
        t = model.types
        self._aid = {n: i + 1 for i, n in enumerate(t.agent_names)}
        self._eid = {n: i for i, n in enumerate(t.edge_names)}
        self.globals = dict(t.globals)
        self.globals.update(globals_ or {})
        self.globals_last_change = 0
        self._params = None
        if model.param_dtype is not None:
            self._params = np.zeros((), dtype=model.param_dtype)
            for pname, default in t.params:
                self._params[pname] = default
            for k, v in (params or {}).items():
                self._params[k] = v
        # multi-rank initialisation phase: what is added is also kept on the host, so that finish_init(distribute=True) can
        # hand it out from rank 0 (distribute!, src/MPI.jl:11-84); released by finish_init
        self._stage = None
        self._unstageable = None
        if _handle is not None:
            self.h = _handle
            return
        self.h = self._create_handle()
        r, w = C.c_int(0), C.c_int(1)
        if hasattr(self.lib, "vb_comm_rank") and self.lib.vb_comm_rank(C.byref(r), C.byref(w)) == 0 and w.value > 1:
            self._stage = {"agents": {}, "edges": {}}

    def _create_handle(self):
        model, t = self.model, self.model.types
        at = (_AgentTypeDesc * max(1, len(t.agent_names)))()
        for i, n in enumerate(t.agent_names):
            dt = t.agent_dtypes[n]
            at[i] = _AgentTypeDesc(n.encode(), dt.itemsize if dt is not None else 0,
                                   sum(_AGENT_HINTS[h] for h in t.agent_hints[n]))
        et = (_EdgeTypeDesc * max(1, len(t.edge_names)))()
        for i, n in enumerate(t.edge_names):
            dt = t.edge_dtypes[n]
            tgt = t.edge_kw[n]["target"]
            et[i] = _EdgeTypeDesc(n.encode(), dt.itemsize if dt is not None else 0,
                                  sum(_EDGE_HINTS[h] for h in t.edge_hints[n]),
                                  self._aid[tgt] if tgt is not None else 0, t.edge_kw[n]["size"])
        self._keep = (at, et)
        md = _ModelDesc(model.name.encode(), len(t.agent_names), at, len(t.edge_names), et,
                        self._params.nbytes if self._params is not None else 0)
        h = C.c_void_p()
        pbuf = self._params.tobytes() if self._params is not None else None
        self.backend.check(self.lib.vb_sim_create(C.byref(md), pbuf, C.byref(h)))
        self.lib.vb_set_config(h, C.c_int(int(config.asserts_enabled)), C.c_int(1))
        return h

    # -- helpers --
    def _ck(self, rc):
        self.backend.check(rc)

    def type_id(self, name: str) -> int:
        return self._aid[name]

    def _ref(self, name: str) -> int:
        if name in self._aid:
            return self._aid[name]
        if name in self._eid:
            return EDGE_REF + self._eid[name]
        raise AssertionError(f"type {name} is not registered")

    def _refs(self, names) -> "C.Array":
        if isinstance(names, str):
            names = [names]
        names = list(names or [])
        arr = (C.c_int * max(1, len(names)))(*[self._ref(n) for n in names])
        return arr, len(names)

    def _adt(self, name):
        return self.model.types.agent_dtypes[name]

    def _edt(self, name):
        return self.model.types.edge_dtypes[name]

    def _state_bytes(self, dt: Optional[np.dtype], values, n: int) -> Optional[np.ndarray]:
        if dt is None:
            return None
        arr = np.asarray(values, dtype=dt) if not (isinstance(values, np.ndarray) and values.dtype == dt) else values
        arr = np.ascontiguousarray(arr).reshape(-1)
        assert arr.shape[0] == n
        return arr

    # -- lifecycle --
    def _log_begin(self, what: str, **kw) -> None:
        if self.logger is not None:
            self.logger.begin(what, **kw)

    def _log_end(self, what: str) -> None:
        if self.logger is not None:
            self.logger.end(what)

    def finish_simulation(self):
        self._log_begin("finish_simulation!")
        if getattr(self, "h", None) is not None:
            self.lib.vb_sim_destroy(self.h)
            self.h = None
        if getattr(self, "logger", None) is not None:
            self.logger.end("finish_simulation!")
            self.logger.close()
            self.logger = None

    def __del__(self):
        try:
            self.finish_simulation()
        except Exception:
            pass

    def copy_simulation(self) -> "Simulation":
        h = C.c_void_p()
        self._ck(self.lib.vb_sim_copy(self.h, C.byref(h)))
        s = Simulation(self.model, None, dict(self.globals), self.backend, _handle=h)
        s._params = None if self._params is None else self._params.copy()
        s._inited = getattr(self, "_inited", False)
        return s

    def param(self, name: str):
        return self._params[name].item() if self._params[name].shape == () else self._params[name]

    def set_param(self, name: str, value) -> "Simulation":
        self._params[name] = value
        self._ck(self.lib.vb_set_param(self.h, self._params.tobytes(), C.c_uint32(self._params.nbytes)))
        return self

    def set_uniform_offset(self, type_name: str, offset: int):
        """multi-GPU: global index of this rank's first agent of the type (keys ctx.uniform like a single-rank run)"""
        self._ck(self.lib.vb_set_uniform_offset(self.h, C.c_int(self._aid[type_name]), C.c_uint64(int(offset))))

    def disable_transition_checks(self, disable: bool):
        self._ck(self.lib.vb_disable_transition_checks(self.h, C.c_int(int(disable))))

    # -- globals (src/Global.jl:10-74, host-side) --
    def get_global(self, name):
        return self.globals[name]

    def set_global(self, name, value):
        self.globals[name] = value
        self.globals_last_change = self.num_transitions()

    def push_global(self, name, value):
        self.globals[name] = list(self.globals[name]) + [value]
        self.globals_last_change = self.num_transitions()

    def modify_global(self, name, f):
        self.set_global(name, f(self.globals[name]))

    # -- init phase --
    def add_agents(self, type_name: str, states=None, n: Optional[int] = None) -> np.ndarray:
        dt = self._adt(type_name)
        if dt is None:
            n = int(n if n is not None else (len(states) if states is not None else 1))
            buf = None
        else:
            arr = np.asarray(states, dtype=dt).reshape(-1)
            n = arr.shape[0]
            buf = np.ascontiguousarray(arr)
        ids = np.zeros(n, dtype=np.uint64)
        self._ck(self.lib.vb_add_agents(self.h, C.c_int(self._aid[type_name]), buf.ctypes.data_as(C.c_void_p) if buf is not None else None,
                                        C.c_uint64(n), ids.ctypes.data_as(C.c_void_p)))
        if self._stage is not None:
            self._stage["agents"].setdefault(self._aid[type_name], []).append((ids.copy(), None if buf is None else buf.copy()))
        return ids

    def add_agents_device(self, type_name: str, dev_ptr: int, n: int) -> None:
        """Bulk add from a device buffer of n AoS records (no ids returned: slots first..first+n-1 in order)."""
        self._ck(self.lib.vb_add_agents(self.h, C.c_int(self._aid[type_name]), C.c_void_p(dev_ptr), C.c_uint64(n), None))
        self._unstageable = "add_agents_device"

    def add_edges_device(self, edge_name: str, from_ptr: int, to_ptr: int, n: int, states_ptr: int = 0) -> None:
        """Bulk add from device buffers of AgentIDs (uint64) in call order."""
        self._ck(self.lib.vb_add_edges(self.h, C.c_int(self._eid[edge_name]), C.c_void_p(from_ptr), C.c_void_p(to_ptr),
                                       C.c_void_p(states_ptr) if states_ptr else None, C.c_uint64(n)))
        self._unstageable = "add_edges_device"

    def add_agent_per_process(self, type_name: str, state=None) -> int:
        """add_agent_per_process!(sim, agent) (src/Agent.jl:363-388): one agent on every rank, outside of transitions"""
        dt = self._adt(type_name)
        buf = None
        if dt is not None:
            buf = np.array([state if isinstance(state, (tuple, np.void)) else (state,)], dtype=dt)
        out = C.c_uint64()
        self._ck(self.lib.vb_add_agent_per_process(self.h, C.c_int(self._aid[type_name]), buf.ctypes.data_as(C.c_void_p) if buf is not None else None,
                                                   C.byref(out)))
        return out.value

    def add_agent(self, type_name: str, state=None) -> int:
        dt = self._adt(type_name)
        if dt is None:
            return int(self.add_agents(type_name, None, 1)[0])
        if not isinstance(state, (tuple, np.void)):
            state = (state,)
        return int(self.add_agents(type_name, np.array([state], dtype=dt))[0])

    def add_edges(self, from_ids, to_ids, edge_name: str, states=None):
        to = np.ascontiguousarray(np.asarray(to_ids, dtype=np.uint64).reshape(-1))
        n = to.shape[0]
        fr = np.ascontiguousarray(np.broadcast_to(np.asarray(from_ids, dtype=np.uint64), (n,)))
        dt = self._edt(edge_name)
        buf = None
        if dt is not None:
            if states is None:
                raise AssertionError(f"edge type {edge_name} has a state")
            arr = np.asarray(states, dtype=dt)
            buf = np.ascontiguousarray(np.broadcast_to(arr.reshape(-1) if arr.ndim else arr, (n,)))
        self._ck(self.lib.vb_add_edges(self.h, C.c_int(self._eid[edge_name]), fr.ctypes.data_as(C.c_void_p), to.ctypes.data_as(C.c_void_p),
                                       buf.ctypes.data_as(C.c_void_p) if buf is not None else None, C.c_uint64(n)))
        if self._stage is not None:
            self._stage["edges"].setdefault(edge_name, []).append((fr.copy(), to.copy(), None if buf is None else buf.copy()))

    def add_edge(self, from_id: int, to_id: int, edge_name: str, state=None):
        dt = self._edt(edge_name)
        if dt is not None and state is not None and not isinstance(state, (tuple, np.void)):
            state = (state,)
        self.add_edges([from_id], [to_id], edge_name, None if dt is None else np.array([state], dtype=dt))

    def remove_edges(self, *args):
        """remove_edges!(sim, to, T) or remove_edges!(sim, from, to, T)  (EdgeMethods.jl:527-599)"""
        if len(args) == 2:
            to, name = args
            fr = 0
        else:
            fr, to, name = args
        if self._stage is not None:
            self._unstageable = "remove_edges! in the initialisation phase"
        self._ck(self.lib.vb_remove_edges(self.h, C.c_int(self._eid[name]), C.c_uint64(int(fr)), C.c_uint64(int(to))))

    def add_raster(self, name: str, dims: Sequence[int], type_name: str, agent_constructor) -> np.ndarray:
        """add_raster!(sim, name, dims, agent_constructor): the constructor is called with the 1-based position
        tuple of every cell in CartesianIndices (column-major) order, or is an array of states in that order."""
        dims = tuple(int(d) for d in dims)
        if not hasattr(self, "_rasters"):
            self._rasters = {}
        self._rasters[name] = dims
        n = int(np.prod(dims))
        dt = self._adt(type_name)
        if callable(agent_constructor):
            states = np.zeros(n, dtype=dt) if dt is not None else None
            for i, pos in enumerate(np.ndindex(*dims[::-1])):
                v = agent_constructor(tuple(p + 1 for p in pos[::-1]))
                if dt is not None:
                    states[i] = v if isinstance(v, (tuple, np.void)) else (v,)
        else:
            states = None if dt is None else np.ascontiguousarray(np.asarray(agent_constructor, dtype=dt).reshape(-1))
        ids = np.zeros(n, dtype=np.uint64)
        d = (C.c_int64 * len(dims))(*dims)
        self._ck(self.lib.vb_add_raster(self.h, name.encode(), C.c_int(len(dims)), d, C.c_int(self._aid[type_name]),
                                        states.ctypes.data_as(C.c_void_p) if states is not None else None,
                                        ids.ctypes.data_as(C.c_void_p)))
        if self._stage is not None:     # the cells are agents like any other; the id grid is handed out by finish_init (broadcastids)
            self._stage["agents"].setdefault(self._aid[type_name], []).append((ids.copy(), None if states is None else states.copy()))
            self._stage.setdefault("rasters", {})[name] = (dims, self._aid[type_name], ids.copy())
        return ids.reshape(dims, order="F")

    def connect_raster_neighbors(self, name: str, edge_name: str, edge_state=None, distance=1, metric: str = "chebyshev",
                                 periodic: bool = True):
        dt = self._edt(edge_name)
        buf = None
        if dt is not None:
            buf = np.array([edge_state if isinstance(edge_state, (tuple, np.void)) else (edge_state,)], dtype=dt)
        self._ck(self.lib.vb_connect_raster_neighbors(self.h, name.encode(), C.c_int(self._eid[edge_name]), C.c_double(float(distance)),
                                                      C.c_int(_METRICS[metric]), C.c_int(int(periodic)),
                                                      buf.ctypes.data_as(C.c_void_p) if buf is not None else None))
        if self._stage is not None:
            if name in self._stage.get("rasters", {}):
                # stage the edges the engine just added (it keeps a raster stencil implicit) in the reference's add order:
                # cells in CartesianIndices order, per cell the stencil offsets in _stencil_core order (src/Raster.jl:82-110,139-167)
                dims, _tid, cells = self._stage["rasters"][name]
                fr, to = raster_neighbor_edges(dims, cells, distance, metric, periodic)
                st = None if buf is None else np.broadcast_to(buf, (fr.shape[0],)).copy()
                self._stage["edges"].setdefault(edge_name, []).append((fr, to, st))
            else:
                self._unstageable = "connect_raster_neighbors on a raster that was not added through add_raster"

    def move_to(self, name: str, aid: int, pos, edge_from_raster: Optional[str], edge_to_raster: Optional[str],
                state_from=None, state_to=None, distance=0, metric: str = "chebyshev", periodic: bool = True,
                only_surrounding: bool = False):
        p = (C.c_int64 * len(pos))(*[int(x) for x in pos])

        def sb(ename, st):
            if ename is None or self._edt(ename) is None:
                return None
            return np.array([st if isinstance(st, (tuple, np.void)) else (st,)], dtype=self._edt(ename))
        bf, bt = sb(edge_from_raster, state_from), sb(edge_to_raster, state_to)
        if self._stage is not None:
            if name in self._stage.get("rasters", {}):     # the edges the engine adds below, staged in move_to!'s order (from-raster edge, then to-raster edge, per cell)
                dims, _tid, grid = self._stage["rasters"][name]
                cells = raster_move_cells(dims, grid, pos, distance, metric, periodic, only_surrounding)
                me = np.full(cells.shape[0], int(aid), dtype=np.uint64)
                if edge_from_raster is not None and edge_from_raster == edge_to_raster:
                    fr, to = np.stack([cells, me], axis=1).reshape(-1), np.stack([me, cells], axis=1).reshape(-1)
                    st = None if bf is None else np.stack([np.broadcast_to(bf, cells.shape), np.broadcast_to(bt, cells.shape)], axis=1).reshape(-1)
                    self._stage["edges"].setdefault(edge_from_raster, []).append((fr, to, st))
                else:
                    if edge_from_raster is not None:
                        self._stage["edges"].setdefault(edge_from_raster, []).append((cells, me, None if bf is None else np.broadcast_to(bf, cells.shape).copy()))
                    if edge_to_raster is not None:
                        self._stage["edges"].setdefault(edge_to_raster, []).append((me, cells, None if bt is None else np.broadcast_to(bt, cells.shape).copy()))
            else:
                self._unstageable = "move_to on a raster that was not added through add_raster"
        self._ck(self.lib.vb_move_to(self.h, name.encode(), C.c_uint64(int(aid)), p,
                                     C.c_int(self._eid[edge_from_raster] if edge_from_raster else -1),
                                     bf.ctypes.data_as(C.c_void_p) if bf is not None else None,
                                     C.c_int(self._eid[edge_to_raster] if edge_to_raster else -1),
                                     bt.ctypes.data_as(C.c_void_p) if bt is not None else None,
                                     C.c_double(float(distance)), C.c_int(_METRICS[metric]), C.c_int(int(periodic)),
                                     C.c_int(int(only_surrounding))))

    def cellid(self, name: str, pos) -> int:
        p = (C.c_int64 * len(pos))(*[int(x) for x in pos])
        out = C.c_uint64()
        self._ck(self.lib.vb_cellid(self.h, name.encode(), p, C.byref(out)))
        return out.value

    def random_pos(self, name: str, weights=None, rng=None) -> tuple:
        """random_pos([rng], sim, raster[, weights]) (src/Raster.jl:512-543): a 1-based position of the raster, uniform or with
        probability proportional to `weights` (same shape as the raster; Julia's vec() order = column-major).  Host side, init phase."""
        dims = self.raster_info(name)
        rng = rng or np.random.default_rng()
        n = int(np.prod(dims))
        if weights is None:
            lin = int(rng.integers(n))
        else:
            w = np.asarray(weights, dtype=np.float64)
            assert tuple(w.shape) == tuple(dims), f"`weights` must have the same dimension as the raster :{name}"
            p = w.reshape(-1, order="F")
            lin = int(rng.choice(n, p=p / p.sum()))
        return tuple(int(i) + 1 for i in np.unravel_index(lin, dims, order="F"))

    def random_cell(self, name: str, weights=None, rng=None) -> int:
        """random_cell([rng], sim, raster[, weights]) (src/Raster.jl:553-577): the id of a random cell."""
        return self.cellid(name, self.random_pos(name, weights, rng))

    def finish_init(self, partition: Optional[dict] = None, return_idmapping: bool = False, partition_algo: str = "EqualAgentNumbers",
                    distribute: bool = True):
        """finish_init!(sim; partition, return_idmapping, partition_algo, distribute) (src/Simulation.jl:403-476).

        One rank: nothing to distribute; `return_idmapping` gives the identity mapping.  Several ranks and `distribute=True` (the
        reference's default): everything the initialisation phase added on rank 0 is handed out - agents by `partition` ({old id:
        rank, 1-based}), in contiguous equal blocks per type (:EqualAgentNumbers) or by graph growing (partition_algo="Metis": Metis itself
        is not installed, graph_growing_partition stands in), every edge to the new owner of its target - and what the other ranks added is discarded, as in the reference.
        `distribute=False` keeps what every rank added itself (SPMD initialisation: each rank adds its own block with ids of its
        own rank; the bench-sized graphs are generated that way, on the device).  Returns the idmapping {old id: new id} when
        `return_idmapping`, else the simulation."""
        self._log_begin("finish_init!")
        rank, world = C.c_int(0), C.c_int(1)
        if hasattr(self.lib, "vb_comm_rank"):
            self.lib.vb_comm_rank(C.byref(rank), C.byref(world))
        idmapping = None
        if distribute and world.value > 1:
            idmapping = self._distribute(partition, partition_algo, rank.value, world.value, return_idmapping)
        self._stage = None
        self._ck(self.lib.vb_finish_init(self.h))
        self._inited = True
        if distribute and return_idmapping and world.value == 1:
            idmapping = {}
            for name in self.model.types.agent_names:
                for i in self.all_agentids(name, all_ranks=False):
                    idmapping[int(i)] = int(i)
        self._log_end("finish_init!")
        return idmapping if (distribute and return_idmapping) else self

    def _distribute(self, partition, partition_algo: str, rank: int, world: int, want_mapping: bool):
        """distribute! (src/MPI.jl:11-84) on the host: rank 0 plans (plan_distribution), the shards travel with
        torch.distributed (initialisation time, host data), every rank rebuilds its engine simulation from its shard."""
        import torch.distributed as dist
        if partition_algo not in ("EqualAgentNumbers", "Metis"):
            raise ValueError("the partition_algo given is unknown")
        use_growing = partition_algo == "Metis" and not partition      # Metis itself is not installed: greedy graph growing stands in
        if self._unstageable is not None or self._stage is None:
            raise AssertionError(f"finish_init(distribute=True) on several ranks needs host-side adds ({self._unstageable or 'no staged init phase'}); "
                                 "use distribute=False for an SPMD initialisation")
        if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() == world):
            raise AssertionError("finish_init(distribute=True) needs torch.distributed initialised with one process per rank")
        shards, meta = None, [None]
        if rank == 0:
            agents = {}
            for tid, chunks in self._stage["agents"].items():
                ids = np.concatenate([c[0] for c in chunks])
                agents[tid] = (ids, None if chunks[0][1] is None else np.concatenate([c[1] for c in chunks]))
            edges = {}
            for name, chunks in self._stage["edges"].items():
                edges[name] = (np.concatenate([c[0] for c in chunks]), np.concatenate([c[1] for c in chunks]),
                               None if chunks[0][2] is None else np.concatenate([c[2] for c in chunks]))
            if use_growing:
                partition = graph_growing_partition(agents, edges, world)
            shards, old, new, bounds = plan_distribution(agents, edges, world, partition or None)
            rasters = {}
            for rname, (dims, tid, ids) in self._stage.get("rasters", {}).items():      # broadcastids (src/MPI.jl:59-73): the grids with the new ids
                rasters[rname] = (dims, tid, new[np.searchsorted(old, ids)])
            meta = [(old, new, bounds, rasters)]
        mine = [None]
        dist.scatter_object_list(mine, shards, src=0)
        dist.broadcast_object_list(meta, src=0)
        old, new, bounds, rasters = meta[0]
        shard = mine[0]
        # rebuild: the initialisation phase of this rank is replaced by its shard
        self.lib.vb_sim_destroy(self.h)
        self.h = self._create_handle()
        self._stage = None
        names = {v: k for k, v in self._aid.items()}
        for tid in sorted(shard["agents"]):
            n, states = shard["agents"][tid]
            if n:
                ids = self.add_agents(names[tid], states, n)
                assert int(ids[0]) == agent_id(tid, rank, 1) and int(ids[-1]) == agent_id(tid, rank, n), "receiver ids differ from the plan"
            if tid in bounds:   # equal blocks: a sharded stochastic run draws what the single-rank run draws
                self.set_uniform_offset(names[tid], bounds[tid][rank])
        for name, (fr, to, states) in shard["edges"].items():
            if to.shape[0]:
                self.add_edges(fr, to, name, states)
        for rname, (dims, tid, ids) in rasters.items():        # every rank knows the whole id grid; its read-outs join the ranks
            d = (C.c_int64 * len(dims))(*dims)
            ids = np.ascontiguousarray(ids, dtype=np.uint64)
            self._ck(self.lib.vb_set_raster(self.h, rname.encode(), C.c_int(len(dims)), d, C.c_int(tid), ids.ctypes.data_as(C.c_void_p)))
            if not hasattr(self, "_rasters"):
                self._rasters = {}
            self._rasters[rname] = tuple(dims)
        return dict(zip(old.tolist(), new.tolist())) if want_mapping else None

    # -- apply! (src/Simulation.jl:720-821) --
    def apply(self, transition: str, call, read, write, add_existing=(), with_edge: Optional[str] = None, seed: int = 0):
        c, nc = self._refs(call)
        r, nr = self._refs(read)
        w, nw = self._refs(write)
        a, na = self._refs(add_existing)
        we = self._eid[with_edge] if with_edge is not None else -1
        self._log_begin("apply!", func=transition, transition=(self.num_transitions() + 1) if self.logger is not None else 0)
        self._ck(self.lib.vb_apply(self.h, transition.encode(), c, nc, r, nr, w, nw, a, na, C.c_int(we), C.c_uint64(seed)))
        self._log_end("apply!")
        return self

    def apply_copy(self, transition: str, call, read, write, **kwargs) -> "Simulation":
        """apply(sim, f, call, read, write; kwargs...) (src/Simulation.jl:847-856): the non-mutating form — copy_simulation, then
        apply! on the copy; returns the copy and leaves `self` untouched (meant for the REPL, not for production loops)."""
        new = self.copy_simulation()
        new.apply(transition, call, read, write, **kwargs)
        return new

    def num_transitions(self) -> int:
        n = C.c_int64()
        self._ck(self.lib.vb_num_transitions(self.h, C.byref(n)))
        return n.value

    def set_read_blocking(self, block_mb: float = -1.0, min_mb: float = -1.0, eager: int = -1) -> None:
        """Policy of the source-blocked read phase (engine-specific tuning knob, no reference counterpart)."""
        self._ck(self.lib.vb_set_read_blocking(self.h, C.c_double(block_mb), C.c_double(min_mb), C.c_int(eager)))

    def set_read_prefilter(self, on: int = -1) -> None:
        """Prefiltered sweeps of reduce transitions with a key (engine-specific knob; results do not depend on it)."""
        self._ck(self.lib.vb_set_read_prefilter(self.h, C.c_int(on)))

    def last_apply_stats(self) -> dict:
        a, b = C.c_double(), C.c_double()
        er, ea, ac, kl = C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_uint64()
        self._ck(self.lib.vb_last_apply_stats(self.h, C.byref(a), C.byref(b), C.byref(er), C.byref(ea), C.byref(ac), C.byref(kl)))
        k = C.c_double()
        self._ck(self.lib.vb_last_kernel_ms(self.h, C.byref(k)))
        nb = C.c_uint32()
        self._ck(self.lib.vb_last_apply_blocks(self.h, C.byref(nb)))
        pf = C.c_int()
        self._ck(self.lib.vb_last_apply_prefiltered(self.h, C.byref(pf)))
        pr = C.c_double(-1.0)
        self._ck(self.lib.vb_last_pass_rate(self.h, C.byref(pr)))
        return {"ms_read_write": a.value, "ms_finish": b.value, "ms_kernel": k.value, "edges_read": er.value,
                "edges_appended": ea.value, "agents_called": ac.value, "kernel_launches": kl.value, "source_blocks": nb.value,
                "prefiltered": bool(pf.value), "pass_rate": pr.value if pr.value >= 0 else None}

    # -- agent queries --
    def num_agents(self, type_name: str) -> int:
        n = C.c_uint64()
        self._ck(self.lib.vb_num_agents(self.h, C.c_int(self._aid[type_name]), C.byref(n)))
        return n.value

    def _all(self, type_name: str):
        n = C.c_uint64()
        t = C.c_int(self._aid[type_name])
        self._ck(self.lib.vb_all_agents(self.h, t, None, None, C.c_uint64(0), C.byref(n)))
        dt = self._adt(type_name)
        states = np.zeros(n.value, dtype=dt) if dt is not None else None
        ids = np.zeros(n.value, dtype=np.uint64)
        self._ck(self.lib.vb_all_agents(self.h, t, states.ctypes.data_as(C.c_void_p) if states is not None else None,
                                        ids.ctypes.data_as(C.c_void_p), C.c_uint64(n.value), C.byref(n)))
        return states, ids

    def _world(self) -> int:
        w, r = C.c_int(1), C.c_int(0)
        if hasattr(self.lib, "vb_comm_rank"):
            self.lib.vb_comm_rank(C.byref(r), C.byref(w))
        return w.value

    def _rank(self) -> int:
        w, r = C.c_int(1), C.c_int(0)
        if hasattr(self.lib, "vb_comm_rank"):
            self.lib.vb_comm_rank(C.byref(r), C.byref(w))
        return r.value

    def _join(self, arr: Optional[np.ndarray]):
        """join (src/MPI.jl:481-517): the ranks' vectors concatenated in rank order, on every rank (collective; host data)"""
        if arr is None or self._world() == 1:
            return arr
        import torch.distributed as dist
        assert dist.is_available() and dist.is_initialized(), "all_ranks=True on several ranks needs torch.distributed (or pass all_ranks=False)"
        parts = [None] * dist.get_world_size()
        dist.all_gather_object(parts, arr)
        return np.concatenate(parts)

    def all_agents(self, type_name: str, all_ranks: bool = True) -> np.ndarray:
        """all_agents(sim, T, [all_ranks = true]) (src/Agent.jl:234-262): the states of the live agents in ascending nr; on several
        ranks joined in rank order (collective) unless all_ranks is false"""
        assert self._adt(type_name) is not None, "all_agents can be only called for agent types that have fields"
        st = self._all(type_name)[0]
        return self._join(st) if all_ranks else st

    def all_agentids(self, type_name: str, all_ranks: bool = True) -> np.ndarray:
        """all_agentids(sim, T, [all_ranks = true]) (src/Agent.jl:287-313), same order as all_agents"""
        ids = self._all(type_name)[1]
        return self._join(ids) if all_ranks else ids

    def agentstate(self, aid: int, type_name: str):
        dt = self._adt(type_name)
        out = np.zeros(1, dtype=dt if dt is not None else np.dtype("u1"))
        self._ck(self.lib.vb_agentstate(self.h, C.c_uint64(int(aid)), C.c_int(self._aid[type_name]), out.ctypes.data_as(C.c_void_p)))
        return out[0] if dt is not None else ()

    def agentstate_flexible(self, aid: int):
        return self.agentstate(aid, self.model.types.agent_names[type_nr(aid) - 1])

    # -- edge queries (EdgeMethods.jl:704-892) --
    def _row(self, to: int, edge_name: str, what: int):
        e = C.c_int(self._eid[edge_name])
        n = C.c_int64()
        self._ck(self.lib.vb_edges_of(self.h, e, C.c_uint64(int(to)), C.c_int(what), None, None, C.c_uint64(0), C.byref(n)))
        if n.value < 0:
            return None, None, -1
        hints = self.model.types.edge_hints[edge_name]
        if what in (ACC_NUM_EDGES, ACC_HAS_EDGE) or ("Stateless" in hints and "IgnoreFrom" in hints):
            return None, None, n.value
        dt = self._edt(edge_name)
        fr = np.zeros(n.value, dtype=np.uint64)
        st = np.zeros(n.value, dtype=dt) if dt is not None else None
        self._ck(self.lib.vb_edges_of(self.h, e, C.c_uint64(int(to)), C.c_int(what), fr.ctypes.data_as(C.c_void_p),
                                      st.ctypes.data_as(C.c_void_p) if st is not None else None, C.c_uint64(n.value), C.byref(n)))
        return fr, st, n.value

    def _single(self, edge_name):
        return "SingleEdge" in self.model.types.edge_hints[edge_name]

    def edges(self, to: int, edge_name: str):
        """Vector of (from, state) pairs; a single pair for :SingleEdge; None for `nothing`."""
        fr, st, n = self._row(to, edge_name, ACC_EDGES)
        if n < 0:
            return None
        es = [(int(fr[i]), st[i]) for i in range(n)]
        return es[0] if self._single(edge_name) else es

    def neighborids(self, to: int, edge_name: str, _what=ACC_NEIGHBORIDS):
        fr, _, n = self._row(to, edge_name, _what)
        if n < 0:
            return None
        return int(fr[0]) if self._single(edge_name) else [int(x) for x in fr]

    def neighborids_iter(self, to: int, edge_name: str):
        return self.neighborids(to, edge_name, ACC_NEIGHBORIDS_ITER)

    def edgestates(self, to: int, edge_name: str, _what=ACC_EDGESTATES):
        _, st, n = self._row(to, edge_name, _what)
        if n < 0:
            return None
        return st[0] if self._single(edge_name) else st

    def edgestates_iter(self, to: int, edge_name: str):
        return self.edgestates(to, edge_name, ACC_EDGESTATES_ITER)

    def neighborstates(self, to: int, edge_name: str, agent_type: str):
        ids = self.neighborids(to, edge_name)
        if ids is None:
            return None
        if self._single(edge_name):
            return self.agentstate(ids, agent_type)
        return [self.agentstate(i, agent_type) for i in ids]

    def neighborstates_iter(self, to: int, edge_name: str, agent_type: str):
        """neighborstates_iter (src/EdgeMethods.jl:783-803): the lazy form; not defined with :SingleEdge or :IgnoreFrom"""
        ids = self.neighborids_iter(to, edge_name)
        return None if ids is None else (self.agentstate(i, agent_type) for i in ids)

    def neighborstates_flexible_iter(self, to: int, edge_name: str):
        ids = self.neighborids_iter(to, edge_name)
        return None if ids is None else (self.agentstate_flexible(i) for i in ids)

    def neighborstates_flexible(self, to: int, edge_name: str):
        ids = self.neighborids(to, edge_name)
        if ids is None:
            return None
        if self._single(edge_name):
            return self.agentstate_flexible(ids)
        return [self.agentstate_flexible(i) for i in ids]

    def num_edges(self, *args, write: bool = False) -> int:
        """num_edges(sim, id, T) (EdgeMethods.jl:850-869) or num_edges(sim, T; write) (Edge.jl:373-389)."""
        if len(args) == 1:
            n = C.c_uint64()
            self._ck(self.lib.vb_num_edges_total(self.h, C.c_int(self._eid[args[0]]), C.c_int(int(write)), C.byref(n)))
            return n.value
        to, name = args
        return self._row(to, name, ACC_NUM_EDGES)[2]

    def has_edge(self, to: int, edge_name: str) -> bool:
        return self._row(to, edge_name, ACC_HAS_EDGE)[2] >= 1

    def all_edges(self, edge_name: str, all_ranks: bool = True):
        """all_edges(sim, T, [all_ranks = true]) (src/EdgeMethods.jl:1005-1027) as three aligned arrays (to, from, states)"""
        to, fr, st = self._all_edges_local(edge_name)
        if all_ranks and self._world() > 1:
            to, fr, st = self._join(to), self._join(fr), self._join(st)
        return to, fr, st

    def _all_edges_local(self, edge_name: str):
        n = C.c_uint64()
        e = C.c_int(self._eid[edge_name])
        self._ck(self.lib.vb_all_edges(self.h, e, None, None, None, C.c_uint64(0), C.byref(n)))
        dt = self._edt(edge_name)
        to = np.zeros(n.value, dtype=np.uint64)
        fr = np.zeros(n.value, dtype=np.uint64)
        st = np.zeros(n.value, dtype=dt) if dt is not None else None
        self._ck(self.lib.vb_all_edges(self.h, e, to.ctypes.data_as(C.c_void_p), fr.ctypes.data_as(C.c_void_p),
                                       st.ctypes.data_as(C.c_void_p) if st is not None else None, C.c_uint64(n.value), C.byref(n)))
        return to, fr, st

    def export_csr(self, edge_name: str, target_type: str, nrows: int):
        off = np.zeros(nrows + 1, dtype=np.uint64)
        e, t = C.c_int(self._eid[edge_name]), C.c_int(self._aid[target_type])
        self._ck(self.lib.vb_export_csr(self.h, e, t, off.ctypes.data_as(C.c_void_p), C.c_uint64(nrows), None, None, C.c_uint64(0)))
        n = int(off[-1])
        dt = self._edt(edge_name)
        fr = np.zeros(n, dtype=np.uint64)
        st = np.zeros(n, dtype=dt) if dt is not None else None
        self._ck(self.lib.vb_export_csr(self.h, e, t, off.ctypes.data_as(C.c_void_p), C.c_uint64(nrows), fr.ctypes.data_as(C.c_void_p),
                                        st.ctypes.data_as(C.c_void_p) if st is not None else None, C.c_uint64(n)))
        return off, fr, st

    # -- mapreduce (AgentMethods.jl:533-565, EdgeMethods.jl:972-994) --
    def mapreduce(self, field: Optional[str], op: str, type_name: str, datatype=None, init=None, equals=None):
        """mapreduce(sim, a -> a.field, op, T; datatype, init).  field=None maps every element to 1 (`_ -> 1`);
        equals=v maps to (a.field == v) (`c -> c.countdown == 0`)."""
        ref = self._ref(type_name)
        dt = self._adt(type_name) if ref < EDGE_REF else self._edt(type_name)
        if ref >= EDGE_REF:
            assert "Stateless" not in self.model.types.edge_hints[type_name], \
                f"mapreduce is not defined for the hint combination of {type_name}"
        if field is None:
            off, fdt = 0, -1
            src = np.dtype("i8")
        else:
            src = dt.fields[field][0]
            off, fdt = dt.fields[field][1], _DT[src]
        if datatype is None:   # val4empty: Helpers.jl:44-50
            if equals is not None:
                datatype = np.dtype("i8")
            elif op in ("&", "|") and src == np.dtype("?"):
                datatype = np.dtype("?")
            else:
                datatype = np.dtype("f8") if src.kind == "f" else np.dtype("i8")
        rdt = np.dtype(datatype)
        out = np.zeros(1, dtype=rdt)
        ib = np.array([init], dtype=rdt) if init is not None else None
        self._ck(self.lib.vb_mapreduce(self.h, C.c_int(ref), C.c_int(off), C.c_int(fdt), C.c_int(int(equals is not None)),
                                       C.c_int64(int(equals) if equals is not None else 0), C.c_int(_OPS[op]), C.c_int(_DT[rdt]),
                                       ib.ctypes.data_as(C.c_void_p) if ib is not None else None, out.ctypes.data_as(C.c_void_p)))
        return out[0].item()

    def mapreduce_fn(self, map_name: str, op: str, type_name: str, datatype=None, init=None):
        """mapreduce(sim, f, op, T; datatype, init) with f a registered map functor (VB_REGISTER_MAP), e.g.
        mapreduce(sim, b -> b.x - b.y, +, Bought) (docs/examples/tutorial1.jl:548) = sim.mapreduce_fn("market_x_minus_y", "+", "Bought")."""
        ref = self._ref(type_name)
        if ref >= EDGE_REF:
            assert "Stateless" not in self.model.types.edge_hints[type_name], \
                f"mapreduce is not defined for the hint combination of {type_name}"
        rdt = np.dtype(datatype if datatype is not None else "f8")
        out = np.zeros(1, dtype=rdt)
        ib = np.array([init], dtype=rdt) if init is not None else None
        self._ck(self.lib.vb_mapreduce_fn(self.h, map_name.encode(), C.c_int(ref), C.c_int(_OPS[op]), C.c_int(_DT[rdt]),
                                          ib.ctypes.data_as(C.c_void_p) if ib is not None else None, out.ctypes.data_as(C.c_void_p)))
        return out[0].item()

    # -- observability (src/REPL.jl:81-168, src/optional/DataFrames.jl:54-163) --
    def show(self) -> str:
        """show(sim): the summary the reference prints at the REPL (model name, agents / edges per type, parameters, rasters, globals)"""
        t = self.model.types
        rank, world = C.c_int(0), C.c_int(1)
        if hasattr(self.lib, "vb_comm_rank"):
            self.lib.vb_comm_rank(C.byref(rank), C.byref(world))
        out = [f"Model Name: {self.model.name}"]
        if t.agent_names:
            out.append("Agent(s):")
        for a in t.agent_names:
            line = f"\t Type {a} with {self.num_agents(a)} agent(s)"
            if world.value > 1:
                line += f" ({len(self.all_agentids(a, all_ranks=False))} on rank {rank.value})"
            out.append(line)
        if t.edge_names:
            out.append("Edge(s):")
        for e in t.edge_names:
            out.append(f"\t Type {e} with {self.num_edges(e)} edge(s)")
        if self._params is not None:
            out.append("Parameter(s):")
            for k in self._params.dtype.names:
                out.append(f"\t :{k} : {self._params[k]}")
        if getattr(self, "_rasters", None):
            out.append("Raster(s):")
            for k, dims in self._rasters.items():
                out.append(f"\t :{k} with dimension {tuple(dims)}")
        if self.globals:
            out.append("Global(s):")
            for k, v in self.globals.items():
                if isinstance(v, (list, np.ndarray)):
                    out.append(f"\t :{k} (empty)" if len(v) == 0 else f"\t :{k} |> last : {v[-1]} (length: {len(v)})")
                else:
                    out.append(f"\t :{k} : {v}")
        n = C.c_int64(0)
        self.lib.vb_num_transitions(self.h, C.byref(n))
        if n.value == 0:
            out.append("Still in initialization process!.")
        return "\n".join(out)

    def dataframe(self, type_name: str, types: bool = False, localnr: bool = False):
        """DataFrame(sim, T; types, localnr) (src/optional/DataFrames.jl:54-163) as a pandas DataFrame: agents -> id + one column per
        field; edges -> from, to + one column per field.  As in the reference only the local partition of a parallel run."""
        import pandas as pd
        if localnr:
            types = True
        names = {v: k for k, v in self._aid.items()}
        nr = lambda ids: (ids & np.uint64((1 << BITS_AGENTNR) - 1)) if localnr else ids          # noqa: E731
        tn = lambda ids: [names.get(int(i) >> SHIFT_TYPE) for i in ids]                            # noqa: E731
        if type_name in self._aid:
            states, ids = self._all(type_name)
            df = pd.DataFrame({"id": nr(ids)})
            if states is not None:
                for f in states.dtype.names:
                    df[f] = states[f]
            return df
        to, fr, st = self.all_edges(type_name, all_ranks=False)
        df = pd.DataFrame()
        if "IgnoreFrom" not in self.model.types.edge_hints[type_name]:
            df["from"] = nr(fr)
            if types:
                df["from_type"] = tn(fr)
        df["to"] = nr(to)
        if types:
            df["to_type"] = tn(to)
        if st is not None:
            for f in st.dtype.names:
                df[f] = st[f]
        return df

    def globals_dataframe(self):
        """GlobalsDataFrame(sim) (src/optional/DataFrames.jl:165-185): the vector-valued globals as columns"""
        import pandas as pd
        cols = {k: list(v) for k, v in self.globals.items() if isinstance(v, (list, np.ndarray))}
        n = max([len(v) for v in cols.values()], default=0)
        return pd.DataFrame({k: v + [None] * (n - len(v)) for k, v in cols.items()})

    # -- raster read-out (Raster.jl:206-387) --
    def raster_info(self, name: str):
        nd = C.c_int()
        dims = (C.c_int64 * 4)()
        self._ck(self.lib.vb_raster_info(self.h, name.encode(), C.byref(nd), dims, None))
        return tuple(dims[i] for i in range(nd.value))

    def rastervalues(self, name: str, field: str, type_name: str) -> np.ndarray:
        dims = self.raster_info(name)
        dt = self._adt(type_name)
        fdt, off = dt.fields[field][0], dt.fields[field][1]
        n = int(np.prod(dims))
        if fdt.subdtype is not None:   # tuple-valued field (e.g. pos::Tuple{Int64,Int64}): one pass per component
            base, shape = fdt.subdtype
            k = int(np.prod(shape))
            out = np.zeros((n, k), dtype=base)
            for j in range(k):
                col = np.zeros(n, dtype=base)
                self._ck(self.lib.vb_rastervalues(self.h, name.encode(), C.c_int(off + j * base.itemsize), C.c_int(_DT[base]),
                                                  col.ctypes.data_as(C.c_void_p)))
                out[:, j] = col
            return out.reshape(dims + tuple(shape), order="F") if len(shape) == 0 else \
                np.stack([out[:, j].reshape(dims, order="F") for j in range(k)], axis=-1)
        out = np.zeros(n, dtype=fdt)
        self._ck(self.lib.vb_rastervalues(self.h, name.encode(), C.c_int(off), C.c_int(_DT[fdt]), out.ctypes.data_as(C.c_void_p)))
        return out.reshape(dims, order="F")

    def calc_rasterstate(self, name: str, field: str, type_name: str) -> np.ndarray:
        return self.rastervalues(name, field, type_name)

    def raster_ids(self, name: str) -> np.ndarray:
        """sim.rasters[name]: the grid of cell ids (on several ranks: the grid handed out by finish_init!, src/MPI.jl:59-73)"""
        dims = self.raster_info(name)
        ids = np.zeros(int(np.prod(dims)), dtype=np.uint64)
        nd, d = C.c_int(), (C.c_int64 * 4)()
        self._ck(self.lib.vb_raster_info(self.h, name.encode(), C.byref(nd), d, ids.ctypes.data_as(C.c_void_p)))
        return ids.reshape(dims, order="F")

    def calc_raster(self, name: str, f, f_returns, accessible=()) -> np.ndarray:
        """calc_raster(sim, raster, f, f_returns, accessible) (src/Raster.jl:206-236), the general form: f(id) — any Python callable
        that uses the query API (agentstate, num_edges, neighborids, ...) on the types listed in `accessible` — is evaluated on the
        host for every cell this rank owns, the ranks' values are joined, cells nobody owns stay zero(f_returns).  As slow as the
        reference's loop; the device forms are calc_rasterstate / calc_rasterstate_fn / calc_raster_num_edges."""
        assert getattr(self, "_inited", False), "calc_raster can be only called after finish_init!"
        ids = self.raster_ids(name)
        flat = ids.reshape(-1, order="F")
        rank = self._rank()
        self.disable_transition_checks(True)
        try:
            mine = [(k, f(int(i))) for k, i in enumerate(flat) if process_nr(int(i)) == rank]
        finally:
            self.disable_transition_checks(False)
        idx = np.array([k for k, _ in mine], dtype=np.int64)
        val = np.array([v for _, v in mine], dtype=f_returns) if mine else np.zeros(0, dtype=f_returns)
        if self._world() > 1:
            idx, val = self._join(idx), self._join(val)
        out = np.zeros(flat.shape[0], dtype=f_returns)
        out[idx] = val
        return out.reshape(ids.shape, order="F")

    def calc_rasterstate_fn(self, name: str, map_name: str, datatype="f8") -> np.ndarray:
        """calc_rasterstate(sim, raster, f, f_returns) (src/Raster.jl:238-280) with f a registered map functor of the cells' type,
        e.g. calc_rasterstate(sim, :raster, c -> c.countdown == 0, Bool) = sim.calc_rasterstate_fn("raster", "pp_has_food", "?")"""
        dims = self.raster_info(name)
        rdt = np.dtype(datatype)
        isf = rdt.kind == "f"
        out = np.zeros(int(np.prod(dims)), dtype="f8" if isf else "i8")
        self._ck(self.lib.vb_calc_rasterstate_fn(self.h, name.encode(), map_name.encode(), C.c_int(int(isf)), out.ctypes.data_as(C.c_void_p)))
        return out.astype(rdt).reshape(dims, order="F")

    def calc_raster_num_edges(self, name: str, edge_name: str) -> np.ndarray:
        dims = self.raster_info(name)
        out = np.zeros(int(np.prod(dims)), dtype=np.int64)
        self._ck(self.lib.vb_calc_raster_num_edges(self.h, name.encode(), C.c_int(self._eid[edge_name]), out.ctypes.data_as(C.c_void_p)))
        return out.reshape(dims, order="F")


def create_simulation(model: Model, params: Optional[dict] = None, globals_: Optional[dict] = None,
                      backend: Optional[Backend] = None, device: int = 0, logging: bool = False,
                      log_path: Optional[str] = None) -> Simulation:
    """create_simulation(model, params, globals; logging) (src/Simulation.jl:261-313); `logging` opens log/<name>_<rank>.log."""
    return Simulation(model, params, globals_, backend, device, logging=logging, log_path=log_path)


def apply(sim: Simulation, transition: str, call, read, write, **kwargs) -> Simulation:
    """apply(sim, f, call, read, write; kwargs...) — the non-mutating apply! of src/Simulation.jl:847-856."""
    return sim.apply_copy(transition, call, read, write, **kwargs)


def add_graph(sim: Simulation, edges_uv, n: Optional[int], agent_type: str, agent_states, edge_type: str, edge_states=None,
              directed: bool = False) -> np.ndarray:
    """add_graph!(sim, graph, agent_constructor, edge_constructor) (src/GraphsSupport.jl:34-57): one agent per vertex, then
    for every edge (u,v) of the graph an edge u->v and, for undirected graphs, v->u, in the graph's edge order.
    `edges_uv` is an (m, 2) array of 0-based vertex pairs, or a networkx graph with vertices 0..n-1 (then `n` and `directed` are
    taken from the graph)."""
    if hasattr(edges_uv, "number_of_nodes") and hasattr(edges_uv, "edges"):          # a networkx (Di)Graph
        g = edges_uv
        n = g.number_of_nodes()
        directed = bool(g.is_directed())
        edges_uv = np.array(list(g.edges()), dtype=np.int64).reshape(-1, 2)
    ids = sim.add_agents(agent_type, agent_states, n)
    uv = np.asarray(edges_uv, dtype=np.int64).reshape(-1, 2)
    if directed:
        fr, to = ids[uv[:, 0]], ids[uv[:, 1]]
        st = edge_states
    else:
        fr = np.stack([ids[uv[:, 0]], ids[uv[:, 1]]], axis=1).reshape(-1)
        to = np.stack([ids[uv[:, 1]], ids[uv[:, 0]]], axis=1).reshape(-1)
        st = None if edge_states is None else np.repeat(np.asarray(edge_states), 2)
    sim.add_edges(fr, to, edge_type, st)
    return ids


def vahanagraph(sim: Simulation, agenttypes=None, edgetypes=None, drop_multiedges: bool = False) -> dict:
    """vahanagraph(sim; agenttypes, edgetypes, drop_multiedges) (src/GraphsSupport.jl:212-289) as plain arrays: the live agents of
    `agenttypes` become vertices 0..nv-1 (type order, then ascending nr: `g2v` maps vertex -> AgentID), every edge of `edgetypes`
    whose two ends are vertices becomes (src, dst) with `edgetype` = index into `edgetypes`; :IgnoreFrom types are skipped as in
    the reference.  Returns {"g2v", "src", "dst", "edgetype"}; `to_networkx` wraps it into a networkx.MultiDiGraph for plotting."""
    t = sim.model.types
    agenttypes = list(agenttypes) if agenttypes is not None else list(t.agent_names)
    edgetypes = list(edgetypes) if edgetypes is not None else list(t.edge_names)
    g2v = np.concatenate([sim.all_agentids(a, all_ranks=False) for a in agenttypes]) if agenttypes else np.zeros(0, dtype=np.uint64)
    order = np.argsort(g2v, kind="stable")
    sorted_ids = g2v[order]

    def vertex(ids):
        k = np.searchsorted(sorted_ids, ids)
        k = np.minimum(k, max(len(sorted_ids) - 1, 0))
        ok = (sorted_ids[k] == ids) if len(sorted_ids) else np.zeros(len(ids), dtype=bool)
        return np.where(ok, order[k] if len(sorted_ids) else 0, -1)
    src, dst, et = [], [], []
    seen = set()
    for idx, name in enumerate(edgetypes):
        if "IgnoreFrom" in t.edge_hints[name]:
            continue
        to, fr, _ = sim.all_edges(name, all_ranks=False)
        f, d = vertex(fr), vertex(to)
        keep = (f >= 0) & (d >= 0)
        f, d = f[keep], d[keep]
        if drop_multiedges:
            sel = []
            for i, pair in enumerate(zip(f.tolist(), d.tolist())):
                if pair not in seen:
                    seen.add(pair)
                    sel.append(i)
            f, d = f[sel], d[sel]
        src.append(f)
        dst.append(d)
        et.append(np.full(len(f), idx, dtype=np.int64))
    cat = lambda xs: np.concatenate(xs) if xs else np.zeros(0, dtype=np.int64)   # noqa: E731
    return {"g2v": g2v, "src": cat(src), "dst": cat(dst), "edgetype": cat(et)}


def vahanasimplegraph(sim: Simulation, agenttypes=None, edgetypes=None) -> dict:
    """vahanasimplegraph(sim; agenttypes, edgetypes) (src/GraphsSupport.jl:291-340): the structure only — a simple digraph has no
    parallel edges, so this is vahanagraph with multi-edges dropped"""
    return vahanagraph(sim, agenttypes, edgetypes, drop_multiedges=True)


def to_networkx(sim: Simulation, agenttypes=None, edgetypes=None, drop_multiedges: bool = False):
    """the vahanagraph as a networkx.MultiDiGraph (vertex attribute `id` = AgentID, edge attribute `edgetype`)"""
    import networkx as nx
    vg = vahanagraph(sim, agenttypes, edgetypes, drop_multiedges)
    g = nx.MultiDiGraph()
    for v, aid in enumerate(vg["g2v"].tolist()):
        g.add_node(v, id=aid)
    for s_, d_, e_ in zip(vg["src"].tolist(), vg["dst"].tolist(), vg["edgetype"].tolist()):
        g.add_edge(s_, d_, edgetype=e_)
    return g
