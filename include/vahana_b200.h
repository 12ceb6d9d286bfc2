/* vahana_b200.h — C-ABI of the B200 execution engine for Vahana.jl's transition hot path.
 *
 * The reference (pure Julia, /root/reference) has no FFI seam; the seam is introduced
 * where `create_model` generates per-model methods (src/Simulation.jl:115-207).  A Julia
 * package re-exporting Vahana's names `ccall`s these symbols (julia/VahanaB200.jl,
 * INTEGRATION.md); the Python mirror in vahana.jl_b200/__init__.py binds them with ctypes.
 *
 * Conventions
 *   - every function returns 0 on success, a VB_ERR_* code otherwise; the message is
 *     available through vb_last_error().  VB_ERR_ASSERT marks conditions the reference
 *     reports as `AssertionError` (@mayassert / @assert), so wrappers can re-raise them.
 *   - plain pointers and sizes only; `*_out` buffers are caller allocated; the library
 *     owns all device memory.  Calls are blocking and not re-entrant per simulation
 *     (the reference is single threaded per rank, src/MPIinit.jl:22).
 *   - agent types are referred to by their 1-based type id (registration order,
 *     src/ModelTypes.jl:84-89); edge types by their 0-based registration index.  In the
 *     mixed call/read/write lists of vb_apply an edge type e is VB_EDGE_REF + e.
 *   - states cross the boundary as arrays of the registered struct (AoS, `size` bytes
 *     each, Julia isbits layout); the engine keeps them in SoA word columns on device.
 */
#ifndef VAHANA_B200_H
#define VAHANA_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct vb_sim vb_sim;
typedef uint64_t vb_agent_id; /* src/Agent.jl:30-64: type:8 | rank:20 | nr:36 */

enum { VB_OK = 0, VB_ERR_ASSERT = 1, VB_ERR_ARG = 2, VB_ERR_CUDA = 3, VB_ERR_STATE = 4, VB_ERR_NOTFOUND = 5 };
enum { VB_EDGE_REF = 256 };

/* hints: src/ModelTypes.jl:93-104 (agents), :165-226 (edges) */
enum { VB_AGENT_IMMORTAL = 1, VB_AGENT_INDEPENDENT = 2 };
enum {
    VB_EDGE_STATELESS = 1, VB_EDGE_IGNORE_FROM = 2, VB_EDGE_SINGLE_EDGE = 4, VB_EDGE_SINGLE_TYPE = 8,
    VB_EDGE_IGNORE_SOURCE_STATE = 16
};
enum { VB_METRIC_CHEBYSHEV = 0, VB_METRIC_EUCLIDEAN = 1, VB_METRIC_MANHATTEN = 2 }; /* src/Raster.jl:83 */
enum { VB_OP_SUM = 0, VB_OP_PROD = 1, VB_OP_MIN = 2, VB_OP_MAX = 3, VB_OP_AND = 4, VB_OP_OR = 5 };
enum { VB_DT_I64 = 0, VB_DT_F64 = 1, VB_DT_BOOL = 2, VB_DT_I32 = 3, VB_DT_F32 = 4, VB_DT_U8 = 5 };

typedef struct {
    const char* name;
    uint32_t size;  /* sizeof the isbits state struct; 0 = stateless */
    uint32_t hints; /* VB_AGENT_* */
} vb_agenttype_desc;

typedef struct {
    const char* name;
    uint32_t size;       /* sizeof the edge state struct; 0 = no fields */
    uint32_t hints;      /* VB_EDGE_* (NumEdgesOnly / HasEdgeOnly already expanded by the wrapper) */
    int32_t target_type; /* 1-based agent type for :SingleType, else 0 (ModelTypes.jl:188-206) */
    uint64_t size_hint;  /* the `size` keyword: preallocation only (EdgeMethods.jl:132,221-227) */
} vb_edgetype_desc;

typedef struct {
    const char* name;
    uint32_t n_agent_types;
    const vb_agenttype_desc* agent_types;
    uint32_t n_edge_types;
    const vb_edgetype_desc* edge_types;
    uint32_t param_size; /* bytes of the model's parameter struct (register_param!, Simulation.jl:587) */
} vb_model_desc;

/* ---- library ------------------------------------------------------------------------ */
const char* vb_last_error(void);
const char* vb_backend(void);                 /* "cuda-sm100a" (product) or "oracle-cpu" (tests only) */
int vb_init(int device);                      /* mpiinit analogue, src/MPIinit.jl:21: rank <-> GPU */
int vb_shutdown(void);
int vb_set_stream(void* cuda_stream);         /* run on the caller's CUDA stream (cudaStream_t) instead of the engine's own */
/* multi-GPU: one process per GPU; the 128-byte NCCL unique id is created on rank 0 and
 * handed to the other ranks by the host launcher (torch.distributed store / MPI / file). */
int vb_comm_unique_id(uint8_t id_out[128]);
int vb_comm_init(int rank, int nranks, const uint8_t id[128]);
int vb_comm_rank(int* rank_out, int* nranks_out);
int vb_set_uniform_offset(vb_sim* sim, int type, uint64_t offset); /* global row of this rank's first agent in the per-agent uniform table */
int vb_halo_bytes(vb_sim* sim, uint64_t* bytes_out); /* bytes this rank received in the halo exchanges of the last apply */
/* Device time of the last apply's halo exchange on its own stream: from behind the entry barrier to behind the last phase's barrier
   (transmit_agents!, src/MPI.jl:155-267); negative when the last apply exchanged nothing through peer memory. */
int vb_last_halo_ms(vb_sim* sim, double* ms_out);

/* ---- lifecycle: create_simulation / copy_simulation / finish_simulation!
 *      (src/Simulation.jl:261, :500, :551) --------------------------------------------- */
int vb_sim_create(const vb_model_desc* model, const void* params, vb_sim** sim_out);
int vb_sim_copy(const vb_sim* sim, vb_sim** sim_out);
int vb_sim_destroy(vb_sim* sim);
int vb_set_param(vb_sim* sim, const void* params, uint32_t size);          /* set_param!, Simulation.jl:602 */
int vb_set_config(vb_sim* sim, int asserts_enabled, int check_readable);   /* src/Vahana.jl:42-84 */
int vb_disable_transition_checks(vb_sim* sim, int disable);                /* src/Helpers.jl:259 */

/* ---- init phase (bulk): add_agent(s)!, add_edge(s)!, rasters, finish_init! ----------- */
int vb_add_agents(vb_sim* sim, int type, const void* states, uint64_t n, vb_agent_id* ids_out); /* AgentMethods.jl:65 */
int vb_add_agent_per_process(vb_sim* sim, int type, const void* state, vb_agent_id* id_out);    /* Agent.jl:363-388 */
int vb_add_edges(vb_sim* sim, int etype, const vb_agent_id* from, const vb_agent_id* to, const void* states,
                 uint64_t n);                                                                    /* EdgeMethods.jl:388-523 */
int vb_remove_edges(vb_sim* sim, int etype, vb_agent_id from /*0 = all*/, vb_agent_id to);      /* EdgeMethods.jl:527-599 */
int vb_add_raster(vb_sim* sim, const char* name, int ndims, const int64_t* dims, int type, const void* states,
                  vb_agent_id* ids_out);                                                         /* Raster.jl:32-54 */
/* The receiver side of broadcastids (src/MPI.jl:59-73, src/Raster.jl:64-75): a raster over agents that exist already — `ids` in
   CartesianIndices (column-major) order, some of them agents of other ranks after finish_init!(distribute = true) handed the cells
   out.  Init phase.  calc_raster / calc_rasterstate / rastervalues of such a raster join the ranks (src/Raster.jl:227,318,378). */
int vb_set_raster(vb_sim* sim, const char* name, int ndims, const int64_t* dims, int type, const vb_agent_id* ids);
int vb_connect_raster_neighbors(vb_sim* sim, const char* name, int etype, double distance, int metric, int periodic,
                                const void* edge_state);                                         /* Raster.jl:139-167 */
int vb_move_to(vb_sim* sim, const char* name, vb_agent_id id, const int64_t* pos, int etype_from /*-1 = nothing*/,
               const void* state_from, int etype_to /*-1 = nothing*/, const void* state_to, double distance,
               int metric, int periodic, int only_surrounding);                                  /* Raster.jl:437-477 */
int vb_cellid(vb_sim* sim, const char* name, const int64_t* pos, vb_agent_id* id_out);           /* Raster.jl:403 */
int vb_finish_init(vb_sim* sim);                                                                 /* Simulation.jl:403-476 */

/* ---- transitions: apply! (src/Simulation.jl:720-821) ----------------------------------
 * `transition` names a functor set registered by a model library (VB_REGISTER_TRANSITION in
 * include/vahana_device.cuh): one compiled launcher per (name, called agent type) — the
 * counterpart of Julia dispatching the closure on the agent's type.  `seed` keys the
 * counter-based uniform table ctx.uniform(k) reads (Philox4x32-10 of (seed, slot, k)).   */
int vb_apply(vb_sim* sim, const char* transition, const int* call, int ncall, const int* read, int nread,
             const int* write, int nwrite, const int* add_existing, int nadd, int with_edge /*-1 = none*/,
             uint64_t seed);
int vb_has_transition(const char* transition, const char* agent_type_name);
int vb_load_model_library(const char* path); /* dlopen a model .so whose static init registers transitions */

/* ---- queries (device -> host) ---------------------------------------------------------- */
int vb_num_agents(vb_sim* sim, int type, uint64_t* n_out);                                 /* Agent.jl:324-343 */
int vb_all_agents(vb_sim* sim, int type, void* states_out, vb_agent_id* ids_out, uint64_t cap, uint64_t* n_out); /* :234-313 */
int vb_agentstate(vb_sim* sim, vb_agent_id id, int type, void* state_out);                 /* AgentMethods.jl:91-154 */
int vb_num_edges_total(vb_sim* sim, int etype, int write, uint64_t* n_out);                /* Edge.jl:373-389 */
/* accessor family, EdgeMethods.jl:704-892.  `what` selects which reference accessor's
 * availability rule (docs/src/performance.md:129-136) is enforced. */
enum { VB_ACC_EDGES = 0, VB_ACC_NEIGHBORIDS = 1, VB_ACC_EDGESTATES = 2, VB_ACC_NUM_EDGES = 3, VB_ACC_HAS_EDGE = 4,
       VB_ACC_NEIGHBORIDS_ITER = 5, VB_ACC_EDGESTATES_ITER = 6 };
int vb_edges_of(vb_sim* sim, int etype, vb_agent_id to, int what, vb_agent_id* from_out, void* states_out,
                uint64_t cap, int64_t* n_out /* -1 = `nothing` */);
int vb_all_edges(vb_sim* sim, int etype, vb_agent_id* to_out, vb_agent_id* from_out, void* states_out, uint64_t cap,
                 uint64_t* n_out);                                                          /* EdgeMethods.jl:1005 */

/* ---- reductions: mapreduce (AgentMethods.jl:533-565, EdgeMethods.jl:972-994) -----------
 * map = "take the field at byte `offset` with dtype `dt`" (the form every reference test and
 * docs example uses, e.g. a -> a.foo), optionally compared `== cmp` first (c -> c.countdown == 0),
 * or a constant 1 (_ -> 1) when dt < 0.  `type_ref` is an agent type id or VB_EDGE_REF + e.
 * init: pointer to one value of the result dtype or NULL (=> identity, Helpers.jl:44-81).  */
int vb_mapreduce(vb_sim* sim, int type_ref, int offset, int dt, int has_cmp, int64_t cmp, int op, int result_dt,
                 const void* init, void* result_out);
/* The same with the map given by name: a functor registered with VB_REGISTER_MAP (include/vahana_device.cuh, vb::MapBase in
 * include/vahana_model.h) for the agent / edge type `type_ref`, i.e. any closure f of mapreduce(sim, f, op, T), e.g.
 * b -> b.x - b.y (docs/examples/tutorial1.jl:548).  result_dt must be floating point iff the functor's Result is.            */
int vb_mapreduce_fn(vb_sim* sim, const char* map_name, int type_ref, int op, int result_dt, const void* init, void* result_out);

/* ---- raster read-out (src/Raster.jl:206-387) ------------------------------------------- */
int vb_rastervalues(vb_sim* sim, const char* name, int offset, int dt, void* out);       /* rastervalues / calc_rasterstate(field) */
int vb_calc_raster_num_edges(vb_sim* sim, const char* name, int etype, int64_t* out);    /* calc_raster(id -> num_edges(sim,id,E)) */
/* calc_rasterstate(sim, raster, f, f_returns) (src/Raster.jl:238-280) with f a registered map functor (VB_REGISTER_MAP) of the cells'
 * agent type: out[i] = f(state of the cell at column-major position i), as double (is_float_out = 1) or int64_t (0).             */
int vb_calc_rasterstate_fn(vb_sim* sim, const char* name, const char* map_name, int is_float_out, void* out);
int vb_raster_info(vb_sim* sim, const char* name, int* ndims_out, int64_t* dims_out, vb_agent_id* ids_out);

/* ---- introspection used by tests/bench -------------------------------------------------- */
int vb_num_transitions(vb_sim* sim, int64_t* n_out);                                    /* Simulation.jl:310,473,816 */
/* CSR of one edge type as stored (rows = target slots of `target_type`): bit-exact parity checks. */
int vb_export_csr(vb_sim* sim, int etype, int target_type, uint64_t* offsets_out /*nrows+1*/, uint64_t nrows,
                  vb_agent_id* from_out, void* states_out, uint64_t cap);
int vb_last_apply_stats(vb_sim* sim, double* ms_read_write, double* ms_finish, uint64_t* edges_read,
                        uint64_t* edges_appended, uint64_t* agents_called, uint64_t* kernel_launches);
uint64_t vb_device_view_bytes(void);               /* bytes of the simulation view uploaded host->device per transition launch */
int vb_last_kernel_ms(vb_sim* sim, double* ms_out); /* CUDA-event time of the transition kernels of the last apply */
/* Policy of the source-blocked read phase of reduce transitions (include/vahana_model.h, vb::ReduceTransition): block size in MB of
   source states (0 disables), activation threshold in MB of the source type's state array, eager = build at first sight instead of
   waiting for a container that survived one apply.  Negative values keep the defaults (VB_BLOCK_MB / VB_BLOCK_MIN_MB / VB_BLOCK_EAGER). */
int vb_set_read_blocking(vb_sim* sim, double block_mb, double min_mb, int eager);
/* Source blocks swept by the read phase of the last apply (0 = direct path): which kernel shape ran (read phase). */
int vb_last_apply_blocks(vb_sim* sim, uint32_t* nblocks_out);
/* Prefiltered sweeps for reduce transitions that name a one-byte key of the neighbour state (kPrefilter, include/vahana_model.h):
   1 = on, 0 = off, negative = default (on; VB_PREFILTER=0 turns it off).  Results do not depend on it.  The second call reports
   whether the last apply's sweeps were prefiltered. */
int vb_set_read_prefilter(vb_sim* sim, int on);
int vb_last_apply_prefiltered(vb_sim* sim, int* on_out);
/* Share of sampled edges whose source key the target's probe let through, as estimated by the last apply that checked (every 16th apply
   of a prefilter-capable reduce transition; negative = never estimated).  Above VB_PF_MAX_PASS (0.30) the engine takes the unfiltered
   sweeps — a key that passes costs a DRAM-random fetch of the exact state, e.g. ϵ = 0.25 in hegselmann.jl:164 — below VB_PF_MIN_PASS
   (0.22) it returns to the prefiltered ones.  The reference has no counterpart (its loop visits every neighbour state). */
int vb_last_pass_rate(vb_sim* sim, double* rate_out);

#ifdef __cplusplus
}
#endif
#endif
