// vahana_device.cuh — device-side API of the B200 engine: what a model's transition functor sees
// (the "Ctx concept"), the kernels that run it, and the registration macro that turns a functor
// into the handle vb_apply launches.
//
// A transition is the CUDA counterpart of the Julia closure `f(state, id, sim)` that the reference
// calls once per agent (src/AgentMethods.jl:193-207, src/Simulation.jl:774-788):
//
//     struct Step : vb::TransitionBase {
//         using State = HKAgent;                           // struct of the called agent type
//         static constexpr bool kCooperative = true;       // warp-per-agent gather (else thread-per-agent)
//         using EdgeWrites = vb::IntList<E_FOO>;           // edge types add_edge is called on (must be in `write`)
//         using AgentWrites = vb::IntList<T_BAR>;          // agent types add_agent is called on
//         template <class Ctx> VB_HD bool operator()(Ctx& ctx, State& self, vb::AgentID id) const;
//     };                                                   // return false == `nothing` (the agent dies)
//
// Ctx concept (identical surface in oracle/vahana_oracle.hpp, class vo::Ctx):
//     param<P>()                                   param(sim, ...)              src/Simulation.jl:587
//     uniform(k)                                   k-th pre-generated uniform of this agent (Philox table)
//     num_edges(E,id) has_edge(E,id)               src/EdgeMethods.jl:850-892
//     for_each_neighbor(E,id,f(from))              neighborids(_iter)           :717-761
//     neighbor_at(E,id,k)                          neighborids(...)[k+1]
//     for_each_edge<S>(E,id,f(from,state))         edges                        :704-715
//     for_each_edgestate<S>(E,id,f(state))         edgestates(_iter)            :804-848
//     agentstate<A>(T,id)  agentfield<F>(T,id,off) agentstate(_flexible)        src/AgentMethods.jl:91-154
//     for_each_neighborstate<A>(E,T,id,f(state))   neighborstates(_iter)        :764-802
//     add_edge(E,from,to[,state])                  add_edge!                    :388-523
//     add_agent<A>(T,state) -> id                  add_agent!                   src/AgentMethods.jl:65-89
//     move_to(raster,id,pos,Efrom,Eto,...) cellid(raster,pos)                   src/Raster.jl:403-477
//     remove_edges(E,to) remove_edges(E,from,to)   remove_edges!                src/EdgeMethods.jl:527-599 (a target of another rank: the request travels)
//     require(cond)                                @assert cond inside the closure: apply! raises an AssertionError
//     sum(x) max(x) min(x) lanes() lane() leader() cooperative group of the agent (1 lane in the oracle)
//
// In a cooperative functor all lanes of the group run the functor for the same agent; for_each_* hand
// each lane a strided share of the row and sum()/max()/min() combine the lanes' partials.  Writes
// (state, add_edge, add_agent) are performed by the leader lane and must sit outside for_each_*.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <type_traits>

#include "vahana_model.h"

namespace vb {

enum : int { MAX_AGENT_TYPES = 48, MAX_EDGE_TYPES = 64, MAX_RASTERS = 4, MAX_PARAM_BYTES = 512 };   // the view travels in the kernel parameter block (<= 32 764 B)
enum : int { MAX_EDGE_WRITES = 6, MAX_AGENT_WRITES = 3, MAX_EDGE_REMOVES = 3 };
enum EdgeKind : uint8_t { KIND_CSR = 0, KIND_COUNT = 1, KIND_FLAG = 2, KIND_STENCIL = 3 };
enum : int { MAX_IMPLICIT_STENCIL = 32 };
enum Mode : int { MODE_DIRECT = 0, MODE_COUNT = 1, MODE_EMIT = 2 };
enum DevError : uint32_t {
    DERR_EDGE_NOT_READABLE = 1, DERR_AGENT_NOT_READABLE = 2, DERR_AGENT_TYPE_MISMATCH = 4, DERR_AGENT_DIED = 8,
    DERR_IMMORTAL_DIED = 16, DERR_ACCESSOR_UNAVAILABLE = 32, DERR_BAD_ID = 64, DERR_EDGE_NOT_DECLARED = 128,
    DERR_SINGLETYPE_MISMATCH = 256, DERR_RASTER_POS = 512, DERR_INDEX = 1024, DERR_REMOTE = 2048,
    DERR_MODEL_ASSERT = 4096
};

// ---- device views of the simulation state (filled by the engine, read by every kernel) -----------
// Composite index: every agent slot of the rank has one dense 32-bit index  comp = base[type] + slot
// (slot = nr - 1).  CSR columns (edge sources), CSR rows of edge types without :SingleType, and the
// append logs all use it: 4 B per reference instead of an 8 B AgentID (SURVEY.md §8d byte model).
struct AgentView {
    uint8_t* state_r;         // SoA word columns: column c at state + c * cap * word
    uint8_t* state_w;
    uint8_t* died_r;          // 1 B per slot, nullptr for :Immortal
    uint8_t* died_w;
    const uint32_t* reuse;    // read.reuseable (slots, 0-based), LIFO: pop from the end
    uint32_t cap;             // column stride of the state buffers = local capacity + ghost capacity
    uint32_t lcap;            // local capacity: slots [0, lcap) are this rank's agents, [lcap, lcap + nghost) are ghosts
    uint32_t nghost;          // remote agents (other ranks) whose state is mirrored here by the halo exchange
    const uint64_t* ghost_ids;// AgentIDs of the ghosts, ascending (rank-major): ghost slot = lcap + position
    uint32_t nslots_r;        // length(read.state)
    uint32_t n_reuse;         // entries of `reuse` still available to this call (after earlier pops)
    uint32_t next0;           // first never-used slot (= nextid - 1) before this call's births
    uint32_t size, word, ncols;
    uint8_t immortal, independent, readable, writeable;
    uint64_t uoffset;         // row offset of this rank's agents in the global per-agent uniform table (multi-GPU)
};
struct EdgeView {
    // read container
    const uint32_t* off;      // KIND_CSR: row offsets [rows + 1]
    const uint32_t* src;      // composite source per entry (nullptr for :IgnoreFrom)
    const uint8_t* st;        // SoA state columns per entry: column c at st + c * st_cap * word
    const uint32_t* cnt;      // KIND_COUNT: edges per row; KIND_FLAG: 0/1 per row
    uint32_t rows;            // rows of the read container
    uint32_t st_cap;          // column stride of `st` (entries)
    // write container (append log) of the running apply
    uint32_t* log_to;         // composite row per appended edge
    uint32_t* log_from;
    uint8_t* log_st;          // SoA columns, stride log_cap
    uint32_t* wcnt;           // KIND_COUNT / KIND_FLAG write container (rows_w entries)
    uint32_t log_cap;
    uint32_t rows_w;
    // multi-GPU: appended edges whose target lives on another rank (or whose source is a remote agent this rank does not mirror
    // yet) travel as AgentIDs: the `storage` of the reference (src/EdgeMethods.jl:396-398), exchanged after the transition
    uint64_t* rlog_to; uint64_t* rlog_from; uint8_t* rlog_st; uint32_t* rlog_dst; uint32_t rlog_cap;
    // remove_edges! records of the running apply: (row, source composite or 0xffffffff = all, append position at call time)
    uint32_t* rm_row; uint32_t* rm_from; uint32_t* rm_mark;
    // multi-GPU: a record whose target lives on another rank keeps its AgentIDs here (rm_to64 = 0 marks a local record) and is
    // exchanged after the transition loop (removeedges_alltoall!, src/MPI.jl:432-479); nullptr on one rank
    uint64_t* rm_to64; uint64_t* rm_from64;
    uint32_t size, word, ncols;
    int32_t target;           // :SingleType target type id, else 0
    uint8_t hints, kind, readable, writeable;
    // KIND_STENCIL: the edges of connect_raster_neighbors! (src/Raster.jl:139-167) kept implicit: row c holds the cells
    // c - s (wrapped or clipped) for every stencil offset s, ordered by (source linear index, s) = the reference's insertion order
    int32_t st_tab;           // index into DeviceSim::stencils: offsets in _stencil_core order (src/Raster.jl:82-96)
    int32_t st_n, st_raster;
    uint32_t st_slot0;        // slot of raster cell 0 (cells occupy consecutive slots)
    uint8_t st_periodic, st_reach;
};
struct RasterView {
    const uint32_t* cells;    // composite index of the cell agent at each position (column-major); 0xffffffff = a cell of another rank
    const uint64_t* cell_ids; // AgentIDs of the cells when some of them live on other ranks (else nullptr)
    int32_t ndims;
    int32_t type;             // agent type of the cells
    int64_t dims[MAX_RASTER_DIMS];
    uint32_t dim32[MAX_RASTER_DIMS];    // dims / column-major strides as 32-bit values (unused dimensions: 1 / 0), and the cell count
    uint32_t stride32[MAX_RASTER_DIMS];
    uint32_t ncells;
};
struct StencilTab {                           // lives in the kernel parameter block (constant bank)
    int32_t lin[MAX_IMPLICIT_STENCIL];        // linear offset of each stencil entry (sum of off[k] * stride[k])
    int8_t off[MAX_IMPLICIT_STENCIL][MAX_RASTER_DIMS];
};
struct DeviceSim {
    AgentView agents[MAX_AGENT_TYPES + 1];   // index = type id (1-based)
    EdgeView edges[MAX_EDGE_TYPES];
    RasterView rasters[MAX_RASTERS];
    StencilTab stencils[MAX_RASTERS];
    uint32_t base[MAX_AGENT_TYPES + 2];      // composite base per type id; base[ntypes + 1] = total
    uint32_t n_agent_types, n_edge_types, n_rasters;
    uint32_t rank, nranks;
    uint32_t check;                           // asserts_enabled && check_readable
    uint32_t* error;                          // device word, OR of DevError bits
    uint64_t seed;
    alignas(16) uint8_t params[MAX_PARAM_BYTES];
};

// per-launch arguments of one transition kernel
struct LaunchArgs {
    const DeviceSim* ds;      // HOST pointer: the launcher copies the view into the kernel's parameter block
    int mode;                 // Mode
    int type;                 // called agent type id
    uint32_t n;               // slots to visit (= nslots_r of the type)
    int in_read, in_write;    // C in read / C in write
    int with_edge;            // -1 or edge type: only agents with an edge of this type are called
    uint32_t* ecount[MAX_EDGE_WRITES];    // COUNT: per-agent add_edge counts; EMIT: exclusive offsets into the log
    uint32_t ebase[MAX_EDGE_WRITES];      // EMIT: log position where this call's appends start
    uint32_t* ercount[MAX_EDGE_WRITES];   // multi-GPU: the same for edges that leave the rank (nullptr on one rank)
    uint32_t erbase[MAX_EDGE_WRITES];
    uint32_t* acount[MAX_AGENT_WRITES];   // COUNT: per-agent add_agent counts; EMIT: exclusive offsets
    uint32_t abase[MAX_AGENT_WRITES];     // EMIT: births of that type by earlier calls of this apply
    uint32_t* rcount[MAX_EDGE_REMOVES];   // COUNT: per-agent remove_edges counts; EMIT: exclusive offsets into the remove log
    uint32_t rbase[MAX_EDGE_REMOVES];     // EMIT: remove-log position where this call's records start
    uint32_t rmark[MAX_EDGE_REMOVES];     // EMIT: append-log length of that edge type before this call (for functors that only remove)
    unsigned long long* stats;            // 1024 counters, 4 words apart: edges read (summed by the host)
    // degree binning of cooperative, write-free transitions (DESIGN.md "read phase"):
    int group;                            // lanes per agent: 0 = default (32 cooperative / 1 otherwise), 8, 32 or 256
    int primary_edge;                     // edge type whose row length classifies an agent (-1 = none)
    uint32_t heavy_min;                   // group < 256: skip agents with >= heavy_min edges (they run in the block pass)
    const uint32_t* rows;                 // group == 256: list of agent slots to run (one block each)
    // source-blocked read phase of a reduce transition (one launch per block of the source-state array):
    const uint32_t* blk_off;              // [n + 1] positions into blk_src of the rows' entries whose source lies in this block
    const uint32_t* blk_src;              // source slots (slot inside the source type = composite - base), rows in order
    uint8_t* blk_acc;                     // accumulators parked between sweeps: SoA words, column c at blk_acc + c * blk_stride * word
    uint32_t blk_stride;
    int blk_first, blk_last;              // first sweep initialises the accumulators, last sweep runs finish()
    const uint32_t* blk_heavy;            // bitmap of rows left to the block-per-agent pass (>= heavy_min entries), or nullptr
    uint32_t blk_ahead;                   // CTAs of look-ahead of the L2 prefetch (about one residency wave), 0 = off
    const uint32_t* blk_rows;             // middle sweeps: ascending list of the rows that own an entry in this block (else nullptr:
    uint32_t blk_nrows;                   //   all n rows are swept); rows outside the list keep their parked accumulator untouched
    const uint32_t* blk_roff;             // with a row list: [blk_nrows + 1] entry positions of the listed rows (their entries are contiguous)
    uint8_t* blk_key;                     // prefilter (F::kPrefilter): one key byte per slot of the source type, written by launch_keys
    uint32_t blk_nkeys;                   //   slots of the source type covered by blk_key (local + ghosts)
    uint32_t blk_key_first;               //   launch_keys: first slot of the range to (re)build, blk_nkeys slots from there
    int blk_prefilter;                    //   sweeps gather keys and fetch the exact state only where may_accept() holds
    // segmented view (prefiltered sweeps): rows longer than the segment length are cut into segments, a warp owns 32
    // consecutive segments, an entry is (source slot - blk_base) | (segment & 31) << 27, accumulators are parked per segment
    const uint32_t* blk_seg_row;          // [blk_nseg] slot of the called agent a segment belongs to | SEG_MULTI (its row has several segments)
    uint32_t blk_nseg;
    uint32_t blk_base;                    // first source slot of the swept block
    const uint32_t* blk_hub_rows;         // blk_op == 1: the rows with several segments, ascending
    const uint32_t* blk_hub_seg;          //   [2 blk_nhub] their segment ranges [first, end)
    uint32_t blk_nhub;
    int blk_op;                           // 0 = sweep one block, 1 = merge the hub rows' segment accumulators and finish them
    cudaStream_t stream;
};
// what the kernel actually receives: launch arguments + the whole simulation view, by value in the
// parameter block (constant bank): no separate upload, no pointer chase to reach a container.
struct KernelArgs {
    LaunchArgs la;
    DeviceSim ds;
};

// host-side descriptor of a compiled transition (one per (name, agent type))
struct TransitionInfo {
    const char* name;
    const char* agent_type;
    uint32_t state_size;
    int cooperative;
    int n_edge_writes, edge_writes[MAX_EDGE_WRITES];
    int n_agent_writes, agent_writes[MAX_AGENT_WRITES];
    int n_edge_removes, edge_removes[MAX_EDGE_REMOVES];
    int primary_edge;         // F::kPrimaryEdge (-1 = none): enables degree binning of the read phase
    cudaError_t (*launch)(const LaunchArgs&);
    // vb::ReduceTransition functors only (else reduce = 0): what the source-blocked path needs to know
    int reduce;               // F::kReduce
    int source_type;          // F::kSourceType
    uint32_t source_size;     // sizeof(F::Source)
    uint32_t acc_bytes;       // F::kAccBytes: bytes of the accumulator parked per row between sweeps
    cudaError_t (*launch_blocked)(const LaunchArgs&);
    cudaError_t (*launch_stencil)(const LaunchArgs&);   // reduce transition whose primary edge type is an implicit raster stencil
    int prefilter;            // F::kPrefilter: the functor names a one-byte key of the source state (include/vahana_model.h)
    cudaError_t (*launch_keys)(const LaunchArgs&);      // fills blk_key[0 .. blk_nkeys) from the source type's read states
    cudaError_t (*launch_passrate)(const LaunchArgs&);  // samples entries of the primary edge type: how many keys may_accept() lets through (la.stats[0..1])
};
// exported by libvahana_b200.so; model libraries call it from static initialisers
extern "C" int vb_register_transition(const TransitionInfo* info);

// a compiled map functor (vb::MapBase): one kernel folds f(element) per CTA into `partial[blockIdx.x]` (double or long long)
struct MapLaunchArgs {
    const uint8_t* cols;      // SoA columns of the mapped states (agents: read buffer; edges: CSR state columns)
    uint32_t stride;          // column stride in elements
    uint64_t n;               // elements
    const uint8_t* died;      // agents of a mortal type: skip died slots (else nullptr)
    int op;                   // vb::OP_*
    void* partial;            // [nblocks] partial results
    uint32_t nblocks;
    cudaStream_t stream;
};
// the same functor evaluated per raster cell without folding (calc_rasterstate, src/Raster.jl:238-280): out[i] = f(state of cell i)
struct MapCellsArgs {
    const uint8_t* cols; uint32_t stride;   // read states of the cells' agent type
    const uint32_t* cells; uint32_t cbase;  // composite index of the cell at each column-major position, and the type's base
    uint64_t n;
    void* out;                              // [n] double or long long
    cudaStream_t stream;
};
struct MapInfo {
    const char* name;
    const char* type_name;    // agent or edge type the map is defined on (registered name)
    uint32_t elem_size;       // sizeof(F::Elem)
    int is_float;             // F::Result is floating point (partials are double) or integral (long long)
    cudaError_t (*launch)(const MapLaunchArgs&);
    cudaError_t (*launch_cells)(const MapCellsArgs&);
};
extern "C" int vb_register_map(const MapInfo* info);

#if defined(__CUDACC__)
// ---- SoA word access ---------------------------------------------------------------------------------
template <int W> struct WordT;
template <> struct WordT<1> { typedef uint8_t type; };
template <> struct WordT<2> { typedef uint16_t type; };
template <> struct WordT<4> { typedef uint32_t type; };
template <> struct WordT<8> { typedef uint64_t type; };
template <> struct WordT<16> { typedef uint4 type; };
template <int S> struct SoaWord { static constexpr int value = (S % 16 == 0) ? 16 : (S % 8 == 0) ? 8 : (S % 4 == 0) ? 4 : (S % 2 == 0) ? 2 : 1; };

template <class T>
__device__ __forceinline__ T soa_load(const uint8_t* __restrict__ cols, uint32_t stride, uint32_t i) {
    constexpr int W = SoaWord<sizeof(T)>::value;
    constexpr int NC = sizeof(T) / W;
    typedef typename WordT<W>::type Word;
    union { T t; Word w[NC]; } u;
#pragma unroll
    for (int c = 0; c < NC; ++c) u.w[c] = reinterpret_cast<const Word*>(cols + (size_t)c * stride * W)[i];
    return u.t;
}
// random gather of one record: bypass L1 allocation and ask L2 for 64 B instead of the default 128 B per miss
// (profiles/microbench/gather.cu: 114 B -> 60 B of DRAM traffic per random 8 B load on B200)
template <int W> __device__ __forceinline__ typename WordT<W>::type ld_gather_word(const void* p) { return *reinterpret_cast<const typename WordT<W>::type*>(p); }
template <> __device__ __forceinline__ uint64_t ld_gather_word<8>(const void* p) {
    uint64_t r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::64B.b64 %0, [%1];" : "=l"(r) : "l"(p));
    return r;
}
template <> __device__ __forceinline__ uint32_t ld_gather_word<4>(const void* p) {
    uint32_t r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::64B.b32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}
template <class T>
__device__ __forceinline__ T soa_gather(const uint8_t* __restrict__ cols, uint32_t stride, uint32_t i) {
    constexpr int W = SoaWord<sizeof(T)>::value;
    constexpr int NC = sizeof(T) / W;
    typedef typename WordT<W>::type Word;
    union { T t; Word w[NC]; } u;
#pragma unroll
    for (int c = 0; c < NC; ++c) u.w[c] = ld_gather_word<W>(cols + (size_t)c * stride * W + (size_t)i * W);
    return u.t;
}
template <class T>
__device__ __forceinline__ void soa_store(uint8_t* __restrict__ cols, uint32_t stride, uint32_t i, const T& v) {
    constexpr int W = SoaWord<sizeof(T)>::value;
    constexpr int NC = sizeof(T) / W;
    typedef typename WordT<W>::type Word;
    union U { T t; Word w[NC]; __device__ U() {} } u;
    u.t = v;
#pragma unroll
    for (int c = 0; c < NC; ++c) reinterpret_cast<Word*>(cols + (size_t)c * stride * W)[i] = u.w[c];
}
// bytes [off, off + sizeof(F)) of a record stored in columns of `word` bytes
template <class F>
__device__ __forceinline__ F soa_load_field(const uint8_t* __restrict__ cols, uint32_t stride, uint32_t word, uint32_t i, uint32_t off) {
    F out;
    uint8_t* o = reinterpret_cast<uint8_t*>(&out);
#pragma unroll
    for (uint32_t b = 0; b < sizeof(F); ++b) {
        const uint32_t p = off + b, c = p / word;
        o[b] = cols[(size_t)c * stride * word + (size_t)i * word + (p - c * word)];
    }
    return out;
}

// ---- the CUDA context -----------------------------------------------------------------------------------
// GROUP lanes cooperate on one agent: 1 (thread per agent) or 32 (warp per agent).
template <class F, int MODE, int GROUP>
class Ctx {
  public:
    const DeviceSim& ds;
    const LaunchArgs& la;
    uint32_t slot;       // slot of the called agent
    uint32_t lane_;
    uint32_t ecnt[F::EdgeWrites::size + 1];
    uint32_t ercnt[F::EdgeWrites::size + 1];   // appended edges that leave the rank
    uint32_t acnt[F::AgentWrites::size + 1];
    uint32_t rcnt[F::EdgeRemoves::size + 1];
    unsigned long long edges_read = 0;

    __device__ Ctx(const DeviceSim& d, const LaunchArgs& l, uint32_t s, uint32_t lane) : ds(d), la(l), slot(s), lane_(lane) {
#pragma unroll
        for (int i = 0; i <= F::EdgeWrites::size; ++i) { ecnt[i] = 0; ercnt[i] = 0; }
#pragma unroll
        for (int i = 0; i <= F::AgentWrites::size; ++i) acnt[i] = 0;
#pragma unroll
        for (int i = 0; i <= F::EdgeRemoves::size; ++i) rcnt[i] = 0;
    }
    __device__ __forceinline__ void fail(uint32_t code) const { atomicOr(ds.error, code); }
    // @assert cond inside a transition closure: apply! raises an AssertionError after the launch (the kernel runs on)
    __device__ __forceinline__ void require(bool cond) const { if (!cond) fail(DERR_MODEL_ASSERT); }

    template <class P> __device__ __forceinline__ const P& param() const { return *reinterpret_cast<const P*>(ds.params); }
    __device__ __forceinline__ double uniform(int k) const { return Philox::uniform(ds.seed, ds.agents[la.type].uoffset + slot, (uint64_t)k); }
    __device__ __forceinline__ int lanes() const { return GROUP; }
    __device__ __forceinline__ int lane() const { return (int)lane_; }
    __device__ __forceinline__ bool leader() const { return lane_ == 0; }
    // group reductions: sub-warp / warp groups shuffle (exited lanes of other groups are not named by a
    // non-exited thread's butterfly partners: groups are aligned powers of two), block groups go through smem
    struct OpSum { template <class T> __device__ __forceinline__ T operator()(T a, T b) const { return a + b; } };
    struct OpMax { template <class T> __device__ __forceinline__ T operator()(T a, T b) const { return a > b ? a : b; } };
    struct OpMin { template <class T> __device__ __forceinline__ T operator()(T a, T b) const { return a < b ? a : b; } };
    template <class T, class Op> __device__ __forceinline__ T reduce(T v, Op op) const {
        if (GROUP == 1) return v;
        constexpr int W = GROUP < 32 ? GROUP : 32;
        // all lanes of one group run the functor in lock step, so the group's own lanes are a safe shuffle mask
        const unsigned mask = GROUP >= 32 ? 0xffffffffu : (((1u << W) - 1u) << ((threadIdx.x & 31u) / W * W));
#pragma unroll
        for (int o = W / 2; o > 0; o >>= 1) v = op(v, __shfl_xor_sync(mask, v, o));
        if (GROUP > 32) {
            __shared__ unsigned long long red[GROUP > 32 ? GROUP / 32 : 1];
            static_assert(sizeof(T) <= 8, "block reduction of values up to 8 bytes");
            __syncthreads();
            if ((threadIdx.x & 31) == 0) { unsigned long long u = 0; memcpy(&u, &v, sizeof(T)); red[threadIdx.x >> 5] = u; }
            __syncthreads();
            T r; { unsigned long long u = red[0]; memcpy(&r, &u, sizeof(T)); }
#pragma unroll
            for (int i = 1; i < GROUP / 32; ++i) { T x; unsigned long long u = red[i]; memcpy(&x, &u, sizeof(T)); r = op(r, x); }
            v = r;
        }
        return v;
    }
    template <class T> __device__ __forceinline__ T sum(T v) const { return reduce(v, OpSum()); }
    template <class T> __device__ __forceinline__ T max(T v) const { return reduce(v, OpMax()); }
    template <class T> __device__ __forceinline__ T min(T v) const { return reduce(v, OpMin()); }
    // vb::ReduceTransition: combine the lanes' partial accumulators (butterfly of merge(); every lane ends with the total)
    template <class A> __device__ __forceinline__ static A shfl_xor_struct(unsigned mask, const A& a, int o) {
        static_assert(sizeof(A) % 4 == 0, "accumulator size must be a multiple of 4 bytes");
        union U { A a; uint32_t w[sizeof(A) / 4]; __device__ U() {} } in, out;
        in.a = a;
#pragma unroll
        for (int i = 0; i < (int)(sizeof(A) / 4); ++i) out.w[i] = __shfl_xor_sync(mask, in.w[i], o);
        return out.a;
    }
    template <class A, class M> __device__ __forceinline__ A reduce_acc(A a, M&& merge) const {
        if (GROUP == 1) return a;
        constexpr int W = GROUP < 32 ? GROUP : 32;
        const unsigned mask = GROUP >= 32 ? 0xffffffffu : (((1u << W) - 1u) << ((threadIdx.x & 31u) / W * W));
#pragma unroll
        for (int o = W / 2; o > 0; o >>= 1) { const A other = shfl_xor_struct(mask, a, o); merge(a, other); }
        if (GROUP > 32) {
            __shared__ uint32_t red_acc[(GROUP > 32 ? GROUP / 32 : 1) * (sizeof(A) / 4)];
            union U { A a; uint32_t w[sizeof(A) / 4]; __device__ U() {} } u;
            __syncthreads();
            if ((threadIdx.x & 31) == 0) { u.a = a; for (int i = 0; i < (int)(sizeof(A) / 4); ++i) red_acc[(threadIdx.x >> 5) * (sizeof(A) / 4) + i] = u.w[i]; }
            __syncthreads();
            for (int i = 0; i < (int)(sizeof(A) / 4); ++i) u.w[i] = red_acc[i];
            A r = u.a;
#pragma unroll
            for (int k = 1; k < GROUP / 32; ++k) { for (int i = 0; i < (int)(sizeof(A) / 4); ++i) u.w[i] = red_acc[k * (sizeof(A) / 4) + i]; merge(r, u.a); }
            a = r;
        }
        return a;
    }

    // -- id <-> composite --
    // slot of a remote agent in the ghost segment of its type (binary search in the sorted ghost table)
    __device__ __forceinline__ bool ghost_slot(const AgentView& av, AgentID id, uint32_t& slot) const {
        uint32_t lo = 0, hi = av.nghost;
        while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (av.ghost_ids[mid] < id) lo = mid + 1; else hi = mid; }
        if (lo >= av.nghost || av.ghost_ids[lo] != id) return false;
        slot = av.lcap + lo;
        return true;
    }
    __device__ __forceinline__ bool comp_of(AgentID id, uint32_t& comp) const {
        const uint32_t t = type_nr(id);
        const uint64_t nr = agent_nr(id);
        if (t < 1 || t > ds.n_agent_types || nr < 1) return false;
        if (process_nr(id) != ds.rank) {
            uint32_t s;
            if (!ghost_slot(ds.agents[t], id, s)) return false;
            comp = ds.base[t] + s;
            return true;
        }
        if (nr > ds.agents[t].lcap) return false;
        comp = ds.base[t] + (uint32_t)(nr - 1);
        return true;
    }
    __device__ __forceinline__ AgentID id_of(uint32_t comp) const {
        uint32_t t = 1;
        while (t < ds.n_agent_types && comp >= ds.base[t + 1]) ++t;
        const uint32_t slot = comp - ds.base[t];
        if (slot >= ds.agents[t].lcap) return ds.agents[t].ghost_ids[slot - ds.agents[t].lcap];
        return agent_id(t, ds.rank, (uint64_t)slot + 1);
    }
    // row of target `id` in the read container of edge type e; false = no container entry
    __device__ __forceinline__ bool row_of(const EdgeView& ev, AgentID id, uint32_t& row) const {
        const uint32_t t = type_nr(id);
        const uint64_t nr = agent_nr(id);
        if (t < 1 || t > ds.n_agent_types || nr < 1) { fail(DERR_BAD_ID); return false; }
        if (process_nr(id) != ds.rank) { fail(DERR_REMOTE); return false; }   // edges live on the rank of their target
        if (ev.target) {
            if ((int)t != ev.target) { if (ds.check) fail(DERR_SINGLETYPE_MISMATCH); return false; }
            row = (uint32_t)(nr - 1);
        } else {
            if (nr > ds.agents[t].lcap) return false;
            row = ds.base[t] + (uint32_t)(nr - 1);
        }
        return row < ev.rows;
    }
    __device__ __forceinline__ const EdgeView& rview(int e) const {
        const EdgeView& ev = ds.edges[e];
        if (ds.check && !ev.readable) fail(DERR_EDGE_NOT_READABLE);   // EdgeMethods.jl:310-316,362-368
        return ev;
    }

    // -- implicit raster stencil (KIND_STENCIL) --
    __device__ __forceinline__ bool stencil_cell(const EdgeView& ev, AgentID id, uint32_t& lin) const {
        const RasterView& rv = ds.rasters[ev.st_raster];
        if ((int)type_nr(id) != rv.type || process_nr(id) != ds.rank) return false;
        const uint64_t slot = agent_nr(id) - 1;
        if (slot < ev.st_slot0 || slot - ev.st_slot0 >= rv.ncells) return false;
        lin = (uint32_t)(slot - ev.st_slot0);
        return true;
    }
    // calls fn(source cell linear index) for the entries first, first + step, ... of row `lin`; returns the row length
    template <class Fn> __device__ __forceinline__ uint32_t stencil_row(const EdgeView& ev, uint32_t lin, uint32_t first, uint32_t step, Fn&& fn) const {
        const RasterView& rv = ds.rasters[ev.st_raster];
        const StencilTab& tab = ds.stencils[ev.st_tab];
        int32_t pos[MAX_RASTER_DIMS];
        bool interior = true;
        {
            uint32_t rest = lin;
#pragma unroll
            for (int k = 0; k < MAX_RASTER_DIMS; ++k) {     // unused dimensions have extent 1: pos = 0, never "interior"-relevant
                const uint32_t d = rv.dim32[k];
                if (k < rv.ndims) {
                    pos[k] = (int32_t)(rest % d); rest /= d;
                    interior &= pos[k] >= (int32_t)ev.st_reach && pos[k] + (int32_t)ev.st_reach < (int32_t)d;
                } else pos[k] = 0;
            }
        }
        if (interior) {   // no wrap, no clip: walking the stencil backwards yields ascending source indices
            uint32_t j = 0;
            for (int si = ev.st_n - 1; si >= 0; --si, ++j) {
                if (j % step != first % step || j < first) continue;
                fn((uint32_t)((int32_t)lin - tab.lin[si]));
            }
            return (uint32_t)ev.st_n;
        }
        unsigned long long key[MAX_IMPLICIT_STENCIL];
        uint32_t n = 0;
        for (int si = 0; si < ev.st_n; ++si) {
            uint32_t l = 0; bool ok = true;
#pragma unroll
            for (int k = 0; k < MAX_RASTER_DIMS; ++k) {
                if (k >= rv.ndims) break;
                int32_t v = pos[k] - tab.off[si][k];
                const int32_t d = (int32_t)rv.dim32[k];
                if (v < 0 || v >= d) { if (!ev.st_periodic) { ok = false; break; } v %= d; if (v < 0) v += d; }
                l += (uint32_t)v * rv.stride32[k];
            }
            if (!ok) continue;
            const unsigned long long kk = ((unsigned long long)l << 6) | (unsigned)si;
            uint32_t j = n++;
            while (j > 0 && key[j - 1] > kk) { key[j] = key[j - 1]; --j; }   // insertion sort: (source index, stencil position)
            key[j] = kk;
        }
        for (uint32_t j = first; j < n; j += step) fn((uint32_t)(key[j] >> 6));
        return n;
    }

    // -- read accessors --
    __device__ __forceinline__ long long num_edges(int e, AgentID id) {
        const EdgeView& ev = rview(e);
        if (ev.hints & EDGE_SINGLE_EDGE) { fail(DERR_ACCESSOR_UNAVAILABLE); return 0; }
        if (ev.kind == KIND_STENCIL) {
            uint32_t lin;
            if (!stencil_cell(ev, id, lin)) return 0;
            return (long long)stencil_row(ev, lin, 0xffffffffu, 1, [](uint32_t) {});
        }
        uint32_t row;
        if (!row_of(ev, id, row)) return 0;
        if (ev.kind != KIND_CSR) return ev.cnt[row];
        return (long long)(ev.off[row + 1] - ev.off[row]);
    }
    __device__ __forceinline__ bool has_edge(int e, AgentID id) {
        const EdgeView& ev = rview(e);
        if ((ev.hints & EDGE_SINGLE_EDGE) && ev.target && ev.kind == KIND_CSR) { fail(DERR_ACCESSOR_UNAVAILABLE); return false; }
        if (ev.kind == KIND_STENCIL) return num_edges(e, id) > 0;
        uint32_t row;
        if (!row_of(ev, id, row)) return false;
        if (ev.kind != KIND_CSR) return ev.cnt[row] != 0;
        return ev.off[row + 1] != ev.off[row];
    }
    template <class Fn> __device__ __forceinline__ void for_each_neighbor(int e, AgentID id, Fn&& fn) {
        const EdgeView& ev = rview(e);
        if (ev.hints & EDGE_IGNORE_FROM) { fail(DERR_ACCESSOR_UNAVAILABLE); return; }
        if (ev.kind == KIND_STENCIL) {
            uint32_t lin;
            if (!stencil_cell(ev, id, lin)) return;
            const uint32_t t = (uint32_t)ds.rasters[ev.st_raster].type;
            const uint32_t n = stencil_row(ev, lin, lane_, GROUP, [&](uint32_t src) { fn(agent_id(t, ds.rank, (uint64_t)ev.st_slot0 + src + 1)); });
            if (lane_ == 0) edges_read += n;
            return;
        }
        uint32_t row;
        if (!row_of(ev, id, row)) return;
        const uint32_t b = ev.off[row], en = ev.off[row + 1];
        for (uint32_t k = b + lane_; k < en; k += GROUP) fn(id_of(ev.src[k]));
        if (lane_ == 0) edges_read += en - b;
    }
    __device__ __forceinline__ AgentID neighbor_at(int e, AgentID id, long long k) {
        const EdgeView& ev = rview(e);
        if (ev.hints & EDGE_IGNORE_FROM) { fail(DERR_ACCESSOR_UNAVAILABLE); return 0; }
        if (ev.kind == KIND_STENCIL) {
            uint32_t lin; AgentID out = 0;
            if (stencil_cell(ev, id, lin) && k >= 0) {
                const uint32_t t = (uint32_t)ds.rasters[ev.st_raster].type;
                uint32_t j = 0;
                stencil_row(ev, lin, 0, 1, [&](uint32_t src) { if ((long long)j++ == k) out = agent_id(t, ds.rank, (uint64_t)ev.st_slot0 + src + 1); });
            }
            if (!out) fail(DERR_INDEX);
            return out;
        }
        uint32_t row;
        if (!row_of(ev, id, row) || k < 0 || (uint32_t)k >= ev.off[row + 1] - ev.off[row]) { fail(DERR_INDEX); return 0; }
        return id_of(ev.src[ev.off[row] + (uint32_t)k]);
    }
    template <class S, class Fn> __device__ __forceinline__ void for_each_edge(int e, AgentID id, Fn&& fn) {
        const EdgeView& ev = rview(e);
        if (ev.hints & (EDGE_IGNORE_FROM | EDGE_STATELESS)) { fail(DERR_ACCESSOR_UNAVAILABLE); return; }
        uint32_t row;
        if (!row_of(ev, id, row)) return;
        const uint32_t b = ev.off[row], en = ev.off[row + 1];
        for (uint32_t k = b + lane_; k < en; k += GROUP) fn(id_of(ev.src[k]), soa_load<S>(ev.st, ev.st_cap, k));
        if (lane_ == 0) edges_read += en - b;
    }
    template <class S, class Fn> __device__ __forceinline__ void for_each_edgestate(int e, AgentID id, Fn&& fn) {
        const EdgeView& ev = rview(e);
        if (ev.hints & EDGE_STATELESS) { fail(DERR_ACCESSOR_UNAVAILABLE); return; }
        uint32_t row;
        if (!row_of(ev, id, row)) return;
        const uint32_t b = ev.off[row], en = ev.off[row + 1];
        for (uint32_t k = b + lane_; k < en; k += GROUP) fn(soa_load<S>(ev.st, ev.st_cap, k));
        if (lane_ == 0) edges_read += en - b;
    }
    __device__ __forceinline__ bool agent_slot(int type, AgentID id, uint32_t& s) const {
        const AgentView& av = ds.agents[type];
        if (ds.check) {
            if ((int)type_nr(id) != type) { fail(DERR_AGENT_TYPE_MISMATCH); return false; }   // AgentMethods.jl:92-94
            if (!av.readable) fail(DERR_AGENT_NOT_READABLE);                                   // :96-101
        }
        const uint64_t nr = agent_nr(id);
        if (process_nr(id) != ds.rank) {                                                        // other rank: mirrored ghost state
            if (!ghost_slot(av, id, s)) { fail(DERR_BAD_ID); return false; }
            return true;
        }
        if (nr < 1 || nr > av.nslots_r) { fail(DERR_BAD_ID); return false; }
        s = (uint32_t)(nr - 1);
        if (ds.check && av.died_r && av.died_r[s]) fail(DERR_AGENT_DIED);                      // :106-110
        return true;
    }
    template <class A> __device__ __forceinline__ A agentstate(int type, AgentID id) {
        uint32_t s;
        if (!agent_slot(type, id, s)) return A{};
        const AgentView& av = ds.agents[type];
        return soa_load<A>(av.state_r, av.cap, s);
    }
    template <class Fd> __device__ __forceinline__ Fd agentfield(int type, AgentID id, int off) {
        uint32_t s;
        if (!agent_slot(type, id, s)) return Fd{};
        const AgentView& av = ds.agents[type];
        return soa_load_field<Fd>(av.state_r, av.cap, av.word, s, (uint32_t)off);
    }
    // neighborstates: column -> slot directly (no id round trip): the hot gather of the read phase
    template <class A, class Fn> __device__ __forceinline__ void for_each_neighborstate(int e, int type, AgentID id, Fn&& fn) {
        const EdgeView& ev = rview(e);
        if (ev.hints & EDGE_IGNORE_FROM) { fail(DERR_ACCESSOR_UNAVAILABLE); return; }
        const AgentView& av = ds.agents[type];
        if (ds.check && !av.readable) fail(DERR_AGENT_NOT_READABLE);
        if (ev.kind == KIND_STENCIL) {   // grid stencil: neighbour cells are adjacent slots, the loads of a warp coalesce / hit L1
            uint32_t lin;
            if (!stencil_cell(ev, id, lin)) return;
            if (type != ds.rasters[ev.st_raster].type) { fail(DERR_AGENT_TYPE_MISMATCH); return; }
            const uint8_t* __restrict__ stp = av.state_r;
            const uint32_t capp = av.cap, s0 = ev.st_slot0;
            const uint32_t n = stencil_row(ev, lin, lane_, GROUP, [&](uint32_t src) { fn(soa_load<A>(stp, capp, s0 + src)); });
            if (lane_ == 0) edges_read += n;
            return;
        }
        uint32_t row;
        if (!row_of(ev, id, row)) return;
        const uint32_t b = __ldcs(ev.off + row), en = __ldcs(ev.off + row + 1);   // CSR is streamed once: evict-first
        const uint32_t tb = ds.base[type];
        const uint32_t* __restrict__ src = ev.src;
        const uint8_t* __restrict__ st = av.state_r;
        const uint32_t cap = av.cap, nsl = av.lcap + av.nghost;   // local slots and ghosts are addressed uniformly
        uint32_t k = b + lane_;
        // four independent gathers in flight per lane; the unsigned compare also rejects sources of another type
        for (; k + 3 * GROUP < en; k += 4 * GROUP) {
            const uint32_t s0 = __ldcs(src + k) - tb, s1 = __ldcs(src + k + GROUP) - tb;
            const uint32_t s2 = __ldcs(src + k + 2 * GROUP) - tb, s3 = __ldcs(src + k + 3 * GROUP) - tb;
            if ((s0 >= nsl) | (s1 >= nsl) | (s2 >= nsl) | (s3 >= nsl)) { fail(DERR_AGENT_TYPE_MISMATCH); continue; }
            const A a0 = soa_gather<A>(st, cap, s0);
            const A a1 = soa_gather<A>(st, cap, s1);
            const A a2 = soa_gather<A>(st, cap, s2);
            const A a3 = soa_gather<A>(st, cap, s3);
            fn(a0); fn(a1); fn(a2); fn(a3);
        }
        for (; k < en; k += GROUP) {
            const uint32_t s0 = __ldcs(src + k) - tb;
            if (s0 >= nsl) { fail(DERR_AGENT_TYPE_MISMATCH); continue; }
            fn(soa_gather<A>(st, cap, s0));
        }
        if (lane_ == 0) edges_read += en - b;
    }

    // -- write side --
    template <class S> __device__ __forceinline__ void add_edge_impl(int e, AgentID from, AgentID to, const S* st) {
        if (lane_ != 0) return;
        const int w = F::EdgeWrites::find(e);
        if (w < 0) { fail(DERR_EDGE_NOT_DECLARED); return; }
        const EdgeView& ev = ds.edges[e];
        uint32_t trow, fcomp = 0;
        const uint32_t tt = type_nr(to);
        const uint64_t tnr = agent_nr(to);
        if (ds.nranks > 1) {
            // the edge travels as AgentIDs if its target lives on another rank, or if its source is a remote agent that is not
            // mirrored here yet (the receiver — possibly this rank — registers the ghost before translating)
            const bool away = process_nr(to) != ds.rank ||
                              (!(ev.hints & EDGE_IGNORE_FROM) && process_nr(from) != ds.rank && !comp_of(from, fcomp));
            if (away) {
                if (MODE == MODE_COUNT) { ercnt[w] += 1; return; }
                const uint32_t pos = la.erbase[w] + la.ercount[w][slot] + ercnt[w];
                ercnt[w] += 1;
                ev.rlog_to[pos] = to;
                if (ev.rlog_from) ev.rlog_from[pos] = from;
                ev.rlog_dst[pos] = process_nr(to);
                if (st && ev.rlog_st) soa_store<S>(ev.rlog_st, ev.rlog_cap, pos, *st);
                return;
            }
        }
        if (MODE == MODE_COUNT) { ecnt[w] += 1; return; }
        if (tt < 1 || tt > ds.n_agent_types || tnr < 1) { fail(DERR_BAD_ID); return; }
        if (process_nr(to) != ds.rank) { fail(DERR_REMOTE); return; }
        if (tnr > ds.agents[tt].lcap) { fail(DERR_BAD_ID); return; }
        if (ev.target) {
            if ((int)tt != ev.target) { fail(DERR_SINGLETYPE_MISMATCH); return; }
            trow = (uint32_t)(tnr - 1);
        } else {
            trow = ds.base[tt] + (uint32_t)(tnr - 1);
        }
        if (!(ev.hints & EDGE_IGNORE_FROM) && !comp_of(from, fcomp)) { fail(DERR_BAD_ID); return; }
        if (ev.kind != KIND_CSR && !ev.log_to) {                                               // order-free fast path
            if (ev.kind == KIND_COUNT) atomicAdd(&ev.wcnt[trow], 1u); else ev.wcnt[trow] = 1u;   // count[to] += 1 / true
            ecnt[w] += 1;
            return;
        }
        const uint32_t pos = la.ebase[w] + la.ecount[w][slot] + ecnt[w];
        ecnt[w] += 1;
        ev.log_to[pos] = trow;
        if (ev.log_from) ev.log_from[pos] = fcomp;
        if (st && ev.log_st) soa_store<S>(ev.log_st, ev.log_cap, pos, *st);
    }
    struct NoState {};
    __device__ __forceinline__ void add_edge(int e, AgentID from, AgentID to) { add_edge_impl<NoState>(e, from, to, nullptr); }
    template <class S> __device__ __forceinline__ void add_edge(int e, AgentID from, AgentID to, const S& st) { add_edge_impl<S>(e, from, to, &st); }

    template <class A> __device__ __forceinline__ AgentID add_agent(int type, const A& a) {
        const int w = F::AgentWrites::find(type);
        if (w < 0) { fail(DERR_EDGE_NOT_DECLARED); return 0; }
        if (MODE == MODE_COUNT) { if (lane_ == 0) acnt[w] += 1; return agent_id((uint32_t)type, ds.rank, 1); }
        const AgentView& av = ds.agents[type];
        // _get_next_id (AgentMethods.jl:37-63): j-th birth of the apply takes reuse[end - j], then fresh slots
        const uint32_t j = la.abase[w] + la.acount[w][slot] + acnt[w];
        acnt[w] += 1;
        const uint32_t s = j < av.n_reuse ? av.reuse[av.n_reuse - 1 - j] : av.next0 + (j - av.n_reuse);
        if (lane_ == 0) {
            uint8_t* dst = (av.independent && s < av.nslots_r) ? av.state_r : av.state_w;   // AgentMethods.jl:79-87
            if (sizeof(A) > 0 && av.size) soa_store<A>(dst, av.cap, s, a);
            if (av.died_w) av.died_w[s] = 0;
        }
        return agent_id((uint32_t)type, ds.rank, (uint64_t)s + 1);
    }

    // remove_edges!(sim, to, T) / remove_edges!(sim, from, to, T)  (EdgeMethods.jl:527-599).  A record removes what is in the
    // write container at the moment of the call: the existing entries and the appends that precede it in call order.
    __device__ __forceinline__ void remove_edges_impl(int e, bool has_from, AgentID from, AgentID to) {
        if (lane_ != 0) return;
        const int r = F::EdgeRemoves::find(e);
        if (r < 0) { fail(DERR_EDGE_NOT_DECLARED); return; }
        if (MODE == MODE_COUNT) { rcnt[r] += 1; return; }
        const EdgeView& ev = ds.edges[e];
        if (has_from && (ev.hints & EDGE_IGNORE_FROM)) { fail(DERR_ACCESSOR_UNAVAILABLE); return; }   // :592-598
        const uint32_t pos = la.rbase[r] + la.rcount[r][slot] + rcnt[r];
        rcnt[r] += 1;
        uint32_t row = 0xffffffffu, fcomp = 0xffffffffu;
        const uint32_t tt = type_nr(to);
        const uint64_t tnr = agent_nr(to);
        const bool remote = process_nr(to) != ds.rank;
        if (remote) {   // queued for the target's rank (EdgeMethods.jl:531-532,547-548,559-560): the record stays inert here
            if (!ev.rm_to64) fail(DERR_REMOTE);
        }
        else if (tt >= 1 && tt <= ds.n_agent_types && tnr >= 1 && tnr <= ds.agents[tt].lcap && (!ev.target || (int)tt == ev.target))
            row = (ev.target ? 0u : ds.base[tt]) + (uint32_t)(tnr - 1);
        if (!remote && has_from && !comp_of(from, fcomp)) row = 0xffffffffu;       // an id that names nobody matches no entry
        if (ev.rm_to64) { ev.rm_to64[pos] = remote ? to : 0ull; ev.rm_from64[pos] = (remote && has_from) ? from : 0ull; }
        const int w = F::EdgeWrites::find(e);
        ev.rm_row[pos] = row;
        ev.rm_from[pos] = has_from ? fcomp : 0xffffffffu;
        ev.rm_mark[pos] = w >= 0 ? la.ebase[w] + la.ecount[w][slot] + ecnt[w] : la.rmark[r];
    }
    __device__ __forceinline__ void remove_edges(int e, AgentID to) { remove_edges_impl(e, false, 0, to); }
    __device__ __forceinline__ void remove_edges(int e, AgentID from, AgentID to) { remove_edges_impl(e, true, from, to); }

    // -- raster --
    __device__ __forceinline__ AgentID cellid(int r, const Pos& p) const {
        const RasterView& rv = ds.rasters[r];
        size_t idx = 0, stride = 1;
        for (int k = 0; k < rv.ndims; ++k) {
            if (p.v[k] < 1 || p.v[k] > rv.dims[k]) { fail(DERR_RASTER_POS); return 0; }
            idx += (size_t)(p.v[k] - 1) * stride;
            stride *= (size_t)rv.dims[k];
        }
        return rv.cell_ids ? rv.cell_ids[idx] : id_of(rv.cells[idx]);
    }
    // move_to! for stateless edge types (Raster.jl:437-477); stencil enumeration as _stencil_core :82-96
    __device__ void move_to(int r, AgentID id, const Pos& p, int e_from, int e_to, double distance = 0, int metric = CHEBYSHEV,
                            bool periodic = true, bool only_surrounding = false) {
        const RasterView& rv = ds.rasters[r];
        if (!only_surrounding) {
            const AgentID cell = cellid(r, p);
            if (cell) {
                if (e_from >= 0) add_edge(e_from, cell, id);
                if (e_to >= 0) add_edge(e_to, id, cell);
            }
        }
        if (distance >= 1) {
            const long long d = (long long)floor(distance);
            long long o[MAX_RASTER_DIMS];
            for (int k = 0; k < rv.ndims; ++k) o[k] = -d;
            while (true) {
                bool zero = true;
                double n2 = 0;
                long long n1 = 0;
                for (int k = 0; k < rv.ndims; ++k) { zero &= o[k] == 0; n2 += (double)(o[k] * o[k]); n1 += o[k] < 0 ? -o[k] : o[k]; }
                bool keep = !zero;
                if (keep && metric == EUCLIDEAN) keep = sqrt(n2) <= distance;
                if (keep && metric == MANHATTEN) keep = (double)n1 <= distance;
                if (keep) {
                    Pos q;
                    bool oob = false;
                    for (int k = 0; k < rv.ndims; ++k) {
                        long long v = p.v[k] + o[k];
                        if (v < 1 || v > rv.dims[k]) { oob = true; long long m = (v - 1) % rv.dims[k]; if (m < 0) m += rv.dims[k]; v = m + 1; }
                        q.v[k] = v;
                    }
                    if (!oob || periodic) {
                        const AgentID cell = cellid(r, q);
                        if (e_from >= 0) add_edge(e_from, cell, id);
                        if (e_to >= 0) add_edge(e_to, id, cell);
                    }
                }
                int k = 0;
                while (k < rv.ndims && ++o[k] > d) { o[k] = -d; ++k; }
                if (k == rv.ndims) break;
            }
        }
    }
};

// ---- the transition kernel: the per-agent loop of transition_with(out)_read! + transition_with_write!
//      (src/AgentMethods.jl:159-245) with GROUP lanes per agent -----------------------------------------
template <class F, int MODE, int GROUP>
__global__ void __launch_bounds__(256) transition_kernel(const __grid_constant__ KernelArgs ka) {
    typedef typename F::State State;
    const LaunchArgs& la = ka.la;
    const DeviceSim& ds = ka.ds;
    uint32_t idx, lane;
    if (GROUP == 256) {                                                // block per agent: heavy rows of the binned read phase
        idx = la.rows[blockIdx.x];
        lane = threadIdx.x;
    } else {
        const uint64_t gtid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
        if (gtid / GROUP >= la.n) return;
        idx = (uint32_t)(gtid / GROUP);
        lane = (uint32_t)(gtid % GROUP);
    }
    const AgentView& av = ds.agents[la.type];
    bool skip = av.died_r && av.died_r[idx];                           // jump over died agents (:199-203)
    if (!skip && la.with_edge >= 0) {                                  // with_edge: only targets of that edge type (:209-229)
        const EdgeView& we = ds.edges[la.with_edge];
        const uint32_t row = ds.base[la.type] + idx;
        skip = !(row < we.rows && (we.kind == KIND_CSR ? we.off[row + 1] != we.off[row] : we.cnt[row] != 0));
    }
    if (GROUP < 256 && !skip && la.heavy_min) {                        // heavy rows are handled by the block pass
        const EdgeView& pe = ds.edges[la.primary_edge];
        const uint32_t row = (pe.target ? 0u : ds.base[la.type]) + idx;
        if (row < pe.rows && pe.off[row + 1] - pe.off[row] >= la.heavy_min) return;
    }
    if (skip) {
        if (lane == 0 && MODE == MODE_COUNT) {
#pragma unroll
            for (int i = 0; i < F::EdgeWrites::size; ++i) { la.ecount[i][idx] = 0; if (la.ercount[i]) la.ercount[i][idx] = 0; }
#pragma unroll
            for (int i = 0; i < F::AgentWrites::size; ++i) la.acount[i][idx] = 0;
#pragma unroll
            for (int i = 0; i < F::EdgeRemoves::size; ++i) la.rcount[i][idx] = 0;
        }
        return;
    }
    State self;
    if (la.in_read && av.size) self = soa_load<State>(av.state_r, av.cap, idx);
    else memset(&self, 0, sizeof(State));                              // Val(T) form (:232-245)
    Ctx<F, MODE, GROUP> ctx(ds, la, idx, lane);
    const AgentID id = agent_id((uint32_t)la.type, ds.rank, (uint64_t)idx + 1);
    const bool alive = F()(ctx, self, id);
    if (lane != 0) return;
    if (MODE == MODE_COUNT) {
#pragma unroll
        for (int i = 0; i < F::EdgeWrites::size; ++i) { la.ecount[i][idx] = ctx.ecnt[i]; if (la.ercount[i]) la.ercount[i][idx] = ctx.ercnt[i]; }
#pragma unroll
        for (int i = 0; i < F::AgentWrites::size; ++i) la.acount[i][idx] = ctx.acnt[i];
#pragma unroll
        for (int i = 0; i < F::EdgeRemoves::size; ++i) la.rcount[i][idx] = ctx.rcnt[i];
        return;
    }
    if (la.in_write) {                                                 // transition_with_write! (:159-181)
        if (alive) {   // died_w was set to died_r by the engine before the launch, so a living agent's flag is already 0
            if (av.size) soa_store<State>(av.independent ? av.state_r : av.state_w, av.cap, idx, self);
        } else if (av.immortal) {
            atomicOr(ds.error, (uint32_t)DERR_IMMORTAL_DIED);
        } else {
            av.died_w[idx] = 1;
        }
    }
    // statistics: spread over 1024 counters (32 B apart) so the L2 atomic units never serialise on one address
    if (ctx.edges_read) atomicAdd(la.stats + ((blockIdx.x & 1023u) << 2), ctx.edges_read);
}

template <class F, int MODE, int GROUP>
cudaError_t launch_one(const KernelArgs& ka) {
    const LaunchArgs& la = ka.la;
    const unsigned threads = 256;
    unsigned blocks;
    if (GROUP == 256) blocks = la.n;   // la.n = number of listed rows
    else blocks = (unsigned)(((unsigned long long)la.n * GROUP + threads - 1) / threads);
    if (blocks == 0) return cudaSuccess;
    transition_kernel<F, MODE, GROUP><<<blocks, threads, 0, la.stream>>>(ka);
    return cudaGetLastError();
}

template <class F, bool COOP> struct LaunchSelect;
template <class F> struct LaunchSelect<F, false> {   // thread per agent: exact sequential semantics inside the functor
    static cudaError_t run(const KernelArgs& ka) {
        switch (ka.la.mode) {
            case MODE_COUNT: return launch_one<F, MODE_COUNT, 1>(ka);
            case MODE_EMIT: return launch_one<F, MODE_EMIT, 1>(ka);
            default: return launch_one<F, MODE_DIRECT, 1>(ka);
        }
    }
};
template <class F> struct LaunchSelect<F, true> {    // cooperative: 8 / 32 / 256 lanes per agent
    static cudaError_t run(const KernelArgs& ka) {
        switch (ka.la.mode) {
            case MODE_COUNT: return launch_one<F, MODE_COUNT, 32>(ka);
            case MODE_EMIT: return launch_one<F, MODE_EMIT, 32>(ka);
            default:
                if (ka.la.group == 1) return launch_one<F, MODE_DIRECT, 1>(ka);
                if (ka.la.group == 8) return launch_one<F, MODE_DIRECT, 8>(ka);
                if (ka.la.group == 256) return launch_one<F, MODE_DIRECT, 256>(ka);
                return launch_one<F, MODE_DIRECT, 32>(ka);
        }
    }
};
template <class F>
cudaError_t launch_transition(const LaunchArgs& la) {
    static thread_local KernelArgs ka;
    ka.la = la;
    ka.ds = *la.ds;
    ka.la.ds = nullptr;
    return LaunchSelect<F, F::kCooperative>::run(ka);
}


// ---- source-blocked read phase of a reduce transition ---------------------------------------------------------------------
// Why: a neighbour-state gather over a state array far larger than L2 pays one DRAM row activation per edge (~55 G random
// 8 B gathers/s on B200, profiles/microbench/gather.cu) — 5-10x the time its algorithmic bytes need.  The engine therefore
// keeps a second copy of the CSR columns split by SOURCE BLOCK (blocks of the source-state array that fit the L2 set-aside)
// and sweeps the called rows once per block: all gathers of a sweep hit L2 (~280 G/s), the accumulator (F::Acc, kAccBytes per
// row) is parked in HBM between sweeps and finish() runs in the last one (profiles/microbench/blocked.cu: 37 ms -> 16 ms).
//
// Kernel shape (profiles/microbench/blocked.cu compares five): a warp owns 32 consecutive rows, 2048 threads per SM.  The rows'
// entries inside the block are contiguous, so the warp gathers them edge-parallel (coalesced index loads, independent L2 hits,
// no divergence however skewed the degrees are) into shared memory and every lane folds its own row's part from there.
// Software-pipelined persistent variants (register-, cp.async- and TMA-staged) issue 2-3x the instructions per row and end up
// issue-bound at the same or a worse time.
namespace blk {
// gather with an L2 evict_last policy: the sweep's source block stays resident while the streams pass through
template <int W> __device__ __forceinline__ typename WordT<W>::type ld_keep_word(const void* p, uint64_t) { return *reinterpret_cast<const typename WordT<W>::type*>(p); }
template <> __device__ __forceinline__ uint64_t ld_keep_word<8>(const void* p, uint64_t pol) {
    uint64_t r; asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.b64 %0, [%1], %2;" : "=l"(r) : "l"(p), "l"(pol)); return r;
}
template <> __device__ __forceinline__ uint32_t ld_keep_word<4>(const void* p, uint64_t pol) {
    uint32_t r; asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.b32 %0, [%1], %2;" : "=r"(r) : "l"(p), "l"(pol)); return r;
}
template <class T> __device__ __forceinline__ T gather_keep(const uint8_t* __restrict__ cols, uint32_t stride, uint32_t i, uint64_t pol) {
    constexpr int W = SoaWord<sizeof(T)>::value;
    constexpr int NC = sizeof(T) / W;
    typedef typename WordT<W>::type Word;
    union U { T t; Word w[NC]; __device__ U() {} } u;
#pragma unroll
    for (int c = 0; c < NC; ++c) u.w[c] = ld_keep_word<W>(cols + (size_t)c * stride * W + (size_t)i * W, pol);
    return u.t;
}
// pull a byte range into L2 (one 128 B line per participating thread and round)
__device__ __forceinline__ void prefetch_l2(const void* p, size_t bytes, uint32_t tid, uint32_t nthreads) {
    const char* b = reinterpret_cast<const char*>(p);
    const size_t first = (size_t)(reinterpret_cast<uintptr_t>(b) & 127u);
    for (size_t o = (size_t)tid * 128; o < bytes + first; o += (size_t)nthreads * 128)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(b - first + o));
}
// streaming (evict-first) column access for the per-row streams
template <class T> __device__ __forceinline__ T soa_load_cs(const uint8_t* __restrict__ cols, uint32_t stride, uint32_t i) {
    constexpr int W = SoaWord<sizeof(T)>::value;
    constexpr int NC = sizeof(T) / W;
    typedef typename WordT<W>::type Word;
    union U { T t; Word w[NC]; __device__ U() {} } u;
#pragma unroll
    for (int c = 0; c < NC; ++c) u.w[c] = __ldcs(reinterpret_cast<const Word*>(cols + (size_t)c * stride * W) + i);
    return u.t;
}
}  // namespace blk

#ifndef VB_BLK_MINCTAS
#define VB_BLK_MINCTAS 6
#endif
template <class F> struct BlockedCfg {
    typedef typename F::State State;
    typedef typename F::Source Source;
    typedef typename F::Acc Acc;
    static constexpr int CHK = 128;         // entries a warp gathers per round (CHK / 32 independent L2 hits per lane)
    // the accumulator is parked as word columns: an 8-byte column for every full 8 bytes, then 4-byte columns
    static_assert(F::kAccBytes % 4 == 0 && F::kAccBytes <= (int)sizeof(Acc), "kAccBytes: a multiple of 4, at most sizeof(Acc)");
    static_assert(sizeof(Acc) % 4 == 0 && sizeof(State) % 4 == 0, "state and accumulator sizes must be multiples of 4 bytes");
    static constexpr int A8 = F::kAccBytes / 8, A4 = (F::kAccBytes % 8) / 4;
    static __device__ __forceinline__ void acc_load(const uint8_t* __restrict__ base, uint32_t stride, uint32_t i, Acc& a) {
        union U { Acc t; uint32_t w[sizeof(Acc) / 4]; __device__ U() {} } u;
#pragma unroll
        for (int c = 0; c < (int)(sizeof(Acc) / 4); ++c) u.w[c] = 0;
#pragma unroll
        for (int c = 0; c < A8; ++c) {
            const uint64_t v = __ldcs(reinterpret_cast<const uint64_t*>(base + (size_t)c * stride * 8) + i);
            u.w[2 * c] = (uint32_t)v; u.w[2 * c + 1] = (uint32_t)(v >> 32);
        }
#pragma unroll
        for (int c = 0; c < A4; ++c) u.w[2 * A8 + c] = __ldcs(reinterpret_cast<const uint32_t*>(base + (size_t)A8 * stride * 8 + (size_t)c * stride * 4) + i);
        a = u.t;
    }
    static __device__ __forceinline__ void acc_store(uint8_t* __restrict__ base, uint32_t stride, uint32_t i, const Acc& a) {
        union U { Acc t; uint32_t w[sizeof(Acc) / 4]; __device__ U() {} } u;
        u.t = a;
#pragma unroll
        for (int c = 0; c < A8; ++c) __stcs(reinterpret_cast<uint64_t*>(base + (size_t)c * stride * 8) + i, (uint64_t)u.w[2 * c] | ((uint64_t)u.w[2 * c + 1] << 32));
#pragma unroll
        for (int c = 0; c < A4; ++c) __stcs(reinterpret_cast<uint32_t*>(base + (size_t)A8 * stride * 8 + (size_t)c * stride * 4) + i, u.w[2 * A8 + c]);
    }
};

template <class F, bool FIRST, bool LAST>
__global__ void __launch_bounds__(256, VB_BLK_MINCTAS) reduce_blocked_kernel(const __grid_constant__ KernelArgs ka) {
    typedef BlockedCfg<F> C;
    typedef typename C::State State;
    typedef typename C::Source Source;
    typedef typename C::Acc Acc;
    constexpr int CHK = C::CHK;
    __shared__ __align__(16) uint8_t val_raw[8][CHK * sizeof(Source)];
    const LaunchArgs& la = ka.la;
    const DeviceSim& ds = ka.ds;
    const uint32_t lane = threadIdx.x & 31;
    Source* val = reinterpret_cast<Source*>(val_raw[threadIdx.x >> 5]);
    const uint64_t gtid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool listed = !FIRST && !LAST && la.blk_rows != nullptr;         // sweep only the rows that own an entry in this block
    const uint32_t nwork = listed ? la.blk_nrows : la.n;
    const uint32_t idx = listed ? (gtid < nwork ? __ldcs(la.blk_rows + gtid) : la.n) : (uint32_t)gtid;
    const AgentView& av = ds.agents[la.type];
    const AgentView& sv = ds.agents[F::kSourceType];
    const uint8_t* __restrict__ src_st = sv.state_r;
    const uint32_t src_cap = sv.cap;
    const uint32_t* __restrict__ gsrc = la.blk_src;
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    // Latency: a row's work is the dependent chain offsets -> indices -> gather, two DRAM round trips before the first L2 hit.
    // Every CTA therefore pulls the streams of the CTA that will run `blk_ahead` CTAs later (about one residency wave) into L2:
    // its offsets, own states and parked accumulators now, its index range at the end (once the two bounding offsets arrived).
    const uint32_t pc = blockIdx.x + la.blk_ahead;
    uint32_t pa = 0, pb = 0;
    const bool ahead = la.blk_ahead && pc < gridDim.x;
    if (ahead && threadIdx.x < 64) {                                       // two warps issue the prefetches (one line per thread and round)
        const uint32_t r0 = pc * 256u, nr = nwork - r0 < 256u ? nwork - r0 : 256u;
        const uint32_t* __restrict__ offs = listed ? la.blk_roff : la.blk_off;
        if (threadIdx.x == 0) { pa = __ldg(offs + r0); pb = __ldg(offs + r0 + nr); }
        blk::prefetch_l2(offs + r0, (size_t)(nr + 1) * 4, threadIdx.x, 64);
        if (listed) blk::prefetch_l2(la.blk_rows + r0, (size_t)nr * 4, threadIdx.x, 64);   // (own states / accumulators of listed rows are scattered)
        else {
            constexpr int SW = SoaWord<sizeof(State)>::value, SC = sizeof(State) / SW;
#pragma unroll
            for (int c = 0; c < SC; ++c) blk::prefetch_l2(av.state_r + ((size_t)c * av.cap + r0) * SW, (size_t)nr * SW, threadIdx.x, 64);
            if (!FIRST) {
#pragma unroll
                for (int c = 0; c < C::A8; ++c) blk::prefetch_l2(la.blk_acc + ((size_t)c * la.blk_stride + r0) * 8, (size_t)nr * 8, threadIdx.x, 64);
#pragma unroll
                for (int c = 0; c < C::A4; ++c) blk::prefetch_l2(la.blk_acc + (size_t)C::A8 * la.blk_stride * 8 + ((size_t)c * la.blk_stride + r0) * 4, (size_t)nr * 4, threadIdx.x, 64);
            }
        }
    }
    // no early exit: the 32 rows of a warp are walked together.  Their entries are contiguous: [lo of lane 0, hi of lane 31)
    // (with a row list the listed rows' entries are still contiguous: the rows in between own none)
    uint32_t lo, hi;
    if (listed) {
        const uint32_t i = gtid < nwork ? (uint32_t)gtid : nwork;
        lo = __ldcs(la.blk_roff + i); hi = gtid < nwork ? __ldcs(la.blk_roff + i + 1) : lo;
    } else {
        const uint32_t row = gtid < nwork ? idx : la.n;                    // rows past the end: an empty range at the very end
        lo = __ldcs(la.blk_off + row); hi = gtid < nwork ? __ldcs(la.blk_off + row + 1) : lo;
    }
    const bool act = gtid < nwork && !(av.died_r && av.died_r[idx]);       // jump over died agents (AgentMethods.jl:199-203)
    const F f{};
    Ctx<F, MODE_DIRECT, 1> ctx(ds, la, idx, 0);
    State self;
    Acc acc;
    if (act) {
        self = blk::soa_load_cs<State>(av.state_r, av.cap, idx);
        if (FIRST) f.init(ctx, self, acc); else C::acc_load(la.blk_acc, la.blk_stride, idx, acc);
    } else {
        memset(&self, 0, sizeof(State));
        memset(&acc, 0, sizeof(Acc));
    }
    const uint32_t e0 = __shfl_sync(0xffffffffu, lo, 0), e1 = __shfl_sync(0xffffffffu, hi, 31);
    const uint32_t len = hi - lo;
    if (!act) hi = lo;                                                    // a died agent's entries are gathered but never folded
    // the warp's entries in chunks of CHK: edge-parallel (coalesced index loads, CHK / 32 independent L2 hits per lane) into
    // shared memory, then every lane folds the part of its own row that lies in the chunk
    for (uint32_t base = e0; base < e1; base += CHK) {
        Source v[CHK / 32];
        uint32_t six[CHK / 32];
#pragma unroll
        for (int j = 0; j < CHK / 32; ++j) { const uint32_t x = base + lane + 32 * j; if (x < e1) six[j] = __ldcs(gsrc + x); }
#pragma unroll
        for (int j = 0; j < CHK / 32; ++j) { const uint32_t x = base + lane + 32 * j; if (x < e1) v[j] = blk::gather_keep<Source>(src_st, src_cap, six[j], pol); }
#pragma unroll
        for (int j = 0; j < CHK / 32; ++j) { const uint32_t x = base + lane + 32 * j; if (x < e1) val[lane + 32 * j] = v[j]; }
        __syncwarp();
        const uint32_t b = lo > base ? lo : base, e = hi < base + CHK ? hi : base + CHK;
        for (uint32_t x = b; x < e; ++x) f.fold(ctx, self, val[x - base], acc);
        __syncwarp();
    }
    if (threadIdx.x < 32 && ahead) {                                       // warp 0: the index range of the CTA `blk_ahead` later
        pa = __shfl_sync(0xffffffffu, pa, 0); pb = __shfl_sync(0xffffffffu, pb, 0);
        blk::prefetch_l2(gsrc + pa, (size_t)(pb - pa) * 4, lane, 32);
    }
    {
        const unsigned total = __reduce_add_sync(0xffffffffu, act ? len : 0u);
        if (lane == 0 && total) atomicAdd(la.stats + ((blockIdx.x & 1023u) << 2), (unsigned long long)total);
    }
    if (!act) return;
    if (!LAST) C::acc_store(la.blk_acc, la.blk_stride, idx, acc);
    else if (!(la.blk_heavy && ((la.blk_heavy[idx >> 5] >> (idx & 31)) & 1u))) {   // heavy rows: done by the block-per-agent pass
        const AgentID id = agent_id((uint32_t)la.type, ds.rank, (uint64_t)idx + 1);
        const bool alive = f.finish(ctx, self, id, acc);
        if (la.in_write) {                                                 // transition_with_write! (AgentMethods.jl:159-181)
            if (alive) { if (av.size) soa_store<State>(av.independent ? av.state_r : av.state_w, av.cap, idx, self); }
            else if (av.immortal) atomicOr(ds.error, (uint32_t)DERR_IMMORTAL_DIED);
            else av.died_w[idx] = 1;
        }
    }
}


// ---- prefiltered sweep (F::kPrefilter, include/vahana_model.h) ----------------------------------------------------------------------
// Same walk as reduce_blocked_kernel — a warp owns 32 consecutive (listed) rows and takes their contiguous entries edge-parallel in
// chunks — but what is gathered for every entry is the one-byte KEY of the source (a column of 1 B per slot: the whole source type or
// half of it fits the L2 set-aside where the 8 B states need eleven blocks).  Entries whose key may_accept() are queued in entry order
// (ballot compaction; a shared counter per row); a flush fetches the exact states of the queue edge-parallel and every lane folds
// the segment of its own row in entry order.  Everything else (parked accumulators, row lists, look-ahead prefetch, finish and the
// write rules) is as in the unfiltered sweep.  Shape from profiles/microbench/prefilter.cu: chunk 64, 8 resident CTAs per SM.
namespace blk {
__device__ __forceinline__ uint32_t ld_key(const uint8_t* p, uint64_t pol) {
    uint32_t r; asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.u8 %0, [%1], %2;" : "=r"(r) : "l"(p), "l"(pol)); return r;
}
}  // namespace blk
template <class F, int RF = 0> struct PrefilterStage {
    static constexpr int CHK = 64, QCAP = CHK + 32;
    uint32_t off[33];                   // entry positions of the warp's 32 rows (+ end)
    uint32_t qcnt[32];                  // queued entries per row
    typename F::Probe probe[32];
    uint32_t qidx[QCAP];                // queued source slots, in entry order
    typename F::Source qval[QCAP];      // their exact states
};
template <class F> struct PrefilterStage<F, 1> : PrefilterStage<F, 0> {
    uint32_t nz[32];                    // RF = 1: the warp's non-empty rows, in order
};
// WPC = warps per CTA.  A CTA lives as long as its slowest warp, and a warp that meets a long power-law row is slow: small CTAs leave
// fewer warps idle (profiles/r1_prefilter/prefilter_shapes2.txt: 11.65 / 10.86 / 10.59 ms with 8 / 4 / 2 warps per CTA).
// RF (row find) = 0: the row that owns an entry by a 5-step search of the 33 offsets in shared memory; RF = 1 (VB_PF_ROWFIND=1,
// experimental: written after round 1's GPU budget was spent, index arithmetic checked on the CPU, not yet run on a GPU): one REDUX
// per 32 entries marks where non-empty rows start, one ballot counts the rows started before, and an entry's row is
// nz[started_before + popc(starts & lanes_le) - 1] — one shared-memory load per entry instead of five dependent ones.
template <class F, bool FIRST, bool LAST, int WPC, int RF = 0>
__global__ void __launch_bounds__(32 * WPC, 64 / WPC) reduce_prefilter_kernel(const __grid_constant__ KernelArgs ka) {
    constexpr uint32_t TPB = 32 * WPC, PFT = TPB < 64 ? TPB : 64;   // rows per CTA; threads that issue the look-ahead prefetches
    typedef BlockedCfg<F> C;
    typedef typename C::State State;
    typedef typename C::Source Source;
    typedef typename C::Acc Acc;
    typedef PrefilterStage<F, RF> Stage;
    constexpr int CHK = Stage::CHK, QCAP = Stage::QCAP, U = CHK / 32;
    __shared__ Stage stages[WPC];
    const LaunchArgs& la = ka.la;
    const DeviceSim& ds = ka.ds;
    const uint32_t lane = threadIdx.x & 31;
    Stage& sm = stages[threadIdx.x >> 5];
    const uint64_t gtid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool listed = !FIRST && !LAST && la.blk_rows != nullptr;
    const uint32_t nwork = listed ? la.blk_nrows : la.n;
    const uint32_t idx = listed ? (gtid < nwork ? __ldcs(la.blk_rows + gtid) : la.n) : (uint32_t)gtid;
    const AgentView& av = ds.agents[la.type];
    const AgentView& sv = ds.agents[F::kSourceType];
    const uint8_t* __restrict__ src_st = sv.state_r;
    const uint32_t src_cap = sv.cap;
    const uint32_t* __restrict__ gsrc = la.blk_src;
    const uint8_t* __restrict__ keys = la.blk_key;
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    // look-ahead prefetch of the streams of the CTA one residency wave later (see reduce_blocked_kernel)
    const uint32_t pc = blockIdx.x + la.blk_ahead;
    uint32_t pa = 0, pb = 0;
    const bool ahead = la.blk_ahead && pc < gridDim.x;
    if (ahead && threadIdx.x < PFT) {
        const uint32_t r0 = pc * TPB, nr = nwork - r0 < TPB ? nwork - r0 : TPB;
        const uint32_t* __restrict__ offs = listed ? la.blk_roff : la.blk_off;
        if (threadIdx.x == 0) { pa = __ldg(offs + r0); pb = __ldg(offs + r0 + nr); }
        blk::prefetch_l2(offs + r0, (size_t)(nr + 1) * 4, threadIdx.x, PFT);
        if (listed) blk::prefetch_l2(la.blk_rows + r0, (size_t)nr * 4, threadIdx.x, PFT);
        else {
            constexpr int SW = SoaWord<sizeof(State)>::value, SC = sizeof(State) / SW;
#pragma unroll
            for (int c = 0; c < SC; ++c) blk::prefetch_l2(av.state_r + ((size_t)c * av.cap + r0) * SW, (size_t)nr * SW, threadIdx.x, PFT);
            if (!FIRST) {
#pragma unroll
                for (int c = 0; c < C::A8; ++c) blk::prefetch_l2(la.blk_acc + ((size_t)c * la.blk_stride + r0) * 8, (size_t)nr * 8, threadIdx.x, PFT);
#pragma unroll
                for (int c = 0; c < C::A4; ++c) blk::prefetch_l2(la.blk_acc + (size_t)C::A8 * la.blk_stride * 8 + ((size_t)c * la.blk_stride + r0) * 4, (size_t)nr * 4, threadIdx.x, PFT);
            }
        }
    }
    uint32_t lo, hi;
    if (listed) {
        const uint32_t i = gtid < nwork ? (uint32_t)gtid : nwork;
        lo = __ldcs(la.blk_roff + i); hi = gtid < nwork ? __ldcs(la.blk_roff + i + 1) : lo;
    } else {
        const uint32_t row = gtid < nwork ? idx : la.n;
        lo = __ldcs(la.blk_off + row); hi = gtid < nwork ? __ldcs(la.blk_off + row + 1) : lo;
    }
    const bool act = gtid < nwork && !(av.died_r && av.died_r[idx]);       // jump over died agents (AgentMethods.jl:199-203)
    const F f{};
    Ctx<F, MODE_DIRECT, 1> ctx(ds, la, idx, 0);
    State self;
    Acc acc;
    if (act) {
        self = blk::soa_load_cs<State>(av.state_r, av.cap, idx);
        if (FIRST) f.init(ctx, self, acc); else C::acc_load(la.blk_acc, la.blk_stride, idx, acc);
        sm.probe[lane] = f.probe(ctx, self);
    } else {
        memset(&self, 0, sizeof(State));
        memset(&acc, 0, sizeof(Acc));
    }
    sm.off[lane] = lo;
    if (lane == 31) sm.off[32] = hi;
    sm.qcnt[lane] = 0;
    const uint32_t alive = __ballot_sync(0xffffffffu, act);               // entries of died rows are never queued
    if constexpr (RF == 1) {
        const uint32_t nzmask = __ballot_sync(0xffffffffu, hi > lo);
        if (hi > lo) sm.nz[__popc(nzmask & ((1u << lane) - 1u))] = lane;
    }
    __syncwarp();
    const uint32_t e0 = sm.off[0], e1 = sm.off[32];
    const uint32_t len = hi - lo;
    const uint32_t lt = (1u << lane) - 1u;
    uint32_t qn = 0;
    for (uint32_t base = e0; base < e1; base += CHK) {
        uint32_t six[U], ks[U];
#pragma unroll
        for (int u = 0; u < U; ++u) { const uint32_t x = base + lane + 32 * u; six[u] = x < e1 ? __ldcs(gsrc + x) : 0u; }
#pragma unroll
        for (int u = 0; u < U; ++u) { const uint32_t x = base + lane + 32 * u; ks[u] = x < e1 ? blk::ld_key(keys + six[u], pol) : 0u; }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t x = base + lane + 32 * u;
            uint32_t r = 0;                                                // the row that owns entry x: the last one starting at or before it
            if constexpr (RF == 0) {
#pragma unroll
                for (int st = 16; st; st >>= 1) if (sm.off[r + st] <= x) r += st;
            } else {
                const bool nonempty = hi > lo;
                const uint32_t x0 = base + 32 * u, d = lo - x0;
                const uint32_t before = __popc(__ballot_sync(0xffffffffu, nonempty && lo < x0));
                const uint32_t starts = __reduce_or_sync(0xffffffffu, (nonempty && lo >= x0 && d < 32u) ? (1u << d) : 0u);
                r = sm.nz[(before + __popc(starts & (lt | (1u << lane))) - 1u) & 31u];
            }
            const bool pass = x < e1 && ((alive >> r) & 1u) && f.may_accept(sm.probe[r], ks[u]);
            const uint32_t m = __ballot_sync(0xffffffffu, pass);
            if (pass) { sm.qidx[qn + __popc(m & lt)] = six[u]; atomicAdd(&sm.qcnt[r], 1u); }
            qn += __popc(m);
        }
        __syncwarp();
        if (qn > (uint32_t)(QCAP - CHK) || base + CHK >= e1) {             // flush: exact states edge-parallel, then the per-row folds
            for (uint32_t i = lane; i < qn; i += 32) sm.qval[i] = soa_gather<Source>(src_st, src_cap, sm.qidx[i]);
            __syncwarp();
            const uint32_t mine = sm.qcnt[lane];
            uint32_t incl = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= (uint32_t)o) incl += t; }
            const uint32_t start = incl - mine;
            for (uint32_t j = 0; j < mine; ++j) f.fold(ctx, self, sm.qval[start + j], acc);
            sm.qcnt[lane] = 0; qn = 0;
            __syncwarp();
        }
    }
    if (threadIdx.x < 32 && ahead) {
        pa = __shfl_sync(0xffffffffu, pa, 0); pb = __shfl_sync(0xffffffffu, pb, 0);
        blk::prefetch_l2(gsrc + pa, (size_t)(pb - pa) * 4, lane, 32);
    }
    {
        const unsigned total = __reduce_add_sync(0xffffffffu, act ? len : 0u);
        if (lane == 0 && total) atomicAdd(la.stats + ((blockIdx.x & 1023u) << 2), (unsigned long long)total);
    }
    if (!act) return;
    if (!LAST) C::acc_store(la.blk_acc, la.blk_stride, idx, acc);
    else if (!(la.blk_heavy && ((la.blk_heavy[idx >> 5] >> (idx & 31)) & 1u))) {   // heavy rows: done by the block-per-agent pass
        const AgentID id = agent_id((uint32_t)la.type, ds.rank, (uint64_t)idx + 1);
        const bool alive_after = f.finish(ctx, self, id, acc);
        if (la.in_write) {                                                 // transition_with_write! (AgentMethods.jl:159-181)
            if (alive_after) { if (av.size) soa_store<State>(av.independent ? av.state_r : av.state_w, av.cap, idx, self); }
            else if (av.immortal) atomicOr(ds.error, (uint32_t)DERR_IMMORTAL_DIED);
            else av.died_w[idx] = 1;
        }
    }
}
// ---- prefiltered sweep over the SEGMENTED view --------------------------------------------------------------------------------------
// Second shape of the prefiltered sweep (profiles/microbench/twophase.cu, profiles/r2_microbench/): the blocked view is built over
// SEGMENTS — a row of more than SEG_LEN entries is cut into segments of SEG_LEN — and every entry carries, above its 27-bit source slot
// (relative to the block), the index of its segment inside the warp's group of 32.  Consequences:
//   * hub rows need no pass of their own: no warp ever walks more than 32 x SEG_LEN entries, the segments' partial accumulators are
//     merged (in segment order) and finished by reduce_hubmerge_kernel;
//   * the row that owns an entry is a shift, its probe comes from the owner lane by shuffle: no offset table in shared memory, no
//     search, no per-row counters; at a flush a lane finds the run of its segment by a binary search over the ascending tags;
//   * U = 4 key gathers per lane and round in flight.
// The queue keeps entry order, so a segment's candidates are folded in entry order; fold() decides on the exact state.
enum : uint32_t { SEG_MULTI = 0x80000000u, SEG_SRC_MASK = 0x07ffffffu, SEG_TAG_SHIFT = 27 };
template <class T> __device__ __forceinline__ T shfl_struct(const T& v, uint32_t src) {
    static_assert(sizeof(T) % 4 == 0, "shuffled structs are whole words");
    union U { T t; uint32_t w[sizeof(T) / 4]; __device__ U() {} } a, b;
    a.t = v;
#pragma unroll
    for (int i = 0; i < (int)(sizeof(T) / 4); ++i) b.w[i] = __shfl_sync(0xffffffffu, a.w[i], src);
    return b.t;
}
template <class F, int U, int QX> struct SegStage {
    static constexpr int QCAP = 32 * U + QX;
    uint32_t qe[QCAP];                  // queued entries (tag | source slot in the block), in entry order
    typename F::Source qv[QCAP];        // their exact states
};
#ifndef VB_SEG_U
#define VB_SEG_U 4
#endif
#ifndef VB_SEG_QX
#define VB_SEG_QX 64
#endif
template <class F, bool FIRST, bool LAST, int WPC, int MINB>
__global__ void __launch_bounds__(32 * WPC, MINB) reduce_segsweep_kernel(const __grid_constant__ KernelArgs ka) {
    typedef BlockedCfg<F> C;
    typedef typename C::State State;
    typedef typename C::Source Source;
    typedef typename C::Acc Acc;
    typedef typename F::Probe Probe;
    constexpr int U = VB_SEG_U, QX = VB_SEG_QX;
    typedef SegStage<F, U, QX> Stage;
    __shared__ Stage stages[WPC];
    const LaunchArgs& la = ka.la;
    const DeviceSim& ds = ka.ds;
    const uint32_t lane = threadIdx.x & 31;
    Stage& sm = stages[threadIdx.x >> 5];
    const uint64_t g = (uint64_t)blockIdx.x * WPC + (threadIdx.x >> 5);
    const uint64_t s0 = g * 32;
    if (s0 >= la.blk_nseg) return;                                        // whole warp
    const uint32_t ns = la.blk_nseg - s0 < 32 ? (uint32_t)(la.blk_nseg - s0) : 32u;
    const AgentView& av = ds.agents[la.type];
    const AgentView& sv = ds.agents[F::kSourceType];
    const uint8_t* __restrict__ src_st = sv.state_r;
    const uint32_t src_cap = sv.cap;
    const uint32_t* __restrict__ gsrc = la.blk_src;
    const uint8_t* __restrict__ keys = la.blk_key + la.blk_base;
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    // look-ahead: pull the offsets / segment rows of the CTA one residency wave later into L2 now, its entry range at the end
    const uint32_t pc = blockIdx.x + la.blk_ahead;
    uint32_t pa = 0, pb = 0;
    const bool ahead = la.blk_ahead && pc < gridDim.x;
    if (ahead && threadIdx.x < 32) {
        const uint64_t r0 = (uint64_t)pc * 32 * WPC;
        const uint32_t nr = la.blk_nseg - r0 < 32u * WPC ? (uint32_t)(la.blk_nseg - r0) : 32u * WPC;
        if (lane == 0) { pa = __ldg(la.blk_off + r0); pb = __ldg(la.blk_off + r0 + nr); }
        blk::prefetch_l2(la.blk_off + r0, (size_t)(nr + 1) * 4, lane, 32);
        blk::prefetch_l2(la.blk_seg_row + r0, (size_t)nr * 4, lane, 32);
        if (!FIRST) {
#pragma unroll
            for (int c = 0; c < C::A8; ++c) blk::prefetch_l2(la.blk_acc + ((size_t)c * la.blk_stride + r0) * 8, (size_t)nr * 8, lane, 32);
#pragma unroll
            for (int c = 0; c < C::A4; ++c) blk::prefetch_l2(la.blk_acc + (size_t)C::A8 * la.blk_stride * 8 + ((size_t)c * la.blk_stride + r0) * 4, (size_t)nr * 4, lane, 32);
        }
    }
    const uint32_t e0 = __ldcs(la.blk_off + s0), e1 = __ldcs(la.blk_off + s0 + ns);
    const uint32_t rowm = lane < ns ? __ldcs(la.blk_seg_row + s0 + lane) : 0u;
    const uint32_t idx = rowm & ~SEG_MULTI;
    const bool act = lane < ns && !(av.died_r && av.died_r[idx]);        // jump over died agents (AgentMethods.jl:199-203)
    const F f{};
    Ctx<F, MODE_DIRECT, 1> ctx(ds, la, idx, 0);
    State self;
    Acc acc;
    Probe probe;
    if (act) {
        self = blk::soa_load_cs<State>(av.state_r, av.cap, idx);
        if (FIRST) f.init(ctx, self, acc); else C::acc_load(la.blk_acc, la.blk_stride, (uint32_t)s0 + lane, acc);
        probe = f.probe(ctx, self);
    } else {
        memset(&self, 0, sizeof(State));
        memset(&acc, 0, sizeof(Acc));
        memset(&probe, 0, sizeof(Probe));
    }
    const uint32_t alive = __ballot_sync(0xffffffffu, act);              // entries of died rows are never queued
    const uint32_t lt = (1u << lane) - 1u;
    uint32_t qn = 0;
    for (uint32_t base = e0; base < e1; base += 32 * U) {
        uint32_t ent[U], ks[U];
#pragma unroll
        for (int u = 0; u < U; ++u) { const uint32_t x = base + lane + 32 * u; ent[u] = x < e1 ? __ldcs(gsrc + x) : 0xffffffffu; }
#pragma unroll
        for (int u = 0; u < U; ++u) ks[u] = ent[u] != 0xffffffffu ? blk::ld_key(keys + (ent[u] & SEG_SRC_MASK), pol) : 0u;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t tag = ent[u] >> SEG_TAG_SHIFT;
            const Probe pr = shfl_struct(probe, tag);
            const bool pass = ent[u] != 0xffffffffu && ((alive >> tag) & 1u) && f.may_accept(pr, ks[u]);
            const uint32_t m = __ballot_sync(0xffffffffu, pass);
            if (pass) sm.qe[qn + __popc(m & lt)] = ent[u];
            qn += __popc(m);
        }
        __syncwarp();
        if (qn > (uint32_t)QX || base + 32 * U >= e1) {                   // flush: exact states edge-parallel, then the per-segment folds
            for (uint32_t i = lane; i < qn; i += 32) sm.qv[i] = soa_gather<Source>(src_st, src_cap, la.blk_base + (sm.qe[i] & SEG_SRC_MASK));
            uint32_t lo = 0, hi = qn;                                     // first queued candidate whose tag is >= my lane
            while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if ((sm.qe[mid] >> SEG_TAG_SHIFT) < lane) lo = mid + 1; else hi = mid; }
            uint32_t end = __shfl_down_sync(0xffffffffu, lo, 1);
            if (lane == 31) end = qn;
            __syncwarp();
            for (uint32_t j = lo; j < end; ++j) f.fold(ctx, self, sm.qv[j], acc);
            qn = 0;
            __syncwarp();
        }
    }
    if (threadIdx.x < 32 && ahead) {
        pa = __shfl_sync(0xffffffffu, pa, 0); pb = __shfl_sync(0xffffffffu, pb, 0);
        blk::prefetch_l2(gsrc + pa, (size_t)(pb - pa) * 4, lane, 32);
    }
    if (lane == 0 && e1 > e0) atomicAdd(la.stats + ((blockIdx.x & 1023u) << 2), (unsigned long long)(e1 - e0));
    if (!act) return;
    if (!LAST || (rowm & SEG_MULTI)) C::acc_store(la.blk_acc, la.blk_stride, (uint32_t)s0 + lane, acc);      // hub rows: finished by reduce_hubmerge_kernel
    else {
        const AgentID id = agent_id((uint32_t)la.type, ds.rank, (uint64_t)idx + 1);
        const bool alive_after = f.finish(ctx, self, id, acc);
        if (la.in_write) {                                                 // transition_with_write! (AgentMethods.jl:159-181)
            if (alive_after) { if (av.size) soa_store<State>(av.independent ? av.state_r : av.state_w, av.cap, idx, self); }
            else if (av.immortal) atomicOr(ds.error, (uint32_t)DERR_IMMORTAL_DIED);
            else av.died_w[idx] = 1;
        }
    }
}
// rows cut into several segments: merge the parked segment accumulators in segment order, then finish() and the write rules
template <class F>
__global__ void __launch_bounds__(128) reduce_hubmerge_kernel(const __grid_constant__ KernelArgs ka) {
    typedef BlockedCfg<F> C;
    typedef typename C::State State;
    typedef typename C::Acc Acc;
    const LaunchArgs& la = ka.la;
    const DeviceSim& ds = ka.ds;
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= la.blk_nhub) return;
    const uint32_t idx = la.blk_hub_rows[t];
    const AgentView& av = ds.agents[la.type];
    if (av.died_r && av.died_r[idx]) return;
    const F f{};
    Ctx<F, MODE_DIRECT, 1> ctx(ds, la, idx, 0);
    State self = soa_load<State>(av.state_r, av.cap, idx);
    Acc acc, part;
    const uint32_t k0 = la.blk_hub_seg[2 * t], k1 = la.blk_hub_seg[2 * t + 1];
    C::acc_load(la.blk_acc, la.blk_stride, k0, acc);
    for (uint32_t k = k0 + 1; k < k1; ++k) { C::acc_load(la.blk_acc, la.blk_stride, k, part); f.merge(acc, part); }
    const AgentID id = agent_id((uint32_t)la.type, ds.rank, (uint64_t)idx + 1);
    const bool alive_after = f.finish(ctx, self, id, acc);
    if (la.in_write) {
        if (alive_after) { if (av.size) soa_store<State>(av.independent ? av.state_r : av.state_w, av.cap, idx, self); }
        else if (av.immortal) atomicOr(ds.error, (uint32_t)DERR_IMMORTAL_DIED);
        else av.died_w[idx] = 1;
    }
}
// key column of the source type: four slots per thread, one packed store
template <class F>
__global__ void __launch_bounds__(256) build_keys_kernel(const __grid_constant__ KernelArgs ka) {
    typedef typename F::Source Source;
    const LaunchArgs& la = ka.la;
    const DeviceSim& ds = ka.ds;
    const AgentView& sv = ds.agents[F::kSourceType];
    // slots [first, end): four per thread on 4-aligned groups (one packed store), single bytes at a ragged head or tail
    const uint64_t first = la.blk_key_first, end = first + la.blk_nkeys;
    const uint64_t i0 = (first & ~3ull) + ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i0 >= end) return;
    const F f{};
    Ctx<F, MODE_DIRECT, 1> ctx(ds, la, (uint32_t)i0, 0);
    if (i0 >= first && i0 + 3 < end) {
        uint32_t packed = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) packed |= (uint32_t)f.key(ctx, soa_load<Source>(sv.state_r, sv.cap, (uint32_t)i0 + j)) << (8 * j);
        *reinterpret_cast<uint32_t*>(la.blk_key + i0) = packed;
    } else
        for (uint64_t i = i0 < first ? first : i0; i < i0 + 4 && i < end; ++i) la.blk_key[i] = f.key(ctx, soa_load<Source>(sv.state_r, sv.cap, (uint32_t)i));
}
template <class F>
cudaError_t launch_keys(const LaunchArgs& la) {
    static thread_local KernelArgs ka;
    ka.la = la;
    ka.ds = *la.ds;
    ka.la.ds = nullptr;
    if (la.blk_nkeys == 0) return cudaSuccess;
    build_keys_kernel<F><<<(unsigned)(((unsigned long long)la.blk_nkeys + 3 + 1023) / 1024), 256, 0, la.stream>>>(ka);
    return cudaGetLastError();
}

// How selective is the prefilter right now?  Samples entries of the primary edge type's CSR (a random row of the called type, a random
// entry of it) and counts how many source keys the row's probe lets through: la.stats[0] += sampled, la.stats[1] += passed.  The engine
// chooses between the prefiltered and the unfiltered sweeps with it (a key that passes costs a DRAM-random fetch of the exact state:
// at ϵ = 0.25 of the docs' HK model half of the neighbours pass and the unfiltered sweeps are faster).
template <class F>
__global__ void __launch_bounds__(256) passrate_kernel(const __grid_constant__ KernelArgs ka) {
    typedef typename F::State State;
    typedef typename F::Source Source;
    const LaunchArgs& la = ka.la;
    const DeviceSim& ds = ka.ds;
    const AgentView& av = ds.agents[la.type];
    const AgentView& sv = ds.agents[F::kSourceType];
    const EdgeView& ev = ds.edges[F::kPrimaryEdge];
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t h = (t + 1) * 0x9e3779b97f4a7c15ull + ds.seed;
    h ^= h >> 32; h *= 0xd6e8feb86659fd93ull; h ^= h >> 32; h *= 0xd6e8feb86659fd93ull; h ^= h >> 32;
    const uint32_t idx = (uint32_t)(((h & 0xffffffffull) * la.n) >> 32);
    const uint32_t row = (ev.target ? 0u : ds.base[la.type]) + idx;
    bool sampled = false, pass = false;
    if (idx < la.n && row < ev.rows && !(av.died_r && av.died_r[idx])) {
        const uint32_t b = ev.off[row], e = ev.off[row + 1];
        if (e > b) {
            const uint32_t k = b + (uint32_t)(((h >> 32) * (uint64_t)(e - b)) >> 32);
            const uint32_t s = ev.src[k] - ds.base[F::kSourceType];
            if (s < sv.lcap + sv.nghost) {
                const F f{};
                Ctx<F, MODE_DIRECT, 1> ctx(ds, la, idx, 0);
                const State self = soa_load<State>(av.state_r, av.cap, idx);
                const Source nb = soa_load<Source>(sv.state_r, sv.cap, s);
                sampled = true;
                pass = f.may_accept(f.probe(ctx, self), (uint32_t)f.key(ctx, nb));
            }
        }
    }
    const unsigned ns = __popc(__ballot_sync(0xffffffffu, sampled)), np = __popc(__ballot_sync(0xffffffffu, pass));
    if ((threadIdx.x & 31) == 0 && ns) { atomicAdd(la.stats, (unsigned long long)ns); atomicAdd(la.stats + 1, (unsigned long long)np); }
}
template <class F>
cudaError_t launch_passrate(const LaunchArgs& la) {
    static thread_local KernelArgs ka;
    ka.la = la;
    ka.ds = *la.ds;
    ka.la.ds = nullptr;
    if (la.n == 0) return cudaSuccess;
    passrate_kernel<F><<<256, 256, 0, la.stream>>>(ka);          // 65 536 samples: +-0.4 % at a pass rate of one half
    return cudaGetLastError();
}

// ---- reduce transition over an implicit raster stencil (KIND_STENCIL): the grid-stencil kernel -----------------------------------
// A thread per cell.  The generic accessor path enumerates a row in the reference's insertion order (sorted keys for border cells,
// strided shares for lane groups); a reduce transition does not depend on the order, so this kernel only decodes the position,
// walks the stencil offsets and folds the neighbour cells' states (adjacent slots: the loads of a warp coalesce and hit L1).
template <class F>
__global__ void __launch_bounds__(256) reduce_stencil_kernel(const __grid_constant__ KernelArgs ka) {
    typedef typename F::State State;
    typedef typename F::Source Source;
    typedef typename F::Acc Acc;
    const LaunchArgs& la = ka.la;
    const DeviceSim& ds = ka.ds;
    const uint64_t gtid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t idx = (uint32_t)gtid;
    const AgentView& av = ds.agents[la.type];
    const bool act = gtid < la.n && !(av.died_r && av.died_r[idx]);
    uint32_t nedges = 0;
    if (act) {
        const EdgeView& ev = ds.edges[F::kPrimaryEdge];
        const RasterView& rv = ds.rasters[ev.st_raster];
        const StencilTab& tab = ds.stencils[ev.st_tab];
        const AgentView& sv = ds.agents[F::kSourceType];
        if (ds.check && (!ev.readable || !sv.readable)) atomicOr(ds.error, (uint32_t)(!ev.readable ? DERR_EDGE_NOT_READABLE : DERR_AGENT_NOT_READABLE));
        State self;
        if (la.in_read && av.size) self = soa_load<State>(av.state_r, av.cap, idx);
        else memset(&self, 0, sizeof(State));
        Ctx<F, MODE_DIRECT, 1> ctx(ds, la, idx, 0);
        const F f{};
        Acc acc;
        f.init(ctx, self, acc);
        const uint32_t lin = idx - ev.st_slot0;
        if (la.type == rv.type && idx >= ev.st_slot0 && lin < rv.ncells) {
            if (F::kSourceType != rv.type) atomicOr(ds.error, (uint32_t)DERR_AGENT_TYPE_MISMATCH);
            else {
                const uint8_t* __restrict__ st = sv.state_r;
                const uint32_t cap = sv.cap, s0 = ev.st_slot0;
                int32_t pos[MAX_RASTER_DIMS];
                bool interior = true;
                uint32_t rest = lin;
#pragma unroll
                for (int k = 0; k < MAX_RASTER_DIMS; ++k) {
                    if (k < rv.ndims) {
                        const uint32_t d = rv.dim32[k];
                        const uint32_t q = (d & (d - 1)) == 0 ? rest >> (31 - __clz(d)) : rest / d;    // power-of-two extents: a shift
                        pos[k] = (int32_t)(rest - q * d); rest = q;
                        interior &= pos[k] >= (int32_t)ev.st_reach && pos[k] + (int32_t)ev.st_reach < (int32_t)d;
                    } else pos[k] = 0;
                }
                if (interior) {
                    for (int si = 0; si < ev.st_n; ++si) f.fold(ctx, self, soa_load<Source>(st, cap, s0 + (uint32_t)((int32_t)lin - tab.lin[si])), acc);
                    nedges = (uint32_t)ev.st_n;
                } else {
                    for (int si = 0; si < ev.st_n; ++si) {
                        uint32_t l = 0; bool ok = true;
#pragma unroll
                        for (int k = 0; k < MAX_RASTER_DIMS; ++k) {
                            if (k >= rv.ndims) break;
                            int32_t v = pos[k] - tab.off[si][k];
                            const int32_t d = (int32_t)rv.dim32[k];
                            if (v < 0 || v >= d) { if (!ev.st_periodic) { ok = false; break; } v %= d; if (v < 0) v += d; }
                            l += (uint32_t)v * rv.stride32[k];
                        }
                        if (!ok) continue;
                        f.fold(ctx, self, soa_load<Source>(st, cap, s0 + l), acc);
                        ++nedges;
                    }
                }
            }
        }
        const AgentID id = agent_id((uint32_t)la.type, ds.rank, (uint64_t)idx + 1);
        const bool alive = f.finish(ctx, self, id, acc);
        if (la.in_write) {                                                 // transition_with_write! (AgentMethods.jl:159-181)
            if (alive) { if (av.size) soa_store<State>(av.independent ? av.state_r : av.state_w, av.cap, idx, self); }
            else if (av.immortal) atomicOr(ds.error, (uint32_t)DERR_IMMORTAL_DIED);
            else av.died_w[idx] = 1;
        }
    }
    const unsigned total = __reduce_add_sync(0xffffffffu, nedges);
    if ((threadIdx.x & 31) == 0 && total) atomicAdd(la.stats + ((blockIdx.x & 1023u) << 2), (unsigned long long)total);
}
// The Moore stencil on a two-dimensional raster, bandwidth-shaped (profiles/microbench/stencil.cu, profiles/r2_microbench/stencil.txt:
// 0.054 ms per generation with a thread per cell and 8 loads, 0.032 ms with this shape at 4096 x 4096): a warp marches along the
// second dimension of a strip of 30 cells of the first (contiguous) one.  Every lane loads ONE state per row — the warp's loads are 32
// adjacent slots — keeps three rows of its column in registers (sliding window) and gets the left / right neighbours by shuffle; lanes 0
// and 31 only carry the halo columns.  No position decode, no div / mod per cell, 8 folds from registers.  Cells outside a clipped
// (non-periodic) raster are marked invalid and skipped.  Requires State == Source (the cells read each other), an :Immortal cell type
// whose slots are exactly the raster's cells, and the full 3 x 3 neighbourhood; everything else runs reduce_stencil_kernel.
template <class T> __device__ __forceinline__ T shfl_any(const T& v, uint32_t src) {
    constexpr int NW = (sizeof(T) + 3) / 4;
    union U { T t; uint32_t w[NW]; __device__ U() {} } a, b;
#pragma unroll
    for (int i = 0; i < NW; ++i) a.w[i] = 0;
    a.t = v;
#pragma unroll
    for (int i = 0; i < NW; ++i) b.w[i] = __shfl_sync(0xffffffffu, a.w[i], src);
    return b.t;
}
template <class F, int RY, bool PERIODIC>
__global__ void __launch_bounds__(256) reduce_stencil_strip_kernel(const __grid_constant__ KernelArgs ka) {
    typedef typename F::State State;
    typedef typename F::Acc Acc;
    const LaunchArgs& la = ka.la;
    const DeviceSim& ds = ka.ds;
    const AgentView& av = ds.agents[la.type];
    const EdgeView& ev = ds.edges[F::kPrimaryEdge];
    const RasterView& rv = ds.rasters[ev.st_raster];
    const uint32_t nx = rv.dim32[0], ny = rv.dim32[1];
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t strips = (nx + 29) / 30, bands = (ny + RY - 1) / RY;
    if (warp >= strips * bands) return;
    const uint32_t sx = warp % strips, by = warp / strips;
    constexpr bool periodic = PERIODIC;                                     // a periodic raster has no invalid neighbours: the tests below fold away
    const int32_t xs = (int32_t)(sx * 30) - 1 + (int32_t)lane;              // the column this lane carries (lanes 0 / 31: halo)
    const bool xin = xs >= 0 && (uint32_t)xs < nx;
    const uint32_t x = (uint32_t)((xs % (int32_t)nx + (int32_t)nx) % (int32_t)nx);
    const bool xvalid = xin || periodic;                                    // a wrapped halo column exists only on a periodic raster
    const bool owner = lane >= 1 && lane <= 30 && xin;
    const uint32_t y0 = by * RY, y1 = y0 + RY < ny ? y0 + RY : ny;
    const uint8_t* __restrict__ st = av.state_r;
    const uint32_t cap = av.cap;
    const F f{};
    auto row = [&](uint32_t y) { return soa_load<State>(st, cap, y * nx + x); };
    // rows above / below a clipped raster are invalid as a whole
    bool upv = periodic || y0 > 0;
    State up = row(y0 ? y0 - 1 : ny - 1), mid = row(y0);
    State upl = shfl_any(up, (lane + 31) & 31), upr = shfl_any(up, (lane + 1) & 31);
    State midl = shfl_any(mid, (lane + 31) & 31), midr = shfl_any(mid, (lane + 1) & 31);
    const bool lv = periodic || __shfl_sync(0xffffffffu, (int)xvalid, (lane + 31) & 31) != 0, rvd = periodic || __shfl_sync(0xffffffffu, (int)xvalid, (lane + 1) & 31) != 0;
    uint32_t nedges = 0;
    State nxt = row(y0 + 1 < ny ? y0 + 1 : 0);                              // the row below is loaded one iteration ahead of its use
    for (uint32_t y = y0; y < y1; ++y) {
        const bool dnv = periodic || y + 1 < ny;
        const State dn = nxt;
        if (y + 1 < y1) nxt = row(y + 2 < ny ? y + 2 : y + 2 - ny);
        const State dnl = shfl_any(dn, (lane + 31) & 31), dnr = shfl_any(dn, (lane + 1) & 31);
        if (owner) {
            const uint32_t idx = y * nx + x;
            Ctx<F, MODE_DIRECT, 1> ctx(ds, la, idx, 0);
            State self = mid;
            Acc a;
            f.init(ctx, self, a);
            if (upv) { if (lv) { f.fold(ctx, self, upl, a); ++nedges; } f.fold(ctx, self, up, a); ++nedges; if (rvd) { f.fold(ctx, self, upr, a); ++nedges; } }
            if (lv) { f.fold(ctx, self, midl, a); ++nedges; }
            if (rvd) { f.fold(ctx, self, midr, a); ++nedges; }
            if (dnv) { if (lv) { f.fold(ctx, self, dnl, a); ++nedges; } f.fold(ctx, self, dn, a); ++nedges; if (rvd) { f.fold(ctx, self, dnr, a); ++nedges; } }
            const AgentID id = agent_id((uint32_t)la.type, ds.rank, (uint64_t)idx + 1);
            const bool alive = f.finish(ctx, self, id, a);
            if (la.in_write) {                                             // transition_with_write! (AgentMethods.jl:159-181)
                if (alive) soa_store<State>(av.independent ? av.state_r : av.state_w, av.cap, idx, self);
                else atomicOr(ds.error, (uint32_t)DERR_IMMORTAL_DIED);     // (the fast path only serves :Immortal cell types)
            }
        }
        up = mid; upl = midl; upr = midr; mid = dn; midl = dnl; midr = dnr; upv = true;
    }
    const unsigned total = __reduce_add_sync(0xffffffffu, nedges);
    if (lane == 0 && total) atomicAdd(la.stats + ((blockIdx.x & 1023u) << 2), (unsigned long long)total);
}
template <class F>
cudaError_t launch_stencil(const LaunchArgs& la) {
    static thread_local KernelArgs ka;
    ka.la = la;
    ka.ds = *la.ds;
    ka.la.ds = nullptr;
    if (la.n == 0) return cudaSuccess;
    if constexpr (std::is_same<typename F::State, typename F::Source>::value) {
        const DeviceSim& ds = ka.ds;
        const EdgeView& ev = ds.edges[F::kPrimaryEdge];
        const AgentView& av = ds.agents[la.type];
        static const bool strip_on = !(getenv("VB_STENCIL_STRIP") && atoi(getenv("VB_STENCIL_STRIP")) == 0);
        if (strip_on && ev.st_raster >= 0) {
            const RasterView& rv = ds.rasters[ev.st_raster];
            const bool ok = rv.ndims == 2 && ev.st_n == 8 && ev.st_reach == 1 && rv.type == la.type && F::kSourceType == la.type && ev.st_slot0 == 0 &&
                            la.n == rv.ncells && av.died_r == nullptr && av.size == sizeof(typename F::State) && la.in_read && rv.dim32[0] >= 3 && rv.dim32[1] >= 3 &&
                            (!ds.check || (ev.readable && av.readable));
            if (ok) {
                constexpr int RY = 32;
                const unsigned long long warps = (unsigned long long)((rv.dim32[0] + 29) / 30) * ((rv.dim32[1] + RY - 1) / RY);
                if (ev.st_periodic) reduce_stencil_strip_kernel<F, RY, true><<<(unsigned)((warps * 32 + 255) / 256), 256, 0, la.stream>>>(ka);
                else reduce_stencil_strip_kernel<F, RY, false><<<(unsigned)((warps * 32 + 255) / 256), 256, 0, la.stream>>>(ka);
                return cudaGetLastError();
            }
        }
    }
    reduce_stencil_kernel<F><<<(unsigned)(((unsigned long long)la.n + 255) / 256), 256, 0, la.stream>>>(ka);
    return cudaGetLastError();
}

// default shape: 2 warps per CTA (parity-tested on the GPU with VB_PF_WARPS=2; 8 and 4 stay selectable through VB_PF_WARPS)
#ifndef VB_PF_WARPS_DEFAULT
#define VB_PF_WARPS_DEFAULT 2
#endif
template <class F, int WPC, int RF = 0>
cudaError_t launch_prefilter(KernelArgs& ka, unsigned long long work) {
    const LaunchArgs& la = ka.la;
    constexpr unsigned TPB = 32 * WPC;
    static int pwave = 0;                                               // resident CTAs of this shape on the whole device
    if (!pwave) {
        int dev = 0, sms = 148, per_sm = 64 / WPC;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, reduce_prefilter_kernel<F, false, false, WPC, RF>, (int)TPB, 0);
        pwave = sms * (per_sm > 0 ? per_sm : 1);
        if (getenv("VB_BLOCK_AHEAD")) pwave = atoi(getenv("VB_BLOCK_AHEAD"));
    }
    ka.la.blk_ahead = (uint32_t)pwave;
    const unsigned grid = (unsigned)((work + TPB - 1) / TPB);
    if (la.blk_first && la.blk_last) reduce_prefilter_kernel<F, true, true, WPC, RF><<<grid, TPB, 0, la.stream>>>(ka);   // all keys in one block
    else if (la.blk_first) reduce_prefilter_kernel<F, true, false, WPC, RF><<<grid, TPB, 0, la.stream>>>(ka);
    else if (la.blk_last) reduce_prefilter_kernel<F, false, true, WPC, RF><<<grid, TPB, 0, la.stream>>>(ka);
    else reduce_prefilter_kernel<F, false, false, WPC, RF><<<grid, TPB, 0, la.stream>>>(ka);
    return cudaGetLastError();
}
#ifndef VB_SEG_WARPS
#define VB_SEG_WARPS 2
#endif
#ifndef VB_SEG_MINB
#define VB_SEG_MINB 24
#endif
template <class F>
cudaError_t launch_segsweep(KernelArgs& ka) {
    const LaunchArgs& la = ka.la;
    constexpr int WPC = VB_SEG_WARPS, MINB = VB_SEG_MINB;
    static int pwave = 0;                                               // resident CTAs of this shape on the whole device
    if (!pwave) {
        int dev = 0, sms = 148, per_sm = MINB;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, reduce_segsweep_kernel<F, false, false, WPC, MINB>, 32 * WPC, 0);
        pwave = sms * (per_sm > 0 ? per_sm : 1);
        if (getenv("VB_BLOCK_AHEAD")) pwave = atoi(getenv("VB_BLOCK_AHEAD"));
    }
    ka.la.blk_ahead = (uint32_t)pwave;
    const unsigned long long groups = ((unsigned long long)la.blk_nseg + 31) / 32;
    const unsigned grid = (unsigned)((groups + WPC - 1) / WPC);
    if (la.blk_first && la.blk_last) reduce_segsweep_kernel<F, true, true, WPC, MINB><<<grid, 32 * WPC, 0, la.stream>>>(ka);
    else if (la.blk_first) reduce_segsweep_kernel<F, true, false, WPC, MINB><<<grid, 32 * WPC, 0, la.stream>>>(ka);
    else if (la.blk_last) reduce_segsweep_kernel<F, false, true, WPC, MINB><<<grid, 32 * WPC, 0, la.stream>>>(ka);
    else reduce_segsweep_kernel<F, false, false, WPC, MINB><<<grid, 32 * WPC, 0, la.stream>>>(ka);
    return cudaGetLastError();
}
template <class F>
cudaError_t launch_blocked(const LaunchArgs& la) {
    static thread_local KernelArgs ka;
    ka.la = la;
    ka.ds = *la.ds;
    ka.la.ds = nullptr;
    if (la.n == 0) return cudaSuccess;
    const bool listed = !la.blk_first && !la.blk_last && la.blk_rows != nullptr;
    if (listed && la.blk_nrows == 0) return cudaSuccess;
    const unsigned grid = (unsigned)(((unsigned long long)(listed ? la.blk_nrows : la.n) + 255) / 256);
    static int wave = 0;                                                // resident CTAs of this kernel on the whole device
    if (!wave) {
        int dev = 0, sms = 148, per_sm = 6;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, reduce_blocked_kernel<F, false, false>, 256, 0);
        wave = sms * (per_sm > 0 ? per_sm : 1);
        if (getenv("VB_BLOCK_AHEAD")) wave = atoi(getenv("VB_BLOCK_AHEAD"));
    }
    ka.la.blk_ahead = (uint32_t)wave;
    if constexpr (F::kPrefilter) {
        if (la.blk_op == 1) {                                           // hub rows of the segmented view
            if (la.blk_nhub == 0) return cudaSuccess;
            reduce_hubmerge_kernel<F><<<(la.blk_nhub + 127) / 128, 128, 0, la.stream>>>(ka);
            return cudaGetLastError();
        }
        if (la.blk_prefilter && la.blk_seg_row) {                       // segmented view
            if (!la.blk_key || la.blk_nseg == 0) return la.blk_key ? cudaSuccess : cudaErrorInvalidValue;
            return launch_segsweep<F>(ka);
        }
        if (la.blk_prefilter) {
            if (!la.blk_key) return cudaErrorInvalidValue;
            static const int wpc = [] { const char* e = getenv("VB_PF_WARPS"); const int v = e ? atoi(e) : VB_PF_WARPS_DEFAULT; return v == 2 || v == 4 ? v : 8; }();
            const unsigned long long work = listed ? la.blk_nrows : la.n;
            static const bool rowfind = getenv("VB_PF_ROWFIND") && atoi(getenv("VB_PF_ROWFIND")) != 0;     // experimental, see reduce_prefilter_kernel
            if (rowfind) return wpc == 8 ? launch_prefilter<F, 8, 1>(ka, work) : launch_prefilter<F, 2, 1>(ka, work);
            if (wpc == 2) return launch_prefilter<F, 2>(ka, work);
            if (wpc == 4) return launch_prefilter<F, 4>(ka, work);
            return launch_prefilter<F, 8>(ka, work);
        }
    }
    if (la.blk_first && la.blk_last) return cudaErrorInvalidValue;      // a single block is the direct path's job
    if (la.blk_first) reduce_blocked_kernel<F, true, false><<<grid, 256, 0, la.stream>>>(ka);
    else if (la.blk_last) reduce_blocked_kernel<F, false, true><<<grid, 256, 0, la.stream>>>(ka);
    else reduce_blocked_kernel<F, false, false><<<grid, 256, 0, la.stream>>>(ka);
    return cudaGetLastError();
}

template <class F>
TransitionInfo make_transition_info(const char* name, const char* agent_type) {
    TransitionInfo ti{};
    ti.name = name;
    ti.agent_type = agent_type;
    ti.state_size = (uint32_t)sizeof(typename F::State);
    ti.cooperative = F::kCooperative;
    ti.n_edge_writes = F::EdgeWrites::size;
    for (int i = 0; i < F::EdgeWrites::size; ++i) ti.edge_writes[i] = F::EdgeWrites::at(i);
    ti.n_agent_writes = F::AgentWrites::size;
    for (int i = 0; i < F::AgentWrites::size; ++i) ti.agent_writes[i] = F::AgentWrites::at(i);
    ti.n_edge_removes = F::EdgeRemoves::size;
    for (int i = 0; i < F::EdgeRemoves::size; ++i) ti.edge_removes[i] = F::EdgeRemoves::at(i);
    ti.primary_edge = F::kPrimaryEdge;
    ti.launch = &launch_transition<F>;
    if constexpr (F::kReduce) {
        ti.reduce = 1;
        ti.source_type = F::kSourceType;
        ti.source_size = (uint32_t)sizeof(typename F::Source);
        ti.acc_bytes = (uint32_t)F::kAccBytes;
        // the source-blocked sweeps park states / accumulators as 4- and 8-byte words
        if constexpr (sizeof(typename F::State) % 4 == 0 && sizeof(typename F::Acc) % 4 == 0 && F::kAccBytes % 4 == 0) ti.launch_blocked = &launch_blocked<F>;
        ti.launch_stencil = &launch_stencil<F>;
        if constexpr (F::kPrefilter) {
            static_assert(sizeof(typename F::Probe) % 4 == 0, "Probe: a multiple of 4 bytes");
            if (ti.launch_blocked) { ti.prefilter = 1; ti.launch_keys = &launch_keys<F>; ti.launch_passrate = &launch_passrate<F>; }
        }
    }
    return ti;
}

// ---- registered map functors of mapreduce (K8) ------------------------------------------------------------------------------------
__device__ __forceinline__ double map_identity(int op, double) {
    return op == OP_PROD ? 1.0 : op == OP_MIN ? INFINITY : op == OP_MAX ? -INFINITY : 0.0;
}
__device__ __forceinline__ long long map_identity(int op, long long) {
    return op == OP_PROD ? 1LL : op == OP_MIN ? 0x7fffffffffffffffLL : op == OP_MAX ? (-0x7fffffffffffffffLL - 1) : op == OP_AND ? -1LL : 0LL;
}
__device__ __forceinline__ double map_fold(double a, double b, int op) {
    return op == OP_SUM ? a + b : op == OP_PROD ? a * b : op == OP_MIN ? fmin(a, b) : fmax(a, b);
}
__device__ __forceinline__ long long map_fold(long long a, long long b, int op) {
    switch (op) {
        case OP_SUM: return (long long)((unsigned long long)a + (unsigned long long)b);
        case OP_PROD: return (long long)((unsigned long long)a * (unsigned long long)b);
        case OP_MIN: return a < b ? a : b;
        case OP_MAX: return a > b ? a : b;
        case OP_AND: return a & b;
        default: return a | b;
    }
}
template <class R> struct MapAcc { typedef long long type; };
template <> struct MapAcc<double> { typedef double type; };
template <> struct MapAcc<float> { typedef double type; };
// grid-stride over the elements (coalesced SoA loads, HBM-bound), lanes -> warps -> CTA; the engine folds the partials
template <class F>
__global__ void __launch_bounds__(256) map_kernel(const MapLaunchArgs a) {
    typedef typename F::Elem Elem;
    typedef typename MapAcc<typename F::Result>::type T;
    __shared__ T sm[8];
    const F f{};
    T acc = map_identity(a.op, T());
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += (uint64_t)gridDim.x * blockDim.x) {
        if (a.died && a.died[i]) continue;
        acc = map_fold((T)f(soa_load<Elem>(a.cols, a.stride, (uint32_t)i)), acc, a.op);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc = map_fold(acc, __shfl_xor_sync(0xffffffffu, acc, o), a.op);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        T v = threadIdx.x < 8 ? sm[threadIdx.x] : map_identity(a.op, T());
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) v = map_fold(v, __shfl_xor_sync(0xffffffffu, v, o), a.op);
        if (threadIdx.x == 0) reinterpret_cast<T*>(a.partial)[blockIdx.x] = v;
    }
}
template <class F>
__global__ void __launch_bounds__(256) map_cells_kernel(const MapCellsArgs a) {
    typedef typename F::Elem Elem;
    typedef typename MapAcc<typename F::Result>::type T;
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    const F f{};
    if (a.cells[i] == 0xffffffffu) { reinterpret_cast<T*>(a.out)[i] = (T)0; return; }     // a cell of another rank: its owner fills it in (joined by the engine)
    reinterpret_cast<T*>(a.out)[i] = (T)f(soa_load<Elem>(a.cols, a.stride, a.cells[i] - a.cbase));
}
template <class F>
cudaError_t launch_map_cells(const MapCellsArgs& a) {
    if (a.n) map_cells_kernel<F><<<(unsigned)((a.n + 255) / 256), 256, 0, a.stream>>>(a);
    return cudaGetLastError();
}
template <class F>
cudaError_t launch_map(const MapLaunchArgs& a) {
    map_kernel<F><<<a.nblocks, 256, 0, a.stream>>>(a);
    return cudaGetLastError();
}
template <class F>
MapInfo make_map_info(const char* name, const char* type_name) {
    static_assert(std::is_base_of<MapBase, F>::value, "map functors derive from vb::MapBase");
    MapInfo mi{};
    mi.name = name; mi.type_name = type_name;
    mi.elem_size = (uint32_t)sizeof(typename F::Elem);
    mi.is_float = std::is_floating_point<typename F::Result>::value ? 1 : 0;
    mi.launch = &launch_map<F>;
    mi.launch_cells = &launch_map_cells<F>;
    return mi;
}

#define VB_CAT2(a, b) a##b
#define VB_CAT(a, b) VB_CAT2(a, b)
// Registers Functor as the map `mname` of mapreduce over agents / edges of type `tname`.
#define VB_REGISTER_MAP(mname, tname, ...)                                                     \
    static const int VB_CAT(vb_regmap_, __COUNTER__) = [] {                                    \
        static const vb::MapInfo mi = vb::make_map_info<__VA_ARGS__>(mname, tname);            \
        return vb_register_map(&mi);                                                           \
    }();
// Registers Functor as the transition `tname` for agents of type `atype` (a string: the registered name).
#define VB_REGISTER_TRANSITION(tname, atype, ...)                                              \
    static const int VB_CAT(vb_reg_, __COUNTER__) = [] {                                       \
        static const vb::TransitionInfo ti = vb::make_transition_info<__VA_ARGS__>(tname, atype); \
        return vb_register_transition(&ti);                                                    \
    }();
#endif  // __CUDACC__

}  // namespace vb
