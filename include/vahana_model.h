// vahana_model.h — types shared by the host engine, the device API and the CPU oracle.
//
// Everything here is plain C++14 that compiles under gcc and nvcc.  Transition functors
// (vahana.jl_b200/csrc/transitions/*.h) are written once against the "Ctx concept"
// documented in include/vahana_device.cuh and are instantiated twice: with the CUDA
// context (product) and with the sequential oracle context (tests / CPU baseline).
//
// Reference semantics restated here:
//   AgentID bit packing          /root/reference/src/Agent.jl:30-118
//   hint names                   /root/reference/src/ModelTypes.jl:81-230
//   reduce-op identities         /root/reference/src/Helpers.jl:44-81
//   stencil metrics              /root/reference/src/Raster.jl:82-110
#pragma once
#include <stdint.h>
#include <stddef.h>

#if defined(__CUDACC__)
#define VB_HD __host__ __device__ __forceinline__
#define VB_D __device__ __forceinline__
#else
#define VB_HD inline
#define VB_D inline
#endif

namespace vb {

typedef uint64_t AgentID;

// Agent.jl:30-64: id = type:8 | rank:20 | nr:36, nr is 1-based.
enum : int { BITS_TYPE = 8, BITS_PROCESS = 20, BITS_AGENTNR = 36 };
enum : int { SHIFT_TYPE = BITS_PROCESS + BITS_AGENTNR, SHIFT_RANK = BITS_AGENTNR };
enum : int { MAX_TYPES = 1 << BITS_TYPE };

VB_HD AgentID agent_id(uint32_t type, uint32_t rank, uint64_t nr) {
    return ((AgentID)type << SHIFT_TYPE) + ((AgentID)rank << SHIFT_RANK) + nr;   // Agent.jl:67-74
}
VB_HD uint32_t type_nr(AgentID id) { return (uint32_t)(id >> SHIFT_TYPE); }                       // Agent.jl:88
VB_HD uint32_t process_nr(AgentID id) { return (uint32_t)((id >> SHIFT_RANK) & ((1u << BITS_PROCESS) - 1)); }  // :105
VB_HD uint64_t agent_nr(AgentID id) { return id & ((1ull << BITS_AGENTNR) - 1); }                 // Agent.jl:113
VB_HD AgentID remove_process(AgentID id) {                                                        // Agent.jl:80-81
    return id & ~((((AgentID)1 << BITS_PROCESS) - 1) << BITS_AGENTNR);
}

// Type references used by apply()'s call/read/write/add_existing lists: agent types are
// their 1-based type id (registration order, ModelTypes.jl:84-89), edge types are
// VB_EDGE_REF + 0-based registration index.
enum : int { EDGE_REF = 256 };

enum AgentHint : uint32_t { AGENT_IMMORTAL = 1, AGENT_INDEPENDENT = 2 };
enum EdgeHint : uint32_t {
    EDGE_STATELESS = 1, EDGE_IGNORE_FROM = 2, EDGE_SINGLE_EDGE = 4, EDGE_SINGLE_TYPE = 8,
    EDGE_IGNORE_SOURCE_STATE = 16
};
enum Metric : int { CHEBYSHEV = 0, EUCLIDEAN = 1, MANHATTEN = 2 };   // Raster.jl:83 (spelling as in the reference)
enum ReduceOp : int { OP_SUM = 0, OP_PROD = 1, OP_MIN = 2, OP_MAX = 3, OP_AND = 4, OP_OR = 5 };
enum DType : int { DT_I64 = 0, DT_F64 = 1, DT_BOOL = 2, DT_I32 = 3, DT_F32 = 4, DT_U8 = 5 };

// Philox4x32-10 — the counter-based generator that stands in for the per-agent
// pre-generated uniform table (north_star: "identical pre-generated per-agent uniform
// draws").  uniform(seed, agent slot, k) is a pure function, so the oracle and the
// kernels read the same table without materialising it.
struct Philox {
    static VB_HD void round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
        const uint32_t n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
        const uint32_t n3 = (uint32_t)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    }
    static VB_HD void gen(uint64_t seed, uint64_t ctr_lo, uint64_t ctr_hi, uint32_t (&out)[4]) {
        uint32_t c[4] = {(uint32_t)ctr_lo, (uint32_t)(ctr_lo >> 32), (uint32_t)ctr_hi, (uint32_t)(ctr_hi >> 32)};
        uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
        for (int i = 0; i < 10; ++i) {
            round(c, k0, k1);
            k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
        }
        out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
    }
    // 53-bit uniform in [0,1): exact in double, identical on host and device.
    static VB_HD double uniform(uint64_t seed, uint64_t a, uint64_t b) {
        uint32_t r[4];
        gen(seed, a, b, r);
        const uint64_t bits = (((uint64_t)r[0] << 32) | r[1]) >> 11;
        return (double)bits * (1.0 / 9007199254740992.0);
    }
    static VB_HD uint64_t u64(uint64_t seed, uint64_t a, uint64_t b) {
        uint32_t r[4];
        gen(seed, a, b, r);
        return ((uint64_t)r[0] << 32) | r[1];
    }
};

// A registered map of mapreduce — the closure `f` of mapreduce(sim, f, op, T) (src/AgentMethods.jl:533-565, src/EdgeMethods.jl:972-994)
// when it is more than a field selector, e.g. `b -> b.x - b.y` (docs/examples/tutorial1.jl:548).  Single source like the transitions:
//     struct XMinusY : vb::MapBase { using Elem = Bought; using Result = double; VB_HD double operator()(const Bought& b) const { return b.x - b.y; } };
// Elem = the agent state (live agents are mapped) or the edge state (every edge is mapped); Result = double or int64_t (Bool: 0 / 1).
struct MapBase {};

// Largest power-of-two word (<= 16 B) that divides a state size: the SoA column width
// the engine stores agent and edge states in (see DESIGN.md "data layout").
VB_HD uint32_t soa_word(uint32_t size) {
    if (size == 0) return 0;
    uint32_t w = 16;
    while (size % w) w >>= 1;
    return w;
}

// AgentID of the g-th agent (0-based, global) of a type whose n agents are spread over nranks in contiguous equal blocks
// (_create_equal_partition, src/Simulation.jl:353-367): larger blocks first, nr counts from 1 inside the owner's block.
VB_HD AgentID block_partition_id(uint32_t type, uint64_t g, uint64_t n, uint32_t nranks) {
    const uint64_t q = n / nranks, r = n % nranks;
    uint64_t p, local;
    if (g < r * (q + 1)) { p = g / (q + 1); local = g % (q + 1); }
    else { p = r + (g - r * (q + 1)) / q; local = (g - r * (q + 1)) % q; }
    return agent_id(type, (uint32_t)p, local + 1);
}

// Compile-time list of type ids: a transition declares the edge / agent types it appends to.
template <int... Is> struct IntList {
    static constexpr int size = sizeof...(Is);
    static VB_HD int at(int i) {
        const int v[sizeof...(Is) + 1] = {Is..., -1};
        return v[i];
    }
    static VB_HD int find(int x) {
        const int v[sizeof...(Is) + 1] = {Is..., -1};
        for (int i = 0; i < (int)sizeof...(Is); ++i) if (v[i] == x) return i;
        return -1;
    }
};
// Defaults every transition functor inherits (see include/vahana_device.cuh).
struct TransitionBase {
    static constexpr bool kCooperative = false;   // true: all lanes of a warp run the functor for one agent
    using EdgeWrites = IntList<>;                 // edge types the functor calls add_edge on
    using AgentWrites = IntList<>;                // agent types the functor calls add_agent on
    using EdgeRemoves = IntList<>;                // edge types the functor calls remove_edges on
    static constexpr int kPrimaryEdge = -1;       // cooperative functors: the edge type whose row the functor walks;
                                                  // lets the engine bin agents by degree (sub-warp / warp / block per agent)
    static constexpr bool kReduce = false;        // true for vb::ReduceTransition functors (init / fold / merge / finish)
};

// A *reduce transition*: the common shape "fold the states of my in-neighbours, then update myself"
// (neighborstates + reduce in the reference's models, e.g. docs/examples/hegselmann.jl:134-144) split into its parts, so that
// the engine is free to choose how the row is walked: lanes of a group with a strided share each (direct path), or one sweep per
// L2-sized block of the source-state array with the accumulator parked in HBM between sweeps (source-blocked path).
//
//     struct Step : vb::ReduceTransition<Step> {
//         using State = HKAgent;  using Source = HKAgent;           // called agent / neighbour state
//         struct Acc { double sum; uint32_t n; };                    // trivially copyable accumulator
//         static constexpr int kAccBytes = 12;                       // leading bytes of Acc that carry information
//         static constexpr int kPrimaryEdge = E_KNOWS, kSourceType = T_HKAGENT;
//         template <class Ctx> VB_HD void init(const Ctx&, const State& self, Acc& a) const;          // a = identity of merge
//         template <class Ctx> VB_HD void fold(const Ctx&, const State& self, const Source& nb, Acc& a) const;
//         VB_HD void merge(Acc& a, const Acc& b) const;              // commutative, associative up to rounding
//         template <class Ctx> VB_HD bool finish(const Ctx&, State& self, vb::AgentID id, const Acc& a) const;   // false = die
//     };
//
// The oracle folds the row strictly left to right (one lane); the GPU paths combine partial accumulators with merge().
//
// Optional *prefilter* (kPrefilter = true) for selective folds — folds that leave the accumulator untouched for most neighbours,
// like the bounded-confidence test of hegselmann.jl:139.  The functor names a one-byte key of a neighbour's state; the engine keeps
// the keys of the source type in a column of their own (1 B per slot: L2-sized where the 8 B states are not), gathers the key of
// every entry and fetches the full state only where may_accept() says the fold could change the accumulator:
//         using Probe = ...;                                                              // what a row needs to judge a key (small, trivially copyable)
//         template <class Ctx> VB_HD uint8_t key(const Ctx&, const Source& nb) const;
//         template <class Ctx> VB_HD Probe probe(const Ctx&, const State& self) const;    // once per row
//         VB_HD bool may_accept(const Probe&, uint32_t key) const;
// Contract: may_accept(probe(self), key(nb)) == false  ==>  fold(self, nb, a) leaves a unchanged.  The decision itself is always
// taken by fold() on the exact state, so results do not depend on the prefilter (tests/test_hk.py pins that, CPU and GPU).
template <class D> struct ReduceTransition : TransitionBase {
    static constexpr bool kCooperative = true;
    static constexpr bool kReduce = true;
    static constexpr bool kPrefilter = false;
    template <class Ctx, class S> VB_HD bool operator()(Ctx& ctx, S& self, AgentID id) const {
        const D& d = static_cast<const D&>(*this);
        typename D::Acc a;
        d.init(ctx, self, a);
        ctx.template for_each_neighborstate<typename D::Source>(D::kPrimaryEdge, D::kSourceType, id,
                                                                [&](const typename D::Source& nb) { d.fold(ctx, self, nb, a); });
        a = ctx.reduce_acc(a, [&](typename D::Acc& x, const typename D::Acc& y) { d.merge(x, y); });
        return d.finish(ctx, self, id, a);
    }
};

// Raster position (up to 4 dimensions, 1-based like CartesianIndex).
enum : int { MAX_RASTER_DIMS = 4 };
struct Pos {
    int64_t v[MAX_RASTER_DIMS];
};

}  // namespace vb
