// vahana_oracle.hpp — CPU ORACLE (test infrastructure, NOT product code).
//
// A sequential C++ restatement of the reference's transition hot path with the
// reference's own container choices (Dict{AgentID,Vector{Edge}} -> open-addressing hash map
// of per-target vectors, Vector for :SingleType, LIFO slot reuse, sequential per-agent loop).
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// load the library built from this file.  The product (vahana.jl_b200/csrc) never includes,
// links or calls it.
//
// Parity status: pinned against the reference's own known-answer tests, restated with the reference's values in
// tests/test_core.py (test/core.jl), test_edges.py (test/edges.jl), test_lifecycle.py (test/remove_agents.jl,
// test/addexisting.jl, test/independent.jl, test/graphs.jl, test/edgesiterator.jl), test_raster.py (test/raster.jl),
// test_remove_edges.py (test/mpi/test_edgetypes.jl, single rank), test_zx_misc.py (test/globals.jl,
// test/parametric_types.jl) and test_zy_full_size.py (random_pos, test/raster.jl:438-459).  The model-level outputs of the
// BASELINE configs (Hegselmann-Krause opinions, Game of Life, predator/prey, SIR) are NOT pinned by any reference test and
// Julia cannot run here: "parity unpinned" for those.  What stands in: independent numpy restatements that this
// oracle matches bit for bit (HK at config 1's full size and on random multigraphs, Game of Life, SIR: tests/test_zzn_sir_numpy.py,
// the market model of docs/examples/tutorial1.jl with an independent Philox4x32-10: tests/test_zzm_market.py), a pure-Python engine
// in the reference's container shapes for predator/prey (tests/test_zzp_pp_restatement.py), and the committed fixtures of
// tests/golden/ (oracle outputs, written by tests/golden/make_golden.py; the SIR and predator/prey ones are reproduced by the
// second implementations alone).
//
// Each function cites the reference lines it follows (paths relative to /root/reference).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <type_traits>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "../include/vahana_model.h"

namespace vo {

using vb::AgentID;

struct AssertionError : std::runtime_error {
    explicit AssertionError(const std::string& m) : std::runtime_error(m) {}
};
struct ArgError : std::runtime_error {
    explicit ArgError(const std::string& m) : std::runtime_error(m) {}
};

// ---------------------------------------------------------------------------------------------
// Open-addressing map AgentID -> V (stand-in for Julia's Dict{AgentID,V}; Julia's Dict is open
// addressing too, so the CPU baseline is not handicapped by node-based buckets).
template <class V>
class FlatMap {
    std::vector<uint64_t> keys_;
    std::vector<uint8_t> st_;  // 0 empty, 1 full, 2 tombstone
    std::vector<V> vals_;
    size_t n_ = 0, used_ = 0;
    static size_t hash(uint64_t k) {
        k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL; k ^= k >> 33;
        return (size_t)k;
    }
    void rehash(size_t cap) {
        std::vector<uint64_t> ok; std::vector<uint8_t> os; std::vector<V> ov;
        ok.swap(keys_); os.swap(st_); ov.swap(vals_);
        keys_.assign(cap, 0); st_.assign(cap, 0); vals_.clear(); vals_.resize(cap);
        n_ = used_ = 0;
        for (size_t i = 0; i < ok.size(); ++i)
            if (os[i] == 1) *insert_slot(ok[i]) = std::move(ov[i]);
    }
    V* insert_slot(uint64_t k) {
        size_t m = keys_.size() - 1, i = hash(k) & m;
        while (st_[i] == 1) i = (i + 1) & m;
        if (st_[i] == 0) ++used_;
        st_[i] = 1; keys_[i] = k; ++n_;
        return &vals_[i];
    }
  public:
    size_t size() const { return n_; }
    V* find(uint64_t k) {
        if (keys_.empty()) return nullptr;
        size_t m = keys_.size() - 1, i = hash(k) & m;
        while (st_[i] != 0) {
            if (st_[i] == 1 && keys_[i] == k) return &vals_[i];
            i = (i + 1) & m;
        }
        return nullptr;
    }
    const V* find(uint64_t k) const { return const_cast<FlatMap*>(this)->find(k); }
    V& get_or_create(uint64_t k) {  // get!(constructor, dict, key)
        if (V* v = find(k)) return *v;
        if (keys_.empty() || (used_ + 1) * 4 > keys_.size() * 3) rehash(keys_.empty() ? 16 : keys_.size() * 2);
        V* v = insert_slot(k);
        *v = V();
        return *v;
    }
    bool erase(uint64_t k) {
        if (keys_.empty()) return false;
        size_t m = keys_.size() - 1, i = hash(k) & m;
        while (st_[i] != 0) {
            if (st_[i] == 1 && keys_[i] == k) { st_[i] = 2; vals_[i] = V(); --n_; return true; }
            i = (i + 1) & m;
        }
        return false;
    }
    void clear() { keys_.clear(); st_.clear(); vals_.clear(); n_ = used_ = 0; }
    template <class F> void for_each(F&& f) {  // slot (hash) order, like iterating a Julia Dict
        for (size_t i = 0; i < keys_.size(); ++i) if (st_[i] == 1) f(keys_[i], vals_[i]);
    }
    template <class F> void for_each(F&& f) const {
        for (size_t i = 0; i < keys_.size(); ++i) if (st_[i] == 1) f(keys_[i], vals_[i]);
    }
};

// ---------------------------------------------------------------------------------------------
struct AgentTypeDesc { std::string name; uint32_t size = 0; uint32_t hints = 0; };
struct EdgeTypeDesc { std::string name; uint32_t size = 0; uint32_t hints = 0; int32_t target = 0; uint64_t size_hint = 0; };

// src/Simulation.jl:13-24 AgentReadWrite
struct AgentRW {
    std::vector<uint8_t> state;       // AoS, `size` bytes per slot (Vector{T})
    std::vector<uint64_t> reuseable;  // Vector{AgentNr}
    std::vector<uint8_t> died;        // Vector{Bool} (empty for immortal types)
    size_t nslots = 0;                // length(state) (tracked separately: stateless T has size 0)
};

// src/Simulation.jl:36-67 AgentFields (single-process subset: no shm/foreign caches)
struct AgentFields {
    AgentTypeDesc desc;
    bool immortal = false, independent = false, stateless = false;
    AgentRW read, write;
    uint64_t nextid = 1;
    int64_t last_change = 0;
    bool writeable = false;   // nodes_attrs[:writeable]
    bool prepared = false;    // MPIWindows.prepared
};

// One per-target container C(B) of the table at src/EdgeMethods.jl:9-41.  Which members are
// meaningful depends on the hints:  Vector{Edge{T}}: from+state | Vector{AgentID}: from |
// Vector{T}: state | Int64: count | Edge{T}/AgentID/T (SingleEdge): one entry, count=1 | Bool: count=1
struct Row {
    std::vector<uint64_t> from;
    std::vector<uint8_t> state;
    int64_t count = 0;
};

struct EdgeContainer {
    FlatMap<Row> dict;           // Dict{AgentID, C(B)}  (no :SingleType)
    std::vector<Row> vec;        // Vector{C(B)} indexed by agent_nr (:SingleType)
    std::vector<uint8_t> assigned;  // isassigned() for non-bits C(B)
};

// src/Simulation.jl:69-100 EdgeFields
struct EdgeFields {
    EdgeTypeDesc desc;
    bool stateless = false, ignorefrom = false, singleedge = false, singletype = false;
    std::shared_ptr<EdgeContainer> read, write;
    std::unordered_map<uint64_t, std::vector<uint64_t>> agentsontarget;
    bool readable = false, writeable = false, add_existing = false;
    int64_t last_change = 0;
    bool bits_container() const { return stateless && ignorefrom; }   // C(B) is Int64 or Bool
};

struct Raster {
    std::string name;
    std::vector<int64_t> dims;
    std::vector<AgentID> ids;   // column-major (CartesianIndices order), Raster.jl:42-49
};

struct Sim;
class Ctx;
// bool f(ctx, state*, id): state is in/out (zero-filled when the called type is not in `read`,
// i.e. the Val(T) form, AgentMethods.jl:232-245); return false = `nothing` (the agent dies).
using TransitionFn = std::function<bool(Ctx&, void*, AgentID)>;

struct MapFn {          // a registered map of mapreduce (vb::MapBase): f(element) as double or int64
    uint32_t elem_size = 0;
    bool is_float = false;
    std::function<void(const uint8_t*, double&, int64_t&)> fn;
};
struct Registry {
    std::map<std::pair<std::string, std::string>, TransitionFn> fns;   // (transition, agent type name)
    std::map<std::pair<std::string, std::string>, MapFn> maps;         // (map, agent / edge type name)
    static Registry& get() { static Registry r; return r; }
};

struct Sim {
    std::string name;
    std::vector<AgentFields> agents;   // index = typeid - 1
    std::vector<EdgeFields> edges;
    std::vector<uint8_t> params;
    std::vector<Raster> rasters;
    bool initialized = false, intransition = false;
    int64_t num_transitions = 0;
    bool asserts_enabled = true;    // src/Vahana.jl:42-65
    bool check_readable = true;     // config.check_readable
    bool all_immortal = true;
    uint32_t rank = 0;
    // per-apply context
    uint64_t seed = 0;
    // stats of the last apply (for the CPU baseline)
    uint64_t st_edges_read = 0, st_edges_appended = 0, st_agents_called = 0;

    AgentFields& A(int type) {
        if (type < 1 || type > (int)agents.size()) throw ArgError("unknown agent type id");
        return agents[type - 1];
    }
    EdgeFields& E(int e) {
        if (e < 0 || e >= (int)edges.size()) throw ArgError("unknown edge type index");
        return edges[e];
    }
    void mayassert(bool c, const char* msg) const { if (asserts_enabled && !c) throw AssertionError(msg); }
};

inline std::unique_ptr<Sim> create_sim(const std::string& name, const std::vector<AgentTypeDesc>& ats,
                                       const std::vector<EdgeTypeDesc>& ets, const void* params, uint32_t psize) {
    auto s = std::make_unique<Sim>();
    s->name = name;
    // register_agenttype!: ModelTypes.jl:81-106
    if (ats.size() >= (size_t)vb::MAX_TYPES) throw AssertionError("maximal number of types already registered");
    for (auto& d : ats) {
        AgentFields f;
        f.desc = d;
        f.immortal = d.hints & vb::AGENT_IMMORTAL;
        f.independent = d.hints & vb::AGENT_INDEPENDENT;
        f.stateless = d.size == 0;
        if (!f.immortal) s->all_immortal = false;
        s->agents.push_back(std::move(f));
    }
    // register_edgetype!: ModelTypes.jl:158-230
    for (auto& d : ets) {
        EdgeFields f;
        f.desc = d;
        f.stateless = d.hints & vb::EDGE_STATELESS;
        f.ignorefrom = d.hints & vb::EDGE_IGNORE_FROM;
        f.singleedge = d.hints & vb::EDGE_SINGLE_EDGE;
        f.singletype = d.hints & vb::EDGE_SINGLE_TYPE;
        if (f.singletype && d.target <= 0) throw AssertionError(":SingleType needs the target keyword");
        if (f.singletype && f.singleedge && !(f.stateless && f.ignorefrom))
            throw AssertionError(":SingleEdge and :SingleType can only be combined with :Stateless and :IgnoreFrom");
        f.read = std::make_shared<EdgeContainer>();
        f.write = std::make_shared<EdgeContainer>();   // before init everything goes to `write`; read aliases it after finish_write!
        if (f.singletype && d.size_hint) {   // init_field! after construction (EdgeMethods.jl:198-239, Simulation.jl create_simulation)
            f.write->vec.resize(d.size_hint);
            f.write->assigned.assign(d.size_hint, f.bits_container() ? 1 : 0);
        }
        s->edges.push_back(std::move(f));
    }
    if (psize) s->params.assign((const uint8_t*)params, (const uint8_t*)params + psize);
    return s;
}

// ---------------------------------------------------------------------------------------------
// Edge container helpers (src/EdgeMethods.jl:165-371)

inline void init_field(EdgeFields& f, EdgeContainer& c) {   // init_field!: EdgeMethods.jl:221-239
    if (f.singletype && f.desc.size_hint) {
        c.vec.resize(f.desc.size_hint);
        c.assigned.assign(f.desc.size_hint, f.bits_container() ? 1 : 0);
    }
}
inline void check_size(EdgeFields& f, EdgeContainer& c, uint64_t nr) {   // _check_size!: EdgeMethods.jl:209-240
    if (f.desc.size_hint) return;
    if (c.vec.size() < nr) {
        c.vec.resize(nr);
        c.assigned.resize(nr, f.bits_container() ? 1 : 0);
    }
}
// _get_agent_container (read form): EdgeMethods.jl:303-330,361-370.  nullptr == `nothing`.
inline Row* get_container(Sim& s, EdgeFields& f, EdgeContainer& c, AgentID to) {
    s.mayassert(f.readable || !s.check_readable, "edge type is not in the `read` argument of apply!");
    if (f.singletype) {
        s.mayassert((int)vb::type_nr(to) == f.desc.target, ":SingleType target type mismatch");
        uint64_t nr = vb::agent_nr(to);
        check_size(f, c, nr);
        if (nr == 0 || nr > c.vec.size()) return nullptr;
        return c.assigned[nr - 1] ? &c.vec[nr - 1] : nullptr;
    }
    return c.dict.find(to);
}
// _get_agent_container!: EdgeMethods.jl:331-360
inline Row& get_container_w(Sim& s, EdgeFields& f, EdgeContainer& c, AgentID to) {
    if (f.singletype) {
        s.mayassert((int)vb::type_nr(to) == f.desc.target, ":SingleType target type mismatch");
        uint64_t nr = vb::agent_nr(to);
        check_size(f, c, nr);
        if (nr == 0 || nr > c.vec.size()) throw AssertionError("edge target outside the `size` of the :SingleType container");
        c.assigned[nr - 1] = 1;
        return c.vec[nr - 1];
    }
    return c.dict.get_or_create(to);
}

// _push_agentsontarget: EdgeMethods.jl:375-384
inline void push_agentsontarget(Sim& s, EdgeFields& f, AgentID from, AgentID to) {
    if (f.ignorefrom || s.all_immortal) return;
    uint32_t t = vb::type_nr(from);
    if (t >= 1 && t <= s.agents.size() && !s.agents[t - 1].immortal) f.agentsontarget[from].push_back(to);
}

// add_edge!: EdgeMethods.jl:388-523 (single process: the `storage` branch is never taken)
inline void add_edge(Sim& s, int e, AgentID from, AgentID to, const void* state) {
    EdgeFields& f = s.E(e);
    s.mayassert(!s.initialized || s.intransition, "add_edge! only in the initialization phase or within a transition");
    s.mayassert(!s.check_readable || !s.initialized || f.writeable, "edge type must be in the `write` argument");   // _can_add :258-265
    {   // the reference dereferences the ids (type table / vector index): id 0 is rejected (test/edges.jl:241-248)
        uint32_t tt = vb::type_nr(to);
        s.mayassert(tt >= 1 && tt <= s.agents.size() && vb::agent_nr(to) >= 1, "invalid target id");
        if (!f.ignorefrom) {
            uint32_t ft = vb::type_nr(from);
            s.mayassert(ft >= 1 && ft <= s.agents.size() && vb::agent_nr(from) >= 1, "invalid source id");
        }
    }
    EdgeContainer& c = *f.write;
    const uint32_t sz = f.desc.size;
    if (f.stateless && f.ignorefrom && !f.singleedge) {          // :388-423  count[to] += 1
        Row& r = get_container_w(s, f, c, to);
        r.count += 1;
        return;
    }
    if (f.singleedge) {                                            // :424-476
        if (s.asserts_enabled && !f.singletype && !(f.ignorefrom && f.stateless)) {   // _can_add :267-293
            if (Row* old = c.dict.find(to)) {
                bool same = true;
                if (!f.ignorefrom && (old->from.empty() || old->from[0] != from)) same = false;
                if (!f.stateless && (old->state.size() != sz || std::memcmp(old->state.data(), state, sz) != 0)) same = false;
                if (!same) throw AssertionError("An edge has already been added to this agent (:SingleEdge)");
            }
        }
        push_agentsontarget(s, f, from, to);
        Row& r = get_container_w(s, f, c, to);
        r.count = 1;
        if (!f.ignorefrom) r.from.assign(1, from);
        if (!f.stateless) r.state.assign((const uint8_t*)state, (const uint8_t*)state + sz);
        return;
    }
    push_agentsontarget(s, f, from, to);                           // :477-523  push!(container, value)
    Row& r = get_container_w(s, f, c, to);
    if (!f.ignorefrom) r.from.push_back(from);
    if (!f.stateless) r.state.insert(r.state.end(), (const uint8_t*)state, (const uint8_t*)state + sz);
    r.count += 1;
}

inline void can_remove_edges(Sim& s, EdgeFields& f) {   // _can_remove_edges: EdgeMethods.jl:101-121
    s.mayassert(!s.initialized || s.intransition, "remove_edges! only in the initialization phase or within a transition");
    s.mayassert(!s.check_readable || !s.initialized || f.writeable, "edge type is not in the `write` argument");
    s.mayassert(!s.check_readable || !s.initialized || f.add_existing, "edge type must be in the `add_existing` keyword");
}
// remove_edges!(sim, to, T): EdgeMethods.jl:527-550
inline void remove_edges_to(Sim& s, int e, AgentID to) {
    EdgeFields& f = s.E(e);
    can_remove_edges(s, f);
    EdgeContainer& c = *f.write;
    if (f.singletype) {
        uint64_t nr = vb::agent_nr(to);
        check_size(f, c, nr);
        if (nr >= 1 && nr <= c.vec.size()) { c.vec[nr - 1] = Row(); c.assigned[nr - 1] = 1; }   // field[nr] = zero(CT)
    } else {
        c.dict.erase(to);
    }
}
// remove_edges!(sim, from, to, T): EdgeMethods.jl:552-599
inline void remove_edges_from_to(Sim& s, int e, AgentID from, AgentID to) {
    EdgeFields& f = s.E(e);
    if (f.ignorefrom) throw AssertionError("remove_edges! with a source agent is not defined for :IgnoreFrom edge types");
    can_remove_edges(s, f);
    EdgeContainer& c = *f.write;
    Row* r = nullptr;
    if (f.singletype) {
        uint64_t nr = vb::agent_nr(to);
        if (nr >= 1 && nr <= c.vec.size() && c.assigned[nr - 1]) r = &c.vec[nr - 1];
    } else {
        r = c.dict.find(to);
    }
    if (!r) return;
    if (f.singleedge) {
        if (!r->from.empty() && r->from[0] == from) remove_edges_to(s, e, to);
        return;
    }
    if (r->from.empty()) return;
    const uint32_t sz = f.desc.size;
    size_t w = 0;
    for (size_t i = 0; i < r->from.size(); ++i) {       // filter(e -> e.from != from, row): the key stays (A-18 quirk)
        if (r->from[i] == from) continue;
        if (w != i) {
            r->from[w] = r->from[i];
            if (!f.stateless) std::memmove(&r->state[w * sz], &r->state[i * sz], sz);
        }
        ++w;
    }
    r->from.resize(w);
    if (!f.stateless) r->state.resize(w * sz);
    r->count = (int64_t)w;
}

// _remove_edges_agent_target!: EdgeMethods.jl:897-920
inline bool remove_edges_agent_target(Sim& s, EdgeFields& f, AgentID to) {
    EdgeContainer& c = *f.write;
    if (f.singletype) {
        if ((int)vb::type_nr(to) != f.desc.target) return false;
        uint64_t nr = vb::agent_nr(to);
        if (c.vec.size() >= nr) { c.vec[nr - 1] = Row(); c.assigned[nr - 1] = 1; return true; }
        return false;
    }
    return c.dict.erase(to);
}
// _remove_edges_agent_source!: EdgeMethods.jl:606-631
inline bool remove_edges_agent_source(Sim& s, int e, const std::vector<AgentID>& from) {
    EdgeFields& f = s.E(e);
    if (f.ignorefrom) return false;
    bool changed = false;
    bool cr = s.check_readable;
    s.check_readable = false;
    for (AgentID fr : from) {
        auto it = f.agentsontarget.find(fr);
        if (it == f.agentsontarget.end()) continue;
        changed = true;
        std::vector<AgentID> targets = it->second;
        for (AgentID t : targets) remove_edges_from_to(s, e, fr, t);
        f.agentsontarget.erase(fr);
    }
    s.check_readable = cr;
    return changed;
}

// _num_edges: EdgeMethods.jl:931-969
inline uint64_t num_edges_total(Sim& s, int e, bool write) {
    EdgeFields& f = s.E(e);
    const EdgeContainer& c = write ? *f.write : *f.read;
    uint64_t n = 0;
    if (f.singletype) {
        for (size_t i = 0; i < c.vec.size(); ++i) if (c.assigned[i]) n += (uint64_t)c.vec[i].count;
    } else {
        c.dict.for_each([&](uint64_t, const Row& r) { n += (uint64_t)r.count; });
    }
    return n;
}

// ---------------------------------------------------------------------------------------------
// Agents (src/AgentMethods.jl)

// _get_next_id: AgentMethods.jl:37-63
inline uint64_t get_next_id(AgentFields& a) {
    if (!a.immortal && !a.read.reuseable.empty()) {
        uint64_t nr = a.read.reuseable.back();    // pop!: take from the END
        a.read.reuseable.pop_back();
        a.write.died[nr - 1] = 0;
        return nr;
    }
    uint64_t nr = a.nextid;
    a.nextid = nr + 1;
    if (nr > a.write.nslots) {
        a.write.nslots = nr;
        a.write.state.resize(nr * a.desc.size);
        if (!a.immortal) a.write.died.resize(nr);
    }
    if (!a.immortal) a.write.died[nr - 1] = 0;
    return nr;
}
// add_agent!: AgentMethods.jl:65-89
inline AgentID add_agent(Sim& s, int type, const void* state) {
    AgentFields& a = s.A(type);
    s.mayassert(!s.initialized || s.intransition, "add_agent! only in the initialization phase or within a transition");
    s.mayassert(!s.initialized || a.writeable, "agent type must be in the `write` argument");
    uint64_t nr = get_next_id(a);
    const uint32_t sz = a.desc.size;
    if (sz) {
        if (a.independent && nr <= a.read.nslots) std::memcpy(&a.read.state[(nr - 1) * sz], state, sz);
        else std::memcpy(&a.write.state[(nr - 1) * sz], state, sz);
    }
    return vb::agent_id((uint32_t)type, s.rank, nr);
}
// agentstate (nompi branch): AgentMethods.jl:91-124
inline const void* agentstate(Sim& s, AgentID id, int type) {
    AgentFields& a = s.A(type);
    s.mayassert((int)vb::type_nr(id) == type, "The id of the agent does not match the given type");
    s.mayassert(a.prepared || !s.check_readable, "agent type must be in the `read` argument of the transition function");
    uint64_t nr = vb::agent_nr(id);
    if (nr < 1 || nr > a.read.nslots) throw AssertionError("agentstate: agent does not exist");   // Julia: BoundsError
    if (!a.immortal) s.mayassert(!a.read.died[nr - 1], "agentstate was requested for an agent that has been removed");
    static const uint64_t zero8 = 0;
    if (a.stateless) return &zero8;
    return &a.read.state[(nr - 1) * a.desc.size];
}

// prepare_write! (agents): AgentMethods.jl:271-298
inline void prepare_write_agent(Sim& s, int type, bool add_existing) {
    AgentFields& a = s.A(type);
    if (a.immortal && s.initialized && !add_existing)
        throw AssertionError("an :Immortal type in `write` must also be in `add_existing` (or `call`)");
    if (!add_existing) {
        a.read = AgentRW();
        a.write = AgentRW();
        a.nextid = 1;
    }
    a.writeable = true;
}

// finish_write! (agents), single process: AgentMethods.jl:300-439
inline void finish_write_agent(Sim& s, int type) {
    AgentFields& a = s.A(type);
    bool must_copy = a.write.nslots > a.read.nslots || !s.initialized || !a.independent;   // :306-308
    if (!a.immortal) {                                                                      // :314-358
        std::vector<AgentID> aids;
        for (uint64_t nr : a.write.reuseable) aids.push_back(vb::agent_id((uint32_t)type, s.rank, nr));
        for (AgentID id : aids)
            for (size_t e = 0; e < s.edges.size(); ++e)
                if (remove_edges_agent_target(s, s.edges[e], id)) s.edges[e].last_change = s.num_transitions;
        if (!aids.empty())
            for (size_t e = 0; e < s.edges.size(); ++e)
                if (remove_edges_agent_source(s, (int)e, aids)) s.edges[e].last_change = s.num_transitions;
    }
    if (!a.stateless && must_copy) {                                                        // :360-382
        if (a.independent && s.initialized)
            std::memcpy(a.write.state.data(), a.read.state.data(), a.read.nslots * a.desc.size);
        a.read.state = a.write.state;
        a.read.nslots = a.write.nslots;
    } else if (a.stateless) {                                                               // :383-390
        a.read.nslots = a.write.nslots;
    }
    if (!a.immortal) a.read.died = a.write.died;                                            // :398-413
    a.read.reuseable.insert(a.read.reuseable.end(), a.write.reuseable.begin(), a.write.reuseable.end());   // :430
    a.write.reuseable.clear();
    a.last_change = s.num_transitions;
    a.writeable = false;
}

// num_agents: Agent.jl:324-343
inline uint64_t num_agents(Sim& s, int type) {
    AgentFields& a = s.A(type);
    if (a.immortal) return a.nextid - 1;
    const std::vector<uint8_t>& d = s.initialized ? a.read.died : a.write.died;
    uint64_t n = 0;
    for (uint8_t x : d) n += !x;
    return n;
}

// prepare_write! / finish_write! (edges): EdgeMethods.jl:639-684
inline void prepare_write_edge(Sim& s, int e, bool in_read, bool add_existing) {
    EdgeFields& f = s.E(e);
    if (add_existing) {
        if (in_read) f.write = std::make_shared<EdgeContainer>(*f.read);   // deepcopy
        else f.write = f.read;                                               // alias
    } else {
        f.write = std::make_shared<EdgeContainer>();
        init_field(f, *f.write);
    }
    f.writeable = true;
    f.add_existing = add_existing;
}
inline void finish_write_edge(Sim& s, int e) {
    EdgeFields& f = s.E(e);
    f.read = f.write;
    f.last_change = s.num_transitions;
    f.writeable = false;
}

// finish_init! (single process): Simulation.jl:403-476
inline void finish_init(Sim& s) {
    if (s.initialized) throw AssertionError("You can not call finish_init! twice for the same simulation");
    for (int rep = 0; rep < 2; ++rep) {
        for (size_t t = 1; t <= s.agents.size(); ++t) finish_write_agent(s, (int)t);
        for (size_t e = 0; e < s.edges.size(); ++e) finish_write_edge(s, (int)e);
    }
    s.initialized = true;
    s.num_transitions = 1;
}

// ---------------------------------------------------------------------------------------------
// Raster (src/Raster.jl)

// _stencil_core: Raster.jl:82-96.  Offsets in Iterators.product order (first dim fastest).
inline std::vector<std::vector<int64_t>> stencil(int metric, int n, double distance) {
    int64_t d = (int64_t)std::floor(distance);
    std::vector<std::vector<int64_t>> out;
    std::vector<int64_t> cur(n, -d);
    if (d < 0) return out;
    while (true) {
        bool zero = true;
        double n2 = 0; int64_t n1 = 0;
        for (int i = 0; i < n; ++i) { zero &= cur[i] == 0; n2 += (double)(cur[i] * cur[i]); n1 += std::llabs(cur[i]); }
        bool keep = !zero;
        if (keep && metric == vb::EUCLIDEAN) keep = std::sqrt(n2) <= distance;
        if (keep && metric == vb::MANHATTEN) keep = (double)n1 <= distance;
        if (keep) out.push_back(cur);
        int i = 0;
        while (i < n && ++cur[i] > d) { cur[i] = -d; ++i; }
        if (i == n) break;
    }
    return out;
}
// _checkpos: Raster.jl:479-499.  Returns false when the position is dropped.
inline bool checkpos(std::vector<int64_t>& pos, const std::vector<int64_t>& dims, bool periodic) {
    bool oob = false;
    for (size_t i = 0; i < dims.size(); ++i) {
        if (pos[i] < 1 || pos[i] > dims[i]) {
            oob = true;
            int64_t m = (pos[i] - 1) % dims[i];
            if (m < 0) m += dims[i];
            pos[i] = m + 1;    // mod1
        }
    }
    return !oob || periodic;
}
inline size_t linear_index(const std::vector<int64_t>& pos, const std::vector<int64_t>& dims) {
    size_t idx = 0, stride = 1;
    for (size_t i = 0; i < dims.size(); ++i) { idx += (size_t)(pos[i] - 1) * stride; stride *= (size_t)dims[i]; }
    return idx;
}
inline Raster& find_raster(Sim& s, const std::string& name) {
    for (auto& r : s.rasters) if (r.name == name) return r;
    throw ArgError("unknown raster " + name);
}
// add_raster!: Raster.jl:32-54 (the agent constructor has already been evaluated by the caller,
// `states` is in CartesianIndices order)
inline Raster& add_raster(Sim& s, const std::string& name, const std::vector<int64_t>& dims, int type, const void* states,
                          AgentID* ids_out) {
    if (s.initialized) throw AssertionError("add_raster! can be only called before finish_init!");
    Raster r;
    r.name = name; r.dims = dims;
    size_t n = 1;
    for (int64_t d : dims) n *= (size_t)d;
    r.ids.resize(n);
    const uint32_t sz = s.A(type).desc.size;
    for (size_t i = 0; i < n; ++i) {
        r.ids[i] = add_agent(s, type, sz ? (const uint8_t*)states + i * sz : nullptr);
        if (ids_out) ids_out[i] = r.ids[i];
    }
    s.rasters.push_back(std::move(r));
    return s.rasters.back();
}
// connect_raster_neighbors!: Raster.jl:139-167 (edge from `org` to `shifted`)
inline void connect_raster_neighbors(Sim& s, const std::string& name, int e, double distance, int metric, bool periodic,
                                     const void* edge_state) {
    Raster& r = find_raster(s, name);
    auto st = stencil(metric, (int)r.dims.size(), distance);
    std::vector<int64_t> org(r.dims.size(), 1), sh(r.dims.size());
    for (size_t i = 0; i < r.ids.size(); ++i) {
        for (auto& o : st) {
            for (size_t k = 0; k < org.size(); ++k) sh[k] = org[k] + o[k];
            if (checkpos(sh, r.dims, periodic)) add_edge(s, e, r.ids[i], r.ids[linear_index(sh, r.dims)], edge_state);
        }
        size_t k = 0;
        while (k < org.size() && ++org[k] > r.dims[k]) { org[k] = 1; ++k; }
    }
}
// move_to!: Raster.jl:437-477
inline void move_to(Sim& s, Raster& r, AgentID id, const int64_t* posv, int e_from, const void* s_from, int e_to,
                    const void* s_to, double distance, int metric, bool periodic, bool only_surrounding) {
    std::vector<int64_t> pos(posv, posv + r.dims.size());
    if (!only_surrounding) {
        for (size_t k = 0; k < pos.size(); ++k)
            if (pos[k] < 1 || pos[k] > r.dims[k]) throw AssertionError("move_to!: position outside the raster");
        AgentID cell = r.ids[linear_index(pos, r.dims)];
        if (e_from >= 0) add_edge(s, e_from, cell, id, s_from);
        if (e_to >= 0) add_edge(s, e_to, id, cell, s_to);
    }
    if (distance >= 1) {
        std::vector<int64_t> sh(pos.size());
        for (auto& o : stencil(metric, (int)r.dims.size(), distance)) {
            for (size_t k = 0; k < pos.size(); ++k) sh[k] = pos[k] + o[k];
            if (!checkpos(sh, r.dims, periodic)) continue;
            AgentID cell = r.ids[linear_index(sh, r.dims)];
            if (e_from >= 0) add_edge(s, e_from, cell, id, s_from);
            if (e_to >= 0) add_edge(s, e_to, id, cell, s_to);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// The transition-side context: the sequential implementation of the Ctx concept documented in
// include/vahana_device.cuh.  Accessors follow src/EdgeMethods.jl:704-892.
class Ctx {
  public:
    Sim& s;
    uint64_t slot = 0;   // 0-based slot of the agent being called (keys the uniform table)
    explicit Ctx(Sim& sim) : s(sim) {}

    template <class P> const P& param() const { return *reinterpret_cast<const P*>(s.params.data()); }
    double uniform(int k) const { return vb::Philox::uniform(s.seed, slot, (uint64_t)k); }
    void require(bool cond) const { if (!cond) throw AssertionError("an assertion inside the transition function failed (ctx.require)"); }   // @assert in a closure
    // cooperative-group surface: the oracle is a group of one
    int lanes() const { return 1; }
    int lane() const { return 0; }
    bool leader() const { return true; }
    template <class T> T sum(T v) const { return v; }
    template <class T> T max(T v) const { return v; }
    template <class T> T min(T v) const { return v; }
    template <class A, class M> A reduce_acc(const A& a, M&&) const { return a; }   // vb::ReduceTransition: one lane, nothing to merge

    static void avail(bool ok, const char* what) { if (!ok) throw AssertionError(std::string(what) + " is not defined for this hint combination"); }

    int64_t num_edges(int e, AgentID id) {                          // EdgeMethods.jl:850-869
        EdgeFields& f = s.E(e);
        avail(!f.singleedge, "num_edges");
        Row* r = get_container(s, f, *f.read, id);
        if (r) s.st_edges_read += 1;
        return r ? r->count : 0;
    }
    bool has_edge(int e, AgentID id) {                              // EdgeMethods.jl:872-892
        EdgeFields& f = s.E(e);
        if (f.ignorefrom && f.stateless) { Row* r = get_container(s, f, *f.read, id); return r && r->count != 0; }
        if (!f.singleedge) return num_edges(e, id) >= 1;
        avail(!f.singletype, "has_edge");
        s.mayassert(f.readable || !s.check_readable, "edge type is not in the `read` argument of apply!");
        return f.read->dict.find(id) != nullptr;
    }
    template <class F> void for_each_neighbor(int e, AgentID id, F&& fn) {   // neighborids(_iter): :717-761
        EdgeFields& f = s.E(e);
        avail(!f.ignorefrom, "neighborids");
        Row* r = get_container(s, f, *f.read, id);
        if (!r) return;
        s.st_edges_read += r->from.size();
        for (size_t i = 0; i < r->from.size(); ++i) fn(r->from[i]);
    }
    AgentID neighbor_at(int e, AgentID id, int64_t k) {             // neighborids(...)[k+1]
        EdgeFields& f = s.E(e);
        avail(!f.ignorefrom, "neighborids");
        Row* r = get_container(s, f, *f.read, id);
        if (!r || k < 0 || (size_t)k >= r->from.size()) throw AssertionError("neighbor_at: index out of range");
        return r->from[(size_t)k];
    }
    template <class S, class F> void for_each_edge(int e, AgentID id, F&& fn) {   // edges: :704-715
        EdgeFields& f = s.E(e);
        avail(!f.stateless && !f.ignorefrom, "edges");
        Row* r = get_container(s, f, *f.read, id);
        if (!r) return;
        s.st_edges_read += r->from.size();
        for (size_t i = 0; i < r->from.size(); ++i) { S st; std::memcpy(&st, &r->state[i * sizeof(S)], sizeof(S)); fn(r->from[i], st); }
    }
    template <class S, class F> void for_each_edgestate(int e, AgentID id, F&& fn) {   // edgestates(_iter): :804-848
        EdgeFields& f = s.E(e);
        avail(!f.stateless, "edgestates");
        Row* r = get_container(s, f, *f.read, id);
        if (!r) return;
        size_t n = r->state.size() / sizeof(S);
        s.st_edges_read += n;
        for (size_t i = 0; i < n; ++i) { S st; std::memcpy(&st, &r->state[i * sizeof(S)], sizeof(S)); fn(st); }
    }
    template <class A> A agentstate(int type, AgentID id) {         // AgentMethods.jl:91-154
        A a{};
        const void* p = vo::agentstate(s, id, type);
        if (sizeof(A) > 0 && !s.A(type).stateless) std::memcpy(&a, p, sizeof(A));
        return a;
    }
    template <class Fd> Fd agentfield(int type, AgentID id, int offset) {   // agentstate_flexible(sim,id).field
        Fd v{};
        const uint8_t* p = (const uint8_t*)vo::agentstate(s, id, type);
        if (offset + sizeof(Fd) > s.A(type).desc.size) throw AssertionError("agentfield: field outside the agent state");
        std::memcpy(&v, p + offset, sizeof(Fd));
        return v;
    }
    template <class A, class F> void for_each_neighborstate(int e, int type, AgentID id, F&& fn) {   // neighborstates: :764-802
        for_each_neighbor(e, id, [&](AgentID from) { fn(agentstate<A>(type, from)); });
    }
    void add_edge(int e, AgentID from, AgentID to) { vo::add_edge(s, e, from, to, nullptr); s.st_edges_appended++; }
    template <class S> void add_edge(int e, AgentID from, AgentID to, const S& st) { vo::add_edge(s, e, from, to, &st); s.st_edges_appended++; }
    template <class A> AgentID add_agent(int type, const A& a) { return vo::add_agent(s, type, &a); }
    void remove_edges(int e, AgentID to) { remove_edges_to(s, e, to); }
    void remove_edges(int e, AgentID from, AgentID to) { remove_edges_from_to(s, e, from, to); }
    AgentID cellid(int raster, const vb::Pos& p) {                  // Raster.jl:403-405
        Raster& r = s.rasters.at((size_t)raster);
        std::vector<int64_t> pos(p.v, p.v + r.dims.size());
        for (size_t k = 0; k < pos.size(); ++k)
            if (pos[k] < 1 || pos[k] > r.dims[k]) throw AssertionError("cellid: position outside the raster");
        return r.ids[linear_index(pos, r.dims)];
    }
    // move_to! with stateless edges (e_from/e_to = -1 for `nothing`)
    void move_to(int raster, AgentID id, const vb::Pos& p, int e_from, int e_to, double distance = 0,
                 int metric = vb::CHEBYSHEV, bool periodic = true, bool only_surrounding = false) {
        Raster& r = s.rasters.at((size_t)raster);
        size_t before = 0;
        (void)before;
        vo::move_to(s, r, id, p.v, e_from, nullptr, e_to, nullptr, distance, metric, periodic, only_surrounding);
    }
};

// ---------------------------------------------------------------------------------------------
// apply!: src/Simulation.jl:720-821 with the transition loops of AgentMethods.jl:159-268
inline bool contains(const std::vector<int>& v, int x) { return std::find(v.begin(), v.end(), x) != v.end(); }

inline void transition_write(Sim& s, int type, uint64_t idx, bool alive, const void* newstate, bool in_write) {
    AgentFields& a = s.A(type);
    if (!in_write) return;                                         // transition_without_write!: :183-191
    const uint32_t sz = a.desc.size;
    if (a.immortal) {                                              // transition_with_write!: :159-181
        s.mayassert(alive, "You can not return `nothing` for immortal agents");
        if (!sz) return;
        if (a.independent) std::memcpy(&a.read.state[(idx - 1) * sz], newstate, sz);
        else std::memcpy(&a.write.state[(idx - 1) * sz], newstate, sz);
    } else if (!alive) {
        a.write.reuseable.push_back(idx);
        a.write.died[idx - 1] = 1;
    } else if (sz) {
        if (a.independent) std::memcpy(&a.read.state[(idx - 1) * sz], newstate, sz);
        else std::memcpy(&a.write.state[(idx - 1) * sz], newstate, sz);
    }
}

inline void apply(Sim& s, const std::string& transition, const std::vector<int>& call, const std::vector<int>& read,
                  const std::vector<int>& write, const std::vector<int>& add_existing, int with_edge, uint64_t seed) {
    if (!s.initialized) throw AssertionError("You must call finish_init! before apply!");
    if (with_edge >= 0 && s.E(with_edge).singletype)
        throw AssertionError("The `with_edge` keyword can only be used for edgetypes without the :SingleType hint");
    for (int c : call) if (contains(add_existing, c)) throw AssertionError("a `call` type can not be element of `add_existing`");
    for (int ae : add_existing) if (!contains(write, ae)) throw AssertionError("type is in `add_existing` but not in `write`");
    // resolve the functor for every called type before touching any state
    std::vector<TransitionFn*> fns;
    for (int c : call) {
        if (c >= vb::EDGE_REF) throw ArgError("`call` must list agent types");
        auto it = Registry::get().fns.find({transition, s.A(c).desc.name});
        if (it == Registry::get().fns.end())
            throw ArgError("transition '" + transition + "' is not registered for agent type " + s.A(c).desc.name);
        fns.push_back(&it->second);
    }
    s.intransition = true;
    s.seed = seed;
    s.st_edges_read = s.st_edges_appended = s.st_agents_called = 0;
    struct Reset { Sim& s; ~Reset() { s.intransition = false; } } reset{s};   // the reference leaves it set after a throw (test/core.jl:295-296)

    for (int r : read) { if (r >= vb::EDGE_REF) s.E(r - vb::EDGE_REF).readable = true; }     // prepare_read! edges: EdgeMethods.jl:665
    for (int r : read) { if (r < vb::EDGE_REF) s.A(r).prepared = true; }                     // prepare_read! agents: AgentMethods.jl:484-501
    struct Unread {
        Sim& s; const std::vector<int>& read;
        ~Unread() { for (int r : read) { if (r >= vb::EDGE_REF) s.edges[r - vb::EDGE_REF].readable = false; else s.agents[r - 1].prepared = false; } }
    } unread{s, read};
    for (int w : write)                                                                       // Simulation.jl:764-768
        if (w >= vb::EDGE_REF && !contains(add_existing, w)) s.E(w - vb::EDGE_REF).agentsontarget.clear();
    for (int w : write) {                                                                     // :770
        bool ae = contains(call, w) || contains(add_existing, w);
        if (w >= vb::EDGE_REF) prepare_write_edge(s, w - vb::EDGE_REF, contains(read, w), ae);
        else prepare_write_agent(s, w, ae);
    }
    Ctx ctx(s);
    std::vector<uint8_t> tmp;
    for (size_t ci = 0; ci < call.size(); ++ci) {                                             // :774-788
        const int C = call[ci];
        AgentFields& a = s.A(C);
        const bool in_write = contains(write, C), in_read = contains(read, C);
        const uint32_t sz = a.desc.size;
        tmp.assign(sz ? sz : 1, 0);
        auto call_one = [&](uint64_t idx) {
            AgentID id = vb::agent_id((uint32_t)C, s.rank, idx);
            if (in_read && sz) std::memcpy(tmp.data(), &a.read.state[(idx - 1) * sz], sz);
            else std::memset(tmp.data(), 0, tmp.size());           // Val(T) form: no state is handed over
            ctx.slot = idx - 1;
            s.st_agents_called++;
            bool alive = (*fns[ci])(ctx, tmp.data(), id);
            transition_write(s, C, idx, alive, tmp.data(), in_write);
        };
        if (with_edge < 0) {                                       // transition_with(out)_read!: AgentMethods.jl:193-207,232-245
            const uint64_t n = a.read.nslots;                      // slots appended during the loop are not visited (A-5)
            for (uint64_t idx = 1; idx <= n; ++idx) {
                if (!a.immortal && a.read.died[idx - 1]) continue;
                call_one(idx);
            }
        } else {                                                   // ..._with_edge!: AgentMethods.jl:209-229,247-268
            std::vector<AgentID> keys;
            s.E(with_edge).read->dict.for_each([&](uint64_t k, const Row&) { if ((int)vb::type_nr(k) == C) keys.push_back(k); });
            for (AgentID k : keys) {
                uint64_t idx = vb::agent_nr(k);
                if (!a.immortal) s.mayassert(!a.read.died[idx - 1], "with_edge: agent has been removed");
                call_one(idx);
            }
        }
    }
    // single process: transmit_remove_edges!/transmit_edges! (:795-800) have nothing to exchange
    for (int w : write) if (w < vb::EDGE_REF) finish_write_agent(s, w);                       // :807
    for (int w : write) if (w >= vb::EDGE_REF) finish_write_edge(s, w - vb::EDGE_REF);        // :809
    s.num_transitions += 1;                                                                   // :816
}

// ---------------------------------------------------------------------------------------------
// mapreduce: AgentMethods.jl:533-565, EdgeMethods.jl:972-994; identities Helpers.jl:44-81.
// The map is "field at byte offset (optionally == cmp), or the constant 1": see include/vahana_b200.h.
struct Value { int dt = vb::DT_I64; int64_t i = 0; double f = 0; };

inline Value load_value(const uint8_t* p, int offset, int dt, bool has_cmp, int64_t cmp, int result_dt) {
    Value v;
    v.dt = result_dt;
    int64_t iv = 0; double fv = 0;
    bool isf = false;
    if (dt < 0) iv = 1;
    else switch (dt) {
        case vb::DT_I64: { int64_t x; std::memcpy(&x, p + offset, 8); iv = x; break; }
        case vb::DT_F64: { double x; std::memcpy(&x, p + offset, 8); fv = x; isf = true; break; }
        case vb::DT_BOOL: case vb::DT_U8: iv = p[offset]; break;
        case vb::DT_I32: { int32_t x; std::memcpy(&x, p + offset, 4); iv = x; break; }
        case vb::DT_F32: { float x; std::memcpy(&x, p + offset, 4); fv = x; isf = true; break; }
        default: throw ArgError("bad dtype");
    }
    if (has_cmp) { iv = isf ? (fv == (double)cmp) : (iv == cmp); isf = false; }
    if (result_dt == vb::DT_F64 || result_dt == vb::DT_F32) v.f = isf ? fv : (double)iv;
    else v.i = isf ? (int64_t)fv : iv;
    return v;
}
inline Value identity(int op, int result_dt) {   // val4empty: Helpers.jl:44-81
    Value v; v.dt = result_dt;
    const bool isf = result_dt == vb::DT_F64 || result_dt == vb::DT_F32;
    const bool isb = result_dt == vb::DT_BOOL;
    switch (op) {
        case vb::OP_SUM: break;
        case vb::OP_PROD: v.i = 1; v.f = 1; break;
        case vb::OP_MAX: if (isf) v.f = -INFINITY; else v.i = isb ? -1 : -INT64_MAX; break;
        case vb::OP_MIN: if (isf) v.f = INFINITY; else v.i = isb ? 1 : INT64_MAX; break;
        case vb::OP_AND: if (isf) throw AssertionError("& is only supported for integer and boolean types"); v.i = isb ? 1 : INT64_MAX; break;
        case vb::OP_OR: if (isf) throw AssertionError("| is only supported for integer and boolean types"); v.i = 0; break;
        default: throw AssertionError("Can not derive the init value for the operator");
    }
    return v;
}
inline void fold(Value& acc, const Value& x, int op) {   // reduced = op(f(x), reduced)
    const bool isf = acc.dt == vb::DT_F64 || acc.dt == vb::DT_F32;
    switch (op) {
        case vb::OP_SUM: if (isf) acc.f = x.f + acc.f; else acc.i = (int64_t)((uint64_t)x.i + (uint64_t)acc.i); break;
        case vb::OP_PROD: if (isf) acc.f = x.f * acc.f; else acc.i = (int64_t)((uint64_t)x.i * (uint64_t)acc.i); break;
        case vb::OP_MIN: if (isf) acc.f = std::fmin(x.f, acc.f); else acc.i = std::min(x.i, acc.i); break;
        case vb::OP_MAX: if (isf) acc.f = std::fmax(x.f, acc.f); else acc.i = std::max(x.i, acc.i); break;
        case vb::OP_AND: acc.i = x.i & acc.i; break;
        case vb::OP_OR: acc.i = x.i | acc.i; break;
    }
}
inline Value mapreduce(Sim& s, int type_ref, int offset, int dt, bool has_cmp, int64_t cmp, int op, int result_dt,
                       const Value* init, const MapFn* mf = nullptr) {
    if (s.intransition) throw AssertionError("You can not call mapreduce inside of a transition function.");
    Value acc = init ? *init : identity(op, result_dt);
    acc.dt = result_dt;
    const bool risf = result_dt == vb::DT_F64 || result_dt == vb::DT_F32;
    if (mf && mf->is_float != risf) throw ArgError("the result datatype does not match the map functor");
    auto load_value = [&](const uint8_t* p, int offset_, int dt_, bool has_cmp_, int64_t cmp_, int result_dt_) -> Value {
        if (!mf) return vo::load_value(p, offset_, dt_, has_cmp_, cmp_, result_dt_);
        Value v; v.dt = result_dt_;
        mf->fn(p, v.f, v.i);                                   // f(state): the closure of mapreduce(sim, f, op, T)
        return v;
    };
    if (type_ref < vb::EDGE_REF) {
        AgentFields& a = s.A(type_ref);
        const uint32_t sz = a.desc.size;
        if (mf && mf->elem_size != sz) throw ArgError("sizeof(Elem) of the map functor does not match the agent type");
        static const uint8_t zeros[64] = {0};
        for (uint64_t i = 1; i < a.nextid; ++i) {
            if (!a.immortal && a.read.died[i - 1]) continue;
            Value x = load_value(sz ? &a.read.state[(i - 1) * sz] : zeros, offset, dt, has_cmp, cmp, result_dt);
            fold(acc, x, op);
        }
    } else {
        EdgeFields& f = s.E(type_ref - vb::EDGE_REF);
        if (f.stateless) throw AssertionError("mapreduce is not defined for :Stateless edge types");
        const uint32_t sz = f.desc.size;
        if (mf && mf->elem_size != sz) throw ArgError("sizeof(Elem) of the map functor does not match the edge type");
        auto row = [&](const Row& r) {
            for (size_t i = 0; i * sz < r.state.size(); ++i) { Value x = load_value(&r.state[i * sz], offset, dt, has_cmp, cmp, result_dt); fold(acc, x, op); }
        };
        if (f.singletype) { for (size_t i = 0; i < f.read->vec.size(); ++i) if (f.read->assigned[i]) row(f.read->vec[i]); }
        else f.read->dict.for_each([&](uint64_t, const Row& r) { row(r); });
    }
    return acc;
}

inline std::unique_ptr<Sim> copy_sim(const Sim& s) {   // copy_simulation: Simulation.jl:500-510 (deepcopy)
    auto c = std::make_unique<Sim>(s);
    for (auto& f : c->edges) {
        bool alias = f.read == f.write;
        f.read = std::make_shared<EdgeContainer>(*f.read);
        f.write = alias ? f.read : std::make_shared<EdgeContainer>(*f.write);
    }
    return c;
}

}  // namespace vo

// Registers a single-source map functor (vb::MapBase) with the oracle.
#define VO_REGISTER_MAP(mname, type_name, ...)                                                                 \
    static const bool VO_CAT(vo_regmap_, __COUNTER__) = [] {                                                   \
        using VoMap = __VA_ARGS__;                                                                             \
        vo::MapFn m;                                                                                           \
        m.elem_size = (uint32_t)sizeof(typename VoMap::Elem);                                                  \
        m.is_float = std::is_floating_point<typename VoMap::Result>::value;                                    \
        m.fn = [](const uint8_t* p, double& f, int64_t& i) {                                                   \
            typename VoMap::Elem e;                                                                            \
            std::memcpy((void*)&e, p, sizeof(e));                                                              \
            const typename VoMap::Result r = VoMap()(e);                                                       \
            if (std::is_floating_point<typename VoMap::Result>::value) f = (double)r; else i = (int64_t)r;     \
        };                                                                                                     \
        vo::Registry::get().maps[{mname, type_name}] = m;                                                      \
        return true;                                                                                           \
    }();
#define VO_CAT2(a, b) a##b
#define VO_CAT(a, b) VO_CAT2(a, b)
// Registers Functor (a single-source transition from vahana.jl_b200/csrc/transitions) with the oracle.
#define VO_REGISTER_TRANSITION(tname, agenttype_name, ...)                                                     \
    static const bool VO_CAT(vo_reg_, __COUNTER__) = [] {                                                      \
        vo::Registry::get().fns[{tname, agenttype_name}] = [](vo::Ctx& ctx, void* st, vb::AgentID id) -> bool { \
            using VoFunctor = __VA_ARGS__;                                                                     \
            typename VoFunctor::State local;                                                                   \
            std::memcpy((void*)&local, st, sizeof(local));                                                     \
            bool alive = VoFunctor()(ctx, local, id);                                                          \
            std::memcpy(st, (void*)&local, sizeof(local));                                                     \
            return alive;                                                                                      \
        };                                                                                                     \
        return true;                                                                                           \
    }();
