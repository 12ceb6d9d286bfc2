// testkit.h — the transition closures of the reference's own test-suite, restated as
// single-source functors so that tests/ can replay test/core.jl, test/edges.jl,
// test/edgesiterator.jl, test/remove_agents.jl, test/addexisting.jl, test/independent.jl,
// test/raster.jl and test/graphs.jl against both the oracle and the CUDA engine.
// Each functor cites the closure it restates (paths relative to /root/reference).
#pragma once
#include "../../../include/vahana_model.h"

namespace testkit {

struct Foo { int64_t foo; };                 // AMortal, AImm, ... (test/core.jl:9-13), Agent (test/edges.jl:13)
struct FooBool { int64_t foo; bool b; };     // ADefault (test/core.jl:14-17)
struct EFoo { int64_t foo; };                // ESDict / EdgeD... (test/core.jl:24, test/edges.jl:15-30)

// do state,_,_ -> nothing   (test/core.jl:144-146,150-152)
template <class S> struct KillAll : vb::TransitionBase {
    using State = S;
    static constexpr bool kCooperative = false;
    template <class Ctx> VB_HD bool operator()(Ctx&, S&, vb::AgentID) const { return false; }
};
// identity / no-op closures (test/core.jl:101-103 on rank 0, test/edges.jl:349-350,377-378)
template <class S> struct Identity : vb::TransitionBase {
    using State = S;
    static constexpr bool kCooperative = false;
    template <class Ctx> VB_HD bool operator()(Ctx&, S&, vb::AgentID) const { return true; }
};
// state.foo < 6 ? state : nothing   (test/core.jl:223-229)
struct KeepFooLt6 : vb::TransitionBase {
    using State = Foo;
    static constexpr bool kCooperative = false;
    template <class Ctx> VB_HD bool operator()(Ctx&, Foo& s, vb::AgentID) const { return s.foo < 6; }
};
// state.foo % 2 == 0 ? state : nothing   (test/core.jl:256-258)
struct KeepEvenFoo : vb::TransitionBase {
    using State = Foo;
    static constexpr bool kCooperative = false;
    template <class Ctx> VB_HD bool operator()(Ctx&, Foo& s, vb::AgentID) const { return s.foo % 2 == 0; }
};
// create_sum_state_neighbors(edgetype): sum of n.foo over neighborstates_flexible (test/core.jl:71-83).
// `foo` is the first field of every agent type of the model, so the flexible lookup is a field read
// at offset 0 of whatever type the neighbour id carries.
template <int E> struct SumStateNeighbors : vb::TransitionBase {
    using State = Foo;
    static constexpr bool kCooperative = true;
    template <class Ctx> VB_HD bool operator()(Ctx& ctx, Foo& self, vb::AgentID id) const {
        int64_t s = 0;
        ctx.for_each_neighbor(E, id, [&](vb::AgentID from) {
            s += ctx.template agentfield<int64_t>((int)vb::type_nr(from), from, 0);
        });
        self.foo = ctx.sum(s);
        return true;
    }
};
// ADefault(state.foo, false)   (test/core.jl:425-427)
struct SetBoolFalse : vb::TransitionBase {
    using State = FooBool;
    static constexpr bool kCooperative = false;
    template <class Ctx> VB_HD bool operator()(Ctx&, FooBool& s, vb::AgentID) const { s.b = false; return true; }
};
// ADefault(state.foo, mod(id, 2) == 1)   (test/core.jl:433-435)
struct SetBoolIdOdd : vb::TransitionBase {
    using State = FooBool;
    static constexpr bool kCooperative = false;
    template <class Ctx> VB_HD bool operator()(Ctx&, FooBool& s, vb::AgentID id) const { s.b = (id % 2) == 1; return true; }
};
// @test num_edges(sim, id, ET) == n  inside the closure (test/edges.jl:338-346): the count is
// stored in the agent so the host can assert on it.
template <int E> struct StoreNumEdges : vb::TransitionBase {
    using State = Foo;
    static constexpr bool kCooperative = false;
    template <class Ctx> VB_HD bool operator()(Ctx& ctx, Foo& s, vb::AgentID id) const { s.foo = ctx.num_edges(E, id); return true; }
};
// add_edges!(sim, id, edges(sim, id, ET))   (test/edges.jl:357-359,367-369)
template <int E> struct ReaddEdges : vb::TransitionBase {
    using State = Foo;
    using EdgeWrites = vb::IntList<E>;
    static constexpr bool kCooperative = false;
    template <class Ctx> VB_HD bool operator()(Ctx& ctx, Foo&, vb::AgentID id) const {
        ctx.template for_each_edge<EFoo>(E, id, [&](vb::AgentID from, const EFoo& st) { ctx.add_edge(E, from, id, st); });
        return true;
    }
};
// add_edge!(sim, id, id, t())  (test/edges.jl:253-266): used for the read/write permission checks
template <int E, bool kStateful> struct AddSelfLoop : vb::TransitionBase {
    using State = Foo;
    using EdgeWrites = vb::IntList<E>;
    static constexpr bool kCooperative = false;
    template <class Ctx> VB_HD bool operator()(Ctx& ctx, Foo&, vb::AgentID id) const {
        if (kStateful) ctx.add_edge(E, id, id, EFoo{0});
        else ctx.add_edge(E, id, id);
        return true;
    }
};

// has_edge(sim, id, t) inside a closure: the read-permission check of test/edges.jl:268-272
template <int E> struct TouchEdge : vb::TransitionBase {
    using State = Foo;
    template <class Ctx> VB_HD bool operator()(Ctx& ctx, Foo&, vb::AgentID id) const { (void)ctx.has_edge(E, id); return true; }
};

// ---- test/mpi/test_agentstate.jl: the state of an edge's source (an agent of another rank under mpiexec) stays fresh ----
// model: Agent{state} (:Immortal in the first half, mortal in the second), EdgeState{state} and NewEdge{state}, both :SingleEdge
// apply!(sim, [Agent], [Agent, E], []) do _, id, sim; e = edges(sim, id, E); isnothing(e) || @test e.state.state * MUL == agentstate(sim, e.from, Agent).state
template <int E, int MUL> struct CheckSourceState : vb::TransitionBase {            // :53-60, :69-75, :97-103
    using State = Foo;
    static constexpr bool kCooperative = false;
    template <class Ctx> VB_HD bool operator()(Ctx& ctx, Foo&, vb::AgentID id) const {
        if (!ctx.has_edge(E, id)) return true;
        ctx.template for_each_edge<EFoo>(E, id, [&](vb::AgentID from, const EFoo& e) {
            const Foo a = ctx.template agentstate<Foo>((int)vb::type_nr(from), from);
            ctx.require(e.foo * MUL == a.foo);
        });
        return true;
    }
};
template <int MUL, int DIV> struct ScaleState : vb::TransitionBase {                 // Agent(state.state * 2) :64-66, Agent(state.state / 2) :79-81
    using State = Foo;
    static constexpr bool kCooperative = false;
    template <class Ctx> VB_HD bool operator()(Ctx&, Foo& s, vb::AgentID) const { s.foo = s.foo * MUL / DIV; return true; }
};
template <int E> struct RequireNoEdge : vb::TransitionBase {                         // @test isnothing(edges(sim, id, NewEdge)) :83-86
    using State = Foo;
    static constexpr bool kCooperative = false;
    template <class Ctx> VB_HD bool operator()(Ctx& ctx, Foo&, vb::AgentID id) const { ctx.require(!ctx.has_edge(E, id)); return true; }
};
template <int FROM_E, int TO_E> struct CopyEdges : vb::TransitionBase {              // add_edge!(sim, e.from, id, NewEdge(e.state.state)) :89-94
    using State = Foo;
    using EdgeWrites = vb::IntList<TO_E>;
    static constexpr bool kCooperative = false;
    template <class Ctx> VB_HD bool operator()(Ctx& ctx, Foo&, vb::AgentID id) const {
        if (!ctx.has_edge(FROM_E, id)) return true;
        ctx.template for_each_edge<EFoo>(FROM_E, id, [&](vb::AgentID from, const EFoo& e) { ctx.add_edge(TO_E, from, id, EFoo{e.foo}); });
        return true;
    }
};

// ---- remove_edges! inside transitions (test/mpi/test_edgetypes.jl:295-449, single process) ----
template <int E> struct RemoveOwnIfEven : vb::TransitionBase {          // :311-317
    using State = Foo;
    using EdgeRemoves = vb::IntList<E>;
    template <class Ctx> VB_HD bool operator()(Ctx& ctx, Foo& s, vb::AgentID id) const { if (s.foo % 2 == 0) ctx.remove_edges(E, id); return true; }
};
template <int E> struct RemoveNeighborsIfEven : vb::TransitionBase {    // :323-330
    using State = Foo;
    using EdgeRemoves = vb::IntList<E>;
    template <class Ctx> VB_HD bool operator()(Ctx& ctx, Foo& s, vb::AgentID id) const {
        if (s.foo % 2 == 0) ctx.remove_edges(E, ctx.neighbor_at(E, id, 0));
        return true;
    }
};
template <int E, bool kStateful> struct RemoveAndReadd : vb::TransitionBase {   // :337-346: remove + add of the same edge keeps it
    using State = Foo;
    using EdgeRemoves = vb::IntList<E>;
    using EdgeWrites = vb::IntList<E>;
    template <class Ctx> VB_HD bool operator()(Ctx& ctx, Foo&, vb::AgentID id) const {
        const vb::AgentID from = ctx.neighbor_at(E, id, 0);
        ctx.remove_edges(E, from, id);
        if (kStateful) ctx.add_edge(E, from, id, EFoo{0}); else ctx.add_edge(E, from, id);
        return true;
    }
};
// The order rule of a removal that names another agent's row (src/Simulation.jl:792-800: removes travel and are applied before the
// new edges arrive): every agent clears the row of its neighbour and then points an edge back at it.  On the cycle i-1 -> i every row
// ends up holding exactly the reversed edge, on one rank (program order) and on several (the removal request reaches the row's rank
// before the new edge does).
template <int E, bool kStateful> struct ClearNeighborRowAndPointBack : vb::TransitionBase {
    using State = Foo;
    using EdgeRemoves = vb::IntList<E>;
    using EdgeWrites = vb::IntList<E>;
    template <class Ctx> VB_HD bool operator()(Ctx& ctx, Foo&, vb::AgentID id) const {
        const vb::AgentID nid = ctx.neighbor_at(E, id, 0);
        ctx.remove_edges(E, nid);
        if (kStateful) ctx.add_edge(E, id, nid, EFoo{7}); else ctx.add_edge(E, id, nid);
        return true;
    }
};
template <int E, bool kSingle> struct RemoveFirstTwoFrom : vb::TransitionBase {   // :385-392
    using State = Foo;
    using EdgeRemoves = vb::IntList<E>;
    template <class Ctx> VB_HD bool operator()(Ctx& ctx, Foo&, vb::AgentID id) const {
        ctx.remove_edges(E, ctx.neighbor_at(E, id, 0), id);
        if (!kSingle) ctx.remove_edges(E, ctx.neighbor_at(E, id, 1), id);
        return true;
    }
};
template <int E> struct RemoveToRandomNeighborIfEven : vb::TransitionBase {   // :399-406: remove_edges!(sim, id, nid, ET), nid = rand(neighborids)
    using State = Foo;
    using EdgeRemoves = vb::IntList<E>;
    template <class Ctx> VB_HD bool operator()(Ctx& ctx, Foo& s, vb::AgentID id) const {
        if (s.foo % 2 == 0) {
            const long long n = ctx.num_edges(E, id);
            long long k = (long long)(ctx.uniform(0) * (double)n);
            if (k >= n) k = n - 1;
            ctx.remove_edges(E, id, ctx.neighbor_at(E, id, k));
        }
        return true;
    }
};
template <int E> struct RemoveFromZero : vb::TransitionBase {            // :428-430
    using State = Foo;
    using EdgeRemoves = vb::IntList<E>;
    template <class Ctx> VB_HD bool operator()(Ctx& ctx, Foo&, vb::AgentID id) const { ctx.remove_edges(E, (vb::AgentID)0, id); return true; }
};

// ---- test/remove_agents.jl ----
struct Empty {};
struct Idx { int64_t idx; };
// num_edges(sim, id, E) == 0 ? nothing : state   (test/remove_agents.jl:52-58)
template <int E> struct DieIfNoEdges : vb::TransitionBase {
    using State = Idx;
    template <class Ctx> VB_HD bool operator()(Ctx& ctx, Idx&, vb::AgentID id) const { return ctx.num_edges(E, id) != 0; }
};

// ---- test/addexisting.jl: ComputeAgent = 1, ConstructedAgent = 2, Connection = 0 ----
struct ConstructAndConnect : vb::TransitionBase {   // :47-50
    using State = Empty;
    using EdgeWrites = vb::IntList<0>;
    using AgentWrites = vb::IntList<2>;
    template <class Ctx> VB_HD bool operator()(Ctx& ctx, Empty&, vb::AgentID id) const {
        const vb::AgentID c = ctx.add_agent(2, Empty{});
        ctx.add_edge(0, c, id);
        return true;
    }
};

// ---- test/independent.jl: agent types 1..3 (Foo), AFooEdge = 0 {foo}, AEdge = 1 ----
template <int K> struct IndepStep : vb::TransitionBase {   // :63-75, :113-125
    using State = Foo;
    using EdgeWrites = vb::IntList<0>;
    template <class Ctx> VB_HD bool operator()(Ctx& ctx, Foo& s, vb::AgentID id) const {
        if (s.foo == K) return false;
        const EFoo st{s.foo};
        ctx.for_each_neighbor(1, id, [&](vb::AgentID nid) {
            ctx.add_edge(0, nid, id, st);
            ctx.add_edge(0, id, nid, st);
        });
        return true;
    }
};
template <int T> struct IndepSpawn : vb::TransitionBase {   // :164-176
    using State = Foo;
    using EdgeWrites = vb::IntList<0>;
    using AgentWrites = vb::IntList<T>;
    template <class Ctx> VB_HD bool operator()(Ctx& ctx, Foo& s, vb::AgentID id) const {
        vb::AgentID nid = ctx.add_agent(T, Foo{s.foo + 10});
        ctx.add_edge(0, nid, id, EFoo{(int64_t)nid});
        ctx.add_edge(0, id, nid, EFoo{(int64_t)nid});
        nid = ctx.add_agent(T, Foo{s.foo + 20});
        ctx.add_edge(0, nid, id, EFoo{(int64_t)nid});
        ctx.add_edge(0, id, nid, EFoo{(int64_t)nid});
        return true;
    }
};

// ---- test/graphs.jl: GraphA = 1 {id, sum_ids_neighbors}, GraphE = 0 ----
struct GraphA { int64_t id; int64_t sum_ids_neighbors; };
// :19-23  `sum(map(a -> a.id, neighborstates(sim, id, GraphE, GraphA)))` as a reduce transition: integer sums are exact in any
// association, so the source-blocked read phase must reproduce the reference's golden value bit for bit
struct SumIds : vb::ReduceTransition<SumIds> {
    using State = GraphA;
    using Source = GraphA;
    struct Acc { int64_t s; };
    static constexpr int kAccBytes = 8;
    static constexpr int kPrimaryEdge = 0;
    static constexpr int kSourceType = 1;
    template <class Ctx> VB_HD void init(const Ctx&, const GraphA&, Acc& a) const { a.s = 0; }
    template <class Ctx> VB_HD void fold(const Ctx&, const GraphA&, const GraphA& nb, Acc& a) const { a.s += nb.id; }
    VB_HD void merge(Acc& a, const Acc& b) const { a.s += b.s; }
    template <class Ctx> VB_HD bool finish(const Ctx&, GraphA& self, vb::AgentID, const Acc& a) const { self.sum_ids_neighbors = a.s; return true; }
};

// ---- test/raster.jl: GridA = 1, Grid3D = 2, Position = 3, MovingAgent = 4; GridE = 0, OnPosition = 1 ----
struct GridA { int64_t pos[2]; bool active; };
struct Grid3D { int64_t pos[3]; bool active; };
struct Position { int64_t ids_sum; };
struct MovingAgent { int64_t value; };
// a.active || any(neighbour.active)   (:25-28, :107-110)
template <class G, int T> struct Diffuse : vb::TransitionBase {
    using State = G;
    template <class Ctx> VB_HD bool operator()(Ctx& ctx, G& a, vb::AgentID id) const {
        bool any = a.active;
        ctx.template for_each_neighborstate<G>(0, T, id, [&](const G& n) { any = any || n.active; });
        a.active = any;
        return true;
    }
};
// probes the per-target row order of a raster edge type: pos := (nr of the k-th neighbour, row length), k = (x + y) mod length
struct GridProbe : vb::TransitionBase {
    using State = GridA;
    template <class Ctx> VB_HD bool operator()(Ctx& ctx, GridA& a, vb::AgentID id) const {
        const int64_t n = ctx.num_edges(0, id);
        if (n > 0) {
            const vb::AgentID nb = ctx.neighbor_at(0, id, (a.pos[0] + a.pos[1]) % n);
            a.pos[0] = (int64_t)vb::agent_nr(nb);
        }
        a.pos[1] = n;
        return true;
    }
};
struct SumOnPos : vb::TransitionBase {   // :243-250 (Val{Position} form)
    using State = Position;
    template <class Ctx> VB_HD bool operator()(Ctx& ctx, Position& p, vb::AgentID id) const {
        int64_t s = 0;
        ctx.for_each_neighbor(1, id, [&](vb::AgentID from) { s += ctx.template agentfield<int64_t>((int)vb::type_nr(from), from, 0); });
        p.ids_sum = s;
        return true;
    }
};
struct ValueOnPos : vb::TransitionBase {   // :251-255 (Val{MovingAgent} form)
    using State = MovingAgent;
    using EdgeWrites = vb::IntList<1>;
    template <class Ctx> VB_HD bool operator()(Ctx& ctx, MovingAgent& m, vb::AgentID id) const {
        const vb::AgentID first = ctx.neighbor_at(1, id, 0);
        const int64_t value = ctx.template agentfield<int64_t>((int)vb::type_nr(first), first, 0);
        vb::Pos p{{value, value, 0, 0}};
        ctx.move_to(0, id, p, 1, 1);
        m.value = value;
        return true;
    }
};

}  // namespace testkit
