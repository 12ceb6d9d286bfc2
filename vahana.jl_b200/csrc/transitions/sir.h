// sir.h — Episim-style SIR contact model (BASELINE config 5).  Episim itself lives in an external repository
// (/root/reference/README.md:60); SURVEY.md Appendix C defines this stand-in with the same structure: a bipartite
// person x location network whose two edge types are rebuilt every step.
//     struct Person state::UInt8; days::UInt8 end   (0 S, 1 I, 2 R)      struct Location n_inf::Int32 end
//     struct Visit infectious::Bool end   (:SingleType target Location, :IgnoreSourceState)
//     struct Exposure risk::Float32 end   (:IgnoreFrom, :SingleType target Person)
//     apply!(sim, visit,  Person,   [Person],           [Visit])       2 x add_edge!(id, loc(u_k), Visit(state == 1))
//     apply!(sim, tally,  Location, [Visit],            [Location])    n_inf = count(e.infectious for e in edgestates)
//     apply!(sim, expose, Location, [Location, Visit],  [Exposure])    per visitor: Exposure(n_inf / n_visitors)
//     apply!(sim, infect, Person,   [Person, Exposure], [Person])      S->I if u < 1 - exp(-beta * sum(risk)); I->R after 10 days
// Uniforms come from the per-agent table ctx.uniform(k) (identical draws in oracle and kernels).
#pragma once
#include <math.h>
#include "../../../include/vahana_model.h"

namespace sir {

struct Person { uint8_t state; uint8_t days; };
struct Location { int32_t n_inf; };
struct Visit { bool infectious; };
struct Exposure { float risk; };
struct Params { double beta; int64_t n_locations; int64_t visits_per_step; int64_t infectious_days; int64_t n_ranks; };
enum : int { T_PERSON = 1, T_LOCATION = 2 };
enum : int { E_VISIT = 0, E_EXPOSURE = 1 };

struct DoVisit : vb::TransitionBase {
    using State = Person;
    using EdgeWrites = vb::IntList<E_VISIT>;
    template <class Ctx>
    VB_HD bool operator()(Ctx& ctx, Person& p, vb::AgentID id) const {
        const Params& pr = ctx.template param<Params>();
        const Visit v{p.state == 1};
        for (int k = 0; k < (int)pr.visits_per_step; ++k) {
            int64_t loc = (int64_t)(ctx.uniform(k) * (double)pr.n_locations);
            if (loc >= pr.n_locations) loc = pr.n_locations - 1;
            // locations are spread over the ranks in contiguous equal blocks (one rank: rank 0, nr = loc + 1)
            ctx.add_edge(E_VISIT, id, vb::block_partition_id(T_LOCATION, (uint64_t)loc, (uint64_t)pr.n_locations, (uint32_t)(pr.n_ranks > 0 ? pr.n_ranks : 1)), v);
        }
        return true;
    }
};
struct Tally : vb::TransitionBase {   // called without Location in `read` (Val form): the state is rebuilt from the edges
    using State = Location;
    template <class Ctx>
    VB_HD bool operator()(Ctx& ctx, Location& l, vb::AgentID id) const {
        int32_t n = 0;
        ctx.template for_each_edgestate<Visit>(E_VISIT, id, [&](const Visit& v) { n += v.infectious ? 1 : 0; });
        l.n_inf = n;
        return true;
    }
};
struct Expose : vb::TransitionBase {
    using State = Location;
    using EdgeWrites = vb::IntList<E_EXPOSURE>;
    template <class Ctx>
    VB_HD bool operator()(Ctx& ctx, Location& l, vb::AgentID id) const {
        const long long nv = ctx.num_edges(E_VISIT, id);
        if (nv == 0) return true;
        const Exposure e{(float)l.n_inf / (float)nv};
        ctx.for_each_neighbor(E_VISIT, id, [&](vb::AgentID visitor) { ctx.add_edge(E_EXPOSURE, id, visitor, e); });
        return true;
    }
};
struct Infect : vb::TransitionBase {
    using State = Person;
    template <class Ctx>
    VB_HD bool operator()(Ctx& ctx, Person& p, vb::AgentID id) const {
        const Params& pr = ctx.template param<Params>();
        if (p.state == 0) {
            float risk = 0.f;
            ctx.template for_each_edgestate<Exposure>(E_EXPOSURE, id, [&](const Exposure& e) { risk += e.risk; });
            const double pinf = 1.0 - exp(-pr.beta * (double)risk);
            if (ctx.uniform(0) < pinf) { p.state = 1; p.days = 0; }
        } else if (p.state == 1) {
            p.days = (uint8_t)(p.days + 1);
            if (p.days >= pr.infectious_days) p.state = 2;
        }
        return true;
    }
};

}  // namespace sir
