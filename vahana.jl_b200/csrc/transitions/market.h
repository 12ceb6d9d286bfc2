// market.h — the market model of the reference's first tutorial (docs/examples/tutorial1.jl, "Excess Demand"): buyers with
// Cobb-Douglas preferences buy two goods from one randomly chosen seller they know; sellers move the price of good y towards
// balanced demand.
//     struct Buyer α::Float64; B::Float64 end              struct Seller p::Float64; d_y::Float64 end       (tutorial1.jl:101-112)
//     struct KnownSeller end   (seller -> buyer, fixed)    struct Bought x::Float64; y::Float64 end  (buyer -> seller, per step)  (:143-158)
//     apply!(sim, calc_demand, Buyer,  [Buyer, Seller, KnownSeller], Bought)                              (:528, :576)
//     apply!(sim, calc_price,  Seller, [Seller, Bought],             Seller)                              (:532, :578)
// `rand(neighborids(...))` (:409) becomes the k-th neighbour with k = floor(u * n), u = ctx.uniform(0): the per-agent uniform table
// that oracle and kernels share.  Both functors are sequential (one lane per agent): calc_price adds the Bought states of a row
// left to right, as reduce(+, edgestates(...)) does (:433), so Float64 results are bit-exact between oracle and kernels.
#pragma once
#include "../../../include/vahana_model.h"

namespace market {

struct Buyer { double alpha; double B; };
struct Seller { double p; double d_y; };
struct Bought { double x; double y; };
enum : int { T_BUYER = 1, T_SELLER = 2 };
enum : int { E_KNOWN_SELLER = 0, E_BOUGHT = 1 };

struct CalcDemand : vb::TransitionBase {   // tutorial1.jl:408-414
    using State = Buyer;
    using EdgeWrites = vb::IntList<E_BOUGHT>;
    template <class Ctx>
    VB_HD bool operator()(Ctx& ctx, Buyer& b, vb::AgentID id) const {
        const long long n = ctx.num_edges(E_KNOWN_SELLER, id);
        if (n == 0) return true;            // the tutorial gives every buyer at least one seller (rand of an empty vector would throw)
        long long k = (long long)(ctx.uniform(0) * (double)n);
        if (k >= n) k = n - 1;
        const vb::AgentID seller = ctx.neighbor_at(E_KNOWN_SELLER, id, k);
        const Seller s = ctx.template agentstate<Seller>(T_SELLER, seller);
        const Bought q{b.B * b.alpha, b.B * (1.0 - b.alpha) / s.p};
        ctx.add_edge(E_BOUGHT, id, seller, q);
        return true;
    }
};
struct CalcPrice : vb::TransitionBase {    // tutorial1.jl:427-435
    using State = Seller;
    template <class Ctx>
    VB_HD bool operator()(Ctx& ctx, Seller& s, vb::AgentID id) const {
        if (!ctx.has_edge(E_BOUGHT, id)) return true;      // isnothing(edgestates(...)): the state stays
        bool first = true;
        Bought q{0.0, 0.0};
        ctx.template for_each_edgestate<Bought>(E_BOUGHT, id, [&](const Bought& e) {
            if (first) { q = e; first = false; } else { q.x = q.x + e.x; q.y = q.y + e.y; }     // reduce(+, sold) starts from the first element
        });
        s = Seller{q.y / q.x * s.p, q.y};
        return true;
    }
};

// the two map closures of the tutorial that are more than a field selector (registered maps of mapreduce, vb::MapBase)
struct XMinusY : vb::MapBase {             // mapreduce(sim, b -> b.x - b.y, +, Bought)   tutorial1.jl:548,577
    using Elem = Bought;
    using Result = double;
    VB_HD double operator()(const Bought& b) const { return b.x - b.y; }
};
struct Revenue : vb::MapBase {             // mapreduce(sim, s -> s.p * s.d_y, +, Seller)  tutorial1.jl:566
    using Elem = Seller;
    using Result = double;
    VB_HD double operator()(const Seller& s) const { return s.p * s.d_y; }
};
struct HasCustomers : vb::MapBase {        // an integral map: mapreduce(sim, s -> s.d_y > 0, +, Seller)
    using Elem = Seller;
    using Result = int64_t;
    VB_HD int64_t operator()(const Seller& s) const { return s.d_y > 0 ? 1 : 0; }
};

}  // namespace market
