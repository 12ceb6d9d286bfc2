// hk.h — Hegselmann–Krause opinion dynamics, the reference's docs example
// (/root/reference/docs/examples/hegselmann.jl:27-45 model, :134-144 transition `step`).
//
// Single-source transition: instantiated with the CUDA context (vahana_device.cuh) by hk.cu
// and with the sequential oracle context by oracle/oracle_models.cpp.
#pragma once
#include "../../../include/vahana_model.h"

namespace hk {

struct HKAgent { double opinion; };     // hegselmann.jl:27-29
struct Params { double eps; };          // register_param!(:ϵ, 0.02), hegselmann.jl:44
enum : int { T_HKAGENT = 1 };           // agent type ids in registration order
enum : int { E_KNOWS = 0 };             // edge types in registration order (struct Knows end, no hints)

// step(agent, id, sim): hegselmann.jl:134-144
//   opinions = map(a -> a.opinion, neighborstates(sim, id, Knows, HKAgent))
//   accepted = filter(o -> abs(o - agent.opinion) < ϵ, opinions);  HKAgent(mean(accepted))
// The neighbour walk is cooperative: each lane of the agent's group folds a strided share of the
// row, ctx.sum() combines the lanes (group of 1 in the oracle => strict left-to-right order).
struct Step : vb::TransitionBase {
    using State = HKAgent;
    static constexpr bool kCooperative = true;
    static constexpr int kPrimaryEdge = E_KNOWS;
    template <class Ctx>
    VB_HD bool operator()(Ctx& ctx, HKAgent& self, vb::AgentID id) const {
        const double eps = ctx.template param<Params>().eps;
        const double own = self.opinion;
        double acc = 0.0;
        long long n = 0;
        ctx.template for_each_neighborstate<HKAgent>(E_KNOWS, T_HKAGENT, id, [&](const HKAgent& nb) {
            if (fabs(nb.opinion - own) < eps) { acc += nb.opinion; n += 1; }
        });
        acc = ctx.sum(acc);
        n = ctx.sum(n);
        self.opinion = acc / (double)n;   // mean(accepted): NaN for an empty set, as in Julia
        return true;
    }
};

}  // namespace hk
