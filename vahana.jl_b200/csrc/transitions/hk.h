// hk.h — Hegselmann–Krause opinion dynamics, the reference's docs example
// (/root/reference/docs/examples/hegselmann.jl:27-45 model, :134-144 transition `step`).
//
// Single-source transition: instantiated with the CUDA context (vahana_device.cuh) by hk.cu
// and with the sequential oracle context by oracle/oracle_models.cpp.
#pragma once
#include <math.h>
#include "../../../include/vahana_model.h"

namespace hk {

struct HKAgent { double opinion; };     // hegselmann.jl:27-29
struct Params { double eps; };          // register_param!(:ϵ, 0.02), hegselmann.jl:44
enum : int { T_HKAGENT = 1 };           // agent type ids in registration order
enum : int { E_KNOWS = 0 };             // edge types in registration order (struct Knows end, no hints)

// step(agent, id, sim): hegselmann.jl:134-144
//   opinions = map(a -> a.opinion, neighborstates(sim, id, Knows, HKAgent))
//   accepted = filter(o -> abs(o - agent.opinion) < ϵ, opinions);  HKAgent(mean(accepted))
// Written as a reduce transition (include/vahana_model.h): fold = the filter + running sum, finish = the mean.  The oracle folds
// the row left to right like the reference's `filter`/`mean`; the GPU combines per-lane or per-source-block partial sums.
//
// Prefilter (include/vahana_model.h): key = the opinion quantised to 1/256 (clamped, monotone, non-expanding), so
// |o_s - o_t| < eps  ==>  |key_s - key_t| <= floor(256 eps) + 1.  At eps = 0.02 that rules out 95 % of the neighbours on one byte.
struct OpinionKey {
    struct Probe { int32_t lo, hi; };
    static VB_HD uint32_t quant(double o) { const double q = o * 256.0; return q >= 255.0 ? 255u : (q > 0.0 ? (uint32_t)q : 0u); }   // NaN -> 0
    static VB_HD Probe probe(double own, double eps) {
        const double w = floor(eps * 256.0) + 1.0;
        const int32_t band = w >= 256.0 ? 256 : (w >= 1.0 ? (int32_t)w : 1);       // eps <= 0 or NaN: fold accepts nothing anyway
        const int32_t k = (int32_t)quant(own);
        Probe p; p.lo = k - band; p.hi = k + band;
        return p;
    }
    static VB_HD bool may_accept(const Probe& p, uint32_t key) { return (int32_t)key >= p.lo && (int32_t)key <= p.hi; }
};
struct Step : vb::ReduceTransition<Step> {
    using State = HKAgent;
    using Source = HKAgent;
    struct Acc { double sum; uint32_t n; };
    static constexpr int kAccBytes = 12;
    static constexpr int kPrimaryEdge = E_KNOWS;
    static constexpr int kSourceType = T_HKAGENT;
    static constexpr bool kPrefilter = true;
    using Probe = OpinionKey::Probe;
    template <class Ctx> VB_HD uint8_t key(const Ctx&, const HKAgent& nb) const { return (uint8_t)OpinionKey::quant(nb.opinion); }
    template <class Ctx> VB_HD Probe probe(const Ctx& ctx, const HKAgent& self) const { return OpinionKey::probe(self.opinion, ctx.template param<Params>().eps); }
    VB_HD bool may_accept(const Probe& p, uint32_t key) const { return OpinionKey::may_accept(p, key); }
    template <class Ctx> VB_HD void init(const Ctx&, const HKAgent&, Acc& a) const { a.sum = 0.0; a.n = 0; }
    template <class Ctx> VB_HD void fold(const Ctx& ctx, const HKAgent& self, const HKAgent& nb, Acc& a) const {
        if (fabs(nb.opinion - self.opinion) < ctx.template param<Params>().eps) { a.sum += nb.opinion; a.n += 1; }
    }
    VB_HD void merge(Acc& a, const Acc& b) const { a.sum += b.sum; a.n += b.n; }
    template <class Ctx> VB_HD bool finish(const Ctx&, HKAgent& self, vb::AgentID, const Acc& a) const {
        self.opinion = a.sum / (double)a.n;   // mean(accepted): NaN for an empty set, as in Julia
        return true;
    }
};

// Test variant with mortality (not in the reference's example): an agent whose accepted neighbourhood averages below `eps * 10`
// returns `nothing`.  Exercises died rows, the dead-agent edge purge and the rebuild of the source-blocked view next to the sweeps.
struct StepOrDie : vb::ReduceTransition<StepOrDie> {
    using State = HKAgent;
    using Source = HKAgent;
    using Acc = Step::Acc;
    static constexpr int kAccBytes = 12;
    static constexpr int kPrimaryEdge = E_KNOWS;
    static constexpr int kSourceType = T_HKAGENT;
    static constexpr bool kPrefilter = true;
    using Probe = OpinionKey::Probe;
    template <class Ctx> VB_HD uint8_t key(const Ctx& c, const HKAgent& nb) const { return Step().key(c, nb); }
    template <class Ctx> VB_HD Probe probe(const Ctx& c, const HKAgent& self) const { return Step().probe(c, self); }
    VB_HD bool may_accept(const Probe& p, uint32_t key) const { return OpinionKey::may_accept(p, key); }
    template <class Ctx> VB_HD void init(const Ctx& c, const HKAgent& s, Acc& a) const { Step().init(c, s, a); }
    template <class Ctx> VB_HD void fold(const Ctx& c, const HKAgent& s, const HKAgent& nb, Acc& a) const { Step().fold(c, s, nb, a); }
    VB_HD void merge(Acc& a, const Acc& b) const { Step().merge(a, b); }
    template <class Ctx> VB_HD bool finish(const Ctx& ctx, HKAgent& self, vb::AgentID, const Acc& a) const {
        if (a.n == 0) return false;                          // every neighbour (and the self loop) is gone or out of range
        self.opinion = a.sum / (double)a.n;
        return self.opinion >= ctx.template param<Params>().eps * 10.0;
    }
};

}  // namespace hk
