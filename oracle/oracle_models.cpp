// oracle_models.cpp — CPU ORACLE (test infrastructure): instantiates the single-source
// transition functors of vahana.jl_b200/csrc/transitions with the sequential oracle context.
#include <cmath>
#include "vahana_oracle.hpp"
#include "../vahana.jl_b200/csrc/transitions/all.h"

#define VB_TRANSITION(tname, atype, ...) VO_REGISTER_TRANSITION(tname, atype, __VA_ARGS__)
#define VB_MAP(mname, tname, ...) VO_REGISTER_MAP(mname, tname, __VA_ARGS__)
#include "../vahana.jl_b200/csrc/transitions/registry.inc"
#undef VB_MAP
#undef VB_TRANSITION

// host generator of the synthetic HK power-law workload (same formulas as the device generator)
#include "../include/vahana_workloads.h"
#include "../vahana.jl_b200/csrc/workloads/hk_powerlaw.h"
extern "C" int vbw_hk_powerlaw_host(uint64_t n, int agent_type, uint64_t seed_graph, uint64_t seed_opinion, double c, uint32_t dmax,
                                    vb_agent_id* from_out, vb_agent_id* to_out, double* opinions_out, uint64_t* n_edges_out) {
    return vbw::hk_powerlaw_host(n, agent_type, seed_graph, seed_opinion, c, dmax, from_out, to_out, opinions_out, n_edges_out);
}
extern "C" int vbw_hk_powerlaw_build_sharded(vb_sim*, int, int, uint64_t, uint64_t, uint64_t, double, uint32_t, uint64_t, uint32_t, uint32_t, uint64_t*) { return VB_ERR_STATE; }
extern "C" int vbw_hk_powerlaw_build(vb_sim*, int, int, uint64_t, uint64_t, uint64_t, double, uint32_t, uint64_t, uint64_t*) { return VB_ERR_STATE; }

// Contract check of a functor's prefilter (include/vahana_model.h) on the CPU: for every pair (self[i], nb[i]) whose fold changes the
// accumulator, may_accept(probe(self), key(nb)) must hold.  Returns the number of violations; *accepted_out / *kept_out count the
// pairs the fold accepted and the pairs the key let through (selectivity).  tests/test_hk.py drives it with adversarial pairs.
namespace {
struct EpsCtx {
    hk::Params p;
    template <class P> const P& param() const { return *reinterpret_cast<const P*>(&p); }
};
}  // namespace
extern "C" uint64_t vbt_hk_prefilter_violations(const double* self, const double* nb, uint64_t n, double eps, uint64_t* accepted_out, uint64_t* kept_out) {
    EpsCtx ctx; ctx.p.eps = eps;
    const hk::Step f{};
    uint64_t bad = 0, acc_n = 0, kept = 0;
    for (uint64_t i = 0; i < n; ++i) {
        hk::HKAgent s{self[i]}, b{nb[i]};
        hk::Step::Acc a; f.init(ctx, s, a);
        f.fold(ctx, s, b, a);
        const bool accepted = a.n != 0;
        const bool may = f.may_accept(f.probe(ctx, s), (uint32_t)f.key(ctx, b));
        acc_n += accepted; kept += may;
        if (accepted && !may) ++bad;
    }
    if (accepted_out) *accepted_out = acc_n;
    if (kept_out) *kept_out = kept;
    return bad;
}
