// hk_powerlaw.h — formulas of the synthetic Hegselmann–Krause power-law workload (include/vahana_workloads.h),
// shared by the device generator, the host generator and the oracle build.
#pragma once
#include <math.h>
#include "../../../include/vahana_model.h"

namespace vbw {

VB_HD uint32_t hk_degree(uint64_t seed, uint64_t i, double c, uint32_t dmax) {
    const double u = vb::Philox::uniform(seed, i, 0);
    const double x = c * pow(1.0 - u, -2.0 / 3.0);
    return x >= (double)dmax ? dmax : (uint32_t)x;
}
VB_HD uint64_t hk_source(uint64_t seed, uint64_t k, uint64_t n) {
    const double v = vb::Philox::uniform(seed, k, 1);
    uint64_t s = (uint64_t)((double)n * v * v);
    return s >= n ? n - 1 : s;
}
VB_HD double hk_opinion(uint64_t seed, uint64_t i) { return vb::Philox::uniform(seed, i, 0); }

// contiguous equal blocks over nranks (src/Simulation.jl:353-367): AgentID of global agent index g
VB_HD uint64_t hk_global_id(int type, uint64_t g, uint64_t n, uint32_t nranks) {
    const uint64_t q = n / nranks, r = n % nranks;
    uint64_t p, local;
    if (g < r * (q + 1)) { p = g / (q + 1); local = g % (q + 1); }
    else { p = r + (g - r * (q + 1)) / q; local = (g - r * (q + 1)) % q; }
    return vb::agent_id((uint32_t)type, (uint32_t)p, local + 1);
}

// host generator (used by both libraries)
inline int hk_powerlaw_host(uint64_t n, int agent_type, uint64_t seed_graph, uint64_t seed_opinion, double c, uint32_t dmax,
                            uint64_t* from_out, uint64_t* to_out, double* opinions_out, uint64_t* n_edges_out) {
    uint64_t e = 0;
    for (uint64_t i = 0; i < n; ++i) {
        const uint32_t d = hk_degree(seed_graph, i, c, dmax);
        if (from_out && to_out) {
            const uint64_t to = vb::agent_id((uint32_t)agent_type, 0, i + 1);
            for (uint32_t k = 0; k < d; ++k) {
                from_out[e + i + k] = vb::agent_id((uint32_t)agent_type, 0, hk_source(seed_graph, e + k, n) + 1);
                to_out[e + i + k] = to;
            }
            from_out[e + i + d] = to;   // self loop
            to_out[e + i + d] = to;
        }
        e += d;
        if (opinions_out) opinions_out[i] = hk_opinion(seed_opinion, i);
    }
    *n_edges_out = e + n;
    return 0;
}

}  // namespace vbw
