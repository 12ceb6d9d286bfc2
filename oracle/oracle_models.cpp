// oracle_models.cpp — CPU ORACLE (test infrastructure): instantiates the single-source
// transition functors of vahana.jl_b200/csrc/transitions with the sequential oracle context.
#include <cmath>
#include "vahana_oracle.hpp"
#include "../vahana.jl_b200/csrc/transitions/all.h"

#define VB_TRANSITION(tname, atype, ...) VO_REGISTER_TRANSITION(tname, atype, __VA_ARGS__)
#include "../vahana.jl_b200/csrc/transitions/registry.inc"
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
